#!/bin/bash
# round-2 first GPU session: phase probe of the step kernel, zero-copy probe, new bench protocol, launch-bounds A/B
mkdir -p gpurun_out
V=$PWD/gym_comm_b200/variants
( OC_B200_LIB=$V/liboc_b200_probe.so timeout 300 python tools/probe_step.py cfg2 ) > gpurun_out/r2_probe_cfg2.txt 2>&1
( OC_B200_LIB=$V/liboc_b200_probe.so timeout 300 python tools/probe_step.py cfg5 ) > gpurun_out/r2_probe_cfg5.txt 2>&1
timeout 300 python tools/zero_copy_probe.py cfg2 > gpurun_out/r2_zero_copy.txt 2>&1
timeout 600 python bench.py --steps 20 --warmup 5 --no-e2e --no-cpu-baseline > gpurun_out/r2_bench_k20.json 2> gpurun_out/r2_bench_k20.err
timeout 600 python bench.py --steps 20 --warmup 5 --no-e2e --no-cpu-baseline --no-workloads > gpurun_out/r2_bench_k20_b.json 2>> gpurun_out/r2_bench_k20.err
timeout 600 python bench.py --no-e2e --no-cpu-baseline --no-workloads > gpurun_out/r2_bench_k2000.json 2>> gpurun_out/r2_bench_k20.err
: > gpurun_out/r2_ab_bounds.txt
for w in cfg2 cfg3 cfg5; do
  for lib in default lb4 lb3; do
    if [ $lib = default ]; then unset OC_B200_LIB; else export OC_B200_LIB=$V/liboc_b200_$lib.so; fi
    timeout 300 python bench.py --workload $w --mode step --single-mode --no-cpu-baseline --no-e2e --no-workloads --steps 64 2>/dev/null |
      python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$w $lib', round(d['ms_per_step']*1e3,3), 'us/step', 'frac', round(d['roofline']['frac'],3), d['repeats'])" >> gpurun_out/r2_ab_bounds.txt
  done
done
unset OC_B200_LIB
cat gpurun_out/r2_probe_cfg2.txt gpurun_out/r2_zero_copy.txt gpurun_out/r2_ab_bounds.txt
python - <<'PY'
import json
for f in ("r2_bench_k20", "r2_bench_k20_b", "r2_bench_k2000"):
    try:
        d = json.loads(open("gpurun_out/%s.json" % f).read().strip().splitlines()[-1])
        print(f, "value %.3g ms/step %.5f frac %.3f reps %s" % (d["value"], d["ms_per_step"], d["roofline"]["frac"], d["repeats"]))
        s = d.get("step_api")
        if s: print("   step_api value %.3g us/step %.3f frac %.3f reps %s" % (s["value"], s["ms_per_step"]*1e3, s["roofline"]["frac"], s["repeats"]))
        for k, v in (d.get("workloads") or {}).items():
            if "error" in v: print("  ", k, v); continue
            print("   %s rollout %.3g (%.2f us, frac %.3f)  step %.3g (%.2f us, frac %.3f)" % (k, v["value"], v["ms_per_step"]*1e3, v["roofline"]["frac"], v["step_api"]["value"], v["step_api"]["ms_per_step"]*1e3, v["step_api"]["roofline"]["frac"]))
    except Exception as ex:
        print(f, "unreadable", ex)
PY
tail -5 gpurun_out/r2_bench_k20.err
