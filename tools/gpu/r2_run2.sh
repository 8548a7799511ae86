#!/bin/bash
# round-2 GPU session 2: new compact kernels + host block path (tests), phase probe, bench protocol
mkdir -p gpurun_out
V=$PWD/gym_comm_b200/variants
timeout 900 python -m pytest tests/test_gpu_compact.py tests/test_gpu_host_env.py tests/test_gpu_cabi_c.py -x -q 2>&1 | tail -15
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
( OC_B200_LIB=$V/liboc_b200_probe.so timeout 300 python tools/probe_step.py cfg2 ) > gpurun_out/r2_probe_cfg2.txt 2>&1
( OC_B200_LIB=$V/liboc_b200_probe.so timeout 300 python tools/probe_step.py cfg5 ) > gpurun_out/r2_probe_cfg5.txt 2>&1
cat gpurun_out/r2_probe_cfg2.txt gpurun_out/r2_probe_cfg5.txt
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r2_bench_k20.json 2> gpurun_out/r2_bench_k20.err
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-workloads > gpurun_out/r2_bench_k20_b.json 2>> gpurun_out/r2_bench_k20.err
timeout 600 python bench.py --no-e2e --no-cpu-baseline --no-workloads > gpurun_out/r2_bench_k2000.json 2>> gpurun_out/r2_bench_k20.err
OC_HOST_ZEROCOPY=0 timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-workloads --single-mode > gpurun_out/r2_bench_k20_nozc.json 2>> gpurun_out/r2_bench_k20.err
python - <<'PY'
import json
for f in ("r2_bench_k20", "r2_bench_k20_b", "r2_bench_k2000", "r2_bench_k20_nozc"):
    try:
        d = json.loads(open("gpurun_out/%s.json" % f).read().strip().splitlines()[-1])
        print(f, "value %.4g us/step %.3f frac %.3f reps %s" % (d["value"], d["ms_per_step"]*1e3, d["roofline"]["frac"], d["repeats"]))
        s = d.get("step_api")
        if s: print("   step_api value %.4g us/step %.3f frac %.3f reps %s" % (s["value"], s["ms_per_step"]*1e3, s["roofline"]["frac"], s["repeats"]))
        for k in ("e2e", "e2e_f32", "e2e_terminal_obs"):
            e = d.get(k)
            if e: print("   %s %s" % (k, {a: e.get(a) for a in ("value", "us_per_step", "steps", "h2d_bytes_per_step", "d2h_bytes_per_step", "finished_envs_per_step", "error")}))
        c = d.get("cpu_baseline")
        if c: print("   cpu", c.get("kind"), c.get("value"), c.get("cores"), "py", (c.get("python_port") or {}).get("value"), "c", (c.get("c_port") or {}).get("value"), c.get("error"))
        print("   clocks", d.get("clocks"))
        for k, v in (d.get("workloads") or {}).items():
            if "error" in v: print("  ", k, v); continue
            print("   %s rollout %.4g (%.2f us, frac %.3f)  step %.4g (%.2f us, frac %.3f) resets/step %.1f" % (k, v["value"], v["ms_per_step"]*1e3, v["roofline"]["frac"], v["step_api"]["value"], v["step_api"]["ms_per_step"]*1e3, v["step_api"]["roofline"]["frac"], v["resets_per_step"]))
    except Exception as ex:
        print(f, "unreadable", repr(ex))
PY
tail -5 gpurun_out/r2_bench_k20.err
