#!/bin/bash
# episode-clock pattern A/B (resets spread over the batch vs on consecutive envs) + an ncu capture of the cfg4 fused launch
cd "$GRAFT_REPO_ROOT"; mkdir -p gpurun_out
for S in spread consecutive; do
  python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-e2e --stagger $S > gpurun_out/r2_bench_stagger_$S.json 2> gpurun_out/r2_bench_stagger_$S.err; echo "$S rc=$?"
done
python - <<'PY'
import json
for S in ("spread", "consecutive"):
    d = json.loads([l for l in open("gpurun_out/r2_bench_stagger_%s.json" % S) if l.startswith("{")][-1])
    print(S, "cfg2 rollout %.3f us  step %.3f us  unchained %.3f us" % (d["ms_per_step"]*1e3, d["step_api"]["ms_per_step"]*1e3, d["step_api_unchained"]["ms_per_step"]*1e3))
    for k, v in d["workloads"].items():
        print("   %s rollout %.3f us (frac %.3f, moved %.3f)  step %.3f us (frac %.3f)" % (k, v["ms_per_step"]*1e3, v["roofline"]["frac"], v["roofline"]["frac_moved"], v["step_api"]["ms_per_step"]*1e3, v["step_api"]["roofline"]["frac"]))
PY
B="python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-e2e --no-workloads --no-chain --seconds 0.05"
cap() {
  local C="$B --workload $2 --mode $3 --single-mode"
  $C > /dev/null 2>> gpurun_out/r2_ncu_plain.err && \
    timeout 600 ncu --set full --clock-control none --import-source on -k regex:$4 -s $5 -c 1 -f -o gpurun_out/r2_prof_$1 $C > gpurun_out/r2_ncu_$1.log 2>&1
  echo "$1 rc=$?"
  ncu -i gpurun_out/r2_prof_$1.ncu-rep --page raw --csv > gpurun_out/r2_raw_$1.csv 2>/dev/null
  rm -f gpurun_out/r2_prof_$1.ncu-rep
}
cap rollout_cfg4_spread cfg4 rollout oc_rollout_kernel 30
cap rollout_cfg5_spread cfg5 rollout oc_rollout_kernel 30
cap rollout_cfg2_spread cfg2 rollout oc_rollout_kernel 30
