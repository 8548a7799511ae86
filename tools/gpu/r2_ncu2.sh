#!/bin/bash
# rollout captures again, skipping past the clock-staggering rollouts (no observations) and the 1-step warm-up launches
mkdir -p gpurun_out
B="python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-e2e --no-workloads --no-chain --seconds 0.05"
cap() {
  local C="$B --workload $2 --mode $3 --single-mode"
  $C > /dev/null 2>> gpurun_out/r2_ncu_plain.err && \
    timeout 600 ncu --set full --clock-control none --import-source on -k regex:$4 -s $5 -c 1 -f -o gpurun_out/r2_prof_$1 $C > gpurun_out/r2_ncu_$1.log 2>&1
  echo "$1 rc=$?"
  ncu -i gpurun_out/r2_prof_$1.ncu-rep --page raw --csv > gpurun_out/r2_raw_$1.csv 2>/dev/null
  if [ "$6" = "src" ]; then ncu -i gpurun_out/r2_prof_$1.ncu-rep --page source --csv > gpurun_out/r2_src_$1.csv 2>/dev/null; fi
  rm -f gpurun_out/r2_prof_$1.ncu-rep
}
cap rollout_cfg2 cfg2 rollout oc_rollout_kernel 30 src
cap rollout_cfg4 cfg4 rollout oc_rollout_kernel 30
cap rollout_cfg3 cfg3 rollout oc_rollout_kernel 30
cap rollout_cfg5 cfg5 rollout oc_rollout_kernel 30
for w in cfg4 cfg2 cfg5; do for s in 0 1; do OC_BENCH_NO_STAGGER=$s python bench.py --workload $w --steps 20 --warmup 5 --no-e2e --no-cpu-baseline --no-workloads --mode rollout --single-mode --seconds 0.3 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('$w no_stagger=$s rollout %.3f us/step frac %.3f resets/step %.1f' % (d['ms_per_step']*1e3, d['roofline']['frac'], d['config']['resets_per_step']))"; done; done
python tools/pcie_ceiling.py
