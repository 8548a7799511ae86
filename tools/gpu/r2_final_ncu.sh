#!/bin/bash
# ncu captures of the FINAL round-2 build (one GPU): fused kernel cfg2 / cfg4, step kernel cfg2 / cfg4, launch list
cd "$GRAFT_REPO_ROOT"; mkdir -p gpurun_out
B="python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-e2e --no-workloads --no-chain --seconds 0.05"
cap() {
  local C="$B --workload $2 --mode $3 --single-mode"
  $C > /dev/null 2>> gpurun_out/r2_ncu_plain.err && \
    timeout 600 ncu --set full --clock-control none --import-source on -k regex:$4 -s $5 -c 1 -f -o gpurun_out/r2_prof_$1 $C > gpurun_out/r2_ncu_$1.log 2>&1
  echo "$1 rc=$?"
  ncu -i gpurun_out/r2_prof_$1.ncu-rep --page raw --csv > gpurun_out/r2_raw_$1.csv 2>/dev/null
  rm -f gpurun_out/r2_prof_$1.ncu-rep
}
cap final_rollout_cfg2 cfg2 rollout oc_rollout_kernel 30
cap final_rollout_cfg3 cfg3 rollout oc_rollout_kernel 30
if [ "$1" = "all" ]; then
cap final_rollout_cfg4 cfg4 rollout oc_rollout_kernel 30
cap final_step_cfg2 cfg2 step oc_step_kernel 40
cap final_step_cfg4 cfg4 step oc_step_kernel 40
C="$B --workload cfg2"
$C > /dev/null 2>&1 && timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r2_final_launches.csv $C > gpurun_out/r2_ncu_launches.log 2>&1; echo "launch list rc=$?"
fi
