#!/bin/bash
# round-2 ncu evidence (one GPU): the launch list of the bench command and `--set full` captures of the dominant kernels.
# Every command is first run to completion WITHOUT ncu (B200_PROFILING.md).  Plain (unchained) step launches: ncu runs
# kernels one at a time, so the chain's overlap cannot be observed under it anyway.
mkdir -p gpurun_out
B="python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-e2e --no-workloads --no-chain --seconds 0.05"
$B > gpurun_out/r2_ncu_plain.json 2> gpurun_out/r2_ncu_plain.err && \
  timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r2_launches.csv $B > gpurun_out/r2_ncu_ll.log 2>&1
echo "launch list rc=$?"
cap() {   # tag workload mode kernel-regex skip
  local C="$B --workload $2 --mode $3 --single-mode"
  $C > /dev/null 2>> gpurun_out/r2_ncu_plain.err && \
    timeout 600 ncu --set full --clock-control none --import-source on -k regex:$4 -s $5 -c 1 -f -o gpurun_out/r2_prof_$1 $C > gpurun_out/r2_ncu_$1.log 2>&1
  echo "$1 rc=$?"
  # summaries are made on the box (the reports are ~21 MB each and only 64 MiB travel back)
  ncu -i gpurun_out/r2_prof_$1.ncu-rep --page raw --csv > gpurun_out/r2_raw_$1.csv 2>/dev/null
  if [ "$6" = "src" ]; then ncu -i gpurun_out/r2_prof_$1.ncu-rep --page source --csv > gpurun_out/r2_src_$1.csv 2>/dev/null; fi
  if [ "$6" != "keep" ]; then rm -f gpurun_out/r2_prof_$1.ncu-rep; fi
}
cap rollout_cfg2 cfg2 rollout oc_rollout_kernel 12 src
cap step_cfg2 cfg2 step oc_step_kernel 40 src
cap rollout_cfg4 cfg4 rollout oc_rollout_kernel 12
cap step_cfg4 cfg4 step oc_step_kernel 40
cap step_cfg3 cfg3 step oc_step_kernel 40
cap step_cfg5 cfg5 step oc_step_kernel 40
cap rollout_cfg3 cfg3 rollout oc_rollout_kernel 12
cap rollout_cfg5 cfg5 rollout oc_rollout_kernel 12
# the compact step kernel (MODE 3), through the one-block host path of tools/e2e_breakdown.py
python tools/e2e_breakdown.py cfg2 > /dev/null 2>&1 && \
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:oc_step_kernel -s 60 -c 1 -f -o gpurun_out/r2_prof_step_compact_cfg2 python tools/e2e_breakdown.py cfg2 > gpurun_out/r2_ncu_compact.log 2>&1
echo "compact rc=$?"
ncu -i gpurun_out/r2_prof_step_compact_cfg2.ncu-rep --page raw --csv > gpurun_out/r2_raw_step_compact_cfg2.csv 2>/dev/null
rm -f gpurun_out/r2_prof_step_compact_cfg2.ncu-rep
du -sh gpurun_out; ls -la gpurun_out | tail -30
