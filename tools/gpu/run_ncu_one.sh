#!/bin/bash
# one ncu capture of the fused kernel: bash tools/gpu/run_ncu_one.sh <workload> <tag>
cd "$GRAFT_REPO_ROOT"; mkdir -p gpurun_out
C="python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-e2e --no-workloads --no-chain --seconds 0.05 --workload $1 --mode rollout --single-mode"
$C > /dev/null 2>&1 && timeout 300 ncu --set full --clock-control none --import-source on -k regex:oc_rollout_kernel -s 30 -c 1 -f -o gpurun_out/prof_$2 $C > gpurun_out/ncu_$2.log 2>&1; echo "rc=$?"
ncu -i gpurun_out/prof_$2.ncu-rep --page raw --csv > gpurun_out/r2_raw_$2.csv 2>/dev/null; rm -f gpurun_out/prof_$2.ncu-rep
