#!/bin/bash
cd "$GRAFT_REPO_ROOT"; mkdir -p gpurun_out
for g in 4 8 2 1; do for v in default noundo; do
  if [ $v = default ]; then unset OC_B200_LIB; else export OC_B200_LIB=$PWD/gym_comm_b200/variants/liboc_b200_$v.so; fi
  OC_ROW_GROUP=$g python bench.py --workload cfg4 --steps 20 --warmup 5 --no-e2e --no-cpu-baseline --no-workloads --single-mode --seconds 0.3 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('cfg4 OC_ROW_GROUP=$g %-8s rollout %.3f us (moved %.3f)' % ('$v', d['ms_per_step']*1e3, d['roofline']['frac_moved']))"
done; done 2>&1 | tee gpurun_out/r2_cfg4_group_undo.txt
