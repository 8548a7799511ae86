#!/bin/bash
# same-box A/B of library builds: bash tools/gpu/run_variants.sh "<workloads>" <variant> <variant> ...   (default = the in-tree build)
cd "$GRAFT_REPO_ROOT"; mkdir -p gpurun_out
WL=$1; shift
for rep in 1 2; do for w in $WL; do for v in "$@"; do
  if [ $v = default ]; then unset OC_B200_LIB; else export OC_B200_LIB=$PWD/gym_comm_b200/variants/liboc_b200_$v.so; fi
  python bench.py --workload $w --steps 20 --warmup 5 --no-e2e --no-cpu-baseline --no-workloads --seconds 0.3 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('$w %-8s rollout %.3f us  step chained %.3f us  unchained %.3f us  replay %.3f us' % ('$v', d['ms_per_step']*1e3, d['step_api']['ms_per_step']*1e3, d['step_api_unchained']['ms_per_step']*1e3, d['replay_api']['ms_per_step']*1e3))"
done; done; done 2>&1 | tee gpurun_out/r2_variants.txt
