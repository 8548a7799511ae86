#!/bin/bash
# same-box A/B of library builds, fused kernel only: bash tools/gpu/run_variants_rollout.sh "<workloads>" <variant> ...
cd "$GRAFT_REPO_ROOT"; mkdir -p gpurun_out
WL=$1; shift
python -m pytest tests/test_gpu_scale.py -m gpu -x -q -k "fused or rollout or replay" 2>&1 | tail -1
for rep in 1 2; do for w in $WL; do for v in "$@"; do
  if [ $v = default ]; then unset OC_B200_LIB; else export OC_B200_LIB=$PWD/gym_comm_b200/variants/liboc_b200_$v.so; fi
  python bench.py --workload $w --steps 20 --warmup 5 --no-e2e --no-cpu-baseline --no-workloads --single-mode --seconds 0.3 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('$w %-8s rollout %.3f us (frac %.3f moved %.3f)' % ('$v', d['ms_per_step']*1e3, d['roofline']['frac'], d['roofline']['frac_moved']))"
done; done; done 2>&1 | tee gpurun_out/r2_variants_rollout.txt
