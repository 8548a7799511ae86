#!/bin/bash
cd "$GRAFT_REPO_ROOT"; mkdir -p gpurun_out
python -m pytest tests/test_gpu_scale.py -m gpu -x -q -k "fused or rollout or replay or properties" 2>&1 | tail -2
python -m pytest tests/test_gpu_golden.py tests/test_gpu_fuzz.py -m gpu -x -q 2>&1 | tail -2
for rep in 1 2; do for w in cfg2 cfg3 cfg4; do for u in 1 0; do
  OC_ROW_UNDO=$u python bench.py --workload $w --steps 20 --warmup 5 --no-e2e --no-cpu-baseline --no-workloads --single-mode --seconds 0.3 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('$w OC_ROW_UNDO=$u rollout %.3f us (frac %.3f moved %.3f)' % (d['ms_per_step']*1e3, d['roofline']['frac'], d['roofline']['frac_moved']))"
done; done; done 2>&1 | tee gpurun_out/r2_undo.txt
