#!/bin/bash
# cost of the reset path: cfg4 / cfg5 (random levels) and cfg2 with the resets of a launch packed into a few warps
# (--stagger consecutive) and spread over the batch (default)
cd "$GRAFT_REPO_ROOT"; mkdir -p gpurun_out
python -m pytest tests/test_gpu_scale.py -m gpu -x -q -k "masked_reset or cfg4 or cfg5" 2>&1 | tail -2
for w in cfg4 cfg5 cfg2; do for s in spread consecutive; do
  python bench.py --workload $w --stagger $s --steps 20 --warmup 5 --no-e2e --no-cpu-baseline --no-workloads --seconds 0.3 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('$w $s rollout %.3f us (moved %.3f)  step chained %.3f us  unchained %.3f us  replay %.3f us' % (d['ms_per_step']*1e3, d['roofline']['frac_moved'], d['step_api']['ms_per_step']*1e3, d['step_api_unchained']['ms_per_step']*1e3, d['replay_api']['ms_per_step']*1e3))"
done; done 2>&1 | tee gpurun_out/r2_reset_path.txt
