#!/bin/bash
cd "$GRAFT_REPO_ROOT"; mkdir -p gpurun_out
python -m pytest tests/test_gpu_golden.py tests/test_gpu_chain.py -m gpu -x -q 2>&1 | tail -2
for w in cfg2 cfg3; do for r in 1 0; do
  OC_OBS_ROT=$r python bench.py --workload $w --steps 20 --warmup 5 --no-e2e --no-cpu-baseline --no-workloads --seconds 0.3 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('$w OC_OBS_ROT=$r rollout %.3f us (moved %.3f)  step chained %.3f us  unchained %.3f us  replay %.3f us' % (d['ms_per_step']*1e3, d['roofline']['frac_moved'], d['step_api']['ms_per_step']*1e3, d['step_api_unchained']['ms_per_step']*1e3, d['replay_api']['ms_per_step']*1e3))"
done; done 2>&1 | tee gpurun_out/r2_obsrot.txt
