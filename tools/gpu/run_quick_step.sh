#!/bin/bash
# quick check after a step-kernel change: chain / compact / host tests + step numbers of cfg2 and cfg4
cd "$GRAFT_REPO_ROOT"; mkdir -p gpurun_out
python -m pytest tests/test_gpu_chain.py tests/test_gpu_compact.py tests/test_gpu_host_env.py tests/test_gpu_golden.py -m gpu -x -q 2>&1 | tail -2
for w in cfg2 cfg4 cfg3 cfg5; do
  python bench.py --workload $w --steps 20 --warmup 5 --no-e2e --no-cpu-baseline --no-workloads --seconds 0.3 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('$w rollout %.3f us (frac %.3f moved %.3f)  step chained %.3f us (frac %.3f)  unchained %.3f us (%.3f)  replay %.3f us' % (d['ms_per_step']*1e3, d['roofline']['frac'], d['roofline']['frac_moved'], d['step_api']['ms_per_step']*1e3, d['step_api']['roofline']['frac'], d['step_api_unchained']['ms_per_step']*1e3, d['step_api_unchained']['roofline']['frac'], d['replay_api']['ms_per_step']*1e3))"
done 2>&1 | tee gpurun_out/r2_quick_step.txt
