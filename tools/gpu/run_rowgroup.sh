#!/bin/bash
# grouped float rows (row sizes that are a multiple of 8 words): parity + OC_ROW_GROUP sweep on cfg4
cd "$GRAFT_REPO_ROOT"; mkdir -p gpurun_out
python -m pytest tests/test_gpu_golden.py tests/test_gpu_chain.py tests/test_gpu_fuzz.py -m gpu -x -q 2>&1 | tail -2
python -m pytest tests/test_gpu_scale.py -m gpu -x -q -k "cfg4 or properties or ragged" 2>&1 | tail -2
for g in 1 2 4 8 16; do
  OC_ROW_GROUP=$g python bench.py --workload cfg4 --steps 20 --warmup 5 --no-e2e --no-cpu-baseline --no-workloads --seconds 0.3 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('cfg4 OC_ROW_GROUP=$g rollout %.3f us (moved %.3f)  step chained %.3f us (frac %.3f)  unchained %.3f us  replay %.3f us' % (d['ms_per_step']*1e3, d['roofline']['frac_moved'], d['step_api']['ms_per_step']*1e3, d['step_api']['roofline']['frac'], d['step_api_unchained']['ms_per_step']*1e3, d['replay_api']['ms_per_step']*1e3))"
done 2>&1 | tee gpurun_out/r2_rowgroup.txt
