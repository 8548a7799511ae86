#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_chain.py -x -q 2>&1 | tail -15
for w in cfg2 cfg3 cfg4 cfg5; do
for c in "" "--no-chain"; do
timeout 300 python bench.py --workload $w --steps 20 --warmup 5 --no-e2e --no-cpu-baseline --no-workloads --mode step --single-mode $c 2>> gpurun_out/r2_run6.err | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('$w $c step_api %.4g us/step %.3f frac %.3f chained %s' % (d['value'], d['ms_per_step']*1e3, d['roofline']['frac'], d['config'].get('mode')), d['repeats']['min_ms'], d['repeats']['max_ms'])"
done; done
timeout 300 python bench.py --steps 2000 --warmup 5 --no-e2e --no-cpu-baseline --no-workloads --mode step --single-mode 2>> gpurun_out/r2_run6.err | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('cfg2 K=2000 step_api %.4g us/step %.3f frac %.3f' % (d['value'], d['ms_per_step']*1e3, d['roofline']['frac']))"
tail -3 gpurun_out/r2_run6.err
