#!/bin/bash
mkdir -p gpurun_out
V=$PWD/gym_comm_b200/variants
timeout 900 python -m pytest tests/test_gpu_compact.py tests/test_gpu_host_env.py tests/test_gpu_golden.py -x -q 2>&1 | tail -5
( OC_B200_LIB=$V/liboc_b200_probe.so timeout 300 python tools/probe_step.py cfg2 ) > gpurun_out/r2_probe_cfg2_b.txt 2>&1
cat gpurun_out/r2_probe_cfg2_b.txt
timeout 600 python tools/e2e_breakdown.py cfg2 > gpurun_out/r2_e2e_breakdown.txt 2>&1
cat gpurun_out/r2_e2e_breakdown.txt
for w in cfg2 cfg3 cfg4 cfg5; do
timeout 600 python bench.py --workload $w --steps 20 --warmup 5 --no-e2e --no-cpu-baseline --no-workloads --mode step --single-mode 2>> gpurun_out/r2_run3.err | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('$w step_api %.4g us/step %.3f frac %.3f' % (d['value'], d['ms_per_step']*1e3, d['roofline']['frac']), d['repeats'])"
done
tail -3 gpurun_out/r2_run3.err
