#!/bin/bash
mkdir -p gpurun_out
V=$PWD/gym_comm_b200/variants
timeout 600 python -m pytest tests/test_gpu_compact.py tests/test_gpu_golden.py -x -q 2>&1 | tail -3
( OC_B200_LIB=$V/liboc_b200_probe.so timeout 300 python tools/probe_step.py cfg2 ) > gpurun_out/r2_probe_cfg2_c.txt 2>&1
cat gpurun_out/r2_probe_cfg2_c.txt
timeout 1200 python tools/step_sweep.py cfg2 cfg4 cfg3 cfg5 2>&1 | tee gpurun_out/r2_step_sweep.txt
