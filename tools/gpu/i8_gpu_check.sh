#!/bin/bash
# Focused GPU check of the host-buffer path (float and compact integer formats, terminal-row gather): the
# plain-C driver (no Python), the numpy parity tests against the C oracle, then the bench line with e2e.
mkdir -p gpurun_out
touch gym_comm_b200/liboc_b200.so oracle/_build/* 2>/dev/null
L=$PWD/gym_comm_b200
nvcc -x c tests/cabi_smoke.c -o /tmp/cabi_smoke -L$L -loc_b200 -Xlinker -rpath,$L 2>/dev/null && timeout 40 /tmp/cabi_smoke 2>&1 | tail -2
timeout 90 python -m pytest tests/test_gpu_host_env.py -x -q 2>&1 | tail -3
timeout 80 python bench.py --steps 3000 --warmup 5 --single-mode --no-cpu-baseline > gpurun_out/r1_bench_i8.json 2> gpurun_out/r1_bench_i8.err; echo bench rc=$?
python -c "
import json; d=json.loads(open('gpurun_out/r1_bench_i8.json').read().strip().splitlines()[-1])
print('e2e', d['e2e']['value']); print('e2e_f32', d['e2e_f32']['value']); print('e2e_term', d['e2e_terminal_obs']); print('value', d['value'])"
