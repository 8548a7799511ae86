#!/bin/bash
for w in cfg2 cfg4 cfg3 cfg5; do for t in 32 64 96 128 160 192 224 256; do OC_BLOCK_THREADS=$t timeout 120 python bench.py --workload $w --steps 20 --warmup 5 --no-e2e --no-cpu-baseline --no-workloads --mode step --single-mode --seconds 0.2 2>/dev/null | python -c "
import json,sys
try:
    d=json.loads(sys.stdin.read().strip().splitlines()[-1])
    print('$w BLOCK_THREADS=$t chained %.3f us/step frac %.3f' % (d['ms_per_step']*1e3, d['roofline']['frac']))
except Exception as ex: print('$w BLOCK_THREADS=$t failed')"; done; done
