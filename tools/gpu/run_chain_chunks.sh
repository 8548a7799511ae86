#!/bin/bash
# chained steps: chunks per CTA (OC_CHAIN_CHUNKS) x CTA size sweep
cd "$GRAFT_REPO_ROOT"; mkdir -p gpurun_out
run() {  # workload, env assignments...
  local w=$1; shift
  env "$@" timeout 120 python bench.py --workload $w --steps 20 --warmup 5 --no-e2e --no-cpu-baseline --no-workloads --mode step --single-mode --seconds 0.2 2>/dev/null | python -c "
import json,sys
try:
    d=json.loads(sys.stdin.read().strip().splitlines()[-1])
    print('$w $* chained %.3f us/step frac %.3f' % (d['ms_per_step']*1e3, d['roofline']['frac']))
except Exception as ex: print('$w $* failed')"
}
for w in cfg2 cfg4; do
  for m in 1 2 3 4 6; do for t in 64 128 224; do run $w OC_CHAIN_CHUNKS=$m OC_BLOCK_THREADS_CHAIN=$t; done; done
done 2>&1 | tee gpurun_out/r2_chain_chunks.txt
for w in cfg3 cfg5; do run $w OC_CHAIN_CHUNKS=1; for m in 1 2 4; do run $w OC_CHAIN_CHUNKS=$m OC_BLOCK_THREADS_CHAIN=128; done; done 2>&1 | tee -a gpurun_out/r2_chain_chunks.txt
