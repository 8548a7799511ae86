#!/bin/bash
# A/B: per-level tables staged in shared memory (TMA bulk load per CTA, what ships) vs read from global memory through L1
# (-DOC_TABLES_GLOBAL, step kernel only).
#   build container:  bash tools/ab_tables_global.sh build     # variant -> gym_comm_b200/variants/liboc_b200_tg.so (travels with gpurun)
#   GPU box:          gpurun -- 'bash tools/ab_tables_global.sh run'   -> gpurun_out/ab_tables_global.txt
set -e
cd "$(dirname "$0")/.."
V=gym_comm_b200/variants
FLAGS="-gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -shared -Xcompiler -fPIC"
if [ "$1" = "build" ]; then
  mkdir -p $V
  nvcc $FLAGS -DOC_TABLES_GLOBAL gym_comm_b200/csrc/oc_kernels.cu -o $V/liboc_b200_tg.so
  ls -la $V
else
  mkdir -p gpurun_out; : > gpurun_out/ab_tables_global.txt
  for w in cfg2 cfg4 cfg3 cfg5; do
    for lib in default tg; do
      if [ $lib = default ]; then unset OC_B200_LIB; else export OC_B200_LIB=$PWD/$V/liboc_b200_$lib.so; fi
      for chain in "" "--no-chain"; do
        python bench.py --workload $w --mode step --single-mode --no-cpu-baseline --no-e2e --no-workloads --steps 20 --warmup 5 --seconds 0.3 $chain 2>/dev/null |
          python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$w $lib ${chain:-chained}', round(d['ms_per_step']*1e3,3), 'us/step', 'frac', round(d['roofline']['frac'],3))" >> gpurun_out/ab_tables_global.txt
      done
    done
  done
  cat gpurun_out/ab_tables_global.txt
fi
