#!/bin/bash
# A/B of the step kernel's launch bounds (compile-time knob OC_STEP_MIN_CTAS, profiles/r1_ptxas_sass.txt).
#   build container:  bash tools/ab_step_bounds.sh build     # variants -> gym_comm_b200/variants/*.so (travel with gpurun)
#   GPU box:          gpurun -- 'bash tools/ab_step_bounds.sh run'   -> gpurun_out/ab_step_bounds.txt
set -e
cd "$(dirname "$0")/.."
V=gym_comm_b200/variants
FLAGS="-gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -shared -Xcompiler -fPIC"
if [ "$1" = "build" ]; then
  mkdir -p $V
  for n in 4 3; do
    nvcc $FLAGS -DOC_STEP_MIN_CTAS=$n gym_comm_b200/csrc/oc_kernels.cu -o $V/liboc_b200_lb$n.so &
  done
  wait; ls -la $V
else
  mkdir -p gpurun_out; : > gpurun_out/ab_step_bounds.txt
  for w in cfg2 cfg3 cfg5; do
    for lib in default lb4 lb3; do
      if [ $lib = default ]; then unset OC_B200_LIB; else export OC_B200_LIB=$PWD/$V/liboc_b200_$lib.so; fi
      python bench.py --workload $w --mode step --single-mode --no-cpu-baseline --no-e2e --steps 20000 2>/dev/null |
        python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$w $lib', round(d['ms_per_step']*1e3,3), 'us/step', 'frac', round(d['roofline']['frac'],3))" >> gpurun_out/ab_step_bounds.txt
    done
  done
  cat gpurun_out/ab_step_bounds.txt
fi
