#!/bin/bash
# A/B of the shared-memory row formats on the GPU box: prints rollout / step-API microseconds per step
# for each BASELINE workload under the default format and the OC_ROW_* overrides.
#   gpurun -- 'bash tools/ab_row_formats.sh'
show() { python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$1', 'rollout_us', round(d['ms_per_step']*1e3,2), 'step_us', round(d['step_api']['ms_per_step']*1e3,2))" 2>/dev/null || echo "$1 failed"; }
run() { w=$1; name=$2; shift; shift; env "$@" python bench.py --no-cpu-baseline --no-e2e --steps 2000 --warmup 20 --workload $w 2>/dev/null | show "$w/$name"; }
for w in cfg2 cfg3 cfg4 cfg5; do
    run $w default OC_AB=1
    run $w bytes OC_ROW_FORMAT=b
    run $w f16 OC_ROW_FORMAT=f OC_ROW_ENVS=16
    run $w f8 OC_ROW_FORMAT=f OC_ROW_ENVS=8
done
