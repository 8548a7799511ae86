#!/bin/bash
# host-buffer path A/B (pieces, direct-write share) + correctness of the host-block path + a bench line
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
python -m pytest tests/test_gpu_host_env.py tests/test_gpu_compact.py tests/test_gpu_cabi_c.py -x -q -m gpu 2>&1 | tail -3
for D in 0 5 10 15 20 30 50 100; do
  echo "== OC_HOST_DIRECT_PCT=$D"
  OC_HOST_DIRECT_PCT=$D python tools/e2e_breakdown.py 2>&1 | grep -E "D oc_step|E Overcooked|G NO_SYNC"
done > gpurun_out/r2_e2e_direct.txt 2>&1
cat gpurun_out/r2_e2e_direct.txt
python bench.py --steps 20 --warmup 5 > gpurun_out/r2_bench_n1_c.json 2> gpurun_out/r2_bench_n1_c.err; echo bench rc=$?
python - <<'PY'
import json
d=json.loads([l for l in open('gpurun_out/r2_bench_n1_c.json') if l.startswith('{')][-1])
print('rollout %.4g step %.4g unchained %.4g e2e %.4g (%.1f us) term %.4g f32 %.4g' % (d['value'], d['step_api']['value'], d['step_api_unchained']['value'], d['e2e']['value'], d['e2e']['us_per_step'], d['e2e_terminal_obs']['value'], d['e2e_f32']['value']))
PY
