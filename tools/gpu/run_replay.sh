#!/bin/bash
cd "$GRAFT_REPO_ROOT"; mkdir -p gpurun_out
python -m pytest tests/test_gpu_scale.py -m gpu -x -q -k "replay or fused or rollout" 2>&1 | tail -3
python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-e2e --no-workloads > gpurun_out/r2_bench_replay.json 2> gpurun_out/r2_bench_replay.err; echo rc=$?
python - <<'PY'
import json
d = json.loads([l for l in open("gpurun_out/r2_bench_replay.json") if l.startswith("{")][-1])
for k in ("step_api", "step_api_unchained", "replay_api"):
    s = d[k]; print(k, "%.4g  %.3f us/step  frac %.3f moved %.3f" % (s["value"], s["ms_per_step"]*1e3, s["roofline"]["frac"], s["roofline"]["frac_moved"]))
print("rollout %.4g %.3f us" % (d["value"], d["ms_per_step"]*1e3))
PY
