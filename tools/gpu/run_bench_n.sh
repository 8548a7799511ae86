#!/bin/bash
# the driver's multi-GPU bench command: bash tools/gpu/run_bench_n.sh N
N=$1; cd "$GRAFT_REPO_ROOT"; mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
( time timeout 900 $TR bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/r2_bench_n${N}_final.json 2> gpurun_out/r2_bench_n${N}_final.err ) 2>&1 | grep real
python - <<PY
import json
N=$N
d = json.loads([l for l in open("gpurun_out/r2_bench_n%d_final.json" % N) if l.startswith("{")][-1])
print("N=%d rollout %.4g us/step %.3f frac %.3f" % (N, d["value"], d["ms_per_step"]*1e3, d["roofline"]["frac"]), d["repeats"])
for k in ("step_api", "step_api_unchained", "replay_api"):
    s = d[k]; print("  %s %.4g us/step %.3f" % (k, s["value"], s["ms_per_step"]*1e3))
for k in ("e2e", "e2e_f32", "e2e_terminal_obs"):
    e = d.get(k); print(" ", k, {a: e.get(a) for a in ("value", "us_per_step", "error")})
for k, v in (d.get("workloads") or {}).items():
    print("   %s rollout %.4g  step %.4g" % (k, v.get("value", 0), v.get("step_api", {}).get("value", 0)), v.get("error"))
PY
tail -n 3 gpurun_out/r2_bench_n${N}_final.err
