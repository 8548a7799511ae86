#!/bin/bash
# multi-GPU session: bash tools/gpu/r2_run_n.sh N
N=$1
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
timeout 600 $TR bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/r2_bench_n$N.json 2> gpurun_out/r2_bench_n$N.err
timeout 300 $TR tools/pcie_ceiling.py > gpurun_out/r2_pcie_ceiling_n$N.json 2> gpurun_out/r2_pcie_n$N.err
timeout 600 $TR train_ppo.py --iters 60 --log-every 20 > gpurun_out/r2_train_n${N}_tomato.jsonl 2> gpurun_out/r2_train_n${N}_tomato.err
timeout 600 $TR train_ppo.py --json-path configs/env_args100on.json --envs 131072 --iters 30 --log-every 10 > gpurun_out/r2_train_n${N}_cfg5.jsonl 2> gpurun_out/r2_train_n${N}_cfg5.err
OC_PPO_GRAPH_DP=0 timeout 600 $TR train_ppo.py --iters 30 --log-every 10 > gpurun_out/r2_train_n${N}_tomato_eagerdp.jsonl 2>> gpurun_out/r2_train_n${N}_tomato.err
python - <<PY
import json
N=$N
try:
    d = json.loads(open("gpurun_out/r2_bench_n%d.json" % N).read().strip().splitlines()[-1])
    print("N=%d rollout %.4g us/step %.3f frac %.3f" % (N, d["value"], d["ms_per_step"]*1e3, d["roofline"]["frac"]), d["repeats"])
    s = d["step_api"]; print("  step_api %.4g us/step %.3f" % (s["value"], s["ms_per_step"]*1e3))
    for k in ("e2e", "e2e_f32", "e2e_terminal_obs"):
        e = d.get(k); print(" ", k, {a: e.get(a) for a in ("value", "us_per_step", "error")})
    for k, v in (d.get("workloads") or {}).items():
        print("   %s rollout %.4g  step %.4g" % (k, v.get("value", 0), v.get("step_api", {}).get("value", 0)), v.get("error"))
except Exception as ex: print("bench unreadable", ex)
for f in ("pcie_ceiling_n%d.json" % N,):
    try: print(open("gpurun_out/r2_" + f).read().strip()[-600:])
    except Exception as ex: print(f, ex)
for f in ("train_n%d_tomato" % N, "train_n%d_cfg5" % N, "train_n%d_tomato_eagerdp" % N):
    try:
        rows=[json.loads(l) for l in open("gpurun_out/r2_%s.jsonl" % f) if l.startswith("{")]
        print(f, " | ".join("it%d %.0fs %.1fM/s deliv %.4f" % (r["iter"], r["wall_s"], r["agent_steps_per_s"]/1e6, r["delivered_frac"]) for r in rows))
    except Exception as ex: print(f, ex)
PY
tail -q -n 3 gpurun_out/r2_bench_n$N.err gpurun_out/r2_train_n${N}_tomato.err gpurun_out/r2_train_n${N}_cfg5.err
