#!/bin/bash
mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests -m gpu -x -q ) 2>&1 | tail -12
timeout 600 python train_ppo.py --iters 60 --log-every 20 > gpurun_out/r2_train_n1_graph.jsonl 2> gpurun_out/r2_train_n1.err
timeout 600 python train_ppo.py --iters 30 --log-every 10 --no-graph > gpurun_out/r2_train_n1_eager.jsonl 2>> gpurun_out/r2_train_n1.err
tail -2 gpurun_out/r2_train_n1_graph.jsonl; tail -1 gpurun_out/r2_train_n1_eager.jsonl; tail -5 gpurun_out/r2_train_n1.err
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r2_bench_k20_c.json 2> gpurun_out/r2_bench_k20_c.err
python - <<'PY'
import json
d = json.loads(open("gpurun_out/r2_bench_k20_c.json").read().strip().splitlines()[-1])
print("rollout %.4g us/step %.3f frac %.3f" % (d["value"], d["ms_per_step"]*1e3, d["roofline"]["frac"]))
s = d["step_api"]; print("step_api %.4g us/step %.3f frac %.3f" % (s["value"], s["ms_per_step"]*1e3, s["roofline"]["frac"]))
for k in ("e2e", "e2e_f32", "e2e_terminal_obs"):
    e = d.get(k); print(k, {a: e.get(a) for a in ("value", "us_per_step", "steps", "finished_envs_per_step", "error")})
for k, v in (d.get("workloads") or {}).items():
    print("   %s rollout %.4g (%.2f us, frac %.3f)  step %.4g (%.2f us, frac %.3f)" % (k, v["value"], v["ms_per_step"]*1e3, v["roofline"]["frac"], v["step_api"]["value"], v["step_api"]["ms_per_step"]*1e3, v["step_api"]["roofline"]["frac"]))
PY
tail -3 gpurun_out/r2_bench_k20_c.err
