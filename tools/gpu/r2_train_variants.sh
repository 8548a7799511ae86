#!/bin/bash
mkdir -p gpurun_out
run() { tag=$1; shift; timeout 300 python train_ppo.py --log-every 10 "$@" > gpurun_out/r2_train_$tag.jsonl 2>gpurun_out/r2_train_$tag.err; python - <<PY
import json
rows=[json.loads(l) for l in open("gpurun_out/r2_train_$tag.jsonl") if l.startswith("{")]
print("$tag:", " | ".join("it%d %.0fs %.1fM/s deliv %.4f len %.1f" % (r["iter"], r["wall_s"], r["agent_steps_per_s"]/1e6, r["delivered_frac"], r["ep_len_mean"]) for r in rows))
PY
}
run default --iters 80
run b262k_e4 --iters 100 --batch-size 262144
run b262k_e2 --iters 140 --batch-size 262144 --epochs 2
run b524k_e2_lr2 --iters 160 --batch-size 524288 --epochs 2 --lr 2e-3
run b262k_e2_lr2 --iters 140 --batch-size 262144 --epochs 2 --lr 2e-3
