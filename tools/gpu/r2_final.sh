#!/bin/bash
# round-2 final GPU check (one GPU): parity suite, smoke, the driver's bench command, the reference arm
mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests -m gpu -x -q ) 2>&1 | tail -6
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
python bench.py --steps 20 --warmup 5 > gpurun_out/r2_bench_n1.json 2> gpurun_out/r2_bench_n1.err; echo bench rc=$?
python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-e2e > gpurun_out/r2_bench_n1_again.json 2>> gpurun_out/r2_bench_n1.err
python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/r2_bench_ref.json 2>/dev/null; echo ref rc=$?
python - <<'PY'
import json
for f in ("r2_bench_n1", "r2_bench_n1_again"):
    d = json.loads(open("gpurun_out/%s.json" % f).read().strip().splitlines()[-1])
    print(f, "rollout %.4g us/step %.3f frac %.3f moved %.3f" % (d["value"], d["ms_per_step"]*1e3, d["roofline"]["frac"], d["roofline"]["frac_moved"]), d["repeats"]["n"], d["clocks"])
    s = d["step_api"]; print("  step_api %.4g us/step %.3f frac %.3f" % (s["value"], s["ms_per_step"]*1e3, s["roofline"]["frac"]))
    for k in ("step_api_unchained", "replay_api"):
        s = d.get(k)
        if s: print("  %s %.4g us/step %.3f frac %.3f moved %.3f" % (k, s["value"], s["ms_per_step"]*1e3, s["roofline"]["frac"], s["roofline"]["frac_moved"]))
    for k in ("e2e", "e2e_f32", "e2e_terminal_obs"):
        e = d.get(k)
        if e: print(" ", k, {a: e.get(a) for a in ("value", "us_per_step", "steps", "error")})
    c = d.get("cpu_baseline")
    if c: print("  cpu", c.get("kind"), c.get("value"), c.get("cores"), "py", (c.get("python_port") or {}).get("value"), "c", (c.get("c_port") or {}).get("value"))
    for k, v in (d.get("workloads") or {}).items():
        print("   %s rollout %.4g (%.2f us, frac %.3f)  step %.4g (%.2f us, frac %.3f)" % (k, v["value"], v["ms_per_step"]*1e3, v["roofline"]["frac"], v["step_api"]["value"], v["step_api"]["ms_per_step"]*1e3, v["step_api"]["roofline"]["frac"]))
d = json.loads(open("gpurun_out/r2_bench_ref.json").read().strip().splitlines()[-1])
print("reference arm:", d["value"], d["cpu_baseline"]["kind"], d["cpu_baseline"]["cores"], d["wall_s"])
PY
tail -3 gpurun_out/r2_bench_n1.err
