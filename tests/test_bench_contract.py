"""bench.py's CPU arm (`--impl reference`) runs without a GPU and prints the contract's JSON line."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_prints_contract_line():
    out = subprocess.check_output([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "1",
                                   "--steps", "5", "--warmup", "3", "--cpu-seconds", "2"], cwd=ROOT, timeout=300).decode()
    line = json.loads(out.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["unit"] == "agent-steps/s" and line["higher_is_better"] is True
    assert line["metric"] == "env agent-steps/sec incl. obs" and line["value"] > 0
    assert line["e2e"] == {"value": line["value"], "unit": "agent-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    cb = line["cpu_baseline"]
    assert cb["cores"] >= 1 and cb["value"] == line["value"] and "sample" in cb
    from oracle import ref_harness
    if ref_harness.reference_available():               # /root/reference here, oracle/_ref (staged copy) on the GPU box
        assert cb["kind"] == "reference" and cb["python_port"]["kind"] == "port"
        assert cb["python_port"]["value"] > cb["value"]  # the port sheds the reference's template-object overheads
    else:
        assert cb["kind"] == "port"
    assert cb["c_port"]["value"] > cb["value"]          # the C port is reported beside it
    assert line["config"]["workload"].startswith("cfg2: open-divider_tomato, 65536 envs/GPU")
    # the reference arm runs on OUR arm's config: both print the object bench.headline_config builds (static, so the
    # driver's same-config check compares like with like)
    sys.path.insert(0, ROOT)
    import bench
    assert line["config"] == bench.headline_config("cfg2", 65536, "rollout", 64, "spread")
    assert {"mode_desc", "l2", "episode_clocks", "rollout_ring_slots"} <= set(line["config"])


def test_nonzero_rank_of_reference_arm_exits_quietly():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    out = subprocess.check_output([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2"],
                                  cwd=ROOT, env=env, timeout=120).decode()
    assert out.strip() == ""


def test_spread_stagger_is_a_permutation_of_the_clocks():
    """bench.py's steady state: env e starts at (e * M) mod T -- every block of T consecutive envs holds every clock
    exactly once (so exactly E / T envs finish per step, as with e mod T), and neighbours are far apart in time."""
    sys.path.insert(0, ROOT)
    import bench
    for T in (7, 200, 500, 900, 1000):
        M = bench.spread_multiplier(T)
        clocks = sorted((e * M) % T for e in range(T))
        assert clocks == list(range(T)), T
        if T >= 200:                        # 32 consecutive envs (one warp): few of them finish inside any 20-step launch
            block = [(e * M) % T for e in range(32)]
            worst = max(sum(1 for c in block if (c - w) % T < 20) for w in range(T))
            assert worst <= 5, (T, M, worst)   # e mod T puts 20 of them into one launch
