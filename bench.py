#!/usr/bin/env python
"""Benchmark of the Overcooked env step + observation path (BASELINE.json metric).

    python bench.py --gpus N --steps K --warmup W            # ours (CUDA, sm_100a)
    python bench.py --impl reference --steps K --warmup W    # CPU arm: the unmodified reference on all host cores

A "step" is one pass of the hot path over one batch: every env of the workload advances one
timestep and every agent's observation is featurised.  Headline workload (BASELINE.json configs[1]):
open-divider_tomato, 2 agents, comm on (C=10), T=500, 65,536 lock-step envs per GPU, uniform random
actions, auto-reset, episode clocks staggered ((e * M) mod T) so that about E/T envs finish in EVERY step.  N GPUs =
N independent shards (weak scaling, no collective on the step path; torch.distributed only for the
barrier and the max-over-ranks of the device time).

How a number is taken (the same for every workload and mode):
  * the K-step region is enqueued behind a device-side gate (a spin kernel), so the host's enqueue
    latency is not inside the CUDA-event bracket: the bracket holds exactly K steps of GPU work;
  * the region is repeated (barrier + synchronize on both sides of every repeat), each repeat's
    duration is the max over ranks, and `value` uses the MEDIAN repeat (min / max beside it);
  * observations go round-robin to a ring of rollout-buffer slots larger than L2, actions come from
    a pool larger than L2; only the packed state (64 B/env) is L2-resident, by design.

The JSON line carries: `value` (fused rollout `oc_rollout`, Philox actions drawn on the device),
`step_api` (one `oc_step` launch per step, actions read from HBM, CUDA graphs), `workloads` (the
other BASELINE configs: cfg3, cfg4, cfg5 -- both modes each), `e2e` (the public host-buffer VecEnv
API, H2D + D2H inside the timed region), `roofline`, `cpu_baseline`, `clocks`.

Only the ``cpu_baseline`` leg and ``--impl reference`` touch ``oracle/`` (as the thing timed
beside us, never as our result).
"""
from __future__ import annotations

import argparse
import json
import math
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

WORKLOADS = {
    # BASELINE.json configs[1]
    "cfg2": dict(level="open-divider_tomato", num_agents=2, max_num_timesteps=500, communication_on=True,
                 num_communication=10, fow_radius=2, envs=65536),
    # configs[2]
    "cfg3": dict(level="partial-divider_salad", num_agents=3, max_num_timesteps=500, communication_on=True,
                 num_communication=10, fow_radius=2, envs=262144),
    # configs[3]: spread/env_args20on_allergic.json
    "cfg4": dict(level="random-open-divider_salad_small_cramped", num_agents=2, max_num_timesteps=900,
                 communication_on=True, num_communication=8, fow_radius=10, envs=65536,
                 ego_config=dict(CAN_MOVE=False, ALLERGIC=True, BLIND=False),
                 partner_config=dict(CAN_MOVE=True, ALLERGIC=False, BLIND=True)),
    # configs[4]: spread/env_args100on.json scale (131,072 envs per GPU = 1,048,576 on 8 GPUs)
    "cfg5": dict(level="random-salad-superwide", num_agents=2, max_num_timesteps=900, communication_on=True,
                 num_communication=100, fow_radius=2, envs=131072),
}
METRIC = "env agent-steps/sec incl. obs"


def workload_namespace(w):
    from gym_comm_b200.arglist import namespace_from_dict
    d = {k: v for k, v in w.items() if k != "envs"}
    return namespace_from_dict(d)


def workload_config(wname, E, mode):
    """The `config` object both arms print (no env needed: F = 23 + S + 2C)."""
    from gym_comm_b200 import levels_data
    w = WORKLOADS[wname]
    text = levels_data.LEVELS[w["level"]]
    S = len(levels_data.SUBTASKS[tuple(text.split("\n\n")[1].split("\n"))])
    A, C = w["num_agents"], w["num_communication"]
    F = 23 + S + 2 * C
    return {"workload": "%s: %s, %d envs/GPU, uniform random (nav, comm) actions, auto-reset, obs f32 [E,%d,%d]" %
                        (wname, w["level"], E, A, F),
            "mode": mode, "envs_per_gpu": E, "num_agents": A, "obs_width": F,
            "max_num_timesteps": w["max_num_timesteps"], "num_communication": C,
            "fow_radius": w["fow_radius"]}, A, F


def ring_and_pool(E, A, F, ring_slots):
    """(R, P): slots of the observation ring (at least 4x L2, at most ~2 GB) and batches of the action pool (> L2)."""
    slot = E * A * F * 4
    R = max(4, min(ring_slots, int(2.0e9 // slot)))
    P = max(R, min(1024, int(math.ceil(256e6 / (E * A * 8)))))
    return R, P


def headline_config(wname, E, mode, ring_slots, stagger):
    """The complete `config` object of the JSON line -- static, so that `--impl reference` prints the SAME object
    (the reference arm runs on this arm's config)."""
    cfg, A, F = workload_config(wname, E, mode)
    R, P = ring_and_pool(E, A, F, ring_slots)
    T = int(WORKLOADS[wname]["max_num_timesteps"])
    cfg.update(
        mode_desc=DESC[mode], rollout_ring_slots=R,
        episode_clocks=("env e starts at (e * %d) mod T: about E/T envs finish in every step, spread over the batch"
                        % spread_multiplier(T)) if stagger == "spread"
        else "env e starts at e mod T: about E/T CONSECUTIVE envs finish in every step",
        l2="inputs/outputs larger than L2: the obs ring (%d x %.1f MB) is rewritten round-robin and the action pool is "
           "%.2f GB; only the %.1f MB packed state stays L2-resident (by design)"
           % (R, E * A * F * 4 / 1e6, P * E * A * 8 / 1e9, E * 64 / 1e6))
    return cfg


def bytes_per_env_step(A, F):
    """obs out A*4*F + actions in A*2*4 + reward out A*4 + done 4 (u8 padded) +
    packed state read+write 2*64."""
    return A * 4 * F + A * 8 + A * 4 + 4 + 128


# ------------------------------------------------------------------------------- clocks
class ClockSampler:
    """nvidia-smi streaming at 50 ms (the recipe's clocks line) for the duration of a timed region."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu = gpu_index
        self.proc = None
        self.lines = []
        self._thr = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append((time.perf_counter(), line.decode().strip()))

    def __enter__(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "50"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL)
            self._thr = threading.Thread(target=self._pump, daemon=True)
            self._thr.start()
            time.sleep(0.25)          # let the first samples arrive before the region starts
        except Exception:
            self.proc = None
        self.t0 = time.perf_counter()
        return self

    def __exit__(self, *a):
        self.t1 = time.perf_counter()
        if self.proc is not None:
            time.sleep(0.06)
            self.proc.terminate()
            try:
                self.proc.wait(timeout=5)
            except Exception:
                self.proc.kill()

    def summary(self):
        rows = [l.split(",") for t, l in self.lines if self.t0 <= t <= self.t1 + 0.06]
        rows = [[x.strip() for x in r] for r in rows if len(r) >= 7]
        if not rows:       # region shorter than the sampling period: take whatever was seen
            rows = [[x.strip() for x in l.split(",")] for _, l in self.lines][-3:]
            rows = [r for r in rows if len(r) >= 7]

        def num(x):
            try:
                return float(x)
            except Exception:
                return None
        sm = [num(r[0]) for r in rows if num(r[0]) is not None]
        mx = [num(r[1]) for r in rows if num(r[1]) is not None]
        pw = [num(r[2]) for r in rows if num(r[2]) is not None]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({names[i] for r in rows for i in range(4) if r[3 + i].lower().startswith("active")})
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "power_w_max": max(pw) if pw else None, "reasons": reasons, "samples": len(rows)}


# ------------------------------------------------------------------------------- CPU arm
def cpu_reference_arm(wname, seconds):
    from oracle import cpu_baseline
    return cpu_baseline.run_all_cores(WORKLOADS[wname], workload_namespace(WORKLOADS[wname]), seconds)


def load_json(*path):
    try:
        return json.load(open(os.path.join(ROOT, *path)))
    except Exception:
        return {}


def spread_multiplier(T: int) -> int:
    """Smallest M >= 0.618 T that is coprime to T: e -> (e * M) mod T is a bijection on every block of T envs."""
    M = max(1, int(round(0.618 * T)))
    while math.gcd(M, T) != 1:
        M += 1
    return M


# ------------------------------------------------------------------------------- one workload on this rank's GPU
class Bench:
    """One workload on this rank's GPU: env handle, rollout-buffer ring, action pool, the timing protocol."""

    def __init__(self, wname, E, rank, world, dev, ring_slots, peak, peak_src, no_graph=False, stagger="spread"):
        self.stagger = stagger
        import torch
        from gym_comm_b200.vec_env import OvercookedVecEnv
        self.torch, self.wname, self.w = torch, wname, WORKLOADS[wname]
        self.rank, self.world, self.dev = rank, world, dev
        self.peak, self.peak_src, self.no_graph = peak, peak_src, no_graph
        self.chain = None            # True: the step API's K steps form one chain (head + chained launches) per graph
        self.ns = workload_namespace(self.w)
        self.E, self.A = E, self.ns.num_agents
        self.env = OvercookedVecEnv(self.ns, num_envs=E, device=dev, seed=1234 + rank, auto_reset=True)
        self.F = F = self.env.obs_width
        A = self.A
        self.bpes = bytes_per_env_step(A, F)
        # ring of rollout-buffer slots the observations go to, round-robin; pool of pre-drawn action batches resident
        # in HBM, cycled through by the step API: both larger than L2
        self.R, self.P = R, P = ring_and_pool(E, A, F, ring_slots)
        gen = torch.Generator(device=dev)
        gen.manual_seed(99 + rank)
        self.actions = torch.stack([torch.randint(0, 4, (P, E, A), generator=gen, device=dev, dtype=torch.int32),
                                    torch.randint(0, self.ns.num_communication, (P, E, A), generator=gen, device=dev,
                                                  dtype=torch.int32)], -1).contiguous()
        self.obs_ring = torch.empty((R, E, A, F), dtype=torch.float32, device=dev)
        self.rew_ring = torch.empty((R, E, A), dtype=torch.float32, device=dev)
        self.done_ring = torch.empty((R, E), dtype=torch.uint8, device=dev)
        self.env.reset()
        self.resets_per_step = 0.0
        if os.environ.get("OC_BENCH_NO_STAGGER") != "1":      # diagnostic: lock-step clocks, no env finishes inside the region
            self.stagger_clocks()
        self.traffic = load_json("profiles", "traffic.json")

    def close(self):
        self.env.close()
        del self.actions, self.obs_ring, self.rew_ring, self.done_ring
        self.torch.cuda.empty_cache()

    def stagger_clocks(self):
        """Steady state of a long run: env e's episode clock starts at (e * M) mod T (through the C ABI's state
        export / import), then one full episode length of fused rollout, so every env has crossed a reset
        and about E/T envs finish -- reset in place, random placements redrawn -- in every later step.
        M (coprime to T, near 0.618 T) permutes the clocks inside every block of T consecutive envs: the E/T
        resets of a step are spread over the batch, as the asynchronous episode ends of a real run are.  With
        M = 1 (`--stagger consecutive`) the same number of resets falls on CONSECUTIVE envs, i.e. a few warps pay
        for all resets of a whole launch and the one-wave kernels wait for them (random levels: +25 %)."""
        torch, env = self.torch, self.env
        T = int(self.w["max_num_timesteps"])
        st = env.get_state()
        M = spread_multiplier(T) if self.stagger == "spread" else 1
        t = ((torch.arange(self.E, device=self.dev, dtype=torch.int64) * M) % T).to(torch.int32)
        st[:, 0] = (st[:, 0] & ~0xFFFF) | t
        env.set_state(st)
        left = T
        while left > 0:
            n = min(left, self.R)
            env.rollout(n, done_out=self.done_ring[:n])
            left -= n
        torch.cuda.synchronize(self.dev)
        self.resets_per_step = float(self.done_ring[:min(T, self.R)].float().sum().item()) / min(T, self.R)

    def barrier(self):
        torch = self.torch
        torch.cuda.synchronize(self.dev)
        if self.world > 1:
            torch.distributed.barrier()
        torch.cuda.synchronize(self.dev)

    def do_step(self, i, chain=None):
        R, P = self.R, self.P
        self.env.step(self.actions[i % P], obs_out=self.obs_ring[i % R], rew_out=self.rew_ring[i % R],
                      done_out=self.done_ring[i % R], chain=chain)

    def do_rollout(self, n):
        self.env.rollout(n, obs_out=self.obs_ring[:n], rew_out=self.rew_ring[:n], done_out=self.done_ring[:n])

    def do_replay(self, n, first):
        """oc_replay: n steps in ONE launch on a caller-given action sequence (a contiguous slice of the pool)."""
        first %= max(1, self.P - n + 1)
        self.env.replay(self.actions[first:first + n], obs_out=self.obs_ring[:n], rew_out=self.rew_ring[:n],
                        done_out=self.done_ring[:n])

    def measure(self, mode, K, W, seconds=0.7, min_reps=30, max_reps=1500, diagnostics=True, clocks=True):
        """Times EXACTLY K steps of `mode` per repeat on the device: gate kernel, event, K steps, event; barrier +
        synchronize on both sides of every repeat; per-repeat max over ranks; median over repeats."""
        torch, env, dev, E, A, F, R, P = self.torch, self.env, self.dev, self.E, self.A, self.F, self.R, self.P
        if mode == "step":
            for i in range(W):
                self.do_step(i)
        elif mode == "replay":
            for i in range(W):
                self.do_replay(1, i)
        else:
            for _ in range(W):
                self.do_rollout(1)
        self.barrier()
        launches0 = env.launch_count()
        # CUDA graphs for the step API: one launch is a few us of GPU work, shorter than a Python -> ctypes call, so
        # the launches are captured in graphs of min(K, R) steps (each step reads its own action batch and writes its
        # own ring slot) and replayed.
        graphs, tail_graph = [], None
        G = min(K, R)
        if mode == "step" and not self.no_graph:
            for g0 in range(0, min(P, 4 * G), G):
                g = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g):
                    for j in range(G):
                        self.do_step(g0 + j, (("head" if j == 0 else "next") if self.chain else None))
                graphs.append(g)
            if K % G:
                tail_graph = torch.cuda.CUDAGraph()
                with torch.cuda.graph(tail_graph):
                    for j in range(K % G):
                        self.do_step(j, (("head" if j == 0 else "next") if self.chain else None))
            self.barrier()
        capture_launches = env.launch_count() - launches0

        def enqueue(rep):
            """exactly K steps; returns the number of kernel launches"""
            if mode == "step" and graphs:
                for j in range(K // G):
                    graphs[(rep * (K // G) + j) % len(graphs)].replay()
                if tail_graph is not None:
                    tail_graph.replay()
                return K
            if mode == "step":
                for i in range(K):
                    self.do_step(rep * K + i, (("head" if i == 0 else "next") if self.chain else None))
                return K
            i, n_l = 0, 0
            while i < K:
                n = min(R, K - i)
                if mode == "replay":
                    self.do_replay(n, rep * K + i)
                else:
                    self.do_rollout(n)
                i += n
                n_l += 1
            return n_l

        host_calls = (K // G + 1) if (mode == "step" and graphs) else (K if mode == "step" else (K + R - 1) // R)
        gate_cycles = int(min(max(3e5, host_calls * 4e4), 4e8))       # >= 0.15 ms; ~20 us of head start per host call
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)

        def one_rep(rep):
            self.barrier()
            torch.cuda._sleep(gate_cycles)            # device-side gate: the host runs ahead while the GPU spins
            e0.record()
            n_l = enqueue(rep)
            e1.record()
            torch.cuda.synchronize(dev)
            return e0.elapsed_time(e1), n_l

        t_w = time.perf_counter()
        for rep in range(3):                          # untimed: graph upload, allocator, clocks
            _, nlaunch = one_rep(rep)
        rep_wall = (time.perf_counter() - t_w) / 3
        nrep = int(min(max(min_reps, seconds / max(rep_wall, 1e-6)), max_reps))
        if self.world > 1:                            # every rank must run the same number of repeats (barriers inside)
            nr = torch.tensor([nrep], dtype=torch.int64, device=dev)
            torch.distributed.all_reduce(nr, op=torch.distributed.ReduceOp.MIN)
            nrep = int(nr.item())
        ms = []
        sampler = ClockSampler(dev.index) if clocks else None
        if sampler is not None:
            sampler.__enter__()
        for rep in range(nrep):
            ms.append(one_rep(3 + rep)[0])
        if sampler is not None:
            sampler.__exit__()
        tm = torch.tensor(ms, dtype=torch.float64, device=dev)
        if self.world > 1:
            torch.distributed.all_reduce(tm, op=torch.distributed.ReduceOp.MAX)
        ms_max = sorted(tm.tolist())
        total_ms = statistics.median(ms_max)
        value = float(E) * A * K * self.world / (total_ms / 1e3)
        local_ms = statistics.median(ms)

        # roofline: algorithmic bytes per launch / average launch duration over the timed region (the bracket holds
        # nothing but back-to-back launches of this one kernel, so total / launches is its average duration incl.
        # the inter-launch gap -- an upper bound on the kernel time)
        bpes = self.bpes
        avg_launch_ms = local_ms / nlaunch
        spl = K / nlaunch
        achieved = float(bpes) * E * spl / (avg_launch_ms / 1e3) / 1e9
        moved = bpes if mode == "step" else A * 4 * F + A * 4 + 1 + (A * 8 if mode == "replay" else 0)
        tr = self.traffic.get(self.wname + ":" + mode)
        roof = {"bound": "hbm", "achieved": achieved, "peak": self.peak, "unit": "GB/s", "frac": achieved / self.peak,
                "traffic": (tr["dram_bytes_per_env_step"] * E * spl) if tr else None,
                "traffic_source": (tr["source"] if tr else None), "peak_source": self.peak_src,
                "kernel": "oc_step_kernel" if mode == "step" else "oc_rollout_kernel",      # oc_replay launches the rollout kernel
                "bytes_per_env_step": bpes, "envs_per_launch": E, "steps_per_launch": spl,
                # what this kernel really moves per env-step: the fused rollout keeps state and actions on chip, so
                # only obs + reward + done cross HBM (that is why its algorithmic frac can exceed 1)
                "bytes_moved_per_env_step": moved,
                "achieved_moved": float(moved) * E * spl / (avg_launch_ms / 1e3) / 1e9,
                "frac_moved": float(moved) * E * spl / (avg_launch_ms / 1e3) / 1e9 / self.peak,
                "launches_in_timed_region": nlaunch, "avg_launch_us": avg_launch_ms * 1e3,
                "us_per_step": local_ms * 1e3 / K,
                "timing": "CUDA events on the launching stream around K steps enqueued behind a device-side gate; "
                          "median of %d repeats" % nrep}
        if roof["frac_moved"] > 1.0:
            roof["note"] = ("a write-only stream: the peak is the measured COPY bandwidth (read+write); ncu reports the "
                            "same launch at 88.7 % of the device's own DRAM peak (profiles/r1_ncu_cfg5_rollout_summary.txt)")
        if diagnostics:
            # L2-flushed variant of the same launch (diagnostic): state, actions and obs lines all cold
            flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)
            fl_ms = []
            nfl = min(K, R)
            for i in range(6):
                flush.fill_(i & 0xFF)
                torch.cuda._sleep(int(2e5))
                e0.record()
                if mode == "step":
                    self.do_step(i)
                elif mode == "replay":
                    self.do_replay(nfl, i)
                else:
                    self.do_rollout(nfl)
                e1.record()
                torch.cuda.synchronize(dev)
                fl_ms.append(e0.elapsed_time(e1))
            del flush
            roof["l2_flushed_launch_us"] = statistics.median(fl_ms) * 1e3
            roof["l2_flushed_steps_per_launch"] = 1 if mode == "step" else nfl
        return {"value": value, "ms_per_step": total_ms / K, "total_ms": total_ms, "roofline": roof,
                "repeats": {"n": nrep, "min_ms": ms_max[0], "median_ms": total_ms, "max_ms": ms_max[-1],
                            "p10_ms": ms_max[len(ms_max) // 10], "p90_ms": ms_max[(len(ms_max) * 9) // 10]},
                "gpu_launches": int(nlaunch), "captured_launches": int(capture_launches),
                "clocks": sampler.summary() if sampler is not None else None,
                "cuda_graphs": bool(graphs), "resets_per_step": self.resets_per_step,
                "chained": bool(self.chain) if mode == "step" else None}


DESC = {"rollout": "fused oc_rollout (up to R steps per launch, Philox actions drawn on the device, state on chip)",
        "step": "C-ABI oc_step, one launch per step, actions read from an HBM pool, CUDA graphs of min(K, R) launches; the "
                "steps of a graph form a chain (OC_FLAG_CHAIN_HEAD + OC_FLAG_CHAINED: a launch starts when the previous one "
                "has stored its states) unless --no-chain"}


# ------------------------------------------------------------------------------- main
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=2000)
    ap.add_argument("--warmup", type=int, default=20)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="cfg2", choices=sorted(WORKLOADS))
    ap.add_argument("--envs", type=int, default=0, help="envs per GPU (default: the workload's)")
    ap.add_argument("--ring", type=int, default=64, help="rollout-buffer slots the obs are written to")
    ap.add_argument("--mode", default="rollout", choices=["step", "rollout"],
                    help="headline mode. rollout: fused oc_rollout, the synthetic random-action rollout of SURVEY 8d; "
                         "step: one oc_step launch per step, actions read from HBM.  The other mode is measured too.")
    ap.add_argument("--single-mode", action="store_true", help="measure only --mode")
    ap.add_argument("--seconds", type=float, default=0.7, help="wall-clock budget of the repeats of one measurement")
    ap.add_argument("--cpu-seconds", type=float, default=20.0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-workloads", action="store_true", help="skip the other BASELINE configs (cfg3, cfg4, cfg5)")
    ap.add_argument("--no-graph", action="store_true", help="launch every step from Python instead of CUDA graphs")
    ap.add_argument("--stagger", default="spread", choices=["spread", "consecutive"],
                    help="episode-clock pattern of the steady state: (e * M) mod T with M coprime to T (the resets of a "
                         "step are spread over the batch) or e mod T (they fall on consecutive envs)")
    ap.add_argument("--no-chain", action="store_true",
                    help="step API: plain launches (each waits for the previous grid to retire) instead of chained ones")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    w = WORKLOADS[args.workload]

    if args.impl == "reference":
        if rank != 0:
            return 0
        t0 = time.time()
        res = cpu_reference_arm(args.workload, max(2.0, min(args.cpu_seconds, 60.0)))
        line = {"impl": "reference", "metric": METRIC, "value": res["value"],
                "unit": "agent-steps/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
                "ms_per_step": None, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": "int32+f64", "data": "synthetic",
                "config": headline_config(args.workload, args.envs or w["envs"], args.mode, max(2, args.ring), args.stagger),
                "cpu_baseline": res,
                "e2e": {"value": res["value"], "unit": "agent-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                "wall_s": time.time() - t0}
        print(json.dumps(line))
        return 0

    import torch
    import torch.distributed as dist
    from gym_comm_b200.sharding import bind_cpu_to_device

    assert torch.cuda.is_available(), "bench.py needs a CUDA device (no CPU fallback for the product path)"
    torch.cuda.set_device(local_rank)
    affinity = bind_cpu_to_device(local_rank) if world > 1 else False      # NUMA-local pinned buffers (e2e path)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    peaks = load_json("MEASURED_PEAKS.json")
    peak = float(peaks.get("hbm_gbs", 6650.0))
    peak_src = "measured (MEASURED_PEAKS.json hbm_gbs)" if "hbm_gbs" in peaks else "fallback 6650 GB/s (B200_PROFILING.md)"
    K, W_ = args.steps, max(args.warmup, 3)
    other_mode = "step" if args.mode == "rollout" else "rollout"

    E = args.envs or w["envs"]
    b = Bench(args.workload, E, rank, world, dev, max(2, args.ring), peak, peak_src, args.no_graph, args.stagger)
    b.chain = not args.no_chain
    A, F, R, P = b.A, b.F, b.R, b.P
    ns = b.ns
    primary = b.measure(args.mode, K, W_, seconds=args.seconds)
    secondary = None if args.single_mode else b.measure(other_mode, K, W_, seconds=args.seconds)
    # the same per-step launches WITHOUT chaining: what a closed loop gets (a policy that needs step N's observations
    # before it can produce step N+1's actions -- train_ppo.py's graphed rollout): each launch waits for the previous grid
    plain = None
    if not args.single_mode and not args.no_chain:
        b.chain = False
        plain = b.measure("step", K, W_, seconds=0.3, min_reps=10, diagnostics=False, clocks=False)
        b.chain = True
    # open-loop sequences in one launch: oc_replay = the fused kernel reading a caller-given action sequence
    replay = None
    if not args.single_mode:
        replay = b.measure("replay", K, W_, seconds=0.3, min_reps=10, diagnostics=False, clocks=False)

    # ---- e2e: the reference-facing call with HOST buffers: OvercookedHostVecEnv.step = C ABI oc_step_host*
    # (pinned numpy buffers; every step copies the actions host->device, runs the step kernel, copies
    # observations / rewards / dones device->host and synchronises before returning).  Timed for a fixed
    # wall-clock window (>= 0.25 s), whatever K is.
    host_action_src = b.actions[:8].cpu().numpy()

    def measure_e2e(fmt, term=False, window=0.3):
        import numpy as np
        from gym_comm_b200.host_env import OvercookedHostVecEnv
        henv = OvercookedHostVecEnv(ns, num_envs=E, device_index=local_rank, seed=1234 + rank, auto_reset=True,
                                    terminal_observations=term, obs_format=fmt)
        try:
            host_actions = []                              # the steps' inputs live in pinned host memory, in the dtype
            for i in range(8):                             # the entry point takes (u8 pairs on the one-block path)
                pa = henv.pinned_array((E, A, 2), henv.action_dtype)
                pa[...] = host_action_src[i]
                host_actions.append(pa)
            henv.reset()
            # steady state of a long run: env e starts e mod T steps into its episode, so that about E / T envs
            # finish in every step (and, with `term`, their terminal rows are delivered)
            T = int(w["max_num_timesteps"])
            henv.stagger_clocks(T, spread_multiplier(T) if args.stagger == "spread" else 1)
            for i in range(5):
                henv.step(host_actions[i % 8])
            b.barrier()
            nfin, nsampled, n = 0, 0, 0
            t0 = time.perf_counter()
            while True:
                for i in range(25):
                    d_ = henv.step(host_actions[(n + i) % 8])[2]   # returns with obs / reward / done valid on the host
                nfin += int(np.count_nonzero(d_))                  # bookkeeping of the bench itself: sampled, one step in 25
                nsampled += 1
                n += 25
                if time.perf_counter() - t0 >= window:
                    break
            dt = time.perf_counter() - t0
            h2d, d2h = henv.h2d_bytes_per_step, henv.d2h_bytes_per_step
            launches = henv.kernel_launches_per_step
            desc = henv.transfer_desc
        finally:
            henv.close()
        # every rank ran its own number of steps inside the same window: the job's rate is the sum of the ranks' rates
        rate = torch.tensor([float(E) * A * n / dt], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(rate, op=dist.ReduceOp.SUM)
        entry = "oc_step_host_i8" if fmt == "i8" else "oc_step_host"
        return {"value": float(rate.item()), "unit": "agent-steps/s",
                "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                "steps": n, "window_s": dt, "us_per_step": dt / n * 1e6,
                "obs_format": ("int8 [E,A,F-1] + f32 clock [E]" if fmt == "i8" else "f32 [E,A,F]"),
                "lossless": True,       # i8: the integer keys as the integers the reference's get_observation2 builds
                                        # (int64 arrays), the clock still f32 -- value-identical to the float rows
                                        # (tests/test_gpu_host_env.py, tests/cabi_smoke.c); nothing is quantised
                "api": "OvercookedHostVecEnv(obs_format=%r).step = C ABI %s, pinned numpy buffers, results valid on return" % (fmt, entry),
                "transfers": desc, "gpu_launches_per_step": launches,
                "terminal_observations": bool(term), "finished_envs_per_step": nfin / nsampled,
                "cpu_affinity": "nvml (GPU-local cores)" if affinity else "none"}

    e2e = e2e_f32 = e2e_term = None
    if not args.no_e2e:
        e2e_f32 = measure_e2e("f32")
        try:                                   # the full SB3 contract: infos[e]["terminal_observation"] for finished envs
            e2e_term = measure_e2e("i8", term=True)
        except Exception as ex:
            e2e_term = {"error": repr(ex)}
        try:
            e2e = measure_e2e("i8")
        except Exception as ex:                # every rank takes the same path: the failure modes are build-level
            e2e = dict(e2e_f32, note="compact format failed (%r); this is the float format" % (ex,))
    b.close()

    # ---- the other BASELINE configs, same protocol, shorter repeat budget
    others = {}
    if not args.no_workloads:
        for name in ("cfg3", "cfg4", "cfg5"):
            if name == args.workload:
                continue
            try:
                ob = Bench(name, WORKLOADS[name]["envs"], rank, world, dev, max(2, args.ring), peak, peak_src, args.no_graph,
                           args.stagger)
                ob.chain = not args.no_chain
                r1 = ob.measure("rollout", K, W_, seconds=0.3, min_reps=10, diagnostics=False, clocks=False)
                r2 = ob.measure("step", K, W_, seconds=0.3, min_reps=10, diagnostics=False, clocks=False)
                others[name] = {"config": workload_config(name, ob.E, "rollout")[0], "value": r1["value"],
                                "unit": "agent-steps/s", "ms_per_step": r1["ms_per_step"], "roofline": r1["roofline"],
                                "repeats": r1["repeats"], "gpu_launches": r1["gpu_launches"],
                                "resets_per_step": r1["resets_per_step"],
                                "step_api": {"value": r2["value"], "ms_per_step": r2["ms_per_step"],
                                             "roofline": r2["roofline"], "repeats": r2["repeats"],
                                             "gpu_launches": r2["gpu_launches"]}}
                ob.close()
            except Exception as ex:            # a workload that does not fit must not hide the headline
                others[name] = {"error": repr(ex)}

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        try:
            cpu = cpu_reference_arm(args.workload, args.cpu_seconds)
        except Exception as ex:  # the checker failing must not hide the GPU number
            cpu = {"error": repr(ex)}

    if rank == 0:
        line = {
            "metric": METRIC, "value": primary["value"], "unit": "agent-steps/s",
            "n_gpus": world, "steps": K, "warmup": W_, "ms_per_step": primary["ms_per_step"],
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "int32+f64", "data": "synthetic",
            "config": headline_config(args.workload, E, args.mode, max(2, args.ring), args.stagger),
            "protocol": {"cuda_graphs": primary["cuda_graphs"], "resets_per_step": primary["resets_per_step"],
                         "repeats": primary["repeats"]["n"], "statistic": "median repeat, max over ranks per repeat"},
            "repeats": primary["repeats"],
            "roofline": primary["roofline"], "cpu_baseline": cpu, "e2e": e2e, "e2e_f32": e2e_f32, "e2e_terminal_obs": e2e_term,
            "gpu_launches": primary["gpu_launches"], "clocks": primary["clocks"],
        }
        if secondary is not None:
            line[other_mode + "_api"] = {"desc": DESC[other_mode], "value": secondary["value"], "unit": "agent-steps/s",
                                         "ms_per_step": secondary["ms_per_step"], "roofline": secondary["roofline"],
                                         "repeats": secondary["repeats"],
                                         "gpu_launches": secondary["gpu_launches"], "clocks": secondary["clocks"]}
        if plain is not None:
            line["step_api_unchained"] = {
                "desc": "C-ABI oc_step, one launch per step, plain launches with programmatic dependent launch (each grid "
                        "waits for the previous one): the closed-loop case, where step N+1's actions depend on step N's output",
                "value": plain["value"], "unit": "agent-steps/s", "ms_per_step": plain["ms_per_step"],
                "roofline": plain["roofline"], "repeats": plain["repeats"], "gpu_launches": plain["gpu_launches"]}
        if replay is not None:
            line["replay_api"] = {
                "desc": "C-ABI oc_replay: up to R steps per launch on a caller-given action sequence int32 [n, E, A, 2] read from "
                        "HBM (recorded / scripted / pre-drawn actions), state on chip, auto-reset on",
                "value": replay["value"], "unit": "agent-steps/s", "ms_per_step": replay["ms_per_step"],
                "roofline": replay["roofline"], "repeats": replay["repeats"], "gpu_launches": replay["gpu_launches"]}
        if others:
            line["workloads"] = others
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
