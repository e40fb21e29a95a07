#!/usr/bin/env python
"""bench.py -- MCTS simulations/s (and executed env-steps/s) of the root-parallel hot path on B200.

Workload (BASELINE.json config C5, the configuration the metric "at 1/2/4/8 B200" is quoted on):
vanilla MCTS on CartPole-v1, 65 536 independent trees PER GPU x nsim=1000, C=1, gamma=0.99,
max_depth=500, one random rollout per leaf, 256 distinct roots per GPU-group (each root searched by
trees/256 trees on every rank; root statistics merged by a per-rank segmented sum + one NCCL all-reduce).
One "step" = one full pass: reset trees -> 1000 simulations per tree -> root statistics -> merge -> arg-max.

    python bench.py [--gpus N] [--steps K] [--warmup W]            # this repo's CUDA engine
    python bench.py --impl reference [--gpus N] ...                # the reference's CPU path (Python port, all cores)
    torchrun ... bench.py --gpus N ...                             # N ranks, one per GPU (weak scaling)

Prints ONE JSON line (rank 0).
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOAD = "C5: vanilla MCTS CartPole-v1, 65536 trees/GPU x nsim=1000, C=1, gamma=0.99, max_depth=500, 1 rollout/leaf"
FLOP_PER_STEP_CARTPOLE = 36
BYTES_PER_LEVEL_A2 = 12 * 2 + 36     # B_level = 12*A + 36
STATE_BYTES = 16


def parse():
    p = argparse.ArgumentParser()
    p.add_argument("--gpus", type=int, default=1)
    p.add_argument("--steps", type=int, default=5)
    p.add_argument("--warmup", type=int, default=3)
    p.add_argument("--impl", default="isla", choices=["isla", "reference"])
    p.add_argument("--trees", type=int, default=65536, help="trees per GPU")
    p.add_argument("--nsim", type=int, default=1000)
    p.add_argument("--roots", type=int, default=256, help="distinct roots (root groups)")
    p.add_argument("--max_depth", type=int, default=500)
    p.add_argument("--c", type=float, default=1.0)
    p.add_argument("--gamma", type=float, default=0.99)
    p.add_argument("--precision", default="f32", choices=["f32", "f64"])
    p.add_argument("--cpu_trees", type=int, default=0, help="trees in the CPU-baseline sample (0 = auto)")
    p.add_argument("--no_cpu_baseline", action="store_true")
    p.add_argument("--no_extras", action="store_true")
    return p.parse_args()


# --------------------------------------------------------------------------- clocks
class ClockSampler:
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index=0):
        self.samples, self.proc, self.index = [], None, index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.samples.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, smax, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for s in self.samples:
            f = [x.strip() for x in s.split(",")]
            if len(f) < 6:
                continue
            try:
                sm.append(float(f[0])); smax.append(float(f[1]))
            except ValueError:
                continue
            for n, v in zip(names, f[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(smax) if smax else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# --------------------------------------------------------------------------- CPU baseline (oracle = checker/baseline only)
def _cpu_fit_worker(args):
    from oracle import py_port
    n_sim, C, gamma, max_depth, seed = args
    dt, sims, steps = py_port.time_vanilla_fit("CartPole-v1", n_sim, C, gamma, max_depth, seed)
    return dt, sims, steps


def cpu_baseline_python(args, cores, trees):
    """The reference's CPU path (pure-Python port, oracle/py_port.py), `cores` processes, `trees` fits."""
    import multiprocessing as mp
    jobs = [(args.nsim, args.c, args.gamma, args.max_depth, s) for s in range(trees)]
    t0 = time.perf_counter()
    if cores == 1:
        res = [_cpu_fit_worker(j) for j in jobs]
    else:
        with mp.get_context("fork").Pool(cores) as pool:
            res = pool.map(_cpu_fit_worker, jobs, chunksize=1)
    wall = time.perf_counter() - t0
    sims = sum(r[1] for r in res)
    steps = sum(r[2] for r in res)
    return wall, sims, steps


def cpu_baseline_native(args, threads, trees):
    """Extra, stronger baseline: the plain-C oracle on all host threads (not the reference's speed)."""
    import numpy as np
    from oracle import coracle
    roots = np.stack([coracle.env_reset(0, 0, i) for i in range(trees)])
    cfg = coracle.make_config(0, n_actions=2, max_depth=args.max_depth, C_=args.c, gamma=args.gamma, seed=0)
    t0 = time.perf_counter()
    _, _, cnt = coracle.run_many(cfg, roots, args.nsim, n_threads=threads)
    wall = time.perf_counter() - t0
    return wall, cnt["sims"], cnt["env_steps"]


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    trees = args.cpu_trees or cores
    for _ in range(min(args.warmup, 1)):
        cpu_baseline_python(args, cores, max(1, cores // 4))
    walls, sims, steps = [], 0, 0
    for _ in range(args.steps):
        w, s, st = cpu_baseline_python(args, cores, trees)
        walls.append(w); sims += s; steps += st
    total = sum(walls)
    value = sims / total
    line = {
        "metric": "mcts_simulations_per_second", "value": value, "unit": "sims/s", "impl": "reference",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * total / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": WORKLOAD, "sample": f"{trees} trees x nsim={args.nsim} per step"},
        "env_steps_per_s": steps / total,
        "cpu_baseline": {"value": value, "unit": "sims/s", "cores": cores, "kind": "port",
                         "sample": f"{trees} independent CartPole fits x nsim={args.nsim} per step on {cores} processes "
                                   f"(oracle/py_port.py: pure-Python port of MctsHash.fit, same per-node work as the reference)"},
        "e2e": {"value": value, "unit": "sims/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# --------------------------------------------------------------------------- this repo's engine
def run_isla(args):
    import numpy as np
    import torch
    import torch.distributed as dist

    from islamcts_b200 import _lib
    from islamcts_b200.batch import RootParallelSearch

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- the engine has no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    pg = None
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
        pg = dist.group.WORLD
    prec = _lib.PRECISION_F64 if args.precision == "f64" else _lib.PRECISION_F32
    search = RootParallelSearch("CartPole-v1", trees_per_rank=args.trees, n_sim=args.nsim, C=args.c, gamma=args.gamma,
                                max_depth=args.max_depth, n_roots=args.roots, seed=0, precision=prec, device=dev,
                                process_group=pg, rank=rank, world_size=world)
    # synthetic roots: CartPole reset distribution U(-0.05,0.05)^4, seed = root id (same on every rank)
    g = torch.Generator().manual_seed(1234)
    roots_host = ((torch.rand((args.roots, 4), generator=g, dtype=torch.float64) - 0.5) * 0.1).pin_memory()
    tree_roots = search.expand_roots(roots_host)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    def timed(fn, steps):
        barrier()
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ev0.record()
        for _ in range(steps):
            fn()
        ev1.record()
        barrier()
        ms = torch.tensor([ev0.elapsed_time(ev1)], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item())

    # ---- device-resident pass
    for _ in range(args.warmup):
        search.search_device(tree_roots)
    cnt0 = search.engine.counters()
    clocks = ClockSampler(local)
    if rank == 0:
        clocks.start()
    ms = timed(lambda: search.search_device(tree_roots), args.steps)
    clk = clocks.stop() if rank == 0 else None
    cnt1 = search.engine.counters()
    # counters are reset by set_roots each pass: cnt1 is the last pass only
    sims_per_pass = args.trees * args.nsim
    total_sims = sims_per_pass * world * args.steps
    value = total_sims / (ms * 1e-3)

    # ---- kernel-only timing of the dominant kernel (the search kernel), same stream, CUDA events
    eng = search.engine
    k_ms = []
    for _ in range(max(1, min(args.steps, 3))):
        eng.set_roots(tree_roots)
        torch.cuda.synchronize(dev)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); eng.run(args.nsim); e1.record()
        torch.cuda.synchronize(dev)
        k_ms.append(e0.elapsed_time(e1))
    kern_ms = sum(k_ms) / len(k_ms)
    cnt = eng.counters()
    levels_per_sim = cnt["levels"] / max(1, cnt["sims"])
    alg_bytes = cnt["levels"] * BYTES_PER_LEVEL_A2 + cnt["sims"] * STATE_BYTES
    achieved_gbs = alg_bytes / (kern_ms * 1e-3) / 1e9
    peak, peak_src = 6650.0, "fallback"
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            peak = float(json.load(f)["hbm_gbs"]); peak_src = "measured"
    except Exception:
        pass

    # ---- end to end through the public API with HOST buffers (H2D roots, D2H actions + statistics)
    for _ in range(2):
        search.fit(roots_host)
    e2e_ms = timed(lambda: search.fit(roots_host), args.steps)
    e2e_value = total_sims / (e2e_ms * 1e-3)

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    line = {
        "metric": "mcts_simulations_per_second", "value": value, "unit": "sims/s", "n_gpus": world,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f64 tree statistics / " + args.precision + " env", "data": "synthetic",
        "config": {"workload": WORKLOAD, "trees_per_gpu": args.trees, "nsim": args.nsim, "roots": args.roots,
                   "C": args.c, "gamma": args.gamma, "max_depth": args.max_depth,
                   "l2": "working set (tree arena %.1f GB/GPU) larger than L2; trees are rebuilt every step"
                         % (eng.layout.total_bytes / 1e9)},
        "env_steps_per_s": cnt["env_steps"] * world / (kern_ms * 1e-3),
        "levels_per_sim": levels_per_sim, "rollout_steps_per_sim": cnt["rollout_steps"] / max(1, cnt["sims"]),
        "roofline": {"bound": "hbm", "achieved": achieved_gbs, "peak": peak, "unit": "GB/s", "frac": achieved_gbs / peak,
                     "traffic": None, "kernel": "tpt_search_kernel<CartPole,%s>" % args.precision,
                     "kernel_ms": kern_ms, "peak_source": peak_src,
                     "algorithmic_bytes": "levels*(12*A+36) + sims*16 (SURVEY 8d), A=2"},
        "e2e": {"value": e2e_value, "unit": "sims/s", "h2d_bytes_per_step": search.h2d_bytes(),
                "d2h_bytes_per_step": search.d2h_bytes(), "ms_per_step": e2e_ms / args.steps},
        "gpu_launches": 5 * args.steps,
        "clocks": clk,
    }
    if not args.no_cpu_baseline and world == 1:
        cores_avail = os.cpu_count() or 1
        w, s, st = cpu_baseline_python(args, 1, args.cpu_trees or 3)
        line["cpu_baseline"] = {"value": s / w, "unit": "sims/s", "cores": 1, "kind": "port",
                                "env_steps_per_s": st / w,
                                "sample": f"{args.cpu_trees or 3} CartPole fits x nsim={args.nsim}, 1 core "
                                          f"(oracle/py_port.py, pure-Python port of MctsHash.fit)"}
        nt = min(cores_avail, 64)
        w2, s2, st2 = cpu_baseline_native(args, nt, nt * 8)
        line["cpu_baseline_native_c"] = {"value": s2 / w2, "unit": "sims/s", "cores": nt, "kind": "port",
                                          "sample": f"{nt * 8} trees x nsim={args.nsim}, plain-C oracle on {nt} threads"}
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    args = parse()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_isla(args)


if __name__ == "__main__":
    main()
