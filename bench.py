#!/usr/bin/env python
"""bench.py -- MCTS simulations/s (and executed env-steps/s) of the root-parallel hot path on B200.

Workload (BASELINE.json config C5, the configuration the metric "at 1/2/4/8 B200" is quoted on):
vanilla MCTS on CartPole-v1, 65 536 independent trees IN TOTAL x nsim=1000, C=1, gamma=0.99, max_depth=500,
one random rollout per leaf, sharded over the N GPUs (65 536 / N trees per GPU: STRONG scaling), 256 distinct
roots x 256 trees (root statistics merged by one fixed-order kernel per rank + ONE NCCL all-reduce of the packed
[256, 2, 2] float64 array).  One "step" = one full pass: reset trees -> 1000 simulations per tree -> per-root
merge -> all-reduce -> arg-max.  Extras: the other two root layouts of SURVEY 8(e) (1 shared root, 65 536 distinct
roots) with their message sizes, the weak-scaling figure (65 536 trees PER GPU), configs C1-C4 with the reference's
CPU speed beside them, the rollout micro-benchmarks.

    python bench.py [--gpus N] [--steps K] [--warmup W]            # this repo's CUDA engine
    python bench.py --impl reference [--gpus N] ...                # the reference's own CPU classes, all host cores
    torchrun ... bench.py --gpus N ...                             # N ranks, one per GPU

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

WORKLOAD = "C5: vanilla MCTS CartPole-v1, 65536 trees x nsim=1000 sharded over the GPUs, C=1, gamma=0.99, max_depth=500, 1 rollout/leaf"
FLOP_PER_STEP_CARTPOLE = 36
BYTES_PER_LEVEL_A2 = 12 * 2 + 36     # B_level = 12*A + 36
STATE_BYTES = 16


def parse():
    p = argparse.ArgumentParser()
    p.add_argument("--gpus", type=int, default=1)
    p.add_argument("--steps", type=int, default=5)
    p.add_argument("--warmup", type=int, default=3)
    p.add_argument("--impl", default="isla", choices=["isla", "reference"])
    p.add_argument("--trees", type=int, default=65536, help="trees of the whole job (sharded over the GPUs)")
    p.add_argument("--weak", action="store_true", help="headline = weak scaling (--trees per GPU) instead of strong")
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


def cpu_baseline_port(args, cores, trees):
    """Fallback CPU arm: the pure-Python port of MctsHash.fit (oracle/py_port.py), `cores` processes, `trees` fits."""
    import multiprocessing as mp
    jobs = [(args.nsim, args.c, args.gamma, args.max_depth, s) for s in range(trees)]
    t0 = time.perf_counter()
    if cores == 1:
        res = [_cpu_fit_worker(j) for j in jobs]
    else:
        with mp.get_context("fork").Pool(cores) as pool:
            res = pool.map(_cpu_fit_worker, jobs, chunksize=1)
    wall = time.perf_counter() - t0
    return wall, sum(r[1] for r in res), sum(r[2] for r in res)


def cpu_baseline(args, cores, trees):
    """The reference's CPU path for the headline config: its OWN MctsHash.fit (unmodified files staged under
    baseline/_ref, imported through oracle/ref_loader.py) when present -> kind "reference"; else the port."""
    from oracle import ref_bench
    default_cfg = (args.nsim, args.c, args.gamma, args.max_depth) == (1000, 1.0, 0.99, 500)
    if ref_bench.reference_root() is not None and default_cfg:
        w, s, st = ref_bench.time_many("c5", trees, cores)
        return w, s, st, "reference", "the reference's own MctsHash.fit (islaMcts/agents/mcts_hash.py, staged unmodified in baseline/_ref) on the float64 env restatement"
    w, s, st = cpu_baseline_port(args, cores, trees)
    return w, s, st, "port", "oracle/py_port.py, pure-Python port of MctsHash.fit (same per-node work as the reference)"


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
        cpu_baseline(args, cores, max(1, cores // 4))
    walls, sims, steps = [], 0, 0
    kind = note = None
    for _ in range(args.steps):
        w, s, st, kind, note = cpu_baseline(args, cores, trees)
        walls.append(w); sims += s; steps += st
    total = sum(walls)
    value = sims / total
    line = {
        "metric": "mcts_simulations_per_second", "value": value, "unit": "sims/s", "impl": "reference",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * total / args.steps,
        "higher_is_better": True, "scaling": "weak" if args.weak else "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": WORKLOAD, "sample": f"{trees} trees x nsim={args.nsim} per step"},
        "env_steps_per_s": steps / total,
        "cpu_baseline": {"value": value, "unit": "sims/s", "cores": cores, "kind": kind,
                         "sample": f"{trees} independent CartPole fits x nsim={args.nsim} per step on {cores} processes ({note})"},
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
    from islamcts_b200 import engine as E
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

    def make_search(n_roots, weak=False, trees=None):
        kw = dict(trees_per_rank=trees or args.trees) if weak else dict(total_trees=trees or args.trees)
        return RootParallelSearch("CartPole-v1", n_sim=args.nsim, C=args.c, gamma=args.gamma, max_depth=args.max_depth,
                                  n_roots=n_roots, seed=0, precision=prec, device=dev, process_group=pg, rank=rank,
                                  world_size=world, **kw)

    def synthetic_roots(n_roots):
        # CartPole reset distribution U(-0.05,0.05)^4, same on every rank
        g = torch.Generator().manual_seed(1234)
        return ((torch.rand((n_roots, 4), generator=g, dtype=torch.float64) - 0.5) * 0.1).pin_memory()

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

    search = make_search(args.roots, weak=args.weak)
    roots_host = synthetic_roots(args.roots)
    tree_roots = search.expand_roots(roots_host)
    trees_job = search.trees * world

    # ---- device-resident pass
    for _ in range(args.warmup):
        search.search_device(tree_roots)
    clocks = ClockSampler(local)
    if rank == 0:
        clocks.start()
    l0 = E.kernel_launches()
    ms = timed(lambda: search.search_device(tree_roots), args.steps)
    launches = (E.kernel_launches() - l0) * world          # counted by the library; every rank launches the same kernels
    clk = clocks.stop() if rank == 0 else None
    total_sims = trees_job * args.nsim * args.steps
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

    # ---- the other root layouts of SURVEY 8(e) and the other scaling mode (same kernel; short runs)
    layouts = {}
    msg_main = search.merge_message_bytes()
    arena_gb = eng.layout.total_bytes / 1e9
    del search, eng, tree_roots
    torch.cuda.empty_cache()
    if not args.no_extras:
        steps_x = max(2, min(args.steps, 3))
        for name, n_roots in (("one_shared_root", 1), ("all_distinct_roots", args.trees)):
            try:
                sx = make_search(n_roots)
            except ValueError:
                continue
            rx = sx.expand_roots(synthetic_roots(n_roots))
            sx.search_device(rx)
            mx = timed(lambda: sx.search_device(rx), steps_x)
            layouts[name] = {"roots": n_roots, "sims_per_s": sx.trees * world * args.nsim * steps_x / (mx * 1e-3),
                             "ms_per_step": mx / steps_x, "allreduce_message_bytes": sx.merge_message_bytes()}
            del sx, rx
            torch.cuda.empty_cache()
        if world > 1 and not args.weak:
            sx = make_search(args.roots, weak=True)
            rx = sx.expand_roots(synthetic_roots(args.roots))
            sx.search_device(rx)
            mx = timed(lambda: sx.search_device(rx), steps_x)
            layouts["weak_scaling"] = {"trees_per_gpu": sx.trees, "roots": args.roots,
                                       "sims_per_s": sx.trees * world * args.nsim * steps_x / (mx * 1e-3),
                                       "ms_per_step": mx / steps_x, "allreduce_message_bytes": sx.merge_message_bytes()}
            del sx, rx
            torch.cuda.empty_cache()

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # MUFU (SFU) rate of the search kernel: 4 per re-evaluated level (two reciprocals, two square roots of the fp32 UCB
    # screening) + 1 per env step (the fast division of the float32 CartPole step) against SMs x 16 lanes x clock
    sm_count = torch.cuda.get_device_properties(dev).multi_processor_count
    mufu_per_s = (cnt["levels"] * 4 + cnt["env_steps"]) / (kern_ms * 1e-3)
    line = {
        "metric": "mcts_simulations_per_second", "value": value, "unit": "sims/s", "n_gpus": world,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True,
        "scaling": "weak" if args.weak else "strong", "vs_baseline": None,
        "dtype": "f64 tree statistics / " + args.precision + " env", "data": "synthetic",
        "config": {"workload": WORKLOAD, "trees_total": trees_job, "trees_per_gpu": trees_job // world, "nsim": args.nsim,
                   "roots": args.roots, "C": args.c, "gamma": args.gamma, "max_depth": args.max_depth,
                   "allreduce_message_bytes": msg_main,
                   "l2": "working set (tree arena %.1f GB/GPU) larger than L2; trees are rebuilt every step"
                         % arena_gb},
        "env_steps_per_s": cnt["env_steps"] * world / (kern_ms * 1e-3),
        "levels_per_sim": levels_per_sim, "rollout_steps_per_sim": cnt["rollout_steps"] / max(1, cnt["sims"]),
        "roofline": {"bound": "hbm", "achieved": achieved_gbs, "peak": peak, "unit": "GB/s", "frac": achieved_gbs / peak,
                     "traffic": None, "kernel": "tpt_search_kernel<CartPole,%s>" % args.precision,
                     "kernel_ms": kern_ms, "peak_source": peak_src,
                     "algorithmic_bytes": "levels*(12*A+36) + sims*16 (SURVEY 8d), A=2",
                     "note": "SURVEY's byte model counts 60 B per visited level; the kernel moves 36 (implicit visit counts, "
                             "returns read from a table instead of a stored row), so `frac` can exceed the share of DRAM time: see dram_frac",
                     "mufu_per_s": mufu_per_s, "mufu_frac": mufu_per_s / (sm_count * 16 * 1.965e9)},
        "e2e": {"value": e2e_value, "unit": "sims/s", "h2d_bytes_per_step": args.roots * 4 * 8,
                "d2h_bytes_per_step": args.roots * (4 + 2 * 16), "ms_per_step": e2e_ms / args.steps},
        "gpu_launches": launches,
        "clocks": clk,
    }
    if layouts:
        line["root_layouts"] = layouts
    try:
        with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
            tr = json.load(f)
        if tr.get("nsim") == args.nsim and tr.get("trees") == trees_job // world:
            line["roofline"]["traffic"] = tr["dram_bytes_per_launch"]
            line["roofline"]["traffic_source"] = tr["source"]
            line["roofline"]["dram_frac"] = tr["dram_bytes_per_launch"] / (kern_ms * 1e-3) / 1e9 / peak
    except Exception:
        pass
    if not args.no_extras and world == 1:
        line["extras"] = other_configs(dev)
    if not args.no_cpu_baseline and world == 1:
        cores_avail = os.cpu_count() or 1
        n_fit = args.cpu_trees or 3
        w, s, st, kind, note = cpu_baseline(args, 1, n_fit)
        line["cpu_baseline"] = {"value": s / w, "unit": "sims/s", "cores": 1, "kind": kind, "env_steps_per_s": st / w,
                                "sample": f"{n_fit} CartPole fits x nsim={args.nsim}, 1 core ({note})"}
        nt = min(cores_avail, 64)
        w2, s2, st2 = cpu_baseline_native(args, nt, nt * 8)
        line["cpu_baseline_native_c"] = {"value": s2 / w2, "unit": "sims/s", "cores": nt, "kind": "port",
                                          "sample": f"{nt * 8} trees x nsim={args.nsim}, plain-C oracle on {nt} threads"}
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def other_configs(dev):
    """Secondary measurements (NOT the headline): BASELINE configs C1-C4 on the CTA-per-tree engine and the
    rollout micro-benchmark R of SURVEY 8(d).  CUDA events, 1 warm-up + best of 3."""
    import torch
    from islamcts_b200 import _lib
    from islamcts_b200 import engine as E
    from islamcts_b200.engine import Engine

    def timeit(fn, reps=3):
        fn(); torch.cuda.synchronize(dev)
        best = 1e30
        for _ in range(reps):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); fn(); e1.record(); torch.cuda.synchronize(dev)
            best = min(best, e0.elapsed_time(e1))
        return best * 1e-3

    def ref_cpu(config, fits=1, n_sim=None):
        """The reference's own class for that config on ONE host core (oracle/ref_bench.py; HEAD semantics)."""
        try:
            from oracle import ref_bench
            if ref_bench.reference_root() is None:
                return {"unavailable": "baseline/_ref not staged"}
            w, s_, st = ref_bench.time_many(config, fits, 1, n_sim)
            sc = ref_bench.CONFIGS[config]
            return {"sims_per_s": s_ / w, "env_steps_per_s": st / w, "cores": 1, "kind": "reference",
                    "sample": "%d x %s.fit(), nsim=%d (reference HEAD: %s)" % (fits, sc["agent"], n_sim or sc["n_sim"],
                              "budgeted rollouts" if sc["agent"] == "MctsHash" else "zero-length rollouts")}
        except Exception as e:   # the baseline is reported, never required
            return {"unavailable": repr(e)[:200]}

    out = {}
    g = torch.Generator().manual_seed(7)
    # C1: vanilla FrozenLake nsim=100 C=0.5 gamma=1 max_depth=1000: one tree (the library's own kernel choice, as Mcts.fit()
    # gets it), and 65 536 trees (thread-per-tree)
    for name, nt, kern in (("c1_frozenlake_1tree", 1, _lib.KERNEL_AUTO), ("c1_frozenlake_65536trees", 65536, _lib.KERNEL_THREAD_PER_TREE)):
        eng = Engine(_lib.ENV_FROZENLAKE, n_trees=nt, sim_capacity=100, max_depth=1000, C_=0.5, gamma=1.0, kernel=kern, seed=1, device=dev)
        roots = torch.zeros((nt, 1), dtype=torch.float64, device=dev)
        t = timeit(lambda: eng.set_roots(roots).run(100))
        out[name] = {"sims_per_s": nt * 100 / t, "ms": t * 1e3}
    # the reference's Mcts.fit passes max_depth as the depth (zero-length rollouts): the same semantics on the device
    eng = Engine(_lib.ENV_FROZENLAKE, n_trees=1, sim_capacity=100, max_depth=1000, C_=0.5, gamma=1.0, budget_mode=_lib.BUDGET_ZERO,
                 kernel=_lib.KERNEL_THREAD_PER_TREE, seed=1, device=dev)
    roots = torch.zeros((1, 1), dtype=torch.float64, device=dev)
    t = timeit(lambda: eng.set_roots(roots).run(100))
    out["c1_frozenlake_1tree_reference_head"] = {"sims_per_s": 100 / t, "ms": t * 1e3, "cpu_reference": ref_cpu("c1", fits=20)}
    # C2: one CartPole tree, nsim=1000, 4096 rollouts per leaf
    eng = Engine(_lib.ENV_CARTPOLE, n_trees=1, sim_capacity=1000, max_depth=500, C_=1.0, gamma=0.99, kernel=_lib.KERNEL_CTA_PER_TREE,
                 rollouts_per_leaf=4096, seed=1, device=dev)
    roots = (torch.rand((1, 4), generator=g, dtype=torch.float64) - 0.5) * 0.1
    t = timeit(lambda: eng.set_roots(roots).run(1000))
    c = eng.counters()
    out["c2_cartpole_1tree_4096rollouts"] = {"sims_per_s": 1000 / t, "env_steps_per_s": c["env_steps"] / t, "ms": t * 1e3,
                                             "cpu_reference": ref_cpu("c2", fits=1)}
    # the same shape as a batch of 296 trees (2 CTAs per SM): what the GPU does with C2 when it is not alone
    eng = Engine(_lib.ENV_CARTPOLE, n_trees=296, sim_capacity=1000, max_depth=500, C_=1.0, gamma=0.99, kernel=_lib.KERNEL_CTA_PER_TREE,
                 rollouts_per_leaf=4096, seed=1, device=dev)
    roots = (torch.rand((296, 4), generator=g, dtype=torch.float64) - 0.5) * 0.1
    t = timeit(lambda: eng.set_roots(roots).run(1000), reps=2)
    c = eng.counters()
    out["c2_cartpole_296trees_4096rollouts"] = {"sims_per_s": 296 * 1000 / t, "env_steps_per_s": c["env_steps"] / t, "ms": t * 1e3}
    # C3: APW Pendulum k=60 alpha=0.8 nsim=10000 max_depth=200 ; C4: DPW MountainCarContinuous alpha=0.5 nsim=10000 max_depth=500
    eng = Engine(_lib.ENV_PENDULUM, agent=_lib.AGENT_APW, n_trees=1, sim_capacity=10000, max_depth=200, C_=1.0, gamma=0.99, alpha1=0.8,
                 k1=60.0, kernel=_lib.KERNEL_CTA_PER_TREE, rollouts_per_leaf=1, seed=1, device=dev)
    roots = torch.tensor([[1.0, 0.3]], dtype=torch.float64)
    t = timeit(lambda: eng.set_roots(roots).run(10000), reps=2)
    c = eng.counters()
    out["c3_apw_pendulum_1tree"] = {"sims_per_s": 10000 / t, "env_steps_per_s": c["env_steps"] / t, "ms": t * 1e3}
    engz = Engine(_lib.ENV_PENDULUM, agent=_lib.AGENT_APW, n_trees=1, sim_capacity=10000, max_depth=200, C_=1.0, gamma=0.99, alpha1=0.8,
                  k1=60.0, kernel=_lib.KERNEL_CTA_PER_TREE, rollouts_per_leaf=1, budget_mode=_lib.BUDGET_ZERO, seed=1, device=dev)
    tz = timeit(lambda: engz.set_roots(roots).run(10000), reps=2)
    out["c3_apw_pendulum_1tree_reference_head"] = {"sims_per_s": 10000 / tz, "ms": tz * 1e3, "cpu_reference": ref_cpu("c3")}
    del engz
    # post-fit analytics of the sweep loop (sweep.py:199-200) on that 10 000-child tree: device kernel vs decoding the whole tree
    import time as _time
    eng.depth_stats(0); torch.cuda.synchronize(dev)
    t0 = _time.perf_counter(); eng.depth_stats(0); t_dev = _time.perf_counter() - t0
    t0 = _time.perf_counter(); eng.export_tree(0); t_host = _time.perf_counter() - t0
    out["analytics_c3_tree"] = {"depth_stats_device_ms": t_dev * 1e3, "full_tree_decode_host_ms": t_host * 1e3}
    eng = Engine(_lib.ENV_MOUNTAINCAR, agent=_lib.AGENT_DPW, n_trees=1, sim_capacity=10000, max_depth=500, C_=1.0, gamma=0.99,
                 alpha1=0.5, k1=1.0, alpha2=0.5, k2=1.0, kernel=_lib.KERNEL_CTA_PER_TREE, seed=1, device=dev)
    roots = torch.tensor([[-0.5, 0.0]], dtype=torch.float64)
    t = timeit(lambda: eng.set_roots(roots).run(10000), reps=2)
    c = eng.counters()
    out["c4_dpw_mountaincar_1tree"] = {"sims_per_s": 10000 / t, "env_steps_per_s": c["env_steps"] / t, "ms": t * 1e3}
    engz = Engine(_lib.ENV_MOUNTAINCAR, agent=_lib.AGENT_DPW, n_trees=1, sim_capacity=10000, max_depth=500, C_=1.0, gamma=0.99,
                  alpha1=0.5, k1=1.0, alpha2=0.5, k2=1.0, kernel=_lib.KERNEL_CTA_PER_TREE, budget_mode=_lib.BUDGET_ZERO, seed=1, device=dev)
    tz = timeit(lambda: engz.set_roots(roots).run(10000), reps=2)
    # the same search with 64 simulations in flight (opt-in tree-parallel mode, virtual loss): a different search, statistics
    # compared in tests/test_gpu_tree_cta.py::test_tree_parallel_virtual_loss_statistics
    engp = Engine(_lib.ENV_MOUNTAINCAR, agent=_lib.AGENT_DPW, n_trees=1, sim_capacity=10000, max_depth=500, C_=1.0, gamma=0.99,
                  alpha1=0.5, k1=1.0, alpha2=0.5, k2=1.0, kernel=_lib.KERNEL_CTA_PER_TREE, seed=1, device=dev, tree_parallel=64, virtual_value=-5.0)
    tp_ = timeit(lambda: engp.set_roots(roots).run(10000), reps=2)
    cp_ = engp.counters()
    out["c4_dpw_mountaincar_1tree_tree_parallel64"] = {"sims_per_s": 10000 / tp_, "env_steps_per_s": cp_["env_steps"] / tp_, "ms": tp_ * 1e3}
    del engp
    out["c4_dpw_mountaincar_1tree_reference_head"] = {"sims_per_s": 10000 / tz, "ms": tz * 1e3, "cpu_reference": ref_cpu("c4")}
    del engz
    # CurveEnv (the env sweep.py / sweep_vanilla.py / examples/example_curve.py of the reference actually run), from an
    # on-track pose (three braking right turns after reset): APW random expansion, APW + voo as in example_curve.py:32-45,
    # vanilla MctsHash on DiscreteCurveEnv([5, 19]) (sweep_vanilla.py:177 defaults) as one tree and as 4096 trees
    pose = torch.tensor([[-5.70993649, 25.49759526, 2.5, 0.0, 3.0]], dtype=torch.float64)
    eng = Engine(_lib.ENV_CURVE, agent=_lib.AGENT_APW, n_trees=1, sim_capacity=1000, max_depth=500, C_=11.0, gamma=0.95, alpha1=0.5,
                 k1=3.0, kernel=_lib.KERNEL_CTA_PER_TREE, rollouts_per_leaf=32, seed=1, device=dev)
    t = timeit(lambda: eng.set_roots(pose).run(1000))
    c = eng.counters()
    out["curve_apw_1tree_32rollouts"] = {"sims_per_s": 1000 / t, "env_steps_per_s": c["env_steps"] / t, "ms": t * 1e3}
    eng = Engine(_lib.ENV_CURVE, agent=_lib.AGENT_APW, n_trees=1, sim_capacity=100, max_depth=500, C_=11.837, gamma=0.4, alpha1=0.0,
                 k1=95.0, expansion=_lib.EXPAND_VOO, epsilon=0.5, voo_n_samples=100, voo_max_try=1000,
                 kernel=_lib.KERNEL_CTA_PER_TREE, seed=1, device=dev)
    t = timeit(lambda: eng.set_roots(pose).run(100))
    out["curve_apw_voo_example_curve_py"] = {"sims_per_s": 100 / t, "ms": t * 1e3}
    for name, nt in (("curve_vanilla_5x19_1tree", 1), ("curve_vanilla_5x19_4096trees", 4096)):
        eng = Engine(_lib.ENV_CURVE, agent=_lib.AGENT_VANILLA, n_trees=nt, sim_capacity=1000, max_depth=500, C_=11.0, gamma=0.95,
                     action_grid=(5, 19), kernel=_lib.KERNEL_CTA_PER_TREE, seed=1, device=dev)
        roots = pose.repeat(nt, 1)
        t = timeit(lambda: eng.set_roots(roots).run(1000), reps=2)
        c = eng.counters()
        out[name] = {"sims_per_s": nt * 1000 / t, "env_steps_per_s": c["env_steps"] / t, "ms": t * 1e3}
    n = 1 << 21
    poses = pose.repeat(n, 1)
    for name, grid in (("rollout_curve_box", None), ("rollout_curve_grid_5x19", (5, 19))):
        steps = []
        def run_curve():
            ret, st = E.rollout(_lib.ENV_CURVE, poses, budget=500, gamma=0.95, seed=1234, sim=0, device=dev, action_grid=grid)
            steps[:] = [st]
        t = timeit(run_curve, reps=3)
        total = int(steps[-1].sum().item())
        out[name] = {"trajectories": n, "env_steps_per_s": total / t, "mean_len": total / n, "ms": t * 1e3}
    # whole episodes in lock-step (sweep.py's loop for --n_episodes 100 at once): CartPole, vanilla, nsim=1000, 100 real steps
    from islamcts_b200.episodes import EpisodeBatch
    eb = EpisodeBatch("CartPole-v1", n_episodes=100, n_sim=1000, max_steps=100, seed=1, max_depth=500, C_=1.0, gamma=0.99, device=dev)
    st0 = (torch.rand((100, 4), generator=g, dtype=torch.float64) - 0.5) * 0.1
    eb.run(st0)
    torch.cuda.synchronize(dev)
    import time as _t
    t0 = _t.perf_counter(); res = eb.run(st0); dt = _t.perf_counter() - t0
    out["episodes_cartpole_100_lockstep"] = {"real_steps": res["real_steps"], "seconds": dt, "real_steps_per_s": res["real_steps"] / dt,
                                             "sims_per_s": res["real_steps"] * 1000 / dt, "mean_steps": float(res["n_steps"].mean())}
    # Agent(param).fit() wall latency through the drop-in classes, a NEW agent per fit as the reference's episode loop builds
    # one per real step (sweep.py:186): host work (env identification, engine from the pool, re-key, reset), the search
    # launch, the arg-max epilogue and the read-back of the action and root statistics.  `without_engine_pool` re-creates
    # the device engine for every agent (what round 1 did).
    out["agent_fit_latency"] = agent_fit_latency(dev)
    # R: rollout micro-benchmark, 2^24 trajectories from the reset distribution, random policy, budget = max_depth
    flop = {0: 36, 1: 27, 2: 27, 3: 6}
    mufu = {0: 4, 1: 2, 2: 1, 3: 0}     # CartPole 1 sin + 1 cos + 2 div, Pendulum 1 sin + 1 mod, MountainCar 1 cos
    peak_fp32 = 148 * 128 * 2 * 1.965e9
    peak_sfu = 148 * 16 * 1.965e9
    for env_id, name, budget, gamma in ((0, "cartpole", 500, 0.99), (1, "pendulum", 200, 0.99), (2, "mountaincar", 500, 0.99), (3, "frozenlake", 1000, 1.0)):
        n = 1 << 24 if env_id in (0, 3) else 1 << 21
        steps = []
        def run():
            ret, st = E.rollout(env_id, None, n=n, budget=budget, gamma=gamma, seed=1234, sim=0, device=dev)
            steps[:] = [st]
        t = timeit(run, reps=5)
        total = int(steps[-1].sum().item())
        out["rollout_" + name] = {"trajectories": n, "env_steps_per_s": total / t, "mean_len": total / n, "ms": t * 1e3,
                                  "fp32_roofline_frac": total / t * flop[env_id] / peak_fp32,
                                  "roofline": {"bound": "fp32 issue", "achieved": total / t * flop[env_id] / 1e12, "peak": peak_fp32 / 1e12,
                                               "unit": "TFLOP/s", "frac": total / t * flop[env_id] / peak_fp32,
                                               "mufu_frac": total / t * mufu[env_id] / peak_sfu}}
    return out


def agent_fit_latency(dev):
    import time as _t
    import warnings
    import torch
    from islamcts_b200 import envs
    from islamcts_b200.agents import Mcts, MctsApw, MctsDpw, MctsHash
    from islamcts_b200.agents.abstract_mcts import clear_engine_pool
    from islamcts_b200.agents.parameters.dpw_parameters import DpwParameters
    from islamcts_b200.agents.parameters.mcts_parameters import MctsParameters
    from islamcts_b200.agents.parameters.pw_parameters import PwParameters
    from islamcts_b200.utils.action_selection_functions import continuous_default_policy, ucb1
    warnings.simplefilter("ignore")

    def common(env, obs, n_sim, C, gamma, max_depth, **kw):
        return dict(root_data=obs, env=env, n_sim=n_sim, C=C, action_selection_fn=ucb1, gamma=gamma,
                    rollout_selection_fn=continuous_default_policy, max_depth=max_depth,
                    n_actions=getattr(env.action_space, "n", None), seed=1, device=str(dev), **kw)

    cases = {
        "c1_mcts_frozenlake_nsim100": ("FrozenLake-v1", lambda e, o: Mcts(MctsParameters(**common(e, o, 100, 0.5, 1.0, 1000)))),
        "c2_mctshash_cartpole_nsim1000": ("CartPole-v1", lambda e, o: MctsHash(MctsParameters(**common(e, o, 1000, 1.0, 0.99, 500)))),
        "c2_mctshash_cartpole_nsim1000_4096rollouts": ("CartPole-v1", lambda e, o: MctsHash(MctsParameters(
            **common(e, o, 1000, 1.0, 0.99, 500, rollouts_per_leaf=4096)))),
        "c3_mctsapw_pendulum_nsim10000": ("Pendulum-v1", lambda e, o: MctsApw(PwParameters(
            **common(e, o, 10000, 1.0, 0.99, 200), alpha=0.8, k=60, action_expansion_function=continuous_default_policy))),
        "c4_mctsdpw_mountaincar_nsim10000": ("MountainCarContinuous-v0", lambda e, o: MctsDpw(DpwParameters(
            **common(e, o, 10000, 1.0, 0.99, 500), alphaApw=0.5, kApw=1, alphaSpw=0.5, kSpw=1))),
    }
    res = {}
    for name, (env_name, make_agent) in cases.items():
        env = envs.make(env_name, api=4, seed=3, device=str(dev))
        obs = env.reset()

        def one_fit():
            ag = make_agent(env.unwrapped, obs)
            ag.root.data = obs
            return ag.fit()

        reps = 3 if "nsim10000" in name and "dpw" in name else 8
        entry = {}
        for label, pooled in (("ms_per_fit", True), ("ms_per_fit_without_engine_pool", False)):
            one_fit(); one_fit()
            torch.cuda.synchronize(dev)
            ts = []
            for _ in range(reps):
                if not pooled:
                    clear_engine_pool()
                t0 = _t.perf_counter(); one_fit(); ts.append(_t.perf_counter() - t0)
            ts.sort()
            entry[label] = 1e3 * ts[len(ts) // 2]
        res[name] = entry
    clear_engine_pool()
    return res


def main():
    args = parse()
    # stdout carries exactly ONE JSON line: libraries that print there (NCCL's version banner) are sent to stderr
    real_stdout = os.dup(1)
    os.dup2(2, 1)
    sys.stdout = os.fdopen(real_stdout, "w", buffering=1)
    if args.impl == "reference":
        run_reference(args)
    else:
        run_isla(args)


if __name__ == "__main__":
    main()
