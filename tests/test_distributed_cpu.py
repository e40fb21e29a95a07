"""N > 1 host logic on CPU: world_size-2 gloo.  Sharding arithmetic (tree ids, root groups) and the merge
all-reduce of the root statistics -- the device search itself is covered by the -m gpu tests."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from islamcts_b200 import sharding
    from oracle import coracle
    trees_per_rank, n_roots, n_sim, A = 8, 4, 40, 2
    plan = sharding.plan(rank, world, trees_per_rank, n_roots)
    roots = np.stack([coracle.env_reset(0, 3, i) for i in range(n_roots)])
    # the CPU oracle stands in for the device search (same tree ids / Philox keys as a rank would use)
    cfg = coracle.make_config(0, n_actions=A, max_depth=50, C_=1.0, gamma=0.95, seed=9, tree_id=plan.tree_id_base)
    na, tot, _ = coracle.run_many(cfg, roots[plan.group], n_sim, n_threads=2)
    mna, mtot = sharding.segmented_sum(torch.as_tensor(na), torch.as_tensor(tot), torch.as_tensor(plan.group), n_roots)
    sharding.all_reduce_root_stats(mna, mtot, dist.group.WORLD)
    q.put((rank, plan.tree_id_base, mna.numpy(), mtot.numpy(), sharding.pick(mna, mtot).numpy()))
    dist.destroy_process_group()


def test_two_rank_merge_equals_single_process_sum():
    world, port = 2, 29531 + os.getpid() % 200
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    out = sorted(q.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    sys.path.insert(0, ROOT)
    from islamcts_b200 import sharding
    from oracle import coracle
    trees_per_rank, n_roots, n_sim, A = 8, 4, 40, 2
    roots = np.stack([coracle.env_reset(0, 3, i) for i in range(n_roots)])
    ena = np.zeros((n_roots, A), np.int64); etot = np.zeros((n_roots, A))
    for r in range(world):
        plan = sharding.plan(r, world, trees_per_rank, n_roots)
        assert plan.tree_id_base == r * trees_per_rank and out[r][1] == plan.tree_id_base
        cfg = coracle.make_config(0, n_actions=A, max_depth=50, C_=1.0, gamma=0.95, seed=9, tree_id=plan.tree_id_base)
        na, tot, _ = coracle.run_many(cfg, roots[plan.group], n_sim, n_threads=2)
        np.add.at(ena, plan.group, na); np.add.at(etot, plan.group, tot)
    for r in range(world):                      # every rank holds the same merged statistics and picks the same actions
        np.testing.assert_array_equal(out[r][2], ena)
        np.testing.assert_allclose(out[r][3], etot, rtol=1e-12)
        np.testing.assert_array_equal(out[r][4], np.argmax(etot / ena, axis=1))
    assert ena.sum() == world * trees_per_rank * n_sim


def test_plan_arithmetic():
    sys.path.insert(0, ROOT)
    from islamcts_b200 import sharding
    p = sharding.plan(3, 8, 65536, 256)
    assert p.tree_id_base == 3 * 65536 and len(p.group) == 65536 and p.group.max() == 255
    assert (np.bincount(p.group) == 256).all()
    with pytest.raises(ValueError):
        sharding.plan(0, 2, 10, 3)
