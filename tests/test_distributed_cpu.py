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


def _worker_children(rank, world, port, q):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from islamcts_b200 import sharding
    from oracle import coracle
    # two APW trees per rank on the same Pendulum root (the oracle stands in for the device search); genetic2 adds the
    # same centre / bound actions to every tree, so the union has rows to sum
    root = coracle.env_reset(1, 3, 0)
    acts, nas, tots = [], [], []
    for t in range(2):
        tree = coracle.Tree(coracle.make_config(1, agent=1, max_depth=20, C_=2.0, gamma=0.95, alpha1=0.5, k1=2.0, expansion=2,
                                                epsilon=0.5, seed=4, tree_id=rank * 2 + t), root)
        tree.run(60 if rank == 0 else 25)      # ranks end up with different numbers of children
        a, na, tot = tree.root_children()
        acts.append(a); nas.append(na); tots.append(tot)
    la, lna, ltot = sharding.merge_root_children(torch.as_tensor(np.concatenate(acts))[:, None], torch.as_tensor(np.concatenate(nas)),
                                                 torch.as_tensor(np.concatenate(tots)))
    ga, gna, gtot = sharding.all_gather_root_children(la, lna, ltot, dist.group.WORLD)
    q.put((rank, np.concatenate(acts), np.concatenate(nas), np.concatenate(tots), ga.numpy(), gna.numpy(), gtot.numpy()))
    dist.destroy_process_group()


def test_two_rank_union_of_continuous_root_children():
    world, port = 2, 29731 + os.getpid() % 200
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker_children, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    out = sorted((q.get(timeout=120) for _ in range(world)), key=lambda o: o[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    # expected union: python dict keyed by the action's bytes, rank order then tree order then insertion order
    exp = {}
    for o in out:
        for a, n, t in zip(o[1], o[2], o[3]):
            k = np.float64(a).tobytes()
            if k not in exp:
                exp[k] = [float(a), 0, 0.0]
            exp[k][1] += int(n); exp[k][2] += float(t)
    ea = np.array([v[0] for v in exp.values()]); en = np.array([v[1] for v in exp.values()]); et = np.array([v[2] for v in exp.values()])
    assert len(ea) < sum(len(o[1]) for o in out)        # some actions were shared between trees
    for o in out:                                        # every rank holds the same union
        np.testing.assert_array_equal(o[4][:, 0], ea)
        np.testing.assert_array_equal(o[5], en)
        np.testing.assert_allclose(o[6], et, rtol=1e-12)
    assert en.sum() == 2 * 60 + 2 * 25
