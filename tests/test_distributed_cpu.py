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
        q_ = etot / ena
        assert (q_[np.arange(n_roots), out[r][4]] == q_.max(axis=1)).all()
    np.testing.assert_array_equal(out[0][4], out[1][4])
    assert ena.sum() == world * trees_per_rank * n_sim


def _worker_strong(rank, world, port, q, total_trees, n_roots):
    """Strong scaling (config C5 as named): the job's trees are split over the ranks; the packed [n_roots, A, 2] array
    is filled for the rank's own roots (zeros elsewhere) and ONE all-reduce is merge and gather at once."""
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from islamcts_b200 import sharding
    from oracle import coracle
    n_sim, A = 30, 2
    plan = sharding.plan_total(rank, world, total_trees, n_roots)
    roots = np.stack([coracle.env_reset(0, 3, i) for i in range(n_roots)])
    cfg = coracle.make_config(0, n_actions=A, max_depth=50, C_=1.0, gamma=0.95, seed=9, tree_id=plan.tree_id_base)
    na, tot, _ = coracle.run_many(cfg, roots[plan.root_offset + plan.group], n_sim, n_threads=2)
    lna, ltot = sharding.segmented_sum(torch.as_tensor(na), torch.as_tensor(tot), torch.as_tensor(plan.group), plan.local_roots)
    packed = torch.zeros((n_roots, A, 2), dtype=torch.float64)
    packed[plan.root_offset: plan.root_offset + plan.local_roots, :, 0] = lna.to(torch.float64)
    packed[plan.root_offset: plan.root_offset + plan.local_roots, :, 1] = ltot
    sharding.all_reduce_packed(packed, dist.group.WORLD)
    mna, mtot = packed[..., 0].to(torch.int64), packed[..., 1]
    q.put((rank, plan.tree_id_base, mna.numpy(), mtot.numpy(), sharding.pick(mna, mtot, seed=5).numpy()))
    dist.destroy_process_group()


@pytest.mark.parametrize("total_trees,n_roots", [(16, 4), (16, 1), (16, 16)])
def test_two_rank_strong_scaling_equals_one_rank(total_trees, n_roots):
    """The three root layouts of SURVEY 8(e) -- one shared root, batched ensembles, all distinct -- sharded over two ranks
    give the statistics of the unsharded job (same global tree ids, hence the same Philox streams)."""
    world, port = 2, 29931 + (os.getpid() + 7 * n_roots) % 200
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker_strong, args=(r, world, port, q, total_trees, n_roots)) for r in range(world)]
    for p in procs:
        p.start()
    out = sorted(q.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    sys.path.insert(0, ROOT)
    from islamcts_b200 import sharding
    from oracle import coracle
    n_sim, A = 30, 2
    one = sharding.plan_total(0, 1, total_trees, n_roots)
    roots = np.stack([coracle.env_reset(0, 3, i) for i in range(n_roots)])
    cfg = coracle.make_config(0, n_actions=A, max_depth=50, C_=1.0, gamma=0.95, seed=9, tree_id=0)
    na, tot, _ = coracle.run_many(cfg, roots[one.group], n_sim, n_threads=2)
    ena = np.zeros((n_roots, A), np.int64); etot = np.zeros((n_roots, A))
    np.add.at(ena, one.group, na); np.add.at(etot, one.group, tot)
    assert out[0][1] == 0 and out[1][1] == total_trees // 2
    for r in range(world):
        np.testing.assert_array_equal(out[r][2], ena)
        np.testing.assert_allclose(out[r][3], etot, rtol=1e-12)
    np.testing.assert_array_equal(out[0][4], out[1][4])       # same seed, same statistics: same picks on every rank


def test_plan_total_arithmetic():
    sys.path.insert(0, ROOT)
    from islamcts_b200 import sharding
    p = sharding.plan_total(3, 8, 65536, 256)          # 32 whole roots per rank
    assert (p.trees_per_rank, p.tree_id_base, p.local_roots, p.root_offset) == (8192, 3 * 8192, 32, 96)
    assert (np.bincount(p.group) == 256).all() and (np.diff(p.group) >= 0).all()
    p = sharding.plan_total(5, 8, 65536, 1)            # one shared root: every rank holds an eighth of it
    assert (p.local_roots, p.root_offset, int(p.group.max())) == (1, 0, 0)
    p = sharding.plan_total(5, 8, 65536, 4)            # a root spans two ranks
    assert (p.local_roots, p.root_offset) == (1, 2)
    p = sharding.plan_total(1, 2, 65536, 65536)        # all distinct
    assert (p.local_roots, p.root_offset, p.tree_id_base) == (32768, 32768, 32768)
    with pytest.raises(ValueError):
        sharding.plan_total(0, 3, 65536, 256)


def test_pick_breaks_exact_ties_uniformly_and_reproducibly():
    """mcts_hash.py:42-50 draws uniformly among the children with the maximal q; sparse-reward roots (all q == 0 on
    FrozenLake) must not always answer action 0 (round-1 advisor finding)."""
    sys.path.insert(0, ROOT)
    from islamcts_b200 import sharding
    n = 4000
    mna = torch.full((n, 4), 25, dtype=torch.int64)
    mtot = torch.zeros((n, 4), dtype=torch.float64)
    a = sharding.pick(mna, mtot, seed=11)
    assert torch.equal(a, sharding.pick(mna, mtot, seed=11))            # ranks with the same seed agree
    assert not torch.equal(a, sharding.pick(mna, mtot, seed=12))
    share = np.bincount(a.numpy(), minlength=4) / n
    assert (abs(share - 0.25) < 0.04).all(), share
    # a unique maximum is always taken; unvisited children never win; two-way ties stay inside the tied set
    mtot[:, 2] = 5.0
    assert (sharding.pick(mna, mtot, seed=3) == 2).all()
    mtot[:, 0] = 5.0
    mna[:, 3] = 0
    b = sharding.pick(mna, mtot, seed=3).numpy()
    assert set(np.unique(b)) == {0, 2} and abs((b == 0).mean() - 0.5) < 0.05
    i, _ = sharding.pick_continuous(torch.arange(3.0)[:, None], torch.tensor([2, 2, 2]), torch.tensor([1.0, 4.0, 4.0]), seed=1)
    assert i in (1, 2)


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
