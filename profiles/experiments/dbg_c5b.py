import sys, numpy as np, torch
sys.path.insert(0, ".")
from islamcts_b200.engine import Engine
from islamcts_b200 import _lib
from oracle import coracle
print("lib", _lib.LIB_PATH)
n_sim = 1000
rng = np.random.default_rng(2024)
roots = (rng.random((65536, 4)) - 0.5) * 0.1
i = 30162
c1 = coracle.make_config(0, n_actions=2, max_depth=500, C_=1.0, gamma=0.99, seed=21, tree_id=1000 + i)
for kernel in (1, 2):
    for ns in (949, 950):
        e2 = Engine(0, n_trees=1, sim_capacity=n_sim, max_depth=500, C_=1.0, gamma=0.99, precision=1, kernel=kernel, seed=21, tree_id_base=1000 + i)
        e2.set_roots(roots[i:i+1]).run(ns)
        t2 = coracle.Tree(c1, roots[i]); t2.run(ns)
        sd, so = e2.export_tree(0), t2.serialise()
        dt = np.abs(sd["total"] - so["total"]); idx = np.argwhere(dt > 1e-9).reshape(-1)
        print("kernel", kernel, "nsim", ns, "records", len(dt), "differing", len(idx), "counters dev", e2.counters(), "oracle", t2.counters())
        if len(idx):
            k0 = idx[-1]
            for k in range(max(0, k0 - 3), min(len(dt), k0 + 6)):
                print("   rec", k, "kind", so["kind"][k], sd["kind"][k], "depth", so["depth"][k], "count", so["count"][k], sd["count"][k], "term", so["terminal"][k], sd["terminal"][k], "total", so["total"][k], sd["total"][k])
