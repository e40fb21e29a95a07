import sys; sys.path.insert(0,'.')
import numpy as np
from oracle import coracle
from islamcts_b200.engine import Engine
kw=dict(env_id=2, agent=3, max_depth=50, C=1.0, gamma=0.99, alpha1=0.5, k1=1.0, alpha2=0.5, k2=1.0)
root=coracle.env_reset(2,5,4)
print('root',root, root.astype(np.float32))
for n_sim in list(range(40,161,4)):
    eng=Engine(2, agent=3, n_trees=1, sim_capacity=200, max_depth=50, C_=1.0, gamma=0.99, alpha1=0.5,k1=1.0,alpha2=0.5,k2=1.0, precision=1, kernel=2, seed=2024, tree_id_base=74)
    eng.set_roots(root[None,:]).run(n_sim)
    t=coracle.Tree(coracle.make_config(2, agent=3, max_depth=50, C_=1.0, gamma=0.99, alpha1=0.5,k1=1.0,alpha2=0.5,k2=1.0, seed=2024, tree_id=74), root)
    t.run(n_sim)
    g=eng.export_tree(0); e=t.serialise()
    same = len(g['count'])==len(e['count']) and (g['count']==e['count']).all() and np.allclose(g['total'],e['total'],rtol=1e-12) and np.allclose(g['action'],e['action'])
    if not same:
        print('first diff at n_sim',n_sim)
        for i in range(max(len(g['count']),len(e['count']))):
            a=[g[k][i] if i<len(g['count']) else None for k in ('kind','depth','count','total','action')]
            b=[e[k][i] if i<len(e['count']) else None for k in ('kind','depth','count','total','action')]
            print(i,a,b, "" if a==b else "<<<")
        print(eng.tree_meta(0), t.counters())
        break
else:
    print('no diff up to 40')
