import sys; sys.path.insert(0,'.')
import numpy as np
from oracle import coracle
from islamcts_b200.engine import Engine
n_trees=24
roots=np.stack([coracle.env_reset(2,5,i) for i in range(n_trees)])
def run(n_sim, nt=n_trees, base=70, off=0):
    eng=Engine(2, agent=3, n_trees=nt, sim_capacity=160, max_depth=50, C_=1.0, gamma=0.99, alpha1=0.5,k1=1.0,alpha2=0.5,k2=1.0, precision=1, kernel=2, seed=2024, tree_id_base=base)
    eng.set_roots(roots[off:off+nt]).run(n_sim)
    bad=[]
    for i in range(nt):
        t=coracle.Tree(coracle.make_config(2, agent=3, max_depth=50, C_=1.0, gamma=0.99, alpha1=0.5,k1=1.0,alpha2=0.5,k2=1.0, seed=2024, tree_id=base+i), roots[off+i])
        t.run(n_sim)
        g=eng.export_tree(i); e=t.serialise()
        same = len(g['count'])==len(e['count']) and (g['count']==e['count']).all() and np.allclose(g['total'],e['total'],rtol=1e-12) and np.allclose(g['action'],e['action'])
        if not same: bad.append(i)
    return bad, eng
for n_sim in (20,40,80,120,160):
    bad,eng=run(n_sim)
    print('n_sim',n_sim,'bad trees',bad, eng.counters())
bad,eng=run(160, nt=1, base=74, off=4); print('single tree 74, cap160:', bad)
bad,eng=run(160, nt=5, base=70, off=0); print('5 trees:', bad)
