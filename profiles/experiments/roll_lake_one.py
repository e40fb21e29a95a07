import sys, torch
sys.path.insert(0, ".")
from islamcts_b200 import engine as E
dev = torch.device("cuda:0")
for _ in range(2):
    ret, st = E.rollout(3, None, n=1 << 24, budget=1000, gamma=1.0, seed=1234, sim=0, device=dev)
torch.cuda.synchronize()
print(int(st.sum().item()))
