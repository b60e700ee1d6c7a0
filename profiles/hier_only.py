import torch, sys, os
sys.path.insert(0, os.getcwd())
from hhmarl_2d_b200.env_hier import VecHighLevelEnv
n=8192
henv = VecHighLevelEnv(n, device=0, seed=2, autoreset=True)
henv.reset()
g = torch.Generator(device="cuda"); g.manual_seed(77)
cmd = torch.randint(0, 3, (8, n, 3), device="cuda", generator=g).to(torch.int32)
for w in range(2): henv.step(cmd[w])
torch.cuda.synchronize()
e0,e1=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
e0.record()
for k in range(3): henv.step(cmd[2+k])
e1.record(); torch.cuda.synchronize()
print("ms per commander step", e0.elapsed_time(e1)/3)
