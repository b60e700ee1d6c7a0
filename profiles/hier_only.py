"""profiles/hier_only.py -- ms per commander step of VecHighLevelEnv at 8 192 arenas: eager launches against the CUDA-graph replay, one handle against groups of arenas on their own streams."""
import os
import sys

sys.path.insert(0, os.getcwd())
import torch  # noqa: E402
from hhmarl_2d_b200.env_hier import VecHighLevelEnv  # noqa: E402

n = 8192
for use_graph, groups in ((False, 1), (True, 1), (True, 2), (True, 4), (True, 8)):
    henv = VecHighLevelEnv(n, device=0, seed=2, autoreset=True, groups=groups)
    henv.use_cuda_graph = use_graph
    henv.reset()
    g = torch.Generator(device="cuda"); g.manual_seed(77)
    cmd = torch.randint(0, 3, (10, n, 3), device="cuda", generator=g).to(torch.int32)
    for w in range(4):
        henv.step(cmd[w])
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for k in range(5):
        henv.step(cmd[4 + k])
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 5
    print(f"use_cuda_graph={use_graph} groups={groups}: {ms:.3f} ms per commander step -> {n / ms / 1e3:.2f} M commander-steps/s")
