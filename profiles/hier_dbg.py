"""profiles/hier_dbg.py -- commander-step time of VecHighLevelEnv after the things bench.py does before its hier leg."""
import os, sys, time
sys.path.insert(0, os.getcwd())
import torch
from hhmarl_2d_b200 import VecLowLevelEnv, make_args
from hhmarl_2d_b200.env_hier import VecHighLevelEnv
n = 8192
dev = torch.device("cuda", 0)

def hier(label):
    henv = VecHighLevelEnv(n, device=0, seed=2, arena_base=0, autoreset=True)
    henv.reset()
    gh = torch.Generator(device=dev); gh.manual_seed(77)
    cmd = torch.randint(0, 3, (8, n, 3), device=dev, generator=gh).to(torch.int32)
    for k in range(5):
        henv.step(cmd[k])
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for k in range(5):
        henv.step(cmd[(5 + k) % 8])
    e1.record(); torch.cuda.synchronize()
    print(f"{label}: {e0.elapsed_time(e1) / 5:.3f} ms per commander step", flush=True)

hier("fresh process")
env = VecLowLevelEnv(n, make_args(level=3), device=0, seed=0, autoreset=True)
env.reset()
act = torch.zeros((n, 2, 4), dtype=torch.int32, device=dev)
for _ in range(50):
    env.step(act)
torch.cuda.synchronize()
hier("after a low-level env stepped")
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
for _ in range(20):
    flush.fill_(1); env.step(act)
torch.cuda.synchronize()
hier("after L2 flushes")
torch.cuda._sleep(20_000_000)
torch.cuda.synchronize()
hier("after a device-side spin")
import numpy as np
a_host, *outs = env.host_buffers()
for _ in range(20):
    env.step_host(a_host, out=tuple(outs))
hier("after zero-copy host steps")
