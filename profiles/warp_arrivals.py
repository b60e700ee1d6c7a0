"""profiles/warp_arrivals.py -- which warp a stage of the v4 step kernel waits for (-DHH_V4_PROFILE build, under gpurun):
    HH_LIB_PATH=$PWD/build/lib_v4prof.so python profiles/warp_arrivals.py
Lane 0 of every warp stamps clock64() when it reaches each CTA barrier; per stage the table gives, for every warp, the median
time from the stage's start (the latest arrival at the previous barrier) to its own arrival, and how often it was the last."""
import ctypes
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from hhmarl_2d_b200 import VecLowLevelEnv, make_args, _native as nat  # noqa: E402

n = 8192
env = VecLowLevelEnv(n, make_args(level=3), device=0, seed=0)
env.reset()
g = torch.Generator(device="cuda"); g.manual_seed(1)
acts = torch.stack([torch.randint(0, 13, (64, n, 2), device="cuda", generator=g), torch.randint(0, 9, (64, n, 2), device="cuda", generator=g),
                    torch.randint(0, 2, (64, n, 2), device="cuda", generator=g), torch.randint(0, 2, (64, n, 2), device="cuda", generator=g)],
                   dim=-1).to(torch.int32).contiguous()
L = nat.lib()
L.hh_debug_v4_warp_arrivals.argtypes = [ctypes.c_void_p, ctypes.c_int32]
ctas = 256
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
rows = []
for t in range(70):
    flush.fill_(t & 0xFF)
    env.step(acts[t % 64])
    if t >= 30:
        buf = np.zeros((ctas, 8, 16), np.int64)
        assert L.hh_debug_v4_warp_arrivals(buf.ctypes.data, ctas) == 0
        rows.append(buf)
a = np.concatenate(rows)                     # [samples, 8 warps, 16 barriers]
nb = 9                                       # barriers of a step without an auto-reset in the CTA
n_all = len(a)
a = a[(a[:, :, 15] == nb).all(1)]
print(f"{len(a)} of {n_all} CTA-steps without an auto-reset (9 barriers)")
names = ["S0 load", "S1 pretick|draws", "S2 actions", "S4 moves", "S5 geometry", "S6 resolve", "S7 commit", "S8 pairs", "S9 rows"]
print("cycles from the start of a stage (latest arrival at the previous barrier) to each warp's arrival; [share of CTAs where the warp was last]")
print("stage              " + "".join(f"   warp {w}    " for w in range(8)))
start = None
for k in range(nb):
    arr = a[:, :, k].astype(np.float64)
    ok = (arr > 0).all(1)
    if start is None:
        base = arr.min(1)
    else:
        base = start
    rel = arr - base[:, None]
    last = arr.argmax(1)
    line = f"{names[k]:18s} "
    for w in range(8):
        line += f"{int(np.median(rel[ok, w])):6d} [{100 * np.mean(last[ok] == w):3.0f}%] "
    print(line)
    start = arr.max(1)
