"""profiles/stage_clocks.py -- per-stage latency of the v4 step kernel from a -DHH_V4_PROFILE build (run under gpurun):
    HH_LIB_PATH=$PWD/build/lib_v4prof.so python profiles/stage_clocks.py [arenas]
Thread 0 of every CTA records clock64() at each stage boundary; the table is the median / max over CTAs and steps."""
import ctypes
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from hhmarl_2d_b200 import VecLowLevelEnv, make_args, _native as nat  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
apc = int(os.environ.get("HH_V4_ARENAS", "32"))
env = VecLowLevelEnv(n, make_args(level=3), device=0, seed=0)
env.reset()
g = torch.Generator(device="cuda"); g.manual_seed(1)
acts = torch.stack([torch.randint(0, 13, (64, n, 2), device="cuda", generator=g), torch.randint(0, 9, (64, n, 2), device="cuda", generator=g),
                    torch.randint(0, 2, (64, n, 2), device="cuda", generator=g), torch.randint(0, 2, (64, n, 2), device="cuda", generator=g)],
                   dim=-1).to(torch.int32).contiguous()
L = nat.lib()
L.hh_debug_v4_profile.argtypes = [ctypes.c_void_p, ctypes.c_int32]
ctas = min(4096, (n + apc - 1) // apc)
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
rows = []
for t in range(60):
    flush.fill_(t & 0xFF)
    env.step(acts[t % 64])
    if t >= 20:
        buf = np.zeros((ctas, 16), np.int64)
        assert L.hh_debug_v4_profile(buf.ctypes.data, ctas) == 0
        rows.append(np.diff(buf[:, :11], axis=1))
d = np.concatenate(rows)            # [steps * ctas, 10]
names = ["S0 load", "S1 pretick|draws", "S2 actions", "S4 moves", "S5 geometry", "S6 resolve", "S7 commit/feat", "S8 pairs",
         "S9 rows", "S10 store"]
print(f"{n} arenas, {ctas} CTAs of {apc} arenas; cycles per stage (thread 0 of each CTA, incl. the barrier wait)")
for k, nm in enumerate(names):
    print(f"  {nm:18s} median {int(np.median(d[:, k])):6d}  p90 {int(np.percentile(d[:, k], 90)):6d}  max {int(d[:, k].max()):7d}")
tot = d.sum(1)
print(f"  {'total':18s} median {int(np.median(tot)):6d}  p90 {int(np.percentile(tot, 90)):6d}  max {int(tot.max()):7d}   "
      f"({np.median(tot) / 1.965e3:.1f} us at 1965 MHz)")
