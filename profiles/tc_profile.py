"""profiles/tc_profile.py -- per-role clock64() stamps of the tcgen05 policy forward (csrc/hh_policy_tc.cu): where a CTA's time
goes (MMA issue per segment, waits for the weight stream, the epilogue phases).  python profiles/tc_profile.py [rows] [cluster]"""
import ctypes
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from hhmarl_2d_b200 import _native as nat  # noqa: E402
from hhmarl_2d_b200 import models as M  # noqa: E402
from hhmarl_2d_b200.fused_forward import FusedPolicyPair  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
L = nat.lib()
L.hh_policy_tc_profile.argtypes = [ctypes.c_void_p]
cluster = 2 if L.hh_policy_tc_pair() else 1
L.hh_policy_tc_mode.restype = ctypes.c_int32
TM = 128 if L.hh_policy_tc_mode() == 2 else 64
torch.manual_seed(0)
m1, m2 = M.build_policy_pair("fight")
m1.cuda(); m2.cuda()
f1 = torch.rand(B, 57, device="cuda"); f2 = torch.rand(B, 57, device="cuda")
fu = FusedPolicyPair(m1, m2, precision=2)
for _ in range(3):
    fu.forward(f1, f2)
tiles = (B + TM - 1) // TM
tiles = (tiles + cluster - 1) // cluster * cluster
names = {1: "seg0 L1h0 start", 2: "seg0 end", 3: "seg1 L1h1 start", 4: "seg1 end", 5: "seg2 ATT start", 6: "seg2 end",
         7: "seg3 SHh0 start", 8: "seg3 end", 9: "seg4 SHh1 start", 10: "seg4 end", 11: "seg5 HEADa start", 12: "seg5 end",
         13: "seg6 HEADb start", 14: "seg6 end", 16: "epi L1h0 begin", 17: "epi L1h0 end", 18: "epi L1h1 begin", 19: "epi L1h1 end",
         20: "epi ATT begin", 21: "epi ATT end", 22: "epi SH begin", 23: "epi SHh0 end", 24: "epi SHh1 end", 25: "epi HEAD begin",
         26: "epi done", 29: "producer done"}
if TM == 128:    # the 128-row form: layer 1's upper half first, the shared layer's first half split around the attention block
    names.update({1: "seg0 L1 hi-half start", 3: "seg1 L1 lo-half start", 7: "seg3 SHh0a start", 9: "seg4 SHh0b start",
                  11: "seg5 SHh1 start", 13: "seg6 HEADa start", 16: "epi L1 hi begin", 17: "epi L1 hi end", 18: "epi L1 lo begin",
                  19: "epi L1 lo end", 14: "CTA exit (TMEM freed)"})
L.hh_policy_tc_debug.argtypes = [ctypes.c_int32]
modes = [(0, "normal")] + ([(1, "DEBUG no weight copies (timing only)"), (2, "DEBUG no MMAs (timing only)")] if len(sys.argv) > 2 else [])
for flags, label in modes:
    L.hh_policy_tc_debug(flags)
    buf = torch.zeros(4 * tiles * 32, dtype=torch.int64, device="cuda")
    L.hh_policy_tc_profile(buf.data_ptr())
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    fu.forward(f1, f2)
    e1.record()
    torch.cuda.synchronize()
    L.hh_policy_tc_profile(None)
    L.hh_policy_tc_debug(0)
    s = buf.view(4, tiles, 32).cpu().double()
    print(f"== {label}: rows {B}, {'CTA pairs (cta_group::2)' if cluster == 2 else f'one CTA per {TM}-row tile'}: launch {e0.elapsed_time(e1) * 1e3:.1f} us "
          f"(with stamps); cycles relative to the start stamp of warp 1 of the same CTA, median over the CTAs of [actor chain 0 | critic chain 1]")
    for rank in range(cluster):
        sr = s[:, rank::cluster, :]
        rel = sr - sr[:, :, 0:1]
        if cluster == 2:
            print(f" CTA rank {rank} of the pair ({'MMA issue' if rank == 0 else 'relay: segment stamps = its waits for its own stages'})")
        for k in sorted(names):
            if flags and not (7 <= k <= 10):
                continue
            a, c = rel[0, :, k].median().item(), rel[1, :, k].median().item()
            print(f"  {names[k]:22s} {a:9.0f} {c:9.0f}")
        print(f"  warp 1 waiting for weight stages (sum):     {sr[0, :, 15].median().item():9.0f} {sr[1, :, 15].median().item():9.0f}")
        print(f"  producer waiting for free slots (sum):      {sr[0, :, 28].median().item():9.0f} {sr[1, :, 28].median().item():9.0f}")
        if TM == 128:
            # how the rounds of tiles follow each other on an SM: entry / exit clocks per SM id (prof 27 / 14 / 30), wall clock (31)
            allc = s.view(-1, 32)
            smid, entry, start, done, exit_, gt = allc[:, 30], allc[:, 27], allc[:, 0], allc[:, 26], allc[:, 14], allc[:, 31]
            print(f"  CTA entry -> set-up done (barriers, TMEM alloc, row map): {(start - entry).median().item():9.0f} cycles;"
                  f"  output done -> exit: {(exit_ - done).median().item():9.0f}")
            gaps, firsts, seconds = [], [], []
            for sm in smid.unique().tolist():
                idx = (smid == sm).nonzero().flatten()
                if idx.numel() == 2:
                    a, b = (idx[0], idx[1]) if entry[idx[0]] < entry[idx[1]] else (idx[1], idx[0])
                    gaps.append((entry[b] - exit_[a]).item())
                    firsts.append((exit_[a] - entry[a]).item())
                    seconds.append((exit_[b] - entry[b]).item())
            import statistics as st_
            if gaps:
                print(f"  SMs with two tiles: {len(gaps)}; first tile entry -> exit {st_.median(firsts):9.0f}, second {st_.median(seconds):9.0f}, "
                      f"exit of the first -> entry of the second {st_.median(gaps):9.0f} cycles")
            print(f"  wall clock, first CTA entry -> last CTA entry: {(gt.max() - gt.min()).item() / 1e3:9.1f} us")
        else:
            print(f"  warp 1 cycles in MMA issue / in commits:    {sr[0, :, 30].median().item():9.0f} {sr[0, :, 31].median().item():9.0f}")
