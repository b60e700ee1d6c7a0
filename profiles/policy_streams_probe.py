"""profiles/policy_streams_probe.py -- how well do policy-forward launches of parts of the batch pack onto the SMs when they come
from several streams of one CUDA graph (no env step, no glue)?  8 192 rows; g parts on g streams, 20 launches per stream."""
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import torch  # noqa: E402
from hhmarl_2d_b200 import models as M  # noqa: E402
from hhmarl_2d_b200.fused_forward import FusedPolicyPair  # noqa: E402

B, T = 8192, 20
torch.manual_seed(0)
m1, m2 = M.build_policy_pair("fight")
m1.cuda(); m2.cuda()
fu = FusedPolicyPair(m1, m2, precision=2)
f1 = torch.rand(B, 57, device="cuda"); f2 = torch.rand(B, 57, device="cuda")
out = (torch.empty((B, 26), device="cuda"), torch.empty((B,), device="cuda"), torch.empty((B, 24), device="cuda"), torch.empty((B,), device="cuda"))
for g in (1, 2, 4, 8, 16):
    n = B // g
    streams = [torch.cuda.Stream() for _ in range(g)]

    def body():
        cur = torch.cuda.current_stream()
        for k, s in enumerate(streams):
            lo, hi = k * n, (k + 1) * n
            s.wait_stream(cur)
            with torch.cuda.stream(s):
                for _ in range(T):
                    fu.forward(f1[lo:hi], f2[lo:hi], out=tuple(o[lo:hi] for o in out))
        for s in streams:
            cur.wait_stream(s)

    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        body()
    torch.cuda.current_stream().wait_stream(side)
    gr = torch.cuda.CUDAGraph()
    with torch.cuda.graph(gr):
        body()
    for _ in range(2):
        gr.replay()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        gr.replay()
    e1.record()
    torch.cuda.synchronize()
    print(f"parts {g:2d} ({n // 128 * 4:3d} CTAs per launch): {e0.elapsed_time(e1) * 1e3 / (10 * T):7.1f} us per 8 192 rows")
