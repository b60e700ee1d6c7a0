"""profiles/policy_forward_probe.py -- times csrc/hh_policy.cu alone (CUDA events) against the packed cuBLAS forward;
run under ncu for the kernel's counters:  ncu --set full -k regex:policy_forward -c 2 python profiles/policy_forward_probe.py"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from hhmarl_2d_b200 import models as M  # noqa: E402
from hhmarl_2d_b200.fused_forward import FusedPolicyPair, PackedPolicyPair  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
torch.manual_seed(0)
m1, m2 = M.build_policy_pair("fight")
m1.cuda(); m2.cuda()
f1 = torch.rand(B, 57, device="cuda"); f2 = torch.rand(B, 57, device="cuda")


def timeit(fn, n=30):
    for _ in range(5):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3


flops = 2 * B * 2 * (57 * 1000 + 100 * 100 + 150 * 150 + 2 * 500 * 500 + 500 * 27)
for prec, name in ((2, "tcgen05 H3"), (0, "fused 3xTF32"), (1, "fused TF32")):
    fu = FusedPolicyPair(m1, m2, precision=prec)
    us = timeit(lambda: fu.forward(f1, f2))
    if prec == 2:
        with torch.no_grad():
            ref = m1.forward_flat(f1)
        out = fu.forward(f1, f2)
        print("tcgen05 H3 max |logits - torch fp32|", (out[0] - ref[0]).abs().max().item(), "value", (out[1] - ref[1]).abs().max().item())
    print(f"{name:14s} {us:8.1f} us  ({flops / us / 1e6:6.1f} TFLOP/s useful)")
pk = PackedPolicyPair(m1, m2)
for tf32 in (False, True):
    torch.backends.cuda.matmul.allow_tf32 = tf32
    us = timeit(lambda: pk.forward(f1, f2))
    print(f"{'cuBLAS ' + ('TF32' if tf32 else 'fp32'):14s} {us:8.1f} us  ({flops / us / 1e6:6.1f} TFLOP/s useful)")
