#!/bin/bash
# profiles/sampler_groups.sh -- the sampler's groups (halves / quarters of the batch on their own streams inside the fragment graph):
# ms per tick at 8 192 arenas, level 3
for co in ""; do
for g in 1 2 4 8; do HH_STEP_CARVEOUT=$co HH_SAMPLER_GROUPS=$g timeout 300 python - <<PY
import torch, os, sys
sys.path.insert(0, ".")
from hhmarl_2d_b200 import VecLowLevelEnv, make_args, VecSampler, TorchPolicy
from hhmarl_2d_b200 import models as M
torch.manual_seed(0)
m1, m2 = M.build_policy_pair("fight"); m1.cuda(); m2.cuda()
env = VecLowLevelEnv(8192, make_args(level=3), device=0, seed=1, autoreset=True)
smp = VecSampler(env, TorchPolicy(m1, 1), TorchPolicy(m2, 2), fragment_len=20, use_cuda_graph=True)
for _ in range(3): smp.collect()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10): smp.collect()
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 200
print("carveout", os.environ.get("HH_STEP_CARVEOUT") or "default", "groups", smp.groups, "ms per tick %.4f" % ms, "-> %.1f M env-steps/s" % (8192 / ms / 1e3))
PY
done
done
