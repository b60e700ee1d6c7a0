"""profiles/ppo_learner_probe.py -- PPOLearner.update on a level-5 fragment batch of 8 192 arenas x 20 ticks (the bench's ppo leg):
eager minibatches against CUDA-graph replays.  python profiles/ppo_learner_probe.py [arenas]"""
import os
import sys
import time

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import torch  # noqa: E402
from hhmarl_2d_b200 import VecLowLevelEnv, make_args, VecSampler, TorchPolicy, PPOLearner  # noqa: E402
from hhmarl_2d_b200 import models as M  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
for use_graph in (False, True):
    torch.manual_seed(0)
    m1, m2 = M.build_policy_pair("fight")
    m1.cuda(); m2.cuda()
    env = VecLowLevelEnv(n, make_args(level=5), device=0, seed=5, autoreset=True, allow_standin_opponents=True)
    smp = VecSampler(env, TorchPolicy(m1, 1), TorchPolicy(m2, 2), fragment_len=20, use_cuda_graph=True)
    learner = PPOLearner(m1, m2, num_sgd_iter=1, sgd_minibatch_size=8192, use_cuda_graph=use_graph)
    for _ in range(2):
        learner.update(smp.collect())
        smp.refresh_policy()
    b = smp.collect()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(3):
        st = learner.update(b)
    torch.cuda.synchronize()
    t = (time.perf_counter() - t0) / 3
    print(f"use_cuda_graph={use_graph}: {t * 1e3:8.2f} ms per update, {st['minibatches']} minibatches -> {t * 1e3 / st['minibatches']:.3f} ms per minibatch")
    del learner, smp, env
