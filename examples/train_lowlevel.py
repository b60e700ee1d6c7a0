#!/usr/bin/env python
"""examples/train_lowlevel.py -- on-device PPO for the low-level policies (the role of the reference's
train_hetero.py, with RLlib replaced by hhmarl_2d_b200's sampler + learner).

    python examples/train_lowlevel.py --level 3 --arenas 4096 --epochs 20
    torchrun --nproc-per-node 8 examples/train_lowlevel.py --level 5 --arenas 8192      # arenas sharded, NCCL grads

Flags follow config.py of the reference where they exist (--level, --agent_mode, --epochs, --mini_batch_size,
--glob_frac, --rew_scale, ...).  Every `--export_every` epochs the two policies are written as
policies/L{level}_AC{i}_{mode}.pt (train_hetero.py:98-107), which levels 4/5 then load as opponents.
"""
import argparse
import os
import sys
import time

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from hhmarl_2d_b200 import PPOLearner, TorchPolicy, VecLowLevelEnv, VecSampler, make_args  # noqa: E402
from hhmarl_2d_b200 import checkpoint, models  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--level", type=int, default=1)
    ap.add_argument("--agent_mode", default="fight")
    ap.add_argument("--epochs", type=int, default=10)
    ap.add_argument("--arenas", type=int, default=2048, help="arenas per GPU")
    ap.add_argument("--fragment", type=int, default=40, help="ticks per rollout fragment (multiple of 20)")
    ap.add_argument("--mini_batch_size", type=int, default=8192)
    ap.add_argument("--num_sgd_iter", type=int, default=2)
    ap.add_argument("--glob_frac", type=float, default=0.0)
    ap.add_argument("--rew_scale", type=float, default=1.0)
    ap.add_argument("--policy_dir", default="policies")
    ap.add_argument("--export_every", type=int, default=0)
    ap.add_argument("--tf32", action="store_true")
    ap.add_argument("--restore", action="store_true",
                    help="warm-start from the exported policies of level-1 (the reference restores L{level-1}, config.py:60-75) "
                         "and, if present, resume this level's training state")
    ap.add_argument("--standin_opponents", action="store_true",
                    help="levels 4/5 without trained opponent files: seeded random-weight stand-ins (benchmarks only)")
    a = ap.parse_args()
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank, local = int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    torch.manual_seed(0)                       # identical initial weights on every rank
    m1, m2 = models.build_policy_pair(a.agent_mode)
    m1.to(dev); m2.to(dev)
    args = make_args(level=a.level, agent_mode=a.agent_mode, glob_frac=a.glob_frac, rew_scale=a.rew_scale)
    opp = None
    if a.level >= 4 and not a.standin_opponents:
        # like env_base._get_policies: missing L3 / L4 policy files are an error, not a silent fallback
        opp = checkpoint.load_opponent_policies(a.policy_dir, a.level, a.agent_mode, dev)
    if a.restore and a.level > 1:
        prev1, prev2 = checkpoint.load_pair(a.policy_dir, a.level - 1, a.agent_mode, dev)
        m1.load_state_dict(prev1.state_dict()); m2.load_state_dict(prev2.state_dict())
    env = VecLowLevelEnv(a.arenas, args, device=local, seed=0, arena_base=rank * a.arenas, opponent_policies=opp,
                         allow_standin_opponents=a.standin_opponents)
    smp = VecSampler(env, TorchPolicy(m1, 1), TorchPolicy(m2, 2), fragment_len=a.fragment,
                     use_cuda_graph=a.level <= 4, allow_tf32=a.tf32)
    learner = PPOLearner(m1, m2, num_sgd_iter=a.num_sgd_iter, sgd_minibatch_size=a.mini_batch_size)
    start = 0
    if a.restore and os.path.exists(checkpoint.training_state_path(a.policy_dir, a.level, a.agent_mode)):
        start = checkpoint.load_training_state(a.policy_dir, a.level, a.agent_mode, learner, smp)
    for ep in range(start, a.epochs):
        t0 = time.time()
        batch = smp.collect()
        torch.cuda.synchronize()
        t1 = time.time()
        st = learner.update(batch)
        smp.refresh_policy()
        torch.cuda.synchronize()
        t2 = time.time()
        rew = batch["rew"].sum(0).mean(0)
        done = float(batch["done"].float().mean())
        if rank == 0:
            print(f"epoch {ep:4d}  reward/fragment ac1 {float(rew[0]):+.3f} ac2 {float(rew[1]):+.3f}  episodes/arena-tick {done:.4f}  "
                  f"loss {st['loss']:+.4f} kl {st['kl'][0]:.4f}/{st['kl'][1]:.4f}  sample {1e-6 * smp.env_steps_per_fragment * world / (t1 - t0):.1f} M steps/s  "
                  f"learn {t2 - t1:.2f} s", flush=True)
        if a.export_every and (ep + 1) % a.export_every == 0 and rank == 0 and a.level >= 3:
            checkpoint.save_policies(a.policy_dir, a.level, a.agent_mode, m1, m2)
            checkpoint.save_training_state(a.policy_dir, a.level, a.agent_mode, learner, smp)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
