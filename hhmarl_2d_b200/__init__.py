"""hhmarl_2d_b200 -- B200-native (sm_100a) implementation of the hot path of IDSIA/hhmarl_2D:
the 2-D air-combat low-level environment step (warsim/simulator + envs/env_hetero.py) for
thousands of arenas in lock-step, behind the reference's own Env API.

  VecLowLevelEnv  batched env on torch CUDA tensors (or host numpy buffers)
  LowLevelEnv     drop-in for envs/env_hetero.py::LowLevelEnv (dict API, one arena)
"""
from .config import make_args, HORIZON_BY_LEVEL  # noqa: F401
from .vec_env import VecLowLevelEnv  # noqa: F401
from .env_hetero import LowLevelEnv  # noqa: F401


def __getattr__(name):   # torch-dependent pieces are imported lazily
    if name in ("VecSampler", "TorchPolicy"):
        from . import sampler
        return getattr(sampler, name)
    if name == "PPOLearner":
        from .ppo import PPOLearner
        return PPOLearner
    if name in ("EvalStats", "evaluate"):
        from . import evaluation
        return getattr(evaluation, name)
    if name in ("TraceRecorder", "HierTraceRecorder"):
        from . import trace
        return getattr(trace, name)
    if name == "OpponentPolicies":
        from .opponents import OpponentPolicies
        return OpponentPolicies
    raise AttributeError(name)


__all__ = ["VecLowLevelEnv", "LowLevelEnv", "make_args", "HORIZON_BY_LEVEL", "VecSampler", "TorchPolicy", "PPOLearner",
           "OpponentPolicies", "EvalStats", "evaluate", "TraceRecorder", "HierTraceRecorder"]
