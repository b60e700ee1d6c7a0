"""CPU-side checks (no GPU): the C-ABI library loads and exports every symbol declared in
include/hhmarl_b200.h, argument validation happens before any CUDA call, and the host-side
mirror of the reference interface (spaces, args) has the reference's shapes."""
import ctypes
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    from hhmarl_2d_b200 import _native as nat
    hdr = open(os.path.join(ROOT, "include", "hhmarl_b200.h")).read()
    declared = set(re.findall(r"\b(hh_[a-z_]+)\s*\(", hdr))
    assert declared == set(nat.EXPORTS), declared ^ set(nat.EXPORTS)
    L = nat.lib()
    for name in declared:
        assert hasattr(L, name), name
    assert b"sm_100a" in L.hh_version()


def test_create_validates_arguments_without_a_gpu():
    from hhmarl_2d_b200 import _native as nat
    L = nat.lib()
    h = nat.VP()
    for bad in (dict(level=0), dict(level=9), dict(level=1, agent_mode=2), dict(level=1, map_size=0.0, horizon=10),
                dict(level=1, map_size=0.3, horizon=0)):
        cfg = nat.HHConfig(**{"map_size": 0.3, "horizon": 150, **bad})
        assert L.hh_create(ctypes.byref(cfg), 8, 0, ctypes.byref(h)) == -1
        assert L.hh_last_error()
    cfg = nat.HHConfig(level=1, map_size=0.3, horizon=150)
    assert L.hh_create(ctypes.byref(cfg), 0, 0, ctypes.byref(h)) == -1
    assert L.hh_step(None, None, None, None, None, None, None) == -1
    assert L.hh_step_range(None, 0, 32, None, None, None, None, None, None) == -1
    assert L.hh_n_arenas(None) == 0 and L.hh_obs_dim(None, 1) == 0


def test_config_struct_layout_matches_header():
    from hhmarl_2d_b200 import _native as nat
    assert ctypes.sizeof(nat.HHConfig) == 8 * 4 + 3 * 8 + 2 * 8
    assert ctypes.sizeof(nat.HHStateView) == 8 * len(nat.STATE_FIELDS) == 8 * 34


def test_args_and_spaces_mirror_reference():
    from hhmarl_2d_b200 import make_args, HORIZON_BY_LEVEL
    from hhmarl_2d_b200.spaces import Box, MultiDiscrete
    assert HORIZON_BY_LEVEL == {1: 150, 2: 200, 3: 300, 4: 350, 5: 400}      # config.py:95
    a = make_args(level=3)
    assert (a.horizon, a.total_num, a.map_size, a.agent_mode) == (300, 4, 0.3, "fight")
    assert a.env_config["args"] is a
    b = Box(np.zeros(26), np.ones(26), dtype=np.float32)
    assert b.shape == (26,) and b.contains(np.full(26, 0.5, np.float32))
    m = MultiDiscrete([13, 9, 2, 2])
    assert m.contains(np.array([12, 8, 1, 1])) and not m.contains(np.array([13, 0, 0, 0]))


def test_missing_library_fails_loudly(monkeypatch):
    from hhmarl_2d_b200 import _native as nat
    monkeypatch.setattr(nat, "_lib", None)
    monkeypatch.setattr(nat, "LIB_PATH", "/nonexistent/libhhmarl_b200.so")
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        nat.lib()


def test_policy_forward_validates_arguments_without_a_gpu():
    """hh_policy_forward / _ex reject inconsistent chain descriptions before any CUDA call; struct layouts match."""
    from hhmarl_2d_b200 import _native as nat
    L = nat.lib()
    assert ctypes.sizeof(nat.HHPolicyChain) == 8 * 8 + 8 * 4
    assert ctypes.sizeof(nat.HHPolicyChainEx) == 13 * 8 + 16 * 4 + 8 * 8  # 15 int32 + padding, then the precision-2 images
    one = (nat.HHPolicyChainEx * 1)()
    assert L.hh_policy_forward_ex(0, one, 0, None) == -1            # no chains
    assert L.hh_policy_forward_ex(9, one, 0, None) == -1            # more than 8
    assert L.hh_policy_forward_ex(1, one, 3, None) == -1            # unknown precision
    assert L.hh_policy_forward_ex(1, one, 0, None) == -1            # null pointers
    assert b"chain" in L.hh_policy_last_error()
    assert L.hh_policy_forward_ex(1, one, 2, None) == -1            # tcgen05 path: same checks before any CUDA call
    assert L.hh_policy_image_bytes(32, 512) == 32 * 512 * 64        # one K = 16 step of n columns: hi + lo halves
    assert L.hh_policy_pack(None, 504, 512, 512, 512, 256, 0, 32, 1, None, None, None) == -1
    assert L.hh_policy_tc_pair() in (0, 1)
    four = (nat.HHPolicyChain * 4)()
    assert L.hh_policy_forward(0, four, None, None, 0, None) == -1
    assert L.hh_step_host_begin(None, None) == -1 and L.hh_step_host_end(None, None, None, None, None) == -1
    assert L.hh_set_host_mode(None, 0) == -1
    assert L.hh_policy_rows_by_key(0, None, 3, None, None, None, None) == -1
