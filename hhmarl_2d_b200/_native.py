"""ctypes binding of libhhmarl_b200.so (C ABI: include/hhmarl_b200.h).

The library is the product: there is NO Python/CPU fallback.  If the shared object is missing
or cannot be loaded this module raises -- it never silently routes around the CUDA path.
"""
from __future__ import annotations

import ctypes
import os
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(_HERE, "csrc")
LIB_PATH = os.environ.get("HH_LIB_PATH") or os.path.join(CSRC, "libhhmarl_b200.so")
SOURCES = ["hh_api.cu", "hh_hier.cu", "hh_policy.cu", "hh_policy_tc.cu"]
HEADERS = ["hh_policy_tc.h", "hh_quad.cuh", "hh_cta.cuh", "hh_v4.cuh", "hh_state_pack.h", "hh_core.cuh", "hh_geodesic.cuh", os.path.join("..", "..", "include", "hhmarl_b200.h")]

NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-std=c++17", "-lineinfo",
              "--expt-relaxed-constexpr", "-Xcompiler", "-fPIC", "-shared"]

I32 = ctypes.c_int32
U64 = ctypes.c_uint64
D = ctypes.c_double
VP = ctypes.c_void_p


class HHConfig(ctypes.Structure):
    _fields_ = [("level", I32), ("agent_mode", I32), ("horizon", I32), ("esc_dist_rew", I32),
                ("friendly_kill", I32), ("friendly_punish", I32), ("autoreset", I32), ("reserved", I32),
                ("map_size", D), ("rew_scale", D), ("glob_frac", D), ("seed", U64), ("arena_base", U64)]


_F64_4 = ("lat", "lon", "heading", "speed", "new_heading", "new_speed")
_I32_4 = ("cannon_remain", "cannon_burst", "cannon_max", "missile_remain", "rocket_max", "missile_wait",
          "alive", "has_missile", "opp_to_attack")
_F64_2 = ("r_lat", "r_lon", "r_heading", "r_new_heading")
_I32_2 = ("r_alive", "r_age", "r_target", "r_id")
_I32_1 = ("steps", "alive_agents", "alive_opps", "escaping", "escaping_time", "next_unit_id", "policy_set",
          "opp_mode", "error")
_U64_1 = ("draws_g", "draws_c")
STATE_FIELDS = ([(n, "f8", 4) for n in _F64_4] + [(n, "i4", 4) for n in _I32_4] + [(n, "f8", 2) for n in _F64_2]
                + [(n, "i4", 2) for n in _I32_2] + [(n, "i4", 1) for n in _I32_1] + [(n, "u8", 1) for n in _U64_1])


class HHHierConfig(ctypes.Structure):
    _fields_ = [("horizon", I32), ("level", I32), ("friendly_kill", I32), ("hier_action_assess", I32),
                ("hier_opp_fight_ratio", I32), ("autoreset", I32), ("map_size", D), ("rew_scale", D), ("glob_frac", D),
                ("seed", U64), ("arena_base", U64)]


class HHHierArena(ctypes.Structure):
    _fields_ = ([(n, D * 6) for n in ("lat", "lon", "hdg", "spd", "nhdg", "nspd", "rlat", "rlon", "rhdg", "rnhdg")]
                + [("ota_dn", (D * 3) * 6), ("rewards", D * 3), ("dg", U64)]
                + [(n, I32 * 6) for n in ("crem", "burst", "mrem", "mwait", "rid")]
                + [(n, I32) for n in ("steps", "alive_ag", "alive_op", "next_id", "sub", "kill_event",
                                      "situation_event", "active", "err")]
                + [("dc", ctypes.c_uint32), ("ca", ctypes.c_int8 * 6)]
                + [(n, ctypes.c_uint8 * 6) for n in ("alive", "hasm", "actype", "ralive", "rage", "rtgt", "ota_n")]
                + [("ota_id", (ctypes.c_uint8 * 3) * 6)])


class HHPolicyChain(ctypes.Structure):
    _fields_ = ([(n, VP) for n in ("x", "w1", "b1", "watt", "batt", "wh", "bh", "out")]
                + [(n, I32) for n in ("ldx", "d_in", "k1_pad", "att_lo", "att_n", "att_pad", "n_out", "ld_out")])


class HHPolicyChainEx(ctypes.Structure):
    _fields_ = ([(n, VP) for n in ("x", "w1", "b1", "watt", "batt", "ws", "bs", "wh", "bh", "out", "rows", "range_dev", "act_out")]
                + [(n, I32) for n in ("n_rows", "ldx", "d_in", "k1_pad", "att_lo", "att_n", "att_pad", "n_out", "ld_out", "n_heads")]
                + [("head", I32 * 4), ("ld_act", I32)]
                + [(n, VP) for n in ("img_w1", "img_att", "img_ws", "img_wh", "us_w1", "us_att", "us_ws", "us_wh")])


class HHStateView(ctypes.Structure):
    _fields_ = [(n, VP) for n, _, _ in STATE_FIELDS]


def needs_build() -> bool:
    if not os.path.exists(LIB_PATH):
        return True
    t = os.path.getmtime(LIB_PATH)
    return any(os.path.getmtime(os.path.join(CSRC, f)) > t for f in SOURCES + HEADERS)


def build(force: bool = False, verbose: bool = False) -> str:
    """Compile the CUDA library for sm_100a with nvcc (cross-compiles without a GPU)."""
    if force or needs_build():
        cmd = ["nvcc"] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-o", LIB_PATH] + SOURCES
        subprocess.run(cmd, cwd=CSRC, check=True)
    return LIB_PATH


_lib = None


def lib() -> ctypes.CDLL:
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(nvcc, sm_100a). hhmarl_2d_b200 has no CPU fallback.")
    _lib = bind(ctypes.CDLL(LIB_PATH))
    return _lib


class _Partial:
    """Proxy used by bind(partial=True): prototypes of entry points a library does not export are dropped."""

    def __init__(self, L):
        object.__setattr__(self, "_L", L)

    def __getattr__(self, name):
        try:
            return getattr(self._L, name)
        except AttributeError:
            return _Missing()


class _Missing:
    argtypes = restype = None


def bind(L: ctypes.CDLL, partial: bool = False) -> ctypes.CDLL:
    """Attach the prototypes of include/hhmarl_b200.h to a loaded library.  `partial` is for the test harness
    (tests/emu), which implements only the host subset of the ABI; the product library must export everything."""
    real = L
    if partial:
        L = _Partial(L)
    P = ctypes.POINTER
    L.hh_create.argtypes = [P(HHConfig), I32, I32, P(VP)]
    L.hh_create.restype = ctypes.c_int
    L.hh_destroy.argtypes = [VP]
    L.hh_destroy.restype = None
    L.hh_n_arenas.argtypes = [VP]
    L.hh_n_arenas.restype = I32
    L.hh_obs_dim.argtypes = [VP, I32]
    L.hh_obs_dim.restype = I32
    L.hh_reset.argtypes = [VP, VP, VP, VP, VP]
    L.hh_reset.restype = ctypes.c_int
    L.hh_step.argtypes = [VP, VP, VP, VP, VP, VP, VP]
    L.hh_step.restype = ctypes.c_int
    L.hh_step_range.argtypes = [VP, ctypes.c_int32, ctypes.c_int32, VP, VP, VP, VP, VP, VP]
    L.hh_step_range.restype = ctypes.c_int
    L.hh_step_range_central.argtypes = [VP, ctypes.c_int32, ctypes.c_int32, VP, VP, VP, VP, VP, VP, VP, ctypes.c_int32, VP]
    L.hh_step_range_central.restype = ctypes.c_int
    L.hh_multicat_forward.argtypes = [I32, I32, VP, VP, I32, VP, I32, VP, I32, VP, VP, VP, VP]
    L.hh_multicat_forward.restype = ctypes.c_int
    L.hh_multicat_backward.argtypes = [I32, I32, VP, VP, I32, VP, I32, VP, I32, VP, VP, VP, VP, VP]
    L.hh_multicat_backward.restype = ctypes.c_int
    F32 = ctypes.c_float
    L.hh_ppo_loss.argtypes = [I32, VP, VP, VP, VP, VP, VP, VP, VP, F32, F32, F32, F32, VP, VP, VP]
    L.hh_ppo_loss.restype = ctypes.c_int
    L.hh_fragment_prepare.argtypes = [ctypes.c_int32, ctypes.c_int32, ctypes.c_int32, VP, VP, VP, VP, VP]
    L.hh_fragment_prepare.restype = ctypes.c_int
    L.hh_fragment_writeback.argtypes = [ctypes.c_int32, ctypes.c_int32, ctypes.c_int32, VP, VP, VP, VP]
    L.hh_fragment_writeback.restype = ctypes.c_int
    L.hh_step_begin.argtypes = [VP, VP, VP, VP, VP, VP]
    L.hh_step_begin.restype = ctypes.c_int
    L.hh_step_finish.argtypes = [VP, VP, VP, VP, VP, VP, VP]
    L.hh_step_finish.restype = ctypes.c_int
    L.hh_reset_host.argtypes = [VP, VP, VP, VP]
    L.hh_reset_host.restype = ctypes.c_int
    L.hh_step_host.argtypes = [VP, VP, VP, VP, VP, VP]
    L.hh_step_host.restype = ctypes.c_int
    L.hh_host_buffers.argtypes = [VP] + [P(VP)] * 5
    L.hh_host_buffers.restype = ctypes.c_int
    L.hh_step_host_begin.argtypes = [VP, VP]
    L.hh_step_host_begin.restype = ctypes.c_int
    L.hh_step_host_end.argtypes = [VP, VP, VP, VP, VP]
    L.hh_step_host_end.restype = ctypes.c_int
    L.hh_set_host_mode.argtypes = [VP, I32]
    L.hh_set_host_mode.restype = ctypes.c_int
    L.hh_get_state.argtypes = [VP, P(HHStateView)]
    L.hh_get_state.restype = ctypes.c_int
    L.hh_set_state.argtypes = [VP, P(HHStateView)]
    L.hh_set_state.restype = ctypes.c_int
    L.hh_launch_count.argtypes = [VP]
    L.hh_launch_count.restype = U64
    L.hh_gae.argtypes = [I32, I32, VP, VP, VP, VP, ctypes.c_float, ctypes.c_float, VP, VP, VP]
    L.hh_gae.restype = ctypes.c_int
    L.hh_gae_agents.argtypes = [I32, I32, I32, VP, VP, VP, VP, ctypes.c_float, ctypes.c_float, VP, VP, VP]
    L.hh_gae_agents.restype = ctypes.c_int
    L.hh_sample_actions.argtypes = [I32, VP, VP, U64, U64, VP, I32, VP, VP, VP]
    L.hh_sample_actions.restype = ctypes.c_int
    L.hh_pack_central.argtypes = [I32, I32, I32, VP, VP, VP, VP, VP]
    L.hh_pack_central.restype = ctypes.c_int
    L.hh_policy_forward.argtypes = [I32, P(HHPolicyChain), VP, VP, I32, VP]
    L.hh_policy_forward.restype = ctypes.c_int
    L.hh_policy_forward_ex.argtypes = [I32, P(HHPolicyChainEx), I32, VP]
    L.hh_policy_forward_ex.restype = ctypes.c_int
    L.hh_policy_pack.argtypes = [VP, I32, I32, I32, I32, I32, I32, I32, I32, VP, VP, VP]
    L.hh_policy_pack.restype = ctypes.c_int
    L.hh_policy_image_bytes.argtypes = [I32, I32]
    L.hh_policy_image_bytes.restype = ctypes.c_int64
    L.hh_policy_tc_pair.argtypes = []
    L.hh_policy_tc_pair.restype = I32
    L.hh_policy_tc_mode.argtypes = []
    L.hh_policy_tc_mode.restype = I32
    L.hh_policy_rows_by_key.argtypes = [I32, VP, I32, P(I32), VP, VP, VP]
    L.hh_policy_rows_by_key.restype = ctypes.c_int
    L.hh_policy_last_error.restype = ctypes.c_char_p
    L.hh_debug_geodesic.argtypes = [I32, I32, VP, VP]
    L.hh_debug_geodesic.restype = ctypes.c_int
    L.hh_hier_create.argtypes = [P(HHHierConfig), I32, I32, P(VP)]
    L.hh_hier_create.restype = ctypes.c_int
    L.hh_hier_destroy.argtypes = [VP]
    L.hh_hier_destroy.restype = None
    L.hh_hier_reset.argtypes = [VP, VP, VP, VP]
    L.hh_hier_reset.restype = ctypes.c_int
    for fn in ("hh_hier_begin", "hh_hier_agents", "hh_hier_tick"):
        getattr(L, fn).argtypes = [VP, VP, VP, VP, VP]
        getattr(L, fn).restype = ctypes.c_int
    L.hh_hier_end.argtypes = [VP, VP, VP, VP, VP, VP]
    L.hh_hier_end.restype = ctypes.c_int
    L.hh_hier_policy_rows.argtypes = [VP, VP, VP, VP, VP]
    L.hh_hier_policy_rows.restype = ctypes.c_int
    L.hh_hier_eval_info.argtypes = [VP, VP, VP]
    L.hh_hier_eval_info.restype = ctypes.c_int
    L.hh_hier_get_state.argtypes = [VP, VP]
    L.hh_hier_get_state.restype = ctypes.c_int
    L.hh_hier_set_state.argtypes = [VP, VP]
    L.hh_hier_set_state.restype = ctypes.c_int
    L.hh_hier_launch_count.argtypes = [VP]
    L.hh_hier_launch_count.restype = U64
    L.hh_hier_last_error.restype = ctypes.c_char_p
    L.hh_last_error.restype = ctypes.c_char_p
    L.hh_version.restype = ctypes.c_char_p
    return real


EXPORTS = ["hh_create", "hh_destroy", "hh_n_arenas", "hh_obs_dim", "hh_reset", "hh_step", "hh_step_range", "hh_step_range_central", "hh_fragment_prepare", "hh_fragment_writeback", "hh_multicat_forward", "hh_multicat_backward", "hh_ppo_loss", "hh_step_begin", "hh_step_finish", "hh_reset_host",
           "hh_step_host", "hh_step_host_begin", "hh_step_host_end", "hh_host_buffers", "hh_set_host_mode", "hh_get_state", "hh_set_state", "hh_launch_count", "hh_gae", "hh_gae_agents", "hh_sample_actions", "hh_pack_central", "hh_debug_geodesic", "hh_last_error", "hh_version", "hh_policy_forward", "hh_policy_forward_ex", "hh_policy_pack", "hh_policy_image_bytes", "hh_policy_tc_pair", "hh_policy_tc_mode", "hh_policy_rows_by_key", "hh_policy_last_error",
           "hh_hier_create", "hh_hier_destroy", "hh_hier_reset", "hh_hier_begin", "hh_hier_agents", "hh_hier_tick",
           "hh_hier_end", "hh_hier_policy_rows", "hh_hier_eval_info", "hh_hier_get_state", "hh_hier_set_state", "hh_hier_launch_count", "hh_hier_last_error"]


def check(rc: int, what: str):
    if rc != 0:
        msg = (lib().hh_hier_last_error() if what.startswith("hh_hier") else
               lib().hh_policy_last_error() if what.startswith("hh_policy") else lib().hh_last_error()).decode()
        if rc == -1:
            raise ValueError(f"{what}: {msg}")
        raise RuntimeError(f"{what}: {msg} (code {rc})")
