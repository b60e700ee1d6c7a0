"""TEST INFRASTRUCTURE: builds tests/emu/libhh_emu.so (the v4 step schedule of hhmarl_2d_b200/csrc/hh_v4.cuh compiled
for the host, see hh_emu.cpp) and lets a test run VecLowLevelEnv's host API against it.  The product never loads it."""
import contextlib
import ctypes
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
CSRC = os.path.join(ROOT, "hhmarl_2d_b200", "csrc")
LIB = os.path.join(HERE, "libhh_emu.so")
DEPS = [os.path.join(HERE, f) for f in ("hh_emu.cpp", "hh_host_shim.h")] + [
    os.path.join(CSRC, f) for f in ("hh_v4.cuh", "hh_quad.cuh", "hh_core.cuh", "hh_geodesic.cuh", "hh_state_pack.h")]


def cuda_include():
    for d in (os.environ.get("CUDA_HOME"), "/usr/local/cuda"):
        if d and os.path.exists(os.path.join(d, "include", "cuda_runtime.h")):
            return os.path.join(d, "include")
    return None


def build(arenas_per_cta: int | None = None) -> str:
    lib = LIB if arenas_per_cta is None else LIB.replace(".so", f"_{arenas_per_cta}.so")
    if os.path.exists(lib) and all(os.path.getmtime(d) <= os.path.getmtime(lib) for d in DEPS):
        return lib
    inc = cuda_include()
    if inc is None:
        raise RuntimeError("cuda_runtime.h not found (needed for the vector types)")
    cmd = ["g++", "-O2", "-std=c++17", "-ffp-contract=off", "-fPIC", "-shared", "-Wno-attributes", "-I", inc]
    if arenas_per_cta is not None:
        cmd.append(f"-DHH_V4_ARENAS={arenas_per_cta}")
    subprocess.run(cmd + ["-o", lib, os.path.join(HERE, "hh_emu.cpp")], check=True)
    return lib


@contextlib.contextmanager
def emulated(reverse: bool = False, arenas_per_cta: int | None = None):
    """Inside the block hhmarl_2d_b200._native.lib() is the CPU emulation of the v4 schedule (host API only)."""
    from hhmarl_2d_b200 import _native as nat
    L = nat.bind(ctypes.CDLL(build(arenas_per_cta)), partial=True)
    L.hh_emu_set_reverse(1 if reverse else 0)
    saved = nat._lib
    nat._lib = L
    try:
        yield L
    finally:
        nat._lib = saved
        L.hh_emu_set_reverse(0)
