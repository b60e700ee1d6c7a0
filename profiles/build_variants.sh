#!/bin/bash
# profiles/build_variants.sh -- the library variants that the profiles/run_*.sh scripts select with HH_LIB_PATH
# (cross-compiles here; build/ is git-ignored but travels to the GPU box with the gpurun snapshot).
set -e
cd "$(dirname "$0")/../hhmarl_2d_b200/csrc"
mkdir -p ../../build
FLAGS="-gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo --expt-relaxed-constexpr -Xcompiler -fPIC -shared"
SRC="hh_api.cu hh_hier.cu hh_policy.cu"
nvcc $FLAGS -DHH_V4_PROFILE -o ../../build/lib_v4prof.so $SRC      # per-stage clock64() stamps (profiles/stage_clocks.py)
nvcc $FLAGS -DHH_V4_MIN_CTAS=3 -o ../../build/lib_v4occ3.so $SRC   # 3 CTAs per SM (80 registers)
nvcc $FLAGS -DHH_V4_WARM -o ../../build/lib_v4warm.so $SRC        # S2 idle warp pre-touches geo::direct_short (i-cache)
nvcc $FLAGS -DHH_V4_ROLLED -o ../../build/lib_v4rolled.so $SRC    # observation-row copy loops not unrolled (code size)
nvcc $FLAGS -DHH_V4_ROLLED -DHH_V4_WARM -o ../../build/lib_v4rolled_warm.so $SRC
# others used during round 1: -DHH_V4_ARENAS=16|64 (arenas per CTA), -DHH_PF_WARPS=16 (policy kernel warps)
# round-2 preparation: tcgen05 / TMEM probe (stand-alone binary, see the header of the source)
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -o ../../build/tcgen05_probe ../../profiles/tcgen05_probe.cu
ls -la ../../build
