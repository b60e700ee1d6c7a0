// hh_host_shim.h -- TEST INFRASTRUCTURE.  Lets g++ compile the device headers of hhmarl_2d_b200/csrc so that the
// v4 step schedule (hh_v4.cuh) can be executed stage by stage on a CPU and compared with the oracle
// (tests/test_emu_v4.py).  Nothing in the product includes this file; the product is the nvcc build.
#pragma once
#include <cuda_runtime.h>   // vector types (double2, int4, float4 ...), __device__ / __forceinline__ as no-ops
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include <new>      // every standard header the harness uses comes BEFORE __noinline__ is defined:
#include <string>   // libstdc++ spells the attribute __attribute__((__noinline__)) itself
#include <vector>

#ifndef __noinline__
#define __noinline__ __attribute__((noinline))
#endif
#ifndef __restrict__
#define __restrict__
#endif

// ---- arithmetic intrinsics (round-to-nearest, never fused: g++ is run with -ffp-contract=off)
static inline double __dadd_rn(double a, double b) { return a + b; }
static inline double __dmul_rn(double a, double b) { return a * b; }
static inline uint32_t __umulhi(uint32_t a, uint32_t b) { return (uint32_t)(((uint64_t)a * (uint64_t)b) >> 32); }
static inline int __popc(unsigned int x) { return __builtin_popcount(x); }
static inline double rsqrt(double x) { return 1.0 / sqrt(x); }
static inline void sincospi(double x, double* s, double* c) {
  // callers pass |x| <= 0.25 (exact quadrant reduction is done before); pi * x is accurate to 1 ulp there
  const double a = 3.14159265358979323846 * x;
  *s = sin(a);
  *c = cos(a);
}
static inline int atomicOr(int* p, int v) {
  const int old = *p;
  *p = old | v;
  return old;
}
static inline int min(int a, int b) { return a < b ? a : b; }

// ---- warp collectives appear only in the quad kernels' helpers (hh_quad.cuh), which the emulation never calls
struct EmuDim3 { unsigned x, y, z; };
static const EmuDim3 threadIdx = {0, 0, 0};
template <typename T>
static inline T __shfl_sync(unsigned, T, int, int = 32) { abort(); }
template <typename T>
static inline T __shfl_xor_sync(unsigned, T, int, int = 32) { abort(); }
static inline unsigned __ballot_sync(unsigned, bool) { abort(); }
static inline int __ffs(int x) { return __builtin_ffs(x); }
