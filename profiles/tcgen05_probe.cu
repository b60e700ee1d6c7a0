// profiles/tcgen05_probe.cu -- round-2 preparation for the policy-forward kernel on tcgen05 / TMEM (DESIGN.md section 8,
// item 1).  A stand-alone probe, NOT part of the library: it pins down, on the first GPU minute of round 2, the facts the
// fused kernel's design depends on and that cannot be checked without a GPU:
//   1. the shared-memory matrix descriptor for the K-major, no-swizzle canonical layout of 32-bit operands
//      (8-row x 16-byte core matrices; leading byte offset = distance of the two 16-byte K chunks of one K = 8 MMA,
//      stride byte offset = distance of consecutive 8-row groups) and the kind::tf32 instruction descriptor;
//   2. where D[row][col] of an M = 128 and of an M = 64 MMA lands in TMEM (lane, column) -- the probe dumps all 128 lanes
//      and searches the mapping;
//   3. how the tensor core narrows fp32 operands to TF32 (truncation or round-to-nearest): decides whether the raw fp32
//      activation tile can serve as the "hi" operand of the 3xTF32 split;
//   4. the accuracy of the 3xTF32 product (A_hi B_hi + A_lo B_hi + A_hi B_lo) against fp64;
//   5. the issue rate of back-to-back M = 64 / M = 128, N = 256 TF32 MMAs from shared memory on all SMs.
//   6. two chained layers on one CTA (accumulator -> tcgen05.ld -> tanh -> rewritten A tile -> next MMAs): the ordering
//      (tcgen05.wait::ld, fence.proxy.async, tcgen05 fences around the CTA barrier) the fused kernel relies on.
// Build + run (one GPU, under a timeout -- every wait in the kernel is bounded, a wrong descriptor cannot hang it):
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -o build/tcgen05_probe profiles/tcgen05_probe.cu
//   timeout 120 build/tcgen05_probe            (add `swap` to exchange the descriptor's two byte-offset fields)
#include <cuda_runtime.h>

#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#define CK(x)                                                                                  \
  do {                                                                                         \
    cudaError_t e_ = (x);                                                                      \
    if (e_ != cudaSuccess) {                                                                   \
      printf("CUDA error %s at %s:%d: %s\n", #x, __FILE__, __LINE__, cudaGetErrorString(e_)); \
      return 1;                                                                                \
    }                                                                                          \
  } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// 64-bit shared-memory matrix descriptor (cute/arch/mma_sm100_desc.hpp: SmemDescriptor): start address, leading and stride byte
// offsets in 16-byte units; version = 1 for sm_100; base offset 0; layout type 0 = no swizzle
__device__ __forceinline__ uint64_t smem_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = (uint64_t)((saddr >> 4) & 0x3FFFu);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFFu) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFFu) << 32;
  d |= (uint64_t)1 << 46;
  return d;
}

// 32-bit instruction descriptor (InstrDescriptor): D = F32 (1 << 4), A = B = TF32 (2 << 7, 2 << 10), both K-major,
// N >> 3 at bit 17, M >> 4 at bit 24
__host__ __device__ constexpr uint32_t instr_desc_tf32(int M, int N) {
  return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(tmem_d),
      "l"(da), "l"(db), "r"(idesc), "r"(accumulate)
      : "memory");
}

__device__ __forceinline__ float tf32_trunc(float v) { return __uint_as_float(__float_as_uint(v) & 0xFFFFE000u); }
__device__ __forceinline__ float tf32_rn(float v) {   // round to nearest (ties away), finite inputs
  return __uint_as_float((__float_as_uint(v) + 0x1000u) & 0xFFFFE000u);
}

// canonical K-major no-swizzle placement of element (row, k) of a [rows x K] operand, in floats
__device__ __forceinline__ int canon(int row, int k, int rows) { return (k >> 2) * (rows * 4) + row * 4 + (k & 3); }

constexpr int kN = 256;   // columns of D = TMEM columns allocated
constexpr int kK = 64;    // K of the staged tile: 8 MMAs of K = 8

// mode 0: D = A B (one MMA chain);  mode 1: 3xTF32 with the RAW fp32 values as "hi" operands (lo = v - trunc(v));
// mode 2: 3xTF32 with hi = rn_tf32(v) stored explicitly (lo = v - hi);  reps > 1 repeats the chain (issue-rate measurement)
template <int M>
__global__ void __launch_bounds__(128) probe_kernel(const float* __restrict__ A, const float* __restrict__ Bt,
                                                    float* __restrict__ tmem_dump, int* __restrict__ status, int mode,
                                                    int reps, int dump, int swap) {
  extern __shared__ __align__(1024) uint8_t smem[];
  float* sA = reinterpret_cast<float*>(smem);
  float* sB = sA + M * kK;
  float* sAl = sB + kN * kK;
  float* sBl = sAl + M * kK;
  __shared__ __align__(8) uint64_t bar;
  __shared__ uint32_t tmem_base_s;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

  for (int i = tid; i < M * kK; i += 128) {
    const int r = i / kK, k = i % kK;
    const float v = A[i];
    const float hi = mode == 2 ? tf32_rn(v) : (mode == 1 ? tf32_trunc(v) : v);
    sA[canon(r, k, M)] = mode == 2 ? hi : v;
    if (mode) sAl[canon(r, k, M)] = v - hi;
  }
  for (int i = tid; i < kN * kK; i += 128) {
    const int n = i / kK, k = i % kK;
    const float v = Bt[i];
    const float hi = mode == 2 ? tf32_rn(v) : (mode == 1 ? tf32_trunc(v) : v);
    sB[canon(n, k, kN)] = mode == 2 ? hi : v;
    if (mode) sBl[canon(n, k, kN)] = v - hi;
  }
  if (tid == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy stores -> visible to the tensor core
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)), "r"(kN));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = tmem_base_s;

  if (warp == 1 && lane == 0) {
    const uint32_t idesc = instr_desc_tf32(M, kN);
    const uint32_t lboA = M * 16, lboB = kN * 16, sbo = 128;
    // `swap` exchanges the two offset fields of the descriptors (if the first run finds no rows, try `tcgen05_probe swap`)
    auto desc = [&](uint32_t addr, uint32_t lbo) { return swap ? smem_desc(addr, sbo, lbo) : smem_desc(addr, lbo, sbo); };
    for (int rep = 0; rep < reps; ++rep) {
#pragma unroll
      for (int ks = 0; ks < kK / 8; ++ks) {
        const uint32_t offA = ks * 2 * lboA, offB = ks * 2 * lboB;
        const uint64_t da = desc(smem_u32(sA) + offA, lboA), db = desc(smem_u32(sB) + offB, lboB);
        umma_tf32(tmem, da, db, idesc, (rep | ks) ? 1u : 0u);
        if (mode) {
          umma_tf32(tmem, desc(smem_u32(sAl) + offA, lboA), db, idesc, 1u);
          umma_tf32(tmem, da, desc(smem_u32(sBl) + offB, lboB), idesc, 1u);
        }
      }
    }
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
  }
  // bounded wait for the MMAs (phase 0 of the barrier)
  uint32_t ok = 0;
  for (long long spin = 0; spin < (1ll << 24) && !ok; ++spin)
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}\n"
                 : "=r"(ok)
                 : "r"(smem_u32(&bar)), "r"(0u)
                 : "memory");
  if (!ok && tid == 0) atomicExch(status, 1);   // the MMAs never completed
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");

  if (dump && ok) {   // warp w reads lanes 32 w .. 32 w + 31, eight columns at a time
    for (int c = 0; c < kN; c += 8) {
      uint32_t r[8];
      const uint32_t taddr = tmem + ((uint32_t)(warp * 32) << 16) + (uint32_t)c;
      asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                   : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
                   : "r"(taddr));
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
      for (int j = 0; j < 8; ++j) tmem_dump[(size_t)blockIdx.x * 128 * kN + (size_t)tid * kN + c + j] = __uint_as_float(r[j]);
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(kN));
}

// Two chained layers on one CTA, M = 128 (the mechanics the fused policy kernel needs): D1 = A B in TMEM; every thread reads
// ITS row of D1 (lane = row for M = 128), applies tanh and writes the first kK columns back INTO the A tile in the canonical
// layout; then D2 = tanh(D1[:, :kK]) B overwrites the accumulator.  Exercises tcgen05.ld -> generic-proxy smem store ->
// fence.proxy.async -> tcgen05.mma ordering and the second phase of the completion barrier.
__global__ void __launch_bounds__(128) chain_kernel(const float* __restrict__ A, const float* __restrict__ Bt,
                                                    float* __restrict__ tmem_dump, int* __restrict__ status) {
  constexpr int M = 128;
  extern __shared__ __align__(1024) uint8_t smem[];
  float* sA = reinterpret_cast<float*>(smem);
  float* sB = sA + M * kK;
  __shared__ __align__(8) uint64_t bar;
  __shared__ uint32_t tmem_base_s;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  for (int i = tid; i < M * kK; i += 128) sA[canon(i / kK, i % kK, M)] = A[i];
  for (int i = tid; i < kN * kK; i += 128) sB[canon(i / kK, i % kK, kN)] = Bt[i];
  if (tid == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)), "r"(kN));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = tmem_base_s;
  const uint32_t idesc = instr_desc_tf32(M, kN), lboA = M * 16, lboB = kN * 16, sbo = 128;
  uint32_t ok = 1;
  for (int layer = 0; layer < 2 && ok; ++layer) {
    if (warp == 1 && lane == 0) {
#pragma unroll
      for (int ks = 0; ks < kK / 8; ++ks)
        umma_tf32(tmem, smem_desc(smem_u32(sA) + ks * 2 * lboA, lboA, sbo), smem_desc(smem_u32(sB) + ks * 2 * lboB, lboB, sbo), idesc,
                  ks ? 1u : 0u);
      asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
    }
    ok = 0;
    for (long long spin = 0; spin < (1ll << 24) && !ok; ++spin)
      asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}\n"
                   : "=r"(ok)
                   : "r"(smem_u32(&bar)), "r"((uint32_t)layer)
                   : "memory");
    if (!ok) {
      if (tid == 0) atomicExch(status, 1 + layer);
      break;
    }
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const int ncols = layer == 0 ? kK : kN;          // layer 0: only the columns that feed layer 1
    for (int c = 0; c < ncols; c += 8) {
      uint32_t r[8];
      const uint32_t taddr = tmem + ((uint32_t)(warp * 32) << 16) + (uint32_t)c;
      asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                   : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
                   : "r"(taddr));
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
      for (int j = 0; j < 8; ++j) {
        if (layer == 0) sA[canon(tid, c + j, M)] = tanhf(__uint_as_float(r[j]));     // row = lane = tid for M = 128
        else tmem_dump[(size_t)tid * kN + c + j] = __uint_as_float(r[j]);
      }
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");     // the rewritten A tile -> visible to the tensor core
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  }
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(kN));
}

static float h_trunc(float v) { uint32_t u; memcpy(&u, &v, 4); u &= 0xFFFFE000u; memcpy(&v, &u, 4); return v; }
static float h_rn(float v) { uint32_t u; memcpy(&u, &v, 4); u = (u + 0x1000u) & 0xFFFFE000u; memcpy(&v, &u, 4); return v; }

template <int M>
static int run(const std::vector<float>& A, const std::vector<float>& Bt, float* dA, float* dB, float* dD, int* dS, int swap) {
  const size_t smem = (size_t)(2 * (M + kN) * kK) * sizeof(float);
  CK(cudaFuncSetAttribute(probe_kernel<M>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  std::vector<float> D(128 * kN);
  std::vector<double> exact((size_t)M * kN), rt((size_t)M * kN), rr((size_t)M * kN);
  for (int r = 0; r < M; ++r)
    for (int n = 0; n < kN; ++n) {
      double e = 0, t = 0, q = 0;
      for (int k = 0; k < kK; ++k) {
        const float a = A[r * kK + k], b = Bt[n * kK + k];
        e += (double)a * b;
        t += (double)h_trunc(a) * h_trunc(b);
        q += (double)h_rn(a) * h_rn(b);
      }
      exact[(size_t)r * kN + n] = e; rt[(size_t)r * kN + n] = t; rr[(size_t)r * kN + n] = q;
    }
  for (int mode = 0; mode < 3; ++mode) {
    CK(cudaMemset(dD, 0, 128 * kN * sizeof(float)));
    CK(cudaMemset(dS, 0, sizeof(int)));
    probe_kernel<M><<<1, 128, smem>>>(dA, dB, dD, dS, mode, 1, 1, swap);
    CK(cudaDeviceSynchronize());
    int st = 0;
    CK(cudaMemcpy(&st, dS, sizeof(int), cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(D.data(), dD, D.size() * sizeof(float), cudaMemcpyDeviceToHost));
    if (st) { printf("M=%d mode %d: the MMAs did not complete (barrier never flipped)\n", M, mode); continue; }
    if (mode == 0) {
      // where does D[r][n] live?  hypothesis for M = 128: lane r, column n.  Otherwise search rows by their first column.
      std::vector<int> lane_of(M, -1);
      int identity = 1;
      for (int r = 0; r < M; ++r) {
        for (int l = 0; l < 128 && lane_of[r] < 0; ++l) {
          int good = 1;
          for (int n = 0; n < 8 && good; ++n) good = std::fabs(D[(size_t)l * kN + n] - rt[(size_t)r * kN + n]) < 2e-2 * (1 + std::fabs(rt[(size_t)r * kN + n]));
          if (good) lane_of[r] = l;
        }
        identity &= lane_of[r] == r;
      }
      printf("M=%d: TMEM lane of D row r: %s;", M, identity ? "lane = r" : "NOT the identity:");
      if (!identity)
        for (int r = 0; r < M; r += (M == 64 ? 4 : 16)) printf(" r%d->%d", r, lane_of[r]);
      double et = 0, er = 0;
      int missing = 0;
      for (int r = 0; r < M; ++r) {
        if (lane_of[r] < 0) { ++missing; continue; }
        for (int n = 0; n < kN; ++n) {
          const double d = D[(size_t)lane_of[r] * kN + n];
          et = std::fmax(et, std::fabs(d - rt[(size_t)r * kN + n]));
          er = std::fmax(er, std::fabs(d - rr[(size_t)r * kN + n]));
        }
      }
      printf(" rows not found %d; max |D - ref| with truncated inputs %.3e, with round-to-nearest inputs %.3e -> operands are %s\n",
             missing, et, er, et < er ? "TRUNCATED to TF32" : "ROUNDED to TF32");
      if (missing) return 0;
      // keep the mapping for the 3x modes
      static std::vector<int> keep;
      keep = lane_of;
    } else {
      double worst = 0, scale = 0;
      for (int r = 0; r < M; ++r)
        for (int n = 0; n < kN; ++n) {
          // rows were found in mode 0; M = 64 may not be the identity, so search by value again through the exact product
          double best = 1e30;
          for (int l = 0; l < 128; ++l) best = std::fmin(best, std::fabs((double)D[(size_t)l * kN + n] - exact[(size_t)r * kN + n]));
          worst = std::fmax(worst, best);
          scale = std::fmax(scale, std::fabs(exact[(size_t)r * kN + n]));
        }
      printf("M=%d 3xTF32 (%s): max |D - exact| = %.3e (max |exact| %.2f; fp32 rounding of the result alone ~%.1e)\n", M,
             mode == 1 ? "raw fp32 as hi, lo = v - trunc(v)" : "hi = rn_tf32(v) stored, lo = v - hi", worst, scale, scale * 6e-8);
    }
  }
  // issue rate on every SM: reps chains of 8 (x3) MMAs, no dump
  int dev = 0, sms = 0;
  CK(cudaGetDevice(&dev));
  CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  for (int mode = 0; mode < 2; ++mode) {
    const int reps = 2000;
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    probe_kernel<M><<<sms, 128, smem>>>(dA, dB, dD, dS, mode, 10, 0, swap);
    CK(cudaDeviceSynchronize());
    CK(cudaEventRecord(e0));
    probe_kernel<M><<<sms, 128, smem>>>(dA, dB, dD, dS, mode, reps, 0, swap);
    CK(cudaEventRecord(e1));
    CK(cudaDeviceSynchronize());
    float ms = 0;
    CK(cudaEventElapsedTime(&ms, e0, e1));
    const double mmas = (double)sms * reps * (kK / 8) * (mode ? 3 : 1);
    printf("M=%d N=%d K=8 kind::tf32 from shared memory, %d SMs, %s: %.1f TFLOP/s (%.3f ms, %.1f ns per MMA per SM)\n", M, kN, sms,
           mode ? "3 MMAs per K step (3xTF32)" : "1 MMA per K step", mmas * 2.0 * M * kN * 8 / (ms * 1e-3) / 1e12, ms,
           ms * 1e6 / (mmas / sms));
  }
  return 0;
}

int main(int argc, char** argv) {
  const int swap = argc > 1 && !strcmp(argv[1], "swap");
  std::vector<float> A(128 * kK), Bt(kN * kK);
  srand(7);
  for (auto& v : A) v = (float)rand() / RAND_MAX * 2.0f - 1.0f;
  for (auto& v : Bt) v = (float)rand() / RAND_MAX * 2.0f - 1.0f;
  float *dA, *dB, *dD;
  int* dS;
  CK(cudaMalloc(&dA, A.size() * sizeof(float)));
  CK(cudaMalloc(&dB, Bt.size() * sizeof(float)));
  CK(cudaMalloc(&dD, (size_t)148 * 128 * kN * sizeof(float)));
  CK(cudaMalloc(&dS, sizeof(int)));
  CK(cudaMemcpy(dA, A.data(), A.size() * sizeof(float), cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dB, Bt.data(), Bt.size() * sizeof(float), cudaMemcpyHostToDevice));
  if (run<128>(A, Bt, dA, dB, dD, dS, swap)) return 1;
  {   // two chained layers (only meaningful once the single product above is right)
    const size_t smem = (size_t)((128 + kN) * kK) * sizeof(float);
    CK(cudaFuncSetAttribute(chain_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    CK(cudaMemset(dD, 0, 128 * kN * sizeof(float)));
    CK(cudaMemset(dS, 0, sizeof(int)));
    chain_kernel<<<1, 128, smem>>>(dA, dB, dD, dS);
    CK(cudaDeviceSynchronize());
    int st = 0;
    std::vector<float> D(128 * kN);
    CK(cudaMemcpy(&st, dS, sizeof(int), cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(D.data(), dD, D.size() * sizeof(float), cudaMemcpyDeviceToHost));
    if (st) printf("chain: layer %d never completed\n", st);
    else {
      double worst = 0, scale = 0;
      for (int r = 0; r < 128; ++r) {
        double h[kK];
        for (int c = 0; c < kK; ++c) {
          double acc = 0;
          for (int k = 0; k < kK; ++k) acc += (double)h_trunc(A[r * kK + k]) * h_trunc(Bt[c * kK + k]);
          h[c] = std::tanh(acc);
        }
        for (int n = 0; n < kN; ++n) {
          double acc = 0;
          for (int k = 0; k < kK; ++k) acc += h[k] * (double)h_trunc(Bt[n * kK + k]);
          worst = std::fmax(worst, std::fabs(acc - D[(size_t)r * kN + n]));
          scale = std::fmax(scale, std::fabs(acc));
        }
      }
      printf("chain (M=128): D2 = tanh(D1[:, :%d]) B, max |D2 - ref| = %.3e (max |ref| %.2f; TF32 operands -> expect ~1e-2)\n", kK, worst, scale);
    }
  }
  if (run<64>(A, Bt, dA, dB, dD, dS, swap)) return 1;
  return 0;
}
