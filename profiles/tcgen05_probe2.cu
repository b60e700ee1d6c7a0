// profiles/tcgen05_probe2.cu -- second stand-alone probe for the tcgen05 policy-forward kernel (round 2).  The first
// probe (tcgen05_probe.cu, profiles/r2a_tcgen05_probe.txt) settled the shared-memory descriptor, the TMEM placement of an
// M = 64 accumulator (row r -> lane 32 (r / 16) + r % 16), operand truncation and the kind::tf32 rate (M = 64 runs at HALF
// the M = 128 rate).  This one checks what the fp16 split formulation ("H3") needs:
//   1. kind::f16, M = 64, N = 256, K = 16 from shared memory in the K-major no-swizzle canonical layout of 16-bit operands
//      (8-row x 16-byte core matrices = 8 x 8 halves; LBO = rows * 16 B between the two K halves, SBO = 128 B);
//   2. the accuracy of  A B ~ 2^-(s + t) (Ah Bh + Al Bh + Ah Bl)  with  Ah = rn_f16(2^s a), Al = rn_f16(2^s a - Ah)  (and the
//      same for B with 2^t): power-of-two prescales keep the low parts in fp16's normal range;
//   3. operands brought in by cp.async.bulk (the TMA engine's 1-D copy, SASS UBLKCP) completing on an mbarrier;
//   4. the register mapping of tcgen05.ld.16x256b for the M = 64 accumulator (16 lanes per warp quadrant);
//   5. the issue rate of the 3-MMA f16 step on all SMs;
//   6. a 2-CTA cluster in which each CTA fetches half of a weight stage and multicasts it to both (halves L2 reads).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -o build/tcgen05_probe2 profiles/tcgen05_probe2.cu
#include <cuda_fp16.h>
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
__device__ __forceinline__ uint64_t smem_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = (uint64_t)((saddr >> 4) & 0x3FFFu);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFFu) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFFu) << 32;
  d |= (uint64_t)1 << 46;
  return d;
}
// kind::f16: D = F32 (1 << 4), A = B = F16 (format 0), K-major, N >> 3 at bit 17, M >> 4 at bit 24
__host__ __device__ constexpr uint32_t instr_desc_f16(int M, int N) {
  return (1u << 4) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
__device__ __forceinline__ void umma_f16(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(tmem_d),
      "l"(da), "l"(db), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ uint32_t bounded_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok = 0;
  for (long long spin = 0; spin < (1ll << 24) && !ok; ++spin)
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}\n"
                 : "=r"(ok)
                 : "r"(smem_u32(bar)), "r"(parity)
                 : "memory");
  return ok;
}

constexpr int kM = 64, kN = 256, kK = 64;                 // 4 MMAs of K = 16 per term
constexpr int kABytes = kM * kK * 2, kBBytes = kN * kK * 2;
constexpr int kImage = 2 * kABytes + 2 * kBBytes;         // Ah | Al | Bh | Bl, each in the canonical layout

// host + device: byte offset of element (row, k) of a [rows x K] 16-bit operand in the canonical K-major layout
__host__ __device__ inline int canon16(int row, int k, int rows) { return (k >> 3) * (rows * 16) + row * 16 + (k & 7) * 2; }

__global__ void __launch_bounds__(128) h3_kernel(const uint8_t* __restrict__ image, float* __restrict__ dump32,
                                                 float* __restrict__ dump16, int* __restrict__ status, int reps, int dump) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ __align__(8) uint64_t bar_full, bar_mma;
  __shared__ uint32_t tmem_base_s;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  if (tid == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar_full)));
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar_mma)));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)), "r"(kN));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = tmem_base_s;

  if (tid == 0) {   // the whole operand image with ONE bulk copy (TMA engine), completing on bar_full
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&bar_full)), "r"((uint32_t)kImage) : "memory");
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(smem)),
                 "l"(image), "r"((uint32_t)kImage), "r"(smem_u32(&bar_full))
                 : "memory");
  }
  if (warp == 1 && lane == 0) {
    if (!bounded_wait(&bar_full, 0)) atomicExch(status, 2);
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t idesc = instr_desc_f16(kM, kN);
    const uint32_t lboA = kM * 16, lboB = kN * 16, sbo = 128;
    const uint32_t sAh = smem_u32(smem), sAl = sAh + kABytes, sBh = sAl + kABytes, sBl = sBh + kBBytes;
    for (int rep = 0; rep < reps; ++rep) {
#pragma unroll
      for (int ks = 0; ks < kK / 16; ++ks) {
        const uint32_t oA = ks * 2 * lboA, oB = ks * 2 * lboB;
        const uint64_t dah = smem_desc(sAh + oA, lboA, sbo), dal = smem_desc(sAl + oA, lboA, sbo);
        const uint64_t dbh = smem_desc(sBh + oB, lboB, sbo), dbl = smem_desc(sBl + oB, lboB, sbo);
        umma_f16(tmem, dal, dbh, idesc, (rep | ks) ? 1u : 0u);   // small terms first
        umma_f16(tmem, dah, dbl, idesc, 1u);
        umma_f16(tmem, dah, dbh, idesc, 1u);
      }
    }
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar_mma)) : "memory");
  }
  const uint32_t ok = bounded_wait(&bar_mma, 0);
  if (!ok && tid == 0) atomicExch(status, 1);
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  if (dump && ok) {
    for (int c = 0; c < kN; c += 8) {   // (a) 32x32b: thread = lane, 8 consecutive columns
      uint32_t r[8];
      const uint32_t taddr = tmem + ((uint32_t)(warp * 32) << 16) + (uint32_t)c;
      asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                   : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
                   : "r"(taddr));
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
      for (int j = 0; j < 8; ++j) dump32[(size_t)tid * kN + c + j] = __uint_as_float(r[j]);
    }
    for (int c = 0; c < kN; c += 8) {   // (b) 16x256b.x1: 16 lanes x 8 columns over the 32 threads, 4 registers each
      uint32_t r[4];
      const uint32_t taddr = tmem + ((uint32_t)(warp * 32) << 16) + (uint32_t)c;
      asm volatile("tcgen05.ld.sync.aligned.16x256b.x1.b32 {%0, %1, %2, %3}, [%4];"
                   : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
                   : "r"(taddr));
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
      for (int j = 0; j < 4; ++j) dump16[((size_t)(c / 8) * 128 + tid) * 4 + j] = __uint_as_float(r[j]);
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(kN));
}

// 6. cluster of two CTAs: CTA r fetches half r of a 16 KB stage and multicasts it into BOTH CTAs' shared memory; each
// CTA's barrier expects the whole stage.  Each CTA then writes what it holds to global memory for the host to compare.
constexpr int kStage = 16384;
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(128)
mcast_kernel(const uint8_t* __restrict__ src, uint8_t* __restrict__ out, int* __restrict__ status) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ __align__(8) uint64_t bar;
  const int tid = threadIdx.x;
  uint32_t rank;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(rank));
  if (tid == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&bar)), "r"((uint32_t)kStage) : "memory");
  }
  // both barriers are initialised and armed before either CTA's copy can signal the peer
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
  if (tid == 0) {
    const uint32_t half = kStage / 2;
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1], %2, [%3], %4;" ::"r"(
            smem_u32(smem) + rank * half),
        "l"(src + (size_t)(blockIdx.x / 2) * kStage + rank * half), "r"(half), "r"(smem_u32(&bar)), "h"((uint16_t)3)
        : "memory");
  }
  const uint32_t ok = bounded_wait(&bar, 0);
  if (!ok && tid == 0) atomicExch(status, 3);
  for (int i = tid; i < kStage / 4; i += 128)
    reinterpret_cast<uint32_t*>(out + (size_t)blockIdx.x * kStage)[i] = ok ? reinterpret_cast<uint32_t*>(smem)[i] : 0u;
  // a CTA must not exit while its peer may still be writing into its shared memory
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}

int main() {
  std::vector<float> A(kM * kK), B(kN * kK);
  srand(11);
  for (auto& v : A) v = (float)rand() / RAND_MAX * 2.0f - 1.0f;             // activations: tanh outputs / observations
  for (auto& v : B) v = ((float)rand() / RAND_MAX * 2.0f - 1.0f) * 0.09f;   // weights of an orthogonal 500 x 500 layer
  for (int i = 0; i < 64; ++i) A[i * 3 % A.size()] *= 1e-4f;                // and a few small ones
  for (int i = 0; i < 64; ++i) B[i * 7 % B.size()] *= 1e-3f;
  const int sA = 12;                                                        // activation prescale 2^12
  float bmax = 0;
  for (float v : B) bmax = std::fmax(bmax, std::fabs(v));
  int sB = 0;
  while (std::ldexp(bmax, sB + 1) < 16384.0f) ++sB;                         // max |2^sB b| in [2^13, 2^14)
  std::vector<uint8_t> image(kImage);
  auto put = [&](int base, int off, float v) { __half h = __float2half_rn(v); memcpy(&image[base + off], &h, 2); };
  double rep_err = 0;
  for (int r = 0; r < kM; ++r)
    for (int k = 0; k < kK; ++k) {
      const float v = std::ldexp(A[r * kK + k], sA);
      const float hi = __half2float(__float2half_rn(v)), lo = __half2float(__float2half_rn(v - hi));
      put(0, canon16(r, k, kM), hi);
      put(kABytes, canon16(r, k, kM), lo);
      if (v != 0) rep_err = std::fmax(rep_err, std::fabs(((double)hi + lo - v) / v));
    }
  for (int n = 0; n < kN; ++n)
    for (int k = 0; k < kK; ++k) {
      const float v = std::ldexp(B[n * kK + k], sB);
      const float hi = __half2float(__float2half_rn(v)), lo = __half2float(__float2half_rn(v - hi));
      put(2 * kABytes, canon16(n, k, kN), hi);
      put(2 * kABytes + kBBytes, canon16(n, k, kN), lo);
      if (v != 0) rep_err = std::fmax(rep_err, std::fabs(((double)hi + lo - v) / v));
    }
  printf("prescales 2^%d (A) 2^%d (B); worst relative error of hi + lo as a representation of the fp32 value: %.3e\n", sA, sB, rep_err);
  uint8_t* dImg;
  float *d32, *d16;
  int* dS;
  CK(cudaMalloc(&dImg, kImage));
  CK(cudaMalloc(&d32, 128 * kN * sizeof(float)));
  CK(cudaMalloc(&d16, (kN / 8) * 128 * 4 * sizeof(float)));
  CK(cudaMalloc(&dS, sizeof(int)));
  CK(cudaMemcpy(dImg, image.data(), kImage, cudaMemcpyHostToDevice));
  CK(cudaMemset(d32, 0, 128 * kN * sizeof(float)));
  CK(cudaMemset(d16, 0, (kN / 8) * 128 * 4 * sizeof(float)));
  CK(cudaMemset(dS, 0, sizeof(int)));
  CK(cudaFuncSetAttribute(h3_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kImage));
  h3_kernel<<<1, 128, kImage>>>(dImg, d32, d16, dS, 1, 1);
  CK(cudaDeviceSynchronize());
  int st = 0;
  CK(cudaMemcpy(&st, dS, sizeof(int), cudaMemcpyDeviceToHost));
  if (st) printf("h3: did not complete (status %d: 1 = MMA barrier, 2 = bulk-copy barrier)\n", st);
  else {
    std::vector<float> D32(128 * kN), D16((kN / 8) * 128 * 4);
    CK(cudaMemcpy(D32.data(), d32, D32.size() * sizeof(float), cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(D16.data(), d16, D16.size() * sizeof(float), cudaMemcpyDeviceToHost));
    const double unscale = std::ldexp(1.0, -(sA + sB));
    double worst = 0, scale = 0, worst_fp32 = 0;
    std::vector<double> exact((size_t)kM * kN);
    for (int r = 0; r < kM; ++r)
      for (int n = 0; n < kN; ++n) {
        double e = 0;
        float f = 0;
        for (int k = 0; k < kK; ++k) {
          e += (double)A[r * kK + k] * B[n * kK + k];
          f = fmaf(A[r * kK + k], B[n * kK + k], f);
        }
        exact[(size_t)r * kN + n] = e;
        const int lane = 32 * (r / 16) + r % 16;
        worst = std::fmax(worst, std::fabs(D32[(size_t)lane * kN + n] * unscale - e));
        worst_fp32 = std::fmax(worst_fp32, std::fabs((double)f - e));
        scale = std::fmax(scale, std::fabs(e));
      }
    printf("h3 (kind::f16 M=64 N=256 K=16, 3 MMAs per step, operands by cp.async.bulk): max |D - exact| = %.3e (max |exact| %.3f; a "
           "sequential fp32 fma loop: %.3e)\n", worst, scale, worst_fp32);
    // 16x256b.x1 mapping: expected thread t of warp q, register j -> row 16 q + t / 4 + 8 (j / 2), column c + 2 (t % 4) + j % 2
    int bad = 0, total = 0;
    for (int cb = 0; cb < kN / 8; ++cb)
      for (int t = 0; t < 128; ++t)
        for (int j = 0; j < 4; ++j) {
          const int q = t / 32, tt = t % 32, row = 16 * q + tt / 4 + 8 * (j / 2), col = cb * 8 + 2 * (tt % 4) + j % 2;
          const double want = exact[(size_t)row * kN + col];
          ++total;
          if (std::fabs(D16[((size_t)cb * 128 + t) * 4 + j] * unscale - want) > 1e-4 * (1 + std::fabs(want))) ++bad;
        }
    printf("tcgen05.ld.16x256b.x1 mapping (thread t, reg j) -> (row 16 q + t/4 + 8 (j/2), col c + 2 (t%%4) + j%%2): %d of %d mismatches\n", bad, total);
    if (bad) {   // print what thread 0..7 of warp 0 actually hold for the first column block
      for (int t = 0; t < 8; ++t)
        for (int j = 0; j < 4; ++j) {
          const double v = D16[((size_t)0 * 128 + t) * 4 + j] * unscale;
          int fr = -1, fc = -1;
          for (int r = 0; r < kM && fr < 0; ++r)
            for (int n = 0; n < 16; ++n)
              if (std::fabs(v - exact[(size_t)r * kN + n]) < 1e-5 * (1 + std::fabs(v))) { fr = r; fc = n; break; }
          printf("  t%d j%d -> row %d col %d\n", t, j, fr, fc);
        }
    }
  }
  {   // issue rate on every SM
    int dev = 0, sms = 0;
    CK(cudaGetDevice(&dev));
    CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    const int reps = 4000;
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    h3_kernel<<<sms, 128, kImage>>>(dImg, d32, d16, dS, 10, 0);
    CK(cudaDeviceSynchronize());
    CK(cudaEventRecord(e0));
    h3_kernel<<<sms, 128, kImage>>>(dImg, d32, d16, dS, reps, 0);
    CK(cudaEventRecord(e1));
    CK(cudaDeviceSynchronize());
    float ms = 0;
    CK(cudaEventElapsedTime(&ms, e0, e1));
    const double mmas = (double)sms * reps * (kK / 16) * 3;
    printf("M=64 N=256 K=16 kind::f16 from shared memory, %d SMs, 3 MMAs per K step: %.1f TFLOP/s of MMA work = %.1f TFLOP/s fp32-equivalent "
           "(%.3f ms, %.1f ns per MMA per SM)\n", sms, mmas * 2.0 * kM * kN * 16 / (ms * 1e-3) / 1e12,
           mmas * 2.0 * kM * kN * 16 / 3 / (ms * 1e-3) / 1e12, ms, ms * 1e6 / (mmas / sms));
  }
  {   // multicast
    const int n_clusters = 8;
    std::vector<uint8_t> src((size_t)n_clusters * kStage), out((size_t)2 * n_clusters * kStage);
    for (size_t i = 0; i < src.size(); ++i) src[i] = (uint8_t)((i * 2654435761u) >> 13);
    uint8_t *dSrc, *dOut;
    CK(cudaMalloc(&dSrc, src.size()));
    CK(cudaMalloc(&dOut, out.size()));
    CK(cudaMemcpy(dSrc, src.data(), src.size(), cudaMemcpyHostToDevice));
    CK(cudaMemset(dOut, 0, out.size()));
    CK(cudaMemset(dS, 0, sizeof(int)));
    CK(cudaFuncSetAttribute(mcast_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kStage));
    mcast_kernel<<<2 * n_clusters, 128, kStage>>>(dSrc, dOut, dS);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) printf("multicast kernel: %s\n", cudaGetErrorString(e));
    else {
      CK(cudaMemcpy(&st, dS, sizeof(int), cudaMemcpyDeviceToHost));
      CK(cudaMemcpy(out.data(), dOut, out.size(), cudaMemcpyDeviceToHost));
      size_t bad = 0;
      for (int b = 0; b < 2 * n_clusters; ++b)
        for (int i = 0; i < kStage; ++i) bad += out[(size_t)b * kStage + i] != src[(size_t)(b / 2) * kStage + i];
      printf("cluster multicast (2 CTAs, each fetches half a 16 KB stage for both): status %d, %zu of %zu bytes wrong\n", st, bad, out.size());
    }
  }
  return 0;
}
