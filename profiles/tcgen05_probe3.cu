// profiles/tcgen05_probe3.cu -- third probe: the CTA-PAIR form of the MMA (tcgen05.mma.cta_group::2) that the policy forward
// needs to halve its weight stream (each CTA of a pair holds 64 rows of A and HALF of the B columns; one MMA of M = 128
// feeds both SMs' tensor cores).  Checks, on a cluster of two CTAs:
//   1. tcgen05.alloc / mma / commit with cta_group::2 (allocation by one warp of EACH CTA; the MMA issued by CTA 0 only;
//      the commit multicast to both CTAs' barriers);
//   2. which operand rows each CTA supplies and where D lands in each CTA's tensor memory (expected from CUTLASS'
//      tmem_frg_2sm: row m of the CTA's 64 rows in lane m for columns [0, N/2) and in lane 64 + m for columns [N/2, N));
//   3. the issue rate of the 3-MMA f16 step in that form;
//   4. whether a bulk copy may complete on the PEER CTA's mbarrier (destination local, barrier remote), which would let
//      each CTA fetch its own weight half and still signal the issuing CTA directly.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -o build/tcgen05_probe3 profiles/tcgen05_probe3.cu
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
__host__ __device__ constexpr uint32_t instr_desc_f16(int M, int N) {
  return (1u << 4) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
__device__ __forceinline__ void umma2_f16(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(tmem_d),
      "l"(da), "l"(db), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ uint32_t bounded_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok = 0;
  for (long long spin = 0; spin < (1ll << 22) && !ok; ++spin)
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}\n"
                 : "=r"(ok)
                 : "r"(bar), "r"(parity)
                 : "memory");
  return ok;
}
__device__ __forceinline__ void cluster_sync() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}

constexpr int kM = 128, kN = 256, kK = 64;           // pair tile; each CTA: 64 rows of A, 128 rows (columns of D) of B
constexpr int kAB = 64 * kK * 2, kBB = (kN / 2) * kK * 2;
__host__ __device__ inline int canon16(int row, int k, int rows) { return (k >> 3) * (rows * 16) + row * 16 + (k & 7) * 2; }

// A [128][kK], Bt [256][kK] halves (row-major, fp16 already); CTA r stages A rows 64 r .. and Bt rows 128 r ..
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(128)
pair_kernel(const __half* __restrict__ A, const __half* __restrict__ Bt, float* __restrict__ dump, int* __restrict__ status, int reps,
            int do_dump) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t *sA = smem, *sB = smem + kAB;
  __shared__ __align__(8) uint64_t bar_mma;
  __shared__ uint32_t tmem_base_s;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  uint32_t rank;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(rank));
  const int pair = blockIdx.x >> 1;
  (void)pair;
  for (int i = tid; i < 64 * kK; i += 128) {
    const int r = i / kK, k = i % kK;
    *reinterpret_cast<__half*>(sA + canon16(r, k, 64)) = A[(size_t)(rank * 64 + r) * kK + k];
  }
  for (int i = tid; i < (kN / 2) * kK; i += 128) {
    const int n = i / kK, k = i % kK;
    *reinterpret_cast<__half*>(sB + canon16(n, k, kN / 2)) = Bt[(size_t)(rank * (kN / 2) + n) * kK + k];
  }
  if (tid == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar_mma)));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)), "r"(kN));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  cluster_sync();        // both CTAs' operands, barriers and allocations are in place
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = tmem_base_s;

  if (rank == 0 && warp == 1 && lane == 0) {
    const uint32_t idesc = instr_desc_f16(kM, kN);
    const uint32_t lboA = 64 * 16, lboB = (kN / 2) * 16, sbo = 128;
    for (int rep = 0; rep < reps; ++rep) {
#pragma unroll
      for (int ks = 0; ks < kK / 16; ++ks) {
        const uint64_t da = smem_desc(smem_u32(sA) + ks * 2 * lboA, lboA, sbo), db = smem_desc(smem_u32(sB) + ks * 2 * lboB, lboB, sbo);
        umma2_f16(tmem, da, db, idesc, (rep | ks) ? 1u : 0u);
        if (reps > 1) {   // the 3-MMA step of the H3 product (same operands: only the rate matters here)
          umma2_f16(tmem, da, db, idesc, 1u);
          umma2_f16(tmem, da, db, idesc, 1u);
        }
      }
    }
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(
                     smem_u32(&bar_mma)),
                 "h"((uint16_t)3)
                 : "memory");
  }
  const uint32_t ok = bounded_wait(smem_u32(&bar_mma), 0);
  if (!ok && tid == 0) atomicExch(status, 1 + (int)rank);
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  if (do_dump && ok) {
    for (int c = 0; c < kN; c += 8) {
      uint32_t r[8];
      const uint32_t taddr = tmem + ((uint32_t)(warp * 32) << 16) + (uint32_t)c;
      asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                   : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
                   : "r"(taddr));
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
      for (int j = 0; j < 8; ++j) dump[((size_t)blockIdx.x * 128 + tid) * kN + c + j] = __uint_as_float(r[j]);
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  cluster_sync();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(kN));
}

// 4. bulk copy with a LOCAL destination completing on the PEER's barrier: CTA 1 fetches into its own shared memory and
// names CTA 0's barrier; CTA 0 waits on that barrier, then (after a cluster barrier) CTA 1's data is checked.
constexpr int kBytes = 8192;
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(64)
remote_bar_kernel(const uint8_t* __restrict__ src, uint8_t* __restrict__ out, int* __restrict__ status) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ __align__(8) uint64_t bar;
  const int tid = threadIdx.x;
  uint32_t rank;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(rank));
  if (tid == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    if (rank == 0) asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&bar)), "r"((uint32_t)kBytes) : "memory");
  }
  cluster_sync();
  if (rank == 1 && tid == 0) {
    uint32_t remote_bar;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(remote_bar) : "r"(smem_u32(&bar)), "r"(0u));
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(smem)), "l"(src),
                 "r"((uint32_t)kBytes), "r"(remote_bar)
                 : "memory");
  }
  uint32_t ok = 1;
  if (rank == 0) {
    ok = bounded_wait(smem_u32(&bar), 0);
    if (!ok && tid == 0) atomicExch(status, 7);
  }
  cluster_sync();     // CTA 0 has seen the completion -> CTA 1's bytes must be there
  if (rank == 1)
    for (int i = tid; i < kBytes / 4; i += 64) reinterpret_cast<uint32_t*>(out)[i] = reinterpret_cast<uint32_t*>(smem)[i];
  cluster_sync();
}

int main() {
  std::vector<float> A(kM * kK), B(kN * kK);
  srand(5);
  for (auto& v : A) v = (float)(rand() % 17 - 8) * 0.125f;    // exactly representable in fp16; products sum exactly in fp32
  for (auto& v : B) v = (float)(rand() % 13 - 6) * 0.25f;
  std::vector<__half> hA(A.size()), hB(B.size());
  for (size_t i = 0; i < A.size(); ++i) hA[i] = __float2half_rn(A[i]);
  for (size_t i = 0; i < B.size(); ++i) hB[i] = __float2half_rn(B[i]);
  __half *dA, *dB;
  float* dD;
  int* dS;
  const int n_pairs_rate = 74;
  CK(cudaMalloc(&dA, hA.size() * 2));
  CK(cudaMalloc(&dB, hB.size() * 2));
  CK(cudaMalloc(&dD, (size_t)2 * n_pairs_rate * 128 * kN * sizeof(float)));
  CK(cudaMalloc(&dS, sizeof(int)));
  CK(cudaMemcpy(dA, hA.data(), hA.size() * 2, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dB, hB.data(), hB.size() * 2, cudaMemcpyHostToDevice));
  CK(cudaMemset(dD, 0, (size_t)2 * 128 * kN * sizeof(float)));
  CK(cudaMemset(dS, 0, sizeof(int)));
  const int smem = kAB + kBB;
  CK(cudaFuncSetAttribute(pair_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  pair_kernel<<<2, 128, smem>>>(dA, dB, dD, dS, 1, 1);
  cudaError_t e = cudaDeviceSynchronize();
  int st = 0;
  if (e != cudaSuccess) {
    printf("pair kernel: %s\n", cudaGetErrorString(e));
    return 1;
  }
  CK(cudaMemcpy(&st, dS, sizeof(int), cudaMemcpyDeviceToHost));
  if (st) printf("pair kernel: the MMAs did not complete in CTA %d (barrier never flipped)\n", st - 1);
  else {
    std::vector<float> D((size_t)2 * 128 * kN);
    CK(cudaMemcpy(D.data(), dD, D.size() * sizeof(float), cudaMemcpyDeviceToHost));
    std::vector<float> exact((size_t)kM * kN);
    for (int r = 0; r < kM; ++r)
      for (int n = 0; n < kN; ++n) {
        float s = 0;
        for (int k = 0; k < kK; ++k) s += A[r * kK + k] * B[n * kK + k];
        exact[(size_t)r * kN + n] = s;
      }
    // expected: CTA c, row m (global row 64 c + m), column n -> lane m + 64 (n / 128), TMEM column n % 128
    int bad = 0;
    for (int c = 0; c < 2; ++c)
      for (int m = 0; m < 64; ++m)
        for (int n = 0; n < kN; ++n)
          bad += D[((size_t)c * 128 + m + 64 * (n / 128)) * kN + (n % 128)] != exact[(size_t)(64 * c + m) * kN + n];
    printf("cta_group::2 M=128 N=256: D[64 c + m][n] at CTA c, lane m + 64 (n / 128), column n %% 128: %d of %d mismatches\n", bad, 2 * 64 * kN);
    if (bad) {   // search: where do a few probe elements live?
      for (int c = 0; c < 2; ++c)
        for (int m : {0, 1, 17, 40}) {
          for (int n : {0, 5, 128, 200}) {
            const float want = exact[(size_t)(64 * c + m) * kN + n];
            int found = 0;
            for (int cc = 0; cc < 2 && found < 3; ++cc)
              for (int l = 0; l < 128 && found < 3; ++l)
                for (int col = 0; col < kN && found < 3; ++col)
                  if (D[((size_t)cc * 128 + l) * kN + col] == want && want != 0) {
                    printf("  D[%d][%d] = %g found at CTA %d lane %d col %d\n", 64 * c + m, n, want, cc, l, col);
                    ++found;
                  }
          }
        }
    }
  }
  {   // rate on 74 pairs
    const int reps = 4000;
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    pair_kernel<<<2 * n_pairs_rate, 128, smem>>>(dA, dB, dD, dS, 10, 0);
    CK(cudaDeviceSynchronize());
    CK(cudaEventRecord(e0));
    pair_kernel<<<2 * n_pairs_rate, 128, smem>>>(dA, dB, dD, dS, reps, 0);
    CK(cudaEventRecord(e1));
    e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("rate kernel: %s\n", cudaGetErrorString(e)); return 1; }
    float ms = 0;
    CK(cudaEventElapsedTime(&ms, e0, e1));
    const double mmas = (double)n_pairs_rate * reps * (kK / 16) * 3;
    printf("cta_group::2 M=128 N=256 K=16 kind::f16, %d CTA pairs, 3 MMAs per K step: %.1f TFLOP/s of MMA work = %.1f TFLOP/s fp32-equivalent "
           "(%.3f ms, %.1f ns per MMA per pair)\n", n_pairs_rate, mmas * 2.0 * kM * kN * 16 / (ms * 1e-3) / 1e12,
           mmas * 2.0 * kM * kN * 16 / 3 / (ms * 1e-3) / 1e12, ms, ms * 1e6 / (mmas / n_pairs_rate));
  }
  {   // bulk copy completing on the peer's barrier
    std::vector<uint8_t> src(kBytes), out(kBytes);
    for (int i = 0; i < kBytes; ++i) src[i] = (uint8_t)((i * 2654435761u) >> 11);
    uint8_t *dSrc, *dOut;
    CK(cudaMalloc(&dSrc, kBytes));
    CK(cudaMalloc(&dOut, kBytes));
    CK(cudaMemcpy(dSrc, src.data(), kBytes, cudaMemcpyHostToDevice));
    CK(cudaMemset(dOut, 0, kBytes));
    CK(cudaMemset(dS, 0, sizeof(int)));
    CK(cudaFuncSetAttribute(remote_bar_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kBytes));
    remote_bar_kernel<<<2, 64, kBytes>>>(dSrc, dOut, dS);
    e = cudaDeviceSynchronize();
    if (e != cudaSuccess) printf("remote-barrier bulk copy: %s\n", cudaGetErrorString(e));
    else {
      CK(cudaMemcpy(&st, dS, sizeof(int), cudaMemcpyDeviceToHost));
      CK(cudaMemcpy(out.data(), dOut, kBytes, cudaMemcpyDeviceToHost));
      int bad = 0;
      for (int i = 0; i < kBytes; ++i) bad += out[i] != src[i];
      printf("bulk copy (destination in CTA 1, completion on CTA 0's barrier): status %d (7 = CTA 0 never saw it), %d of %d bytes wrong\n", st, bad,
             kBytes);
    }
  }
  return 0;
}
