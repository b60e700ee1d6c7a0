// profiles/tcgen05_probe4.cu -- fourth probe: the A operand of a kind::f16 MMA taken from TENSOR MEMORY (M = 128), which is
// what a 128-row tile of the policy forward needs (hi halves of the activations in shared memory, lo halves in TMEM):
//   1. layout of a 16-bit A operand in TMEM: lane = row, 32-bit column c of the operand holds elements k = 2 c, 2 c + 1
//      (written here with tcgen05.st.32x32b by the thread that owns the lane);
//   2. D = A B with A from TMEM, B from shared memory (K-major canonical layout), checked exactly;
//   3. issue rate of M = 128 N = 256 K = 16: A from TMEM vs A from shared memory, on all SMs.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -o build/tcgen05_probe4 profiles/tcgen05_probe4.cu
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
__device__ __forceinline__ uint32_t bounded_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok = 0;
  for (long long spin = 0; spin < (1ll << 22) && !ok; ++spin)
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}\n"
                 : "=r"(ok)
                 : "r"(bar), "r"(parity)
                 : "memory");
  return ok;
}
constexpr int kM = 128, kN = 256, kK = 64;
constexpr int kColsD = kN, kColsA = kK / 2;       // TMEM: accumulator columns, then the A operand (2 halves per column)
__host__ __device__ inline int canon16(int row, int k, int rows) { return (k >> 3) * (rows * 16) + row * 16 + (k & 7) * 2; }

// mode 0: A from TMEM, mode 1: A from shared memory
__global__ void __launch_bounds__(128) ts_kernel(const __half* __restrict__ A, const __half* __restrict__ Bt, float* __restrict__ dump,
                                                 int* __restrict__ status, int reps, int do_dump, int mode) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t *sB = smem, *sA = smem + kN * kK * 2;
  __shared__ __align__(8) uint64_t bar;
  __shared__ uint32_t tmem_base_s;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  for (int i = tid; i < kN * kK; i += 128) *reinterpret_cast<__half*>(sB + canon16(i / kK, i % kK, kN)) = Bt[i];
  for (int i = tid; i < kM * kK; i += 128) *reinterpret_cast<__half*>(sA + canon16(i / kK, i % kK, kM)) = A[i];
  if (tid == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)), "r"(512));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = tmem_base_s;
  // A into TMEM: thread tid owns lane tid = row tid; column kColsD + c holds (A[row][2c], A[row][2c+1])
  for (int c = 0; c < kColsA; c += 8) {
    uint32_t r[8];
    for (int j = 0; j < 8; ++j) {
      const __half2 h = __halves2half2(A[(size_t)tid * kK + 2 * (c + j)], A[(size_t)tid * kK + 2 * (c + j) + 1]);
      r[j] = *reinterpret_cast<const uint32_t*>(&h);
    }
    const uint32_t taddr = tmem + ((uint32_t)(warp * 32) << 16) + (uint32_t)(kColsD + c);
    asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"r"(taddr), "r"(r[0]), "r"(r[1]),
                 "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7])
                 : "memory");
  }
  asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");

  if (warp == 1) {
    const uint32_t idesc = instr_desc_f16(kM, kN);
    const uint32_t lboA = kM * 16, lboB = kN * 16, sbo = 128;
    for (int rep = 0; rep < reps; ++rep) {
#pragma unroll
      for (int ks = 0; ks < kK / 16; ++ks) {
        const uint64_t db = smem_desc(smem_u32(sB) + ks * 2 * lboB, lboB, sbo);
        const uint32_t acc = (rep | ks) ? 1u : 0u;
        const int n_mma = reps > 1 ? 3 : 1;
        for (int q = 0; q < n_mma; ++q) {
          if (mode == 0) {
            const uint32_t ta = tmem + (uint32_t)(kColsD + ks * 8);      // 16 halves = 8 columns per K step
            asm volatile(
                "{\n\t.reg .pred p, e;\n\telect.sync _|e, 0xffffffff;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                "@e tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}\n" ::"r"(tmem),
                "r"(ta), "l"(db), "r"(idesc), "r"(q ? 1u : acc)
                : "memory");
          } else {
            const uint64_t da = smem_desc(smem_u32(sA) + ks * 2 * lboA, lboA, sbo);
            asm volatile(
                "{\n\t.reg .pred p, e;\n\telect.sync _|e, 0xffffffff;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                "@e tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(tmem),
                "l"(da), "l"(db), "r"(idesc), "r"(q ? 1u : acc)
                : "memory");
          }
        }
      }
    }
    asm volatile("{\n\t.reg .pred e;\n\telect.sync _|e, 0xffffffff;\n\t@e tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n\t}\n" ::"r"(
                     smem_u32(&bar))
                 : "memory");
  }
  const uint32_t ok = bounded_wait(smem_u32(&bar), 0);
  if (!ok && tid == 0) atomicExch(status, 1);
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  if (do_dump && ok) {
    for (int c = 0; c < kN; c += 8) {
      uint32_t r[8];
      const uint32_t taddr = tmem + ((uint32_t)(warp * 32) << 16) + (uint32_t)c;
      asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                   : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
                   : "r"(taddr));
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
      for (int j = 0; j < 8; ++j) dump[(size_t)tid * kN + c + j] = __uint_as_float(r[j]);
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512));
}

int main() {
  std::vector<float> A(kM * kK), B(kN * kK);
  srand(9);
  for (auto& v : A) v = (float)(rand() % 17 - 8) * 0.125f;
  for (auto& v : B) v = (float)(rand() % 13 - 6) * 0.25f;
  std::vector<__half> hA(A.size()), hB(B.size());
  for (size_t i = 0; i < A.size(); ++i) hA[i] = __float2half_rn(A[i]);
  for (size_t i = 0; i < B.size(); ++i) hB[i] = __float2half_rn(B[i]);
  __half *dA, *dB;
  float* dD;
  int* dS;
  CK(cudaMalloc(&dA, hA.size() * 2));
  CK(cudaMalloc(&dB, hB.size() * 2));
  CK(cudaMalloc(&dD, (size_t)128 * kN * sizeof(float)));
  CK(cudaMalloc(&dS, sizeof(int)));
  CK(cudaMemcpy(dA, hA.data(), hA.size() * 2, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dB, hB.data(), hB.size() * 2, cudaMemcpyHostToDevice));
  const int smem = (kN + kM) * kK * 2;
  CK(cudaFuncSetAttribute(ts_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  std::vector<float> exact((size_t)kM * kN);
  for (int r = 0; r < kM; ++r)
    for (int n = 0; n < kN; ++n) {
      float s = 0;
      for (int k = 0; k < kK; ++k) s += A[r * kK + k] * B[n * kK + k];
      exact[(size_t)r * kN + n] = s;
    }
  for (int mode = 0; mode < 2; ++mode) {
    CK(cudaMemset(dD, 0, (size_t)128 * kN * sizeof(float)));
    CK(cudaMemset(dS, 0, sizeof(int)));
    ts_kernel<<<1, 128, smem>>>(dA, dB, dD, dS, 1, 1, mode);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("mode %d: %s\n", mode, cudaGetErrorString(e)); return 1; }
    int st = 0;
    CK(cudaMemcpy(&st, dS, sizeof(int), cudaMemcpyDeviceToHost));
    std::vector<float> D((size_t)128 * kN);
    CK(cudaMemcpy(D.data(), dD, D.size() * sizeof(float), cudaMemcpyDeviceToHost));
    int bad = 0;
    for (size_t i = 0; i < D.size(); ++i) bad += D[i] != exact[i];
    printf("M=128 N=256 K=16 kind::f16, A from %s: %s, %d of %zu elements differ from the exact product\n", mode ? "shared memory" : "TENSOR MEMORY",
           st ? "MMAs did not complete" : "completed", bad, D.size());
    int dev = 0, sms = 0;
    CK(cudaGetDevice(&dev));
    CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    const int reps = 4000;
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    ts_kernel<<<sms, 128, smem>>>(dA, dB, dD, dS, 10, 0, mode);
    CK(cudaDeviceSynchronize());
    CK(cudaEventRecord(e0));
    ts_kernel<<<sms, 128, smem>>>(dA, dB, dD, dS, reps, 0, mode);
    CK(cudaEventRecord(e1));
    CK(cudaDeviceSynchronize());
    float ms = 0;
    CK(cudaEventElapsedTime(&ms, e0, e1));
    const double mmas = (double)sms * reps * (kK / 16) * 3;
    printf("  rate on %d SMs, 3 MMAs per K step: %.1f TFLOP/s of MMA work (%.1f ns per MMA per SM)\n", sms,
           mmas * 2.0 * kM * kN * 16 / (ms * 1e-3) / 1e12, ms * 1e6 / (mmas / sms));
  }
  return 0;
}
