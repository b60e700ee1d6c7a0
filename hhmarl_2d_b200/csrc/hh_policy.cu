// hh_policy.cu -- fused forward of the sampler's four network chains (Fight1/Fight2 or Esc1/Esc2: actor and central
// critic of both policies, models/ac_models_hetero.py:86-103, 256-291, 368-404) for the one-token-per-sequence case of
// the rollout (seq_lens = [1] * B, where the 2-head attention is out_proj(v_proj(x))).
//
// Per chain and row:  x[57|66] -> H = tanh(x W1 + b1) [500] -> (Fight only) the last 100 / 150 columns pass through the
// folded attention matrix, residual add, L2 normalisation -> Z = tanh(in Ws + bs) with the process-wide SHARED_LAYER
// (ac_models_hetero.py:22-27) -> head (26 / 24 logits, or the value).  The four chains are independent given x
// (the attention matrix is block diagonal over the actor / critic halves), so the grid is (row tiles of 64) x 4.
//
// One CTA = 64 rows of one chain, 8 warps, the 64 x 500 activation tile resident in shared memory (fp32, row
// stride 516 so that the A-fragment loads of mma.m16n8k8 are bank-conflict free); every GEMM of the chain reads its A
// operand from that tile and its B operand (weights, zero-padded to multiples of 8 so that no bounds checks are
// needed, stored in MMA fragment order so that a warp's load is two full lines) straight from L2 into registers, one
// k-step ahead of the MMAs -- each weight element is used by exactly one warp of the CTA, for all four 16-row tiles.  (A 4-stage cp.async ring through shared memory was measured slower:
// the per-k-step barrier costs more than the L2 latency it hides, profiles/README.md.)  Tensor cores through mma.sync TF32 with the 3xTF32 split (a = a_hi + a_lo, b = b_hi + b_lo,
// a b ~ a_lo b_hi + a_hi b_lo + a_hi b_hi, fp32 accumulate), which reproduces fp32 GEMM results to ~1e-6
// (precision = 0, default); precision = 1 is plain TF32 (one MMA per product).
// What this replaces in the sampler: 9 cuBLAS SGEMMs + ~25 element-wise kernels per tick (profiles/r1l_launches_summary.md).
#include <cuda_runtime.h>
#include <stdint.h>

#include <string>

#include "../../include/hhmarl_b200.h"
#include "hh_policy_tc.h"

namespace hh {
namespace pf {

constexpr int TM = 64;        // rows per CTA
constexpr int LDA = 516;      // activation tile row stride (floats): 516 mod 32 = 4
constexpr int XLD = 100;      // input tile row stride: 100 mod 32 = 4, >= 72 (66 padded to 72)
constexpr int NP = 504;       // K extent of the 500-wide layers: padded to a multiple of 8
constexpr int NW = 512;       // N extent (row stride) of their weight matrices: 64 n-tiles, 8 per warp, no guards
#ifndef HH_PF_WARPS
#define HH_PF_WARPS 8
#endif
constexpr int kWarps = HH_PF_WARPS;   // 8 or 16: each warp owns 64 / kWarps of the 64 n-tiles of a 500-wide layer
constexpr int kThreads = 32 * kWarps;
constexpr int NTW = 64 / kWarps;                  // n-tiles per warp, wide layers
constexpr int NTA = (19 + kWarps - 1) / kWarps;   // attention block: <= 19 n-tiles

constexpr int kMaxChains = 8;
struct Chain {
  const float *x, *w1, *b1, *watt, *batt, *ws, *bs, *wh, *bh;
  float* out;                 // logits / value, may be null when act_out is given
  const int* rows;            // optional gather: local row r is global row rows[begin + r] (input AND output)
  const int* range_dev;       // optional {begin, count} in device memory (data-dependent row lists without a host sync)
  int* act_out;               // optional per-head argmax, int32 [.., 4]
  int n_rows, ldx, d_in, k1_pad, att_lo, att_n, att_pad, n_out, ld_out, n_heads, head[4], ld_act;
};
struct Args {
  Chain c[kMaxChains];
};

// TF32 operands are the upper 19 bits of an fp32 word; the tensor core ignores the rest, i.e. passing the raw bits
// truncates.  (cvt.rna.tf32.f32 is emulated with ~5 ALU instructions on sm_100a -- it was 45 % of this kernel's
// instruction stream, profiles/r1m_*.)  Split for 3xTF32: hi = x with the low 13 mantissa bits cleared, lo = x - hi
// (exact); lo is truncated by the hardware, which leaves a relative error of <= 2^-20 per operand.
__device__ __forceinline__ uint32_t f2tf32(float x) { return __float_as_uint(x); }
__device__ __forceinline__ void split_tf32(float x, uint32_t& hi, uint32_t& lo) {
  hi = __float_as_uint(x) & 0xffffe000u;
  lo = __float_as_uint(x - __uint_as_float(hi));
}
// (tanh through the SFU -- 1 - 2 / (__expf(2x) + 1) -- was measured SLOWER than libdevice's tanhf here: 433 vs 419 us.)
__device__ __forceinline__ void mma_tf32(float (&d)[4], const uint32_t (&a)[4], const uint32_t (&b)[2]) {
  asm("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}

// acc[i][m] += A[64 x k_pad] (shared, row stride lda) x W[k_pad x 8 n_tiles] (global, row stride ldw) for the n-tiles
// warp + 8 i (i < NT) of this warp and the four 16-row tiles m.  Fragment layouts of mma.m16n8k8 (row.col):
//   A: a0 (g, t) a1 (g + 8, t) a2 (g, t + 4) a3 (g + 8, t + 4);  B: b0 (k = t, n = g) b1 (k = t + 4, n = g);
//   C: c0 (g, 2t) c1 (g, 2t + 1) c2 (g + 8, 2t) c3 (g + 8, 2t + 1);   g = lane / 4, t = lane % 4.
// acc[i][m] += A[64 x k_pad] (shared, row stride lda) x W[k_pad x 8 n_tiles] (global, row stride ldw) for the n-tiles
// warp + kWarps i (i < NT) of this warp and the four 16-row tiles m.  Fragment layouts of mma.m16n8k8 (row.col):
//   A: a0 (g, t) a1 (g + 8, t) a2 (g, t + 4) a3 (g + 8, t + 4);  B: b0 (k = t, n = g) b1 (k = t + 4, n = g);
//   C: c0 (g, 2t) c1 (g, 2t + 1) c2 (g + 8, 2t) c3 (g + 8, 2t + 1);   g = lane / 4, t = lane % 4.
template <int NT, bool X3, bool GUARD>
__device__ __forceinline__ void gemm_tile(const float* __restrict__ A, int lda, int k_pad, const float* __restrict__ W,
                                          int n_tiles, float (&acc)[NT][4][4]) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, g = lane >> 2, t = lane & 3;
#pragma unroll
  for (int i = 0; i < NT; ++i)
#pragma unroll
    for (int m = 0; m < 4; ++m)
#pragma unroll
      for (int q = 0; q < 4; ++q) acc[i][m][q] = 0.0f;
  bool on[NT];
#pragma unroll
  for (int i = 0; i < NT; ++i) on[i] = !GUARD || warp + kWarps * i < n_tiles;
  // W is stored in FRAGMENT ORDER (fused_forward.FusedPolicyPair): [k-step][n-tile][lane][2] = the (b0, b1) pair of
  // every lane of the 8 x 8 block, so one LDG.64 per n-tile reads two full 128-byte lines.  (Row-major weights cost
  // four lines = four L1 wavefronts per 4-byte load and made the L1 the bottleneck: 34 % tensor-pipe use, r1m.)
  const float2* wp = reinterpret_cast<const float2*>(W) + (size_t)warp * 32 + lane;
  const size_t w_step = (size_t)n_tiles * 32;     // float2 elements per k-step
  const float* ap = A + (size_t)g * lda + t;
  const int n_steps = k_pad >> 3;
  // The weights of k-step j + 2 are requested while k-step j computes: a three-slot register ring (the loop is
  // unrolled by three so that the slots are compile-time registers).  One k-step of lead was not enough to cover the
  // L2 round trip with two warps per scheduler (long-scoreboard was the top stall, profiles/r1m_*).
  float2 bq[3][NT];
#pragma unroll
  for (int s = 0; s < 2; ++s)
#pragma unroll
    for (int i = 0; i < NT; ++i)
      bq[s][i] = (on[i] && s < n_steps) ? __ldg(wp + (size_t)s * w_step + 32 * kWarps * i) : make_float2(0.0f, 0.0f);
#pragma unroll 1
  for (int j = 0; j < n_steps; j += 3) {
#pragma unroll
    for (int s = 0; s < 3; ++s) {
      if (j + s < n_steps) {           // warp-uniform
        if (j + s + 2 < n_steps) {
#pragma unroll
          for (int i = 0; i < NT; ++i)
            bq[(s + 2) % 3][i] = on[i] ? __ldg(wp + (size_t)(j + s + 2) * w_step + 32 * kWarps * i) : make_float2(0.0f, 0.0f);
        }
        uint32_t bh[NT][2], bl[NT][2], ah[4][4], al[4][4];
        const float* apk = ap + (j + s) * 8;
#pragma unroll
        for (int m = 0; m < 4; ++m) {
          const float a0 = apk[m * 16 * lda], a1 = apk[(m * 16 + 8) * lda], a2 = apk[m * 16 * lda + 4],
                      a3 = apk[(m * 16 + 8) * lda + 4];
          if (X3) {
            split_tf32(a0, ah[m][0], al[m][0]);
            split_tf32(a1, ah[m][1], al[m][1]);
            split_tf32(a2, ah[m][2], al[m][2]);
            split_tf32(a3, ah[m][3], al[m][3]);
          } else {
            ah[m][0] = f2tf32(a0); ah[m][1] = f2tf32(a1); ah[m][2] = f2tf32(a2); ah[m][3] = f2tf32(a3);
          }
        }
#pragma unroll
        for (int i = 0; i < NT; ++i) {
          if (X3) {
            split_tf32(bq[s][i].x, bh[i][0], bl[i][0]);
            split_tf32(bq[s][i].y, bh[i][1], bl[i][1]);
          } else {
            bh[i][0] = f2tf32(bq[s][i].x);
            bh[i][1] = f2tf32(bq[s][i].y);
          }
        }
        // small terms first; consecutive MMAs hit different accumulators (4 NT independent chains per pass)
        if (X3) {
#pragma unroll
          for (int m = 0; m < 4; ++m)
#pragma unroll
            for (int i = 0; i < NT; ++i)
              if (!GUARD || on[i]) mma_tf32(acc[i][m], al[m], bh[i]);
#pragma unroll
          for (int m = 0; m < 4; ++m)
#pragma unroll
            for (int i = 0; i < NT; ++i)
              if (!GUARD || on[i]) mma_tf32(acc[i][m], ah[m], bl[i]);
        }
#pragma unroll
        for (int m = 0; m < 4; ++m)
#pragma unroll
          for (int i = 0; i < NT; ++i)
            if (!GUARD || on[i]) mma_tf32(acc[i][m], ah[m], bh[i]);
      }
    }
  }
}

// store f(acc + bias) of this warp's tiles into the activation tile at column offset col0 (columns < n_cols only)
template <int NT, bool TANH>
__device__ __forceinline__ void store_tile(float* __restrict__ S, int col0, int n_cols, const float* __restrict__ bias,
                                           int n_tiles, const float (&acc)[NT][4][4], bool add_residual) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, g = lane >> 2, t = lane & 3;
#pragma unroll
  for (int i = 0; i < NT; ++i) {
    if (warp + kWarps * i >= n_tiles) continue;
    const int c = (warp + kWarps * i) * 8 + 2 * t;
#pragma unroll
    for (int m = 0; m < 4; ++m)
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const int col = c + (q & 1), row = m * 16 + g + ((q >> 1) ? 8 : 0);
        if (col < n_cols) {
          float v = acc[i][m][q] + __ldg(bias + col);
          float* p = S + (size_t)row * LDA + col0 + col;
          if (add_residual) v += *p;
          *p = TANH ? tanhf(v) : v;
        }
      }
  }
}

template <bool X3>
__global__ void __launch_bounds__(kThreads, 1) policy_forward_kernel(Args args) {
  extern __shared__ __align__(16) float smem[];
  float* act = smem;                    // [TM][LDA] activation tile
  float* xin = smem + TM * LDA;         // [TM][XLD] input tile; later the staged logits [TM][33]
  __shared__ int rowmap[TM];            // global row of every tile row, -1 beyond the chain's row list
  const Chain& C = args.c[blockIdx.y];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, g = lane >> 2, t = lane & 3;
  int beg = 0, cnt = C.n_rows;
  if (C.range_dev) {
    beg = C.range_dev[0];
    cnt = C.range_dev[1];
  }
  const int row0 = blockIdx.x * TM;
  if (row0 >= cnt) return;              // CTA-uniform
  if (tid < TM) {
    const int lr = row0 + tid;
    rowmap[tid] = lr < cnt ? (C.rows ? C.rows[beg + lr] : beg + lr) : -1;
  }
  __syncthreads();

  // input rows (zero beyond d_in and beyond the row list); the activation tile's padding columns start as zero
  for (int i = tid; i < TM * XLD; i += kThreads) {
    const int r = i / XLD, c = i - r * XLD;
    const int gr = rowmap[r];
    xin[i] = (c < C.d_in && gr >= 0) ? __ldg(C.x + (size_t)gr * C.ldx + c) : 0.0f;
  }
  for (int i = tid; i < TM * (LDA - 500); i += kThreads) {
    const int r = i / (LDA - 500), c = 500 + i - r * (LDA - 500);
    act[r * LDA + c] = 0.0f;
  }
  __syncthreads();

  {  // H = tanh(x W1 + b1)
    float acc[NTW][4][4];
    gemm_tile<NTW, X3, false>(xin, XLD, C.k1_pad, C.w1, NW / 8, acc);
    store_tile<NTW, true>(act, 0, 500, C.b1, NW / 8, acc, false);
  }
  __syncthreads();

  if (C.att_n > 0) {  // single-token attention: r = full + (full Wa + ba), then L2-normalise the block
    float acc[NTA][4][4];
    const int n_tiles = C.att_pad / 8;
    gemm_tile<NTA, X3, true>(act + C.att_lo, LDA, C.att_pad, C.watt, n_tiles, acc);
    __syncthreads();                     // every warp has read the block before anyone overwrites it
    store_tile<NTA, false>(act, C.att_lo, C.att_n, C.batt, n_tiles, acc, true);
    __syncthreads();
    for (int r = warp * (TM / kWarps); r < (warp + 1) * (TM / kWarps); ++r) {   // F.normalize: x / max(||x||_2, 1e-12)
      float* p = act + (size_t)r * LDA + C.att_lo;
      float ss = 0.0f;
      for (int c = lane; c < C.att_n; c += 32) ss += p[c] * p[c];
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, o);
      const float inv = 1.0f / fmaxf(sqrtf(ss), 1e-12f);
      for (int c = lane; c < C.att_n; c += 32) p[c] *= inv;
    }
    __syncthreads();
  }

  {  // Z = tanh(in Ws + bs), written back in place once every warp is done reading
    float acc[NTW][4][4];
    gemm_tile<NTW, X3, false>(act, LDA, NP, C.ws, NW / 8, acc);
    __syncthreads();
    store_tile<NTW, true>(act, 0, 500, C.bs, NW / 8, acc, false);
  }
  __syncthreads();

  {  // head: logits or value (+ optional per-head argmax, env_base.py:373-382)
    float acc[1][4][4];
    gemm_tile<1, X3, true>(act, LDA, NP, C.wh, 4, acc);
    float* lg = xin;                     // [TM][33]
    if (warp < 4) {
      const int c = warp * 8 + 2 * t;
#pragma unroll
      for (int m = 0; m < 4; ++m)
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const int col = c + (q & 1), r = m * 16 + g + ((q >> 1) ? 8 : 0);
          const int gr = rowmap[r];
          if (col < C.n_out) {
            const float v = acc[0][m][q] + __ldg(C.bh + col);
            if (gr >= 0 && C.out) C.out[(size_t)gr * C.ld_out + col] = v;
            lg[r * 33 + col] = v;
          }
        }
    }
    if (C.act_out) {
      __syncthreads();
      if (tid < TM && rowmap[tid] >= 0) {
        const float* row = lg + tid * 33;
        int4 a = make_int4(0, 0, 0, 0);
        int o = 0;
#pragma unroll
        for (int h = 0; h < 4; ++h) {
          if (h < C.n_heads) {
            int best = 0;
            float bv = row[o];
            for (int k = 1; k < C.head[h]; ++k)
              if (row[o + k] > bv) { bv = row[o + k]; best = k; }      // first maximum, like torch.argmax
            (h == 0 ? a.x : h == 1 ? a.y : h == 2 ? a.z : a.w) = best;
            o += C.head[h];
          }
        }
        reinterpret_cast<int4*>(C.act_out)[(size_t)rowmap[tid] * C.ld_act] = a;
      }
    }
  }
}

}  // namespace pf
}  // namespace hh

namespace hh {
namespace pf {
// Row lists per key without a host round trip (level 5: which frozen policy set an arena's opponents use,
// env_hetero.py:55-59): rows[k][0 .. count_k) = the arenas whose key equals keys[k], ranges[k] = {k n, count_k}.
// The order inside a list is whatever the atomics give; every row's result is independent of its neighbours.
__global__ void rows_init_kernel(int n, int n_keys, int* __restrict__ ranges) {
  const int k = threadIdx.x;
  if (k < n_keys) {
    ranges[2 * k] = k * n;
    ranges[2 * k + 1] = 0;
  }
}
__global__ void rows_fill_kernel(int n, int n_keys, int4 keys, const uint8_t* __restrict__ key, int* __restrict__ rows,
                                 int* __restrict__ ranges) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int v = key[i];
  const int k = v == keys.x ? 0 : (v == keys.y && n_keys > 1) ? 1 : (v == keys.z && n_keys > 2) ? 2 : (v == keys.w && n_keys > 3) ? 3 : -1;
  if (k < 0) return;
  const int pos = atomicAdd(&ranges[2 * k + 1], 1);
  rows[(size_t)k * n + pos] = i;
}
}  // namespace pf
}  // namespace hh

static thread_local std::string g_pf_error;
extern "C" const char* hh_policy_last_error(void) { return g_pf_error.c_str(); }

static int launch_policy(const hh::pf::Args& a, int n_chains, int max_rows, int precision, void* stream) {
  using namespace hh::pf;
  const size_t smem = sizeof(float) * (TM * LDA + TM * XLD);
  static bool opted_dev[64][2] = {};       // the opt-in is per device (one process may drive several)
  int dev = 0;
  cudaError_t ce = cudaGetDevice(&dev);
  if (ce != cudaSuccess || dev < 0 || dev >= 64) {
    g_pf_error = "hh_policy_forward: cudaGetDevice failed";
    return -2;
  }
  bool* opted = opted_dev[dev];
  if (!opted[precision]) {
    ce = precision == 0
             ? cudaFuncSetAttribute(policy_forward_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)
             : cudaFuncSetAttribute(policy_forward_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (ce != cudaSuccess) {
      g_pf_error = std::string("cudaFuncSetAttribute(policy_forward_kernel): ") + cudaGetErrorString(ce);
      return -2;
    }
    opted[precision] = true;
  }
  const dim3 grid((max_rows + TM - 1) / TM, n_chains);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (precision == 0) policy_forward_kernel<true><<<grid, kThreads, smem, st>>>(a);
  else policy_forward_kernel<false><<<grid, kThreads, smem, st>>>(a);
  ce = cudaGetLastError();
  if (ce != cudaSuccess) {
    g_pf_error = std::string("policy_forward_kernel launch: ") + cudaGetErrorString(ce);
    return -2;
  }
  return 0;
}

static bool chain_ok(const hh_policy_chain_ex& s, bool tc) {
  if (!s.x || !s.b1 || !s.bs || !s.bh || (!s.out && !s.act_out)) return false;
  if (!tc && (!s.w1 || !s.ws || !s.wh)) return false;
  if (s.n_rows < 0 || s.d_in <= 0 || s.d_in > 72 || s.k1_pad % 8 || s.k1_pad < s.d_in || s.k1_pad > 72) return false;
  if (s.n_out <= 0 || s.n_out > 32) return false;
  if (s.att_n > 0 && ((!tc && !s.watt) || !s.batt || s.att_pad % 8 || s.att_pad < s.att_n || s.att_lo + s.att_n != 500 || s.att_pad > 152))
    return false;
  if (s.act_out) {
    if (s.n_heads < 1 || s.n_heads > 4) return false;
    int tot = 0;
    for (int h = 0; h < s.n_heads; ++h) {
      if (s.head[h] < 1) return false;
      tot += s.head[h];
    }
    if (tot != s.n_out) return false;
  }
  return true;
}

extern "C" int hh_policy_forward_ex(int32_t n_chains, const hh_policy_chain_ex* chains, int32_t precision, void* stream) {
  using namespace hh::pf;
  if (n_chains <= 0 || n_chains > kMaxChains || !chains || precision < 0 || precision > 2) {
    g_pf_error = "hh_policy_forward_ex: bad argument";
    return -1;
  }
  if (precision == 2) {   // tcgen05 / TMEM path (hh_policy_tc.cu)
    int rows = 0;
    for (int i = 0; i < n_chains; ++i) {
      if (!chain_ok(chains[i], true)) {
        g_pf_error = "hh_policy_forward_ex: inconsistent chain description";
        return -1;
      }
      if (chains[i].n_rows > rows) rows = chains[i].n_rows;
    }
    if (rows == 0) return 0;
    return hh_pf_tc_launch(chains, n_chains, rows, stream, g_pf_error);
  }
  Args a;
  int max_rows = 0;
  for (int i = 0; i < n_chains; ++i) {
    const hh_policy_chain_ex& s = chains[i];
    if (!chain_ok(s, false)) {
      g_pf_error = "hh_policy_forward_ex: inconsistent chain description";
      return -1;
    }
    Chain& c = a.c[i];
    c.x = s.x; c.w1 = s.w1; c.b1 = s.b1; c.watt = s.watt; c.batt = s.batt; c.ws = s.ws; c.bs = s.bs; c.wh = s.wh; c.bh = s.bh;
    c.out = s.out; c.rows = s.rows; c.range_dev = s.range_dev; c.act_out = s.act_out;
    c.n_rows = s.n_rows; c.ldx = s.ldx; c.d_in = s.d_in; c.k1_pad = s.k1_pad; c.att_lo = s.att_lo; c.att_n = s.att_n;
    c.att_pad = s.att_pad; c.n_out = s.n_out; c.ld_out = s.ld_out; c.n_heads = s.n_heads;
    for (int h = 0; h < 4; ++h) c.head[h] = s.head[h];
    c.ld_act = s.ld_act > 0 ? s.ld_act : 1;
    if (s.n_rows > max_rows) max_rows = s.n_rows;   // with range_dev, n_rows is the capacity of the row list
  }
  if (max_rows == 0) return 0;
  return launch_policy(a, n_chains, max_rows, precision, stream);
}

extern "C" int hh_policy_forward(int32_t n_rows, const hh_policy_chain* chains, const float* ws_dev, const float* bs_dev,
                                 int32_t precision, void* stream) {
  if (n_rows <= 0 || !chains || !ws_dev || !bs_dev || (precision != 0 && precision != 1)) {
    g_pf_error = "hh_policy_forward: bad argument";
    return -1;
  }
  hh_policy_chain_ex ex[4];
  for (int i = 0; i < 4; ++i) {
    const hh_policy_chain& s = chains[i];
    hh_policy_chain_ex& c = ex[i];
    c = hh_policy_chain_ex{};
    c.x = s.x; c.w1 = s.w1; c.b1 = s.b1; c.watt = s.watt; c.batt = s.batt; c.ws = ws_dev; c.bs = bs_dev; c.wh = s.wh; c.bh = s.bh;
    c.out = s.out;
    c.n_rows = n_rows; c.ldx = s.ldx; c.d_in = s.d_in; c.k1_pad = s.k1_pad; c.att_lo = s.att_lo; c.att_n = s.att_n;
    c.att_pad = s.att_pad; c.n_out = s.n_out; c.ld_out = s.ld_out;
  }
  return hh_policy_forward_ex(4, ex, precision, stream);
}

extern "C" int hh_policy_rows_by_key(int32_t n, const uint8_t* key_dev, int32_t n_keys, const int32_t* keys_host,
                                     int32_t* rows_dev, int32_t* ranges_dev, void* stream) {
  using namespace hh::pf;
  if (n <= 0 || !key_dev || n_keys < 1 || n_keys > 4 || !keys_host || !rows_dev || !ranges_dev) {
    g_pf_error = "hh_policy_rows_by_key: bad argument";
    return -1;
  }
  int4 keys = make_int4(keys_host[0], n_keys > 1 ? keys_host[1] : -1, n_keys > 2 ? keys_host[2] : -1, n_keys > 3 ? keys_host[3] : -1);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  rows_init_kernel<<<1, 32, 0, st>>>(n, n_keys, ranges_dev);
  rows_fill_kernel<<<(n + 255) / 256, 256, 0, st>>>(n, n_keys, keys, key_dev, rows_dev, ranges_dev);
  const cudaError_t ce = cudaGetLastError();
  if (ce != cudaSuccess) {
    g_pf_error = std::string("hh_policy_rows_by_key launch: ") + cudaGetErrorString(ce);
    return -2;
  }
  return 0;
}

extern "C" int hh_policy_pack(const float* w_dev, int32_t k_rows, int32_t n_cols, int32_t ldw, int32_t n_total, int32_t n_chunk,
                              int32_t row_shift, int32_t ksteps, int32_t kps, void* image_dev, float* unscale_dev, void* stream) {
  return hh_pf_tc_pack(w_dev, k_rows, n_cols, ldw, n_total, n_chunk, row_shift, ksteps, kps, image_dev, unscale_dev, stream, g_pf_error);
}
