// hh_policy_tc.cu -- the policy forward (models/ac_models_hetero.py:86-103, 256-291, 368-404; same chains as hh_policy.cu) on
// Blackwell's 5th-generation tensor cores: tcgen05.mma with the accumulators in tensor memory, operands in shared memory,
// weights streamed by the TMA engine's bulk copies.  precision = 2 of hh_policy_forward_ex; fp32-equivalent results.
//
// Arithmetic ("H3").  kind::f16 MMAs run at twice the kind::tf32 rate and their operands take half the shared memory, so
// every fp32 operand v is carried as TWO halves of a power-of-two multiple:  hi = rn_f16(2^s v),  lo = rn_f16(2^s v - hi)
// (hi + lo = 2^s v to 2^-22 relative; the prescale keeps lo in fp16's normal range), and a product is three MMAs into
// one fp32 accumulator:  A_lo B_hi + A_hi B_lo + A_hi B_hi.  Activations use s = 12 (|v| <= 1 after tanh / L2
// normalisation; observations are in [0, 1]); each weight matrix gets the s that puts max |w| in [2^13, 2^14)
// (hh_policy_pack, on the device).  The epilogue multiplies by 2^-(12 + s).  Measured (profiles/r2b_tcgen05_probe2.txt):
// max error 3.7e-7 on sums of magnitude 1 (a sequential fp32 fma loop: 2.4e-7); 349 TFLOP/s fp32-equivalent at M = 64.
//
// One CTA = 64 rows of one chain (a 128-row tile of hi + lo activations would need 258 KB; measured in r2a: an M = 64 MMA takes
// the same time as an M = 128 one, which is why the 2x of kind::f16 matters).  Shared memory: activation tile hi | lo as
// [64 x 512] halves in the K-major no-swizzle canonical layout (8-row x 16-byte core matrices: element (r, k) at
// (k / 8) * 1024 + r * 16 + (k % 8) * 2), input tile [64 x 80] the same way, a 4-slot ring of 16 KB weight stages.
// Warp roles: warp 0 streams the weight images (cp.async.bulk -> mbarrier), one thread of warp 1 issues every MMA, warps
// 2..9 are the epilogue (tcgen05.ld, bias, tanh / attention residual + L2 normalisation / head, hi + lo split, store back
// into the activation tile for the next layer).  A 500-wide layer is two N = 256 halves in two TMEM regions, so that
// the first half's epilogue overlaps the second half's MMAs and the next layer starts on the first half's columns.
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdlib.h>

#include <string>

#include "../../include/hhmarl_b200.h"
#include "hh_policy_tc.h"

namespace hh {
namespace tc {

constexpr int TM = 64;                     // rows per CTA
constexpr int KA = 512;                    // K extent of the activation tile (500 padded)
constexpr int KX = 80;                     // K extent of the input tile (<= 72 inputs, padded to a multiple of 16)
constexpr uint32_t ACT_BYTES = TM * KA * 2;
constexpr uint32_t X_BYTES = TM * KX * 2;
constexpr uint32_t STAGE_BYTES = 16384;
constexpr int NSTAGE = 4;
constexpr uint32_t OFF_ACT_HI = 0, OFF_ACT_LO = ACT_BYTES, OFF_X_HI = 2 * ACT_BYTES, OFF_X_LO = OFF_X_HI + X_BYTES;
constexpr uint32_t OFF_RING = OFF_X_LO + X_BYTES;
constexpr uint32_t SMEM_BYTES = OFF_RING + NSTAGE * STAGE_BYTES;      // 217 088
constexpr int kThreads = 320;
constexpr int kEpiThreads = 256;
constexpr float ACT_SCALE = 4096.0f, ACT_UNSCALE = 1.0f / 4096.0f;   // 2^12
constexpr float ACT_CLAMP = 15.99f;        // |2^12 v| must stay below fp16's 65504
constexpr uint32_t LBO_A = TM * 16, SBO = 128;
constexpr int kTmemCols = 512;

constexpr int kMaxSeg = 8;
struct Seg {                 // one run of MMAs over consecutive K steps into one accumulator region
  const uint8_t* w;          // packed weight image of this run (global)
  uint16_t n;                // MMA N
  uint16_t ksteps;           // K = 16 steps
  uint16_t kps;              // K steps per ring stage (stage bytes = kps * n * 64)
  uint16_t tmem_col;         // first accumulator column
  uint16_t a_k0;             // first K column of the A operand inside its tile (a multiple of 8; of 16 in the M = 128 form)
  uint8_t a_src;             // 0 = input tile, 1 = activation tile
  uint8_t first;             // the first MMA overwrites the accumulator
  uint8_t wait_act;          // activation barrier to wait for before the first MMA (0xff = none)
  uint8_t commit_acc;        // accumulator barrier to commit to after the last MMA (0xff = none)
  uint8_t wait_free;         // M = 128 form: "accumulator drained" barrier to wait for before the first MMA (0xff = none)
};
struct Chain {
  const float *x, *b1, *batt, *bs, *bh, *us_w1, *us_att, *us_ws, *us_wh;
  float* out;
  const int* rows;
  const int* range_dev;
  int* act_out;
  int n_rows, ldx, d_in, att_lo, att_n, att_pad, n_out, ld_out, n_heads, head[4], ld_act, n_seg;
  Seg seg[kMaxSeg];
};
struct Args {
  Chain c[8];
  unsigned long long* prof;   // optional clock64() stamps, 32 per CTA (hh_policy_tc_profile; null = off)
  int debug;                  // timing experiments only (results are garbage): 1 = no weight copies, 2 = no MMAs
  uint32_t zero;              // always 0; a value the compiler cannot fold (orders the epilogue math after the accumulator release)
};

// barrier slots
constexpr int B_FULL = 0, B_EMPTY = NSTAGE, B_ACC = 2 * NSTAGE, B_ACT = 2 * NSTAGE + 6, B_COUNT = 2 * NSTAGE + 6 + 6;
// accumulator barriers: 0 L1 half 0, 1 L1 half 1, 2 attention, 3 shared half 0, 4 shared half 1, 5 head
// activation barriers:  0 H[:, :256] ready, 1 H[:, 256:], 2 attention block rewritten, 3 Z[:, :256], 4 Z[:, 256:], 5 input tile

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t smem_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = (uint64_t)((saddr >> 4) & 0x3FFFu);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFFu) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFFu) << 32;
  d |= (uint64_t)1 << 46;        // descriptor version of sm_100; no swizzle, base offset 0
  return d;
}
// kind::f16 instruction descriptor: D = F32, A = B = F16, both K-major, N >> 3 at bit 17, M >> 4 at bit 24
__device__ __forceinline__ uint32_t instr_desc_f16(uint32_t M, uint32_t N) { return (1u << 4) | ((N >> 3) << 17) | ((M >> 4) << 24); }
__device__ __forceinline__ void umma_f16(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(tmem_d),
      "l"(da), "l"(db), "r"(idesc), "r"(accumulate)
      : "memory");
}
// The issuing WARP runs converged and one elected lane issues (elect.sync inside the asm): with warp-uniform operands the
// compiler keeps the descriptors in uniform registers.  (Issuing from `if (lane == 0)` made it wrap every UTCHMMA in an
// ELECT + 7x R2UR.BROADCAST waterfall loop -- ~130 cycles of the issuing thread per MMA, profiles/README.md round 2.)
__device__ __forceinline__ void umma_f16_elect(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p, q;\n\telect.sync _|q, 0xffffffff;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "@q tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(tmem_d),
      "l"(da), "l"(db), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma_commit_elect(uint32_t bar) {
  asm volatile("{\n\t.reg .pred q;\n\telect.sync _|q, 0xffffffff;\n\t"
               "@q tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n\t}\n" ::"r"(bar)
               : "memory");
}
// lean forms for the issue loops: descriptors as (low word, constant high word); only the low word (address, LBO) changes
constexpr uint32_t kDescHi = (SBO >> 4) | (1u << 14);       // stride byte offset 128, descriptor version 1, no swizzle
__device__ __forceinline__ uint32_t desc_lo(uint32_t saddr, uint32_t lbo_bytes) { return ((saddr >> 4) & 0x3FFFu) | ((lbo_bytes >> 4) << 16); }
template <int GROUP, bool ACC1>
__device__ __forceinline__ void umma_lean(uint32_t tmem_d, uint32_t a_lo, uint32_t b_lo, uint32_t idesc, uint32_t accumulate) {
  if (GROUP == 2) {
    if (ACC1)
      asm volatile("{\n\t.reg .pred p, q;\n\t.reg .b64 da, db;\n\tmov.b64 da, {%1, %3};\n\tmov.b64 db, {%2, %3};\n\telect.sync _|q, 0xffffffff;\n\t"
                   "setp.eq.u32 p, 0, 0;\n\t@q tcgen05.mma.cta_group::2.kind::f16 [%0], da, db, %4, p;\n\t}\n" ::"r"(tmem_d),
                   "r"(a_lo), "r"(b_lo), "r"(kDescHi), "r"(idesc)
                   : "memory");
    else
      asm volatile("{\n\t.reg .pred p, q;\n\t.reg .b64 da, db;\n\tmov.b64 da, {%1, %3};\n\tmov.b64 db, {%2, %3};\n\telect.sync _|q, 0xffffffff;\n\t"
                   "setp.ne.b32 p, %5, 0;\n\t@q tcgen05.mma.cta_group::2.kind::f16 [%0], da, db, %4, p;\n\t}\n" ::"r"(tmem_d),
                   "r"(a_lo), "r"(b_lo), "r"(kDescHi), "r"(idesc), "r"(accumulate)
                   : "memory");
  } else {
    if (ACC1)
      asm volatile("{\n\t.reg .pred p, q;\n\t.reg .b64 da, db;\n\tmov.b64 da, {%1, %3};\n\tmov.b64 db, {%2, %3};\n\telect.sync _|q, 0xffffffff;\n\t"
                   "setp.eq.u32 p, 0, 0;\n\t@q tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %4, p;\n\t}\n" ::"r"(tmem_d),
                   "r"(a_lo), "r"(b_lo), "r"(kDescHi), "r"(idesc)
                   : "memory");
    else
      asm volatile("{\n\t.reg .pred p, q;\n\t.reg .b64 da, db;\n\tmov.b64 da, {%1, %3};\n\tmov.b64 db, {%2, %3};\n\telect.sync _|q, 0xffffffff;\n\t"
                   "setp.ne.b32 p, %5, 0;\n\t@q tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %4, p;\n\t}\n" ::"r"(tmem_d),
                   "r"(a_lo), "r"(b_lo), "r"(kDescHi), "r"(idesc), "r"(accumulate)
                   : "memory");
  }
}
// One K = 16 step of the split product in ONE asm block: A_lo B_hi (accumulate flag), A_hi B_lo, A_hi B_hi.  The five
// descriptor words enter once (each vector -> uniform register move costs the issuing warp ~10 cycles; as three separate
// blocks that was ~100 cycles per MMA -- more than a cta_group::2 MMA takes to execute).
template <int GROUP>
__device__ __forceinline__ void umma3_lean(uint32_t tmem_d, uint32_t a_lo, uint32_t a_hi, uint32_t b_hi, uint32_t b_lo, uint32_t idesc,
                                           uint32_t accumulate) {
  if (GROUP == 2)
    asm volatile(
        "{\n\t.reg .pred p, t, q;\n\t.reg .b32 hi;\n\t.reg .b64 dal, dah, dbh, dbl;\n\t"
        "mov.u32 hi, 0x4008;\n\t"
        "mov.b64 dal, {%1, hi};\n\tmov.b64 dah, {%2, hi};\n\tmov.b64 dbh, {%3, hi};\n\tmov.b64 dbl, {%4, hi};\n\t"
        "elect.sync _|q, 0xffffffff;\n\tsetp.ne.b32 p, %6, 0;\n\tsetp.eq.u32 t, 0, 0;\n\t"
        "@q tcgen05.mma.cta_group::2.kind::f16 [%0], dal, dbh, %5, p;\n\t"
        "@q tcgen05.mma.cta_group::2.kind::f16 [%0], dah, dbl, %5, t;\n\t"
        "@q tcgen05.mma.cta_group::2.kind::f16 [%0], dah, dbh, %5, t;\n\t}\n" ::"r"(tmem_d),
        "r"(a_lo), "r"(a_hi), "r"(b_hi), "r"(b_lo), "r"(idesc), "r"(accumulate)
        : "memory");
  else
    asm volatile(
        "{\n\t.reg .pred p, t, q;\n\t.reg .b32 hi;\n\t.reg .b64 dal, dah, dbh, dbl;\n\t"
        "mov.u32 hi, 0x4008;\n\t"
        "mov.b64 dal, {%1, hi};\n\tmov.b64 dah, {%2, hi};\n\tmov.b64 dbh, {%3, hi};\n\tmov.b64 dbl, {%4, hi};\n\t"
        "elect.sync _|q, 0xffffffff;\n\tsetp.ne.b32 p, %6, 0;\n\tsetp.eq.u32 t, 0, 0;\n\t"
        "@q tcgen05.mma.cta_group::1.kind::f16 [%0], dal, dbh, %5, p;\n\t"
        "@q tcgen05.mma.cta_group::1.kind::f16 [%0], dah, dbl, %5, t;\n\t"
        "@q tcgen05.mma.cta_group::1.kind::f16 [%0], dah, dbh, %5, t;\n\t}\n" ::"r"(tmem_d),
        "r"(a_lo), "r"(a_hi), "r"(b_hi), "r"(b_lo), "r"(idesc), "r"(accumulate)
        : "memory");
}
static_assert(kDescHi == 0x4008u, "the descriptor's constant high word is spelled out in umma3_lean");
// warp-uniform wait: every lane polls, the loop condition is a vote -> control flow (and everything computed under it) stays
// warp-uniform for the compiler
__device__ __forceinline__ void mbar_wait_uniform(uint32_t bar, uint32_t parity) {
  for (uint32_t spin = 0;; ++spin) {
    uint32_t ok;
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}\n"
                 : "=r"(ok)
                 : "r"(bar), "r"(parity)
                 : "memory");
    if (__all_sync(0xffffffffu, ok)) break;
    if (spin > (1u << 28)) __trap();
  }
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void fence_tc_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void fence_tc_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
// Every wait is bounded: a protocol error traps (the launch fails with an error) instead of hanging the device.
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok = 0;
  for (uint32_t spin = 0; !ok; ++spin) {
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}\n"
                 : "=r"(ok)
                 : "r"(bar), "r"(parity)
                 : "memory");
    if (!ok && spin > (1u << 24)) __trap();
  }
}
// Busy poll (test_wait never suspends the thread): for the single-lane roles whose barriers are signalled from the peer
// CTA / the tensor cores, where the wake-up of a suspended try_wait is what the stage round trip waits for.
__device__ __forceinline__ void mbar_spin(uint32_t bar, uint32_t parity) {
  uint32_t ok = 0;
  for (uint32_t spin = 0; !ok; ++spin) {
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}\n"
                 : "=r"(ok)
                 : "r"(bar), "r"(parity)
                 : "memory");
    if (!ok && spin > (1u << 28)) __trap();
  }
}
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(src),
               "r"(bytes), "r"(bar)
               : "memory");
}
__device__ __forceinline__ void bulk_g2s_mc(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar, uint16_t mask) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1], %2, [%3], %4;" ::"r"(dst),
               "l"(src), "r"(bytes), "r"(bar), "h"(mask)
               : "memory");
}
__device__ __forceinline__ void umma_commit_mc(uint32_t bar, uint16_t mask) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(bar), "h"(mask)
               : "memory");
}
__device__ __forceinline__ void cluster_sync() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ uint32_t cluster_rank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void tmem_ld_16x256b_x4(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.16x256b.x4.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
        "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld_16x256b_x1(uint32_t taddr, uint32_t (&r)[4]) {
  asm volatile("tcgen05.ld.sync.aligned.16x256b.x1.b32 {%0, %1, %2, %3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
               : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// byte offset of element (row, k) inside a 64-row operand tile
__device__ __forceinline__ uint32_t canon(int row, int k) { return (uint32_t)((k >> 3) * (TM * 16) + row * 16 + (k & 7) * 2); }

// store the pair (v0, v1) = elements (row, k), (row, k + 1), k even, as hi / lo halves of 2^12 v
__device__ __forceinline__ void store_pair(uint8_t* hi_tile, uint8_t* lo_tile, int row, int k, float v0, float v1) {
  const float s0 = v0 * ACT_SCALE, s1 = v1 * ACT_SCALE;
  const __half2 h = __floats2half2_rn(s0, s1);
  const float2 hf = __half22float2(h);
  const __half2 l = __floats2half2_rn(s0 - hf.x, s1 - hf.y);
  const uint32_t o = canon(row, k);
  *reinterpret_cast<__half2*>(hi_tile + o) = h;
  *reinterpret_cast<__half2*>(lo_tile + o) = l;
}
__device__ __forceinline__ float2 load_pair(const uint8_t* hi_tile, const uint8_t* lo_tile, int row, int k) {
  const uint32_t o = canon(row, k);
  const float2 h = __half22float2(*reinterpret_cast<const __half2*>(hi_tile + o));
  const float2 l = __half22float2(*reinterpret_cast<const __half2*>(lo_tile + o));
  return make_float2((h.x + l.x) * ACT_UNSCALE, (h.y + l.y) * ACT_UNSCALE);
}

// tanh(x) from a = 2 log2(e) x:  1 - 2 / (2^a + 1), two SFU operations (abs error < 4e-7; saturates correctly at +-inf).
// -DHH_TC_EXACT_TANH selects libdevice's tanhf (about four times the instructions; the epilogue is issue-bound).
constexpr float kTwoLog2e = 2.8853900817779268f;
__device__ __forceinline__ float tanh_from_scaled(float a) {
#ifdef HH_TC_EXACT_TANH
  return tanhf(a * (1.0f / kTwoLog2e));
#else
  float e, r;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(a));
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(e + 1.0f));
  return fmaf(-2.0f, r, 1.0f);
#endif
}
// 2^12 tanh(x) from a = 2 log2(e) x in one fma after the two SFU operations (the activation prescale folded in)
__device__ __forceinline__ float tanh_scaled_4096(float a) {
#ifdef HH_TC_EXACT_TANH
  return ACT_SCALE * tanhf(a * (1.0f / kTwoLog2e));
#else
  float e, r;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(a));
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(e + 1.0f));
  return fmaf(-2.0f * ACT_SCALE, r, ACT_SCALE);
#endif
}
// Four at a time with ONE reciprocal: the conversions to fp16 share the 16-lane XU pipe with ex2 / rcp, which bounds the tanh
// epilogues (3 XU operations per element); 1 / a_i = (1 / (a0 a1 a2 a3)) * (the other three) trades 0.75 of them for ~3 FMA-pipe
// operations.  The exponent is clamped to 27 (tanh is 1.0f there already: identical results) so that the product stays finite.
__device__ __forceinline__ void tanh_scaled_4096_x4(float& x0, float& x1, float& x2, float& x3) {
#ifdef HH_TC_EXACT_TANH
  x0 = tanh_scaled_4096(x0); x1 = tanh_scaled_4096(x1); x2 = tanh_scaled_4096(x2); x3 = tanh_scaled_4096(x3);
#else
  float e0, e1, e2, e3, r;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e0) : "f"(fminf(x0, 27.0f)));
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e1) : "f"(fminf(x1, 27.0f)));
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e2) : "f"(fminf(x2, 27.0f)));
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e3) : "f"(fminf(x3, 27.0f)));
  const float a0 = e0 + 1.0f, a1 = e1 + 1.0f, a2 = e2 + 1.0f, a3 = e3 + 1.0f;
  const float p01 = a0 * a1, p23 = a2 * a3;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(p01 * p23));
  const float r01 = r * p23, r23 = r * p01;      // 1 / (a0 a1), 1 / (a2 a3)
  x0 = fmaf(-2.0f * ACT_SCALE, r01 * a1, ACT_SCALE);
  x1 = fmaf(-2.0f * ACT_SCALE, r01 * a0, ACT_SCALE);
  x2 = fmaf(-2.0f * ACT_SCALE, r23 * a3, ACT_SCALE);
  x3 = fmaf(-2.0f * ACT_SCALE, r23 * a2, ACT_SCALE);
#endif
}
__device__ __forceinline__ void split_pair(float v0, float v1, __half2& h, __half2& l) {
  const float s0 = v0 * ACT_SCALE, s1 = v1 * ACT_SCALE;
  h = __floats2half2_rn(s0, s1);
  const float2 hf = __half22float2(h);
  l = __floats2half2_rn(s0 - hf.x, s1 - hf.y);
}

// epilogue of one 256-column half of a 500-wide layer: act[:, k0 + c] = tanh(acc * us + bias[k0 + c]); this warp owns
// the 16 rows of its lane quadrant and 128 of the 256 columns (64 values per thread).  With hold_bar != 0 the values are
// computed first and stored only once that barrier has completed (the layer is rewritten in place: nothing may be
// stored before every MMA that reads the old tile has finished).
template <bool HOLD>
__device__ __forceinline__ void epi_tanh_half(uint8_t* smem, uint32_t tmem, int tmem_col, int k0, const float* __restrict__ bias,
                                              float us, int q, int part, int lane, uint32_t hold_bar) {
  const uint32_t taddr = tmem + ((uint32_t)(32 * q) << 16) + (uint32_t)(tmem_col + part * 128);
  const int rlo = 16 * q + (lane >> 2), cpair = 2 * (lane & 3);
  uint8_t *hi = smem + OFF_ACT_HI, *lo = smem + OFF_ACT_LO;
  const float us2 = us * kTwoLog2e;
  __half2 hh[32], ll[32];
  uint32_t r[2][16];
  tmem_ld_16x256b_x4(taddr, r[0]);
#pragma unroll
  for (int c = 0; c < 4; ++c) {
    tmem_ld_wait();
    if (c < 3) tmem_ld_16x256b_x4(taddr + 32 * (c + 1), r[(c + 1) & 1]);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int k = k0 + part * 128 + 32 * c + 8 * i + cpair;
      float2 b = __ldg(reinterpret_cast<const float2*>(bias + k));
      b.x *= kTwoLog2e;
      b.y *= kTwoLog2e;
      const uint32_t* rr = r[c & 1] + 4 * i;
      split_pair(tanh_from_scaled(fmaf(__uint_as_float(rr[0]), us2, b.x)), tanh_from_scaled(fmaf(__uint_as_float(rr[1]), us2, b.y)),
                 hh[8 * c + 2 * i], ll[8 * c + 2 * i]);
      split_pair(tanh_from_scaled(fmaf(__uint_as_float(rr[2]), us2, b.x)), tanh_from_scaled(fmaf(__uint_as_float(rr[3]), us2, b.y)),
                 hh[8 * c + 2 * i + 1], ll[8 * c + 2 * i + 1]);
    }
  }
  if (HOLD) mbar_wait(hold_bar, 0);
#pragma unroll
  for (int c = 0; c < 4; ++c)
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int k = k0 + part * 128 + 32 * c + 8 * i + cpair;
      const uint32_t o = canon(rlo, k);
      *reinterpret_cast<__half2*>(hi + o) = hh[8 * c + 2 * i];
      *reinterpret_cast<__half2*>(lo + o) = ll[8 * c + 2 * i];
      *reinterpret_cast<__half2*>(hi + o + 128) = hh[8 * c + 2 * i + 1];     // row + 8
      *reinterpret_cast<__half2*>(lo + o + 128) = ll[8 * c + 2 * i + 1];
    }
}

// CS = CTAs per cluster: the CS row tiles of a cluster belong to the same chain and consume the same weight stream, so
// each CTA fetches 1 / CS of every stage and multicasts it into all of them (the L2 reads drop by CS; a slot is reused
// once the MMAs of ALL CS CTAs have read it).
template <int CS>
__global__ void __launch_bounds__(kThreads, 1) policy_forward_tc_kernel(const __grid_constant__ Args args) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ __align__(8) uint64_t bars[B_COUNT];
  __shared__ uint32_t tmem_base_s;
  __shared__ int rowmap[TM];
  __shared__ float ssum[2][TM];
  const Chain& C = args.c[blockIdx.y];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  int beg = 0, cnt = C.n_rows;
  if (C.range_dev) {
    beg = C.range_dev[0];
    cnt = C.range_dev[1];
  }
  const int row0 = blockIdx.x * TM;
  if ((int)(blockIdx.x / CS) * CS * TM >= cnt) return;   // cluster-uniform: a CTA without rows still relays its share of the weights
  unsigned long long* prof = args.prof ? args.prof + 32 * (size_t)(blockIdx.y * gridDim.x + blockIdx.x) : nullptr;
  constexpr uint16_t kMask = (uint16_t)((1u << CS) - 1u);
  if (tid < TM) {
    const int lr = row0 + tid;
    rowmap[tid] = lr < cnt ? (C.rows ? C.rows[beg + lr] : beg + lr) : -1;
  }
  const uint32_t bar0 = smem_u32(bars);
  if (tid == 0) {
    for (int i = 0; i < 2 * NSTAGE + 6; ++i) mbar_init(bar0 + 8 * i, (i >= B_EMPTY && i < B_EMPTY + NSTAGE) ? CS : 1);
    for (int i = 0; i < 6; ++i) mbar_init(bar0 + 8 * (B_ACT + i), kEpiThreads);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)), "r"(kTmemCols));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  fence_tc_before();
  __syncthreads();
  fence_tc_after();
  if (CS > 1) cluster_sync();           // every CTA's barriers are initialised before a peer can signal them
  const uint32_t tmem = tmem_base_s;

  if (warp == 0) {
    if (lane == 0) {   // ---- weight stream: the stages of every segment, in the order the MMAs consume them
      uint32_t slot = 0, phase = 0;
      const uint32_t rank = CS > 1 ? cluster_rank() : 0u;
      const uint32_t ring = smem_u32(smem + OFF_RING);
      long long stall = 0;
      for (int s = 0; s < C.n_seg; ++s) {
        const Seg& g = C.seg[s];
        const uint32_t bytes = (uint32_t)g.kps * g.n * 64u, part = bytes / CS;
        const uint8_t* src = g.w + rank * part;
        const int n_stage = g.ksteps / g.kps;
        for (int k = 0; k < n_stage; ++k) {
          const long long t0 = prof ? clock64() : 0;
          mbar_wait(bar0 + 8 * (B_EMPTY + slot), phase ^ 1);
          if (prof) stall += clock64() - t0;
          mbar_expect_tx(bar0 + 8 * (B_FULL + slot), bytes);
          if (CS == 1) bulk_g2s(ring + slot * STAGE_BYTES, src, bytes, bar0 + 8 * (B_FULL + slot));
          else bulk_g2s_mc(ring + slot * STAGE_BYTES + rank * part, src, part, bar0 + 8 * (B_FULL + slot), kMask);
          src += bytes;
          if (++slot == NSTAGE) {
            slot = 0;
            phase ^= 1;
          }
        }
      }
      if (prof) {
        prof[28] = (unsigned long long)stall;
        prof[29] = (unsigned long long)clock64();
      }
    }
  } else if (warp == 1) {
    {   // ---- MMA issue by the converged warp (one elected lane inside the asm).  Descriptors advance by adding to their 14-bit address field (16-byte units).
      uint32_t slot = 0, phase = 0;
      long long stall = 0;
      if (prof && lane == 0) prof[0] = (unsigned long long)clock64();
      mbar_wait_uniform(bar0 + 8 * (B_ACT + 5), 0);   // the input tile is in shared memory
      for (int s = 0; s < C.n_seg; ++s) {
        const Seg& g = C.seg[s];
        const uint32_t n = g.n, kps = g.kps, n_stage = g.ksteps / kps;
        if (g.wait_act != 0xff) mbar_wait_uniform(bar0 + 8 * (B_ACT + g.wait_act), 0);
        if (prof && lane == 0) prof[1 + 2 * s] = (unsigned long long)clock64();
        fence_tc_after();
        const uint32_t idesc = instr_desc_f16(TM, n);
        const uint32_t a_addr = smem_u32(smem + (g.a_src ? OFF_ACT_HI : OFF_X_HI)) + (uint32_t)(g.a_k0 >> 3) * (TM * 16);
        uint32_t a_hi = desc_lo(a_addr, LBO_A), a_lo = desc_lo(a_addr + (g.a_src ? ACT_BYTES : X_BYTES), LBO_A);
        const uint32_t b_seg = desc_lo(smem_u32(smem + OFF_RING), 16u * n);   // leading byte offset of B = 16 n bytes
        const uint32_t lo_off = 2u * n, step_off = 4u * n;                // B lo block, next K step (16-byte units)
        const uint32_t d = tmem + g.tmem_col;
        uint32_t acc = g.first ? 0u : 1u;
        for (uint32_t k = 0; k < n_stage; ++k) {
          const long long t0 = prof ? clock64() : 0;
          mbar_wait_uniform(bar0 + 8 * (B_FULL + slot), phase);
          if (prof) stall += clock64() - t0;
          fence_tc_after();
          uint32_t b = b_seg + slot * (STAGE_BYTES >> 4);
          for (uint32_t j = 0; j < kps; ++j) {
            umma3_lean<1>(d, a_lo, a_hi, b, b + lo_off, idesc, acc);   // small terms first
            acc = 1u;
            a_hi += (2 * LBO_A) >> 4;
            a_lo += (2 * LBO_A) >> 4;
            b += step_off;
          }
          // the slot is free once these MMAs have read it -- in every CTA of the cluster
          if (CS == 1) umma_commit_elect(bar0 + 8 * (B_EMPTY + slot));
          else umma_commit_mc(bar0 + 8 * (B_EMPTY + slot), kMask);
          if (++slot == NSTAGE) {
            slot = 0;
            phase ^= 1;
          }
        }
        if (g.commit_acc != 0xff) umma_commit_elect(bar0 + 8 * (B_ACC + g.commit_acc));
        if (prof && lane == 0) prof[2 + 2 * s] = (unsigned long long)clock64();
      }
      if (prof && lane == 0) prof[15] = (unsigned long long)stall;
    }
  } else {
    // ---- epilogue warps 2..9: lane quadrant q = warp % 4 (the TMEM lanes a warp can read), column part (warp - 2) / 4
    const int et = tid - 64;
    const int q = warp & 3, part = (warp - 2) >> 2;
    const int rlo = 16 * q + (lane >> 2), cpair = 2 * (lane & 3);
    uint8_t *hi = smem + OFF_ACT_HI, *lo = smem + OFF_ACT_LO;
    // input rows -> hi / lo halves of 2^12 x in the canonical layout (zero beyond d_in and beyond the row list)
    {
      constexpr int kIter = TM * (KX / 2) / kEpiThreads;     // 10 pairs per thread: all loads in flight before the first use
      float v0[kIter], v1[kIter];
#pragma unroll
      for (int it = 0; it < kIter; ++it) {
        const int i = et + it * kEpiThreads, r = i / (KX / 2), c = 2 * (i - r * (KX / 2));
        const int gr = rowmap[r];
        const float* xr = C.x + (size_t)(gr >= 0 ? gr : 0) * C.ldx;
        v0[it] = (gr >= 0 && c < C.d_in) ? __ldg(xr + c) : 0.0f;
        v1[it] = (gr >= 0 && c + 1 < C.d_in) ? __ldg(xr + c + 1) : 0.0f;
      }
#pragma unroll
      for (int it = 0; it < kIter; ++it) {
        const int i = et + it * kEpiThreads, r = i / (KX / 2), c = 2 * (i - r * (KX / 2));
        store_pair(smem + OFF_X_HI, smem + OFF_X_LO, r, c, fminf(fmaxf(v0[it], -ACT_CLAMP), ACT_CLAMP),
                   fminf(fmaxf(v1[it], -ACT_CLAMP), ACT_CLAMP));
      }
    }
    fence_async_smem();
    mbar_arrive(bar0 + 8 * (B_ACT + 5));
    {  // H = tanh(x W1 + b1)
      const float us = __ldg(C.us_w1);
      for (int h = 0; h < 2; ++h) {
        mbar_wait(bar0 + 8 * (B_ACC + h), 0);
        if (prof && et == 0) prof[16 + 2 * h] = (unsigned long long)clock64();
        fence_tc_after();
        epi_tanh_half<false>(smem, tmem, 256 * h, 256 * h, C.b1, us, q, part, lane, 0u);
        fence_async_smem();
        fence_tc_before();
        mbar_arrive(bar0 + 8 * (B_ACT + h));
        if (prof && et == 0) prof[17 + 2 * h] = (unsigned long long)clock64();
      }
    }
    if (C.att_n > 0) {   // single-token attention: r = full + (full Wa + ba), then L2-normalise the block
      const float us = __ldg(C.us_att);
      mbar_wait(bar0 + 8 * (B_ACC + 2), 0);
      if (prof && et == 0) prof[20] = (unsigned long long)clock64();
      fence_tc_after();
      const uint32_t taddr = tmem + ((uint32_t)(32 * q) << 16);
      const int n_chunk = (C.att_pad + 31) >> 5, c_mid = (n_chunk + 1) >> 1;       // 32-column chunks: part 0 the first half
      const int c_beg = part ? c_mid : 0, c_end = part ? n_chunk : c_mid;
      float v[3][16];
      float ss0 = 0.0f, ss1 = 0.0f;
#pragma unroll
      for (int ci = 0; ci < 3; ++ci) {
        const int chunk = c_beg + ci;
        if (chunk < c_end) {                                // warp-uniform
          uint32_t r[16];
          tmem_ld_16x256b_x4(taddr + 32 * chunk, r);
          tmem_ld_wait();
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const int col = 32 * chunk + 8 * i + cpair;
            if (col < C.att_n) {                            // att_n is even: a pair is valid or invalid as a whole
              const float2 b = __ldg(reinterpret_cast<const float2*>(C.batt + col));
              const float2 p0 = load_pair(hi, lo, rlo, C.att_lo + col), p1 = load_pair(hi, lo, rlo + 8, C.att_lo + col);
              v[ci][4 * i] = fmaf(__uint_as_float(r[4 * i]), us, b.x) + p0.x;
              v[ci][4 * i + 1] = fmaf(__uint_as_float(r[4 * i + 1]), us, b.y) + p0.y;
              v[ci][4 * i + 2] = fmaf(__uint_as_float(r[4 * i + 2]), us, b.x) + p1.x;
              v[ci][4 * i + 3] = fmaf(__uint_as_float(r[4 * i + 3]), us, b.y) + p1.y;
            } else {
              v[ci][4 * i] = v[ci][4 * i + 1] = v[ci][4 * i + 2] = v[ci][4 * i + 3] = 0.0f;
            }
            ss0 += v[ci][4 * i] * v[ci][4 * i] + v[ci][4 * i + 1] * v[ci][4 * i + 1];
            ss1 += v[ci][4 * i + 2] * v[ci][4 * i + 2] + v[ci][4 * i + 3] * v[ci][4 * i + 3];
          }
        }
      }
      ss0 += __shfl_xor_sync(0xffffffffu, ss0, 1);
      ss0 += __shfl_xor_sync(0xffffffffu, ss0, 2);
      ss1 += __shfl_xor_sync(0xffffffffu, ss1, 1);
      ss1 += __shfl_xor_sync(0xffffffffu, ss1, 2);
      if ((lane & 3) == 0) {
        ssum[part][rlo] = ss0;
        ssum[part][rlo + 8] = ss1;
      }
      asm volatile("bar.sync 1, %0;" ::"n"(kEpiThreads) : "memory");       // the eight epilogue warps
      const float inv0 = 1.0f / fmaxf(sqrtf(ssum[0][rlo] + ssum[1][rlo]), 1e-12f);          // F.normalize: x / max(||x||_2, 1e-12)
      const float inv1 = 1.0f / fmaxf(sqrtf(ssum[0][rlo + 8] + ssum[1][rlo + 8]), 1e-12f);
#pragma unroll
      for (int ci = 0; ci < 3; ++ci) {
        const int chunk = c_beg + ci;
        if (chunk < c_end) {
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const int col = 32 * chunk + 8 * i + cpair;
            if (col < C.att_n) {
              store_pair(hi, lo, rlo, C.att_lo + col, v[ci][4 * i] * inv0, v[ci][4 * i + 1] * inv0);
              store_pair(hi, lo, rlo + 8, C.att_lo + col, v[ci][4 * i + 2] * inv1, v[ci][4 * i + 3] * inv1);
            }
          }
        }
      }
      fence_async_smem();
      fence_tc_before();
      mbar_arrive(bar0 + 8 * (B_ACT + 2));
      if (prof && et == 0) prof[21] = (unsigned long long)clock64();
    }
    {  // Z = tanh(in Ws + bs), in place: the first half is computed while the second half's MMAs run and stored once they
       // have all read the tile
      const float us = __ldg(C.us_ws);
      mbar_wait(bar0 + 8 * (B_ACC + 3), 0);
      if (prof && et == 0) prof[22] = (unsigned long long)clock64();
      fence_tc_after();
      epi_tanh_half<true>(smem, tmem, 0, 0, C.bs, us, q, part, lane, bar0 + 8 * (B_ACC + 4));
      fence_async_smem();
      fence_tc_before();
      mbar_arrive(bar0 + 8 * (B_ACT + 3));
      if (prof && et == 0) prof[23] = (unsigned long long)clock64();
      fence_tc_after();
      epi_tanh_half<false>(smem, tmem, 256, 256, C.bs, us, q, part, lane, 0u);
      fence_async_smem();
      fence_tc_before();
      mbar_arrive(bar0 + 8 * (B_ACT + 4));
      if (prof && et == 0) prof[24] = (unsigned long long)clock64();
    }
    if (part == 0) {  // head: logits or value (+ optional per-head argmax, env_base.py:373-382)
      const float us = __ldg(C.us_wh);
      mbar_wait(bar0 + 8 * (B_ACC + 5), 0);
      if (prof && et == 0) prof[25] = (unsigned long long)clock64();
      fence_tc_after();
      float* lg = reinterpret_cast<float*>(smem + OFF_X_HI);       // [TM][33]; the input tile is dead by now
      uint32_t r[16];
      tmem_ld_16x256b_x4(tmem + ((uint32_t)(32 * q) << 16), r);
      tmem_ld_wait();
      const int g0 = rowmap[rlo], g1 = rowmap[rlo + 8];
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 2; ++j) {
          const int col = 8 * i + cpair + j;
          if (col < C.n_out) {
            const float b = __ldg(C.bh + col);
            const float v0 = fmaf(__uint_as_float(r[4 * i + j]), us, b), v1 = fmaf(__uint_as_float(r[4 * i + 2 + j]), us, b);
            if (C.out) {
              if (g0 >= 0) C.out[(size_t)g0 * C.ld_out + col] = v0;
              if (g1 >= 0) C.out[(size_t)g1 * C.ld_out + col] = v1;
            }
            lg[rlo * 33 + col] = v0;
            lg[(rlo + 8) * 33 + col] = v1;
          }
        }
      if (C.act_out) {
        __syncwarp();
        const int row = 16 * q + lane;
        if (lane < 16 && rowmap[row] >= 0) {
          const float* lr = lg + row * 33;
          int4 a = make_int4(0, 0, 0, 0);
          int o = 0;
#pragma unroll
          for (int h = 0; h < 4; ++h) {
            if (h < C.n_heads) {
              int best = 0;
              float bv = lr[o];
              for (int k = 1; k < C.head[h]; ++k)
                if (lr[o + k] > bv) { bv = lr[o + k]; best = k; }      // first maximum, like torch.argmax
              (h == 0 ? a.x : h == 1 ? a.y : h == 2 ? a.z : a.w) = best;
              o += C.head[h];
            }
          }
          reinterpret_cast<int4*>(C.act_out)[(size_t)rowmap[row] * C.ld_act] = a;
        }
      }
    }
    if (prof && et == 0) prof[26] = (unsigned long long)clock64();
  }
  fence_tc_before();
  __syncthreads();
  if (CS > 1) cluster_sync();           // no CTA leaves while a peer may still write into its shared memory
  if (warp == 1) {
    __syncwarp();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(kTmemCols));
  }
}

// =====================================================================================================================
// CTA-PAIR form (HH_TC_PAIR=1; see g_pair below for why it is not the default): two CTAs of a cluster = two 64-row tiles of the same chain share ONE weight stream.  Every
// MMA is a tcgen05.mma.cta_group::2 of M = 128 issued by the pair's CTA 0: each CTA supplies its own 64 activation rows
// and HALF of the weight columns of the step (its ring holds 8 KB per K step instead of 16), so the bytes an SM pulls
// from L2 per row halve -- the single-CTA kernel above is bound by exactly that stream -- and the pair's tensor cores run
// at the M = 128 rate (profiles/r2i_tcgen05_probe3.txt: 2.0x the M = 64 rate).  Accumulator of an N-wide step in each CTA's
// tensor memory: its row m in lane m for columns [0, N/2) and in lane 64 + m for columns [N/2, N) (N/2 TMEM columns).
// Protocol: each CTA streams its own half into its own ring; CTA 1 relays "stage landed" to CTA 0's full barrier (a bulk
// copy cannot complete on a remote barrier, probe3); CTA 0's commits are multicast to both CTAs' empty / accumulator
// barriers; both CTAs' epilogue warps arrive (one lane per warp) on CTA 0's activation barriers.
constexpr int NSTAGE2 = 8;
constexpr uint32_t STAGE2_BYTES = 8192;
static_assert(NSTAGE2 * STAGE2_BYTES == NSTAGE * STAGE_BYTES, "same ring footprint");
constexpr int P_FULL = 0, P_EMPTY = NSTAGE2, P_ACC = 2 * NSTAGE2, P_ACT = 2 * NSTAGE2 + 6, P_COUNT = 2 * NSTAGE2 + 12;
constexpr int kEpiWarps = kEpiThreads / 32;

__device__ __forceinline__ void umma2_f16(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(tmem_d),
      "l"(da), "l"(db), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma2_f16_elect(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p, q;\n\telect.sync _|q, 0xffffffff;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "@q tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(tmem_d),
      "l"(da), "l"(db), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma2_commit_both_elect(uint32_t bar) {
  asm volatile("{\n\t.reg .pred q;\n\telect.sync _|q, 0xffffffff;\n\t"
               "@q tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;\n\t}\n" ::"r"(bar),
               "h"((uint16_t)3)
               : "memory");
}
__device__ __forceinline__ void umma2_commit_both(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(bar),
               "h"((uint16_t)3)
               : "memory");
}
__device__ __forceinline__ uint32_t map_to_cta(uint32_t local_addr, uint32_t cta) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(local_addr), "r"(cta));
  return r;
}
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_addr) {
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
// the relay's arrive carries no data of its own (the stage was written by the TMA engine and is read by the tensor
// cores): no cluster-scope release fence on the signalling thread
__device__ __forceinline__ void mbar_arrive_cluster_relaxed(uint32_t cluster_addr) {
  asm volatile("mbarrier.arrive.relaxed.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
// wait that also acquires what a peer CTA released (its epilogue's shared-memory writes, its relay)
__device__ __forceinline__ void mbar_wait_cluster(uint32_t bar, uint32_t parity) {
  uint32_t ok = 0;
  for (uint32_t spin = 0; !ok; ++spin) {
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}\n"
                 : "=r"(ok)
                 : "r"(bar), "r"(parity)
                 : "memory");
    if (!ok && spin > (1u << 24)) __trap();
  }
}
__device__ __forceinline__ void tmem_ld_32x32b_x16(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
        "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld_32x32b_x8(uint32_t taddr, uint32_t (&r)[8]) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
               : "r"(taddr));
}
// this warp's part of a phase is done: its shared-memory writes are visible to the tensor cores, its TMEM reads are
// complete; one lane tells CTA 0
__device__ __forceinline__ void warp_arrive_leader(uint32_t leader_bar, int lane) {
  fence_async_smem();
  fence_tc_before();
  __syncwarp();
  if (lane == 0) mbar_arrive_cluster(leader_bar);
}

// epilogue of one N = 256 step of a 500-wide layer in the pair layout: the thread owns row 32 (q & 1) + lane and 64
// consecutive activation columns k0 + 128 (q >> 1) + 64 part + [0, 64)
template <bool HOLD>
__device__ __forceinline__ void epi2_tanh_half(uint8_t* smem, uint32_t tmem, int tmem_col, int k0, const float* __restrict__ bias,
                                               float us, int q, int part, int lane, uint32_t hold_bar) {
  const int row = 32 * (q & 1) + lane;
  const int kb = k0 + 128 * (q >> 1) + 64 * part;
  const uint32_t taddr = tmem + ((uint32_t)(32 * q) << 16) + (uint32_t)(tmem_col + 64 * part);
  const float us2 = us * kTwoLog2e;
  uint4 hh[8], ll[8];
  uint32_t r[2][16];
  tmem_ld_32x32b_x16(taddr, r[0]);
#pragma unroll
  for (int c = 0; c < 4; ++c) {
    tmem_ld_wait();
    if (c < 3) tmem_ld_32x32b_x16(taddr + 16 * (c + 1), r[(c + 1) & 1]);
#pragma unroll
    for (int g = 0; g < 2; ++g) {
      const int k = kb + 16 * c + 8 * g;
      const float4 b0 = __ldg(reinterpret_cast<const float4*>(bias + k)), b1 = __ldg(reinterpret_cast<const float4*>(bias + k + 4));
      const float bb[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
      __half2 h[4], l[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float v0 = tanh_from_scaled(fmaf(__uint_as_float(r[c & 1][8 * g + 2 * j]), us2, bb[2 * j] * kTwoLog2e));
        const float v1 = tanh_from_scaled(fmaf(__uint_as_float(r[c & 1][8 * g + 2 * j + 1]), us2, bb[2 * j + 1] * kTwoLog2e));
        split_pair(v0, v1, h[j], l[j]);
      }
      hh[2 * c + g] = make_uint4(*reinterpret_cast<uint32_t*>(&h[0]), *reinterpret_cast<uint32_t*>(&h[1]),
                                 *reinterpret_cast<uint32_t*>(&h[2]), *reinterpret_cast<uint32_t*>(&h[3]));
      ll[2 * c + g] = make_uint4(*reinterpret_cast<uint32_t*>(&l[0]), *reinterpret_cast<uint32_t*>(&l[1]),
                                 *reinterpret_cast<uint32_t*>(&l[2]), *reinterpret_cast<uint32_t*>(&l[3]));
    }
  }
  if (HOLD) mbar_wait(hold_bar, 0);
  uint8_t *hi = smem + OFF_ACT_HI, *lo = smem + OFF_ACT_LO;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const uint32_t o = (uint32_t)(((kb >> 3) + i) * (TM * 16) + row * 16);       // one core-matrix row = 8 columns = 16 bytes
    *reinterpret_cast<uint4*>(hi + o) = hh[i];
    *reinterpret_cast<uint4*>(lo + o) = ll[i];
  }
}

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kThreads, 1) policy_forward_pair_kernel(const __grid_constant__ Args args) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ __align__(8) uint64_t bars[P_COUNT];
  __shared__ uint32_t tmem_base_s;
  __shared__ int rowmap[TM];
  __shared__ float ssum[4][TM];
  const Chain& C = args.c[blockIdx.y];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  int beg = 0, cnt = C.n_rows;
  if (C.range_dev) {
    beg = C.range_dev[0];
    cnt = C.range_dev[1];
  }
  const int row0 = blockIdx.x * TM;
  if ((int)(blockIdx.x >> 1) * 2 * TM >= cnt) return;   // pair-uniform: a CTA without rows still supplies its half of the weights
  unsigned long long* prof = args.prof ? args.prof + 32 * (size_t)(blockIdx.y * gridDim.x + blockIdx.x) : nullptr;
  const uint32_t rank = cluster_rank();
  if (tid < TM) {
    const int lr = row0 + tid;
    rowmap[tid] = lr < cnt ? (C.rows ? C.rows[beg + lr] : beg + lr) : -1;
  }
  const uint32_t bar0 = smem_u32(bars);
  if (tid == 0) {
    for (int i = 0; i < NSTAGE2; ++i) {
      mbar_init(bar0 + 8 * (P_FULL + i), rank == 0 ? 2 : 1);     // CTA 0: its own copy + CTA 1's relay
      mbar_init(bar0 + 8 * (P_EMPTY + i), 1);
    }
    for (int i = 0; i < 6; ++i) {
      mbar_init(bar0 + 8 * (P_ACC + i), 1);
      mbar_init(bar0 + 8 * (P_ACT + i), 2 * kEpiWarps);          // one arrival per epilogue warp of both CTAs
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)), "r"(kTmemCols));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;");
  }
  fence_tc_before();
  __syncthreads();
  cluster_sync();                       // both CTAs' barriers and allocations exist before anyone signals a peer
  fence_tc_after();
  const uint32_t tmem = tmem_base_s;
  const uint32_t ring = smem_u32(smem + OFF_RING);

  if (warp == 0) {
    if (lane == 0) {   // ---- this CTA's half of the weight stream
      uint32_t slot = 0, phase = 0;
      long long stall = 0;
      for (int s = 0; s < C.n_seg; ++s) {
        const Seg& g = C.seg[s];
        const uint32_t bytes = (uint32_t)g.kps * g.n * 32u;                 // per CTA and stage
        const uint8_t* src = g.w + rank * bytes;
        const int n_stage = g.ksteps / g.kps;
        for (int k = 0; k < n_stage; ++k) {
          const long long t0 = prof ? clock64() : 0;
          if (!(slot & 1u)) mbar_spin(bar0 + 8 * (P_EMPTY + slot + 1), phase ^ 1);   // slots are released in pairs (odd slot's barrier)
          if (prof) stall += clock64() - t0;
          if (args.debug & 1) {
            mbar_arrive(bar0 + 8 * (P_FULL + slot));
          } else {
            mbar_expect_tx(bar0 + 8 * (P_FULL + slot), bytes);
            bulk_g2s(ring + slot * STAGE2_BYTES, src, bytes, bar0 + 8 * (P_FULL + slot));
          }
          src += 2 * bytes;
          if (++slot == NSTAGE2) {
            slot = 0;
            phase ^= 1;
          }
        }
      }
      if (prof) {
        prof[28] = (unsigned long long)stall;
        prof[29] = (unsigned long long)clock64();
      }
    }
  } else if (warp == 1) {
    if (rank == 1) {
     if (lane == 0) {   // ---- relay: "my half of stage k has landed" -> CTA 0's full barrier
      uint32_t slot = 0, phase = 0;
      long long stall = 0;
      const uint32_t leader_full = map_to_cta(bar0 + 8 * P_FULL, 0);
      if (prof) prof[0] = (unsigned long long)clock64();
      for (int s = 0; s < C.n_seg; ++s) {
        const int n_stage = C.seg[s].ksteps / C.seg[s].kps;
        if (prof) prof[1 + 2 * s] = (unsigned long long)clock64();
        for (int k = 0; k < n_stage; ++k) {
          const long long t0 = prof ? clock64() : 0;
          mbar_spin(bar0 + 8 * (P_FULL + slot), phase);
          if (prof) stall += clock64() - t0;
          mbar_arrive_cluster_relaxed(leader_full + 8 * slot);
          if (++slot == NSTAGE2) {
            slot = 0;
            phase ^= 1;
          }
        }
        if (prof) prof[2 + 2 * s] = (unsigned long long)clock64();
      }
      if (prof) prof[15] = (unsigned long long)stall;
     }
    } else {   // ---- MMA issue for the pair (CTA 0): the converged warp, one elected lane inside the asm
      uint32_t slot = 0, phase = 0;
      long long stall = 0, t_mma = 0, t_commit = 0;
      if (prof && lane == 0) prof[0] = (unsigned long long)clock64();
      mbar_wait_cluster(bar0 + 8 * (P_ACT + 5), 0);           // both input tiles are in shared memory
      for (int s = 0; s < C.n_seg; ++s) {
        const Seg& g = C.seg[s];
        const uint32_t n = g.n, kps = g.kps, n_stage = g.ksteps / kps;
        if (g.wait_act != 0xff) mbar_wait_cluster(bar0 + 8 * (P_ACT + g.wait_act), 0);
        if (prof && lane == 0) prof[1 + 2 * s] = (unsigned long long)clock64();
        fence_tc_after();
        const uint32_t idesc = instr_desc_f16(2 * TM, n);
        const uint32_t a_addr = smem_u32(smem + (g.a_src ? OFF_ACT_HI : OFF_X_HI)) + (uint32_t)(g.a_k0 >> 3) * (TM * 16);
        uint32_t a_hi = desc_lo(a_addr, LBO_A), a_lo = desc_lo(a_addr + (g.a_src ? ACT_BYTES : X_BYTES), LBO_A);
        const uint32_t b_seg = desc_lo(ring, 8u * n);                     // a CTA holds n / 2 weight columns: LBO = 16 (n / 2) bytes
        const uint32_t lo_off = n, step_off = 2u * n;                     // B lo block, next K step (16-byte units)
        const uint32_t d = tmem + g.tmem_col;
        uint32_t acc = g.first ? 0u : 1u;
        for (uint32_t k = 0; k < n_stage; ++k) {
          const long long t0 = prof ? clock64() : 0;
          mbar_wait_uniform(bar0 + 8 * (P_FULL + slot), phase);   // the tensor cores read the stage (async proxy): no thread-level acquire
          if (prof) stall += clock64() - t0;
          fence_tc_after();
          uint32_t b = b_seg + slot * (STAGE2_BYTES >> 4);
          const long long t1 = prof ? clock64() : 0;
          for (uint32_t j = 0; j < kps && !(args.debug & 2); ++j) {
            umma3_lean<2>(d, a_lo, a_hi, b, b + lo_off, idesc, acc);   // small terms first
            acc = 1u;
            a_hi += (2 * LBO_A) >> 4;
            a_lo += (2 * LBO_A) >> 4;
            b += step_off;
          }
          const long long t2 = prof ? clock64() : 0;
          // both rings' slots are free once these MMAs have read them.  Slots are released in PAIRS (one multicast commit
          // per two stages, on the odd slot's barrier, which the producers wait for before refilling the even slot)
          if (slot & 1u) umma2_commit_both_elect(bar0 + 8 * (P_EMPTY + slot));
          if (prof) {
            t_mma += t2 - t1;
            t_commit += clock64() - t2;
          }
          if (++slot == NSTAGE2) {
            slot = 0;
            phase ^= 1;
          }
        }
        if (g.commit_acc != 0xff) umma2_commit_both_elect(bar0 + 8 * (P_ACC + g.commit_acc));
        if (prof && lane == 0) prof[2 + 2 * s] = (unsigned long long)clock64();
      }
      if (prof && lane == 0) {
        prof[15] = (unsigned long long)stall;
        prof[30] = (unsigned long long)t_mma;
        prof[31] = (unsigned long long)t_commit;
      }
    }
  } else {
    // ---- epilogue warps 2..9: TMEM lane quadrant q = warp % 4 -> rows 32 (q & 1) + lane, column half q >> 1; part (warp - 2) / 4
    const int et = tid - 64;
    const int q = warp & 3, part = (warp - 2) >> 2;
    const int row = 32 * (q & 1) + lane;
    uint8_t *hi = smem + OFF_ACT_HI, *lo = smem + OFF_ACT_LO;
    const uint32_t act0 = map_to_cta(bar0 + 8 * P_ACT, 0);     // CTA 0's activation barriers
    {
      constexpr int kIter = TM * (KX / 2) / kEpiThreads;
      float v0[kIter], v1[kIter];
#pragma unroll
      for (int it = 0; it < kIter; ++it) {
        const int i = et + it * kEpiThreads, r = i / (KX / 2), c = 2 * (i - r * (KX / 2));
        const int gr = rowmap[r];
        const float* xr = C.x + (size_t)(gr >= 0 ? gr : 0) * C.ldx;
        v0[it] = (gr >= 0 && c < C.d_in) ? __ldg(xr + c) : 0.0f;
        v1[it] = (gr >= 0 && c + 1 < C.d_in) ? __ldg(xr + c + 1) : 0.0f;
      }
#pragma unroll
      for (int it = 0; it < kIter; ++it) {
        const int i = et + it * kEpiThreads, r = i / (KX / 2), c = 2 * (i - r * (KX / 2));
        store_pair(smem + OFF_X_HI, smem + OFF_X_LO, r, c, fminf(fmaxf(v0[it], -ACT_CLAMP), ACT_CLAMP),
                   fminf(fmaxf(v1[it], -ACT_CLAMP), ACT_CLAMP));
      }
    }
    warp_arrive_leader(act0 + 8 * 5, lane);
    {  // H = tanh(x W1 + b1)
      const float us = __ldg(C.us_w1);
      for (int h = 0; h < 2; ++h) {
        mbar_wait(bar0 + 8 * (P_ACC + h), 0);
        if (prof && et == 0) prof[16 + 2 * h] = (unsigned long long)clock64();
        fence_tc_after();
        epi2_tanh_half<false>(smem, tmem, 128 * h, 256 * h, C.b1, us, q, part, lane, 0u);
        warp_arrive_leader(act0 + 8 * h, lane);
        if (prof && et == 0) prof[17 + 2 * h] = (unsigned long long)clock64();
      }
    }
    if (C.att_n > 0) {   // single-token attention: r = full + (full Wa + ba), then L2-normalise the block
      const float us = __ldg(C.us_att);
      mbar_wait(bar0 + 8 * (P_ACC + 2), 0);
      if (prof && et == 0) prof[20] = (unsigned long long)clock64();
      fence_tc_after();
      const int n_sub = C.att_pad >> 1;                                   // this lane half's columns: D column n_sub (q >> 1) + t
      const int t_mid = (((n_sub >> 3) + 1) >> 1) << 3;                    // part 0: TMEM columns [0, t_mid), part 1: [t_mid, n_sub)
      const int t_beg = part ? t_mid : 0, t_end = part ? n_sub : t_mid;
      const uint32_t taddr = tmem + ((uint32_t)(32 * q) << 16) + 256u;
      float v[5][8];
      float ss = 0.0f;
#pragma unroll
      for (int u = 0; u < 5; ++u) {
        const int t0 = t_beg + 8 * u;
        if (t0 < t_end) {                                                  // warp-uniform
          uint32_t r[8];
          tmem_ld_32x32b_x8(taddr + t0, r);
          tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const int col = n_sub * (q >> 1) + t0 + j;
            float x = 0.0f;
            if (col < C.att_n) {
              const uint32_t o = canon(row, C.att_lo + col);
              const float res = (__half2float(*reinterpret_cast<const __half*>(hi + o)) + __half2float(*reinterpret_cast<const __half*>(lo + o))) *
                                ACT_UNSCALE;
              x = fmaf(__uint_as_float(r[j]), us, __ldg(C.batt + col)) + res;
            }
            v[u][j] = x;
            ss += x * x;
          }
        }
      }
      ssum[2 * (q >> 1) + part][row] = ss;
      asm volatile("bar.sync 1, %0;" ::"n"(kEpiThreads) : "memory");       // the eight epilogue warps
      const float inv = 1.0f / fmaxf(sqrtf(ssum[0][row] + ssum[1][row] + ssum[2][row] + ssum[3][row]), 1e-12f);   // F.normalize
#pragma unroll
      for (int u = 0; u < 5; ++u) {
        const int t0 = t_beg + 8 * u;
        if (t0 < t_end) {
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const int col = n_sub * (q >> 1) + t0 + j;
            if (col < C.att_n) {
              const float sv = v[u][j] * inv * ACT_SCALE;
              const __half h = __float2half_rn(sv);
              const uint32_t o = canon(row, C.att_lo + col);
              *reinterpret_cast<__half*>(hi + o) = h;
              *reinterpret_cast<__half*>(lo + o) = __float2half_rn(sv - __half2float(h));
            }
          }
        }
      }
      warp_arrive_leader(act0 + 8 * 2, lane);
      if (prof && et == 0) prof[21] = (unsigned long long)clock64();
    }
    {  // Z = tanh(in Ws + bs), in place: the first half is computed while the second half's MMAs run and stored once they
       // have all read the tile
      const float us = __ldg(C.us_ws);
      mbar_wait(bar0 + 8 * (P_ACC + 3), 0);
      if (prof && et == 0) prof[22] = (unsigned long long)clock64();
      fence_tc_after();
      epi2_tanh_half<true>(smem, tmem, 0, 0, C.bs, us, q, part, lane, bar0 + 8 * (P_ACC + 4));
      warp_arrive_leader(act0 + 8 * 3, lane);
      if (prof && et == 0) prof[23] = (unsigned long long)clock64();
      fence_tc_after();
      epi2_tanh_half<false>(smem, tmem, 128, 256, C.bs, us, q, part, lane, 0u);
      warp_arrive_leader(act0 + 8 * 4, lane);
      if (prof && et == 0) prof[24] = (unsigned long long)clock64();
    }
    if (part == 0) {  // head: logits or value (+ optional per-head argmax, env_base.py:373-382); N = 32: 16 columns per lane half
      const float us = __ldg(C.us_wh);
      mbar_wait(bar0 + 8 * (P_ACC + 5), 0);
      if (prof && et == 0) prof[25] = (unsigned long long)clock64();
      fence_tc_after();
      float* lg = reinterpret_cast<float*>(smem + OFF_X_HI);       // [TM][33]; the input tile is dead by now
      uint32_t r[16];
      tmem_ld_32x32b_x16(tmem + ((uint32_t)(32 * q) << 16) + 384u, r);
      tmem_ld_wait();
      const int gr = rowmap[row];
#pragma unroll
      for (int j = 0; j < 16; ++j) {
        const int col = 16 * (q >> 1) + j;
        if (col < C.n_out) {
          const float x = fmaf(__uint_as_float(r[j]), us, __ldg(C.bh + col));
          if (C.out && gr >= 0) C.out[(size_t)gr * C.ld_out + col] = x;
          lg[row * 33 + col] = x;
        }
      }
      if (C.act_out) {
        asm volatile("bar.sync 2, 128;" ::: "memory");             // the four part-0 warps hold a row's 32 columns between them
        if (q < 2 && gr >= 0) {
          const float* lr = lg + row * 33;
          int4 a = make_int4(0, 0, 0, 0);
          int o = 0;
#pragma unroll
          for (int h = 0; h < 4; ++h) {
            if (h < C.n_heads) {
              int best = 0;
              float bv = lr[o];
              for (int k = 1; k < C.head[h]; ++k)
                if (lr[o + k] > bv) { bv = lr[o + k]; best = k; }      // first maximum, like torch.argmax
              (h == 0 ? a.x : h == 1 ? a.y : h == 2 ? a.z : a.w) = best;
              o += C.head[h];
            }
          }
          reinterpret_cast<int4*>(C.act_out)[(size_t)gr * C.ld_act] = a;
        }
      }
    }
    if (prof && et == 0) prof[26] = (unsigned long long)clock64();
  }
  fence_tc_before();
  __syncthreads();
  cluster_sync();                       // neither CTA leaves (or frees tensor memory) while the pair's MMAs may still run
  if (warp == 1) {
    __syncwarp();
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(kTmemCols));
  }
}

// =====================================================================================================================
// M = 128 form (default): one CTA = 128 rows of one chain, so every weight byte and every MMA serves twice the rows of the
// 64-row form above (an M = 128 kind::f16 MMA takes the time of an M = 64 one, profiles/r2v_tcgen05_probe4.txt).  A 128-row
// tile of hi + lo activations does not fit in shared memory; the hi halves live there ([128 x 512] halves, canonical layout with
// 128 rows: LBO 2048, SBO 128) and the LO halves in TENSOR MEMORY, columns [256, 512): the A operand of the A_lo B_hi MMA is read
// from TMEM (lane = row, 32-bit column c = elements k = 2 c, 2 c + 1), written there by the epilogue with tcgen05.st.  That leaves
// ONE accumulator region, columns [0, 256): a 500-wide layer is two N = 256 halves through the same region, the epilogue
// DRAINS it into registers (64 values per thread: thread = row, 16 epilogue warps = 4 lane quadrants x 4 column parts) and frees
// it at once, so the next half's MMAs run under the tanh / split / store work; the shared layer's first half is held in
// registers until the second half's MMAs have read the tile it overwrites.  The input tile's lo halves sit in TMEM columns
// [256, 296) (the head of the lo region, rewritten only by layer 1's LAST epilogue: layer 1 computes its upper half first).
// Schedule (segment table in hh_pf_tc_launch): L1 upper half, L1 lower half, attention (MMAs under the lower half's epilogue),
// shared layer half 0 below the attention block (under the attention epilogue), half 0 from the block on, half 1, head on the
// first / second 256 columns.  Weight images are those of the 64-row form (the attention image starts on a K = 16 step).
constexpr int TM2 = 128;
constexpr uint32_t ACT2_BYTES = TM2 * KA * 2;              // 131 072: hi halves only
constexpr uint32_t X2_BYTES = TM2 * KX * 2;                // 20 480
constexpr uint32_t OFF2_ACT = 0, OFF2_X = ACT2_BYTES, OFF2_RING = ACT2_BYTES + X2_BYTES;
constexpr uint32_t SMEM2_BYTES = OFF2_RING + NSTAGE * STAGE_BYTES;   // 217 088
constexpr uint32_t LBO_A2 = TM2 * 16;
constexpr int kEpiWarps2 = 16, kThreads2 = 64 + 32 * kEpiWarps2;     // 576
// x_lo aliases the lo halves of activation columns 0 .. 79: layer 1 writes its upper half (TMEM columns 384 ..) first, while the
// MMAs of its lower half still read x_lo; the lower half's epilogue runs after all of layer 1's MMAs
constexpr uint32_t kAccCol = 0, kLoCol = 256, kXLoCol = 256;
constexpr int Q_FULL = 0, Q_EMPTY = NSTAGE, Q_ACC = 2 * NSTAGE, Q_ACT = Q_ACC + 6, Q_FREE = Q_ACT + 6, Q_COUNT = Q_FREE + 6;

__device__ __forceinline__ void tmem_st_32x32b_x16(uint32_t taddr, const uint32_t (&r)[16]) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};" ::"r"(taddr),
               "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]), "r"(r[10]),
               "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
               : "memory");
}
__device__ __forceinline__ void tmem_st_32x32b_x2(uint32_t taddr, uint32_t r0, uint32_t r1) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x2.b32 [%0], {%1, %2};" ::"r"(taddr), "r"(r0), "r"(r1) : "memory");
}
__device__ __forceinline__ void tmem_ld_32x32b_x2(uint32_t taddr, uint32_t& r0, uint32_t& r1) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x2.b32 {%0, %1}, [%2];" : "=r"(r0), "=r"(r1) : "r"(taddr));
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
// one K = 16 step: A_lo (tensor memory) B_hi with the accumulate flag, then A_hi (shared memory) B_lo, A_hi B_hi
__device__ __forceinline__ void umma3_m128(uint32_t tmem_d, uint32_t ta_lo, uint32_t a_hi, uint32_t b_hi, uint32_t b_lo, uint32_t idesc,
                                           uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p, t, q;\n\t.reg .b32 hi;\n\t.reg .b64 dah, dbh, dbl;\n\t"
      "mov.u32 hi, 0x4008;\n\t"
      "mov.b64 dah, {%2, hi};\n\tmov.b64 dbh, {%3, hi};\n\tmov.b64 dbl, {%4, hi};\n\t"
      "elect.sync _|q, 0xffffffff;\n\tsetp.ne.b32 p, %6, 0;\n\tsetp.eq.u32 t, 0, 0;\n\t"
      "@q tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], dbh, %5, p;\n\t"
      "@q tcgen05.mma.cta_group::1.kind::f16 [%0], dah, dbl, %5, t;\n\t"
      "@q tcgen05.mma.cta_group::1.kind::f16 [%0], dah, dbh, %5, t;\n\t}\n" ::"r"(tmem_d),
      "r"(ta_lo), "r"(a_hi), "r"(b_hi), "r"(b_lo), "r"(idesc), "r"(accumulate)
      : "memory");
}
// a warp's part of a phase is done (its shared-memory and tensor-memory writes are visible to the tensor cores, its TMEM reads
// complete): one lane arrives
__device__ __forceinline__ void warp_arrive_local(uint32_t bar, int lane) {
  fence_async_smem();
  tmem_st_wait();
  fence_tc_before();
  __syncwarp();
  if (lane == 0) mbar_arrive(bar);
}

// epilogue of one N = 256 half: thread = row 32 q + lane, columns 64 part + [0, 64) of the half.  Drains the accumulator, frees
// it (free_bar), computes tanh(acc * us + bias) and its hi / lo split, and -- once hold_bar (if any) has completed -- stores hi
// into the activation tile and lo into tensor memory at activation columns k0 + 64 part + [0, 64).
template <bool HOLD>
__device__ __forceinline__ void epi128_tanh_half(uint8_t* smem, uint32_t tmem, int k0, const float* __restrict__ bias, float us, int q,
                                                 int part, int lane, uint32_t free_bar, uint32_t hold_bar, uint32_t zero_rt,
                                                 unsigned long long* stamp = nullptr) {
  const int row = 32 * q + lane, kb = k0 + 64 * part;
  const uint32_t lane_base = tmem + ((uint32_t)(32 * q) << 16);
  float us2 = us * kTwoLog2e;
  uint32_t raw[64];
#pragma unroll
  for (int c = 0; c < 4; ++c) tmem_ld_32x32b_x16(lane_base + kAccCol + 64 * part + 16 * c, *reinterpret_cast<uint32_t(*)[16]>(&raw[16 * c]));
  tmem_ld_wait();
  fence_tc_before();
  __syncwarp();
  if (lane == 0) mbar_arrive(free_bar);         // the accumulator region can take the next MMAs
  {  // Without a dependence ptxas hoists the tanh math between the four loads above (each has its own scoreboard), and the arrive
     // -- the next segment's MMAs -- comes ~3 000 cycles late.  The scale picks up (clock & 0) read after the arrive.
    uint32_t fence_clk;
    asm volatile("mov.u32 %0, %%clock;" : "=r"(fence_clk)::"memory");
    us2 = __uint_as_float(__float_as_uint(us2) | (fence_clk & zero_rt));
  }
  if (stamp) stamp[0] = (unsigned long long)clock64();
  uint4 hh[8];
  uint32_t ll[32];
#pragma unroll
  for (int g = 0; g < 8; ++g) {
    const int k = kb + 8 * g;
    const float4 b0 = __ldg(reinterpret_cast<const float4*>(bias + k)), b1 = __ldg(reinterpret_cast<const float4*>(bias + k + 4));
    const float bb[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
    __half2 h[4], l[4];
    float sv[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) sv[j] = fmaf(__uint_as_float(raw[8 * g + j]), us2, bb[j] * kTwoLog2e);
    tanh_scaled_4096_x4(sv[0], sv[1], sv[2], sv[3]);
    tanh_scaled_4096_x4(sv[4], sv[5], sv[6], sv[7]);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float s0 = sv[2 * j], s1 = sv[2 * j + 1];
      h[j] = __floats2half2_rn(s0, s1);
      const float2 hf = __half22float2(h[j]);
      l[j] = __floats2half2_rn(s0 - hf.x, s1 - hf.y);
      ll[4 * g + j] = *reinterpret_cast<uint32_t*>(&l[j]);
    }
    hh[g] = make_uint4(*reinterpret_cast<uint32_t*>(&h[0]), *reinterpret_cast<uint32_t*>(&h[1]), *reinterpret_cast<uint32_t*>(&h[2]),
                       *reinterpret_cast<uint32_t*>(&h[3]));
  }
  if (stamp) stamp[1] = (unsigned long long)clock64();
  if (HOLD) mbar_wait(hold_bar, 0);
  uint8_t* hi = smem + OFF2_ACT;
#pragma unroll
  for (int g = 0; g < 8; ++g) *reinterpret_cast<uint4*>(hi + (uint32_t)(((kb >> 3) + g) * (TM2 * 16) + row * 16)) = hh[g];
  tmem_st_32x32b_x16(lane_base + kLoCol + (kb >> 1), *reinterpret_cast<uint32_t(*)[16]>(&ll[0]));
  tmem_st_32x32b_x16(lane_base + kLoCol + (kb >> 1) + 16, *reinterpret_cast<uint32_t(*)[16]>(&ll[16]));
}

__global__ void __launch_bounds__(kThreads2, 1) policy_forward_m128_kernel(const __grid_constant__ Args args) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ __align__(8) uint64_t bars[Q_COUNT];
  __shared__ uint32_t tmem_base_s;
  __shared__ int rowmap[TM2];
  __shared__ float ssum[4][TM2];
  const Chain& C = args.c[blockIdx.y];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  int beg = 0, cnt = C.n_rows;
  if (C.range_dev) {
    beg = C.range_dev[0];
    cnt = C.range_dev[1];
  }
  const int row0 = blockIdx.x * TM2;
  if (row0 >= cnt) return;              // CTA-uniform
  unsigned long long* prof = args.prof ? args.prof + 32 * (size_t)(blockIdx.y * gridDim.x + blockIdx.x) : nullptr;
  if (prof && tid == 0) {   // CTA entry: SM clock, SM id, wall clock (how the two rounds of tiles follow each other on an SM)
    uint32_t smid;
    unsigned long long gt;
    asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(gt));
    prof[27] = (unsigned long long)clock64();
    prof[30] = smid;
    prof[31] = gt;
  }
  if (tid < TM2) {
    const int lr = row0 + tid;
    rowmap[tid] = lr < cnt ? (C.rows ? C.rows[beg + lr] : beg + lr) : -1;
  }
  const uint32_t bar0 = smem_u32(bars);
  if (tid == 0) {
    for (int i = 0; i < 2 * NSTAGE + 6; ++i) mbar_init(bar0 + 8 * i, 1);
    for (int i = 0; i < 12; ++i) mbar_init(bar0 + 8 * (Q_ACT + i), kEpiWarps2);   // activation + accumulator-free: one lane per warp
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)), "r"(kTmemCols));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  fence_tc_before();
  __syncthreads();
  fence_tc_after();
  const uint32_t tmem = tmem_base_s;

  if (warp == 0) {
    if (lane == 0) {   // ---- weight stream
      uint32_t slot = 0, phase = 0;
      const uint32_t ring = smem_u32(smem + OFF2_RING);
      long long stall = 0;
      for (int s = 0; s < C.n_seg; ++s) {
        const Seg& g = C.seg[s];
        const uint32_t bytes = (uint32_t)g.kps * g.n * 64u;
        const uint8_t* src = g.w;
        const int n_stage = g.ksteps / g.kps;
        for (int k = 0; k < n_stage; ++k) {
          const long long t0 = prof ? clock64() : 0;
          mbar_wait(bar0 + 8 * (Q_EMPTY + slot), phase ^ 1);
          if (prof) stall += clock64() - t0;
          mbar_expect_tx(bar0 + 8 * (Q_FULL + slot), bytes);
          bulk_g2s(ring + slot * STAGE_BYTES, src, bytes, bar0 + 8 * (Q_FULL + slot));
          src += bytes;
          if (++slot == NSTAGE) {
            slot = 0;
            phase ^= 1;
          }
        }
      }
      if (prof) {
        prof[28] = (unsigned long long)stall;
        prof[29] = (unsigned long long)clock64();
      }
    }
  } else if (warp == 1) {
    {   // ---- MMA issue by the converged warp (one elected lane inside the asm)
      uint32_t slot = 0, phase = 0;
      long long stall = 0;
      if (prof && lane == 0) prof[0] = (unsigned long long)clock64();
      mbar_wait_uniform(bar0 + 8 * (Q_ACT + 5), 0);   // the input tile is in shared / tensor memory
      for (int s = 0; s < C.n_seg; ++s) {
        const Seg& g = C.seg[s];
        const uint32_t n = g.n, kps = g.kps, n_stage = g.ksteps / kps;
        if (g.wait_act != 0xff) mbar_wait_uniform(bar0 + 8 * (Q_ACT + g.wait_act), 0);
        if (g.wait_free != 0xff) mbar_wait_uniform(bar0 + 8 * (Q_FREE + g.wait_free), 0);
        if (prof && lane == 0 && s < 7) prof[1 + 2 * s] = (unsigned long long)clock64();
        fence_tc_after();
        const uint32_t idesc = instr_desc_f16(TM2, n);
        const uint32_t a_addr = smem_u32(smem + (g.a_src ? OFF2_ACT : OFF2_X)) + (uint32_t)(g.a_k0 >> 3) * (TM2 * 16);
        uint32_t a_hi = desc_lo(a_addr, LBO_A2);
        uint32_t ta_lo = tmem + (g.a_src ? kLoCol : kXLoCol) + (uint32_t)(g.a_k0 >> 1);
        const uint32_t b_seg = desc_lo(smem_u32(smem + OFF2_RING), 16u * n);
        const uint32_t lo_off = 2u * n, step_off = 4u * n;
        const uint32_t d = tmem + kAccCol;
        uint32_t acc = g.first ? 0u : 1u;
        for (uint32_t k = 0; k < n_stage; ++k) {
          const long long t0 = prof ? clock64() : 0;
          mbar_wait_uniform(bar0 + 8 * (Q_FULL + slot), phase);
          if (prof) stall += clock64() - t0;
          fence_tc_after();
          uint32_t b = b_seg + slot * (STAGE_BYTES >> 4);
          for (uint32_t j = 0; j < kps; ++j) {
            umma3_m128(d, ta_lo, a_hi, b, b + lo_off, idesc, acc);
            acc = 1u;
            a_hi += (2 * LBO_A2) >> 4;
            ta_lo += 8;
            b += step_off;
          }
          umma_commit_elect(bar0 + 8 * (Q_EMPTY + slot));
          if (++slot == NSTAGE) {
            slot = 0;
            phase ^= 1;
          }
        }
        if (g.commit_acc != 0xff) umma_commit_elect(bar0 + 8 * (Q_ACC + g.commit_acc));
        if (prof && lane == 0 && s < 7) prof[2 + 2 * s] = (unsigned long long)clock64();
      }
      if (prof && lane == 0) prof[15] = (unsigned long long)stall;
    }
  } else {
    // ---- epilogue warps 2..17: TMEM lane quadrant q = warp % 4 (rows 32 q + lane), column part (warp - 2) / 4
    const int et = tid - 64;
    const int q = warp & 3, part = (warp - 2) >> 2;
    const int row = 32 * q + lane;
    const uint32_t lane_base = tmem + ((uint32_t)(32 * q) << 16);
    uint8_t* hi = smem + OFF2_ACT;
    {  // input tile -> hi halves (shared memory) and lo halves (tensor memory) of 2^12 x.  Coalesced: consecutive threads read
       // consecutive floats of a row; the lo halves pass through the (still unused) activation tile as [128][80] halves, because
       // only the thread that owns a TMEM lane (= row) can write it
      __half* lo_stage = reinterpret_cast<__half*>(smem + OFF2_ACT);
      const int d_in = C.d_in;
      // thread = one column c of the tile and every sixth row (480 of the 512 threads): the column's predicate, global offset and
      // shared-memory offset are loop constants, all 22 loads are in flight before the first use
      constexpr int kRowsPerPass = 32 * kEpiWarps2 / KX;            // 6
      constexpr int kIter = (TM2 + kRowsPerPass - 1) / kRowsPerPass; // 22
      if (et < kRowsPerPass * KX) {
        const int r0 = et / KX, c = et - r0 * KX;
        const bool col_ok = c < d_in;
        const float* xc = C.x + c;
        float xv[kIter];
#pragma unroll
        for (int i = 0; i < kIter; ++i) {
          const int r = r0 + kRowsPerPass * i;
          const int gr = r < TM2 ? rowmap[r] : -1;
          xv[i] = (col_ok && gr >= 0) ? __ldg(xc + (size_t)gr * C.ldx) : 0.0f;
        }
        uint8_t* hi_c = smem + OFF2_X + (uint32_t)((c >> 3) * (TM2 * 16) + (c & 7) * 2);
#pragma unroll
        for (int i = 0; i < kIter; ++i) {
          const int r = r0 + kRowsPerPass * i;
          if (r < TM2) {
            const float sv = fminf(fmaxf(xv[i], -ACT_CLAMP), ACT_CLAMP) * ACT_SCALE;
            const __half h = __float2half_rn(sv);
            *reinterpret_cast<__half*>(hi_c + r * 16) = h;
            lo_stage[r * KX + c] = __float2half_rn(sv - __half2float(h));
          }
        }
      }
      asm volatile("bar.sync 1, %0;" ::"n"(32 * kEpiWarps2) : "memory");
      uint32_t lo_w[10];
#pragma unroll
      for (int j = 0; j < 10; ++j) lo_w[j] = *reinterpret_cast<const uint32_t*>(lo_stage + row * KX + 20 * part + 2 * j);
#pragma unroll
      for (int j = 0; j < 5; ++j) tmem_st_32x32b_x2(lane_base + kXLoCol + 10 * part + 2 * j, lo_w[2 * j], lo_w[2 * j + 1]);
      warp_arrive_local(bar0 + 8 * (Q_ACT + 5), lane);
    }
    {  // H = tanh(x W1 + b1)
      const float us = __ldg(C.us_w1);
      for (int h = 0; h < 2; ++h) {                    // processing order: columns 256 .. 511, then 0 .. 255 (see the segment table)
        mbar_wait(bar0 + 8 * (Q_ACC + h), 0);
        if (prof && et == 0) prof[16 + 2 * h] = (unsigned long long)clock64();
        fence_tc_after();
        epi128_tanh_half<false>(smem, tmem, 256 * (1 - h), C.b1, us, q, part, lane, bar0 + 8 * (Q_FREE + h), 0u, args.zero);
        warp_arrive_local(bar0 + 8 * (Q_ACT + h), lane);
        if (prof && et == 0) prof[17 + 2 * h] = (unsigned long long)clock64();
      }
    }
    if (C.att_n > 0) {   // single-token attention: r = full + (full Wa + ba), then L2-normalise the block
      const float us = __ldg(C.us_att);
      mbar_wait(bar0 + 8 * (Q_ACC + 2), 0);
      if (prof && et == 0) prof[20] = (unsigned long long)clock64();
      fence_tc_after();
      // this warp's columns of the block: [c_beg, c_end), a multiple of 8 per part (att_pad / 4 rounded up to 8: 32 or 40)
      const int per = (((C.att_pad + 3) >> 2) + 7) & ~7;
      const int c_beg = part * per, c_end = min((int)C.att_pad, c_beg + per);
      const int klo = (C.att_lo + c_beg) >> 1;            // first column of the lo operand that belongs to this part
      uint32_t racc[40], rlo[20];
#pragma unroll
      for (int u = 0; u < 5; ++u) {
        if (c_beg + 8 * u < c_end) {                       // warp-uniform
          tmem_ld_32x32b_x8(lane_base + kAccCol + c_beg + 8 * u, *reinterpret_cast<uint32_t(*)[8]>(&racc[8 * u]));
          // lo pairs of the same 8 columns = 4 TMEM columns (att_lo + att_pad <= 512, checked on the host: inside the lo region)
          asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0, %1, %2, %3}, [%4];"
                       : "=r"(rlo[4 * u]), "=r"(rlo[4 * u + 1]), "=r"(rlo[4 * u + 2]), "=r"(rlo[4 * u + 3])
                       : "r"(lane_base + kLoCol + (uint32_t)(klo + 4 * u)));
        }
      }
      tmem_ld_wait();
      fence_tc_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(bar0 + 8 * (Q_FREE + 2));
      float v[40];
      float ss = 0.0f;
#pragma unroll
      for (int u = 0; u < 5; ++u) {
        const bool whole = c_beg + 8 * u + 8 <= c_end && c_beg + 8 * u + 8 <= C.att_n;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const int col = c_beg + 8 * u + 2 * j, k = C.att_lo + col;
          float v0 = 0.0f, v1 = 0.0f;
          if (whole || (col < c_end && col < C.att_n)) {
            const float2 b = __ldg(reinterpret_cast<const float2*>(C.batt + col));
            const float2 hf = __half22float2(*reinterpret_cast<const __half2*>(hi + (uint32_t)((k >> 3) * (TM2 * 16) + row * 16 + (k & 7) * 2)));
            const uint32_t lw = rlo[4 * u + j];
            const float2 lf = __half22float2(*reinterpret_cast<const __half2*>(&lw));
            v0 = fmaf(__uint_as_float(racc[8 * u + 2 * j]), us, b.x) + (hf.x + lf.x) * ACT_UNSCALE;
            v1 = fmaf(__uint_as_float(racc[8 * u + 2 * j + 1]), us, b.y) + (hf.y + lf.y) * ACT_UNSCALE;
          }
          v[8 * u + 2 * j] = v0;
          v[8 * u + 2 * j + 1] = v1;
          ss += v0 * v0 + v1 * v1;
        }
      }
      ssum[part][row] = ss;
      asm volatile("bar.sync 1, %0;" ::"n"(32 * kEpiWarps2) : "memory");       // the sixteen epilogue warps
      const float inv = 1.0f / fmaxf(sqrtf(ssum[0][row] + ssum[1][row] + ssum[2][row] + ssum[3][row]), 1e-12f);   // F.normalize
#pragma unroll
      for (int u = 0; u < 5; ++u) {
        const int col8 = c_beg + 8 * u;
        if (col8 + 8 <= c_end && col8 + 8 <= C.att_n) {      // a whole group of eight columns (warp-uniform): one x4 store of lo
          uint32_t lw[4];
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const int k = C.att_lo + col8 + 2 * j;
            __half2 h, l;
            split_pair(v[8 * u + 2 * j] * inv, v[8 * u + 2 * j + 1] * inv, h, l);
            *reinterpret_cast<__half2*>(hi + (uint32_t)((k >> 3) * (TM2 * 16) + row * 16 + (k & 7) * 2)) = h;
            lw[j] = *reinterpret_cast<uint32_t*>(&l);
          }
          asm volatile("tcgen05.st.sync.aligned.32x32b.x4.b32 [%0], {%1, %2, %3, %4};" ::"r"(lane_base + kLoCol + (uint32_t)((C.att_lo + col8) >> 1)),
                       "r"(lw[0]), "r"(lw[1]), "r"(lw[2]), "r"(lw[3])
                       : "memory");
        } else {
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const int col = col8 + 2 * j, k = C.att_lo + col;
            if (col < c_end && col < C.att_n) {
              __half2 h, l;
              split_pair(v[8 * u + 2 * j] * inv, v[8 * u + 2 * j + 1] * inv, h, l);
              *reinterpret_cast<__half2*>(hi + (uint32_t)((k >> 3) * (TM2 * 16) + row * 16 + (k & 7) * 2)) = h;
              // one 32-bit column of the lo operand = this pair (att_lo is even)
              asm volatile("tcgen05.st.sync.aligned.32x32b.x1.b32 [%0], {%1};" ::"r"(lane_base + kLoCol + (k >> 1)),
                           "r"(*reinterpret_cast<uint32_t*>(&l))
                           : "memory");
            }
          }
        }
      }
      warp_arrive_local(bar0 + 8 * (Q_ACT + 2), lane);
      if (prof && et == 0) prof[21] = (unsigned long long)clock64();
    }
    {  // Z = tanh(in Ws + bs), in place: the first half is drained and computed while the second half's MMAs run and stored
       // once they have all read the tile
      const float us = __ldg(C.us_ws);
      mbar_wait(bar0 + 8 * (Q_ACC + 3), 0);
      if (prof && et == 0) prof[22] = (unsigned long long)clock64();
      fence_tc_after();
      epi128_tanh_half<true>(smem, tmem, 0, C.bs, us, q, part, lane, bar0 + 8 * (Q_FREE + 3), bar0 + 8 * (Q_ACC + 4), args.zero,
                             nullptr);
      warp_arrive_local(bar0 + 8 * (Q_ACT + 3), lane);
      if (prof && et == 0) prof[23] = (unsigned long long)clock64();
      fence_tc_after();
      epi128_tanh_half<false>(smem, tmem, 256, C.bs, us, q, part, lane, bar0 + 8 * (Q_FREE + 4), 0u, args.zero);
      warp_arrive_local(bar0 + 8 * (Q_ACT + 4), lane);
      if (prof && et == 0) prof[24] = (unsigned long long)clock64();
    }
    {  // head: logits or value (+ optional per-head argmax, env_base.py:373-382).  The four part-0 warps read the accumulator
       // (thread = row, 32 columns) and stage the rows; all sixteen warps write them out, a row's columns on consecutive lanes
      float* lg = reinterpret_cast<float*>(smem + OFF2_X);       // [TM2][33]; the input tile is dead by now
      if (part == 0) {
        const float us = __ldg(C.us_wh);
        mbar_wait(bar0 + 8 * (Q_ACC + 5), 0);
        if (prof && et == 0) prof[25] = (unsigned long long)clock64();
        fence_tc_after();
        uint32_t r[32];
        tmem_ld_32x32b_x16(lane_base + kAccCol, *reinterpret_cast<uint32_t(*)[16]>(&r[0]));
        tmem_ld_32x32b_x16(lane_base + kAccCol + 16, *reinterpret_cast<uint32_t(*)[16]>(&r[16]));
        tmem_ld_wait();
#pragma unroll
        for (int col = 0; col < 32; ++col) lg[row * 33 + col] = col < C.n_out ? fmaf(__uint_as_float(r[col]), us, __ldg(C.bh + col)) : 0.0f;
      }
      asm volatile("bar.sync 1, %0;" ::"n"(32 * kEpiWarps2) : "memory");
      const int ew = warp - 2;                                    // 0 .. 15: rows 8 ew .. 8 ew + 7
      if (C.out && lane < C.n_out) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const int r = 8 * ew + i, gr = rowmap[r];
          if (gr >= 0) C.out[(size_t)gr * C.ld_out + lane] = lg[r * 33 + lane];
        }
      }
      if (C.act_out && lane < 8) {
        const int r = 8 * ew + lane, gr = rowmap[r];
        if (gr >= 0) {
          const float* lr = lg + r * 33;
          int4 a = make_int4(0, 0, 0, 0);
          int o = 0;
#pragma unroll
          for (int h = 0; h < 4; ++h) {
            if (h < C.n_heads) {
              int best = 0;
              float bv = lr[o];
              for (int k = 1; k < C.head[h]; ++k)
                if (lr[o + k] > bv) { bv = lr[o + k]; best = k; }      // first maximum, like torch.argmax
              (h == 0 ? a.x : h == 1 ? a.y : h == 2 ? a.z : a.w) = best;
              o += C.head[h];
            }
          }
          reinterpret_cast<int4*>(C.act_out)[(size_t)gr * C.ld_act] = a;
        }
      }
    }
    if (prof && et == 0) prof[26] = (unsigned long long)clock64();
  }
  fence_tc_before();
  __syncthreads();
  if (warp == 1) {
    __syncwarp();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(kTmemCols));
    if (prof && lane == 0) prof[14] = (unsigned long long)clock64();      // CTA exit
  }
}

// ---- operand images -----------------------------------------------------------------------------------------------
// us[0] = 2^-(12 + s), us[1] = 2^s with s such that max |2^s w| lies in [2^13, 2^14)
__global__ void pack_scale_kernel(const float* __restrict__ w, int k_rows, int ldw, int n_cols, float* __restrict__ us) {
  __shared__ float red[32];
  float m = 0.0f;
  for (int i = threadIdx.x; i < k_rows * n_cols; i += blockDim.x) m = fmaxf(m, fabsf(w[(size_t)(i / n_cols) * ldw + i % n_cols]));
  for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = m;
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int i = 1; i < (int)(blockDim.x >> 5); ++i) m = fmaxf(m, red[i]);
    int s = 0;
    if (m > 0.0f && isfinite(m)) s = 13 - ilogbf(m);
    s = max(-60, min(60, s));
    us[0] = ldexpf(1.0f, -(12 + s));
    us[1] = ldexpf(1.0f, s);
  }
}
// image = [chunk][stage][cta r < split][kstep j < kps]{ hi [2 k-halves][n_sub][8], lo [2][n_sub][8] } halves, n_sub = n_chunk / split:
// every (stage, CTA) part is contiguous = one bulk copy.  Image row k' = kstep * 16 + half * 8 + kk holds w row k' - row_shift
// (zero outside [0, k_rows)); column c of chunk ch holds w column ch * n_chunk + c (zero from n_cols on); CTA r of a pair
// owns the columns [r n_sub, (r + 1) n_sub) of the chunk.
__global__ void pack_image_kernel(const float* __restrict__ w, int k_rows, int n_cols, int ldw, int n_total, int n_chunk, int row_shift,
                                  int ksteps, int kps, int split, const float* __restrict__ us, __half* __restrict__ img) {
  const int total = ksteps * 16 * n_total;
  const float scale = us[1];
  const int n_sub = n_chunk / split, n_stage = ksteps / kps;
  for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < total; e += gridDim.x * blockDim.x) {
    int t = e;
    const int kk = t & 7; t >>= 3;
    const int c = t % n_chunk; t /= n_chunk;
    const int h = t & 1; t >>= 1;
    const int ks = t % ksteps;
    const int chunk = t / ksteps;
    const int row = ks * 16 + h * 8 + kk - row_shift, col = chunk * n_chunk + c;
    const float v = (row >= 0 && row < k_rows && col < n_cols) ? w[(size_t)row * ldw + col] * scale : 0.0f;
    const __half hi = __float2half_rn(v);
    const __half lo = __float2half_rn(v - __half2float(hi));
    const int r = c / n_sub, cs = c - r * n_sub, stage = ks / kps, j = ks - stage * kps;
    const size_t blk = (((size_t)(chunk * n_stage + stage) * split + r) * kps + j) * ((size_t)n_sub * 32);
    const size_t base = blk + (size_t)(h * n_sub + cs) * 8 + kk;
    img[base] = hi;
    img[base + (size_t)n_sub * 16] = lo;
  }
}

}  // namespace tc
}  // namespace hh

// tuning / profiling knobs of the tcgen05 path (tests and profiles/ only)
static unsigned long long* g_prof = nullptr;
// 0 (default): one CTA per 64-row tile; 1 (HH_TC_PAIR=1): the CTA-pair kernel (cta_group::2).  Measured at 8 192 rows x 4 chains
// (profiles/README.md, round 2): 143 us against 185 us -- the pair form halves the weight bytes per SM and its MMAs run at
// twice the rate, but every MMA costs the issuing warp ~100 cycles (seven vector -> uniform register moves per UTCHMMA), more
// than a cta_group::2 MMA takes to execute, and the pair's extra signalling (relay, multicast commits) is on the critical
// path; it stays in the library as a tested variant.
static int g_mode = [] {            // HH_TC_MODE = m128 (default) | m64 | pair   (HH_TC_PAIR=1 is the older spelling of pair)
  const char* e = getenv("HH_TC_MODE");
  const char* p = getenv("HH_TC_PAIR");
  if (e && std::string(e) == "m64") return 0;
  if ((e && std::string(e) == "pair") || (p && atoi(p) == 1)) return 1;
  return 2;
}();
extern "C" int hh_policy_tc_profile(unsigned long long* stamps_dev) {   // 32 clock64() stamps per CTA, null = off
  g_prof = stamps_dev;
  return 0;
}
static int g_debug = 0;
extern "C" int hh_policy_tc_debug(int32_t flags) {   // timing experiments (profiles/tc_profile.py): results are garbage
  g_debug = flags;
  return 0;
}
// 0: 64-row tiles, one CTA each; 1: the CTA-pair form; 2 (default): 128-row tiles with the lo halves in tensor memory.  Decides the
// layout hh_policy_pack produces (pair: every stage split into two column halves) and the attention block's geometry
extern "C" int32_t hh_policy_tc_mode(void) { return g_mode; }
extern "C" int32_t hh_policy_tc_pair(void) { return g_mode == 1; }

extern "C" int64_t hh_policy_image_bytes(int32_t ksteps, int32_t n_total) { return (int64_t)ksteps * n_total * 64; }

int hh_pf_tc_pack(const float* w_dev, int k_rows, int n_cols, int ldw, int n_total, int n_chunk, int row_shift, int ksteps, int kps,
                  void* image_dev, float* unscale_dev, void* stream, std::string& err) {
  using namespace hh::tc;
  const int g_pair = g_mode == 1;
  const int split = 1 + g_pair;
  if (!w_dev || !image_dev || !unscale_dev || k_rows <= 0 || n_cols <= 0 || ldw < n_cols || n_total < n_cols || n_chunk <= 0 ||
      n_total % n_chunk || n_chunk % (8 * split) || (g_pair && n_chunk % 16) || n_chunk > 256 || ksteps <= 0 || kps <= 0 || ksteps % kps ||
      row_shift < 0 || (uint32_t)kps * n_chunk * 64u > STAGE_BYTES) {
    err = "hh_policy_pack: bad argument";
    return -1;
  }
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  pack_scale_kernel<<<1, 1024, 0, st>>>(w_dev, k_rows, ldw, n_cols, unscale_dev);
  const int total = ksteps * 16 * n_total;
  pack_image_kernel<<<(total + 255) / 256, 256, 0, st>>>(w_dev, k_rows, n_cols, ldw, n_total, n_chunk, row_shift, ksteps, kps, split,
                                                         unscale_dev, static_cast<__half*>(image_dev));
  const cudaError_t ce = cudaGetLastError();
  if (ce != cudaSuccess) {
    err = std::string("hh_policy_pack launch: ") + cudaGetErrorString(ce);
    return -2;
  }
  return 0;
}

int hh_pf_tc_launch(const hh_policy_chain_ex* chains, int n_chains, int max_rows, void* stream, std::string& err) {
  using namespace hh::tc;
  Args a;
  const int pair = g_mode == 1, m128 = g_mode == 2;
  for (int i = 0; i < n_chains; ++i) {
    const hh_policy_chain_ex& s = chains[i];
    if (!s.img_w1 || !s.img_ws || !s.img_wh || !s.us_w1 || !s.us_ws || !s.us_wh || (s.att_n > 0 && (!s.img_att || !s.us_att))) {
      err = "hh_policy_forward_ex(precision = 2): the chain carries no packed operand images (hh_policy_pack)";
      return -1;
    }
    if (s.d_in > KX || (s.att_n & 1) || (s.att_lo & 1) || s.att_pad > 256 || (s.att_n > 0 && s.att_lo + s.att_n > 504)) {
      err = "hh_policy_forward_ex(precision = 2): unsupported chain shape";
      return -1;
    }
    Chain& c = a.c[i];
    c.x = s.x; c.b1 = s.b1; c.batt = s.batt; c.bs = s.bs; c.bh = s.bh;
    c.us_w1 = s.us_w1; c.us_att = s.us_att; c.us_ws = s.us_ws; c.us_wh = s.us_wh;
    c.out = s.out; c.rows = s.rows; c.range_dev = s.range_dev; c.act_out = s.act_out;
    c.n_rows = s.n_rows; c.ldx = s.ldx; c.d_in = s.d_in; c.att_lo = s.att_lo; c.att_n = s.att_n;
    // the attention step's N: the block padded to the MMA's granularity (8 columns; 16 for the pair form)
    const int att_nn = s.att_n > 0 ? ((pair || m128) ? (s.att_n + 15) / 16 * 16 : (s.att_n + 7) / 8 * 8) : 0;
    c.att_pad = att_nn;
    c.n_out = s.n_out; c.ld_out = s.ld_out; c.n_heads = s.n_heads;
    for (int h = 0; h < 4; ++h) c.head[h] = s.head[h];
    c.ld_act = s.ld_act > 0 ? s.ld_act : 1;
    const int k1s = (s.d_in + 15) / 16;
    int n = 0;
    auto seg = [&](const void* w, size_t off, int nn, int ksteps, int kps, int col, int a_src, int a_k0, int first, int wait, int commit,
                   int wait_free = 0xff) {
      Seg& g = c.seg[n++];
      g.w = static_cast<const uint8_t*>(w) + off;
      g.n = (uint16_t)nn; g.ksteps = (uint16_t)ksteps; g.kps = (uint16_t)kps; g.tmem_col = (uint16_t)col;
      g.a_k0 = (uint16_t)a_k0;
      g.a_src = (uint8_t)a_src; g.first = (uint8_t)first; g.wait_act = (uint8_t)wait; g.commit_acc = (uint8_t)commit;
      g.wait_free = (uint8_t)wait_free;
    };
    if (m128) {
      if (s.att_n > 0 && s.att_lo + att_nn > KA) {
        err = "hh_policy_forward_ex(precision = 2): attention block beyond the activation tile";
        return -1;
      }
      // one accumulator region: every segment waits for the drain of the previous one (accumulator-free barriers 0 .. 4:
      // after layer 1 half 0 / half 1, the attention block, the shared layer's half 0 / half 1)
      // Layer 1 runs its upper half (columns 256 .. 511) first: the attention block lies there, so its MMAs start as soon as
      // the lower half has left the accumulator and run under that half's tanh epilogue.
      seg(s.img_w1, (size_t)k1s * 256 * 64, 256, k1s, 1, 0, 0, 0, 1, 0xff, 0);
      seg(s.img_w1, 0, 256, k1s, 1, 0, 0, 0, 1, 0xff, 1, 0);
      int act_ready = 1, free_ready = 1;
      if (s.att_n > 0) {
        const int k0 = s.att_lo & ~15, ks = (s.att_lo + s.att_n - k0 + 15) / 16;
        // reads the first-processed half's activations (barrier 0) unless the block reaches below column 256
        seg(s.img_att, 0, att_nn, ks, 1, 0, 1, k0, 1, k0 >= 256 ? 0 : 1, 2, 1);
        act_ready = 2; free_ready = 2;
      }
      if (s.att_n > 0 && s.att_lo >= 16) {
        // The shared layer's K steps below the attention block do not read its output: they start as soon as the block has
        // left the accumulator and run under the attention epilogue (residual, L2 normalisation, operand rewrite).
        const int ks0 = s.att_lo / 16;
        seg(s.img_ws, 0, 256, ks0, 1, 0, 1, 0, 1, 1, 0xff, 2);
        seg(s.img_ws, (size_t)ks0 * 256 * 64, 256, KA / 16 - ks0, 1, 0, 1, 16 * ks0, 0, 2, 3);
      } else {
        seg(s.img_ws, 0, 256, KA / 16, 1, 0, 1, 0, 1, act_ready, 3, free_ready);
      }
      seg(s.img_ws, (size_t)(KA / 16) * 256 * 64, 256, KA / 16, 1, 0, 1, 0, 1, 0xff, 4, 3);
      seg(s.img_wh, 0, 32, 16, 8, 0, 1, 0, 1, 3, 0xff, 4);
      seg(s.img_wh, (size_t)16 * 32 * 64, 32, 16, 8, 0, 1, 256, 0, 4, 5);
      c.n_seg = n;
      continue;
    }
    // accumulator regions (TMEM columns): single-CTA form 0 / 256 (an N = 256 step takes 256 columns), pair form 0 / 128 /
    // 256 (attention) / 384 (head) (N / 2 columns per step)
    const int colB = pair ? 128 : 256, colAtt = pair ? 256 : 0, colHead = pair ? 384 : 0;
    seg(s.img_w1, 0, 256, k1s, 1, 0, 0, 0, 1, 0xff, 0);
    seg(s.img_w1, (size_t)k1s * 256 * 64, 256, k1s, 1, colB, 0, 0, 1, 0xff, 1);
    int act_ready = 1;
    if (s.att_n > 0) {
      const int k0 = s.att_lo & ~7, ks = (s.att_lo + s.att_n - k0 + 15) / 16;
      seg(s.img_att, 0, att_nn, ks, 1, colAtt, 1, k0, 1, 1, 2);
      act_ready = 2;
    }
    seg(s.img_ws, 0, 256, KA / 16, 1, 0, 1, 0, 1, act_ready, 3);
    seg(s.img_ws, (size_t)(KA / 16) * 256 * 64, 256, KA / 16, 1, colB, 1, 0, 1, 0xff, 4);
    seg(s.img_wh, 0, 32, 16, 8, colHead, 1, 0, 1, 3, 0xff);
    seg(s.img_wh, (size_t)16 * 32 * 64, 32, 16, 8, colHead, 1, 256, 0, 4, 5);
    c.n_seg = n;
  }
  a.prof = g_prof;
  a.debug = g_debug;
  a.zero = 0;
  static bool opted_dev[64] = {};
  int dev = 0;
  cudaError_t ce = cudaGetDevice(&dev);
  if (ce != cudaSuccess || dev < 0 || dev >= 64) {
    err = "hh_policy_forward_ex: cudaGetDevice failed";
    return -2;
  }
  if (!opted_dev[dev]) {
    ce = cudaFuncSetAttribute(policy_forward_tc_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_BYTES);
    if (ce == cudaSuccess) ce = cudaFuncSetAttribute(policy_forward_pair_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_BYTES);
    if (ce == cudaSuccess) ce = cudaFuncSetAttribute(policy_forward_m128_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM2_BYTES);
    if (ce != cudaSuccess) {
      err = std::string("cudaFuncSetAttribute(policy_forward kernels): ") + cudaGetErrorString(ce);
      return -2;
    }
    opted_dev[dev] = true;
  }
  const int tiles = (max_rows + TM - 1) / TM;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (m128) policy_forward_m128_kernel<<<dim3((unsigned)((max_rows + TM2 - 1) / TM2), (unsigned)n_chains), kThreads2, SMEM2_BYTES, st>>>(a);
  else if (pair) policy_forward_pair_kernel<<<dim3((unsigned)((tiles + 1) / 2 * 2), (unsigned)n_chains), kThreads, SMEM_BYTES, st>>>(a);
  else policy_forward_tc_kernel<1><<<dim3((unsigned)tiles, (unsigned)n_chains), kThreads, SMEM_BYTES, st>>>(a);
  ce = cudaGetLastError();
  if (ce != cudaSuccess) {
    err = std::string("policy_forward (tcgen05) launch: ") + cudaGetErrorString(ce);
    return -2;
  }
  return 0;
}
