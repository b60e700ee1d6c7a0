// hh_api.cu -- fused step / reset kernels and the C ABI declared in include/hhmarl_b200.h.
//
// Kernel shape: one thread per arena, 32-thread CTAs (one warp) so that the 256 warps of the
// N = 8192 headline configuration spread over all 148 SMs; the arena's 320 B of state are read
// with 16-byte coalesced loads into registers, advanced through
//   action decode -> scripted opponents -> tick (kinematics, cannon, rockets; WGS84 FP64) ->
//   rewards / out-of-bounds / termination -> (auto-reset) -> observations
// and written back once.  Observations are staged through shared memory so that the [N][26]
// and [N][24] float rows leave the SM as contiguous 16-byte stores.
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include <new>
#include <string>
#include <vector>

#include "../../include/hhmarl_b200.h"
#include "hh_env.cuh"

namespace hh {

constexpr int kThreads = 32;

template <int MODE>
struct ObsDims {
  static constexpr int D1 = MODE == 0 ? OBS_AC1 : OBS_ESC_AC1;
  static constexpr int D2 = MODE == 0 ? OBS_AC2 : OBS_ESC_AC2;
};

// contiguous, coalesced copy of `n_floats` staged floats (16-byte aligned on both sides)
__device__ __forceinline__ void flush_rows(float* __restrict__ dst, const float* __restrict__ src, int n_floats,
                                           int lane) {
  const int n4 = n_floats >> 2;
  const float4* s4 = reinterpret_cast<const float4*>(src);
  float4* d4 = reinterpret_cast<float4*>(dst);
  for (int k = lane; k < n4; k += kThreads) d4[k] = s4[k];
  for (int k = (n4 << 2) + lane; k < n_floats; k += kThreads) dst[k] = src[k];
}

template <int MODE>
__device__ __forceinline__ void write_agent_obs(Arena& A, const Geom& g, float* obs1, float* obs2, float* s1,
                                                float* s2, int arena0, int n_valid, bool valid) {
  constexpr int D1 = ObsDims<MODE>::D1, D2 = ObsDims<MODE>::D2;
  const int lane = threadIdx.x;
  if (valid) {
    HVec hv[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) hv[u] = heading_vec(A.hdg[u]);
    unit_observation<0, MODE>(A, g, hv, s1 + lane * D1);
    unit_observation<1, MODE>(A, g, hv, s2 + lane * D2);
  }
  __syncwarp();
  if (obs1) flush_rows(obs1 + (size_t)arena0 * D1, s1, n_valid * D1, lane);
  if (obs2) flush_rows(obs2 + (size_t)arena0 * D2, s2, n_valid * D2, lane);
}

// ------------------------------------------------------------------------------------------
// step: levels 1-3 (scripted opponents) in one launch
// ------------------------------------------------------------------------------------------
template <int LEVEL, int MODE>
__global__ void __launch_bounds__(kThreads)
step_kernel(StatePtrs S, Params P, const int32_t* __restrict__ actions, float* __restrict__ obs1,
            float* __restrict__ obs2, float* __restrict__ rew_out, uint8_t* __restrict__ done_out) {
  __shared__ __align__(16) float s1[kThreads * ObsDims<MODE>::D1];
  __shared__ __align__(16) float s2[kThreads * ObsDims<MODE>::D2];
  const int arena0 = blockIdx.x * kThreads;
  const int a = arena0 + threadIdx.x;
  const bool valid = a < P.n_arenas;
  const int n_valid = min(kThreads, P.n_arenas - arena0);
  const Geom g = make_geom(P.map_size);
  Arena A;
  if (valid) {
    load_arena(S, a, A);
    const Rng rng{P.seed_lo, P.seed_hi, P.arena_base + (uint32_t)a};
    const int4* ap = reinterpret_cast<const int4*>(actions) + 2 * (size_t)a;
    const int4 act0 = ap[0], act1 = ap[1];

    // ---- LowLevelEnv._take_action, env_hetero.py:105-186
    A.steps += 1;
    double rew[2] = {0.0, 0.0};
    double opp_focus[2] = {0.0, 0.0};
    const bool present[2] = {A.alive[0], A.alive[1]};
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      if (A.alive[i]) {
        const int t = A.ota[i];
        if (t != 0 && pick4(A.alive, t - 1)) {  // opp_stats[i][0], env_hetero.py:169-170
          double lat_t = pick4(A.lat, t - 1), lon_t = pick4(A.lon, t - 1);
          opp_focus[i] = focus_norm_from_deg(
              focus_deg(heading_vec(pick4(A.hdg, t - 1)), lat_t, lon_t, A.lat[i], A.lon[i]));
        }
        take_base_action<MODE>(A, rng, i, t, i == 0 ? act0 : act1, rew[i]);
      }
    }
#pragma unroll
    for (int i = 2; i < 4; ++i) {
      if (A.alive[i]) {
        if (LEVEL == 1) opp_missile_rule(A, rng, g, i);
        else if (LEVEL == 2) opp_level2(A, rng, g, i);
        else opp_level3(A, rng, g, i);
      }
    }

    // ---- CmanoSimulator.do_tick + _get_rewards
    Kills K;
#pragma unroll
    for (int j = 0; j < 4; ++j) { K.killer[j] = 0; K.by_rocket[j] = false; }
    do_tick(A, rng, P, K);
    assemble_rewards<MODE>(A, P, g, K, opp_focus, present, rew);

    // ---- HHMARLBaseEnv.step, env_base.py:89-90
    const bool done = A.alive_ag <= 0 || A.alive_op <= 0 || A.steps >= P.horizon;
    if (rew_out) reinterpret_cast<float2*>(rew_out)[a] = make_float2((float)rew[0], (float)rew[1]);
    if (done_out) done_out[a] = done ? 1 : 0;
    if (done && P.autoreset) reset_arena(A, rng, P);
  }
  write_agent_obs<MODE>(A, g, obs1, obs2, s1, s2, arena0, n_valid, valid);
  if (valid) store_arena(S, a, A);
}

// ------------------------------------------------------------------------------------------
// reset (masked)
// ------------------------------------------------------------------------------------------
template <int MODE>
__global__ void __launch_bounds__(kThreads)
reset_kernel(StatePtrs S, Params P, const uint8_t* __restrict__ mask, int first_time, float* __restrict__ obs1,
             float* __restrict__ obs2) {
  __shared__ __align__(16) float s1[kThreads * ObsDims<MODE>::D1];
  __shared__ __align__(16) float s2[kThreads * ObsDims<MODE>::D2];
  const int arena0 = blockIdx.x * kThreads;
  const int a = arena0 + threadIdx.x;
  const bool valid = a < P.n_arenas;
  const int n_valid = min(kThreads, P.n_arenas - arena0);
  const Geom g = make_geom(P.map_size);
  Arena A;
  if (valid) {
    const Rng rng{P.seed_lo, P.seed_hi, P.arena_base + (uint32_t)a};
    if (first_time) {
      A.dg = 0;
      A.dc = 0;
      A.err = 0;
      reset_arena(A, rng, P);
    } else {
      load_arena(S, a, A);
      if (!mask || mask[a]) reset_arena(A, rng, P);
    }
  }
  write_agent_obs<MODE>(A, g, obs1, obs2, s1, s2, arena0, n_valid, valid);
  if (valid) store_arena(S, a, A);
}

}  // namespace hh

// ============================================================================================
// host side: handle, C ABI
// ============================================================================================
using namespace hh;

static thread_local std::string g_last_error;

static int fail(int code, const std::string& msg) {
  g_last_error = msg;
  return code;
}
#define HH_CUDA(expr)                                                                          \
  do {                                                                                         \
    cudaError_t _e = (expr);                                                                   \
    if (_e != cudaSuccess)                                                                     \
      return fail(-2, std::string(#expr) + ": " + cudaGetErrorString(_e));                     \
  } while (0)

struct hh_env {
  hh_config cfg;
  int n = 0;
  int device = 0;
  StatePtrs S{};
  void* slab = nullptr;
  size_t slab_bytes = 0;
  Params P{};
  bool initialised = false;
  uint64_t launches = 0;
  // host-variant staging
  cudaStream_t hstream = nullptr;
  int32_t* d_actions = nullptr;
  float *d_obs1 = nullptr, *d_obs2 = nullptr, *d_rew = nullptr;
  uint8_t *d_done = nullptr, *d_mask = nullptr;
  void* pinned = nullptr;
  size_t pinned_bytes = 0;
};

static size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

static int obs_dim(const hh_config& c, int agent) {
  if (c.agent_mode == 0) return agent == 1 ? OBS_AC1 : OBS_AC2;
  return agent == 1 ? OBS_ESC_AC1 : OBS_ESC_AC2;
}

extern "C" const char* hh_last_error(void) { return g_last_error.c_str(); }
extern "C" const char* hh_version(void) { return "hhmarl_2d_b200 0.1 (sm_100a)"; }

extern "C" int hh_create(const hh_config* cfg, int32_t n_arenas, int32_t device, hh_env** out) {
  if (!cfg || !out) return fail(-1, "hh_create: null argument");
  if (n_arenas <= 0) return fail(-1, "hh_create: n_arenas must be positive");
  if (cfg->level < 1 || cfg->level > 3)
    return fail(-1, "hh_create: level must be 1..3 for the fused scripted-opponent kernel");
  if (cfg->agent_mode != 0 && cfg->agent_mode != 1) return fail(-1, "hh_create: agent_mode must be 0 or 1");
  if (!(cfg->map_size > 0)) return fail(-1, "hh_create: map_size must be positive");
  if (cfg->horizon <= 0 || cfg->horizon > 65535) return fail(-1, "hh_create: horizon out of range");
  HH_CUDA(cudaSetDevice(device));
  hh_env* e = new (std::nothrow) hh_env();
  if (!e) return fail(-3, "hh_create: out of host memory");
  e->cfg = *cfg;
  e->n = n_arenas;
  e->device = device;
  const size_t N = (size_t)n_arenas;
  // one slab, every array 256-byte aligned
  size_t off = 0;
  auto take = [&](size_t bytes) {
    size_t o = off;
    off = align_up(off + bytes, 256);
    return o;
  };
  size_t o_f64[6], o_ac, o_r[4], o_ri, o_meta, o_dg;
  for (int k = 0; k < 6; ++k) o_f64[k] = take(N * 4 * sizeof(double));
  o_ac = take(N * 4 * sizeof(uint2));
  for (int k = 0; k < 4; ++k) o_r[k] = take(N * 2 * sizeof(double));
  o_ri = take(N * 2 * sizeof(uint32_t));
  o_meta = take(N * sizeof(uint4));
  o_dg = take(N * sizeof(unsigned long long));
  e->slab_bytes = off;
  cudaError_t ce = cudaMalloc(&e->slab, e->slab_bytes);
  if (ce != cudaSuccess) {
    delete e;
    return fail(-2, std::string("cudaMalloc state slab: ") + cudaGetErrorString(ce));
  }
  cudaMemset(e->slab, 0, e->slab_bytes);
  char* base = static_cast<char*>(e->slab);
  e->S.lat = (double*)(base + o_f64[0]);
  e->S.lon = (double*)(base + o_f64[1]);
  e->S.hdg = (double*)(base + o_f64[2]);
  e->S.spd = (double*)(base + o_f64[3]);
  e->S.nhdg = (double*)(base + o_f64[4]);
  e->S.nspd = (double*)(base + o_f64[5]);
  e->S.acint = (uint2*)(base + o_ac);
  e->S.rlat = (double*)(base + o_r[0]);
  e->S.rlon = (double*)(base + o_r[1]);
  e->S.rhdg = (double*)(base + o_r[2]);
  e->S.rnhdg = (double*)(base + o_r[3]);
  e->S.rint = (uint32_t*)(base + o_ri);
  e->S.meta = (uint4*)(base + o_meta);
  e->S.draws_g = (unsigned long long*)(base + o_dg);
  Params& P = e->P;
  P.n_arenas = n_arenas;
  P.level = cfg->level;
  P.agent_mode = cfg->agent_mode;
  P.horizon = cfg->horizon;
  P.esc_dist_rew = cfg->esc_dist_rew;
  P.friendly_kill = cfg->friendly_kill;
  P.friendly_punish = cfg->friendly_punish;
  P.autoreset = cfg->autoreset;
  P.map_size = cfg->map_size;
  P.rew_scale = cfg->rew_scale;
  P.glob_frac = cfg->glob_frac;
  P.seed_lo = (uint32_t)cfg->seed;
  P.seed_hi = (uint32_t)(cfg->seed >> 32);
  P.arena_base = (uint32_t)cfg->arena_base;
  *out = e;
  return 0;
}

extern "C" void hh_destroy(hh_env* e) {
  if (!e) return;
  cudaSetDevice(e->device);
  if (e->slab) cudaFree(e->slab);
  if (e->d_actions) cudaFree(e->d_actions);
  if (e->d_obs1) cudaFree(e->d_obs1);
  if (e->d_obs2) cudaFree(e->d_obs2);
  if (e->d_rew) cudaFree(e->d_rew);
  if (e->d_done) cudaFree(e->d_done);
  if (e->d_mask) cudaFree(e->d_mask);
  if (e->pinned) cudaFreeHost(e->pinned);
  if (e->hstream) cudaStreamDestroy(e->hstream);
  delete e;
}

extern "C" int32_t hh_n_arenas(const hh_env* e) { return e ? e->n : 0; }
extern "C" int32_t hh_obs_dim(const hh_env* e, int32_t agent) {
  if (!e || (agent != 1 && agent != 2)) return 0;
  return obs_dim(e->cfg, agent);
}
extern "C" uint64_t hh_launch_count(const hh_env* e) { return e ? e->launches : 0; }

extern "C" int hh_reset(hh_env* e, const uint8_t* mask_dev, float* obs1, float* obs2, void* stream) {
  if (!e) return fail(-1, "hh_reset: null env");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int blocks = (e->n + kThreads - 1) / kThreads;
  const int first = e->initialised ? 0 : 1;
  if (e->cfg.agent_mode == 0)
    reset_kernel<0><<<blocks, kThreads, 0, st>>>(e->S, e->P, mask_dev, first, obs1, obs2);
  else
    reset_kernel<1><<<blocks, kThreads, 0, st>>>(e->S, e->P, mask_dev, first, obs1, obs2);
  HH_CUDA(cudaGetLastError());
  e->initialised = true;
  e->launches += 1;
  return 0;
}

template <int LEVEL>
static void launch_step(hh_env* e, const int32_t* actions, float* obs1, float* obs2, float* rew, uint8_t* done,
                        cudaStream_t st) {
  const int blocks = (e->n + kThreads - 1) / kThreads;
  if (e->cfg.agent_mode == 0)
    step_kernel<LEVEL, 0><<<blocks, kThreads, 0, st>>>(e->S, e->P, actions, obs1, obs2, rew, done);
  else
    step_kernel<LEVEL, 1><<<blocks, kThreads, 0, st>>>(e->S, e->P, actions, obs1, obs2, rew, done);
}

extern "C" int hh_step(hh_env* e, const int32_t* actions_dev, float* obs1, float* obs2, float* rew,
                       uint8_t* done, void* stream) {
  if (!e) return fail(-1, "hh_step: null env");
  if (!e->initialised) return fail(-4, "hh_step: call hh_reset first");
  if (!actions_dev) return fail(-1, "hh_step: null actions");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  switch (e->cfg.level) {
    case 1: launch_step<1>(e, actions_dev, obs1, obs2, rew, done, st); break;
    case 2: launch_step<2>(e, actions_dev, obs1, obs2, rew, done, st); break;
    default: launch_step<3>(e, actions_dev, obs1, obs2, rew, done, st); break;
  }
  HH_CUDA(cudaGetLastError());
  e->launches += 1;
  return 0;
}

// ------------------------------------------------------------------------------------------ host variants
static int ensure_staging(hh_env* e) {
  if (e->hstream) return 0;
  const size_t N = (size_t)e->n;
  const int d1 = obs_dim(e->cfg, 1), d2 = obs_dim(e->cfg, 2);
  HH_CUDA(cudaSetDevice(e->device));
  HH_CUDA(cudaStreamCreateWithFlags(&e->hstream, cudaStreamNonBlocking));
  HH_CUDA(cudaMalloc(&e->d_actions, N * 8 * sizeof(int32_t)));
  HH_CUDA(cudaMalloc(&e->d_obs1, N * d1 * sizeof(float)));
  HH_CUDA(cudaMalloc(&e->d_obs2, N * d2 * sizeof(float)));
  HH_CUDA(cudaMalloc(&e->d_rew, N * 2 * sizeof(float)));
  HH_CUDA(cudaMalloc(&e->d_done, N));
  HH_CUDA(cudaMalloc(&e->d_mask, N));
  e->pinned_bytes = N * (8 * sizeof(int32_t) + (d1 + d2 + 2) * sizeof(float) + 2);
  HH_CUDA(cudaMallocHost(&e->pinned, e->pinned_bytes));
  return 0;
}

extern "C" int hh_reset_host(hh_env* e, const uint8_t* mask_host, float* obs1_host, float* obs2_host) {
  if (!e) return fail(-1, "hh_reset_host: null env");
  int rc = ensure_staging(e);
  if (rc) return rc;
  const size_t N = (size_t)e->n;
  const int d1 = obs_dim(e->cfg, 1), d2 = obs_dim(e->cfg, 2);
  if (mask_host) HH_CUDA(cudaMemcpyAsync(e->d_mask, mask_host, N, cudaMemcpyHostToDevice, e->hstream));
  rc = hh_reset(e, mask_host ? e->d_mask : nullptr, e->d_obs1, e->d_obs2, e->hstream);
  if (rc) return rc;
  if (obs1_host) HH_CUDA(cudaMemcpyAsync(obs1_host, e->d_obs1, N * d1 * sizeof(float), cudaMemcpyDeviceToHost, e->hstream));
  if (obs2_host) HH_CUDA(cudaMemcpyAsync(obs2_host, e->d_obs2, N * d2 * sizeof(float), cudaMemcpyDeviceToHost, e->hstream));
  HH_CUDA(cudaStreamSynchronize(e->hstream));
  return 0;
}

extern "C" int hh_step_host(hh_env* e, const int32_t* actions_host, float* obs1_host, float* obs2_host,
                            float* rew_host, uint8_t* done_host) {
  if (!e) return fail(-1, "hh_step_host: null env");
  if (!actions_host) return fail(-1, "hh_step_host: null actions");
  int rc = ensure_staging(e);
  if (rc) return rc;
  const size_t N = (size_t)e->n;
  const int d1 = obs_dim(e->cfg, 1), d2 = obs_dim(e->cfg, 2);
  // pageable caller buffers are bounced through the handle's pinned slab so that the copies
  // are true async DMA; pinned caller buffers (cudaHostRegister'ed / torch pin_memory) would
  // also work directly.
  char* pin = static_cast<char*>(e->pinned);
  int32_t* p_act = reinterpret_cast<int32_t*>(pin);
  float* p_obs1 = reinterpret_cast<float*>(pin + N * 8 * sizeof(int32_t));
  float* p_obs2 = p_obs1 + N * d1;
  float* p_rew = p_obs2 + N * d2;
  uint8_t* p_done = reinterpret_cast<uint8_t*>(p_rew + N * 2);
  memcpy(p_act, actions_host, N * 8 * sizeof(int32_t));
  HH_CUDA(cudaMemcpyAsync(e->d_actions, p_act, N * 8 * sizeof(int32_t), cudaMemcpyHostToDevice, e->hstream));
  rc = hh_step(e, e->d_actions, e->d_obs1, e->d_obs2, e->d_rew, e->d_done, e->hstream);
  if (rc) return rc;
  if (obs1_host) HH_CUDA(cudaMemcpyAsync(p_obs1, e->d_obs1, N * d1 * sizeof(float), cudaMemcpyDeviceToHost, e->hstream));
  if (obs2_host) HH_CUDA(cudaMemcpyAsync(p_obs2, e->d_obs2, N * d2 * sizeof(float), cudaMemcpyDeviceToHost, e->hstream));
  if (rew_host) HH_CUDA(cudaMemcpyAsync(p_rew, e->d_rew, N * 2 * sizeof(float), cudaMemcpyDeviceToHost, e->hstream));
  if (done_host) HH_CUDA(cudaMemcpyAsync(p_done, e->d_done, N, cudaMemcpyDeviceToHost, e->hstream));
  HH_CUDA(cudaStreamSynchronize(e->hstream));
  if (obs1_host) memcpy(obs1_host, p_obs1, N * d1 * sizeof(float));
  if (obs2_host) memcpy(obs2_host, p_obs2, N * d2 * sizeof(float));
  if (rew_host) memcpy(rew_host, p_rew, N * 2 * sizeof(float));
  if (done_host) memcpy(done_host, p_done, N);
  return 0;
}

// ------------------------------------------------------------------------------------------ state access
namespace {
struct HostPacked {
  std::vector<double> f64[6], r[4];
  std::vector<uint2> ac;
  std::vector<uint32_t> ri;
  std::vector<uint4> meta;
  std::vector<unsigned long long> dg;
  explicit HostPacked(size_t N) : ac(N * 4), ri(N * 2), meta(N), dg(N) {
    for (auto& v : f64) v.resize(N * 4);
    for (auto& v : r) v.resize(N * 2);
  }
};
}  // namespace

extern "C" int hh_get_state(hh_env* e, hh_state_view* o) {
  if (!e || !o) return fail(-1, "hh_get_state: null argument");
  const size_t N = (size_t)e->n;
  HH_CUDA(cudaSetDevice(e->device));
  HH_CUDA(cudaDeviceSynchronize());
  HostPacked h(N);
  double* dsrc[6] = {e->S.lat, e->S.lon, e->S.hdg, e->S.spd, e->S.nhdg, e->S.nspd};
  double* ddst[6] = {o->lat, o->lon, o->heading, o->speed, o->new_heading, o->new_speed};
  for (int k = 0; k < 6; ++k)
    if (ddst[k]) HH_CUDA(cudaMemcpy(ddst[k], dsrc[k], N * 4 * sizeof(double), cudaMemcpyDeviceToHost));
  double* rsrc[4] = {e->S.rlat, e->S.rlon, e->S.rhdg, e->S.rnhdg};
  double* rdst[4] = {o->r_lat, o->r_lon, o->r_heading, o->r_new_heading};
  for (int k = 0; k < 4; ++k)
    if (rdst[k]) HH_CUDA(cudaMemcpy(rdst[k], rsrc[k], N * 2 * sizeof(double), cudaMemcpyDeviceToHost));
  HH_CUDA(cudaMemcpy(h.ac.data(), e->S.acint, N * 4 * sizeof(uint2), cudaMemcpyDeviceToHost));
  HH_CUDA(cudaMemcpy(h.ri.data(), e->S.rint, N * 2 * sizeof(uint32_t), cudaMemcpyDeviceToHost));
  HH_CUDA(cudaMemcpy(h.meta.data(), e->S.meta, N * sizeof(uint4), cudaMemcpyDeviceToHost));
  HH_CUDA(cudaMemcpy(h.dg.data(), e->S.draws_g, N * sizeof(unsigned long long), cudaMemcpyDeviceToHost));
  for (size_t a = 0; a < N; ++a) {
    for (int u = 0; u < 4; ++u) {
      const uint2 w = h.ac[a * 4 + u];
      const size_t i = a * 4 + u;
      if (o->cannon_remain) o->cannon_remain[i] = w.x & 0xFFFF;
      if (o->cannon_burst) o->cannon_burst[i] = (w.x >> 16) & 0xFF;
      if (o->missile_remain) o->missile_remain[i] = (w.x >> 24) & 0xFF;
      if (o->cannon_max) o->cannon_max[i] = w.y & 0xFFFF;
      if (o->missile_wait) o->missile_wait[i] = (w.y >> 16) & 0xFF;
      if (o->rocket_max) o->rocket_max[i] = (w.y >> 24) & 0xF;
      if (o->alive) o->alive[i] = (w.y >> 28) & 1;
      if (o->has_missile) o->has_missile[i] = (w.y >> 29) & 1;
    }
    for (int s = 0; s < 2; ++s) {
      const uint32_t w = h.ri[a * 2 + s];
      const size_t i = a * 2 + s;
      if (o->r_alive) o->r_alive[i] = w & 1;
      if (o->r_age) o->r_age[i] = (w >> 1) & 0xF;
      if (o->r_target) o->r_target[i] = (w >> 5) & 0x7;
      if (o->r_id) o->r_id[i] = (w >> 8) & 0xFFFF;
    }
    const uint4 m = h.meta[a];
    if (o->steps) o->steps[a] = m.x & 0xFFFF;
    if (o->alive_agents) o->alive_agents[a] = (m.x >> 16) & 0xF;
    if (o->alive_opps) o->alive_opps[a] = (m.x >> 20) & 0xF;
    if (o->escaping) o->escaping[a] = (m.x >> 24) & 1;
    if (o->policy_set) o->policy_set[a] = (m.x >> 25) & 0x7;
    if (o->opp_mode) o->opp_mode[a] = (m.x >> 28) & 1;
    if (o->escaping_time) o->escaping_time[a] = m.y & 0xFF;
    if (o->next_unit_id) o->next_unit_id[a] = (m.y >> 8) & 0xFF;
    if (o->opp_to_attack) {
      const int t0 = (m.y >> 24) & 3, t1 = (m.y >> 26) & 3;
      o->opp_to_attack[a * 4 + 0] = t0 ? t0 + 2 : 0;
      o->opp_to_attack[a * 4 + 1] = t1 ? t1 + 2 : 0;
      o->opp_to_attack[a * 4 + 2] = (m.y >> 28) & 3;
      o->opp_to_attack[a * 4 + 3] = (m.y >> 30) & 3;
    }
    if (o->draws_c) o->draws_c[a] = m.z;
    if (o->error) o->error[a] = (int32_t)m.w;
    if (o->draws_g) o->draws_g[a] = h.dg[a];
  }
  return 0;
}

extern "C" int hh_set_state(hh_env* e, const hh_state_view* in) {
  if (!e || !in) return fail(-1, "hh_set_state: null argument");
  const size_t N = (size_t)e->n;
  const void* need[] = {in->lat, in->lon, in->heading, in->speed, in->new_heading, in->new_speed,
                        in->cannon_remain, in->cannon_burst, in->cannon_max, in->missile_remain,
                        in->rocket_max, in->missile_wait, in->alive, in->has_missile, in->opp_to_attack,
                        in->r_lat, in->r_lon, in->r_heading, in->r_new_heading, in->r_alive, in->r_age,
                        in->r_target, in->r_id, in->steps, in->alive_agents, in->alive_opps, in->escaping,
                        in->escaping_time, in->next_unit_id, in->policy_set, in->opp_mode, in->error,
                        in->draws_g, in->draws_c};
  for (const void* p : need)
    if (!p) return fail(-1, "hh_set_state: every field of the view must be provided");
  HH_CUDA(cudaSetDevice(e->device));
  HH_CUDA(cudaDeviceSynchronize());
  HostPacked h(N);
  for (size_t a = 0; a < N; ++a) {
    for (int u = 0; u < 4; ++u) {
      const size_t i = a * 4 + u;
      uint2 w;
      w.x = (uint32_t)in->cannon_remain[i] | ((uint32_t)in->cannon_burst[i] << 16) | ((uint32_t)in->missile_remain[i] << 24);
      w.y = (uint32_t)in->cannon_max[i] | ((uint32_t)in->missile_wait[i] << 16) | ((uint32_t)in->rocket_max[i] << 24) |
            ((uint32_t)(in->alive[i] & 1) << 28) | ((uint32_t)(in->has_missile[i] & 1) << 29);
      h.ac[i] = w;
    }
    for (int s = 0; s < 2; ++s) {
      const size_t i = a * 2 + s;
      h.ri[i] = (uint32_t)(in->r_alive[i] & 1) | ((uint32_t)in->r_age[i] << 1) | ((uint32_t)in->r_target[i] << 5) |
                ((uint32_t)in->r_id[i] << 8);
    }
    uint4 m;
    m.x = (uint32_t)in->steps[a] | ((uint32_t)in->alive_agents[a] << 16) | ((uint32_t)in->alive_opps[a] << 20) |
          ((uint32_t)(in->escaping[a] & 1) << 24) | ((uint32_t)in->policy_set[a] << 25) | ((uint32_t)(in->opp_mode[a] & 1) << 28);
    const int t0 = in->opp_to_attack[a * 4 + 0], t1 = in->opp_to_attack[a * 4 + 1];
    m.y = (uint32_t)in->escaping_time[a] | ((uint32_t)in->next_unit_id[a] << 8) | ((uint32_t)(t0 ? t0 - 2 : 0) << 24) |
          ((uint32_t)(t1 ? t1 - 2 : 0) << 26) | ((uint32_t)in->opp_to_attack[a * 4 + 2] << 28) |
          ((uint32_t)in->opp_to_attack[a * 4 + 3] << 30);
    m.z = (uint32_t)in->draws_c[a];
    m.w = (uint32_t)in->error[a];
    h.meta[a] = m;
    h.dg[a] = in->draws_g[a];
  }
  const double* dsrc[6] = {in->lat, in->lon, in->heading, in->speed, in->new_heading, in->new_speed};
  double* ddst[6] = {e->S.lat, e->S.lon, e->S.hdg, e->S.spd, e->S.nhdg, e->S.nspd};
  for (int k = 0; k < 6; ++k) HH_CUDA(cudaMemcpy(ddst[k], dsrc[k], N * 4 * sizeof(double), cudaMemcpyHostToDevice));
  const double* rsrc[4] = {in->r_lat, in->r_lon, in->r_heading, in->r_new_heading};
  double* rdst[4] = {e->S.rlat, e->S.rlon, e->S.rhdg, e->S.rnhdg};
  for (int k = 0; k < 4; ++k) HH_CUDA(cudaMemcpy(rdst[k], rsrc[k], N * 2 * sizeof(double), cudaMemcpyHostToDevice));
  HH_CUDA(cudaMemcpy(e->S.acint, h.ac.data(), N * 4 * sizeof(uint2), cudaMemcpyHostToDevice));
  HH_CUDA(cudaMemcpy(e->S.rint, h.ri.data(), N * 2 * sizeof(uint32_t), cudaMemcpyHostToDevice));
  HH_CUDA(cudaMemcpy(e->S.meta, h.meta.data(), N * sizeof(uint4), cudaMemcpyHostToDevice));
  HH_CUDA(cudaMemcpy(e->S.draws_g, h.dg.data(), N * sizeof(unsigned long long), cudaMemcpyHostToDevice));
  e->initialised = true;
  return 0;
}
