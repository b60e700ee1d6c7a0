// hh_emu.cpp -- TEST INFRASTRUCTURE: the host subset of the C ABI (include/hhmarl_b200.h) implemented by running
// the v4 step schedule of hhmarl_2d_b200/csrc/hh_v4.cuh -- the very stage functions the CUDA kernel is made of --
// on a CPU, CTA by CTA, role by role.  tests/test_emu_v4.py drives it with the same parity checks as the GPU tests
// so that the semantics of a kernel revision are pinned to the oracle before any GPU time is spent, and so that
// cross-thread hazards between roles of one stage show up (the roles' threads can be run in reverse order).
// It is NOT a fallback: hhmarl_2d_b200 never loads it (the package raises if libhhmarl_b200.so is missing).
#include "hh_host_shim.h"

#include <string.h>

#include <new>
#include <string>
#include <vector>

#include "../../hhmarl_2d_b200/csrc/hh_v4.cuh"
#include "../../hhmarl_2d_b200/csrc/hh_state_pack.h"

using namespace hh;

static thread_local std::string g_err;
static int fail(int code, const char* msg) {
  g_err = msg;
  return code;
}

struct hh_env {
  hh_config cfg;
  int n = 0;
  Params P{};
  StatePtrs S{};
  std::vector<double> f64[6], r[4];
  std::vector<uint2> ac;
  std::vector<uint32_t> ri;
  std::vector<uint4> meta;
  std::vector<unsigned long long> dg;
  bool initialised = false;
  uint64_t launches = 0;
  v4::Smem* sm = nullptr;
};

static int obs_dim(const hh_config& c, int agent) {
  if (c.agent_mode == 0) return agent == 1 ? OBS_AC1 : OBS_AC2;
  return agent == 1 ? OBS_ESC_AC1 : OBS_ESC_AC2;
}

extern "C" const char* hh_last_error(void) { return g_err.c_str(); }
extern "C" const char* hh_version(void) { return "hhmarl_2d_b200 v4 step schedule, CPU emulation (test harness)"; }
extern "C" void hh_emu_set_reverse(int on) { v4::emu_reverse = on != 0; }
extern "C" int hh_emu_smem_bytes(void) { return (int)sizeof(v4::Smem); }

extern "C" int hh_create(const hh_config* cfg, int32_t n_arenas, int32_t, hh_env** out) {
  if (!cfg || !out) return fail(-1, "hh_create: null argument");
  if (n_arenas <= 0) return fail(-1, "hh_create: n_arenas must be positive");
  if (cfg->level < 1 || cfg->level > 3) return fail(-1, "hh_create (emulation): levels 1..3 only");
  hh_env* e = new hh_env();
  e->cfg = *cfg;
  e->n = n_arenas;
  const size_t N = (size_t)n_arenas;
  for (auto& v : e->f64) v.assign(N * 4, 0.0);
  for (auto& v : e->r) v.assign(N * 2, 0.0);
  e->ac.assign(N * 4, uint2{0, 0});
  e->ri.assign(N * 2, 0u);
  e->meta.assign(N, uint4{0, 0, 0, 0});
  e->dg.assign(N, 0ull);
  e->S.lat = e->f64[0].data(); e->S.lon = e->f64[1].data(); e->S.hdg = e->f64[2].data(); e->S.spd = e->f64[3].data();
  e->S.nhdg = e->f64[4].data(); e->S.nspd = e->f64[5].data();
  e->S.acint = e->ac.data();
  e->S.rlat = e->r[0].data(); e->S.rlon = e->r[1].data(); e->S.rhdg = e->r[2].data(); e->S.rnhdg = e->r[3].data();
  e->S.rint = e->ri.data();
  e->S.meta = e->meta.data();
  e->S.draws_g = e->dg.data();
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
  P.geom = make_geom(P.map_size);
  P.short_moves = 1;
  e->sm = new v4::Smem();
  *out = e;
  return 0;
}
extern "C" void hh_destroy(hh_env* e) {
  if (!e) return;
  delete e->sm;
  delete e;
}
extern "C" int32_t hh_n_arenas(const hh_env* e) { return e ? e->n : 0; }
extern "C" int32_t hh_obs_dim(const hh_env* e, int32_t agent) {
  if (!e || (agent != 1 && agent != 2)) return 0;
  return obs_dim(e->cfg, agent);
}
extern "C" uint64_t hh_launch_count(const hh_env* e) { return e ? e->launches : 0; }

static int n_blocks(const hh_env* e) { return (e->n + v4::kArenas - 1) / v4::kArenas; }

extern "C" int hh_reset_host(hh_env* e, const uint8_t* mask, float* obs1, float* obs2) {
  if (!e) return fail(-1, "hh_reset_host: null env");
  const int first = e->initialised ? 0 : 1;
  for (int b = 0; b < n_blocks(e); ++b) {
    memset(e->sm, 0xCD, sizeof(v4::Smem));   // poison: a stage must not rely on what another launch left behind
    if (e->cfg.agent_mode == 0) v4::reset_body<0>(*e->sm, e->S, e->P, mask, first, obs1, obs2, b);
    else v4::reset_body<1>(*e->sm, e->S, e->P, mask, first, obs1, obs2, b);
  }
  e->initialised = true;
  e->launches += 1;
  return 0;
}

template <int LEVEL>
static void run_step(hh_env* e, const int32_t* actions, float* obs1, float* obs2, float* rew, uint8_t* done) {
  for (int b = 0; b < n_blocks(e); ++b) {
    memset(e->sm, 0xCD, sizeof(v4::Smem));
    if (e->cfg.agent_mode == 0) v4::step_body<LEVEL, 0>(*e->sm, e->S, e->P, actions, obs1, obs2, rew, done, b);
    else v4::step_body<LEVEL, 1>(*e->sm, e->S, e->P, actions, obs1, obs2, rew, done, b);
  }
}

extern "C" int hh_step_host(hh_env* e, const int32_t* actions, float* obs1, float* obs2, float* rew, uint8_t* done) {
  if (!e) return fail(-1, "hh_step_host: null env");
  if (!e->initialised) return fail(-4, "hh_step: call hh_reset first");
  if (!actions) return fail(-1, "hh_step_host: null actions");
  switch (e->cfg.level) {
    case 1: run_step<1>(e, actions, obs1, obs2, rew, done); break;
    case 2: run_step<2>(e, actions, obs1, obs2, rew, done); break;
    default: run_step<3>(e, actions, obs1, obs2, rew, done); break;
  }
  e->launches += 1;
  return 0;
}

extern "C" int hh_get_state(hh_env* e, hh_state_view* o) {
  if (!e || !o) return fail(-1, "hh_get_state: null argument");
  const size_t N = (size_t)e->n;
  double* ddst[6] = {o->lat, o->lon, o->heading, o->speed, o->new_heading, o->new_speed};
  for (int k = 0; k < 6; ++k)
    if (ddst[k]) memcpy(ddst[k], e->f64[k].data(), N * 4 * sizeof(double));
  double* rdst[4] = {o->r_lat, o->r_lon, o->r_heading, o->r_new_heading};
  for (int k = 0; k < 4; ++k)
    if (rdst[k]) memcpy(rdst[k], e->r[k].data(), N * 2 * sizeof(double));
  unpack_state(N, e->ac.data(), e->ri.data(), e->meta.data(), e->dg.data(), o);
  return 0;
}

extern "C" int hh_set_state(hh_env* e, const hh_state_view* in) {
  if (!e || !in) return fail(-1, "hh_set_state: null argument");
  const size_t N = (size_t)e->n;
  const double* dsrc[6] = {in->lat, in->lon, in->heading, in->speed, in->new_heading, in->new_speed};
  for (int k = 0; k < 6; ++k) memcpy(e->f64[k].data(), dsrc[k], N * 4 * sizeof(double));
  const double* rsrc[4] = {in->r_lat, in->r_lon, in->r_heading, in->r_new_heading};
  for (int k = 0; k < 4; ++k) memcpy(e->r[k].data(), rsrc[k], N * 2 * sizeof(double));
  pack_state(N, in, e->ac.data(), e->ri.data(), e->meta.data(), e->dg.data());
  e->initialised = true;
  return 0;
}

// geodesic spot checks: mode 0 direct (lat, lon, azi, s12) -> (lat2, lon2); mode 3 direct_short with the
// sine / cosine of the azimuth computed the way the step kernel does for aircraft (heading vector)
extern "C" int hh_debug_geodesic(int32_t mode, int32_t n, const double* in, double* out) {
  for (int i = 0; i < n; ++i) {
    const double* p = in + 4 * (size_t)i;
    double2 q;
    if (mode == 0) q = geo::direct(p[0], p[1], p[2], p[3]);
    else if (mode == 1) q = geo::inverse(p[0], p[1], p[2], p[3]);
    else if (mode == 2) q = geo::inverse_local(p[0], p[1], p[2], p[3]);
    else if (mode == 3) {
      const HVec hv = heading_vec(p[2]);
      q = geo::direct_short(p[0], p[1], p[2], hv.c, hv.s, p[3]);
    } else {
      double sa, ca;
      geo::sincosd(geo::ang_round(geo::ang_normalize(p[2])), sa, ca);
      q = geo::direct_short(p[0], p[1], p[2], sa, ca, p[3]);
    }
    out[2 * (size_t)i] = q.x;
    out[2 * (size_t)i + 1] = q.y;
  }
  return 0;
}

// scalar helpers of hh_core.cuh, one call per row of in[n][6] -> out[n][2] (property tests, tests/test_emu_v4.py)
extern "C" int hh_emu_scalar(int32_t op, int32_t n, const double* in, double* out) {
  const Geom g = make_geom(0.3);
  for (int i = 0; i < n; ++i) {
    const double* p = in + 6 * (size_t)i;
    double a = 0.0, b = 0.0;
    switch (op) {
      case 0: a = pymod(p[0], p[1]); break;
      case 1: a = signed_heading_diff(p[0], p[1]); break;
      case 2: a = normalize_angle(p[0]); break;
      case 3: a = angle_in_radar_range(p[0], p[1]) ? 1.0 : 0.0; break;
      case 4: a = hdg_feature(p[0]); break;
      case 5: a = focus_deg(heading_vec(p[0]), p[1], p[2], p[3], p[4]); break;
      case 6: {
        a = (double)correct_angle_sign(p[1], p[2], p[0], p[3], p[4]);
        double sx, cx;
        angle_sign_offsets(p[0], sx, cx);
        b = (double)angle_sign_from(p[1], p[2], sx, cx, p[3], p[4]);
        break;
      }
      case 7: a = rocket_speed((int)p[0]); break;
      case 8: rel_pos(g, p[0], p[1], a, b); break;
      case 9: a = in_boundary(g, p[0], p[1]) ? 1.0 : 0.0; break;
      case 10: a = hdiff_norm(heading_vec(p[0]), heading_vec(p[1])); break;
      case 11: a = fmod_small(p[0], p[1]); break;
      default: return -1;
    }
    out[2 * (size_t)i] = a;
    out[2 * (size_t)i + 1] = b;
  }
  return 0;
}
