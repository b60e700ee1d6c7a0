// hh_v4.cuh -- "v4" fused step for levels 1-3: the step as a short chain of STAGES, each stage a set of ROLES that
// run concurrently on disjoint warp ranges of a 256-thread CTA (32 arenas staged in shared memory).
//
// v3 (hh_cta.cuh) already maps every phase to the threads that have work in it, but its phases run one after the
// other, and at 8 192 arenas the launch time is the length of ONE CTA's dependent instruction chain (~6.9 k warp
// instructions at ~10 cycles each, profiles/r1h_step_kernel_ncu.md) while the SM sits at 17 % issue utilisation.
// v4 shortens the chain instead of the instruction count:
//
//   * the G-stream random numbers a step can consume (<= 13: SURVEY A.5 draw sites) are drawn AHEAD by 4 threads per
//     arena into a shared-memory table (Philox is counter based), concurrently with the pre-tick relations; the
//     arena-serial stages (scripted opponents, missile noise) then only index the table -- same draws, same order;
//   * the rocket pipeline (heading noise, rate limit, WGS84 move, proximity geometry) runs on its own warps next to
//     the aircraft pipeline; the move is speculative and committed only if the rocket survives the resolution;
//   * kill resolution, rocket resolution, out-of-bounds, rewards and termination are ONE arena-mapped stage;
//   * the observation is split into per-unit features (unit-mapped), the 8 (fight) / 14 (escape) arccos-bearing pair
//     features per arena (one thread each) and a copy-only assembly;
//   * a unit's move uses geo::direct_short with the heading vector that the observation needs anyway (one sincos
//     per unit and step instead of three).
//
// Every stage is a function of (context, role-local thread index); HH_ROLE / HH_BARRIER turn the schedule into
// `if (tid in range)` + __syncthreads() on the device and into plain loops on the host, which is how
// tests/emu/ runs this very file against the oracle on a CPU (test infrastructure; the product is the CUDA build).
// Semantics, RNG draw order and state layout are those of hh_cta.cuh / hh_quad.cuh.
#pragma once
#include "hh_quad.cuh"

namespace hh {
namespace v4 {

#ifndef HH_V4_ARENAS
#define HH_V4_ARENAS 32
#endif
constexpr int kArenas = HH_V4_ARENAS;   // arenas per CTA (multiple of 8: every role covers whole warps)
constexpr int A1 = kArenas, A2 = 2 * kArenas, A4 = 4 * kArenas, A8 = 8 * kArenas;
constexpr int kThreads = A8;
constexpr int kDraws = 16;              // G draws prepared per arena and step (max consumed: 13)
constexpr int kPF = 16;                 // pair-feature slots per arena

struct Smem {
  // aircraft, index = local arena * 4 + unit
  double lat[A4], lon[A4], hdg[A4], spd[A4], nhdg[A4], nspd[A4];
  double nlat[A4], nlon[A4];
  double hvc[A4], hvs[A4], hvn[A4];     // heading vector of the post-tick heading
  double rel_focus[A4], near_dn[A4];
  // rockets, index = local arena * 2 + slot (slot = AC1 shooter unit / 2)
  double rlat[A2], rlon[A2], rhdg[A2], rnhdg[A2], rmlat[A2], rmlon[A2], rmh[A2];
  double rnd[A1 * kDraws];              // G-stream draws dg0 + 0 .. dg0 + 15
  double pf[A1 * kPF];                  // pair features, degrees (hdiff: normalised)
  double pdn[A4];                       // [al*4 + au*2 + {0,1}] normalised distance agent -> nearest / second enemy
  double rew[A2], opp_focus[A2];
  double sgn_x[A2], sgn_y[A2];          // level 3: heading offsets of the opponents' turn-sign test
  unsigned long long dg[A1], dg0[A1], dg_out[A1];
  alignas(16) int4 act[A2];
  alignas(16) float obs1[A1 * OBS_ESC_AC1];
  alignas(16) float obs2[A1 * OBS_ESC_AC2];
  alignas(16) float uf[A4 * 4];         // per-unit features: lat_rel, lon_rel, speed, heading
  int crem[A4], burst[A4], cmax[A4], mrem[A4], rmax[A4], mwait[A4], ota[A4];
  int near_t[A4], inr[A4];
  int po[A4];                           // [al*4 + au*2 + {0,1}] nearest / second enemy unit (-1: none)
  int rage[A2], rtgt[A2], rid[A2];
  int steps[A1], alive_ag[A1], alive_op[A1], esc_time[A1], next_id[A1], pset[A1], opp_mode[A1], err[A1], escaping[A1];
  int dgi[A1];                          // G draws consumed by the action phase
  int done[A1], alive_fin[A1], alive_pre[A1];
  unsigned int dc[A1];
  unsigned char alive[A4], hasm[A4], upd[A4], firing[A4], launched[A4], shot[A4], oob[A4];
  unsigned char ralive[A2], rocket0[A2], hit_t[A2], hit_f[A2], exploded[A2];
  unsigned char want0[A1];
  int any_reset;                        // some arena of the CTA auto-resets in this launch
};

struct Ctx {
  Smem& S;
  const StatePtrs& G;
  const Params& P;
  const int32_t* actions;
  float *obs1, *obs2, *rew_out;
  uint8_t* done_out;
  int arena0, n_valid;
  Geom g;
  // optional second destination of the observations (DUAL launches only): the next tick's central-critic rows
  // [7 action columns | own obs | other obs] of both policies (central_critic_observer, train_hetero.py:162-181), row stride cen_ld
  float* cen1 = nullptr;
  float* cen2 = nullptr;
  int cen_ld = 0;
  // arenas beyond N shadow the last valid one (they compute, never store)
  __device__ __forceinline__ int arena_of(int al) const { return arena0 + (al < n_valid ? al : n_valid - 1); }
  __device__ __forceinline__ Rng rng(int al) const {
    return Rng{P.seed_lo, P.seed_hi, P.arena_base + (uint32_t)arena_of(al)};
  }
  // G-stream draw number i of this step (table, or on the spot beyond it)
  __device__ __forceinline__ double draw(int al, int i) const {
    return i < kDraws ? S.rnd[al * kDraws + i] : g_random_at(rng(al), S.dg0[al] + (unsigned long long)i);
  }
};

__device__ __forceinline__ HVec sm_hv(const Smem& S, int i) {
  HVec h;
  h.c = S.hvc[i];
  h.s = S.hvs[i];
  h.n = S.hvn[i];
  return h;
}
__device__ __forceinline__ int alive_mask(const Smem& S, int b) {
  return S.alive[b] | (S.alive[b + 1] << 1) | (S.alive[b + 2] << 2) | (S.alive[b + 3] << 3);
}
// nearest live enemy of unit u (arena base b), env_base.py:400-422; ties -> lower id
__device__ __forceinline__ int sm_nearest(const Smem& S, const Geom& g, int b, int u, int alive_m, double& dn) {
  const int e0 = u < 2 ? 2 : 0, e1 = e0 + 1;
  const double d0 = g.inv_diag * dist_raw(S.lat[b + u], S.lon[b + u], S.lat[b + e0], S.lon[b + e0]);
  const double d1 = g.inv_diag * dist_raw(S.lat[b + u], S.lon[b + u], S.lat[b + e1], S.lon[b + e1]);
  int best = -1;
  dn = 0.0;
  if ((alive_m >> e0) & 1) { best = e0; dn = d0; }
  if (((alive_m >> e1) & 1) && (best < 0 || d1 < d0)) { best = e1; dn = d1; }
  return best;
}

// Bulky code that a stage executes rarely is kept OUT OF LINE: a stage's common path then is a compact, sequential
// instruction stream (the step is bound by instruction fetch), not a series of jumps over it.
static __device__ __noinline__ bool cold_cannon_range(double lat_s, double lon_s, double hdg_s, double lat_t, double lon_t,
                                                      double range_km, double half_width) {
  return unit_in_cannon_range(lat_s, lon_s, hdg_s, lat_t, lon_t, range_km, half_width);
}
static __device__ __noinline__ bool cold_within_1km(double lat_r, double lon_r, double lat_t, double lon_t) {
  return within_1km(lat_r, lon_r, lat_t, lon_t);
}
static __device__ __noinline__ bool cold_launch_gate(double lat_s, double lon_s, double hdg_s, double lat_t, double lon_t) {
  return launch_gate(lat_s, lon_s, hdg_s, lat_t, lon_t);
}

// ===================================================================================== S0: load
__device__ __forceinline__ void s0_load_unit(const Ctx& C, int t) {
  Smem& S = C.S;
  const int al = t >> 2, u = t & 3;
  Lane L;
  load_lane(C.G, C.arena_of(al), u, L);
  S.lat[t] = L.lat; S.lon[t] = L.lon; S.hdg[t] = L.hdg; S.spd[t] = L.spd; S.nhdg[t] = L.nhdg; S.nspd[t] = L.nspd;
  S.crem[t] = L.crem; S.burst[t] = L.burst; S.cmax[t] = L.cmax; S.mrem[t] = L.mrem; S.rmax[t] = L.rmax;
  S.mwait[t] = L.mwait; S.alive[t] = L.alive; S.hasm[t] = L.hasm; S.ota[t] = L.ota;
  if ((u & 1) == 0) {
    const int rs = al * 2 + (u >> 1);
    S.rlat[rs] = L.rlat; S.rlon[rs] = L.rlon; S.rhdg[rs] = L.rhdg; S.rnhdg[rs] = L.rnhdg;
    S.ralive[rs] = L.ralive; S.rage[rs] = L.rage; S.rtgt[rs] = L.rtgt; S.rid[rs] = L.rid;
  }
  if (u == 0) {
    S.steps[al] = L.steps + 1;                    // self.steps += 1 (env_hetero.py:114)
    S.alive_ag[al] = L.alive_ag; S.alive_op[al] = L.alive_op; S.esc_time[al] = L.esc_time; S.next_id[al] = L.next_id;
    S.pset[al] = L.pset; S.opp_mode[al] = L.opp_mode; S.err[al] = L.err; S.escaping[al] = L.escaping;
    S.dg[al] = L.dg; S.dg0[al] = L.dg; S.dc[al] = L.dc;
    if (al == 0) S.any_reset = 0;
  }
}
__device__ __forceinline__ void s0_load_actions(const Ctx& C, int t) {
  C.S.act[t] = reinterpret_cast<const int4*>(C.actions)[(size_t)C.arena_of(t >> 1) * 2 + (t & 1)];
}

// ===================================================================================== S1: pre-tick relations | draws
// agents: opp_stats[i][0] = focus(opp_to_attack -> self) (env_hetero.py:169-170); opponents: nearest agent and,
// at level 3, focus(self -> agent) and the turn sign (env_hetero.py:251-260)
template <int LEVEL>
__device__ __forceinline__ void s1_pretick(const Ctx& C, int t) {
  Smem& S = C.S;
  const int al = t >> 2, u = t & 3, ub = al * 4;
  const int alive_m = alive_mask(S, ub);
  double ndn = 0.0;
  int nt = -1;
  if (u >= 2 && S.alive[t]) nt = sm_nearest(S, C.g, ub, u, alive_m, ndn);
  int rx = -1, ry = -1;
  if (u < 2) {
    const int o = S.ota[t];
    if (S.alive[t] && o != 0 && ((alive_m >> (o - 1)) & 1)) { rx = o - 1; ry = u; }
  } else if (LEVEL == 3 && nt >= 0) {
    rx = u;
    ry = nt;
  }
  double rf = 0.0;
  if (rx >= 0)
    rf = focus_deg(heading_vec(S.hdg[ub + rx]), S.lat[ub + rx], S.lon[ub + rx], S.lat[ub + ry], S.lon[ub + ry]);
  S.near_t[t] = nt;
  S.near_dn[t] = ndn;
  S.rel_focus[t] = rf;
  S.upd[t] = S.alive[t];            // CmanoSimulator.do_tick's snapshot: nothing dies in the action phase
  S.launched[t] = 0;
  if (u < 2) S.rew[al * 2 + u] = 0.0;
  if (u == 0) {
    S.alive_pre[al] = alive_m;
    // agent 1's launch attempt consumes the step's first G draw (env_base.py:228-230)
    S.want0[al] = S.alive[t] && S.act[al * 2].w != 0 && S.ota[t] != 0 && S.mrem[t] > 0 && !S.hasm[t] && S.mwait[t] == 0;
  }
}
template <int LEVEL>
__device__ __forceinline__ void s1_draws(const Ctx& C, int t) {
  constexpr int kNeed = LEVEL == 1 ? 8 : kDraws;   // level 1: 1 (agent) + 2 (opponents) + 2 (noise)
  const int al = t >> 2, j = (t & 3) * 4;
  if (j >= kNeed) return;
  const Rng rng = C.rng(al);
  const unsigned long long base = C.S.dg0[al] + (unsigned long long)j;
#pragma unroll
  for (int q = 0; q < 4; ++q) C.S.rnd[al * kDraws + j + q] = g_random_at(rng, base + q);
}
// level 3: the heading part of the opponents' turn-sign test (env_base.py:469-475) runs on the draw warps, after the
// draws: the unit warps' chain (nearest-agent scan, heading vector, focus angle) is the longer one
__device__ __forceinline__ void s1_sign_offsets(const Ctx& C, int t) {
  const int al = t >> 2, k = t & 3;
  if (k >= 2) return;
  double sx = 0.0, cx = 0.0;
  if (C.S.alive[al * 4 + 2 + k]) angle_sign_offsets(C.S.hdg[al * 4 + 2 + k], sx, cx);
  C.S.sgn_x[al * 2 + k] = sx;
  C.S.sgn_y[al * 2 + k] = cx;
}

// ===================================================================================== S2: actions
// Rafale.fire_missile (ac1.py:72-79) + Rocket.__init__ (rocket_unit.py:23-30) for AC1 unit us -> rocket slot rs
__device__ __forceinline__ void launch(Smem& S, int us, int rs, int b, int tgt) {
  if (!S.hasm[us] && S.mrem[us] > 0) {
    const int tq = tgt < 0 ? 0 : tgt;
    if (cold_launch_gate(S.lat[us], S.lon[us], S.hdg[us], S.lat[b + tq], S.lon[b + tq])) {
      S.rlat[rs] = S.lat[us]; S.rlon[rs] = S.lon[us]; S.rhdg[rs] = S.hdg[us]; S.rnhdg[rs] = S.hdg[us];
      S.ralive[rs] = 1; S.rage[rs] = 0; S.rtgt[rs] = tq + 1;
      S.hasm[us] = 1;
      S.mrem[us] -= 1;
      S.launched[us] = 1;
    }
  }
}
// agents' _take_base_action (env_base.py:214-238), one thread per agent
template <int MODE>
__device__ __forceinline__ void s2_agents(const Ctx& C, int t) {
  Smem& S = C.S;
  const int al = t >> 1, au = t & 1, b = al * 4, us = b + au;
  S.opp_focus[t] = focus_norm_from_deg(S.rel_focus[us]);
  if (!S.alive[us]) return;
  const int4 act = S.act[t];
  const double h = pymod(S.hdg[us] + (double)((act.x - 6) * 15), 360.0);
  if (h >= 360.0 || h < 0.0) atomicOr(&S.err[al], ERR_HEADING);
  S.nhdg[us] = h;
  S.nspd[us] = 100.0 + ((max_speed(au) - 100.0) / 8.0) * (double)act.y;
  if (act.z != 0 && S.crem[us] > 0) {
    const int bt = is_ac1(au) ? 5 : 3;
    S.burst[us] = S.crem[us] < bt ? S.crem[us] : bt;
    if (MODE == 1 && S.crem[us] < 90) S.rew[t] -= 0.1;
  }
  if (au == 0 && S.want0[al]) {
    const int new_wait = randint_from(7, 17, C.draw(al, 0));   // drawn iff attempted (env_base.py:228-230)
    launch(S, us, al * 2, b, S.ota[us] - 1);
    S.mwait[us] = new_wait;
    if (MODE == 1 && S.mrem[us] < 3) S.rew[t] -= 0.1;
  }
  if (S.mwait[us] > 0 && !S.hasm[us]) S.mwait[us] -= 1;        // env_base.py:235-236
}
// scripted opponents (env_hetero.py:118-158), one thread per arena, ids 3 then 4 (shared escape state, draw order)
template <int LEVEL>
__device__ __forceinline__ void s2_script(const Ctx& C, int t) {
  Smem& S = C.S;
  const int al = t, b = al * 4;
  Lane L;                      // only the arena scalars of the Lane are used by scripted_opponent_g
  L.steps = S.steps[al];
  L.escaping = S.escaping[al] != 0;
  L.esc_time = S.esc_time[al];
  L.dg = S.want0[al] ? 1 : 0;  // index into this step's draws
  auto next = [&C, al](unsigned long long i) { return C.draw(al, (int)i); };
#pragma unroll 1
  for (int k = 2; k < 4; ++k) {
    const int us = b + k;
    const bool k_alive = S.alive[us];
    const int k_near = S.near_t[us];
    int k_sign = 1;
    if (LEVEL == 3 && k_alive && k_near >= 0)
      k_sign = angle_sign_from(S.lat[us], S.lon[us], S.sgn_x[al * 2 + k - 2], S.sgn_y[al * 2 + k - 2], S.lat[b + k_near],
                               S.lon[b + k_near]);
    const OppDecision d = scripted_opponent_g<LEVEL>(L, next, C.g, k, k_alive, S.hasm[us], S.mwait[us], S.lat[us], S.lon[us],
                                                     S.hdg[us], k_near, S.near_dn[us], S.rel_focus[us], k_sign);
    if (!k_alive) continue;
    if (d.set_hs) {
      if (d.heading >= 360.0 || d.heading < 0.0) atomicOr(&S.err[al], ERR_HEADING);
      if (d.speed > max_speed(k) || d.speed < 0.0) atomicOr(&S.err[al], ERR_SPEED);
      S.nhdg[us] = d.heading;
      S.nspd[us] = d.speed;
    }
    if (d.fire) {
      const int bt = is_ac1(k) ? 5 : 3;
      S.burst[us] = S.crem[us] < bt ? S.crem[us] : bt;
    }
    if (d.want_missile) {
      if (k == 2) launch(S, us, al * 2 + 1, b, d.tgt);
      S.mwait[us] = LEVEL == 3 ? 10 : 5;   // never decremented
    }
  }
  S.escaping[al] = L.escaping;
  S.esc_time[al] = L.esc_time;
  S.dgi[al] = (int)L.dg;
}

// ===================================================================================== S4: moves
// aircraft: rate limits (ac1.py:82-99 / ac2.py:69-86), burst bookkeeping, Unit.update (cmano_simulator.py:65-72)
__device__ __forceinline__ void s4_move_unit(const Ctx& C, int t) {
  Smem& S = C.S;
  const int al = t >> 2, u = t & 3, ub = al * 4;
  if (S.launched[t]) S.rid[al * 2 + (u >> 1)] = S.next_id[al] + (u == 2 ? (int)S.launched[ub] : 0);   // ids in shooter order
  const bool upd = S.upd[t];
  const double max_deg = is_ac1(u) ? 5.0 : 3.5, max_kn = is_ac1(u) ? 35.0 : 28.0;
  double hdg = S.hdg[t], spd = S.spd[t];
  if (upd) {
    const double nh = S.nhdg[t], ns = S.nspd[t];
    if (hdg != nh) {
      const double dd = signed_heading_diff(hdg, nh);
      hdg = fabs(dd) <= max_deg ? nh : pymod(hdg + (dd >= 0.0 ? max_deg : -max_deg), 360.0);
    }
    if (spd != ns) {
      const double dd = ns - spd;
      spd = fabs(dd) <= max_kn ? ns : spd + (dd >= 0.0 ? max_kn : -max_kn);
    }
    S.hdg[t] = hdg;
    S.spd[t] = spd;
  }
  const bool firing = upd && S.burst[t] > 0;
  S.firing[t] = firing;
  if (firing) {
    S.burst[t] -= 1;
    S.crem[t] = S.crem[t] > 0 ? S.crem[t] - 1 : 0;
  }
  const HVec hv = heading_vec(hdg);     // (cos, sin) of 90 deg - heading = (sin, cos) of the azimuth
  S.hvc[t] = hv.c;
  S.hvs[t] = hv.s;
  S.hvn[t] = hv.n;
  double nlat = S.lat[t], nlon = S.lon[t];
  if (upd && spd > 0.0) {
    const double2 q = geo::direct_short(nlat, nlon, hdg, hv.c, hv.s, spd * kKnotsToMs * 1.0);
    nlat = q.x;
    nlon = q.y;
  }
  S.nlat[t] = nlat;
  S.nlon[t] = nlon;
}
// rockets: the shooter's heading noise (ac1.py:117-128), then the speculative Rocket.update move
// (rocket_unit.py:58-73); committed in S7 unless the rocket explodes or expires in this tick
__device__ __forceinline__ void s4_move_rocket(const Ctx& C, int t) {
  Smem& S = C.S;
  const int al = t >> 1, slot = t & 1, us = al * 4 + slot * 2;
  const bool r0 = S.ralive[t];          // snapshot includes rockets launched in this step
  S.rocket0[t] = r0;
  S.exploded[t] = 0;
  if (!r0) return;
  double nh = S.rnhdg[t];
  if (S.upd[us] && S.hasm[us]) {
    const int first = (slot == 1 && S.upd[al * 4] && S.hasm[al * 4] && S.ralive[al * 2]) ? 1 : 0;   // id order
    nh = clip(__dmul_rn(S.rhdg[t], uniform_from(0.95, 1.05, C.draw(al, S.dgi[al] + first))), 0.0, 359.0);
    S.rnhdg[t] = nh;
  }
  double h = S.rhdg[t];
  if (h != nh) {
    const double dd = signed_heading_diff(h, nh);
    h = fabs(dd) <= 10.0 ? nh : h + (dd >= 0.0 ? 10.0 : -10.0);   // no modulo (rocket_unit.py:62-66)
  }
  double sa, ca;
  geo::sincosd(geo::ang_round(geo::ang_normalize(h)), sa, ca);
  const double2 q = geo::direct_short(S.rlat[t], S.rlon[t], h, sa, ca, rocket_speed(S.rage[t]) * kKnotsToMs * 1.0);
  S.rmh[t] = h;
  S.rmlat[t] = q.x;
  S.rmlon[t] = q.y;
}

// ===================================================================================== S5: engagement geometry
// cannon (ac1.py:105-115): shooter u sees lower ids at their NEW position, higher ids at the OLD one
__device__ __forceinline__ void s5_cannon(const Ctx& C, int t) {
  Smem& S = C.S;
  const int al = t >> 2, u = t & 3, ub = al * 4;
  int in_range = 0;
  if (S.firing[t]) {
    const int alive0 = S.alive_pre[al];
    const double range = is_ac1(u) ? 2.0 : 4.5, half_w = (is_ac1(u) ? 10.0 : 7.0) / 2.0;
    // candidates first (three squared distances), then the full test for the few that are near: the warp makes as
    // many passes through the geodesic code as its busiest lane has candidates, usually 0 or 1
    int cand = 0;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const bool ok = j != u && ((alive0 >> j) & 1) && (C.P.friendly_kill || ((u < 2) != (j < 2)));
      const double jl = j < u ? S.nlat[ub + j] : S.lat[ub + j];
      const double jo = j < u ? S.nlon[ub + j] : S.lon[ub + j];
      if (ok && maybe_within_km(S.lat[t], S.lon[t], jl, jo, range)) cand |= 1 << j;
    }
#pragma unroll 1
    while (cand) {
      const int j = __ffs(cand) - 1;
      cand &= cand - 1;
      const double jl = j < u ? S.nlat[ub + j] : S.lat[ub + j];
      const double jo = j < u ? S.nlon[ub + j] : S.lon[ub + j];
      if (cold_cannon_range(S.lat[t], S.lon[t], S.hdg[t], jl, jo, range, half_w)) in_range |= 1 << j;
    }
  }
  S.inr[t] = in_range;
  // out-of-bounds test of the post-tick position (env_base.py:244-255); combined with the alive mask in S6
  const bool upd = S.upd[t];
  S.oob[t] = !in_boundary(C.g, upd ? S.nlat[t] : S.lat[t], upd ? S.nlon[t] : S.lon[t]);
}
// rocket proximity (rocket_unit.py:39,49): every aircraft has already moved
__device__ __forceinline__ void s5_rocket(const Ctx& C, int t) {
  Smem& S = C.S;
  const int b = (t >> 1) * 4;
  bool ht = false, hf = false;
  if (S.rocket0[t]) {
    const int tq = S.rtgt[t] > 0 ? S.rtgt[t] - 1 : 0;
    ht = maybe_within_km(S.rlat[t], S.rlon[t], S.nlat[b + tq], S.nlon[b + tq], 1.0) &&
         cold_within_1km(S.rlat[t], S.rlon[t], S.nlat[b + tq], S.nlon[b + tq]);
    if (C.P.friendly_kill)   // "friendly" is always id 2
      hf = maybe_within_km(S.rlat[t], S.rlon[t], S.nlat[b + 1], S.nlon[b + 1], 1.0) &&
           cold_within_1km(S.rlat[t], S.rlon[t], S.nlat[b + 1], S.nlon[b + 1]);
  }
  S.hit_t[t] = ht;
  S.hit_f[t] = hf;
}

// cannon kill resolution in (shooter, target) id order; C-stream draws only for live in-range targets (ac1.py:105-115)
__device__ __forceinline__ void cold_kill_resolution(const Ctx& C, int al, int& alive_m, unsigned& killer_pack) {
  Smem& S = C.S;
  const int b = al * 4;
  const Rng rng = C.rng(al);
  unsigned dc = S.dc[al];
#pragma unroll 1
  for (int k = 0; k < 4; ++k) {
    const int row = S.inr[b + k];
    if (row == 0) continue;
    const double p_hit = (k & 1) ? 0.9 / (3.0 / 1.0) : 0.75 / (5.0 / 1.0);
#pragma unroll 1
    for (int j = 0; j < 4; ++j) {
      if (((row >> j) & 1) && ((alive_m >> j) & 1)) {
        if (c_random_at(rng, dc++) < p_hit) {
          alive_m &= ~(1 << j);
          killer_pack |= (unsigned)(k + 1) << (4 * j);
        }
      }
    }
  }
  S.dc[al] = dc;
}

// ===================================================================================== S6: resolution (arena-mapped)
// kill resolution in (shooter, target) id order with C-stream draws, rockets in launch order, out-of-bounds,
// _combat_rewards / _get_rewards (env_base.py:240-310, env_hetero.py:188-225), termination (env_base.py:89-90)
template <int MODE>
__device__ __forceinline__ void s6_resolve(const Ctx& C, int t) {
  Smem& S = C.S;
  const Params& P = C.P;
  const int al = t, b = al * 4;
  int alive_m = S.alive_pre[al];
  unsigned killer_pack = 0;      // 4 bits per victim: killer id (0 = none)
  if (S.inr[b] | S.inr[b + 1] | S.inr[b + 2] | S.inr[b + 3]) cold_kill_resolution(C, al, alive_m, killer_pack);
  // missile bookkeeping of the shooters (ac1.py:117-128); the noise itself was applied in S4
  {
    int noise = 0;
#pragma unroll
    for (int s = 0; s < 2; ++s) {
      const int us = b + 2 * s;
      if (S.upd[us] && S.hasm[us]) {
        if (!S.rocket0[al * 2 + s]) S.hasm[us] = 0;
        else noise += 1;
      }
    }
    S.dg[al] = S.dg0[al] + (unsigned long long)(S.dgi[al] + noise);
  }
  S.next_id[al] += S.launched[b] + S.launched[b + 2];
  // rockets in launch (id) order (rocket_unit.py:37-56)
  int by_rocket_m = 0;
  {
    const int r0 = S.rocket0[al * 2], r2 = S.rocket0[al * 2 + 1];
    if (r0 || r2) {
      const int first = (r0 && r2) ? (S.rid[al * 2] < S.rid[al * 2 + 1] ? 0 : 1) : (r0 ? 0 : 1);
#pragma unroll 1
      for (int n = 0; n < 2; ++n) {
        const int s = n == 0 ? first : 1 - first, rs = al * 2 + s;
        if (!S.rocket0[rs]) continue;
        const int tt = S.rtgt[rs] > 0 ? S.rtgt[rs] - 1 : 0;
        if (S.hit_t[rs] && ((alive_m >> tt) & 1)) {
          alive_m &= ~(1 << tt);
          killer_pack |= (unsigned)(2 * s + 1) << (4 * tt);
          by_rocket_m |= 1 << tt;
          S.exploded[rs] = 1;
        } else if (S.hit_f[rs] && ((alive_m >> 1) & 1)) {
          alive_m &= ~2;
          killer_pack |= (unsigned)(2 * s + 1) << 4;
          by_rocket_m |= 2;
          S.exploded[rs] = 1;
        }
      }
    }
  }
  const double s = P.rew_scale;
  const int oob_m = alive_m & (S.oob[b] | (S.oob[b + 1] << 1) | (S.oob[b + 2] << 2) | (S.oob[b + 3] << 3));
  alive_m &= ~oob_m;
  const int present_m = (S.upd[b] ? 1 : 0) | (S.upd[b + 1] ? 2 : 0);   // alive at step start (reward-dict membership)
  double rews0 = 0.0, rews1 = 0.0;
  int destroyed_m = oob_m & 3;
  if (oob_m & 1) rews0 += -5.0 * s;
  if (oob_m & 2) rews1 += -5.0 * s;
  int alive_ag = S.alive_ag[al] - __popc(oob_m & 3), alive_op = S.alive_op[al] - __popc(oob_m & 12);
  if (killer_pack != 0) {
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int k = (killer_pack >> (4 * j)) & 0xF;
      if (k == 0) continue;
      double add0 = 0.0, add1 = 0.0;
      if (k <= 2) {
        if (j >= 2) {
          if (MODE == 0) {
            double r;
            if ((by_rocket_m >> j) & 1)
              r = (1.0 + 0.5 * ((double)S.mrem[b] / (double)S.rmax[b])) * s;       // only agent 1 carries missiles
            else
              r = ((0.5 + 0.5 * ((double)S.crem[b + k - 1] / (double)S.cmax[b + k - 1])) +
                   (0.5 + 0.5 * S.opp_focus[al * 2 + k - 1])) * s;
            if (k == 1) add0 = r; else add1 = r;
          }
          alive_op -= 1;
        } else {
          if (k == 1) add0 = -2.0 * s; else add1 = -2.0 * s;
          if (P.friendly_punish) {
            if (j == 0) add0 += -2.0 * s; else add1 += -2.0 * s;
            destroyed_m |= 1 << j;
          }
          alive_ag -= 1;
        }
      } else {
        if (j < 2) {
          if (j == 0) add0 = -2.0 * s; else add1 = -2.0 * s;
          destroyed_m |= 1 << j;
          alive_ag -= 1;
        } else {
          alive_op -= 1;
        }
      }
      rews0 += add0;
      rews1 += add1;
    }
  }
  if (MODE == 1 && P.esc_dist_rew) {           // env_hetero.py:198-214, on the post-tick positions
    double pl[4], po[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      pl[j] = S.upd[b + j] ? S.nlat[b + j] : S.lat[b + j];
      po[j] = S.upd[b + j] ? S.nlon[b + j] : S.lon[b + j];
    }
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      if (!((alive_m >> i) & 1)) continue;
      const double d2 = dist_raw(pl[i], po[i], pl[2], po[2]);
      const double d3 = dist_raw(pl[i], po[i], pl[3], po[3]);
      const bool has2 = (alive_m >> 2) & 1, has3 = (alive_m >> 3) & 1;
      const bool swap = has2 && has3 && (C.g.inv_diag * d3) < (C.g.inv_diag * d2);
      const double first = has2 ? (swap ? d3 : d2) : d3, second = swap ? d2 : d3;
      const int n = (int)has2 + (int)has3;
      double mine = 0.0;
      for (int j = 1; j <= n; ++j) {
        const double od = j == 1 ? first : second;
        if (od < 0.06) {
          mine += -0.02 / j;
          if (S.spd[b + i] < 200.0) mine += -0.02 / j;
        } else if (od > 0.13) {
          mine += 0.02 / j;
          if (S.spd[b + i] > 500.0) mine += 0.02 / j;
        }
      }
      if (i == 0) rews0 += mine; else rews1 += mine;
    }
  }
  double r0 = S.rew[al * 2], r1 = S.rew[al * 2 + 1];
  const bool share = P.glob_frac > 0.0 && MODE == 0;
  if ((present_m & 1) && (((alive_m >> 0) & 1) || (destroyed_m & 1))) r0 += share ? rews0 + P.glob_frac * rews1 : rews0;
  if ((present_m & 2) && (((alive_m >> 1) & 1) || (destroyed_m & 2))) r1 += share ? rews1 + P.glob_frac * rews0 : rews1;
  S.alive_ag[al] = alive_ag;
  S.alive_op[al] = alive_op;
  S.alive_fin[al] = alive_m;
  const bool done = alive_ag <= 0 || alive_op <= 0 || S.steps[al] >= P.horizon;
  S.done[al] = done ? 1 : 0;
  if (done && P.autoreset) S.any_reset = 1;      // every writer stores the same value
  if (al < C.n_valid) {
    const int a = C.arena0 + al;
    if (C.rew_out) reinterpret_cast<float2*>(C.rew_out)[a] = make_float2((float)r0, (float)r1);
    if (C.done_out) C.done_out[a] = done ? 1 : 0;
  }
}

// ===================================================================================== S7: commit, auto-reset, per-unit features
__device__ __forceinline__ void unit_features(const Ctx& C, int t) {
  Smem& S = C.S;
  const int u = t & 3;
  double x, y;
  rel_pos(C.g, S.lat[t], S.lon[t], x, y);
  float4 f;
  f.x = (float)x;
  f.y = (float)y;
  f.z = (float)clip(S.spd[t] / max_speed(u), 0.0, 1.0);
  f.w = (float)hdg_feature(S.hdg[t]);
  reinterpret_cast<float4*>(S.uf)[t] = f;
  S.shot[t] = S.burst[t] > 0 || (is_ac1(u) && S.hasm[t]);      // env_base.py:208-211
}
// HHMARLBaseEnv.reset for one unit of arena al (reset_lane, hh_quad.cuh); arena scalars by unit 0
// in-launch auto-reset: the 16 G draws from dg on were prepared by s7a_reset_draws (8 threads per arena)
__device__ __forceinline__ void s7a_reset_draws(const Ctx& C, int t) {
  Smem& S = C.S;
  const int al = t >> 3, j = (t & 7) * 2;
  if (!(S.done[al] && C.P.autoreset)) return;
  const Rng rng = C.rng(al);
  const unsigned long long base = S.dg[al] + (unsigned long long)j;
  S.rnd[al * kDraws + j] = g_random_at(rng, base);
  S.rnd[al * kDraws + j + 1] = g_random_at(rng, base + 1);
}
// (forceinline: a noinline callee taking the context by reference forces Ctx -- and with it every stage's accesses to
// it -- into local memory: 552-byte stack frame, launch 24 -> 32 us, profiles/r1o_*)
__device__ __forceinline__ void reset_unit(const Ctx& C, int t, unsigned long long dg, unsigned int dc, int err,
                                           bool from_table) {
  Smem& S = C.S;
  const int al = t >> 2, u = t & 3;
  Lane L;
  L.dg = dg;
  L.dc = dc;
  L.err = err;
  if (from_table) {
    const double* tab = S.rnd + al * kDraws;
    const Rng rng = C.rng(al);
    reset_lane_g(L, [tab, dg, &rng](unsigned long long i) { return i - dg < (unsigned long long)kDraws ? tab[i - dg] : g_random_at(rng, i); },
                 C.P, u);
  } else {
    reset_lane(L, C.rng(al), C.P, u);
  }
  S.lat[t] = L.lat; S.lon[t] = L.lon; S.hdg[t] = L.hdg; S.spd[t] = L.spd; S.nhdg[t] = L.nhdg; S.nspd[t] = L.nspd;
  S.crem[t] = L.crem; S.burst[t] = 0; S.cmax[t] = L.cmax; S.mrem[t] = L.mrem; S.rmax[t] = L.rmax; S.mwait[t] = 0;
  S.alive[t] = 1; S.hasm[t] = 0; S.ota[t] = 0;
  const HVec hv = heading_vec(L.hdg);
  S.hvc[t] = hv.c;
  S.hvs[t] = hv.s;
  S.hvn[t] = hv.n;
  if (u == 0) {
    S.steps[al] = 0; S.alive_ag[al] = 2; S.alive_op[al] = 2; S.esc_time[al] = 0; S.escaping[al] = 0; S.next_id[al] = 5;
    S.pset[al] = L.pset; S.opp_mode[al] = L.opp_mode; S.dg_out[al] = L.dg;
  }
}
__device__ __forceinline__ void s7_commit_unit(const Ctx& C, int t) {
  Smem& S = C.S;
  const int al = t >> 2, u = t & 3;
  if (S.done[al] && C.P.autoreset) {
    reset_unit(C, t, S.dg[al], S.dc[al], S.err[al], true);
  } else {
    if (S.upd[t]) {
      S.lat[t] = S.nlat[t];
      S.lon[t] = S.nlon[t];
    }
    S.alive[t] = (S.alive_fin[al] >> u) & 1;
    if (u == 0) S.dg_out[al] = S.dg[al];
  }
  unit_features(C, t);
}
__device__ __forceinline__ void s7_commit_rocket(const Ctx& C, int t) {
  Smem& S = C.S;
  const int al = t >> 1;
  if (S.done[al] && C.P.autoreset) {
    S.rlat[t] = 0.0; S.rlon[t] = 0.0; S.rhdg[t] = 0.0; S.rnhdg[t] = 0.0;
    S.ralive[t] = 0; S.rage[t] = 0; S.rtgt[t] = 0; S.rid[t] = 0;
  } else if (S.rocket0[t]) {
    if (S.exploded[t] || S.rage[t] > 10) {      // rocket_unit.py:44,54,57-59
      S.ralive[t] = 0;
    } else {
      S.rhdg[t] = S.rmh[t];
      S.rlat[t] = S.rmlat[t];
      S.rlon[t] = S.rmlon[t];
      S.rage[t] += 1;
    }
  }
}

// ===================================================================================== S8: pair features
// Per arena: for each agent (nearest enemy o) focus(self->o), focus(o->self), hdiff(self, o); the team-mate pair
// focus(1->2), focus(2->1); in escape mode the same three for the second enemy.  One task per thread; every task
// is the angle between a heading vector and either a line of sight (env_base.py:424-432) or another heading vector
// (:448-456), so the tasks only select their two vectors and share ONE arccos evaluation (no divergent copies).
template <int MODE>
__device__ __forceinline__ void s8_pairs(const Ctx& C, int t) {
  Smem& S = C.S;
  constexpr int kTasks = MODE == 0 ? 8 : 14;
  const int al = t >> 3, b = al * 4;
  const int alive_m = alive_mask(S, b);
#pragma unroll 1
  for (int j = t & 7; j < kTasks; j += 8) {
    int x = -1, y = -1;          // unit slots: angle at x towards y (line of sight) or between the headings of x and y
    bool heading_pair = false;
    if (j == 6 || j == 7) {
      if ((alive_m & 3) == 3) { x = b + (j - 6); y = b + (7 - j); }
    } else {
      const int jj = j < 6 ? j : j - 8, au = jj / 3, kind = jj - 3 * au, second = j >= 8;
      const int us = b + au;
      double dn = 0.0;
      int o = ((alive_m >> au) & 1) ? sm_nearest(S, C.g, b, au, alive_m, dn) : -1;
      if (second && o >= 0) {
        const int o2 = o == 2 ? 3 : 2;          // the other enemy, second in the sorted list if alive
        o = ((alive_m >> o2) & 1) ? o2 : -1;
        if (o >= 0) dn = C.g.inv_diag * dist_raw(S.lat[us], S.lon[us], S.lat[b + o], S.lon[b + o]);
      }
      if (kind == 0) {
        S.po[b + au * 2 + second] = o;
        S.pdn[b + au * 2 + second] = dn;
      }
      if (o >= 0) {
        x = kind == 1 ? b + o : us;
        y = kind == 1 ? us : b + o;
        heading_pair = kind == 2;
      }
    }
    double v = 0.0;
    if (x >= 0) {
      const HVec a = sm_hv(S, x);
      double w0 = S.lon[y] - S.lon[x], w1 = S.lat[y] - S.lat[x];
      double wn = sqrt(w0 * w0 + w1 * w1);
      if (heading_pair) { w0 = S.hvc[y]; w1 = S.hvs[y]; wn = S.hvn[y]; }
      const double cs = clip((a.c * w0 + a.s * w1) / (a.n * wn + 1e-10), -1.0, 1.0);
      const double deg = m::acos_(cs) * (180.0 / geo::kPi);
      v = heading_pair ? clip(deg / 180.0, 0.0, 1.0) : deg;
    }
    S.pf[al * kPF + j] = v;
  }
}

// ===================================================================================== S9: observation rows
// lowlevel_state (env_hetero.py:65-103): fight_state_values / esc_state_values (env_base.py:111-164),
// opp_ac_values (:185-212), friendly_ac_values (:166-183) assembled from the unit and pair features.
// Three threads per agent row: own part, enemy block(s), friend block.
__device__ __forceinline__ int put_unit(const Smem& S, int us, float* out) {
  const float4 f = reinterpret_cast<const float4*>(S.uf)[us];
  out[0] = f.x; out[1] = f.y; out[2] = f.z; out[3] = f.w;
  return 4;
}
template <int MODE>
__device__ __forceinline__ void s9_rows(const Ctx& C, int t) {
  Smem& S = C.S;
  constexpr int D1 = MODE == 0 ? OBS_AC1 : OBS_ESC_AC1, D2 = MODE == 0 ? OBS_AC2 : OBS_ESC_AC2;
  const int part = t / A2, ag = t - part * A2;       // part: 0 own, 1 enemies, 2 friend (warp-uniform)
  const int al = ag >> 1, au = ag & 1, b = al * 4, us = b + au;
  float* row = au == 0 ? S.obs1 + al * D1 : S.obs2 + al * D2;
  const int n_own = MODE == 0 ? (au == 0 ? 12 : 10) : (au == 0 ? 7 : 6);
  const int n_enemy = MODE == 0 ? 9 : 18;
  const int o = S.po[b + au * 2];
  if (o < 0) {                  // dead, or no enemy left: the whole row is zero (env_hetero.py:77-79)
    const int lo = part == 0 ? 0 : (part == 1 ? n_own : n_own + n_enemy);
    const int hi = part == 0 ? n_own : (part == 1 ? n_own + n_enemy : n_own + n_enemy + 5);
    for (int k = lo; k < hi; ++k) row[k] = 0.0f;
    return;
  }
  const double* pf = S.pf + al * kPF + au * 3;
  const double f_uo = pf[0], f_ou = pf[1], hd = pf[2], dn = S.pdn[b + au * 2];
  if (part == 0) {
    float* out = row;
    int n = put_unit(S, us, out);
    if (MODE == 0) {
      out[n++] = (float)focus_norm_from_deg(f_uo);
      out[n++] = (float)aspect_from_deg(f_ou);
      out[n++] = (float)hd;
      out[n++] = (float)dn;
      out[n++] = (float)clip((double)S.crem[us] / (double)S.cmax[us], 0.0, 1.0);
      if (au == 0) {
        out[n++] = (float)clip((double)S.mrem[us] / (double)S.rmax[us], 0.0, 1.0);
        out[n++] = S.mwait[us] == 0 ? 1.0f : 0.0f;
        out[n++] = (S.hasm[us] || S.burst[us] > 0) ? 1.0f : 0.0f;
      } else {
        out[n++] = S.burst[us] > 0 ? 1.0f : 0.0f;
      }
    } else {
      out[n++] = (float)clip((double)S.crem[us] / (double)S.cmax[us], 0.0, 1.0);
      if (au == 0) out[n++] = (float)clip((double)S.mrem[us] / (double)S.rmax[us], 0.0, 1.0);
      out[n++] = S.shot[us] ? 1.0f : 0.0f;
    }
  } else if (part == 1) {
    float* out = row + n_own;
    int n = put_unit(S, b + o, out);
    out[n++] = (float)hd;
    if (MODE == 0) {
      out[n++] = (float)focus_norm_from_deg(f_ou);
      out[n++] = (float)aspect_from_deg(f_uo);
    } else {
      out[n++] = (float)focus_norm_from_deg(f_uo);
      out[n++] = (float)focus_norm_from_deg(f_ou);
    }
    out[n++] = (float)dn;
    out[n++] = S.shot[b + o] ? 1.0f : 0.0f;
    if (MODE == 1) {
      const int o2 = S.po[b + au * 2 + 1];
      if (o2 >= 0) {
        const double* pq = S.pf + al * kPF + 8 + au * 3;
        n += put_unit(S, b + o2, out + n);
        out[n++] = (float)pq[2];
        out[n++] = (float)focus_norm_from_deg(pq[0]);
        out[n++] = (float)focus_norm_from_deg(pq[1]);
        out[n++] = (float)S.pdn[b + au * 2 + 1];
        out[n++] = S.shot[b + o2] ? 1.0f : 0.0f;
      } else {
        for (int z = 0; z < 9; ++z) out[n++] = 0.0f;
      }
    }
  } else {
    float* out = row + n_own + n_enemy;
    const int fri = b + (au ^ 1);
    if (S.alive[fri]) {
      const float4 f = reinterpret_cast<const float4*>(S.uf)[fri];
      out[0] = f.x;
      out[1] = f.y;
      out[2] = (float)focus_norm_from_deg(S.pf[al * kPF + 6 + au]);        // focus(self -> friend)
      out[3] = (float)focus_norm_from_deg(S.pf[al * kPF + 7 - au]);        // focus(friend -> self)
      out[4] = (float)(C.g.inv_diag * dist_raw(S.lat[us], S.lon[us], S.lat[fri], S.lon[fri]));
    } else {
      for (int k = 0; k < 5; ++k) out[k] = 0.0f;
    }
  }
}

// ===================================================================================== S10: store
__device__ __forceinline__ void s10_store_unit(const Ctx& C, int t) {
  const Smem& S = C.S;
  const StatePtrs& G = C.G;
  const int al = t >> 2, u = t & 3, b = al * 4;
  if (al >= C.n_valid) return;
  const int a = C.arena0 + al;
  const size_t i = (size_t)a * 4 + u;
  G.lat[i] = S.lat[t];
  G.lon[i] = S.lon[t];
  G.hdg[i] = S.hdg[t];
  G.spd[i] = S.spd[t];
  G.nhdg[i] = S.nhdg[t];
  G.nspd[i] = S.nspd[t];
  uint2 w;
  w.x = (uint32_t)S.crem[t] | ((uint32_t)S.burst[t] << 16) | ((uint32_t)S.mrem[t] << 24);
  w.y = (uint32_t)S.cmax[t] | ((uint32_t)S.mwait[t] << 16) | ((uint32_t)S.rmax[t] << 24) | ((uint32_t)S.alive[t] << 28) |
        ((uint32_t)S.hasm[t] << 29);
  G.acint[i] = w;
  if ((u & 1) == 0) {
    const int rs = al * 2 + (u >> 1);
    const size_t r = (size_t)a * 2 + (u >> 1);
    G.rlat[r] = S.rlat[rs];
    G.rlon[r] = S.rlon[rs];
    G.rhdg[r] = S.rhdg[rs];
    G.rnhdg[r] = S.rnhdg[rs];
    G.rint[r] = (uint32_t)S.ralive[rs] | ((uint32_t)S.rage[rs] << 1) | ((uint32_t)S.rtgt[rs] << 5) | ((uint32_t)S.rid[rs] << 8);
  }
  if (u == 0) {
    // agents store target id - 2, opponents the id itself (load_lane)
    // opp_to_attack after lowlevel_state: the agents' nearest enemy (S8), None for the opponents at levels 1-3
    uint32_t packed = 0;
#pragma unroll
    for (int k = 0; k < 2; ++k) {
      const int o = S.po[b + k * 2];          // enemy unit 2 / 3, or -1
      packed |= (uint32_t)(o < 0 ? 0 : o - 1) << (2 * k);
    }
    uint4 m;
    m.x = (uint32_t)S.steps[al] | ((uint32_t)S.alive_ag[al] << 16) | ((uint32_t)S.alive_op[al] << 20) |
          ((uint32_t)(S.escaping[al] != 0) << 24) | ((uint32_t)S.pset[al] << 25) | ((uint32_t)S.opp_mode[al] << 28);
    m.y = (uint32_t)S.esc_time[al] | ((uint32_t)S.next_id[al] << 8) | (packed << 24);
    m.z = S.dc[al];
    m.w = (uint32_t)S.err[al];
    G.meta[a] = m;
    G.draws_g[a] = S.dg_out[al];
  }
}
// contiguous, coalesced copy of the staged observation rows (16-byte aligned on both sides).
// HH_V4_ROLLED (variant build, round-2 experiment): the copy loops make at most one trip (<= 208 float4 for 256 threads), the
// compiler's 4x unrolling only spreads ~80 executed instructions over 6 KB of code (profiles/README.md: 16 % of the kernel's
// no_instruction samples)
#if defined(HH_V4_ROLLED) && defined(__CUDACC__)
#define HH_NO_UNROLL _Pragma("unroll 1")
#else
#define HH_NO_UNROLL
#endif
template <bool CENTRAL = false>
__device__ __forceinline__ void s10_store_rows(const Ctx& C, int t, int n_threads, int d1, int d2) {
  const Smem& S = C.S;
  if (CENTRAL && C.cen1) {
    const int w = d1 + d2, nf = C.n_valid * w;
    HH_NO_UNROLL
    for (int k = t; k < nf; k += n_threads) {
      const int a = k / w, c = k - a * w;
      const float v = c < d1 ? S.obs1[a * d1 + c] : S.obs2[a * d2 + (c - d1)];
      const size_t row = (size_t)(C.arena0 + a) * C.cen_ld + 7;
      C.cen1[row + c] = v;                                   // [.. | obs1 | obs2]
      C.cen2[row + (c < d1 ? d2 + c : c - d1)] = v;          // [.. | obs2 | obs1]
    }
  }
  if (C.obs1) {
    const int nf = C.n_valid * d1, n4 = nf >> 2;
    float* dst = C.obs1 + (size_t)C.arena0 * d1;
    HH_NO_UNROLL
    for (int k = t; k < n4; k += n_threads) reinterpret_cast<float4*>(dst)[k] = reinterpret_cast<const float4*>(S.obs1)[k];
    HH_NO_UNROLL
    for (int k = (n4 << 2) + t; k < nf; k += n_threads) dst[k] = S.obs1[k];
  }
  if (C.obs2) {
    const int nf = C.n_valid * d2, n4 = nf >> 2;
    float* dst = C.obs2 + (size_t)C.arena0 * d2;
    HH_NO_UNROLL
    for (int k = t; k < n4; k += n_threads) reinterpret_cast<float4*>(dst)[k] = reinterpret_cast<const float4*>(S.obs2)[k];
    HH_NO_UNROLL
    for (int k = (n4 << 2) + t; k < nf; k += n_threads) dst[k] = S.obs2[k];
  }
}

// ===================================================================================== schedules
// A role is a half-open range of CTA threads [lo, lo + n) executing `call` with the role-local index t.
#if defined(__CUDACC__)
#define HH_ROLE(lo, n, call)                               \
  if (tid >= (lo) && tid < (lo) + (n)) {                   \
    const int t = tid - (lo);                              \
    call;                                                  \
  }
// DUAL (a template parameter of step_body): two 256-thread sub-blocks of kArenas arenas each share one 512-thread CTA -- the same
// per-arena code, each sub-block with its own Smem and its own named barrier.  Halves the number of CTAs of a launch, which
// matters when the launch shares the GPU with the policy forward's whole-SM CTAs (the sampler's grouped mode).
#define HH_BARRIER()                                                                       \
  do {                                                                                     \
    if (DUAL) asm volatile("bar.sync %0, %1;" ::"r"(1 + hh_sub), "n"(8 * kArenas) : "memory"); \
    else __syncthreads();                                                                  \
  } while (0)
#define HH_TID_DECL                                                    \
  const int tid = DUAL ? (int)(threadIdx.x % (8 * kArenas)) : (int)threadIdx.x; \
  const int hh_sub = DUAL ? (int)(threadIdx.x / (8 * kArenas)) : 0;    \
  (void)hh_sub;
#else   // host emulation (tests/emu): roles of a stage run one after the other, threads in emu_order
#define HH_ROLE(lo, n, call)                               \
  for (int t_ = 0; t_ < (n); ++t_) {                       \
    const int t = ::hh::v4::emu_reverse ? (n) - 1 - t_ : t_; \
    call;                                                  \
  }
#define HH_BARRIER()
#define HH_TID_DECL
static bool emu_reverse = false;
#endif

// Optional stage clocks (build with -DHH_V4_PROFILE; profiles/ only): thread 0 of each CTA records clock64() at
// every stage boundary, hh_debug_v4_profile() reads them back.
#if defined(HH_V4_PROFILE) && defined(__CUDACC__)
__device__ long long g_stage_clock[4096 * 16];
__device__ long long g_warp_arrive[512 * 8 * 16];   // [CTA][warp][barrier]: when each warp reached each barrier
#define HH_MARK(k) if (tid == 0 && block < 4096) g_stage_clock[block * 16 + (k)] = clock64();
#undef HH_BARRIER
#define HH_BARRIER()                                                                                          \
  do {                                                                                                        \
    if ((tid & 31) == 0 && block < 512 && hh_bar < 16) g_warp_arrive[(block * 8 + (tid >> 5)) * 16 + hh_bar] = clock64(); \
    ++hh_bar;                                                                                                 \
    __syncthreads();                                                                                          \
  } while (0)
#undef HH_TID_DECL
#define HH_TID_DECL const int tid = threadIdx.x; int hh_bar = 0; (void)DUAL;
#else
#define HH_MARK(k)
#endif

// Experiment for round 2 (build with -DHH_V4_WARM, profiles/build_variants.sh; NOT in the default library): a warp that
// idles during S2 makes one dry call of geo::direct_short so that the routine's 58 instruction-cache lines are in the
// SM's L1.5 when the unit and rocket warps reach it in S4 (profiles/README.md: 41 % of the kernel's no_instruction stall
// samples sit in that routine).  The call has no side effect: its result is compared with a value it cannot take.
#if defined(HH_V4_WARM) && defined(__CUDACC__)
__device__ __forceinline__ void s2_warm(const Ctx& C, int t) {
  if (t & 31) return;   // lane 0 of the role's warp fetches for everyone
  const double2 q = geo::direct_short(5.25 + 1e-9 * (double)(C.arena0 & 7), 7.1, 10.0, 0.17364817766693033,
                                      0.984807753012208, 100.0);
  if (q.x == 1234.5) C.S.any_reset = 2;
}
#define HH_WARM_S2 HH_ROLE(A2, A1, s2_warm(C, t))
#else
#define HH_WARM_S2
#endif

template <int LEVEL, int MODE, bool DUAL = false>
__device__ __forceinline__ void step_body(Smem& S, const StatePtrs& G, const Params& P, const int32_t* __restrict__ actions,
                                          float* __restrict__ obs1, float* __restrict__ obs2,
                                          float* __restrict__ rew_out, uint8_t* __restrict__ done_out, int block,
                                          float* cen1 = nullptr, float* cen2 = nullptr, int cen_ld = 0) {
  constexpr int D1 = MODE == 0 ? OBS_AC1 : OBS_ESC_AC1, D2 = MODE == 0 ? OBS_AC2 : OBS_ESC_AC2;
  HH_TID_DECL
  const int arena0 = block * kArenas;
  const int n_valid_ = P.n_arenas - arena0;
  const Ctx C{S, G, P, actions, obs1, obs2, rew_out, done_out, arena0, n_valid_ < kArenas ? n_valid_ : kArenas,
              P.geom, cen1, cen2, cen_ld};
  HH_MARK(0)
  HH_ROLE(0, A4, s0_load_unit(C, t))
  HH_ROLE(A4, A2, s0_load_actions(C, t))
  HH_BARRIER();
  HH_MARK(1)
  HH_ROLE(0, A4, s1_pretick<LEVEL>(C, t))
  HH_ROLE(A4, A4, s1_draws<LEVEL>(C, t); if (LEVEL == 3) s1_sign_offsets(C, t))
  HH_BARRIER();
  HH_MARK(2)
  HH_ROLE(0, A2, s2_agents<MODE>(C, t))
  HH_ROLE(A4, A1, s2_script<LEVEL>(C, t))
  HH_WARM_S2
  HH_BARRIER();
  HH_MARK(3)
  HH_ROLE(0, A4, s4_move_unit(C, t))
  HH_ROLE(A4, A2, s4_move_rocket(C, t))
  HH_BARRIER();
  HH_MARK(4)
  HH_ROLE(0, A4, s5_cannon(C, t))
  HH_ROLE(A4, A2, s5_rocket(C, t))
  HH_BARRIER();
  HH_MARK(5)
  HH_ROLE(0, A1, s6_resolve<MODE>(C, t))
  HH_BARRIER();
  HH_MARK(6)
  if (S.any_reset) {                 // CTA-uniform (written before the barrier above)
    HH_ROLE(0, A8, s7a_reset_draws(C, t))
    HH_BARRIER();
  }
  HH_ROLE(0, A4, s7_commit_unit(C, t))
  HH_ROLE(A4, A2, s7_commit_rocket(C, t))
  HH_BARRIER();
  HH_MARK(7)
  HH_ROLE(0, A8, s8_pairs<MODE>(C, t))
  HH_BARRIER();
  HH_MARK(8)
  HH_ROLE(0, 3 * A2, s9_rows<MODE>(C, t))
  HH_BARRIER();
  HH_MARK(9)
  HH_ROLE(0, A4, s10_store_unit(C, t))     // (storing the state during S9 instead was measured slower: r1o)
  HH_ROLE(0, A8, s10_store_rows<DUAL>(C, t, A8, D1, D2))
  HH_MARK(10)
#if defined(HH_V4_PROFILE) && defined(__CUDACC__)
  if ((tid & 31) == 0 && block < 512) g_warp_arrive[(block * 8 + (tid >> 5)) * 16 + 15] = hh_bar;   // barriers passed
#endif
}

// Masked reset + first observation through the same stages (first_time: state is created, not loaded).
__device__ __forceinline__ bool r_selected(const Ctx& C, const uint8_t* mask, int first_time, int al) {
  return first_time || !mask || mask[C.arena_of(al)];
}
__device__ __forceinline__ void r0_load_unit(const Ctx& C, int t) {
  s0_load_unit(C, t);
  if ((t & 3) == 0) C.S.steps[t >> 2] -= 1;      // s0 counts the step that a reset does not take
}
__device__ __forceinline__ void r1_reset_unit(const Ctx& C, const uint8_t* mask, int first_time, int t) {
  Smem& S = C.S;
  const int al = t >> 2;
  if (r_selected(C, mask, first_time, al)) {
    if (first_time) {
      reset_unit(C, t, 0ull, 0u, 0, false);
      if ((t & 3) == 0) {
        S.dc[al] = 0;
        S.err[al] = 0;
      }
    } else {
      reset_unit(C, t, S.dg[al], S.dc[al], S.err[al], false);
    }
  } else {
    const HVec hv = heading_vec(S.hdg[t]);
    S.hvc[t] = hv.c;
    S.hvs[t] = hv.s;
    S.hvn[t] = hv.n;
    if ((t & 3) == 0) S.dg_out[al] = S.dg[al];
  }
  unit_features(C, t);
}
__device__ __forceinline__ void r1_reset_rocket(const Ctx& C, const uint8_t* mask, int first_time, int t) {
  Smem& S = C.S;
  if (r_selected(C, mask, first_time, t >> 1)) {
    S.rlat[t] = 0.0; S.rlon[t] = 0.0; S.rhdg[t] = 0.0; S.rnhdg[t] = 0.0;
    S.ralive[t] = 0; S.rage[t] = 0; S.rtgt[t] = 0; S.rid[t] = 0;
  }
}
template <int MODE>
__device__ __forceinline__ void reset_body(Smem& S, const StatePtrs& G, const Params& P, const uint8_t* __restrict__ mask,
                                           int first_time, float* __restrict__ obs1, float* __restrict__ obs2, int block) {
  constexpr int D1 = MODE == 0 ? OBS_AC1 : OBS_ESC_AC1, D2 = MODE == 0 ? OBS_AC2 : OBS_ESC_AC2;
  constexpr bool DUAL = false;
  HH_TID_DECL
  const int arena0 = block * kArenas;
  const int n_valid_ = P.n_arenas - arena0;
  const Ctx C{S, G, P, nullptr, obs1, obs2, nullptr, nullptr, arena0, n_valid_ < kArenas ? n_valid_ : kArenas,
              P.geom};
  if (!first_time) {
    HH_ROLE(0, A4, r0_load_unit(C, t))
  }
  HH_BARRIER();
  HH_ROLE(0, A4, r1_reset_unit(C, mask, first_time, t))
  HH_ROLE(A4, A2, r1_reset_rocket(C, mask, first_time, t))
  HH_BARRIER();
  HH_ROLE(0, A8, s8_pairs<MODE>(C, t))
  HH_BARRIER();
  HH_ROLE(0, 3 * A2, s9_rows<MODE>(C, t))
  HH_BARRIER();
  HH_ROLE(0, A4, s10_store_unit(C, t))
  HH_ROLE(0, A8, s10_store_rows(C, t, A8, D1, D2))
}

}  // namespace v4
}  // namespace hh
