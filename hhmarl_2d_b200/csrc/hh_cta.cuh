// hh_cta.cuh -- "v3" fused step for levels 1-3: a CTA of 128 threads advances 32 arenas whose complete
// state is staged in shared memory; every phase of the step is executed by exactly the threads that have
// work in it, with a CTA barrier between phases:
//
//   unit-mapped  (128 threads = 32 arenas x 4 aircraft): load/store, pre-tick relations, rate limits + the
//                WGS84 direct solve, shooter->target cannon geometry, rocket geometry / moves, reset sampling,
//                heading vectors
//   agent-mapped ( 64 threads = 32 arenas x 2 agents)  : action decode, observations
//   arena-mapped ( 32 threads = one warp)              : everything the reference serialises per arena -- scripted
//                opponents with their shared escape state and G-stream draws, kill resolution with C-stream
//                draws, rocket resolution in launch order, rewards, termination
//
// Versus the quad kernel (hh_quad.cuh), where all four lanes of a quad execute the serial parts redundantly
// and the agent-only / shooter-only parts run at 2/4 or 1/4 lane occupancy, this issues about half the warp
// instructions per arena-step, and all warps of the CTA are in the same code region at the same time (the
// step is instruction-fetch bound, profiles/README.md).  Semantics are identical (same helpers, same order of
// RNG draws); tests/test_gpu_parity.py runs both.
#pragma once
#include "hh_quad.cuh"

namespace hh {
namespace cta {

#ifndef HH_CTA_ARENAS
#define HH_CTA_ARENAS 32
#endif
constexpr int kArenas = HH_CTA_ARENAS;        // arenas per CTA (multiple of 8: whole warps in every mapping)
constexpr int kThreads = 4 * kArenas;
constexpr int A4 = 4 * kArenas, A2 = 2 * kArenas, A1 = kArenas;

struct Smem {
  double lat[A4], lon[A4], hdg[A4], spd[A4], nhdg[A4], nspd[A4];
  double rlat[A4], rlon[A4], rhdg[A4], rnhdg[A4];  // rocket of the AC1 shooter in that unit slot
  double nlat[A4], nlon[A4];
  double hvc[A4], hvs[A4], hvn[A4];
  double rel_focus[A4], near_dn[A4];
  double rew[A2], opp_focus[A2];
  unsigned long long dg[A1], dg0[A1];
  int crem[A4], burst[A4], cmax[A4], mrem[A4], rmax[A4], mwait[A4];
  int rage[A4], rtgt[A4], rid[A4], ota[A4];
  int near_t[A4], rel_sign[A4], tgt[A4], new_wait[A4], inr[A4];
  unsigned char alive[A4], hasm[A4], ralive[A4], want[A4], launched[A4], upd[A4], rocket0[A4], hit_t[A4],
      hit_f[A4], exploded[A4], firing[A4];
  int steps[A1], alive_ag[A1], alive_op[A1], esc_time[A1], next_id[A1], pset[A1], opp_mode[A1], err[A1], escaping[A1];
  unsigned int dc[A1];
  unsigned int killer_pack[A1];
  int by_rocket[A1], done[A1], alive_pre[A1];
  int4 act[A2];
  float obs1[A1 * OBS_ESC_AC1], obs2[A1 * OBS_ESC_AC2];
};

__device__ __forceinline__ HVec sm_hv(const Smem& S, int i) {
  HVec h;
  h.c = S.hvc[i];
  h.s = S.hvs[i];
  h.n = S.hvn[i];
  return h;
}

// nearest live enemy of unit slot us (arena base b, unit u), env_base.py:400-422
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

// opp_ac_values (env_base.py:185-212) of enemy q for unit u (slots b+q, b+u)
__device__ __forceinline__ int sm_enemy_block(const Smem& S, const Geom& g, int b, int u, int q, double dq, int omode,
                                              double f_uq, double f_qu, int shot_m, float* out) {
  double x, y;
  rel_pos(g, S.lat[b + q], S.lon[b + q], x, y);
  int n = 0;
  out[n++] = (float)x;
  out[n++] = (float)y;
  out[n++] = (float)clip(S.spd[b + q] / max_speed(q), 0.0, 1.0);
  out[n++] = (float)hdg_feature(S.hdg[b + q]);
  out[n++] = (float)hdiff_norm(sm_hv(S, b + q), sm_hv(S, b + u));
  if (omode == 0) {
    out[n++] = (float)focus_norm_from_deg(f_qu);
    out[n++] = (float)aspect_from_deg(f_uq);
  } else {
    out[n++] = (float)focus_norm_from_deg(f_uq);
    out[n++] = (float)focus_norm_from_deg(f_qu);
  }
  out[n++] = (float)dq;
  out[n++] = ((shot_m >> q) & 1) ? 1.0f : 0.0f;
  return n;
}

// lowlevel_state (env_hetero.py:65-103) of agent u in arena base b; returns opp_to_attack (0 = None)
__device__ __forceinline__ int sm_observation(const Smem& S, const Geom& g, int b, int u, int omode, int alive_m,
                                              int shot_m, float* out) {
  const int len = obs_len(u, omode);
  const int us = b + u;
  double dn;
  const int o = ((alive_m >> u) & 1) ? sm_nearest(S, g, b, u, alive_m, dn) : -1;
  if (o < 0) {
    for (int k = 0; k < len; ++k) out[k] = 0.0f;
    return 0;
  }
  int n = 0;
  double x, y;
  rel_pos(g, S.lat[us], S.lon[us], x, y);
  out[n++] = (float)x;
  out[n++] = (float)y;
  out[n++] = (float)clip(S.spd[us] / max_speed(u), 0.0, 1.0);
  out[n++] = (float)hdg_feature(S.hdg[us]);
  const HVec hu = sm_hv(S, us);
  const double f_uo = focus_deg(hu, S.lat[us], S.lon[us], S.lat[b + o], S.lon[b + o]);
  const double f_ou = focus_deg(sm_hv(S, b + o), S.lat[b + o], S.lon[b + o], S.lat[us], S.lon[us]);
  if (omode == 0) {
    out[n++] = (float)focus_norm_from_deg(f_uo);
    out[n++] = (float)aspect_from_deg(f_ou);
    out[n++] = (float)hdiff_norm(hu, sm_hv(S, b + o));
    out[n++] = (float)dn;
    out[n++] = (float)clip((double)S.crem[us] / (double)S.cmax[us], 0.0, 1.0);
    if (is_ac1(u)) {
      out[n++] = (float)clip((double)S.mrem[us] / (double)S.rmax[us], 0.0, 1.0);
      out[n++] = S.mwait[us] == 0 ? 1.0f : 0.0f;
      out[n++] = (S.hasm[us] || S.burst[us] > 0) ? 1.0f : 0.0f;
    } else {
      out[n++] = S.burst[us] > 0 ? 1.0f : 0.0f;
    }
    n += sm_enemy_block(S, g, b, u, o, dn, 0, f_uo, f_ou, shot_m, out + n);
  } else {
    out[n++] = (float)clip((double)S.crem[us] / (double)S.cmax[us], 0.0, 1.0);
    if (is_ac1(u)) out[n++] = (float)clip((double)S.mrem[us] / (double)S.rmax[us], 0.0, 1.0);
    out[n++] = (S.burst[us] > 0 || (is_ac1(u) && S.hasm[us])) ? 1.0f : 0.0f;
    n += sm_enemy_block(S, g, b, u, o, dn, 1, f_uo, f_ou, shot_m, out + n);
    const int e0 = u < 2 ? 2 : 0;
    const int o2 = (o == e0) ? e0 + 1 : e0;
    if ((alive_m >> o2) & 1) {
      const double dq = g.inv_diag * dist_raw(S.lat[us], S.lon[us], S.lat[b + o2], S.lon[b + o2]);
      const double f_uq = focus_deg(hu, S.lat[us], S.lon[us], S.lat[b + o2], S.lon[b + o2]);
      const double f_qu = focus_deg(sm_hv(S, b + o2), S.lat[b + o2], S.lon[b + o2], S.lat[us], S.lon[us]);
      n += sm_enemy_block(S, g, b, u, o2, dq, 1, f_uq, f_qu, shot_m, out + n);
    } else {
      for (int z = 0; z < 9; ++z) out[n++] = 0.0f;
    }
  }
  const int fri = u ^ 1;
  if ((alive_m >> fri) & 1) {
    rel_pos(g, S.lat[b + fri], S.lon[b + fri], x, y);
    out[n++] = (float)x;
    out[n++] = (float)y;
    out[n++] = (float)focus_norm_from_deg(focus_deg(hu, S.lat[us], S.lon[us], S.lat[b + fri], S.lon[b + fri]));
    out[n++] = (float)focus_norm_from_deg(
        focus_deg(sm_hv(S, b + fri), S.lat[b + fri], S.lon[b + fri], S.lat[us], S.lon[us]));
    out[n++] = (float)(g.inv_diag * dist_raw(S.lat[us], S.lon[us], S.lat[b + fri], S.lon[b + fri]));
  } else {
    for (int k = 0; k < 5; ++k) out[n++] = 0.0f;
  }
  return o + 1;
}

template <int LEVEL, int MODE>
__device__ __forceinline__ void step_body(Smem& S, const StatePtrs& G, const Params& P, const int32_t* __restrict__ actions,
                                          float* __restrict__ obs1, float* __restrict__ obs2,
                                          float* __restrict__ rew_out, uint8_t* __restrict__ done_out) {
  constexpr int D1 = MODE == 0 ? OBS_AC1 : OBS_ESC_AC1, D2 = MODE == 0 ? OBS_AC2 : OBS_ESC_AC2;
  const int tid = threadIdx.x;
  const int arena0 = blockIdx.x * kArenas;
  const int n_valid = min(kArenas, P.n_arenas - arena0);
  const Geom g = make_geom(P.map_size);
  // unit-mapped view
  const int ul = tid >> 2, u = tid & 3;            // local arena, unit
  const int ub = ul * 4;                           // slot of unit 0 of that arena
  const bool uvalid = ul < n_valid;
  const int ua = arena0 + (uvalid ? ul : n_valid - 1);   // tail threads shadow the last valid arena (no stores)

  // ---------------------------------------------------------------- P0: load (unit-mapped)
  {
    Lane L;
    load_lane(G, ua, u, L);
    S.lat[tid] = L.lat; S.lon[tid] = L.lon; S.hdg[tid] = L.hdg; S.spd[tid] = L.spd; S.nhdg[tid] = L.nhdg; S.nspd[tid] = L.nspd;
    S.rlat[tid] = L.rlat; S.rlon[tid] = L.rlon; S.rhdg[tid] = L.rhdg; S.rnhdg[tid] = L.rnhdg;
    S.crem[tid] = L.crem; S.burst[tid] = L.burst; S.cmax[tid] = L.cmax; S.mrem[tid] = L.mrem; S.rmax[tid] = L.rmax;
    S.mwait[tid] = L.mwait; S.alive[tid] = L.alive; S.hasm[tid] = L.hasm; S.ralive[tid] = L.ralive; S.rage[tid] = L.rage;
    S.rtgt[tid] = L.rtgt; S.rid[tid] = L.rid; S.ota[tid] = L.ota;
    if (u == 0) {
      S.steps[ul] = L.steps + 1;                    // self.steps += 1 (env_hetero.py:114)
      S.alive_ag[ul] = L.alive_ag; S.alive_op[ul] = L.alive_op; S.esc_time[ul] = L.esc_time; S.next_id[ul] = L.next_id;
      S.pset[ul] = L.pset; S.opp_mode[ul] = L.opp_mode; S.err[ul] = L.err; S.escaping[ul] = L.escaping;
      S.dg[ul] = L.dg; S.dg0[ul] = L.dg; S.dc[ul] = L.dc;
    }
    if (tid < A2) {
      const int al = tid >> 1;
      const int a = arena0 + (al < n_valid ? al : n_valid - 1);
      S.act[tid] = reinterpret_cast<const int4*>(actions)[(size_t)a * 2 + (tid & 1)];
    }
  }
  __syncthreads();

  // ---------------------------------------------------------------- P1: pre-tick relations (unit-mapped)
  {
    const int alive_m = S.alive[ub] | (S.alive[ub + 1] << 1) | (S.alive[ub + 2] << 2) | (S.alive[ub + 3] << 3);
    if (u == 0) S.alive_pre[ul] = alive_m;
    double ndn = 0.0;
    int nt = -1;
    if (u >= 2 && S.alive[tid]) nt = sm_nearest(S, g, ub, u, alive_m, ndn);
    int rx = -1, ry = -1;
    if (u < 2) {
      const int t = S.ota[tid];
      if (S.alive[tid] && t != 0 && ((alive_m >> (t - 1)) & 1)) { rx = t - 1; ry = u; }
    } else if (LEVEL == 3 && nt >= 0) {
      rx = u;
      ry = nt;
    }
    double rf = 0.0;
    int rs = 1;
    if (rx >= 0) {
      rf = focus_deg(heading_vec(S.hdg[ub + rx]), S.lat[ub + rx], S.lon[ub + rx], S.lat[ub + ry], S.lon[ub + ry]);
      if (u >= 2) rs = correct_angle_sign(S.lat[ub + rx], S.lon[ub + rx], S.hdg[ub + rx], S.lat[ub + ry], S.lon[ub + ry]);
    }
    S.near_t[tid] = nt;
    S.near_dn[tid] = ndn;
    S.rel_focus[tid] = rf;
    S.rel_sign[tid] = rs;
    S.want[tid] = 0;
    S.tgt[tid] = -1;
    S.new_wait[tid] = 0;
    if (tid < A2) S.rew[tid] = 0.0;
  }
  __syncthreads();

  // ---------------------------------------------------------------- P2: actions
  // warps 0-1: agents' _take_base_action (env_base.py:214-238), one thread per agent
  // warp 2   : scripted opponents (env_hetero.py:118-158), one thread per arena, ids 3 then 4
  if (tid < A2) {
    const int al = tid >> 1, au = tid & 1, us = al * 4 + au;
    S.opp_focus[tid] = focus_norm_from_deg(S.rel_focus[us]);
    if (S.alive[us]) {
      const int4 act = S.act[tid];
      const double h = pymod(S.hdg[us] + (double)((act.x - 6) * 15), 360.0);
      if (h >= 360.0 || h < 0.0) atomicOr(&S.err[al], ERR_HEADING);
      S.nhdg[us] = h;
      S.nspd[us] = 100.0 + ((max_speed(au) - 100.0) / 8.0) * (double)act.y;
      if (act.z != 0 && S.crem[us] > 0) {
        const int bt = is_ac1(au) ? 5 : 3;
        S.burst[us] = S.crem[us] < bt ? S.crem[us] : bt;
        if (MODE == 1 && S.crem[us] < 90) S.rew[tid] -= 0.1;
      }
      const bool want = is_ac1(au) && act.w != 0 && S.ota[us] != 0 && S.mrem[us] > 0 && !S.hasm[us] && S.mwait[us] == 0;
      if (want) {
        const Rng rng{P.seed_lo, P.seed_hi, P.arena_base + (uint32_t)(arena0 + min(al, n_valid - 1))};
        S.want[us] = 1;
        S.tgt[us] = S.ota[us] - 1;
        S.new_wait[us] = randint_from(7, 17, g_random_at(rng, S.dg0[al]));   // drawn iff attempted (env_base.py:228-230)
      }
    }
  } else if (tid < A2 + A1) {
    const int al = tid - A2, b = al * 4;
    const Rng rng{P.seed_lo, P.seed_hi, P.arena_base + (uint32_t)(arena0 + min(al, n_valid - 1))};
    // agent 1's attempt consumes one G draw before the opponents' (same predicate as above)
    const int4 a0 = S.act[al * 2];
    const bool want0 = S.alive[b] && a0.w != 0 && S.ota[b] != 0 && S.mrem[b] > 0 && !S.hasm[b] && S.mwait[b] == 0;
    Lane L;                      // only the arena scalars of the Lane are used by scripted_opponent
    L.steps = S.steps[al];
    L.escaping = S.escaping[al] != 0;
    L.esc_time = S.esc_time[al];
    L.dg = S.dg0[al] + (want0 ? 1 : 0);
#pragma unroll 1
    for (int k = 2; k < 4; ++k) {
      const int us = b + k;
      const bool k_alive = S.alive[us];
      const OppDecision d = scripted_opponent<LEVEL>(L, rng, g, k, k_alive, S.hasm[us], S.mwait[us], S.lat[us], S.lon[us],
                                                     S.hdg[us], S.near_t[us], S.near_dn[us], S.rel_focus[us], S.rel_sign[us]);
      if (k_alive) {
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
        S.want[us] = d.want_missile;
        S.tgt[us] = d.tgt;
      }
    }
    S.escaping[al] = L.escaping;
    S.esc_time[al] = L.esc_time;
    S.dg[al] = L.dg;
  }
  __syncthreads();

  // ---------------------------------------------------------------- P3: launches (unit-mapped, rare)
  S.launched[tid] = 0;
  if (S.want[tid]) {
    if (!S.hasm[tid] && S.mrem[tid] > 0) {
      const int t = S.tgt[tid] < 0 ? 0 : S.tgt[tid];
      if (launch_gate(S.lat[tid], S.lon[tid], S.hdg[tid], S.lat[ub + t], S.lon[ub + t])) {
        S.rlat[tid] = S.lat[tid]; S.rlon[tid] = S.lon[tid]; S.rhdg[tid] = S.hdg[tid]; S.rnhdg[tid] = S.hdg[tid];
        S.ralive[tid] = 1; S.rage[tid] = 0; S.rtgt[tid] = t + 1;
        S.hasm[tid] = 1;
        S.mrem[tid] -= 1;
        S.launched[tid] = 1;
      }
    }
  }
  __syncthreads();
  {
    // rocket ids in shooter id order; missile_wait bookkeeping (env_base.py:228-236, env_hetero.py:123,136,158)
    const int l0 = S.launched[ub], l2 = S.launched[ub + 2];
    if (S.launched[tid]) S.rid[tid] = S.next_id[ul] + (u == 2 ? l0 : 0);
    if (S.want[tid]) {
      if (u < 2) {
        S.mwait[tid] = S.new_wait[tid];
        if (MODE == 1 && S.mrem[tid] < 3) S.rew[ul * 2 + u] -= 0.1;
      } else {
        S.mwait[tid] = LEVEL == 3 ? 10 : 5;
      }
    }
    if (u < 2 && S.alive[tid] && S.mwait[tid] > 0 && !S.hasm[tid]) S.mwait[tid] -= 1;
    __syncthreads();
    if (u == 0) S.next_id[ul] += l0 + l2;
  }

  // ---------------------------------------------------------------- P4: rate limits + moves (unit-mapped)
  {
    const bool upd = S.alive[tid];
    S.upd[tid] = upd;
    S.rocket0[tid] = S.ralive[tid];
    const double max_deg = is_ac1(u) ? 5.0 : 3.5, max_kn = is_ac1(u) ? 35.0 : 28.0;
    double hdg = S.hdg[tid], spd = S.spd[tid];
    if (upd) {
      const double nh = S.nhdg[tid], ns = S.nspd[tid];
      if (hdg != nh) {
        const double dd = signed_heading_diff(hdg, nh);
        hdg = fabs(dd) <= max_deg ? nh : pymod(hdg + (dd >= 0.0 ? max_deg : -max_deg), 360.0);
      }
      if (spd != ns) {
        const double dd = ns - spd;
        spd = fabs(dd) <= max_kn ? ns : spd + (dd >= 0.0 ? max_kn : -max_kn);
      }
      S.hdg[tid] = hdg;
      S.spd[tid] = spd;
    }
    const bool firing = upd && S.burst[tid] > 0;
    S.firing[tid] = firing;
    if (firing) {
      S.burst[tid] -= 1;
      S.crem[tid] = S.crem[tid] > 0 ? S.crem[tid] - 1 : 0;
    }
    double nlat = S.lat[tid], nlon = S.lon[tid];
    if (upd && spd > 0.0) {
      const double2 q = geo::direct(nlat, nlon, hdg, spd * kKnotsToMs * 1.0);
      nlat = q.x;
      nlon = q.y;
    }
    S.nlat[tid] = nlat;
    S.nlon[tid] = nlon;
  }
  __syncthreads();

  // ---------------------------------------------------------------- P5: cannon geometry (unit-mapped shooters)
  {
    int in_range = 0;
    if (S.firing[tid]) {
      const int alive0 = S.alive_pre[ul];
      const double range = is_ac1(u) ? 2.0 : 4.5, half_w = (is_ac1(u) ? 10.0 : 7.0) / 2.0;
#pragma unroll 1
      for (int j = 0; j < 4; ++j) {
        if (j == u || !((alive0 >> j) & 1)) continue;
        if (!(P.friendly_kill || ((u < 2) != (j < 2)))) continue;
        const double jl = j < u ? S.nlat[ub + j] : S.lat[ub + j];   // lower ids have already moved
        const double jo = j < u ? S.nlon[ub + j] : S.lon[ub + j];
        if (unit_in_cannon_range(S.lat[tid], S.lon[tid], S.hdg[tid], jl, jo, range, half_w)) in_range |= 1 << j;
      }
    }
    S.inr[tid] = in_range;
  }
  __syncthreads();

  // ---------------------------------------------------------------- P6: kill resolution + missile noise (arena-mapped)
  if (tid < A1) {
    const int al = tid, b = al * 4;
    const Rng rng{P.seed_lo, P.seed_hi, P.arena_base + (uint32_t)(arena0 + min(al, n_valid - 1))};
    int alive_m = S.alive_pre[al];
    unsigned killer_pack = 0;
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
    // missile heading noise (ac1.py:117-128), shooters in id order
    unsigned long long dg = S.dg[al];
#pragma unroll
    for (int k = 0; k < 4; k += 2) {
      const int us = b + k;
      if (S.upd[us] && S.hasm[us]) {
        if (!S.rocket0[us]) S.hasm[us] = 0;
        else S.rnhdg[us] = clip(__dmul_rn(S.rhdg[us], uniform_from(0.95, 1.05, g_random_at(rng, dg++))), 0.0, 359.0);
      }
    }
    S.dg[al] = dg;
    S.killer_pack[al] = killer_pack;
    S.by_rocket[al] = 0;
    S.alive_pre[al] = alive_m;        // from here on: alive mask after the cannon phase
  }
  __syncthreads();

  // ---------------------------------------------------------------- P7: rocket geometry (unit-mapped rocket owners)
  {
    bool ht = false, hf = false;
    if (S.rocket0[tid]) {
      const int t = S.rtgt[tid] > 0 ? S.rtgt[tid] - 1 : 0;
      ht = within_1km(S.rlat[tid], S.rlon[tid], S.nlat[ub + t], S.nlon[ub + t]);     // targets have already moved
      if (P.friendly_kill && ((S.alive_pre[ul] >> 1) & 1))
        hf = within_1km(S.rlat[tid], S.rlon[tid], S.nlat[ub + 1], S.nlon[ub + 1]);   // "friendly" is always id 2
    }
    S.hit_t[tid] = ht;
    S.hit_f[tid] = hf;
    S.exploded[tid] = 0;
  }
  __syncthreads();

  // ---------------------------------------------------------------- P8: rocket resolution in launch order (arena-mapped)
  if (tid < A1) {
    const int al = tid, b = al * 4;
    const int r0 = S.rocket0[b], r2 = S.rocket0[b + 2];
    if (r0 || r2) {
      int alive_m = S.alive_pre[al];
      unsigned killer_pack = S.killer_pack[al];
      int by_rocket = 0;
      const int first = (r0 && r2) ? (S.rid[b] < S.rid[b + 2] ? 0 : 2) : (r0 ? 0 : 2);
#pragma unroll 1
      for (int n = 0; n < 2; ++n) {
        const int s = n == 0 ? first : 2 - first;
        if (!S.rocket0[b + s]) continue;
        const int tt = S.rtgt[b + s] > 0 ? S.rtgt[b + s] - 1 : 0;
        if (S.hit_t[b + s] && ((alive_m >> tt) & 1)) {
          alive_m &= ~(1 << tt);
          killer_pack |= (unsigned)(s + 1) << (4 * tt);
          by_rocket |= 1 << tt;
          S.exploded[b + s] = 1;
        } else if (S.hit_f[b + s] && ((alive_m >> 1) & 1)) {
          alive_m &= ~2;
          killer_pack |= (unsigned)(s + 1) << 4;
          by_rocket |= 2;
          S.exploded[b + s] = 1;
        }
      }
      S.alive_pre[al] = alive_m;
      S.killer_pack[al] = killer_pack;
      S.by_rocket[al] = by_rocket;
    }
  }
  __syncthreads();

  // ---------------------------------------------------------------- P9: rocket moves, commit positions (unit-mapped)
  {
    if (S.rocket0[tid]) {
      if (S.exploded[tid] || S.rage[tid] > 10) {
        S.ralive[tid] = 0;
      } else {
        double h = S.rhdg[tid];
        const double nh = S.rnhdg[tid];
        if (h != nh) {
          const double dd = signed_heading_diff(h, nh);
          h = fabs(dd) <= 10.0 ? nh : h + (dd >= 0.0 ? 10.0 : -10.0);
        }
        const double2 q = geo::direct(S.rlat[tid], S.rlon[tid], h, rocket_speed(S.rage[tid]) * kKnotsToMs * 1.0);
        S.rhdg[tid] = h;
        S.rlat[tid] = q.x;
        S.rlon[tid] = q.y;
        S.rage[tid] += 1;
      }
    }
    if (S.upd[tid]) {
      S.lat[tid] = S.nlat[tid];
      S.lon[tid] = S.nlon[tid];
    }
    const bool al = (S.alive_pre[ul] >> u) & 1;
    S.alive[tid] = al;
    S.firing[tid] = al && !in_boundary(g, S.lat[tid], S.lon[tid]);     // reuse: out-of-bounds flag
  }
  __syncthreads();

  // ---------------------------------------------------------------- P10: rewards, termination (arena-mapped)
  if (tid < A1) {
    const int al = tid, b = al * 4;
    const double s = P.rew_scale;
    const int oob_m = S.firing[b] | (S.firing[b + 1] << 1) | (S.firing[b + 2] << 2) | (S.firing[b + 3] << 3);
    int alive_m = S.alive_pre[al] & ~oob_m;
    const int present_m = (S.upd[b] ? 1 : 0) | (S.upd[b + 1] ? 2 : 0);   // alive at step start (reward-dict membership)
    double rews0 = 0.0, rews1 = 0.0;
    int destroyed_m = oob_m & 3;
    if (oob_m & 1) rews0 += -5.0 * s;
    if (oob_m & 2) rews1 += -5.0 * s;
    int alive_ag = S.alive_ag[al] - __popc(oob_m & 3), alive_op = S.alive_op[al] - __popc(oob_m & 12);
    const unsigned killer_pack = S.killer_pack[al];
    const int by_rocket_m = S.by_rocket[al];
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
                r = (1.0 + 0.5 * ((double)S.mrem[b] / (double)S.rmax[b])) * s;
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
    if (MODE == 1 && P.esc_dist_rew) {
#pragma unroll
      for (int i = 0; i < 2; ++i) {
        if (!((alive_m >> i) & 1)) continue;
        const double d2 = dist_raw(S.lat[b + i], S.lon[b + i], S.lat[b + 2], S.lon[b + 2]);
        const double d3 = dist_raw(S.lat[b + i], S.lon[b + i], S.lat[b + 3], S.lon[b + 3]);
        const bool has2 = (alive_m >> 2) & 1, has3 = (alive_m >> 3) & 1;
        const bool swap = has2 && has3 && (g.inv_diag * d3) < (g.inv_diag * d2);
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
    S.alive_pre[al] = alive_m;
    const bool done = alive_ag <= 0 || alive_op <= 0 || S.steps[al] >= P.horizon;
    S.done[al] = done ? 1 : 0;
    if (al < n_valid) {
      const int a = arena0 + al;
      if (rew_out) reinterpret_cast<float2*>(rew_out)[a] = make_float2((float)r0, (float)r1);
      if (done_out) done_out[a] = done ? 1 : 0;
    }
  }
  __syncthreads();

  // ---------------------------------------------------------------- P11: alive flags / auto-reset (unit-mapped)
  {
    S.alive[tid] = (S.alive_pre[ul] >> u) & 1;
    if (S.done[ul] && P.autoreset) {
      const Rng rng{P.seed_lo, P.seed_hi, P.arena_base + (uint32_t)ua};
      Lane L;
      L.dg = S.dg[ul];
      L.dc = S.dc[ul];
      L.err = S.err[ul];
      reset_lane(L, rng, P, u);
      S.lat[tid] = L.lat; S.lon[tid] = L.lon; S.hdg[tid] = L.hdg; S.spd[tid] = L.spd; S.nhdg[tid] = L.nhdg; S.nspd[tid] = L.nspd;
      S.rlat[tid] = 0.0; S.rlon[tid] = 0.0; S.rhdg[tid] = 0.0; S.rnhdg[tid] = 0.0;
      S.crem[tid] = L.crem; S.burst[tid] = 0; S.cmax[tid] = L.cmax; S.mrem[tid] = L.mrem; S.rmax[tid] = L.rmax; S.mwait[tid] = 0;
      S.alive[tid] = 1; S.hasm[tid] = 0; S.ralive[tid] = 0; S.rage[tid] = 0; S.rtgt[tid] = 0; S.rid[tid] = 0; S.ota[tid] = 0;
      __syncwarp(0xFu << (tid & 28));    // the quad's four lanes read dg above before lane 0 overwrites it
      if (u == 0) {
        S.steps[ul] = 0; S.alive_ag[ul] = 2; S.alive_op[ul] = 2; S.esc_time[ul] = 0; S.escaping[ul] = 0; S.next_id[ul] = 5;
        S.pset[ul] = L.pset; S.opp_mode[ul] = L.opp_mode; S.dg[ul] = L.dg;
      }
    }
  }
  __syncthreads();

  // ---------------------------------------------------------------- P12: heading vectors (unit-mapped)
  {
    const HVec h = heading_vec(S.hdg[tid]);
    S.hvc[tid] = h.c;
    S.hvs[tid] = h.s;
    S.hvn[tid] = h.n;
    S.firing[tid] = S.burst[tid] > 0 || (is_ac1(u) && S.hasm[tid]);   // reuse: shot flag (env_base.py:208-211)
  }
  __syncthreads();

  // ---------------------------------------------------------------- P13: observations (agent-mapped)
  if (tid < A2) {
    const int al = tid >> 1, au = tid & 1, b = al * 4;
    const int alive_m = S.alive[b] | (S.alive[b + 1] << 1) | (S.alive[b + 2] << 2) | (S.alive[b + 3] << 3);
    const int shot_m = S.firing[b] | (S.firing[b + 1] << 1) | (S.firing[b + 2] << 2) | (S.firing[b + 3] << 3);
    float* row = au == 0 ? S.obs1 + al * D1 : S.obs2 + al * D2;
    S.ota[b + au] = sm_observation(S, g, b, au, MODE, alive_m, shot_m, row);
  } else {
    const int k = tid - A2;             // opponents' opp_to_attack stays None at levels 1-3
    const int al = k >> 1, ou = 2 + (k & 1);
    S.ota[al * 4 + ou] = 0;
  }
  __syncthreads();

  // ---------------------------------------------------------------- P14: store (unit-mapped) + observation rows
  {
    Lane L;
    L.lat = S.lat[tid]; L.lon = S.lon[tid]; L.hdg = S.hdg[tid]; L.spd = S.spd[tid]; L.nhdg = S.nhdg[tid]; L.nspd = S.nspd[tid];
    L.rlat = S.rlat[tid]; L.rlon = S.rlon[tid]; L.rhdg = S.rhdg[tid]; L.rnhdg = S.rnhdg[tid];
    L.crem = S.crem[tid]; L.burst = S.burst[tid]; L.cmax = S.cmax[tid]; L.mrem = S.mrem[tid]; L.rmax = S.rmax[tid];
    L.mwait = S.mwait[tid]; L.alive = S.alive[tid]; L.hasm = S.hasm[tid]; L.ralive = S.ralive[tid]; L.rage = S.rage[tid];
    L.rtgt = S.rtgt[tid]; L.rid = S.rid[tid]; L.ota = S.ota[tid];
    L.steps = S.steps[ul]; L.alive_ag = S.alive_ag[ul]; L.alive_op = S.alive_op[ul]; L.esc_time = S.esc_time[ul];
    L.next_id = S.next_id[ul]; L.pset = S.pset[ul]; L.opp_mode = S.opp_mode[ul]; L.err = S.err[ul];
    L.escaping = S.escaping[ul] != 0; L.dg = S.dg[ul]; L.dc = S.dc[ul];
    store_lane(G, ua, u, L, uvalid);
    const int n4_1 = (n_valid * D1) >> 2, n4_2 = (n_valid * D2) >> 2;
    if (obs1) {
      float4* d4 = reinterpret_cast<float4*>(obs1 + (size_t)arena0 * D1);
      const float4* s4 = reinterpret_cast<const float4*>(S.obs1);
      for (int k = tid; k < n4_1; k += kThreads) d4[k] = s4[k];
      for (int k = (n4_1 << 2) + tid; k < n_valid * D1; k += kThreads) obs1[(size_t)arena0 * D1 + k] = S.obs1[k];
    }
    if (obs2) {
      float4* d4 = reinterpret_cast<float4*>(obs2 + (size_t)arena0 * D2);
      const float4* s4 = reinterpret_cast<const float4*>(S.obs2);
      for (int k = tid; k < n4_2; k += kThreads) d4[k] = s4[k];
      for (int k = (n4_2 << 2) + tid; k < n_valid * D2; k += kThreads) obs2[(size_t)arena0 * D2 + k] = S.obs2[k];
    }
  }
}

}  // namespace cta
}  // namespace hh
