// hh_quad.cuh -- the 2-vs-2 low-level arena advanced cooperatively by a QUAD of lanes.
//
// Mapping: 4 consecutive lanes of a warp own one arena; lane u = aircraft id u+1 (ids 1,2 =
// agents AC1/AC2, ids 3,4 = opponents AC1/AC2, env_base.py:556-560) and, for the AC1 lanes
// (u = 0, 2), the single rocket that shooter can have in flight (ac1.py:73).  8 arenas per warp.
//
// What runs in parallel in the lanes: heading/speed rate limits, the WGS84 direct solve of every
// unit's move, the shooter->target cannon geometry, rocket proximity geometry, the flat-plane
// angle features of opponents/agents, reset sampling and the observation vectors.
// What the reference serialises (CmanoSimulator.do_tick's id-ordered updates against a snapshot,
// "dead units still act in the tick they die", RNG consumption only for live in-range targets,
// rockets after aircraft in launch order) is resolved REDUNDANTLY by all four
// lanes from warp ballots / shuffles of the per-lane geometry bits, so the arena scalars (alive
// counters, RNG draw counters, escape flag ...) stay identical in the four lanes without a
// broadcast and no lane idles waiting for a "leader".
//
// All __shfl_sync/__ballot_sync sit at the top level of the step (never under lane-divergent
// control flow); lanes of arenas beyond N recompute the last valid arena and skip the stores.
#pragma once
#include "hh_core.cuh"

namespace hh {

constexpr unsigned kFull = 0xffffffffu;

struct StatePtrs {
  double *lat, *lon, *hdg, *spd, *nhdg, *nspd;  // [N*4]  aircraft kinematics
  uint2* acint;                                  // [N*4]  packed aircraft integers
  double *rlat, *rlon, *rhdg, *rnhdg;            // [N*2]  rocket kinematics (slot = shooter u/2)
  uint32_t* rint;                                // [N*2]  packed rocket integers
  uint4* meta;                                   // [N]    arena scalars + C-stream counter
  unsigned long long* draws_g;                   // [N]    G-stream draw counter
};

struct Lane {
  // own aircraft
  double lat, lon, hdg, spd, nhdg, nspd;
  int crem, burst, cmax, mrem, rmax, mwait;
  bool alive, hasm;
  // own rocket (AC1 lanes only)
  double rlat, rlon, rhdg, rnhdg;
  bool ralive;
  int rage, rtgt, rid;
  int ota;  // own opp_to_attack: 0 = None, else id 1..4
  // arena scalars, replicated in the 4 lanes
  int steps, alive_ag, alive_op, esc_time, next_id, pset, opp_mode, err;
  bool escaping;
  unsigned long long dg;
  unsigned int dc;
};

__device__ __forceinline__ int quad_ballot(bool p) {
  unsigned b = __ballot_sync(kFull, p);
  return (int)((b >> (threadIdx.x & 28 & 31)) & 0xFu);
}
template <typename T>
__device__ __forceinline__ T qshfl(T v, int src) {
  return __shfl_sync(kFull, v, src, 4);
}
__device__ __forceinline__ double pick4d(double a0, double a1, double a2, double a3, int i) {
  return i == 0 ? a0 : (i == 1 ? a1 : (i == 2 ? a2 : a3));
}

// ------------------------------------------------------------------------------------- load / store
__device__ __forceinline__ void load_lane(const StatePtrs& S, int a, int u, Lane& L) {
  const size_t i = (size_t)a * 4 + u;
  L.lat = S.lat[i];
  L.lon = S.lon[i];
  L.hdg = S.hdg[i];
  L.spd = S.spd[i];
  L.nhdg = S.nhdg[i];
  L.nspd = S.nspd[i];
  const uint2 w = S.acint[i];
  L.crem = w.x & 0xFFFF;
  L.burst = (w.x >> 16) & 0xFF;
  L.mrem = (w.x >> 24) & 0xFF;
  L.cmax = w.y & 0xFFFF;
  L.mwait = (w.y >> 16) & 0xFF;
  L.rmax = (w.y >> 24) & 0xF;
  L.alive = (w.y >> 28) & 1;
  L.hasm = (w.y >> 29) & 1;
  L.rlat = L.rlon = L.rhdg = L.rnhdg = 0.0;
  L.ralive = false;
  L.rage = L.rtgt = L.rid = 0;
  if ((u & 1) == 0) {
    const size_t r = (size_t)a * 2 + (u >> 1);
    L.rlat = S.rlat[r];
    L.rlon = S.rlon[r];
    L.rhdg = S.rhdg[r];
    L.rnhdg = S.rnhdg[r];
    const uint32_t rw = S.rint[r];
    L.ralive = rw & 1;
    L.rage = (rw >> 1) & 0xF;
    L.rtgt = (rw >> 5) & 0x7;
    L.rid = (rw >> 8) & 0xFFFF;
  }
  const uint4 m = S.meta[a];
  L.steps = m.x & 0xFFFF;
  L.alive_ag = (m.x >> 16) & 0xF;
  L.alive_op = (m.x >> 20) & 0xF;
  L.escaping = (m.x >> 24) & 1;
  L.pset = (m.x >> 25) & 0x7;
  L.opp_mode = (m.x >> 28) & 1;
  L.esc_time = m.y & 0xFF;
  L.next_id = (m.y >> 8) & 0xFF;
  const int o = (m.y >> (24 + 2 * u)) & 0x3;  // agents store target id - 2, opponents the id itself
  L.ota = o == 0 ? 0 : (u < 2 ? o + 2 : o);
  L.dc = m.z;
  L.err = (int)m.w;
  L.dg = S.draws_g[a];
}

__device__ __forceinline__ void store_lane(const StatePtrs& S, int a, int u, const Lane& L, bool valid) {
  // arena-level words need every lane's opp_to_attack / error bits: gather before predication
  const int o = L.ota == 0 ? 0 : (u < 2 ? L.ota - 2 : L.ota);
  int packed = o << (2 * u);
  packed |= __shfl_xor_sync(kFull, packed, 1, 4);
  packed |= __shfl_xor_sync(kFull, packed, 2, 4);
  int err = L.err;
  err |= __shfl_xor_sync(kFull, err, 1, 4);
  err |= __shfl_xor_sync(kFull, err, 2, 4);
  if (!valid) return;
  const size_t i = (size_t)a * 4 + u;
  S.lat[i] = L.lat;
  S.lon[i] = L.lon;
  S.hdg[i] = L.hdg;
  S.spd[i] = L.spd;
  S.nhdg[i] = L.nhdg;
  S.nspd[i] = L.nspd;
  uint2 w;
  w.x = (uint32_t)L.crem | ((uint32_t)L.burst << 16) | ((uint32_t)L.mrem << 24);
  w.y = (uint32_t)L.cmax | ((uint32_t)L.mwait << 16) | ((uint32_t)L.rmax << 24) | ((uint32_t)L.alive << 28) |
        ((uint32_t)L.hasm << 29);
  S.acint[i] = w;
  if ((u & 1) == 0) {
    const size_t r = (size_t)a * 2 + (u >> 1);
    S.rlat[r] = L.rlat;
    S.rlon[r] = L.rlon;
    S.rhdg[r] = L.rhdg;
    S.rnhdg[r] = L.rnhdg;
    S.rint[r] = (uint32_t)L.ralive | ((uint32_t)L.rage << 1) | ((uint32_t)L.rtgt << 5) | ((uint32_t)L.rid << 8);
  }
  if (u == 0) {
    uint4 m;
    m.x = (uint32_t)L.steps | ((uint32_t)L.alive_ag << 16) | ((uint32_t)L.alive_op << 20) |
          ((uint32_t)L.escaping << 24) | ((uint32_t)L.pset << 25) | ((uint32_t)L.opp_mode << 28);
    m.y = (uint32_t)L.esc_time | ((uint32_t)L.next_id << 8) | ((uint32_t)packed << 24);
    m.z = L.dc;
    m.w = (uint32_t)err;
    S.meta[a] = m;
    S.draws_g[a] = L.dg;
  }
}

// ------------------------------------------------------------------------------------- small unit ops
__device__ __forceinline__ void set_heading(Lane& L, double h) {
  if (h >= 360.0 || h < 0.0) L.err |= ERR_HEADING;  // the reference raises (ac1.py:58-61)
  L.nhdg = h;
}
__device__ __forceinline__ void set_speed(Lane& L, int u, double s) {
  if (s > max_speed(u) || s < 0.0) L.err |= ERR_SPEED;  // ac1.py:63-67
  L.nspd = s;
}
// ac1.py:69-70 / ac2.py:65-66
__device__ __forceinline__ void fire_cannon(Lane& L, int u) {
  const int bt = is_ac1(u) ? 5 : 3;
  L.burst = L.crem < bt ? L.crem : bt;
}

// nearest live enemy of unit u by normalised flat distance (env_base.py:400-422); ties -> lower id
__device__ __forceinline__ int nearest_enemy(const Geom& g, int u, double lat, double lon, const double (&lat4)[4],
                                             const double (&lon4)[4], int alive_m, double& dn) {
  const int e0 = u < 2 ? 2 : 0, e1 = e0 + 1;
  const double la0 = u < 2 ? lat4[2] : lat4[0], lo0 = u < 2 ? lon4[2] : lon4[0];
  const double la1 = u < 2 ? lat4[3] : lat4[1], lo1 = u < 2 ? lon4[3] : lon4[1];
  const double d0 = g.inv_diag * dist_raw(lat, lon, la0, lo0);
  const double d1 = g.inv_diag * dist_raw(lat, lon, la1, lo1);
  int best = -1;
  dn = 0.0;
  if ((alive_m >> e0) & 1) { best = e0; dn = d0; }
  if (((alive_m >> e1) & 1) && (best < 0 || d1 < d0)) { best = e1; dn = d1; }
  return best;
}

// ac1.py:72-79 + Rocket.__init__ (rocket_unit.py:23-30) for the calling lane's own aircraft.
// Returns true when a rocket was put into the sim (its id is assigned afterwards, in id order).
__device__ __forceinline__ bool try_launch(Lane& L, bool want, double tlat, double tlon, int tgt_index) {
  bool launched = false;
  if (want && !L.hasm && L.mrem > 0) {
    if (launch_gate(L.lat, L.lon, L.hdg, tlat, tlon)) {
      L.rlat = L.lat;
      L.rlon = L.lon;
      L.rhdg = L.hdg;
      L.rnhdg = L.hdg;
      L.ralive = true;
      L.rage = 0;
      L.rtgt = tgt_index + 1;
      L.hasm = true;
      L.mrem -= 1;
      launched = true;
    }
  }
  return launched;
}

// ------------------------------------------------------------------------------------- scripted opponents
// One scripted opponent's decision (env_hetero.py:118-158, 227-271), evaluated identically by
// all four lanes for opponent k (so that the shared escape flag / timer / G counter advance in
// every lane); only lane k keeps the per-unit outputs.
struct OppDecision {
  double heading, speed;
  bool set_hs, fire, want_missile;
  int tgt;  // index of the agent to shoot at
};

// `next(i)` returns G-stream draw number i of the arena (L.dg counts them): either computed on the spot
// (scripted_opponent below) or read from a table drawn ahead by other threads (hh_v4.cuh).
template <int LEVEL, class NextDraw>
__device__ __forceinline__ OppDecision scripted_opponent_g(Lane& L, NextDraw next, const Geom& g, int k, bool k_alive,
                                                           bool k_hasm, int k_mwait, double k_lat, double k_lon,
                                                           double k_hdg, int k_near, double k_dn, double k_focus,
                                                           int k_sign) {
  OppDecision d;
  d.heading = k_hdg;
  d.speed = 0.0;
  d.set_hs = false;
  d.fire = false;
  d.want_missile = false;
  d.tgt = k_near;
  if (!k_alive) return d;
  if (LEVEL == 3) {  // __opp_level3
    if (L.steps % 60 == 0 && !L.escaping) {
      L.escaping = randint_from(0, 1, next(L.dg++)) != 0;
      if (L.escaping) L.esc_time = (int)uniform_from(20.0, 30.0, next(L.dg++));
    }
    d.set_hs = true;
    bool fire_m = false;
    if (L.escaping) {  // _escaping_opp
      double y, x;
      rel_pos(g, k_lat, k_lon, y, x);
      const double lo = y < 0.5 ? (x < 0.5 ? 30.0 : 300.0) : (x < 0.5 ? 120.0 : 210.0);
      d.heading = (double)(int)uniform_from(lo, lo + 30.0, next(L.dg++));
      d.speed = (double)(int)uniform_from(300.0, 600.0, next(L.dg++));
      d.fire = randint_from(0, 1, next(L.dg++)) != 0;
      L.esc_time -= 1;
      if (L.esc_time <= 0) L.escaping = false;
      d.tgt = -1;
    } else {  // _hardcoded_opp
      d.speed = (double)(int)uniform_from(100.0, 400.0, next(L.dg++));
      if (k_near >= 0) {
        const double r = uniform_from(0.7, 1.3, next(L.dg++));
        if (k_dn > 0.008 && k_focus > 4.0)
          d.heading = pymod(__dadd_rn(k_hdg, __dmul_rn(__dmul_rn(r, (double)k_sign), k_focus)), 360.0);
        if (k_dn > 0.05)
          d.speed = k_focus < 30.0 ? (double)(int)uniform_from(500.0, 800.0, next(L.dg++))
                                   : (double)(int)uniform_from(100.0, 500.0, next(L.dg++));
        d.fire = k_dn < 0.03 && k_focus < 10.0;
        fire_m = k_dn < 0.09 && k_focus < 5.0;
      }
      if (!is_ac1(k)) d.speed = clip(d.speed, 0.0, 600.0);
    }
    d.want_missile = fire_m && d.tgt >= 0 && !k_hasm && k_mwait == 0 && is_ac1(k);
  } else {
    if (LEVEL == 2) {  // __opp_level2: fire_cannon every step, occasional +-90 deg turn
      d.fire = true;
      bool turn = L.steps <= 5;
      if (!turn) turn = (L.steps % randint_from(35, 45, next(L.dg++))) <= 5;
      if (turn) {
        const int r = randint_from(0, 1, next(L.dg++));
        d.set_hs = true;
        d.heading = pymod(k_hdg + (r ? -90.0 : 90.0), 360.0);
        d.speed = (double)(100 + randint_from(0, 4, next(L.dg++)) * 75);
      }
    }
    // __opp_level1 missile rule (also the tail of __opp_level2); short-circuit order of the reference
    bool w = !k_hasm && (L.steps % 40) < 3;
    if (w) w = randint_from(0, 1, next(L.dg++)) != 0;
    d.want_missile = w && k_mwait == 0 && is_ac1(k) && k_near >= 0;
  }
  return d;
}

template <int LEVEL>
__device__ __forceinline__ OppDecision scripted_opponent(Lane& L, const Rng& rng, const Geom& g, int k, bool k_alive,
                                                         bool k_hasm, int k_mwait, double k_lat, double k_lon,
                                                         double k_hdg, int k_near, double k_dn, double k_focus,
                                                         int k_sign) {
  return scripted_opponent_g<LEVEL>(
      L, [&rng](unsigned long long i) { return g_random_at(rng, i); }, g, k, k_alive, k_hasm, k_mwait, k_lat, k_lon, k_hdg,
      k_near, k_dn, k_focus, k_sign);
}

// ------------------------------------------------------------------------------------- observations
struct World {  // post-tick view of all four aircraft, identical in the four lanes
  double lat[4], lon[4], hdg[4], spd[4];
  HVec hv[4];
  int alive_m, shot_m;  // bit u: aircraft u alive / has burst > 0 or (AC1) a missile in flight
};

__device__ __forceinline__ World gather_world(const Lane& L, int u) {
  World W;
  const HVec mine = heading_vec(L.hdg);
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    W.lat[j] = qshfl(L.lat, j);
    W.lon[j] = qshfl(L.lon, j);
    W.hdg[j] = qshfl(L.hdg, j);
    W.spd[j] = qshfl(L.spd, j);
    W.hv[j].c = qshfl(mine.c, j);
    W.hv[j].s = qshfl(mine.s, j);
    W.hv[j].n = qshfl(mine.n, j);
  }
  W.alive_m = quad_ballot(L.alive);
  W.shot_m = quad_ballot(L.burst > 0 || (is_ac1(u) && L.hasm));
  return W;
}

#define HH_W(field, i) pick4d(W.field[0], W.field[1], W.field[2], W.field[3], (i))
__device__ __forceinline__ HVec pick_hv(const World& W, int i) {
  return i == 0 ? W.hv[0] : (i == 1 ? W.hv[1] : (i == 2 ? W.hv[2] : W.hv[3]));
}

// friendly_ac_values (env_base.py:166-183): 5 values
__device__ __forceinline__ void friend_block(const World& W, const Geom& g, int self_u, float* out) {
  const int fri = self_u ^ 1;  // fri_ac_id: 1<->2, 3<->4 (env_hetero.py:71-75)
  if ((W.alive_m >> fri) & 1) {
    double x, y;
    const double lat_s = HH_W(lat, self_u), lon_s = HH_W(lon, self_u);
    const double lat_f = HH_W(lat, fri), lon_f = HH_W(lon, fri);
    rel_pos(g, lat_f, lon_f, x, y);
    out[0] = (float)x;
    out[1] = (float)y;
    out[2] = (float)focus_norm_from_deg(focus_deg(pick_hv(W, self_u), lat_s, lon_s, lat_f, lon_f));
    out[3] = (float)focus_norm_from_deg(focus_deg(pick_hv(W, fri), lat_f, lon_f, lat_s, lon_s));
    out[4] = (float)(g.inv_diag * dist_raw(lat_s, lon_s, lat_f, lon_f));
  } else {
#pragma unroll
    for (int k = 0; k < 5; ++k) out[k] = 0.0f;
  }
}

__device__ __forceinline__ int obs_len(int u, int omode) {
  return omode == 0 ? (is_ac1(u) ? OBS_AC1 : OBS_AC2) : (is_ac1(u) ? OBS_ESC_AC1 : OBS_ESC_AC2);
}

// 9 values of opp_ac_values (env_base.py:185-212) for enemy q as seen by unit u
__device__ __forceinline__ int enemy_block(const World& W, const Geom& g, int u, int q, double dq, int omode,
                                           double f_uq, double f_qu, float* out) {
  double x, y;
  const HVec hq = pick_hv(W, q), hu = pick_hv(W, u);
  rel_pos(g, HH_W(lat, q), HH_W(lon, q), x, y);
  int n = 0;
  out[n++] = (float)x;
  out[n++] = (float)y;
  out[n++] = (float)clip(HH_W(spd, q) / max_speed(q), 0.0, 1.0);
  out[n++] = (float)hdg_feature(HH_W(hdg, q));
  out[n++] = (float)hdiff_norm(hq, hu);
  if (omode == 0) {
    out[n++] = (float)focus_norm_from_deg(f_qu);
    out[n++] = (float)aspect_from_deg(f_uq);
  } else {
    out[n++] = (float)focus_norm_from_deg(f_uq);
    out[n++] = (float)focus_norm_from_deg(f_qu);
  }
  out[n++] = (float)dq;
  out[n++] = ((W.shot_m >> q) & 1) ? 1.0f : 0.0f;
  return n;
}

// lowlevel_state (env_hetero.py:65-103) for the calling lane's unit u in observation mode omode
// (0 fight: fight_state_values env_base.py:111-135; 1 escape: esc_state_values :137-164).
// Writes obs_len floats to out and returns the new opp_to_attack (0 = None).
__device__ __forceinline__ int unit_observation(const Lane& L, const World& W, const Geom& g, int u, int omode,
                                                float* out) {
  const int len = obs_len(u, omode);
  double dn;
  const int o = L.alive ? nearest_enemy(g, u, L.lat, L.lon, W.lat, W.lon, W.alive_m, dn) : -1;
  if (o < 0) {
    for (int k = 0; k < len; ++k) out[k] = 0.0f;
    return 0;
  }
  int n = 0;
  double x, y;
  rel_pos(g, L.lat, L.lon, x, y);
  out[n++] = (float)x;
  out[n++] = (float)y;
  out[n++] = (float)clip(L.spd / max_speed(u), 0.0, 1.0);
  out[n++] = (float)hdg_feature(L.hdg);
  const HVec hu = pick_hv(W, u);
  const double lat_o = HH_W(lat, o), lon_o = HH_W(lon, o);
  const double f_uo = focus_deg(hu, L.lat, L.lon, lat_o, lon_o);
  const double f_ou = focus_deg(pick_hv(W, o), lat_o, lon_o, L.lat, L.lon);
  if (omode == 0) {
    out[n++] = (float)focus_norm_from_deg(f_uo);
    out[n++] = (float)aspect_from_deg(f_ou);
    out[n++] = (float)hdiff_norm(hu, pick_hv(W, o));
    out[n++] = (float)dn;
    out[n++] = (float)clip((double)L.crem / (double)L.cmax, 0.0, 1.0);
    if (is_ac1(u)) {
      out[n++] = (float)clip((double)L.mrem / (double)L.rmax, 0.0, 1.0);
      out[n++] = L.mwait == 0 ? 1.0f : 0.0f;
      out[n++] = (L.hasm || L.burst > 0) ? 1.0f : 0.0f;
    } else {
      out[n++] = L.burst > 0 ? 1.0f : 0.0f;
    }
    n += enemy_block(W, g, u, o, dn, 0, f_uo, f_ou, out + n);
  } else {
    out[n++] = (float)clip((double)L.crem / (double)L.cmax, 0.0, 1.0);
    if (is_ac1(u)) out[n++] = (float)clip((double)L.mrem / (double)L.rmax, 0.0, 1.0);
    out[n++] = (L.burst > 0 || (is_ac1(u) && L.hasm)) ? 1.0f : 0.0f;
    n += enemy_block(W, g, u, o, dn, 1, f_uo, f_ou, out + n);
    const int e0 = u < 2 ? 2 : 0;
    const int o2 = (o == e0) ? e0 + 1 : e0;  // the other enemy, second in the sorted list if alive
    if ((W.alive_m >> o2) & 1) {
      const double lat_q = HH_W(lat, o2), lon_q = HH_W(lon, o2);
      const double dq = g.inv_diag * dist_raw(L.lat, L.lon, lat_q, lon_q);
      const double f_uq = focus_deg(hu, L.lat, L.lon, lat_q, lon_q);
      const double f_qu = focus_deg(pick_hv(W, o2), lat_q, lon_q, L.lat, L.lon);
      n += enemy_block(W, g, u, o2, dq, 1, f_uq, f_qu, out + n);
    } else {
      for (int z = 0; z < 9; ++z) out[n++] = 0.0f;
    }
  }
  friend_block(W, g, u, out + n);
  return o + 1;
}

// ------------------------------------------------------------------------------------- reset
// HHMARLBaseEnv.reset + _reset_scenario + _sample_state (env_base.py:62-77, 489-585) and
// LowLevelEnv.reset (env_hetero.py:53-60).  Each lane samples its own aircraft; the G-stream
// draw order of the reference is r, then per unit (x, y[, heading]) in id order, then (level 5,
// fight) k -- so every lane knows its draw indices without communication.
// `next(i)` returns G-stream draw number i (absolute index): computed on the spot (reset_lane below) or read from a
// table that several threads filled in parallel (hh_v4.cuh).
template <class NextDraw>
__device__ __forceinline__ void reset_lane_g(Lane& L, NextDraw next, const Params& P, int u) {
  const int level = P.level;
  const unsigned long long base = L.dg;
  const int r = randint_from(1, 2, next(base));
  const int opp_draws = level == 1 ? 2 : 3;
  const int off = u == 0 ? 0 : (u == 1 ? 3 : (u == 2 ? 6 : 6 + opp_draws));
  const unsigned long long d = base + 1 + off;
  const int group = u >> 1, i = u & 1;
  double xw0, xw1, xe0, xe1, y0, y1;
  if (level == 1) { xw0 = 7.12; xw1 = 7.14; xe0 = 7.16; xe1 = 7.17; y0 = 5.1; y1 = 5.11; }
  else if (level == 2) { xw0 = 7.08; xw1 = 7.13; xe0 = 7.18; xe1 = 7.23; y0 = 5.08; y1 = 5.13; }
  else { xw0 = 7.07; xw1 = 7.12; xe0 = 7.18; xe1 = 7.23; y0 = 5.09; y1 = 5.12; }
  const bool west = (group == 0) == (r == 1);  // agents start west when r == 1, opponents east
  const double di = (double)i * 0.1;
  const double x = west ? uniform_from(xw0, xw1, next(d)) : uniform_from(xe0, xe1, next(d));
  const double y = uniform_from(__dadd_rn(y0, di), __dadd_rn(y1, di), next(d + 1));
  int a = 0;
  if (group == 0) {
    const double rr = next(d + 2);
    if (level == 1) a = r == 1 ? randint_from(30, 150, rr) : randint_from(200, 330, rr);
    else if (level == 2) a = r == 1 ? randint_from(0, 180, rr) : randint_from(180, 359, rr);
    else a = r == 1 ? randint_from(0, 270, rr) : randint_from(90, 359, rr);
  } else if (level >= 2) {
    a = randint_from(0, 359, next(d + 2));
  }
  L.lat = y;
  L.lon = x;
  L.hdg = (double)a;
  L.nhdg = (double)a;
  const double sp = (level <= 2 && group == 1) ? 0.0 : 100.0;
  L.spd = sp;
  L.nspd = sp;
  int cm = 200, mr = is_ac1(u) ? 5 : 0;  // ac1.py:30,33 / ac2.py:29,46
  if (level <= 4 && group == 1) { cm = 400; if (is_ac1(u)) mr = 8; }       // env_base.py:567-570
  else if (level == 5) { cm = 300; if (is_ac1(u)) mr = 6; }                // env_base.py:571-574
  L.crem = cm; L.cmax = cm; L.burst = 0;
  L.mrem = mr; L.rmax = mr; L.mwait = 0;
  L.alive = true; L.hasm = false;
  L.rlat = L.rlon = L.rhdg = L.rnhdg = 0.0;
  L.ralive = false; L.rage = 0; L.rtgt = 0; L.rid = 0;
  L.ota = 0;
  L.steps = 0; L.alive_ag = 2; L.alive_op = 2;
  L.escaping = false; L.esc_time = 0; L.next_id = 5;
  L.pset = 0; L.opp_mode = 0;
  unsigned long long used = 1 + 6 + 2 * opp_draws;
  if (level == 5 && P.agent_mode == 0) {
    const int k = randint_from(3, 5, next(base + used));
    used += 1;
    L.pset = k;
    L.opp_mode = k == 5 ? 1 : 0;
  }
  L.dg = base + used;
}

__device__ __forceinline__ void reset_lane(Lane& L, const Rng& rng, const Params& P, int u) {
  reset_lane_g(L, [&rng](unsigned long long i) { return g_random_at(rng, i); }, P, u);
}

}  // namespace hh
