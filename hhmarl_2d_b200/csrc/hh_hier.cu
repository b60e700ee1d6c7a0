// hh_hier.cu -- the 3-vs-3 hierarchical (commander) environment of the reference, envs/env_hier.py
// HighLevelEnv on top of HHMARLBaseEnv (envs/env_base.py) and warsim/simulator/*.
//
// One commander step = _action_assess + up to 16 low-level sub-steps, each of which queries a frozen
// fight/escape policy for every live aircraft (env_hier.py:114-140).  As for levels 4/5 the networks are
// batched across arenas outside this library, so a sub-step is two launches around two network batches:
//   agents (ids 1-3) never observe anything that changes within the sub-step before they act (they only
//   see opponents, who act after them), opponents (ids 4-6) observe the agents' fresh fire decisions:
//     hh_hier_begin  : _action_assess, low-level observations of the agents
//     hh_hier_agents : agents' _take_base_action, low-level observations of the opponents
//     hh_hier_tick   : opponents' _take_base_action, do_tick, rewards, events, next observations of agents
//     hh_hier_end    : termination, (auto-reset), commander observations (34-d) and opp_to_attack lists
// Arenas whose sub-step loop has ended (kill event, surrounding event, 16 sub-steps) idle in lock-step.
//
// Mapping: one thread per arena over an array-of-structs arena record in global memory (~1 KB, L2
// resident); the id-ordered semantics of CmanoSimulator.do_tick are written sequentially.  This first
// version favours exactness over speed: the per-sub-step cost of config 5 is dominated by the six policy
// batches, not by these kernels.
#include <cuda_runtime.h>
#include <stdint.h>

#include <new>
#include <string>

#include "../../include/hhmarl_b200.h"
#include "hh_core.cuh"

namespace hh {
namespace hier {

constexpr int NU = 6, NA = 3;
constexpr int OBS_HL = 34, LL_STRIDE = 30;

struct Arena {  // mirrors hh_hier_arena in include/hhmarl_b200.h (plain data, copied verbatim by get/set_state)
  double lat[NU], lon[NU], hdg[NU], spd[NU], nhdg[NU], nspd[NU];
  double rlat[NU], rlon[NU], rhdg[NU], rnhdg[NU];
  double ota_dn[NU][3];
  double rewards[NA];
  unsigned long long dg;
  int32_t crem[NU], burst[NU], mrem[NU], mwait[NU], rid[NU];
  int32_t steps, alive_ag, alive_op, next_id, sub, kill_event, situation_event, active, err;
  uint32_t dc;
  int8_t ca[NU];  // commander actions, -1 = None
  uint8_t alive[NU], hasm[NU], actype[NU], ralive[NU], rage[NU], rtgt[NU];
  uint8_t ota_n[NU], ota_id[NU][3];
};

struct HParams {
  int n_arenas, horizon, level, friendly_kill, autoreset, action_assess, fight_p, fight_q;
  double map_size, rew_scale, glob_frac;
  uint32_t seed_lo, seed_hi, arena_base;
  int short_moves;   // 1: one-tick moves through geo::direct_tick (default), 0: full Karney direct (HH_SHORT_MOVES=0)
};

__device__ __forceinline__ double g_next(const Rng& r, Arena& A) { return g_random_at(r, A.dg++); }
__device__ __forceinline__ double maxspd(const Arena& A, int i) { return A.actype[i] == 1 ? 900.0 : 600.0; }
__device__ __forceinline__ bool shot(const Arena& A, int i) {
  return A.burst[i] > 0 || (A.actype[i] == 1 && A.hasm[i]);
}
__device__ __forceinline__ double fdeg(const Arena& A, int a, int b) {  // _focus_angle(a, b) in degrees
  return focus_deg(heading_vec(A.hdg[a]), A.lat[a], A.lon[a], A.lat[b], A.lon[b]);
}
__device__ __forceinline__ double draw(const Arena& A, int a, int b) { return dist_raw(A.lat[a], A.lon[a], A.lat[b], A.lon[b]); }

// _nearby_object (env_base.py:400-422): live enemies (or live team-mates) of unit i sorted by d_norm, stable
__device__ int nearby(const Arena& A, const Geom& g, int i, bool friendly, int* ids, double* dn) {
  const bool agent = i < NA;
  const int start = (agent != friendly) ? NA : 0;  // enemies of agents / friends of opponents start at NA
  int n = 0;
  for (int j = start; j < start + NA; ++j) {
    if (j == i || !A.alive[j]) continue;
    const double d = g.inv_diag * draw(A, i, j);
    int k = n;
    while (k > 0 && dn[k - 1] > d) {
      dn[k] = dn[k - 1];
      ids[k] = ids[k - 1];
      --k;
    }
    dn[k] = d;
    ids[k] = j;
    ++n;
  }
  return n;
}

// friendly_ac_values (env_base.py:166-183)
__device__ void friend_values(const Arena& A, const Geom& g, int i, int f, float* out) {
  if (f >= 0 && A.alive[f]) {
    double x, y;
    rel_pos(g, A.lat[f], A.lon[f], x, y);
    out[0] = (float)x;
    out[1] = (float)y;
    out[2] = (float)focus_norm_from_deg(fdeg(A, i, f));
    out[3] = (float)focus_norm_from_deg(fdeg(A, f, i));
    out[4] = (float)(g.inv_diag * draw(A, i, f));
  } else {
    for (int k = 0; k < 5; ++k) out[k] = 0.0f;
  }
}

// opp_ac_values (env_base.py:185-212); mode 0 "fight" (9), 1 "esc" (9), 2 "HighLevel" (10)
__device__ int enemy_values(const Arena& A, const Geom& g, int mode, int o, int i, double dist, float* out) {
  double x, y;
  int n = 0;
  rel_pos(g, A.lat[o], A.lon[o], x, y);
  out[n++] = (float)x;
  out[n++] = (float)y;
  out[n++] = (float)clip(A.spd[o] / maxspd(A, o), 0.0, 1.0);
  out[n++] = (float)hdg_feature(A.hdg[o]);
  out[n++] = (float)hdiff_norm(heading_vec(A.hdg[o]), heading_vec(A.hdg[i]));
  const double f_io = fdeg(A, i, o), f_oi = fdeg(A, o, i);
  if (mode == 0) {
    out[n++] = (float)focus_norm_from_deg(f_oi);
    out[n++] = (float)aspect_from_deg(f_io);
  } else {
    out[n++] = (float)focus_norm_from_deg(f_io);
    out[n++] = (float)focus_norm_from_deg(f_oi);
  }
  if (mode == 2) {
    out[n++] = (float)aspect_from_deg(f_io);
    out[n++] = (float)aspect_from_deg(f_oi);
  }
  out[n++] = (float)dist;
  if (mode != 2) out[n++] = shot(A, o) ? 1.0f : 0.0f;
  return n;
}

// HighLevelEnv.lowlevel_state (env_hier.py:100-112) for unit i; returns the observation length
__device__ int lowlevel_obs(Arena& A, const Geom& g, int i, float* out) {
  int fid[3];
  double fdn[3];
  const int nf = nearby(A, g, i, true, fid, fdn);
  const int fri = nf ? fid[0] : -1;
  const bool ac1 = A.actype[i] == 1;
  int n = 0;
  double x, y;
  rel_pos(g, A.lat[i], A.lon[i], x, y);
  out[n++] = (float)x;
  out[n++] = (float)y;
  out[n++] = (float)clip(A.spd[i] / maxspd(A, i), 0.0, 1.0);
  out[n++] = (float)hdg_feature(A.hdg[i]);
  if (A.ca[i] != 0) {  // fight_state_values (env_base.py:111-135) against the commander's choice
    int idx = A.ca[i] - 1;
    if (idx < 0) idx += A.ota_n[i];
    if (idx < 0 || idx >= A.ota_n[i]) { atomicOr(&A.err, 128); idx = 0; }
    const int o = A.ota_id[i][idx];
    const double dn = A.ota_dn[i][idx];  // stale distance of the commander step
    out[n++] = (float)focus_norm_from_deg(fdeg(A, i, o));
    out[n++] = (float)aspect_from_deg(fdeg(A, o, i));
    out[n++] = (float)hdiff_norm(heading_vec(A.hdg[i]), heading_vec(A.hdg[o]));
    out[n++] = (float)dn;
    out[n++] = (float)clip((double)A.crem[i] / 300.0, 0.0, 1.0);
    if (ac1) {
      out[n++] = (float)clip((double)A.mrem[i] / 8.0, 0.0, 1.0);
      out[n++] = A.mwait[i] == 0 ? 1.0f : 0.0f;
      out[n++] = (A.hasm[i] || A.burst[i] > 0) ? 1.0f : 0.0f;
    } else {
      out[n++] = A.burst[i] > 0 ? 1.0f : 0.0f;
    }
    n += enemy_values(A, g, 0, o, i, dn, out + n);
  } else {  // esc_state_values (env_base.py:137-164) over the stored opponent list
    out[n++] = (float)clip((double)A.crem[i] / 300.0, 0.0, 1.0);
    if (ac1) out[n++] = (float)clip((double)A.mrem[i] / 8.0, 0.0, 1.0);
    out[n++] = shot(A, i) ? 1.0f : 0.0f;
    int filled = 0;
    for (int k = 0; k < A.ota_n[i]; ++k) {
      filled += enemy_values(A, g, 1, A.ota_id[i][k], i, A.ota_dn[i][k], out + n + filled);
      if (filled == 18) break;
    }
    for (; filled < 18; ++filled) out[n + filled] = 0.0f;
    n += 18;
  }
  friend_values(A, g, i, fri, out + n);
  n += 5;
  for (int k = n; k < LL_STRIDE; ++k) out[k] = 0.0f;
  return n;
}

__device__ void write_ll_unit(Arena& A, const Geom& g, int i, float* row, uint8_t* info_out) {
  // bits 1-2 (policy kind) are constant for the whole commander step, bit 0 = the unit queries its policy now
  uint8_t info = (A.ca[i] == 0 ? 2 : 0) | (A.actype[i] == 2 ? 4 : 0);
  if (A.active && A.alive[i]) {
    lowlevel_obs(A, g, i, row);
    info |= 1;
  } else {
    for (int k = 0; k < LL_STRIDE; ++k) row[k] = 0.0f;
  }
  *info_out = info;
}

__device__ void write_ll(Arena& A, const Geom& g, int first, float* ll_obs, uint8_t* ll_info) {
  for (int i = first; i < first + NA; ++i) {
    // bits 1-2 (policy kind) are constant for the whole commander step, bit 0 = the unit queries its policy now
    uint8_t info = (A.ca[i] == 0 ? 2 : 0) | (A.actype[i] == 2 ? 4 : 0);
    float* row = ll_obs + i * LL_STRIDE;
    if (A.active && A.alive[i]) {
      lowlevel_obs(A, g, i, row);
      info |= 1;
    } else {
      for (int k = 0; k < LL_STRIDE; ++k) row[k] = 0.0f;
    }
    ll_info[i] = info;
  }
}

// HighLevelEnv.state (env_hier.py:49-98): commander observations + opp_to_attack lists
__device__ void commander_state(Arena& A, const Geom& g, float* obs) {
  for (int i = 0; i < NU; ++i) {
    A.ota_n[i] = 0;
    int ids[3];
    double dn[3];
    if (i < NA) {
      float* out = obs + i * OBS_HL;
      for (int k = 0; k < OBS_HL; ++k) out[k] = 0.0f;
      if (!A.alive[i]) continue;
      const int n_opps = nearby(A, g, i, false, ids, dn);
      if (!n_opps) continue;
      double x, y;
      rel_pos(g, A.lat[i], A.lon[i], x, y);
      out[0] = (float)x;
      out[1] = (float)y;
      out[2] = (float)clip(A.spd[i] / maxspd(A, i), 0.0, 1.0);
      out[3] = (float)hdg_feature(A.hdg[i]);
      int filled = 0;
      for (int k = 0; k < n_opps; ++k) {
        filled += enemy_values(A, g, 2, ids[k], i, dn[k], out + 4 + filled);
        A.ota_id[i][A.ota_n[i]] = (uint8_t)ids[k];
        A.ota_dn[i][A.ota_n[i]] = dn[k];
        A.ota_n[i] += 1;
        if (filled == 20) break;
      }
      int fid[3];
      double fdn[3];
      const int nf = nearby(A, g, i, true, fid, fdn);
      for (int k = 0; k < nf && k < 2; ++k) friend_values(A, g, i, fid[k], out + 24 + 5 * k);
    } else if (A.alive[i]) {
      const int n = nearby(A, g, i, false, ids, dn);
      for (int k = 0; k < n; ++k) {
        A.ota_id[i][k] = (uint8_t)ids[k];
        A.ota_dn[i][k] = dn[k];
      }
      A.ota_n[i] = (uint8_t)n;
    }
  }
}

// _take_base_action(mode="HighLevel") (env_base.py:214-238) + Rafale.fire_missile (ac1.py:72-79)
__device__ void base_action_hl(Arena& A, const Rng& rng, int i, const int32_t* act) {
  int idx = A.ca[i] - 1;
  if (idx < 0) idx += A.ota_n[i];          // escape (0) indexes [-1]: the LAST listed opponent
  if (idx < 0 || idx >= A.ota_n[i]) { A.err |= 128; return; }
  const int opp = A.ota_id[i][idx];
  const double h = pymod(A.hdg[i] + (double)((act[0] - 6) * 15), 360.0);
  if (h >= 360.0 || h < 0.0) A.err |= ERR_HEADING;
  A.nhdg[i] = h;
  A.nspd[i] = 100.0 + ((maxspd(A, i) - 100.0) / 8.0) * (double)act[1];
  if (act[2] != 0 && A.crem[i] > 0) {
    const int bt = A.actype[i] == 1 ? 5 : 3;
    A.burst[i] = A.crem[i] < bt ? A.crem[i] : bt;
  }
  if (A.actype[i] == 1 && act[3] != 0) {
    if (A.mrem[i] > 0 && !A.hasm[i] && A.mwait[i] == 0) {
      if (launch_gate(A.lat[i], A.lon[i], A.hdg[i], A.lat[opp], A.lon[opp])) {
        A.rlat[i] = A.lat[i];
        A.rlon[i] = A.lon[i];
        A.rhdg[i] = A.hdg[i];
        A.rnhdg[i] = A.hdg[i];
        A.ralive[i] = 1;
        A.rage[i] = 0;
        A.rtgt[i] = (uint8_t)opp;
        A.rid[i] = A.next_id++;
        A.hasm[i] = 1;
        A.mrem[i] -= 1;
      }
      A.mwait[i] = randint_from(8, 12, g_next(rng, A));
    }
  }
  if (A.mwait[i] > 0 && !A.hasm[i]) A.mwait[i] -= 1;
}

struct Events {
  int n, killer[NU], victim[NU];
};

// CmanoSimulator.do_tick (cmano_simulator.py:138-157) with Rafale/RafaleLong/Rocket.update
__device__ void do_tick(Arena& A, const Rng& rng, const HParams& P, Events& ev) {
  uint8_t alive0[NU], ralive0[NU];
  for (int i = 0; i < NU; ++i) { alive0[i] = A.alive[i]; ralive0[i] = A.ralive[i]; }
  ev.n = 0;
  for (int i = 0; i < NU; ++i) {
    if (!alive0[i]) continue;
    const bool ac1 = A.actype[i] == 1;
    const double max_deg = ac1 ? 5.0 : 3.5, max_kn = ac1 ? 35.0 : 28.0;
    if (A.hdg[i] != A.nhdg[i]) {
      const double d = signed_heading_diff(A.hdg[i], A.nhdg[i]);
      A.hdg[i] = fabs(d) <= max_deg ? A.nhdg[i] : pymod(A.hdg[i] + (d >= 0.0 ? max_deg : -max_deg), 360.0);
    }
    if (A.spd[i] != A.nspd[i]) {
      const double d = A.nspd[i] - A.spd[i];
      A.spd[i] = fabs(d) <= max_kn ? A.nspd[i] : A.spd[i] + (d >= 0.0 ? max_kn : -max_kn);
    }
    if (A.burst[i] > 0) {
      A.burst[i] -= 1;
      A.crem[i] = A.crem[i] > 0 ? A.crem[i] - 1 : 0;
      const double range = ac1 ? 2.0 : 4.5, half_w = (ac1 ? 10.0 : 7.0) / 2.0;
      const double p_hit = ac1 ? 0.75 / (5.0 / 1.0) : 0.9 / (3.0 / 1.0);
      uint8_t snap[NU];
      for (int j = 0; j < NU; ++j) snap[j] = A.alive[j];  // list(sim.active_units.values()) at this moment
      for (int j = 0; j < NU; ++j) {
        if (j == i || !snap[j]) continue;
        if (!(P.friendly_kill || ((i < NA) != (j < NA)))) continue;
        if (unit_in_cannon_range(A.lat[i], A.lon[i], A.hdg[i], A.lat[j], A.lon[j], range, half_w)) {
          if (c_random_at(rng, A.dc++) < p_hit) {
            A.alive[j] = 0;
            ev.killer[ev.n] = i;
            ev.victim[ev.n] = j;
            ev.n += 1;
          }
        }
      }
    }
    if (ac1 && A.hasm[i]) {
      if (!A.ralive[i]) A.hasm[i] = 0;
      else A.rnhdg[i] = clip(__dmul_rn(A.rhdg[i], uniform_from(0.95, 1.05, g_next(rng, A))), 0.0, 359.0);
    }
    if (A.spd[i] > 0.0) {
      const double2 q = P.short_moves ? geo::direct_tick(A.lat[i], A.lon[i], A.hdg[i], A.spd[i] * kKnotsToMs * 1.0)
                                     : geo::direct(A.lat[i], A.lon[i], A.hdg[i], A.spd[i] * kKnotsToMs * 1.0);
      A.lat[i] = q.x;
      A.lon[i] = q.y;
    }
  }
  // rockets in launch order
  for (;;) {
    int s = -1;
    for (int k = 0; k < NU; ++k)
      if (ralive0[k] && (s < 0 || A.rid[k] < A.rid[s])) s = k;
    if (s < 0) break;
    ralive0[s] = 0;
    const int t = A.rtgt[s];
    if (within_1km(A.rlat[s], A.rlon[s], A.lat[t], A.lon[t]) && A.alive[t]) {
      A.ralive[s] = 0;
      A.alive[t] = 0;
      ev.killer[ev.n] = s;
      ev.victim[ev.n] = t;
      ev.n += 1;
      continue;
    }
    if (P.friendly_kill) {
      const int f = (s + 1) == 2 ? 0 : 1;  // friendly_id = 1 if source.id == 2 else 2 (rocket_unit.py:46)
      if (A.alive[f] && within_1km(A.rlat[s], A.rlon[s], A.lat[f], A.lon[f])) {
        A.ralive[s] = 0;
        A.alive[f] = 0;
        ev.killer[ev.n] = s;
        ev.victim[ev.n] = f;
        ev.n += 1;
        continue;
      }
    }
    if (A.rage[s] > 10) {
      A.ralive[s] = 0;
      continue;
    }
    if (A.rhdg[s] != A.rnhdg[s]) {
      const double d = signed_heading_diff(A.rhdg[s], A.rnhdg[s]);
      A.rhdg[s] = fabs(d) <= 10.0 ? A.rnhdg[s] : A.rhdg[s] + (d >= 0.0 ? 10.0 : -10.0);
    }
    const double2 q = P.short_moves ? geo::direct_tick(A.rlat[s], A.rlon[s], A.rhdg[s], rocket_speed(A.rage[s]) * kKnotsToMs * 1.0)
                                   : geo::direct(A.rlat[s], A.rlon[s], A.rhdg[s], rocket_speed(A.rage[s]) * kKnotsToMs * 1.0);
    A.rlat[s] = q.x;
    A.rlon[s] = q.y;
    A.rage[s] += 1;
  }
}

// ---- the same tick, split for the staged kernels (tick_kernel below): every unit's rate limits and its one-tick move are
// independent of the other units (a shooter's cannon test reads only POSITIONS of the others and its OWN heading), so they run
// unit-parallel ahead of the id-ordered pass, which commits the positions in id order (targets with a lower id are seen at
// their new position, higher ids at their old one, the shooter itself at its old one: cmano_simulator.py:142, ac1.py:81-133).
struct TickTmp {
  double nlat[NU], nlon[NU];   // aircraft positions after this tick's move
  uint8_t alive0[NU], ralive0[NU], rmove[NU];
};

__device__ void tick_unit_prepare(Arena& A, const HParams& P, int i, TickTmp& T) {   // one thread per (arena, unit)
  T.alive0[i] = A.alive[i];
  T.ralive0[i] = A.ralive[i];
  T.rmove[i] = 0;
  if (!A.alive[i]) return;
  const bool ac1 = A.actype[i] == 1;
  const double max_deg = ac1 ? 5.0 : 3.5, max_kn = ac1 ? 35.0 : 28.0;
  if (A.hdg[i] != A.nhdg[i]) {
    const double d = signed_heading_diff(A.hdg[i], A.nhdg[i]);
    A.hdg[i] = fabs(d) <= max_deg ? A.nhdg[i] : pymod(A.hdg[i] + (d >= 0.0 ? max_deg : -max_deg), 360.0);
  }
  if (A.spd[i] != A.nspd[i]) {
    const double d = A.nspd[i] - A.spd[i];
    A.spd[i] = fabs(d) <= max_kn ? A.nspd[i] : A.spd[i] + (d >= 0.0 ? max_kn : -max_kn);
  }
  T.nlat[i] = A.lat[i];
  T.nlon[i] = A.lon[i];
  if (A.spd[i] > 0.0) {
    const double2 q = P.short_moves ? geo::direct_tick(A.lat[i], A.lon[i], A.hdg[i], A.spd[i] * kKnotsToMs * 1.0)
                                   : geo::direct(A.lat[i], A.lon[i], A.hdg[i], A.spd[i] * kKnotsToMs * 1.0);
    T.nlat[i] = q.x;
    T.nlon[i] = q.y;
  }
}

__device__ void tick_resolve(Arena& A, const Rng& rng, const HParams& P, TickTmp& T, Events& ev) {   // one thread per arena
  ev.n = 0;
  for (int i = 0; i < NU; ++i) {
    if (!T.alive0[i]) continue;
    const bool ac1 = A.actype[i] == 1;
    if (A.burst[i] > 0) {
      A.burst[i] -= 1;
      A.crem[i] = A.crem[i] > 0 ? A.crem[i] - 1 : 0;
      const double range = ac1 ? 2.0 : 4.5, half_w = (ac1 ? 10.0 : 7.0) / 2.0;
      const double p_hit = ac1 ? 0.75 / (5.0 / 1.0) : 0.9 / (3.0 / 1.0);
      uint8_t snap[NU];
      for (int j = 0; j < NU; ++j) snap[j] = A.alive[j];  // list(sim.active_units.values()) at this moment
      for (int j = 0; j < NU; ++j) {
        if (j == i || !snap[j]) continue;
        if (!(P.friendly_kill || ((i < NA) != (j < NA)))) continue;
        if (unit_in_cannon_range(A.lat[i], A.lon[i], A.hdg[i], A.lat[j], A.lon[j], range, half_w)) {
          if (c_random_at(rng, A.dc++) < p_hit) {
            A.alive[j] = 0;
            ev.killer[ev.n] = i;
            ev.victim[ev.n] = j;
            ev.n += 1;
          }
        }
      }
    }
    if (ac1 && A.hasm[i]) {
      if (!A.ralive[i]) A.hasm[i] = 0;
      else A.rnhdg[i] = clip(__dmul_rn(A.rhdg[i], uniform_from(0.95, 1.05, g_next(rng, A))), 0.0, 359.0);
    }
    A.lat[i] = T.nlat[i];
    A.lon[i] = T.nlon[i];
  }
  // rockets in launch order: proximity kills and lifetime here, the surviving rockets' moves unit-parallel afterwards
  for (;;) {
    int s = -1;
    for (int k = 0; k < NU; ++k)
      if (T.ralive0[k] && (s < 0 || A.rid[k] < A.rid[s])) s = k;
    if (s < 0) break;
    T.ralive0[s] = 0;
    const int t = A.rtgt[s];
    if (within_1km(A.rlat[s], A.rlon[s], A.lat[t], A.lon[t]) && A.alive[t]) {
      A.ralive[s] = 0;
      A.alive[t] = 0;
      ev.killer[ev.n] = s;
      ev.victim[ev.n] = t;
      ev.n += 1;
      continue;
    }
    if (P.friendly_kill) {
      const int f = (s + 1) == 2 ? 0 : 1;  // friendly_id = 1 if source.id == 2 else 2 (rocket_unit.py:46)
      if (A.alive[f] && within_1km(A.rlat[s], A.rlon[s], A.lat[f], A.lon[f])) {
        A.ralive[s] = 0;
        A.alive[f] = 0;
        ev.killer[ev.n] = s;
        ev.victim[ev.n] = f;
        ev.n += 1;
        continue;
      }
    }
    if (A.rage[s] > 10) {
      A.ralive[s] = 0;
      continue;
    }
    T.rmove[s] = 1;
  }
}

__device__ void tick_rocket_move(Arena& A, const HParams& P, int s, const TickTmp& T) {   // one thread per (arena, rocket slot)
  if (!T.rmove[s]) return;
  if (A.rhdg[s] != A.rnhdg[s]) {
    const double d = signed_heading_diff(A.rhdg[s], A.rnhdg[s]);
    A.rhdg[s] = fabs(d) <= 10.0 ? A.rnhdg[s] : A.rhdg[s] + (d >= 0.0 ? 10.0 : -10.0);
  }
  const double2 q = P.short_moves ? geo::direct_tick(A.rlat[s], A.rlon[s], A.rhdg[s], rocket_speed(A.rage[s]) * kKnotsToMs * 1.0)
                                 : geo::direct(A.rlat[s], A.rlon[s], A.rhdg[s], rocket_speed(A.rage[s]) * kKnotsToMs * 1.0);
  A.rlat[s] = q.x;
  A.rlon[s] = q.y;
  A.rage[s] += 1;
}

// _combat_rewards(mode="HighLevel") (env_base.py:240-310) + HighLevelEnv._get_rewards (env_hier.py:210-224)
__device__ bool get_rewards(Arena& A, const Geom& g, const HParams& P, const Events& ev) {
  const double s = P.rew_scale;
  double rews[NA] = {0.0, 0.0, 0.0};
  bool destroyed[NA] = {false, false, false}, kill = false;
  for (int i = 0; i < NU; ++i) {
    if (A.alive[i] && !in_boundary(g, A.lat[i], A.lon[i])) {
      A.alive[i] = 0;
      kill = true;
      if (i < NA) { rews[i] += -2.0 * s; destroyed[i] = true; A.alive_ag -= 1; }
      else A.alive_op -= 1;
    }
  }
  for (int k = 0; k < ev.n; ++k) {
    const int kl = ev.killer[k], v = ev.victim[k];
    if (kl < NA) {
      if (v >= NA) { rews[kl] += 1.0; A.alive_op -= 1; }
      else A.alive_ag -= 1;
    } else {
      if (v < NA) { rews[v] += -1.0 * s; destroyed[v] = true; A.alive_ag -= 1; }
      else A.alive_op -= 1;
    }
    kill = true;
  }
  for (int i = 0; i < NA; ++i) {
    if (A.alive[i] || destroyed[i]) {
      if (P.glob_frac > 0.0) {
        double others = 0.0;
        for (int j = 0; j < NA; ++j)
          if (j != i) others += rews[j];
        A.rewards[i] += rews[i] + P.glob_frac * others;
      } else {
        A.rewards[i] += rews[i];
      }
    }
  }
  return kill;
}

// HighLevelEnv._surrounding_event (env_hier.py:192-208)
__device__ bool surrounding_event(const Arena& A) {
  bool event = false;
  for (int i = 0; i < NA && !event; ++i)
    for (int j = NA; j < NU && !event; ++j)
      if (A.alive[i] && A.alive[j]) {
        event = false;
        if (draw(A, i, j) < 0.1)
          if (fdeg(A, i, j) < 15.0 || fdeg(A, j, i) < 15.0) event = true;
      }
  return event;
}

// HighLevelEnv._action_assess (env_hier.py:142-190)
__device__ void action_assess(Arena& A, const Rng& rng, const HParams& P, const int32_t* cmd) {
  for (int i = 0; i < NU; ++i) {
    if (i < NA) A.ca[i] = (int8_t)cmd[i];
    if (A.alive[i]) {
      if (i < NA) {
        A.rewards[i] = 0.0;
        if (A.ca[i] > 0) {
          const int idx = A.ca[i] - 1;
          int opp = -1;
          if (idx < A.ota_n[i]) opp = A.ota_id[i][idx];
          else A.ca[i] = 1;
          if (opp < 0) A.rewards[i] = -0.1;
          if (P.action_assess && opp >= 0)
            A.rewards[i] = (draw(A, i, opp) < 0.1 && fdeg(A, i, opp) < 15.0 && fdeg(A, opp, i) > 40.0) ? 0.1 : 0.0;
        } else if (P.action_assess) {
          if (A.ota_n[i] > 0) {
            const int cl = A.ota_id[i][0];
            if (draw(A, cl, i) < 0.1 && fdeg(A, cl, i) < 15.0 && fdeg(A, i, cl) > 40.0) A.rewards[i] = 0.1;
          } else {
            A.err |= 128;
          }
        }
      } else {
        int ag;
        if (g_next(rng, A) * (double)P.fight_q >= (double)(P.fight_q - P.fight_p)) {   // choices([0,1], [q-p, p])
          const int possible = A.ota_n[i];
          bool other = false;
          if (possible > 1) other = g_next(rng, A) * 4.0 >= 1.0;                         // choices([0,1], [1, 3])
          ag = other ? randint_from(2, possible, g_next(rng, A)) : 1;
        } else {
          ag = 0;
        }
        A.ca[i] = (int8_t)ag;
      }
    } else {
      if (i < NA) A.rewards[i] = 0.0;
      A.ca[i] = -1;
    }
  }
}

// HHMARLBaseEnv.reset + _reset_scenario("HighLevel") + HighLevelEnv._sample_state (env_base.py:62-77,551-585;
// env_hier.py:226-250)
__device__ void reset_arena(Arena& A, const Rng& rng, const HParams& P) {
  A.steps = 0;
  A.alive_ag = A.alive_op = 0;
  A.next_id = 1;
  A.sub = 0;
  A.kill_event = A.situation_event = A.active = 0;
  const int r = randint_from(1, 2, g_next(rng, A));
  for (int u = 0; u < NU; ++u) {
    const int group = u / NA, i = u % NA;
    const bool west = (group == 0) == (r == 1);
    const double step = 0.4 / (double)NA;
    const double x = west ? uniform_from(7.07, 7.22, g_next(rng, A)) : uniform_from(7.28, 7.43, g_next(rng, A));
    const double y = uniform_from(__dadd_rn(5.07, __dmul_rn((double)i, step)), __dadd_rn(5.12, __dmul_rn((double)i, step)),
                                  g_next(rng, A));
    const int a = randint_from(0, 359, g_next(rng, A));
    const int ac = i <= 1 ? i + 1 : randint_from(1, 2, g_next(rng, A));
    A.lat[u] = y;
    A.lon[u] = x;
    A.hdg[u] = A.nhdg[u] = (double)a;
    A.spd[u] = A.nspd[u] = (P.level <= 2 && group == 1) ? 0.0 : 100.0;
    A.crem[u] = 300;
    A.burst[u] = 0;
    A.mrem[u] = ac == 1 ? 8 : 0;
    A.mwait[u] = 0;
    A.alive[u] = 1;
    A.hasm[u] = 0;
    A.actype[u] = (uint8_t)ac;
    A.ralive[u] = 0;
    A.rage[u] = A.rtgt[u] = 0;
    A.rid[u] = 0;
    A.rlat[u] = A.rlon[u] = A.rhdg[u] = A.rnhdg[u] = 0.0;
    A.ota_n[u] = 0;
    A.ca[u] = -1;
    A.next_id += 1;
    if (group == 0) A.alive_ag += 1; else A.alive_op += 1;
  }
  for (int i = 0; i < NA; ++i) A.rewards[i] = 0.0;
}

#define HH_HIER_PROLOGUE                                                   \
  const int a = blockIdx.x * blockDim.x + threadIdx.x;                     \
  if (a >= P.n_arenas) return;                                             \
  Arena& A = arenas[a];                                                    \
  const Geom g = make_geom(P.map_size);                                    \
  const Rng rng{P.seed_lo, P.seed_hi, P.arena_base + (uint32_t)a};

__global__ void __launch_bounds__(64) reset_kernel(Arena* arenas, HParams P, const uint8_t* mask, int first, float* obs) {
  HH_HIER_PROLOGUE
  if (first) {
    A.dg = 0;
    A.dc = 0;
    A.err = 0;
  }
  if (first || !mask || mask[a]) reset_arena(A, rng, P);
  commander_state(A, g, obs + (size_t)a * NA * OBS_HL);
}

// ---- staged kernels of the sub-step loop.  A CTA owns kAr arenas whose records are staged in shared memory (coalesced
// load / store of the ~0.9 KB array-of-structs records); the id-ordered parts run one thread per arena, the per-unit parts
// (rate limits + one-tick geodesic moves, rocket moves, the low-level observations) one thread per (arena, unit).
constexpr int kAr = 32, kHT = kAr * NU;   // 192 threads
static_assert(sizeof(Arena) % 8 == 0, "records are copied in 8-byte words");

struct HStage {
  Arena rec[kAr];
  float rows[kAr][NA][LL_STRIDE];
  TickTmp tmp[kAr];
  uint8_t info[kAr][NA];
};

__device__ __forceinline__ int stage_load(HStage& S, const Arena* arenas, int n_arenas) {
  const int arena0 = blockIdx.x * kAr;
  const int n_valid = min(kAr, n_arenas - arena0);
  const unsigned long long* src = reinterpret_cast<const unsigned long long*>(arenas + arena0);
  unsigned long long* dst = reinterpret_cast<unsigned long long*>(S.rec);
  const int words = n_valid * (int)(sizeof(Arena) / 8);
  for (int k = threadIdx.x; k < words; k += kHT) dst[k] = src[k];
  __syncthreads();
  return n_valid;
}
// writes the records back and the three staged observation rows / info bytes of every arena (units first .. first + 2)
__device__ __forceinline__ void stage_store(HStage& S, Arena* arenas, int n_valid, int first, float* ll_obs, uint8_t* ll_info) {
  __syncthreads();
  const int arena0 = blockIdx.x * kAr;
  unsigned long long* dst = reinterpret_cast<unsigned long long*>(arenas + arena0);
  const unsigned long long* src = reinterpret_cast<const unsigned long long*>(S.rec);
  const int words = n_valid * (int)(sizeof(Arena) / 8);
  for (int k = threadIdx.x; k < words; k += kHT) dst[k] = src[k];
  for (int k = threadIdx.x; k < n_valid * NA * LL_STRIDE; k += kHT) {
    const int al = k / (NA * LL_STRIDE), r = k - al * (NA * LL_STRIDE);
    ll_obs[((size_t)(arena0 + al) * NU + first) * LL_STRIDE + r] = (&S.rows[al][0][0])[r];
  }
  for (int k = threadIdx.x; k < n_valid * NA; k += kHT) {
    const int al = k / NA, u = k - al * NA;
    ll_info[(size_t)(arena0 + al) * NU + first + u] = S.info[al][u];
  }
}
__device__ __forceinline__ void stage_write_ll(HStage& S, const Geom& g, int n_valid, int first) {
  const int t = threadIdx.x;
  if (t < kAr * NA) {
    const int al = t / NA, u = t - al * NA;
    if (al < n_valid) write_ll_unit(S.rec[al], g, first + u, S.rows[al][u], &S.info[al][u]);
  }
}

__global__ void __launch_bounds__(kHT) begin_kernel(Arena* arenas, HParams P, const int32_t* cmd, float* ll_obs, uint8_t* ll_info) {
  extern __shared__ __align__(16) unsigned char hsm[];
  HStage& S = *reinterpret_cast<HStage*>(hsm);
  const int n_valid = stage_load(S, arenas, P.n_arenas);
  const int arena0 = blockIdx.x * kAr, t = threadIdx.x;
  const Geom g = make_geom(P.map_size);
  if (t < n_valid) {
    Arena& A = S.rec[t];
    const int a = arena0 + t;
    const Rng rng{P.seed_lo, P.seed_hi, P.arena_base + (uint32_t)a};
    action_assess(A, rng, P, cmd + (size_t)a * NA);
    A.sub = 0;
    A.kill_event = A.situation_event = 0;
    A.active = 1;
  }
  __syncthreads();
  stage_write_ll(S, g, n_valid, 0);
  stage_store(S, arenas, n_valid, 0, ll_obs, ll_info);
  if (t < n_valid) {
    const Arena& A = S.rec[t];
    const size_t a = (size_t)(arena0 + t);
    for (int i = NA; i < NU; ++i)   // opponents' policy kinds are known already (their observations come later)
      ll_info[a * NU + i] = (A.ca[i] == 0 ? 2 : 0) | (A.actype[i] == 2 ? 4 : 0);
  }
  __syncthreads();                  // the agents' info bytes above were written by other threads of this CTA
  if (t < n_valid) {
    // bit 3 (this call only): alive at the start of the commander step -- a unit that is dead now makes no policy query
    // in any of the step's sub-steps, so the host leaves it out of the batched forwards
    const Arena& A = S.rec[t];
    const size_t a = (size_t)(arena0 + t);
    for (int i = 0; i < NU; ++i)
      if (A.alive[i]) ll_info[a * NU + i] |= 8;
  }
}

__global__ void __launch_bounds__(kHT) agents_kernel(Arena* arenas, HParams P, const int32_t* act, float* ll_obs, uint8_t* ll_info) {
  extern __shared__ __align__(16) unsigned char hsm[];
  HStage& S = *reinterpret_cast<HStage*>(hsm);
  const int n_valid = stage_load(S, arenas, P.n_arenas);
  const int arena0 = blockIdx.x * kAr, t = threadIdx.x;
  const Geom g = make_geom(P.map_size);
  if (t < n_valid) {
    Arena& A = S.rec[t];
    const int a = arena0 + t;
    const Rng rng{P.seed_lo, P.seed_hi, P.arena_base + (uint32_t)a};
    if (A.active)
      for (int i = 0; i < NA; ++i)
        if (A.alive[i]) base_action_hl(A, rng, i, act + ((size_t)a * NU + i) * 4);
  }
  __syncthreads();
  stage_write_ll(S, g, n_valid, NA);
  stage_store(S, arenas, n_valid, NA, ll_obs, ll_info);
}

__global__ void __launch_bounds__(kHT) tick_kernel(Arena* arenas, HParams P, const int32_t* act, float* ll_obs, uint8_t* ll_info) {
  extern __shared__ __align__(16) unsigned char hsm[];
  HStage& S = *reinterpret_cast<HStage*>(hsm);
  __shared__ Events evs[kAr];
  const int n_valid = stage_load(S, arenas, P.n_arenas);
  const int arena0 = blockIdx.x * kAr, t = threadIdx.x;
  const Geom g = make_geom(P.map_size);
  const int ual = t / NU, uu = t - ual * NU;      // (arena, unit) of the unit-parallel stages
  if (t < n_valid) {                              // opponents' _take_base_action (id order)
    Arena& A = S.rec[t];
    const int a = arena0 + t;
    const Rng rng{P.seed_lo, P.seed_hi, P.arena_base + (uint32_t)a};
    if (A.active)
      for (int i = NA; i < NU; ++i)
        if (A.alive[i]) base_action_hl(A, rng, i, act + ((size_t)a * NU + i) * 4);
  }
  __syncthreads();
  if (ual < n_valid && S.rec[ual].active) tick_unit_prepare(S.rec[ual], P, uu, S.tmp[ual]);
  __syncthreads();
  if (t < n_valid && S.rec[t].active) {
    const Rng rng{P.seed_lo, P.seed_hi, P.arena_base + (uint32_t)(arena0 + t)};
    tick_resolve(S.rec[t], rng, P, S.tmp[t], evs[t]);
  }
  __syncthreads();
  if (ual < n_valid && S.rec[ual].active) tick_rocket_move(S.rec[ual], P, uu, S.tmp[ual]);
  __syncthreads();
  if (t < n_valid) {
    Arena& A = S.rec[t];
    if (A.active) {
      A.kill_event = get_rewards(A, g, P, evs[t]) ? 1 : 0;
      if (A.sub > 10) A.situation_event = surrounding_event(A) ? 1 : 0;   // min_sub_steps = 10 (env_hier.py:120,133)
      A.sub += 1;
      A.steps += 1;
      A.active = (A.sub <= 15 && !A.kill_event && !A.situation_event) ? 1 : 0;   // n_sub_steps = 15 (env_hier.py:33,125)
    }
  }
  __syncthreads();
  stage_write_ll(S, g, n_valid, 0);
  stage_store(S, arenas, n_valid, 0, ll_obs, ll_info);
}

// HHMARLBaseEnv.step with args.eval_info (env_base.py:91-107): the step's `info` dict as int32[12] per arena =
// agents_win, opps_win, draw, agent_fight, agent_escape, opp_fight, opp_escape, agent_steps, opp_steps, opp1, opp2,
// opp3.  Reads the record as the last tick left it (aircraft that exist NOW, the commander actions as
// _action_assess left them), so it runs between the last hh_hier_tick and hh_hier_end (which may auto-reset).
__global__ void __launch_bounds__(64) eval_info_kernel(const Arena* arenas, HParams P, int32_t* info) {
  const int a = blockIdx.x * blockDim.x + threadIdx.x;
  if (a >= P.n_arenas) return;
  const Arena& A = arenas[a];
  int32_t o[12];
  for (int k = 0; k < 12; ++k) o[k] = 0;
  const bool before_horizon = A.steps < P.horizon;
  o[0] = (A.alive_op <= 0 && before_horizon) ? 1 : 0;
  o[1] = (A.alive_ag <= 0 && before_horizon) ? 1 : 0;
  o[2] = (!before_horizon && A.alive_ag > 0 && A.alive_op > 0) ? 1 : 0;
  for (int i = 0; i < NU; ++i) {
    if (!A.alive[i]) continue;
    const int v = A.ca[i];   // `if v:` -- None (-1) and 0 are falsy
    if (v > 0) {
      if (i < NA) {
        o[3] += 1;
        o[7] += 1;
        if (v <= 3) o[8 + v] += 1;
      } else {
        o[5] += 1;
        o[8] += 1;
      }
    } else if (i < NA) {
      o[4] += 1;
      o[7] += 1;
    } else {
      o[6] += 1;
      o[8] += 1;
    }
  }
  for (int k = 0; k < 12; ++k) info[(size_t)a * 12 + k] = o[k];
}

__global__ void __launch_bounds__(64) end_kernel(Arena* arenas, HParams P, float* obs, float* rew, uint8_t* done, int32_t* substeps) {
  HH_HIER_PROLOGUE
  const bool d = A.alive_ag <= 0 || A.alive_op <= 0 || A.steps >= P.horizon;
  for (int i = 0; i < NA; ++i) rew[(size_t)a * NA + i] = (float)A.rewards[i];
  done[a] = d ? 1 : 0;
  if (substeps) substeps[a] = A.sub;
  A.active = 0;
  if (d && P.autoreset) reset_arena(A, rng, P);
  commander_state(A, g, obs + (size_t)a * NA * OBS_HL);
}

// Row lists of the frozen low-level policies for one commander step, built on the device (no host synchronisation): which
// network a unit queries (fight / escape x aircraft type) is fixed by hh_hier_begin's info bytes.  List 2 k + h holds the
// flat unit indices (arena * 6 + unit) of policy kind k (0 fight AC1, 1 fight AC2, 2 escape AC1, 3 escape AC2) among the
// agents (h = 0, units 0-2) or the opponents (h = 1, units 3-5) that were alive at the start of the step.
__global__ void policy_rows_init_kernel(int cap, int32_t* __restrict__ ranges) {
  const int k = threadIdx.x;
  if (k < 8) {
    ranges[2 * k] = k * cap;
    ranges[2 * k + 1] = 0;
  }
}
__global__ void policy_rows_fill_kernel(int n_units, int cap, const uint8_t* __restrict__ ll_info, int32_t* __restrict__ rows,
                                        int32_t* __restrict__ ranges) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_units) return;
  const int v = ll_info[i];
  if (!(v & 8)) return;                                 // not alive at the start of the commander step: never queries
  const int bits = v & 6, kind = bits == 0 ? 0 : bits == 4 ? 1 : bits == 2 ? 2 : 3;
  const int list = 2 * kind + ((i % NU) >= NA ? 1 : 0);
  const int pos = atomicAdd(&ranges[2 * list + 1], 1);
  rows[(size_t)list * cap + pos] = i;
}

}  // namespace hier
}  // namespace hh

// ============================================================================================ C ABI
using hh::hier::Arena;
using hh::hier::HParams;

static_assert(sizeof(Arena) == sizeof(hh_hier_arena), "hh_hier_arena in the header must mirror the device record");

static thread_local std::string g_hier_error;
extern "C" const char* hh_hier_last_error(void) { return g_hier_error.c_str(); }

static int hfail(int code, const std::string& msg) {
  g_hier_error = msg;
  return code;
}
#define HHH_CUDA(expr)                                                                          \
  do {                                                                                          \
    cudaError_t _e = (expr);                                                                    \
    if (_e != cudaSuccess) return hfail(-2, std::string(#expr) + ": " + cudaGetErrorString(_e)); \
  } while (0)

struct hh_hier_env {
  HParams P{};
  Arena* arenas = nullptr;
  int device = 0;
  bool initialised = false;
  uint64_t launches = 0;
};

extern "C" int hh_hier_create(const hh_hier_config* c, int32_t n_arenas, int32_t device, hh_hier_env** out) {
  if (!c || !out) return hfail(-1, "hh_hier_create: null argument");
  if (n_arenas <= 0) return hfail(-1, "hh_hier_create: n_arenas must be positive");
  if (!(c->map_size > 0) || c->horizon <= 0) return hfail(-1, "hh_hier_create: bad map_size / horizon");
  if (c->hier_opp_fight_ratio < 0 || c->hier_opp_fight_ratio > 100) return hfail(-1, "hh_hier_create: fight ratio must be 0..100");
  HHH_CUDA(cudaSetDevice(device));
  hh_hier_env* e = new (std::nothrow) hh_hier_env();
  if (!e) return hfail(-3, "out of host memory");
  e->device = device;
  HParams& P = e->P;
  P.n_arenas = n_arenas;
  P.horizon = c->horizon;
  P.level = c->level;
  P.friendly_kill = c->friendly_kill;
  P.autoreset = c->autoreset;
  P.action_assess = c->hier_action_assess;
  int p = c->hier_opp_fight_ratio, q = 100, a = p, b = q;   // Fraction(ratio, 100).as_integer_ratio()
  while (b) { int t = a % b; a = b; b = t; }
  if (a > 0) { p /= a; q /= a; }
  P.fight_p = p;
  P.fight_q = q;
  P.map_size = c->map_size;
  P.rew_scale = c->rew_scale;
  P.glob_frac = c->glob_frac;
  P.seed_lo = (uint32_t)c->seed;
  P.seed_hi = (uint32_t)(c->seed >> 32);
  P.arena_base = (uint32_t)c->arena_base;
  {
    const char* sm = getenv("HH_SHORT_MOVES");
    P.short_moves = (sm && sm[0] == '0') ? 0 : 1;
  }
  cudaError_t ce = cudaMalloc(&e->arenas, sizeof(Arena) * (size_t)n_arenas);
  if (ce != cudaSuccess) {
    delete e;
    return hfail(-2, std::string("cudaMalloc: ") + cudaGetErrorString(ce));
  }
  cudaMemset(e->arenas, 0, sizeof(Arena) * (size_t)n_arenas);
  *out = e;
  return 0;
}

extern "C" void hh_hier_destroy(hh_hier_env* e) {
  if (!e) return;
  cudaSetDevice(e->device);
  if (e->arenas) cudaFree(e->arenas);
  delete e;
}

#define HIER_LAUNCH(kern, ...)                                                                        \
  do {                                                                                                \
    const int blocks = (e->P.n_arenas + 63) / 64;                                                     \
    hh::hier::kern<<<blocks, 64, 0, static_cast<cudaStream_t>(stream)>>>(e->arenas, e->P, __VA_ARGS__); \
    HHH_CUDA(cudaGetLastError());                                                                     \
    e->launches += 1;                                                                                 \
  } while (0)
// the staged kernels of the sub-step loop: kAr arenas per CTA, records in (dynamic) shared memory
#define HIER_LAUNCH_STAGED(kern, ...)                                                                 \
  do {                                                                                                \
    static bool opted[64] = {};                                                                       \
    if (e->device >= 0 && e->device < 64 && !opted[e->device]) {                                      \
      HHH_CUDA(cudaFuncSetAttribute(hh::hier::kern, cudaFuncAttributeMaxDynamicSharedMemorySize,      \
                                    (int)sizeof(hh::hier::HStage)));                                  \
      opted[e->device] = true;                                                                        \
    }                                                                                                 \
    const int blocks = (e->P.n_arenas + hh::hier::kAr - 1) / hh::hier::kAr;                           \
    hh::hier::kern<<<blocks, hh::hier::kHT, sizeof(hh::hier::HStage), static_cast<cudaStream_t>(stream)>>>( \
        e->arenas, e->P, __VA_ARGS__);                                                                \
    HHH_CUDA(cudaGetLastError());                                                                     \
    e->launches += 1;                                                                                 \
  } while (0)

extern "C" int hh_hier_reset(hh_hier_env* e, const uint8_t* mask_dev, float* obs_dev, void* stream) {
  if (!e || !obs_dev) return hfail(-1, "hh_hier_reset: null argument");
  HIER_LAUNCH(reset_kernel, mask_dev, e->initialised ? 0 : 1, obs_dev);
  e->initialised = true;
  return 0;
}
extern "C" int hh_hier_begin(hh_hier_env* e, const int32_t* cmd_dev, float* ll_obs_dev, uint8_t* ll_info_dev, void* stream) {
  if (!e || !cmd_dev || !ll_obs_dev || !ll_info_dev) return hfail(-1, "hh_hier_begin: null argument");
  if (!e->initialised) return hfail(-4, "hh_hier_begin: call hh_hier_reset first");
  HIER_LAUNCH_STAGED(begin_kernel, cmd_dev, ll_obs_dev, ll_info_dev);
  return 0;
}
extern "C" int hh_hier_agents(hh_hier_env* e, const int32_t* act_dev, float* ll_obs_dev, uint8_t* ll_info_dev, void* stream) {
  if (!e || !act_dev || !ll_obs_dev || !ll_info_dev) return hfail(-1, "hh_hier_agents: null argument");
  HIER_LAUNCH_STAGED(agents_kernel, act_dev, ll_obs_dev, ll_info_dev);
  return 0;
}
extern "C" int hh_hier_tick(hh_hier_env* e, const int32_t* act_dev, float* ll_obs_dev, uint8_t* ll_info_dev, void* stream) {
  if (!e || !act_dev || !ll_obs_dev || !ll_info_dev) return hfail(-1, "hh_hier_tick: null argument");
  HIER_LAUNCH_STAGED(tick_kernel, act_dev, ll_obs_dev, ll_info_dev);
  return 0;
}
extern "C" int hh_hier_end(hh_hier_env* e, float* obs_dev, float* rew_dev, uint8_t* done_dev, int32_t* substeps_dev, void* stream) {
  if (!e || !obs_dev || !rew_dev || !done_dev) return hfail(-1, "hh_hier_end: null argument");
  HIER_LAUNCH(end_kernel, obs_dev, rew_dev, done_dev, substeps_dev);
  return 0;
}
extern "C" int hh_hier_policy_rows(hh_hier_env* e, const uint8_t* ll_info_dev, int32_t* rows_dev, int32_t* ranges_dev, void* stream) {
  if (!e || !ll_info_dev || !rows_dev || !ranges_dev) return hfail(-1, "hh_hier_policy_rows: null argument");
  const int n_units = e->P.n_arenas * hh::hier::NU, cap = e->P.n_arenas * hh::hier::NA;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  hh::hier::policy_rows_init_kernel<<<1, 32, 0, st>>>(cap, ranges_dev);
  hh::hier::policy_rows_fill_kernel<<<(n_units + 255) / 256, 256, 0, st>>>(n_units, cap, ll_info_dev, rows_dev, ranges_dev);
  HHH_CUDA(cudaGetLastError());
  e->launches += 2;
  return 0;
}
extern "C" int hh_hier_eval_info(hh_hier_env* e, int32_t* info_dev, void* stream) {
  if (!e || !info_dev) return hfail(-1, "hh_hier_eval_info: null argument");
  if (!e->initialised) return hfail(-4, "hh_hier_eval_info: call hh_hier_reset first");
  const int blocks = (e->P.n_arenas + 63) / 64;
  hh::hier::eval_info_kernel<<<blocks, 64, 0, static_cast<cudaStream_t>(stream)>>>(e->arenas, e->P, info_dev);
  HHH_CUDA(cudaGetLastError());
  e->launches += 1;
  return 0;
}
extern "C" int hh_hier_get_state(hh_hier_env* e, hh_hier_arena* out_host) {
  if (!e || !out_host) return hfail(-1, "hh_hier_get_state: null argument");
  HHH_CUDA(cudaSetDevice(e->device));
  HHH_CUDA(cudaDeviceSynchronize());
  HHH_CUDA(cudaMemcpy(out_host, e->arenas, sizeof(Arena) * (size_t)e->P.n_arenas, cudaMemcpyDeviceToHost));
  return 0;
}
extern "C" int hh_hier_set_state(hh_hier_env* e, const hh_hier_arena* in_host) {
  if (!e || !in_host) return hfail(-1, "hh_hier_set_state: null argument");
  HHH_CUDA(cudaSetDevice(e->device));
  HHH_CUDA(cudaDeviceSynchronize());
  HHH_CUDA(cudaMemcpy(e->arenas, in_host, sizeof(Arena) * (size_t)e->P.n_arenas, cudaMemcpyHostToDevice));
  e->initialised = true;
  return 0;
}
extern "C" uint64_t hh_hier_launch_count(const hh_hier_env* e) { return e ? e->launches : 0; }
