// hh_api.cu -- fused step / reset kernels and the C ABI declared in include/hhmarl_b200.h.
//
// Kernel shape: a QUAD of lanes per arena (lane = aircraft, hh_quad.cuh), 8 arenas per warp,
// 4-warp CTAs (32 arenas) whose warps are re-aligned with a barrier at every phase boundary so that
// they stream the ~160 KB of step code through the instruction caches together (N = 8192 -> 256 CTAs,
// 6.9 warps / SM).  Each lane reads its own aircraft's slice of the 320 B
// struct-of-arrays arena state with coalesced 8-byte loads, the quad advances the arena through
//   action decode -> scripted opponents -> tick (kinematics, cannon, rockets; WGS84 FP64) ->
//   rewards / out-of-bounds / termination -> (auto-reset) -> observations
// exchanging cross-unit data with warp shuffles, and writes the state back once.  Observations
// are staged through shared memory so that the [N][26] and [N][24] float rows leave the SM as
// contiguous 16-byte stores.
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <new>
#include <string>
#include <vector>

#include "../../include/hhmarl_b200.h"
#include "hh_quad.cuh"
#include "hh_cta.cuh"
#include "hh_v4.cuh"
#include "hh_state_pack.h"

namespace hh {

#ifndef HH_CTA_THREADS
#define HH_CTA_THREADS 128
#endif
#ifndef HH_NO_PHASE_SYNC
#define HH_PHASE_SYNC 1
#endif
constexpr int kThreads = HH_CTA_THREADS;     // warps per CTA x 32; 8 arenas per warp
__device__ __forceinline__ void cta_sync() {
  if (kThreads == 32) __syncwarp(); else __syncthreads();
}
// Re-aligns the warps of a CTA at phase boundaries: the step is ~10 k instructions of straight-line code per
// warp; warps that drift apart each stream their own copy through the instruction caches (profiles/README.md).
__device__ __forceinline__ void phase_sync() {
#ifdef HH_PHASE_SYNC
  if (kThreads > 32) __syncthreads();
#endif
}
constexpr int kArenasPerCta = kThreads / 4;

template <int MODE>
struct ObsDims {
  static constexpr int D1 = MODE == 0 ? OBS_AC1 : OBS_ESC_AC1;
  static constexpr int D2 = MODE == 0 ? OBS_AC2 : OBS_ESC_AC2;
};

// contiguous, coalesced copy of `n_floats` staged floats (16-byte aligned on both sides)
__device__ __forceinline__ void flush_rows(float* __restrict__ dst, const float* __restrict__ src, int n_floats) {
  const int n4 = n_floats >> 2;
  const float4* s4 = reinterpret_cast<const float4*>(src);
  float4* d4 = reinterpret_cast<float4*>(dst);
  for (int k = threadIdx.x; k < n4; k += kThreads) d4[k] = s4[k];
  for (int k = (n4 << 2) + threadIdx.x; k < n_floats; k += kThreads) dst[k] = src[k];
}

// agents' observations (lanes 0 and 1 of each quad) -> shared staging -> contiguous rows in HBM
template <int MODE>
__device__ __forceinline__ void write_agent_obs(Lane& L, const Geom& g, int u, float* obs1, float* obs2, float* s1,
                                                float* s2, int arena0, int n_valid) {
  constexpr int D1 = ObsDims<MODE>::D1, D2 = ObsDims<MODE>::D2;
  const int al = threadIdx.x >> 2;
  const World W = gather_world(L, u);
  if (u < 2) L.ota = unit_observation(L, W, g, u, MODE, u == 0 ? s1 + al * D1 : s2 + al * D2);
  cta_sync();
  if (obs1) flush_rows(obs1 + (size_t)arena0 * D1, s1, n_valid * D1);
  if (obs2) flush_rows(obs2 + (size_t)arena0 * D2, s2, n_valid * D2);
}

// ------------------------------------------------------------------------------------------
// step phases (shared by the fused level 1-3 kernel and the split level 4-5 kernels)
// ------------------------------------------------------------------------------------------
struct PreTick {  // pre-tick view of the four aircraft, identical in the four lanes
  double lat4[4], lon4[4], hdg4[4];
  int alive_m;
};
__device__ __forceinline__ PreTick gather_pre(const Lane& L) {
  PreTick p;
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    p.lat4[j] = qshfl(L.lat, j);
    p.lon4[j] = qshfl(L.lon, j);
    p.hdg4[j] = qshfl(L.hdg, j);
  }
  p.alive_m = quad_ballot(L.alive);
  return p;
}
#define HH_P(field, i) pick4d(p.field[0], p.field[1], p.field[2], p.field[3], (i))

// One flat-plane relation per lane, all four lanes in parallel:
//   agents   : opp_stats[i][0] = focus(opp_to_attack -> self)              (env_hetero.py:169-170)
//   opponents: nearest agent; with WITH_L3 also sign / focus(self -> agent) (env_hetero.py:251-260)
template <bool WITH_L3>
__device__ __forceinline__ void pre_tick_relations(const Lane& L, int u, const Geom& g, const PreTick& p, int& near_t,
                                                   double& near_dn, double& rel_focus, int& rel_sign) {
  near_dn = 0.0;
  near_t = -1;
  if (u >= 2 && L.alive) near_t = nearest_enemy(g, u, L.lat, L.lon, p.lat4, p.lon4, p.alive_m, near_dn);
  int rel_x = -1, rel_y = -1;
  if (u < 2) {
    if (L.alive && L.ota != 0 && ((p.alive_m >> (L.ota - 1)) & 1)) { rel_x = L.ota - 1; rel_y = u; }
  } else if (WITH_L3 && near_t >= 0) {
    rel_x = u;
    rel_y = near_t;
  }
  rel_focus = 0.0;
  rel_sign = 1;
  if (rel_x >= 0) {
    const double xlat = HH_P(lat4, rel_x), xlon = HH_P(lon4, rel_x), xh = HH_P(hdg4, rel_x);
    const double ylat = HH_P(lat4, rel_y), ylon = HH_P(lon4, rel_y);
    rel_focus = focus_deg(heading_vec(xh), xlat, xlon, ylat, ylon);
    if (u >= 2) rel_sign = correct_angle_sign(xlat, xlon, xh, ylat, ylon);
  }
}

// _take_base_action (env_base.py:214-238) up to the launch decision, for the lanes with mine == true
// (agents, or frozen-policy opponents at levels 4/5).  Also draws missile_wait = randint(7, 17), which the
// reference consumes iff the launch is attempted (env_base.py:228-230); only an AC1 lane can attempt.
template <int MODE>
__device__ __forceinline__ void base_action(Lane& L, const Rng& rng, int u, bool mine, const int4 act, double& rew,
                                            bool& want_missile, int& tgt, int& new_wait) {
  want_missile = false;
  tgt = -1;
  new_wait = 0;
  if (mine) {
    set_heading(L, pymod(L.hdg + (double)((act.x - 6) * 15), 360.0));
    set_speed(L, u, 100.0 + ((max_speed(u) - 100.0) / 8.0) * (double)act.y);
    if (act.z != 0 && L.crem > 0) {
      fire_cannon(L, u);
      if (MODE == 1 && u < 2 && L.crem < 90) rew -= 0.1;
    }
    want_missile = is_ac1(u) && act.w != 0 && L.ota != 0 && L.mrem > 0 && !L.hasm && L.mwait == 0;
    tgt = L.ota - 1;
  }
  const int draws = __popc(quad_ballot(want_missile));   // 0 or 1: lanes 0/1 or lanes 2/3 act, one AC1 among them
  if (want_missile) new_wait = randint_from(7, 17, g_random_at(rng, L.dg));
  L.dg += draws;
}

// Rafale.fire_missile (ac1.py:72-79) for every lane that wants to launch; rocket ids in shooter id order
__device__ __forceinline__ void launch_phase(Lane& L, int u, const PreTick& p, bool want_missile, int tgt) {
  const int tq = tgt < 0 ? 0 : tgt;
  const bool launched = try_launch(L, want_missile, HH_P(lat4, tq), HH_P(lon4, tq), tq);
  const int launch_m = quad_ballot(launched);
  if (launched) L.rid = L.next_id + __popc(launch_m & ((1 << u) - 1));
  L.next_id += __popc(launch_m);
}

// CmanoSimulator.do_tick (cmano_simulator.py:138-157) + _get_rewards/_combat_rewards (env_hetero.py:188-225,
// env_base.py:240-310).  `rew` enters with the action-phase penalties and leaves with the step reward of
// the calling agent lane.  Returns done (env_base.py:89-90).
// SHORT: moves through geo::direct_tick (the short-arc solve of the v4 kernel) -- the level-4/5 product path; the quad
// level 1-3 kernel keeps the full Karney solve as an independent cross-check (A/B test in tests/test_gpu_parity.py).
template <int MODE, bool SHORT>
__device__ __forceinline__ bool tick_and_rewards(Lane& L, const Rng& rng, const Geom& g, const Params& P, int u,
                                                 const PreTick& p, double opp_focus, bool present, double& rew) {
  const int alive0 = p.alive_m;                  // snapshot: nothing dies in the action phase
  const bool upd = L.alive;
  const bool rocket0 = L.ralive;                 // snapshot includes rockets launched this step
  const double max_deg = is_ac1(u) ? 5.0 : 3.5, max_kn = is_ac1(u) ? 35.0 : 28.0;
  if (upd) {                                     // ac1.py:82-99
    if (L.hdg != L.nhdg) {
      const double delta = signed_heading_diff(L.hdg, L.nhdg);
      L.hdg = fabs(delta) <= max_deg ? L.nhdg : pymod(L.hdg + (delta >= 0.0 ? max_deg : -max_deg), 360.0);
    }
    if (L.spd != L.nspd) {
      const double delta = L.nspd - L.spd;
      L.spd = fabs(delta) <= max_kn ? L.nspd : L.spd + (delta >= 0.0 ? max_kn : -max_kn);
    }
  }
  const bool firing = upd && L.burst > 0;        // ac1.py:101-104
  if (firing) {
    L.burst -= 1;
    L.crem = L.crem > 0 ? L.crem - 1 : 0;
  }
  // every unit's move (Unit.update, cmano_simulator.py:65-72) depends only on itself
  phase_sync();
  double nlat = L.lat, nlon = L.lon;
  if (upd && L.spd > 0.0) {
    const double2 q = (SHORT && P.short_moves) ? geo::direct_tick(L.lat, L.lon, L.hdg, L.spd * kKnotsToMs * 1.0)
                            : geo::direct(L.lat, L.lon, L.hdg, L.spd * kKnotsToMs * 1.0);
    nlat = q.x;
    nlon = q.y;
  }
  // cannon geometry: shooter u sees lower ids at their NEW position, higher ids at the OLD one,
  // itself at its old position with its new heading (ac1.py:105-115, A.3 of SURVEY.md)
  phase_sync();
  int in_range = 0;
  {
    const double range = is_ac1(u) ? 2.0 : 4.5, half_w = (is_ac1(u) ? 10.0 : 7.0) / 2.0;
#pragma unroll 1
    for (int j = 0; j < 4; ++j) {
      const double jl_new = qshfl(nlat, j), jo_new = qshfl(nlon, j);
      const double jl_old = qshfl(L.lat, j), jo_old = qshfl(L.lon, j);   // L.lat/lon are still pre-tick here
      const double jl = j < u ? jl_new : jl_old;
      const double jo = j < u ? jo_new : jo_old;
      const bool group_ok = P.friendly_kill || ((u < 2) != (j < 2));
      if (firing && j != u && ((alive0 >> j) & 1) && group_ok)
        if (unit_in_cannon_range(L.lat, L.lon, L.hdg, jl, jo, range, half_w)) in_range |= 1 << j;
    }
  }
  // kill resolution in (shooter, target) id order; C-stream draws only for live in-range targets
  int alive_m = alive0;
  unsigned killer_pack = 0;  // 4 bits per victim: killer id (0 = none)
  int by_rocket_m = 0;
  {
    const unsigned inr_pack = (unsigned)qshfl(in_range, 0) | ((unsigned)qshfl(in_range, 1) << 4) |
                              ((unsigned)qshfl(in_range, 2) << 8) | ((unsigned)qshfl(in_range, 3) << 12);
    if (inr_pack != 0) {
#pragma unroll 1
      for (int k = 0; k < 4; ++k) {
        const unsigned row = (inr_pack >> (4 * k)) & 0xFu;
        if (row == 0) continue;
        const double p_hit = (k & 1) ? 0.9 / (3.0 / 1.0) : 0.75 / (5.0 / 1.0);
#pragma unroll 1
        for (int j = 0; j < 4; ++j) {
          if (((row >> j) & 1) && ((alive_m >> j) & 1)) {
            if (c_random_at(rng, L.dc++) < p_hit) {
              alive_m &= ~(1 << j);
              killer_pack |= (unsigned)(k + 1) << (4 * j);
            }
          }
        }
      }
    }
  }
  // missile heading noise (ac1.py:117-128): G draws in shooter id order
  {
    const bool noise = upd && is_ac1(u) && L.hasm && rocket0;
    if (upd && is_ac1(u) && L.hasm && !rocket0) L.hasm = false;
    const int noise_m = quad_ballot(noise);
    if (noise) {
      const double f = uniform_from(0.95, 1.05, g_random_at(rng, L.dg + (u == 2 ? (noise_m & 1) : 0)));
      L.rnhdg = clip(__dmul_rn(L.rhdg, f), 0.0, 359.0);
    }
    L.dg += __popc(noise_m);
  }
  if (upd) {
    L.lat = nlat;
    L.lon = nlon;
  }
  // rockets (rocket_unit.py:37-73), after every aircraft, in launch (id) order
  phase_sync();
  {
    const int t = L.rtgt > 0 ? L.rtgt - 1 : 0;
    const double t_lat = qshfl(L.lat, t), t_lon = qshfl(L.lon, t);     // targets have already moved
    const double f_lat = qshfl(L.lat, 1), f_lon = qshfl(L.lon, 1);     // "friendly" is always id 2 (A.6.5)
    bool hit_t = false, hit_f = false;
    if (rocket0) {
      hit_t = within_1km(L.rlat, L.rlon, t_lat, t_lon);
      if (P.friendly_kill && ((alive_m >> 1) & 1)) hit_f = within_1km(L.rlat, L.rlon, f_lat, f_lon);
    }
    const int live_m = quad_ballot(rocket0), hit_t_m = quad_ballot(hit_t), hit_f_m = quad_ballot(hit_f);
    const int rid0 = qshfl(L.rid, 0), rid2 = qshfl(L.rid, 2);
    const int tg0 = qshfl(t, 0), tg2 = qshfl(t, 2);
    int exploded_m = 0;
    if (live_m != 0) {
      const int first = ((live_m & 5) == 5) ? (rid0 < rid2 ? 0 : 2) : ((live_m & 1) ? 0 : 2);
#pragma unroll 1
      for (int n = 0; n < 2; ++n) {
        const int s = n == 0 ? first : 2 - first;
        if (!((live_m >> s) & 1)) continue;
        const int tt = s == 0 ? tg0 : tg2;
        if (((hit_t_m >> s) & 1) && ((alive_m >> tt) & 1)) {
          alive_m &= ~(1 << tt);
          killer_pack |= (unsigned)(s + 1) << (4 * tt);
          by_rocket_m |= 1 << tt;
          exploded_m |= 1 << s;
        } else if (((hit_f_m >> s) & 1) && ((alive_m >> 1) & 1)) {
          alive_m &= ~2;
          killer_pack |= (unsigned)(s + 1) << 4;
          by_rocket_m |= 2;
          exploded_m |= 1 << s;
        }
      }
    }
    if (rocket0) {
      if ((exploded_m >> u) & 1) {
        L.ralive = false;
      } else if (L.rage > 10) {
        L.ralive = false;
      } else {
        if (L.rhdg != L.rnhdg) {
          const double delta = signed_heading_diff(L.rhdg, L.rnhdg);
          L.rhdg = fabs(delta) <= 10.0 ? L.rnhdg : L.rhdg + (delta >= 0.0 ? 10.0 : -10.0);
        }
        const double2 q = (SHORT && P.short_moves) ? geo::direct_tick(L.rlat, L.rlon, L.rhdg, rocket_speed(L.rage) * kKnotsToMs * 1.0)
                                : geo::direct(L.rlat, L.rlon, L.rhdg, rocket_speed(L.rage) * kKnotsToMs * 1.0);
        L.rlat = q.x;
        L.rlon = q.y;
        L.rage += 1;
      }
    }
  }
  L.alive = (alive_m >> u) & 1;

  // ---- rewards, evaluated identically in the four lanes
  phase_sync();
  {
    const double s = P.rew_scale;
    const int oob_m = quad_ballot(L.alive && !in_boundary(g, L.lat, L.lon));
    if ((oob_m >> u) & 1) L.alive = false;
    alive_m &= ~oob_m;
    double rews0 = 0.0, rews1 = 0.0;
    int destroyed_m = oob_m & 3;
    if (oob_m & 1) rews0 += -5.0 * s;
    if (oob_m & 2) rews1 += -5.0 * s;
    L.alive_ag -= __popc(oob_m & 3);
    L.alive_op -= __popc(oob_m & 12);
    const double of0 = qshfl(opp_focus, 0), of1 = qshfl(opp_focus, 1);
    const int cr0 = qshfl(L.crem, 0), cr1 = qshfl(L.crem, 1), cm0 = qshfl(L.cmax, 0), cm1 = qshfl(L.cmax, 1);
    const int mr0 = qshfl(L.mrem, 0), rm0 = qshfl(L.rmax, 0);
    if (killer_pack != 0) {   // quad-uniform but warp-divergent: no shuffles inside
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
                r = (1.0 + 0.5 * ((double)mr0 / (double)rm0)) * s;       // only agent 1 carries missiles
              else
                r = ((0.5 + 0.5 * ((double)(k == 1 ? cr0 : cr1) / (double)(k == 1 ? cm0 : cm1))) +
                     (0.5 + 0.5 * (k == 1 ? of0 : of1))) * s;
              if (k == 1) add0 = r; else add1 = r;
            }
            L.alive_op -= 1;
          } else {
            if (k == 1) add0 = -2.0 * s; else add1 = -2.0 * s;
            if (P.friendly_punish) {
              if (j == 0) add0 += -2.0 * s; else add1 += -2.0 * s;
              destroyed_m |= 1 << j;
            }
            L.alive_ag -= 1;
          }
        } else {
          if (j < 2) {
            if (j == 0) add0 = -2.0 * s; else add1 = -2.0 * s;
            destroyed_m |= 1 << j;
            L.alive_ag -= 1;
          } else {
            L.alive_op -= 1;
          }
        }
        rews0 += add0;
        rews1 += add1;
      }
    }
    if (MODE == 1 && P.esc_dist_rew) {           // env_hetero.py:198-214 (per agent lane)
      const double l2 = qshfl(L.lat, 2), o2 = qshfl(L.lon, 2), l3 = qshfl(L.lat, 3), o3 = qshfl(L.lon, 3);
      double mine = 0.0;
      if (u < 2 && L.alive) {
        const double d2 = dist_raw(L.lat, L.lon, l2, o2), d3 = dist_raw(L.lat, L.lon, l3, o3);
        const bool has2 = (alive_m >> 2) & 1, has3 = (alive_m >> 3) & 1;
        const bool swap = has2 && has3 && (g.inv_diag * d3) < (g.inv_diag * d2);
        const double first = has2 ? (swap ? d3 : d2) : d3, second = swap ? d2 : d3;
        const int n = (int)has2 + (int)has3;
        for (int j = 1; j <= n; ++j) {
          const double od = j == 1 ? first : second;
          if (od < 0.06) {
            mine += -0.02 / j;
            if (L.spd < 200.0) mine += -0.02 / j;
          } else if (od > 0.13) {
            mine += 0.02 / j;
            if (L.spd > 500.0) mine += 0.02 / j;
          }
        }
      }
      rews0 += qshfl(mine, 0);
      rews1 += qshfl(mine, 1);
    }
    if (u < 2 && present && (L.alive || ((destroyed_m >> u) & 1))) {
      const double own = u == 0 ? rews0 : rews1, other = u == 0 ? rews1 : rews0;
      rew += (P.glob_frac > 0.0 && MODE == 0) ? own + P.glob_frac * other : own;
    }
  }
  return L.alive_ag <= 0 || L.alive_op <= 0 || L.steps >= P.horizon;   // env_base.py:89-90
}

// terminal bookkeeping shared by the fused and the finishing kernel
template <int MODE>
__device__ __forceinline__ void finish_step(Lane& L, const Rng& rng, const Geom& g, const Params& P, const StatePtrs& S,
                                            int a, int u, bool valid, bool done, double rew, float* obs1, float* obs2,
                                            float* rew_out, uint8_t* done_out, float* s1, float* s2, int arena0,
                                            int n_valid) {
  const double r1 = qshfl(rew, 1);
  if (valid && u == 0) {
    if (rew_out) reinterpret_cast<float2*>(rew_out)[a] = make_float2((float)rew, (float)r1);
    if (done_out) done_out[a] = done ? 1 : 0;
  }
  if (done && P.autoreset) reset_lane(L, rng, P, u);
  phase_sync();
  write_agent_obs<MODE>(L, g, u, obs1, obs2, s1, s2, arena0, n_valid);
  store_lane(S, a, u, L, valid);
}

#define HH_KERNEL_PROLOGUE                                                                        \
  const int u = threadIdx.x & 3;                                                                  \
  const int arena0 = blockIdx.x * kArenasPerCta;                                                  \
  const int a_raw = arena0 + (threadIdx.x >> 2);                                                  \
  const bool valid = a_raw < P.n_arenas;                                                          \
  const int a = valid ? a_raw : P.n_arenas - 1; /* tail lanes shadow the last arena, never store */ \
  const int n_valid = min(kArenasPerCta, P.n_arenas - arena0);                                    \
  const Geom g = make_geom(P.map_size);                                                           \
  const Rng rng{P.seed_lo, P.seed_hi, P.arena_base + (uint32_t)a};                                \
  Lane L;                                                                                         \
  load_lane(S, a, u, L);

// ------------------------------------------------------------------------------------------
// step, levels 1-3 (scripted opponents): one launch
// ------------------------------------------------------------------------------------------
template <int LEVEL, int MODE>
__global__ void __launch_bounds__(kThreads)
step_kernel(StatePtrs S, Params P, const int32_t* __restrict__ actions, float* __restrict__ obs1,
            float* __restrict__ obs2, float* __restrict__ rew_out, uint8_t* __restrict__ done_out) {
  __shared__ __align__(16) float s1[kArenasPerCta * ObsDims<MODE>::D1];
  __shared__ __align__(16) float s2[kArenasPerCta * ObsDims<MODE>::D2];
  HH_KERNEL_PROLOGUE
  int4 act = make_int4(6, 0, 0, 0);
  if (u < 2) act = reinterpret_cast<const int4*>(actions)[(size_t)a * 2 + u];

  // LowLevelEnv._take_action, env_hetero.py:105-186
  L.steps += 1;
  const bool present = L.alive;       // agents: has an entry in the reward dict (env_hetero.py:168)
  double rew = 0.0;
  const PreTick p = gather_pre(L);
  int near_t, rel_sign;
  double near_dn, rel_focus;
  pre_tick_relations<LEVEL == 3>(L, u, g, p, near_t, near_dn, rel_focus, rel_sign);
  const double opp_focus = u < 2 ? focus_norm_from_deg(rel_focus) : 0.0;  // 0 when no opp_stats entry
  phase_sync();

  bool want_missile;
  int tgt, new_wait;
  base_action<MODE>(L, rng, u, u < 2 && L.alive, act, rew, want_missile, tgt, new_wait);

  // scripted opponents, id order, shared escape state (env_hetero.py:118-158)
#pragma unroll 1
  for (int k = 2; k < 4; ++k) {
    const bool k_alive = (p.alive_m >> k) & 1;
    const bool k_hasm = qshfl((int)L.hasm, k) != 0;
    const int k_mwait = qshfl(L.mwait, k);
    const int k_near = qshfl(near_t, k);
    const double k_dn = qshfl(near_dn, k);
    const double k_focus = qshfl(rel_focus, k);
    const int k_sign = qshfl(rel_sign, k);
    const double k_lat = qshfl(L.lat, k), k_lon = qshfl(L.lon, k), k_hdg = qshfl(L.hdg, k);
    const OppDecision d = scripted_opponent<LEVEL>(L, rng, g, k, k_alive, k_hasm, k_mwait, k_lat, k_lon, k_hdg,
                                                   k_near, k_dn, k_focus, k_sign);
    if (u == k && k_alive) {
      if (d.set_hs) {
        set_heading(L, d.heading);
        set_speed(L, u, d.speed);
      }
      if (d.fire) fire_cannon(L, u);
      want_missile = d.want_missile;
      tgt = d.tgt;
    }
    phase_sync();
  }
  launch_phase(L, u, p, want_missile, tgt);
  phase_sync();
  if (want_missile) {
    if (u < 2) {
      L.mwait = new_wait;
      if (MODE == 1 && L.mrem < 3) rew -= 0.1;
    } else {
      L.mwait = LEVEL == 3 ? 10 : 5;   // env_hetero.py:123,136,158 (never decremented: SURVEY A.6.1)
    }
  }
  if (u < 2 && L.alive && L.mwait > 0 && !L.hasm) L.mwait -= 1;  // env_base.py:235-236

  const bool done = tick_and_rewards<MODE, false>(L, rng, g, P, u, p, opp_focus, present, rew);
  finish_step<MODE>(L, rng, g, P, S, a, u, valid, done, rew, obs1, obs2, rew_out, done_out, s1, s2, arena0, n_valid);
}

// Same step, "v3" work distribution (hh_cta.cuh): 32 arenas per 128-thread CTA staged in shared memory, each
// phase executed by exactly the threads that have work in it.
template <int LEVEL, int MODE>
__global__ void __launch_bounds__(cta::kThreads)
step_kernel_cta(StatePtrs S, Params P, const int32_t* __restrict__ actions, float* __restrict__ obs1,
                float* __restrict__ obs2, float* __restrict__ rew_out, uint8_t* __restrict__ done_out) {
  __shared__ __align__(16) cta::Smem sm;
  cta::step_body<LEVEL, MODE>(sm, S, P, actions, obs1, obs2, rew_out, done_out);
}

// Same step, "v4" schedule (hh_v4.cuh, default): stages of concurrent roles on a 256-thread CTA -- random draws
// prepared ahead, rocket pipeline beside the aircraft pipeline, one arena-serial resolution stage, pair-parallel
// observation features, short-arc direct solve.  2 CTAs per SM (<= 128 registers) so that 8 192 arenas = 256 CTAs
// are resident at once.
#ifndef HH_V4_MIN_CTAS
#define HH_V4_MIN_CTAS (v4::kThreads <= 256 ? 2 : 1)   // <= 128 registers per thread either way
#endif
constexpr int kV4MinCtas = HH_V4_MIN_CTAS;
template <int LEVEL, int MODE>
__global__ void __launch_bounds__(v4::kThreads, kV4MinCtas)
step_kernel_v4(StatePtrs S, Params P, const int32_t* __restrict__ actions, float* __restrict__ obs1,
               float* __restrict__ obs2, float* __restrict__ rew_out, uint8_t* __restrict__ done_out, int block0) {
  // block0 > 0: a launch over a sub-range of the arenas (pipelined host mode; P.n_arenas is then the range's end)
  extern __shared__ __align__(16) unsigned char v4_smem[];
  v4::step_body<LEVEL, MODE>(*reinterpret_cast<v4::Smem*>(v4_smem), S, P, actions, obs1, obs2, rew_out, done_out,
                             blockIdx.x + block0);
}
// two sub-blocks per CTA (v4::step_body<.., DUAL = true>): blocks 2 b and 2 b + 1 of the plain launch
template <int LEVEL, int MODE>
__global__ void __launch_bounds__(2 * v4::kThreads, 1)
step_kernel_v4_dual(StatePtrs S, Params P, const int32_t* __restrict__ actions, float* __restrict__ obs1,
                    float* __restrict__ obs2, float* __restrict__ rew_out, uint8_t* __restrict__ done_out, int block0,
                    float* cen1, float* cen2, int cen_ld) {
  extern __shared__ __align__(16) unsigned char v4_smem[];
  const int sub = threadIdx.x / v4::kThreads;
  const int block = 2 * (int)blockIdx.x + sub + block0;
  if (block * v4::kArenas >= P.n_arenas) return;      // sub-block uniform (an odd number of blocks: the last CTA runs one)
  v4::step_body<LEVEL, MODE, true>(*reinterpret_cast<v4::Smem*>(v4_smem + (size_t)sub * ((sizeof(v4::Smem) + 15) / 16 * 16)), S, P,
                                   actions, obs1, obs2, rew_out, done_out, block, cen1, cen2, cen_ld);
}
template <int LEVEL, int MODE>
static cudaError_t v4_opt_in_smem() {
  cudaError_t ce = cudaFuncSetAttribute(step_kernel_v4<LEVEL, MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                        (int)sizeof(v4::Smem));
  if (ce == cudaSuccess)
    ce = cudaFuncSetAttribute(step_kernel_v4_dual<LEVEL, MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                              (int)(2 * ((sizeof(v4::Smem) + 15) / 16 * 16)));
  return ce;
}

// ------------------------------------------------------------------------------------------
// step, levels 4-5 (frozen-policy opponents, env_base.py:349-398): the opponent's observation is
// needed mid-step -- after the agents' fire decisions, before the tick -- so the step is split:
//   step_begin_kernel : agents' _take_base_action, then lowlevel_state(opp_mode) of both opponents
//   (batched opponent networks, outside this library)
//   step_finish_kernel: opponents' _take_base_action with the per-head argmax, tick, rewards, obs
// ------------------------------------------------------------------------------------------
constexpr int kOppD3 = OBS_ESC_AC1, kOppD4 = OBS_ESC_AC2;   // row strides of the opponent-obs buffers (max over modes)

template <int MODE>
__global__ void __launch_bounds__(kThreads)
step_begin_kernel(StatePtrs S, Params P, const int32_t* __restrict__ actions, float* __restrict__ opp_obs3,
                  float* __restrict__ opp_obs4, uint8_t* __restrict__ pset_out, float* __restrict__ rew_pre) {
  __shared__ __align__(16) float s3[kArenasPerCta * kOppD3];
  __shared__ __align__(16) float s4[kArenasPerCta * kOppD4];
  HH_KERNEL_PROLOGUE
  int4 act = make_int4(6, 0, 0, 0);
  if (u < 2) act = reinterpret_cast<const int4*>(actions)[(size_t)a * 2 + u];
  L.steps += 1;
  double rew = 0.0;
  const PreTick p = gather_pre(L);
  bool want_missile;
  int tgt, new_wait;
  base_action<MODE>(L, rng, u, u < 2 && L.alive, act, rew, want_missile, tgt, new_wait);
  launch_phase(L, u, p, want_missile, tgt);
  if (want_missile) {
    L.mwait = new_wait;
    if (MODE == 1 && L.mrem < 3) rew -= 0.1;
  }
  if (u < 2 && L.alive && L.mwait > 0 && !L.hasm) L.mwait -= 1;
  {
    const double r1 = qshfl(rew, 1);
    if (valid && u == 0 && rew_pre) reinterpret_cast<float2*>(rew_pre)[a] = make_float2((float)rew, (float)r1);
    if (valid && u == 0 && pset_out) pset_out[a] = (uint8_t)L.pset;   // k of env_hetero.py:57 (0 below level 5)
  }
  // _policy_actions -> lowlevel_state(policy_type, opp id): the opponents already see this step's bursts / launches
  const World W = gather_world(L, u);
  const int al = threadIdx.x >> 2;
  if (u >= 2) {
    float* row = u == 2 ? s3 + al * kOppD3 : s4 + al * kOppD4;
    const int len = obs_len(u, L.opp_mode), stride = u == 2 ? kOppD3 : kOppD4;
    // only an opponent that still exists queries its policy (env_hetero.py:160-182), and only that query rewrites its
    // opp_to_attack (lowlevel_state, env_hetero.py:89-96): a destroyed opponent keeps its last value
    const int o_new = unit_observation(L, W, g, u, L.opp_mode, row);
    if (L.alive) L.ota = o_new;
    for (int k = len; k < stride; ++k) row[k] = 0.0f;
  }
  cta_sync();
  if (opp_obs3) flush_rows(opp_obs3 + (size_t)arena0 * kOppD3, s3, n_valid * kOppD3);
  if (opp_obs4) flush_rows(opp_obs4 + (size_t)arena0 * kOppD4, s4, n_valid * kOppD4);
  store_lane(S, a, u, L, valid);
}

template <int MODE>
__global__ void __launch_bounds__(kThreads)
step_finish_kernel(StatePtrs S, Params P, const int32_t* __restrict__ opp_actions, const float* __restrict__ rew_pre,
                   float* __restrict__ obs1, float* __restrict__ obs2, float* __restrict__ rew_out,
                   uint8_t* __restrict__ done_out) {
  __shared__ __align__(16) float s1[kArenasPerCta * ObsDims<MODE>::D1];
  __shared__ __align__(16) float s2[kArenasPerCta * ObsDims<MODE>::D2];
  HH_KERNEL_PROLOGUE
  int4 act = make_int4(6, 0, 0, 0);
  if (u >= 2) act = reinterpret_cast<const int4*>(opp_actions)[(size_t)a * 2 + (u - 2)];
  const bool present = L.alive;
  double rew = 0.0;
  if (u < 2 && rew_pre) rew = (double)rew_pre[(size_t)a * 2 + u];
  const PreTick p = gather_pre(L);
  int near_t, rel_sign;
  double near_dn, rel_focus;
  pre_tick_relations<false>(L, u, g, p, near_t, near_dn, rel_focus, rel_sign);   // positions are still pre-tick
  const double opp_focus = u < 2 ? focus_norm_from_deg(rel_focus) : 0.0;
  bool want_missile;
  int tgt, new_wait;
  base_action<MODE>(L, rng, u, u >= 2 && L.alive, act, rew, want_missile, tgt, new_wait);
  launch_phase(L, u, p, want_missile, tgt);
  if (want_missile) L.mwait = new_wait;
  if (u >= 2 && L.alive && L.mwait > 0 && !L.hasm) L.mwait -= 1;
  const bool done = tick_and_rewards<MODE, true>(L, rng, g, P, u, p, opp_focus, present, rew);
  finish_step<MODE>(L, rng, g, P, S, a, u, valid, done, rew, obs1, obs2, rew_out, done_out, s1, s2, arena0, n_valid);
}

// ------------------------------------------------------------------------------------------
// reset (masked)
// ------------------------------------------------------------------------------------------
template <int MODE>
__global__ void __launch_bounds__(kThreads)
reset_kernel(StatePtrs S, Params P, const uint8_t* __restrict__ mask, int first_time, float* __restrict__ obs1,
             float* __restrict__ obs2) {
  __shared__ __align__(16) float s1[kArenasPerCta * ObsDims<MODE>::D1];
  __shared__ __align__(16) float s2[kArenasPerCta * ObsDims<MODE>::D2];
  const int u = threadIdx.x & 3;
  const int arena0 = blockIdx.x * kArenasPerCta;
  const int a_raw = arena0 + (threadIdx.x >> 2);
  const bool valid = a_raw < P.n_arenas;
  const int a = valid ? a_raw : P.n_arenas - 1;
  const int n_valid = min(kArenasPerCta, P.n_arenas - arena0);
  const Geom g = make_geom(P.map_size);
  const Rng rng{P.seed_lo, P.seed_hi, P.arena_base + (uint32_t)a};
  Lane L;
  if (first_time) {
    L.dg = 0;
    L.dc = 0;
    L.err = 0;
    reset_lane(L, rng, P, u);
  } else {
    load_lane(S, a, u, L);
    if (!mask || mask[a]) reset_lane(L, rng, P, u);
  }
  write_agent_obs<MODE>(L, g, u, obs1, obs2, s1, s2, arena0, n_valid);
  store_lane(S, a, u, L, valid);
}

// ------------------------------------------------------------------------------------------
// generalised advantage estimation over a rollout fragment [T][N][2] (RLlib postprocessing,
// compute_advantages with use_gae=True): one thread per (arena, agent), backward scan over t;
// consecutive threads touch consecutive addresses in every [t] slab.
// ------------------------------------------------------------------------------------------
__global__ void gae_kernel(int T, int n_pairs, int n_agents, const float* __restrict__ rew, const float* __restrict__ vf,
                           const float* __restrict__ last_vf, const uint8_t* __restrict__ done, float gamma,
                           float lam, float* __restrict__ adv, float* __restrict__ vtarg) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;   // i = arena * n_agents + agent
  if (i >= n_pairs) return;
  const int arena = i / n_agents, n_arenas = n_pairs / n_agents;
  float next_v = last_vf[i], gae = 0.0f;
  for (int t = T - 1; t >= 0; --t) {
    const size_t k = (size_t)t * n_pairs + i;
    const float nonterminal = done[(size_t)t * n_arenas + arena] ? 0.0f : 1.0f;   // terminated == truncated here
    const float delta = rew[k] + gamma * next_v * nonterminal - vf[k];
    gae = delta + gamma * lam * nonterminal * gae;
    adv[k] = gae;
    vtarg[k] = gae + vf[k];
    next_v = vf[k];
  }
}

// ------------------------------------------------------------------------------------------
// sampler glue: MultiCategorical sampling of both policies' actions and the central-critic observation layout
// ------------------------------------------------------------------------------------------
// One thread per (arena, agent): inverse-CDF sample of every MultiDiscrete head from its logits (RLlib's
// TorchMultiCategorical.sample / logp), Philox stream 2 keyed by (seed, global arena id), per-arena counter.
__global__ void sample_actions_kernel(int n, const float* __restrict__ logits1, const float* __restrict__ logits2,
                                      uint32_t seed_lo, uint32_t seed_hi, uint32_t arena_base, uint32_t* __restrict__ ctr,
                                      int explore, int32_t* __restrict__ actions, float* __restrict__ logp) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= 2 * n) return;
  const int a = i >> 1, agent = i & 1;
  const float* lg = agent == 0 ? logits1 + (size_t)a * 26 : logits2 + (size_t)a * 24;
  const int n_heads = agent == 0 ? 4 : 3;
  const uint32_t c = ctr[i];
  float lp = 0.0f;
  int o = 0;
  int32_t out[4] = {0, 0, 0, 0};
  for (int h = 0; h < n_heads; ++h) {
    const int w = h == 0 ? 13 : (h == 1 ? 9 : 2);
    float mx = lg[o];
    for (int k = 1; k < w; ++k) mx = fmaxf(mx, lg[o + k]);
    float e[13], sum = 0.0f;
    for (int k = 0; k < w; ++k) {
      e[k] = expf(lg[o + k] - mx);
      sum += e[k];
    }
    int pick = 0;
    if (explore) {
      const float u = (float)philox_u53(seed_lo, seed_hi, c, (uint32_t)h, arena_base + (uint32_t)a, 2u + (uint32_t)agent) * sum;
      float acc = 0.0f;
      pick = w - 1;
      for (int k = 0; k < w; ++k) {
        acc += e[k];
        if (u < acc) { pick = k; break; }
      }
    } else {
      for (int k = 1; k < w; ++k)
        if (lg[o + k] > lg[o + pick]) pick = k;
    }
    out[h] = pick;
    lp += lg[o + pick] - mx - logf(sum);
    o += w;
  }
  ctr[i] = c + 1;
  reinterpret_cast<int4*>(actions)[i] = make_int4(out[0], out[1], out[2], out[3]);
  logp[i] = lp;
}

// ------------------------------------------------------------------------------------------
// The learner's MultiCategorical terms of one policy (RLlib TorchMultiCategorical.logp / entropy / kl, summed over the heads):
// one thread per row.  Forward: logp of the taken actions, entropy, KL(old || new).  Backward: d / d logits from the three
// upstream gradients; per head with p = softmax(z), q = softmax(z_old):
//   d logp / d z_j = [j = a] - p_j,   d H / d z_j = -p_j (log p_j + H),   d KL / d z_j = p_j - q_j.
// Replaces ~35 element-wise torch kernels per head and direction in PPOLearner's minibatch step.
struct MultiCatShape { int n_heads, w[4], n_tot; };
__device__ __forceinline__ void head_stats(const float* z, int w, float& mx, float& lse) {
  mx = z[0];
  for (int k = 1; k < w; ++k) mx = fmaxf(mx, z[k]);
  float s = 0.0f;
  for (int k = 0; k < w; ++k) s += expf(z[k] - mx);
  lse = mx + logf(s);
}
__global__ void multicat_forward_kernel(int n, MultiCatShape sh, const float* __restrict__ logits, int ld, const float* __restrict__ old_logits,
                                        int ld_old, const int32_t* __restrict__ actions, int ld_act, float* __restrict__ logp,
                                        float* __restrict__ ent, float* __restrict__ kl) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float* z = logits + (size_t)i * ld;
  const float* zo = old_logits + (size_t)i * ld_old;
  float lp = 0.0f, en = 0.0f, k_ = 0.0f;
  int o = 0;
  for (int h = 0; h < sh.n_heads; ++h) {
    const int w = sh.w[h];
    float mx, lse, mxo, lseo;
    head_stats(z + o, w, mx, lse);
    head_stats(zo + o, w, mxo, lseo);
    const int a = actions[(size_t)i * ld_act + h];
    lp += z[o + a] - lse;
    for (int k = 0; k < w; ++k) {
      const float l = z[o + k] - lse, lo = zo[o + k] - lseo;
      en -= expf(l) * l;
      k_ += expf(lo) * (lo - l);
    }
    o += w;
  }
  logp[i] = lp;
  ent[i] = en;
  kl[i] = k_;
}
__global__ void multicat_backward_kernel(int n, MultiCatShape sh, const float* __restrict__ logits, int ld, const float* __restrict__ old_logits,
                                         int ld_old, const int32_t* __restrict__ actions, int ld_act, const float* __restrict__ g_logp,
                                         const float* __restrict__ g_ent, const float* __restrict__ g_kl, float* __restrict__ g_logits) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float* z = logits + (size_t)i * ld;
  const float* zo = old_logits + (size_t)i * ld_old;
  float* g = g_logits + (size_t)i * sh.n_tot;
  const float gl = g_logp[i], ge = g_ent[i], gk = g_kl[i];
  int o = 0;
  for (int h = 0; h < sh.n_heads; ++h) {
    const int w = sh.w[h];
    float mx, lse, mxo, lseo;
    head_stats(z + o, w, mx, lse);
    head_stats(zo + o, w, mxo, lseo);
    const int a = actions[(size_t)i * ld_act + h];
    float H = 0.0f;
    for (int k = 0; k < w; ++k) {
      const float l = z[o + k] - lse;
      H -= expf(l) * l;
    }
    for (int k = 0; k < w; ++k) {
      const float l = z[o + k] - lse, p = expf(l), q = expf(zo[o + k] - lseo);
      g[o + k] = gl * ((k == a ? 1.0f : 0.0f) - p) - ge * p * (l + H) + gk * (p - q);
    }
    o += w;
  }
}
// central_critic_observer (train_hetero.py:162-181): flat_p = [act_own | act_other | obs_own | obs_other] with the
// action columns zero at sampling time; writes the observation columns of both policies' inputs.
__global__ void pack_central_kernel(int n, int d1, int d2, const float* __restrict__ obs1, const float* __restrict__ obs2,
                                    float* __restrict__ flat1, float* __restrict__ flat2) {
  const int D = 7 + d1 + d2;
  const size_t total = (size_t)n * (d1 + d2);
  for (size_t k = (size_t)blockIdx.x * blockDim.x + threadIdx.x; k < total; k += (size_t)gridDim.x * blockDim.x) {
    const int a = (int)(k / (d1 + d2)), c = (int)(k % (d1 + d2));
    const float v = c < d1 ? obs1[(size_t)a * d1 + c] : obs2[(size_t)a * d2 + (c - d1)];
    flat1[(size_t)a * D + 7 + c] = v;                                   // [.. | obs1 | obs2]
    flat2[(size_t)a * D + 7 + (c < d1 ? d2 + c : c - d1)] = v;          // [.. | obs2 | obs1]
  }
}

// ------------------------------------------------------------------------------------------
// test access to the device geodesics (hh_debug_geodesic)
// ------------------------------------------------------------------------------------------
__global__ void geodesic_debug_kernel(int mode, int n, const double* __restrict__ in, double* __restrict__ out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const double a = in[i], b = in[n + i], c = in[2 * n + i], d = in[3 * n + i];
  double2 r;
  if (mode == 0) r = geo::direct(a, b, c, d);
  else if (mode == 1) r = geo::inverse(a, b, c, d);
  else if (mode == 2) r = geo::inverse_local(a, b, c, d);
  else if (mode == 3) {          // the v4 step's aircraft move: azimuth through the heading vector (env_base.py:428)
    const HVec hv = heading_vec(c);
    r = geo::direct_short(a, b, c, hv.c, hv.s, d);
  } else {                       // the v4 step's rocket move: azimuth through sincosd
    double sa, ca;
    geo::sincosd(geo::ang_round(geo::ang_normalize(c)), sa, ca);
    r = geo::direct_short(a, b, c, sa, ca, d);
  }
  out[i] = r.x;
  out[n + i] = r.y;
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
  bool mid_step = false;
  int step_impl = 4;      // HH_STEP_IMPL=v4|cta|quad (levels 1-3 fused step): 4 = hh_v4.cuh, 3 = hh_cta.cuh, 2 = hh_quad.cuh
  float* rew_pre = nullptr;
  uint64_t launches = 0;
  // host-variant staging
  cudaStream_t hstream = nullptr;
  void* d_slab = nullptr;
  int32_t* d_actions = nullptr;
  float *d_obs1 = nullptr, *d_obs2 = nullptr, *d_rew = nullptr;
  uint8_t *d_done = nullptr, *d_mask = nullptr;
  void* pinned = nullptr;
  void* pinned_dev = nullptr;   // device-side address of the pinned slab (zero-copy host mode)
  size_t pinned_bytes = 0;
  bool host_pending = false;    // a step enqueued by hh_step_host_begin has not been collected yet
  int host_mode = 1;            // 0: staged (H2D, launch, D2H); 1: zero-copy (kernels read / write the pinned slab);
                                // 2: pipelined (two half-batch launches, the first half's D2H under the second half's kernel)
  cudaStream_t cstream = nullptr;       // copy stream of the pipelined mode
  cudaGraphExec_t pipe_graph = nullptr; // one graph launch per pipelined host step
};

static size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

static int obs_dim(const hh_config& c, int agent) {
  if (c.agent_mode == 0) return agent == 1 ? OBS_AC1 : OBS_AC2;
  return agent == 1 ? OBS_ESC_AC1 : OBS_ESC_AC2;
}

extern "C" const char* hh_last_error(void) { return g_last_error.c_str(); }
extern "C" const char* hh_version(void) { return "hhmarl_2d_b200 0.1 (sm_100a)"; }

extern "C" int hh_gae(int32_t T, int32_t n_arenas, const float* rew_dev, const float* vf_dev, const float* last_vf_dev,
                      const uint8_t* done_dev, float gamma, float lam, float* adv_dev, float* vtarg_dev, void* stream) {
  if (T <= 0 || n_arenas <= 0 || !rew_dev || !vf_dev || !last_vf_dev || !done_dev || !adv_dev || !vtarg_dev)
    return fail(-1, "hh_gae: bad argument");
  const int n_pairs = n_arenas * 2;
  gae_kernel<<<(n_pairs + 127) / 128, 128, 0, static_cast<cudaStream_t>(stream)>>>(T, n_pairs, 2, rew_dev, vf_dev, last_vf_dev,
                                                                                 done_dev, gamma, lam, adv_dev, vtarg_dev);
  HH_CUDA(cudaGetLastError());
  return 0;
}

// The two ends of a rollout fragment in the sampler's central-critic layout (rows [7 action columns | own obs | other obs], D floats):
//   prepare  : the critic sees ZERO actions while sampling -- clear the action columns of all T x N rows of both
//              policies and seed tick 0 with the current central observation rows;
//   writeback: CustomCallback.on_postprocess_trajectory (train_hetero.py:120-160) -- the critic's action columns get the real
//              actions, scaled (a0 / 12, a1 / 8, a2, a3; train_hetero.py:143-146): flat1 = [own1 (4) | own2 (3)], flat2 = [own2 (3) | own1 (4)].
__global__ void fragment_prepare_kernel(int T, int n, int D, float* __restrict__ flat1, float* __restrict__ flat2,
                                        const float* __restrict__ cur1, const float* __restrict__ cur2) {
  const size_t n0 = (size_t)n * D, nz = (size_t)(T - 1) * n * 7;
  for (size_t k = (size_t)blockIdx.x * blockDim.x + threadIdx.x; k < n0 + nz; k += (size_t)gridDim.x * blockDim.x) {
    if (k < n0) {
      flat1[k] = cur1[k];
      flat2[k] = cur2[k];
    } else {
      const size_t j = k - n0, row = n + j / 7;
      flat1[row * D + j % 7] = 0.0f;
      flat2[row * D + j % 7] = 0.0f;
    }
  }
}
__global__ void fragment_writeback_kernel(size_t rows, int D, const int32_t* __restrict__ actions, float* __restrict__ flat1,
                                          float* __restrict__ flat2) {
  for (size_t r = (size_t)blockIdx.x * blockDim.x + threadIdx.x; r < rows; r += (size_t)gridDim.x * blockDim.x) {
    const int4 a1 = reinterpret_cast<const int4*>(actions)[2 * r], a2 = reinterpret_cast<const int4*>(actions)[2 * r + 1];
    const float o1[4] = {(float)a1.x / 12.0f, (float)a1.y / 8.0f, (float)a1.z / 1.0f, (float)a1.w / 1.0f};
    const float o2[3] = {(float)a2.x / 12.0f, (float)a2.y / 8.0f, (float)a2.z / 1.0f};
    float* f1 = flat1 + r * D;
    float* f2 = flat2 + r * D;
    f1[0] = o1[0]; f1[1] = o1[1]; f1[2] = o1[2]; f1[3] = o1[3]; f1[4] = o2[0]; f1[5] = o2[1]; f1[6] = o2[2];
    f2[0] = o2[0]; f2[1] = o2[1]; f2[2] = o2[2]; f2[3] = o1[0]; f2[4] = o1[1]; f2[5] = o1[2]; f2[6] = o1[3];
  }
}
extern "C" int hh_fragment_prepare(int32_t T, int32_t n_arenas, int32_t D, float* flat1_dev, float* flat2_dev, const float* cur1_dev,
                                   const float* cur2_dev, void* stream) {
  if (T <= 0 || n_arenas <= 0 || D < 7 || !flat1_dev || !flat2_dev || !cur1_dev || !cur2_dev)
    return fail(-1, "hh_fragment_prepare: bad argument");
  fragment_prepare_kernel<<<592, 256, 0, static_cast<cudaStream_t>(stream)>>>(T, n_arenas, D, flat1_dev, flat2_dev, cur1_dev, cur2_dev);
  HH_CUDA(cudaGetLastError());
  return 0;
}
extern "C" int hh_fragment_writeback(int32_t T, int32_t n_arenas, int32_t D, const int32_t* actions_dev, float* flat1_dev,
                                     float* flat2_dev, void* stream) {
  if (T <= 0 || n_arenas <= 0 || D < 7 || !actions_dev || !flat1_dev || !flat2_dev)
    return fail(-1, "hh_fragment_writeback: bad argument");
  fragment_writeback_kernel<<<592, 256, 0, static_cast<cudaStream_t>(stream)>>>((size_t)T * n_arenas, D, actions_dev, flat1_dev, flat2_dev);
  HH_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int hh_gae_agents(int32_t T, int32_t n_arenas, int32_t n_agents, const float* rew_dev, const float* vf_dev,
                             const float* last_vf_dev, const uint8_t* done_dev, float gamma, float lam, float* adv_dev,
                             float* vtarg_dev, void* stream) {
  if (T <= 0 || n_arenas <= 0 || n_agents <= 0 || !rew_dev || !vf_dev || !last_vf_dev || !done_dev || !adv_dev || !vtarg_dev)
    return fail(-1, "hh_gae_agents: bad argument");
  const int n_pairs = n_arenas * n_agents;
  gae_kernel<<<(n_pairs + 127) / 128, 128, 0, static_cast<cudaStream_t>(stream)>>>(T, n_pairs, n_agents, rew_dev, vf_dev,
                                                                                 last_vf_dev, done_dev, gamma, lam, adv_dev,
                                                                                 vtarg_dev);
  HH_CUDA(cudaGetLastError());
  return 0;
}

// PPO's per-policy loss (RLlib 2.4 ppo_torch_policy.loss) from the per-row terms, one thread per row:
//   ratio = exp(logp - old_logp), surr = min(adv ratio, adv clamp(ratio, 1 - c, 1 + c)), vfl = clamp((vf - vtarg)^2, 0, vf_clip),
//   loss_i = -surr + kl_coeff kl + vf_coeff vfl - ent_coeff ent;
// sums[0..3] += loss_i, kl_i, vfl_i, ent_i (the caller divides by n), deriv[4][n] = d (mean loss) / d {logp, ent, kl, vf}_i.
__global__ void ppo_loss_kernel(int n, const float* __restrict__ logp, const float* __restrict__ ent, const float* __restrict__ kl,
                                const float* __restrict__ vf, const float* __restrict__ old_logp, const float* __restrict__ adv,
                                const float* __restrict__ vtarg, const float* __restrict__ kl_coeff, float clip, float vf_clip,
                                float vf_coeff, float ent_coeff, float* __restrict__ sums, float* __restrict__ deriv) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  float v[4] = {0.0f, 0.0f, 0.0f, 0.0f};
  if (i < n) {
    const float klc = kl_coeff[0], inv_n = 1.0f / (float)n;
    const float ratio = expf(logp[i] - old_logp[i]), a = adv[i];
    const float rc = fminf(fmaxf(ratio, 1.0f - clip), 1.0f + clip);
    const float s1 = a * ratio, s2 = a * rc;
    const float surr = fminf(s1, s2);
    // d surr / d ratio: torch.min splits a tie evenly between its arguments, clamp passes the gradient on [1 - c, 1 + c]
    const float in_range = (ratio >= 1.0f - clip && ratio <= 1.0f + clip) ? 1.0f : 0.0f;
    const float ds = s1 < s2 ? a : (s1 > s2 ? a * in_range : 0.5f * a + 0.5f * a * in_range);
    const float d = vf[i] - vtarg[i], sq = d * d;
    const float vfl = fminf(fmaxf(sq, 0.0f), vf_clip);
    const float dv = (sq >= 0.0f && sq <= vf_clip) ? 2.0f * d : 0.0f;
    v[0] = -surr + klc * kl[i] + vf_coeff * vfl - ent_coeff * ent[i];
    v[1] = kl[i];
    v[2] = vfl;
    v[3] = ent[i];
    deriv[i] = -ds * ratio * inv_n;
    deriv[(size_t)n + i] = -ent_coeff * inv_n;
    deriv[2 * (size_t)n + i] = klc * inv_n;
    deriv[3 * (size_t)n + i] = vf_coeff * dv * inv_n;
  }
  __shared__ float red[4][4];
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    float x = v[k];
    for (int o = 16; o > 0; o >>= 1) x += __shfl_xor_sync(0xffffffffu, x, o);
    if ((threadIdx.x & 31) == 0) red[k][threadIdx.x >> 5] = x;
  }
  __syncthreads();
  if (threadIdx.x < 4) atomicAdd(sums + threadIdx.x, red[threadIdx.x][0] + red[threadIdx.x][1] + red[threadIdx.x][2] + red[threadIdx.x][3]);
}
extern "C" int hh_ppo_loss(int32_t n_rows, const float* logp_dev, const float* entropy_dev, const float* kl_dev, const float* vf_dev,
                           const float* old_logp_dev, const float* adv_dev, const float* vtarg_dev, const float* kl_coeff_dev,
                           float clip, float vf_clip, float vf_coeff, float ent_coeff, float* sums_dev, float* deriv_dev, void* stream) {
  if (n_rows <= 0 || !logp_dev || !entropy_dev || !kl_dev || !vf_dev || !old_logp_dev || !adv_dev || !vtarg_dev || !kl_coeff_dev ||
      !sums_dev || !deriv_dev)
    return fail(-1, "hh_ppo_loss: bad argument");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  HH_CUDA(cudaMemsetAsync(sums_dev, 0, 4 * sizeof(float), st));
  ppo_loss_kernel<<<(n_rows + 127) / 128, 128, 0, st>>>(n_rows, logp_dev, entropy_dev, kl_dev, vf_dev, old_logp_dev, adv_dev, vtarg_dev,
                                                        kl_coeff_dev, clip, vf_clip, vf_coeff, ent_coeff, sums_dev, deriv_dev);
  HH_CUDA(cudaGetLastError());
  return 0;
}

static int multicat_shape(int32_t n_heads, const int32_t* widths, MultiCatShape& sh) {
  if (n_heads < 1 || n_heads > 4 || !widths) return -1;
  sh.n_heads = n_heads;
  sh.n_tot = 0;
  for (int h = 0; h < 4; ++h) {
    sh.w[h] = h < n_heads ? widths[h] : 0;
    if (h < n_heads && (widths[h] < 1 || widths[h] > 64)) return -1;
    sh.n_tot += sh.w[h];
  }
  return 0;
}
extern "C" int hh_multicat_forward(int32_t n_rows, int32_t n_heads, const int32_t* widths, const float* logits_dev, int32_t ld,
                                   const float* old_logits_dev, int32_t ld_old, const int32_t* actions_dev, int32_t ld_act,
                                   float* logp_dev, float* entropy_dev, float* kl_dev, void* stream) {
  MultiCatShape sh;
  if (n_rows <= 0 || multicat_shape(n_heads, widths, sh) || !logits_dev || !old_logits_dev || !actions_dev || !logp_dev || !entropy_dev ||
      !kl_dev || ld < sh.n_tot || ld_old < sh.n_tot || ld_act < n_heads)
    return fail(-1, "hh_multicat_forward: bad argument");
  multicat_forward_kernel<<<(n_rows + 127) / 128, 128, 0, static_cast<cudaStream_t>(stream)>>>(
      n_rows, sh, logits_dev, ld, old_logits_dev, ld_old, actions_dev, ld_act, logp_dev, entropy_dev, kl_dev);
  HH_CUDA(cudaGetLastError());
  return 0;
}
extern "C" int hh_multicat_backward(int32_t n_rows, int32_t n_heads, const int32_t* widths, const float* logits_dev, int32_t ld,
                                    const float* old_logits_dev, int32_t ld_old, const int32_t* actions_dev, int32_t ld_act,
                                    const float* g_logp_dev, const float* g_entropy_dev, const float* g_kl_dev, float* g_logits_dev,
                                    void* stream) {
  MultiCatShape sh;
  if (n_rows <= 0 || multicat_shape(n_heads, widths, sh) || !logits_dev || !old_logits_dev || !actions_dev || !g_logp_dev ||
      !g_entropy_dev || !g_kl_dev || !g_logits_dev || ld < sh.n_tot || ld_old < sh.n_tot || ld_act < n_heads)
    return fail(-1, "hh_multicat_backward: bad argument");
  multicat_backward_kernel<<<(n_rows + 127) / 128, 128, 0, static_cast<cudaStream_t>(stream)>>>(
      n_rows, sh, logits_dev, ld, old_logits_dev, ld_old, actions_dev, ld_act, g_logp_dev, g_entropy_dev, g_kl_dev, g_logits_dev);
  HH_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int hh_sample_actions(int32_t n_arenas, const float* logits1_dev, const float* logits2_dev, uint64_t seed,
                                 uint64_t arena_base, uint32_t* counters_dev, int32_t explore, int32_t* actions_dev,
                                 float* logp_dev, void* stream) {
  if (n_arenas <= 0 || !logits1_dev || !logits2_dev || !counters_dev || !actions_dev || !logp_dev)
    return fail(-1, "hh_sample_actions: bad argument");
  sample_actions_kernel<<<(2 * n_arenas + 127) / 128, 128, 0, static_cast<cudaStream_t>(stream)>>>(
      n_arenas, logits1_dev, logits2_dev, (uint32_t)seed, (uint32_t)(seed >> 32), (uint32_t)arena_base, counters_dev, explore,
      actions_dev, logp_dev);
  HH_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int hh_pack_central(int32_t n_arenas, int32_t d1, int32_t d2, const float* obs1_dev, const float* obs2_dev,
                               float* flat1_dev, float* flat2_dev, void* stream) {
  if (n_arenas <= 0 || d1 <= 0 || d2 <= 0 || !obs1_dev || !obs2_dev || !flat1_dev || !flat2_dev)
    return fail(-1, "hh_pack_central: bad argument");
  pack_central_kernel<<<592, 256, 0, static_cast<cudaStream_t>(stream)>>>(n_arenas, d1, d2, obs1_dev, obs2_dev, flat1_dev,
                                                                          flat2_dev);
  HH_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int hh_debug_geodesic(int32_t mode, int32_t n, const double* in_host, double* out_host) {
  if (mode < 0 || mode > 4 || n <= 0 || !in_host || !out_host) return fail(-1, "hh_debug_geodesic: bad argument");
  double *d_in = nullptr, *d_out = nullptr;
  HH_CUDA(cudaMalloc(&d_in, sizeof(double) * 4 * n));
  HH_CUDA(cudaMalloc(&d_out, sizeof(double) * 2 * n));
  HH_CUDA(cudaMemcpy(d_in, in_host, sizeof(double) * 4 * n, cudaMemcpyHostToDevice));
  geodesic_debug_kernel<<<(n + 127) / 128, 128>>>(mode, n, d_in, d_out);
  HH_CUDA(cudaGetLastError());
  HH_CUDA(cudaMemcpy(out_host, d_out, sizeof(double) * 2 * n, cudaMemcpyDeviceToHost));
  cudaFree(d_in);
  cudaFree(d_out);
  return 0;
}

#ifdef HH_V4_PROFILE
// profiling builds only (profiles/stage_clocks.py): per-CTA clock64() at the stage boundaries of the last v4 step
extern "C" int hh_debug_v4_profile(long long* out_host, int32_t n_ctas) {
  if (!out_host || n_ctas <= 0 || n_ctas > 4096) return fail(-1, "hh_debug_v4_profile: bad argument");
  HH_CUDA(cudaDeviceSynchronize());
  HH_CUDA(cudaMemcpyFromSymbol(out_host, v4::g_stage_clock, sizeof(long long) * 16 * (size_t)n_ctas));
  return 0;
}
extern "C" int hh_debug_v4_warp_arrivals(long long* out_host, int32_t n_ctas) {
  if (!out_host || n_ctas <= 0 || n_ctas > 512) return fail(-1, "hh_debug_v4_warp_arrivals: bad argument");
  HH_CUDA(cudaDeviceSynchronize());
  HH_CUDA(cudaMemcpyFromSymbol(out_host, v4::g_warp_arrive, sizeof(long long) * 8 * 16 * (size_t)n_ctas));
  return 0;
}
#endif

extern "C" int hh_create(const hh_config* cfg, int32_t n_arenas, int32_t device, hh_env** out) {
  if (!cfg || !out) return fail(-1, "hh_create: null argument");
  if (n_arenas <= 0) return fail(-1, "hh_create: n_arenas must be positive");
  if (cfg->level < 1 || cfg->level > 5) return fail(-1, "hh_create: level must be 1..5");
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
  P.geom = make_geom(P.map_size);
  {
    const char* sm = getenv("HH_SHORT_MOVES");
    P.short_moves = (sm && sm[0] == '0') ? 0 : 1;
  }
  {
    const char* impl = getenv("HH_STEP_IMPL");
    if (2 * sizeof(v4::Smem) > 48 * 1024) {   // opt in to > 48 KB of dynamic shared memory (per device, idempotent; the dual form takes two)
      cudaError_t ce2 = v4_opt_in_smem<1, 0>();
      if (ce2 == cudaSuccess) ce2 = v4_opt_in_smem<1, 1>();
      if (ce2 == cudaSuccess) ce2 = v4_opt_in_smem<2, 0>();
      if (ce2 == cudaSuccess) ce2 = v4_opt_in_smem<2, 1>();
      if (ce2 == cudaSuccess) ce2 = v4_opt_in_smem<3, 0>();
      if (ce2 == cudaSuccess) ce2 = v4_opt_in_smem<3, 1>();
      if (ce2 != cudaSuccess) {
        hh_destroy(e);
        return fail(-2, std::string("cudaFuncSetAttribute(step_kernel_v4, max dynamic smem): ") + cudaGetErrorString(ce2));
      }
    }
    const char* hm = getenv("HH_HOST_MODE");
    // default: zero-copy (+23 % e2e over staged).  The pipelined mode (two half-batch launches, copy-engine D2H of the first half
    // under the second half's kernel, one CUDA graph) is bit-identical but measured SLOWER at 8 192 arenas (74.6 M env-steps/s
    // against 99.6 M): the graph's six nodes on two streams cost more in scheduling latency than the overlap returns
    // (profiles/README.md, round 2); it stays selectable (HH_HOST_MODE=pipelined, hh_set_host_mode(2)).
    e->host_mode = (hm && std::string(hm) == "staged") ? 0 : (hm && std::string(hm) == "pipelined") ? 2 : 1;
    e->step_impl = !impl ? 4 : (std::string(impl) == "quad" ? 2 : (std::string(impl) == "cta" ? 3 : 4));
  }
  *out = e;
  return 0;
}

extern "C" void hh_destroy(hh_env* e) {
  if (!e) return;
  cudaSetDevice(e->device);
  if (e->hstream) cudaStreamSynchronize(e->hstream);   // a step sent with hh_step_host_begin may still be in flight
  if (e->pipe_graph) cudaGraphExecDestroy(e->pipe_graph);
  if (e->cstream) cudaStreamDestroy(e->cstream);
  if (e->slab) cudaFree(e->slab);
  if (e->d_slab) cudaFree(e->d_slab);
  if (e->rew_pre) cudaFree(e->rew_pre);
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
  const int blocks = (e->n + kArenasPerCta - 1) / kArenasPerCta;
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

// arenas [first, end) of the v4 step (first a multiple of v4::kArenas); the whole batch is first = 0, end = n
template <int LEVEL>
static void launch_step_v4_range(hh_env* e, int first, int end, const int32_t* actions, float* obs1, float* obs2, float* rew,
                                 uint8_t* done, cudaStream_t st, bool dual = false, float* cen1 = nullptr, float* cen2 = nullptr,
                                 int cen_ld = 0) {
  Params P = e->P;
  P.n_arenas = end;
  const int block0 = first / v4::kArenas, vblocks = (end - first + v4::kArenas - 1) / v4::kArenas;
  if (dual) {
    const size_t sm2 = 2 * ((sizeof(v4::Smem) + 15) / 16 * 16);
    if (e->cfg.agent_mode == 0)
      step_kernel_v4_dual<LEVEL, 0><<<(vblocks + 1) / 2, 2 * v4::kThreads, sm2, st>>>(e->S, P, actions, obs1, obs2, rew, done, block0,
                                                                                      cen1, cen2, cen_ld);
    else
      step_kernel_v4_dual<LEVEL, 1><<<(vblocks + 1) / 2, 2 * v4::kThreads, sm2, st>>>(e->S, P, actions, obs1, obs2, rew, done, block0,
                                                                                      cen1, cen2, cen_ld);
    return;
  }
  if (e->cfg.agent_mode == 0)
    step_kernel_v4<LEVEL, 0><<<vblocks, v4::kThreads, sizeof(v4::Smem), st>>>(e->S, P, actions, obs1, obs2, rew, done, block0);
  else
    step_kernel_v4<LEVEL, 1><<<vblocks, v4::kThreads, sizeof(v4::Smem), st>>>(e->S, P, actions, obs1, obs2, rew, done, block0);
}

template <int LEVEL>
static void launch_step(hh_env* e, const int32_t* actions, float* obs1, float* obs2, float* rew, uint8_t* done,
                        cudaStream_t st) {
  if (e->step_impl == 4) {
    launch_step_v4_range<LEVEL>(e, 0, e->n, actions, obs1, obs2, rew, done, st);
    return;
  }
  if (e->step_impl == 3) {
    const int cblocks = (e->n + cta::kArenas - 1) / cta::kArenas;
    if (e->cfg.agent_mode == 0)
      step_kernel_cta<LEVEL, 0><<<cblocks, cta::kThreads, 0, st>>>(e->S, e->P, actions, obs1, obs2, rew, done);
    else
      step_kernel_cta<LEVEL, 1><<<cblocks, cta::kThreads, 0, st>>>(e->S, e->P, actions, obs1, obs2, rew, done);
    return;
  }
  const int blocks = (e->n + kArenasPerCta - 1) / kArenasPerCta;
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
  if (e->cfg.level >= 4)
    return fail(-5, "hh_step: levels 4/5 have frozen-policy opponents: use hh_step_begin / hh_step_finish");
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

// arenas [first, first + count) only: the pointers address the WHOLE batch's arrays (arena 0 first).  Lets a caller keep two
// halves of a batch in flight on two streams (the sampler's software pipeline: one half's policy forward runs while the other
// half steps).  Levels 1-3, default step kernel; first must be a multiple of 32.
static int step_range_impl(hh_env* e, int32_t first, int32_t count, const int32_t* actions_dev, float* obs1, float* obs2,
                           float* rew, uint8_t* done, float* cen1, float* cen2, int cen_ld, void* stream);
extern "C" int hh_step_range(hh_env* e, int32_t first, int32_t count, const int32_t* actions_dev, float* obs1, float* obs2,
                             float* rew, uint8_t* done, void* stream) {
  return step_range_impl(e, first, count, actions_dev, obs1, obs2, rew, done, nullptr, nullptr, 0, stream);
}
// ... and the observations ALSO (obs1 / obs2 may be NULL) as the next tick's central-critic rows of both policies:
// central1[a][7 ..] = [obs1 | obs2], central2[a][7 ..] = [obs2 | obs1] (what hh_pack_central would write), row stride ld floats
extern "C" int hh_step_range_central(hh_env* e, int32_t first, int32_t count, const int32_t* actions_dev, float* obs1, float* obs2,
                                     float* rew, uint8_t* done, float* central1, float* central2, int32_t ld, void* stream) {
  if (!central1 || !central2 || (e && ld < 7 + obs_dim(e->cfg, 1) + obs_dim(e->cfg, 2)))
    return fail(-3, "hh_step_range_central: central1 / central2 with a row stride of at least 7 + d1 + d2 floats");
  return step_range_impl(e, first, count, actions_dev, obs1, obs2, rew, done, central1, central2, ld, stream);
}
static int step_range_impl(hh_env* e, int32_t first, int32_t count, const int32_t* actions_dev, float* obs1, float* obs2,
                           float* rew, uint8_t* done, float* cen1, float* cen2, int cen_ld, void* stream) {
  if (!e) return fail(-1, "hh_step_range: null env");
  if (!e->initialised) return fail(-4, "hh_step_range: call hh_reset first");
  if (!actions_dev) return fail(-1, "hh_step_range: null actions");
  if (e->cfg.level >= 4) return fail(-5, "hh_step_range: levels 1-3 only (levels 4/5: hh_step_begin / hh_step_finish)");
  if (e->step_impl != 4) return fail(-5, "hh_step_range: needs the default step kernel (HH_STEP_IMPL=v4)");
  if (first < 0 || count <= 0 || first % v4::kArenas != 0 || first + count > e->n)
    return fail(-3, "hh_step_range: first must be a multiple of 32 and [first, first + count) inside the batch");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  // ranges run next to other kernels (that is what they are for): two sub-blocks per CTA = half as many CTAs spread over the SMs
  static const bool dual_env = [] { const char* v = getenv("HH_STEP_RANGE_DUAL"); return !v || atoi(v) != 0; }();
  const bool dual = dual_env || cen1;      // the central rows are written by the dual form
  switch (e->cfg.level) {
    case 1: launch_step_v4_range<1>(e, first, first + count, actions_dev, obs1, obs2, rew, done, st, dual, cen1, cen2, cen_ld); break;
    case 2: launch_step_v4_range<2>(e, first, first + count, actions_dev, obs1, obs2, rew, done, st, dual, cen1, cen2, cen_ld); break;
    default: launch_step_v4_range<3>(e, first, first + count, actions_dev, obs1, obs2, rew, done, st, dual, cen1, cen2, cen_ld); break;
  }
  HH_CUDA(cudaGetLastError());
  e->launches += 1;
  return 0;
}

extern "C" int hh_step_begin(hh_env* e, const int32_t* actions_dev, float* opp_obs3_dev, float* opp_obs4_dev,
                             uint8_t* policy_set_dev, void* stream) {
  if (!e) return fail(-1, "hh_step_begin: null env");
  if (!e->initialised) return fail(-4, "hh_step_begin: call hh_reset first");
  if (e->cfg.level < 4) return fail(-5, "hh_step_begin: only levels 4/5 split the step");
  if (!actions_dev || !opp_obs3_dev || !opp_obs4_dev) return fail(-1, "hh_step_begin: null argument");
  if (!e->rew_pre) HH_CUDA(cudaMalloc(&e->rew_pre, sizeof(float) * 2 * (size_t)e->n));
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int blocks = (e->n + kArenasPerCta - 1) / kArenasPerCta;
  if (e->cfg.agent_mode == 0)
    step_begin_kernel<0><<<blocks, kThreads, 0, st>>>(e->S, e->P, actions_dev, opp_obs3_dev, opp_obs4_dev, policy_set_dev,
                                                      e->rew_pre);
  else
    step_begin_kernel<1><<<blocks, kThreads, 0, st>>>(e->S, e->P, actions_dev, opp_obs3_dev, opp_obs4_dev, policy_set_dev,
                                                      e->rew_pre);
  HH_CUDA(cudaGetLastError());
  e->launches += 1;
  e->mid_step = true;
  return 0;
}

extern "C" int hh_step_finish(hh_env* e, const int32_t* opp_actions_dev, float* obs1, float* obs2, float* rew,
                              uint8_t* done, void* stream) {
  if (!e) return fail(-1, "hh_step_finish: null env");
  if (!e->mid_step) return fail(-4, "hh_step_finish: call hh_step_begin first");
  if (!opp_actions_dev) return fail(-1, "hh_step_finish: null opponent actions");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int blocks = (e->n + kArenasPerCta - 1) / kArenasPerCta;
  if (e->cfg.agent_mode == 0)
    step_finish_kernel<0><<<blocks, kThreads, 0, st>>>(e->S, e->P, opp_actions_dev, e->rew_pre, obs1, obs2, rew, done);
  else
    step_finish_kernel<1><<<blocks, kThreads, 0, st>>>(e->S, e->P, opp_actions_dev, e->rew_pre, obs1, obs2, rew, done);
  HH_CUDA(cudaGetLastError());
  e->launches += 1;
  e->mid_step = false;
  return 0;
}

// ------------------------------------------------------------------------------------------ host variants
// One pinned host slab and one device slab with the same layout:
//   [ actions i32 N*8 | obs1 f32 N*d1 | obs2 f32 N*d2 | rew f32 N*2 | done u8 N | mask u8 N ]
// so a host step is ONE H2D (actions), the launch, ONE D2H (obs1..done) and one stream sync.  Callers that
// use the slab's own regions (hh_host_buffers) skip the host-side memcpys as well.
namespace {
struct HostLayout {
  size_t o_act, o_obs1, o_obs2, o_rew, o_done, o_mask, total, out_bytes;
};
HostLayout host_layout(const hh_env* e);
}  // namespace

static int ensure_staging(hh_env* e) {
  if (e->hstream) return 0;
  const HostLayout h = host_layout(e);
  HH_CUDA(cudaSetDevice(e->device));
  HH_CUDA(cudaStreamCreateWithFlags(&e->hstream, cudaStreamNonBlocking));
  HH_CUDA(cudaMalloc(&e->d_slab, h.total));
  HH_CUDA(cudaHostAlloc(&e->pinned, h.total, cudaHostAllocMapped | cudaHostAllocPortable));
  memset(e->pinned, 0, h.total);
  HH_CUDA(cudaHostGetDevicePointer(&e->pinned_dev, e->pinned, 0));
  e->pinned_bytes = h.total;
  char* d = static_cast<char*>(e->d_slab);
  e->d_actions = reinterpret_cast<int32_t*>(d + h.o_act);
  e->d_obs1 = reinterpret_cast<float*>(d + h.o_obs1);
  e->d_obs2 = reinterpret_cast<float*>(d + h.o_obs2);
  e->d_rew = reinterpret_cast<float*>(d + h.o_rew);
  e->d_done = reinterpret_cast<uint8_t*>(d + h.o_done);
  e->d_mask = reinterpret_cast<uint8_t*>(d + h.o_mask);
  return 0;
}

namespace {
HostLayout host_layout(const hh_env* e) {
  const size_t N = (size_t)e->n;
  const int d1 = obs_dim(e->cfg, 1), d2 = obs_dim(e->cfg, 2);
  HostLayout h;
  h.o_act = 0;
  h.o_obs1 = align_up(N * 8 * sizeof(int32_t), 256);
  h.o_obs2 = align_up(h.o_obs1 + N * d1 * sizeof(float), 256);   // obs1..done travel in ONE D2H (regions 256-B
  h.o_rew = align_up(h.o_obs2 + N * d2 * sizeof(float), 256);    // aligned: the kernels store rows as float4)
  h.o_done = align_up(h.o_rew + N * 2 * sizeof(float), 256);
  h.out_bytes = h.o_done + N - h.o_obs1;
  h.o_mask = align_up(h.o_done + N, 256);
  h.total = align_up(h.o_mask + N, 256);
  return h;
}
}  // namespace

extern "C" int hh_host_buffers(hh_env* e, int32_t** actions, float** obs1, float** obs2, float** rew, uint8_t** done) {
  if (!e) return fail(-1, "hh_host_buffers: null env");
  int rc = ensure_staging(e);
  if (rc) return rc;
  const HostLayout h = host_layout(e);
  char* p = static_cast<char*>(e->pinned);
  if (actions) *actions = reinterpret_cast<int32_t*>(p + h.o_act);
  if (obs1) *obs1 = reinterpret_cast<float*>(p + h.o_obs1);
  if (obs2) *obs2 = reinterpret_cast<float*>(p + h.o_obs2);
  if (rew) *rew = reinterpret_cast<float*>(p + h.o_rew);
  if (done) *done = reinterpret_cast<uint8_t*>(p + h.o_done);
  return 0;
}

extern "C" int hh_set_host_mode(hh_env* e, int32_t mode) {
  if (!e) return fail(-1, "hh_set_host_mode: null env");
  if (mode < 0 || mode > 2) return fail(-1, "hh_set_host_mode: mode must be 0 (staged copies), 1 (zero-copy) or 2 (pipelined)");
  e->host_mode = mode;
  return 0;
}

extern "C" int hh_reset_host(hh_env* e, const uint8_t* mask_host, float* obs1_host, float* obs2_host) {
  if (!e) return fail(-1, "hh_reset_host: null env");
  int rc = ensure_staging(e);
  if (rc) return rc;
  const size_t N = (size_t)e->n;
  const int d1 = obs_dim(e->cfg, 1), d2 = obs_dim(e->cfg, 2);
  const HostLayout h = host_layout(e);
  char* pin = static_cast<char*>(e->pinned);
  if (mask_host) {
    memcpy(pin + h.o_mask, mask_host, N);
    HH_CUDA(cudaMemcpyAsync(e->d_mask, pin + h.o_mask, N, cudaMemcpyHostToDevice, e->hstream));
  }
  if (e->host_mode >= 1) {
    char* pd = static_cast<char*>(e->pinned_dev);
    rc = hh_reset(e, mask_host ? e->d_mask : nullptr, reinterpret_cast<float*>(pd + h.o_obs1),
                  reinterpret_cast<float*>(pd + h.o_obs2), e->hstream);
    if (rc) return rc;
  } else {
    rc = hh_reset(e, mask_host ? e->d_mask : nullptr, e->d_obs1, e->d_obs2, e->hstream);
    if (rc) return rc;
    HH_CUDA(cudaMemcpyAsync(pin + h.o_obs1, e->d_obs1, h.o_obs2 + N * d2 * sizeof(float) - h.o_obs1, cudaMemcpyDeviceToHost,
                            e->hstream));
  }
  HH_CUDA(cudaStreamSynchronize(e->hstream));
  if (obs1_host && obs1_host != reinterpret_cast<float*>(pin + h.o_obs1)) memcpy(obs1_host, pin + h.o_obs1, N * d1 * sizeof(float));
  if (obs2_host && obs2_host != reinterpret_cast<float*>(pin + h.o_obs2)) memcpy(obs2_host, pin + h.o_obs2, N * d2 * sizeof(float));
  return 0;
}

// Pipelined host step (host mode 2).  A lone CTA of the step kernel needs ~14 us, two co-resident ones ~22 us, and the
// 1.6 MB of observations need ~33 us over PCIe: so the batch is stepped as TWO half-batch launches back to back (one CTA per
// SM each), and the first half's observations travel (copy engine, D2H) while the second half computes.  Actions are read
// from, rewards / done flags written to, the pinned slab directly (small); the whole step is one CUDA-graph launch.
static bool pipelined_applies(const hh_env* e) {
  return e->cfg.level <= 3 && e->step_impl == 4 && e->n >= 2048 && e->initialised;
}
static int pipelined_step(hh_env* e) {
  if (!e->pipe_graph) {
    const HostLayout h = host_layout(e);
    const size_t N = (size_t)e->n;
    const int d1 = obs_dim(e->cfg, 1), d2 = obs_dim(e->cfg, 2);
    const int half = (e->n / 2 + v4::kArenas - 1) / v4::kArenas * v4::kArenas;
    char* pd = static_cast<char*>(e->pinned_dev);
    char* pin = static_cast<char*>(e->pinned);
    const int32_t* act = reinterpret_cast<const int32_t*>(pd + h.o_act);
    float* rew = reinterpret_cast<float*>(pd + h.o_rew);
    uint8_t* done = reinterpret_cast<uint8_t*>(pd + h.o_done);
    if (!e->cstream) HH_CUDA(cudaStreamCreateWithFlags(&e->cstream, cudaStreamNonBlocking));
    cudaEvent_t ev[3];
    for (auto& x : ev) HH_CUDA(cudaEventCreateWithFlags(&x, cudaEventDisableTiming));
    cudaGraph_t graph = nullptr;
    HH_CUDA(cudaStreamBeginCapture(e->hstream, cudaStreamCaptureModeThreadLocal));
    const int bounds[3] = {0, half, e->n};
    for (int c = 0; c < 2; ++c) {
      switch (e->cfg.level) {
        case 1: launch_step_v4_range<1>(e, bounds[c], bounds[c + 1], act, e->d_obs1, e->d_obs2, rew, done, e->hstream); break;
        case 2: launch_step_v4_range<2>(e, bounds[c], bounds[c + 1], act, e->d_obs1, e->d_obs2, rew, done, e->hstream); break;
        default: launch_step_v4_range<3>(e, bounds[c], bounds[c + 1], act, e->d_obs1, e->d_obs2, rew, done, e->hstream); break;
      }
      cudaEventRecord(ev[c], e->hstream);
      cudaStreamWaitEvent(e->cstream, ev[c], 0);
      const size_t a0 = (size_t)bounds[c], cnt = (size_t)(bounds[c + 1] - bounds[c]);
      cudaMemcpyAsync(pin + h.o_obs1 + a0 * d1 * sizeof(float), e->d_obs1 + a0 * d1, cnt * d1 * sizeof(float), cudaMemcpyDeviceToHost,
                      e->cstream);
      cudaMemcpyAsync(pin + h.o_obs2 + a0 * d2 * sizeof(float), e->d_obs2 + a0 * d2, cnt * d2 * sizeof(float), cudaMemcpyDeviceToHost,
                      e->cstream);
    }
    (void)N;
    cudaEventRecord(ev[2], e->cstream);
    cudaStreamWaitEvent(e->hstream, ev[2], 0);
    cudaError_t ce = cudaStreamEndCapture(e->hstream, &graph);
    for (auto& x : ev) cudaEventDestroy(x);
    if (ce != cudaSuccess || !graph) return fail(-2, std::string("pipelined host step: stream capture failed: ") + cudaGetErrorString(ce));
    ce = cudaGraphInstantiate(&e->pipe_graph, graph, 0);
    cudaGraphDestroy(graph);
    if (ce != cudaSuccess) return fail(-2, std::string("pipelined host step: cudaGraphInstantiate: ") + cudaGetErrorString(ce));
  }
  HH_CUDA(cudaGraphLaunch(e->pipe_graph, e->hstream));
  e->launches += 2;
  return 0;
}

// send / poll halves of a host step (RLlib's own asynchronous env interface is BaseEnv.send_actions() / poll()):
// _begin enqueues the step on the handle's private stream and returns; _end waits for it and delivers the results.
// Several handles (e.g. two halves of a batch) can be in flight at once, so that one handle's PCIe traffic and the
// caller's preparation of the next actions overlap another handle's kernel.
extern "C" int hh_step_host_begin(hh_env* e, const int32_t* actions_host) {
  if (!e) return fail(-1, "hh_step_host_begin: null env");
  if (!actions_host) return fail(-1, "hh_step_host_begin: null actions");
  if (e->host_pending) return fail(-4, "hh_step_host_begin: the previous step was not collected (hh_step_host_end)");
  int rc = ensure_staging(e);
  if (rc) return rc;
  const size_t N = (size_t)e->n;
  const HostLayout h = host_layout(e);
  char* pin = static_cast<char*>(e->pinned);
  if (reinterpret_cast<const char*>(actions_host) != pin + h.o_act) memcpy(pin + h.o_act, actions_host, N * 8 * sizeof(int32_t));
  if (e->host_mode == 2 && pipelined_applies(e)) {
    rc = pipelined_step(e);
    if (rc) return rc;
  } else if (e->host_mode >= 1) {
    // zero-copy: the step kernel loads the actions from, and stores observations / rewards / done flags to, the
    // pinned slab through its device-side mapping -- no staging copies, the PCIe traffic rides on the kernel
    char* pd = static_cast<char*>(e->pinned_dev);
    rc = hh_step(e, reinterpret_cast<const int32_t*>(pd + h.o_act), reinterpret_cast<float*>(pd + h.o_obs1),
                 reinterpret_cast<float*>(pd + h.o_obs2), reinterpret_cast<float*>(pd + h.o_rew),
                 reinterpret_cast<uint8_t*>(pd + h.o_done), e->hstream);
    if (rc) return rc;
  } else {
    HH_CUDA(cudaMemcpyAsync(e->d_actions, pin + h.o_act, N * 8 * sizeof(int32_t), cudaMemcpyHostToDevice, e->hstream));
    rc = hh_step(e, e->d_actions, e->d_obs1, e->d_obs2, e->d_rew, e->d_done, e->hstream);
    if (rc) return rc;
    HH_CUDA(cudaMemcpyAsync(pin + h.o_obs1, e->d_obs1, h.out_bytes, cudaMemcpyDeviceToHost, e->hstream));
  }
  e->host_pending = true;
  return 0;
}

extern "C" int hh_step_host_end(hh_env* e, float* obs1_host, float* obs2_host, float* rew_host, uint8_t* done_host) {
  if (!e) return fail(-1, "hh_step_host_end: null env");
  if (!e->host_pending) return fail(-4, "hh_step_host_end: no step in flight (hh_step_host_begin)");
  const size_t N = (size_t)e->n;
  const int d1 = obs_dim(e->cfg, 1), d2 = obs_dim(e->cfg, 2);
  const HostLayout h = host_layout(e);
  char* pin = static_cast<char*>(e->pinned);
  e->host_pending = false;
  HH_CUDA(cudaStreamSynchronize(e->hstream));
  if (obs1_host && reinterpret_cast<char*>(obs1_host) != pin + h.o_obs1) memcpy(obs1_host, pin + h.o_obs1, N * d1 * sizeof(float));
  if (obs2_host && reinterpret_cast<char*>(obs2_host) != pin + h.o_obs2) memcpy(obs2_host, pin + h.o_obs2, N * d2 * sizeof(float));
  if (rew_host && reinterpret_cast<char*>(rew_host) != pin + h.o_rew) memcpy(rew_host, pin + h.o_rew, N * 2 * sizeof(float));
  if (done_host && reinterpret_cast<char*>(done_host) != pin + h.o_done) memcpy(done_host, pin + h.o_done, N);
  return 0;
}

extern "C" int hh_step_host(hh_env* e, const int32_t* actions_host, float* obs1_host, float* obs2_host,
                            float* rew_host, uint8_t* done_host) {
  const int rc = hh_step_host_begin(e, actions_host);
  if (rc) return rc;
  return hh_step_host_end(e, obs1_host, obs2_host, rew_host, done_host);
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
  unpack_state(N, h.ac.data(), h.ri.data(), h.meta.data(), h.dg.data(), o);
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
  pack_state(N, in, h.ac.data(), h.ri.data(), h.meta.data(), h.dg.data());
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
