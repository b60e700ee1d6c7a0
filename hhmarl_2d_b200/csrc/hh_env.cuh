// hh_env.cuh -- device-side 2-vs-2 low-level air-combat arena: reset, action phase, tick,
// rewards, observations.  One thread advances one arena whose whole state lives in registers.
//
// Behavioural contract: SURVEY.md Appendix A; every block cites the reference lines it
// reproduces (paths relative to the reference repo).  The structure is NOT the reference's
// (no unit registry, no event list, no dict lookups): units are fixed register slots, the
// "snapshot + removal" semantics of CmanoSimulator.do_tick (cmano_simulator.py:138-157) become
// alive-bit bookkeeping, and the event list becomes a per-victim killer record.
#pragma once
#include "hh_geodesic.cuh"
#include "hh_state.cuh"

namespace hh {

constexpr double kKnotsToMs = 0.514444;  // cmano_simulator.py:21
constexpr int OBS_AC1 = 26, OBS_AC2 = 24, OBS_ESC_AC1 = 30, OBS_ESC_AC2 = 29;

// rocket_unit.py:16-21 -- scipy quadratic spline through (0,500),(10,2000),(20,1400),(30,600)
// sampled at life_time 0..10 s (oracle/gen_rocket_table.py).
__device__ __forceinline__ double rocket_speed(int life) {
  switch (life) {
    case 0: return 0x1.f400000000000p+8;
    case 1: return 0x1.7b5ffffffffffp+9;
    case 2: return 0x1.f0aaaaaaaaaacp+9;
    case 3: return 0x1.2cf0000000000p+10;
    case 4: return 0x1.5b80000000000p+10;
    case 5: return 0x1.8405555555554p+10;
    case 6: return 0x1.a680000000001p+10;
    case 7: return 0x1.c2f0000000000p+10;
    case 8: return 0x1.d955555555556p+10;
    case 9: return 0x1.e9b0000000000p+10;
    default: return 0x1.f400000000000p+10;
  }
}

// ------------------------------------------------------------------------------------- RNG
// Philox4x32-10, key = seed, counter = (draw_lo, draw_hi, arena_id, stream); SURVEY.md A.5.
__device__ __forceinline__ double philox_u53(uint32_t k0, uint32_t k1, uint32_t c0, uint32_t c1,
                                             uint32_t c2, uint32_t c3) {
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
    uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
    uint32_t n0 = hi1 ^ c1 ^ k0, n2 = hi0 ^ c3 ^ k1;
    c0 = n0; c1 = lo1; c2 = n2; c3 = lo0;
    k0 += 0x9E3779B9u;
    k1 += 0xBB67AE85u;
  }
  return ((double)(c0 >> 5) * 67108864.0 + (double)(c1 >> 6)) * (1.0 / 9007199254740992.0);
}

struct Rng {
  uint32_t k0, k1, arena;
};
__device__ __forceinline__ double g_random(const Rng& r, Arena& A) {
  double v = philox_u53(r.k0, r.k1, (uint32_t)A.dg, (uint32_t)(A.dg >> 32), r.arena, 0u);
  A.dg += 1;
  return v;
}
__device__ __forceinline__ double g_uniform(const Rng& r, Arena& A, double a, double b) {
  return __dadd_rn(a, __dmul_rn(b - a, g_random(r, A)));  // a + (b-a)*random(), never fused
}
__device__ __forceinline__ int g_randint(const Rng& r, Arena& A, int a, int b) {
  return a + (int)(g_random(r, A) * (double)(b - a + 1));
}
__device__ __forceinline__ double c_random(const Rng& r, Arena& A) {
  double v = philox_u53(r.k0, r.k1, A.dc, 0u, r.arena, 1u);
  A.dc += 1;
  return v;
}

// ------------------------------------------------------------------------------------- scalar helpers
__device__ __forceinline__ double pymod(double x, double m) {  // CPython float %, m > 0
  double r = fmod(x, m);
  if (r != 0.0 && r < 0.0) r += m;
  return r;
}
__device__ __forceinline__ double clip(double x, double lo, double hi) {
  return x < lo ? lo : (x > hi ? hi : x);
}
// angles.py:22-29
__device__ __forceinline__ double signed_heading_diff(double actual, double desired) {
  double delta = desired - actual;
  if (delta < -180.0) delta = 360.0 + delta;
  if (delta > 180.0) delta = -360.0 + delta;
  return delta;
}
// angles.py:10-15 applied to an azimuth in (-180, 180]
__device__ __forceinline__ double normalize_angle(double a) {
  while (a >= 360.0) a -= 360.0;
  while (a < 0.0) a += 360.0;
  return a;
}
__device__ __forceinline__ constexpr bool is_ac1(int u) { return (u & 1) == 0; }
__device__ __forceinline__ constexpr double max_speed(int u) { return is_ac1(u) ? 900.0 : 600.0; }

struct Geom {     // per-step constants of the map (env_base.py:43, map_limits.py)
  double ext_lat, ext_lon, top, right;
  double inv_diag;  // (1 - 0) / (sqrt(2 ms^2) - 0)   (env_base.py:439,458-462)
};
__device__ __forceinline__ Geom make_geom(double ms) {
  Geom g;
  g.top = 5.0 + ms;
  g.right = 7.0 + ms;
  g.ext_lat = g.top - 5.0;
  g.ext_lon = g.right - 7.0;
  g.inv_diag = 1.0 / sqrt(2.0 * (ms * ms));
  return g;
}
// map_limits.py:37-40
__device__ __forceinline__ void rel_pos(const Geom& g, double lat, double lon, double& lat_rel,
                                        double& lon_rel) {
  lat_rel = clip((lat - 5.0) / g.ext_lat, 0.0, 1.0);
  lon_rel = clip((lon - 7.0) / g.ext_lon, 0.0, 1.0);
}
// map_limits.py:47-48
__device__ __forceinline__ bool in_boundary(const Geom& g, double lat, double lon) {
  return 7.0 <= lon && lon <= g.right && 5.0 <= lat && lat <= g.top;
}

// heading unit vector of env_base.py:428 / :452: (cos, sin) of ((90 - heading) % 360) * pi/180
struct HVec {
  double c, s, n;
};
__device__ __forceinline__ HVec heading_vec(double heading) {
  HVec h;
  double th = pymod(90.0 - heading, 360.0) * (geo::kPi / 180.0);
  sincos(th, &h.s, &h.c);
  h.n = sqrt(h.c * h.c + h.s * h.s);
  return h;
}
// env_base.py:424-432 -- degrees
__device__ __forceinline__ double focus_deg(const HVec& ha, double lat_a, double lon_a, double lat_b,
                                            double lon_b) {
  double v0 = lon_b - lon_a, v1 = lat_b - lat_a;
  double x = clip((ha.c * v0 + ha.s * v1) / (ha.n * sqrt(v0 * v0 + v1 * v1) + 1e-10), -1.0, 1.0);
  return acos(x) * (180.0 / geo::kPi);
}
__device__ __forceinline__ double focus_norm_from_deg(double deg) { return clip(deg / 180.0, 0.0, 1.0); }
// env_base.py:441-446
__device__ __forceinline__ double aspect_from_deg(double deg) { return clip((180.0 - deg) / 180.0, 0.0, 1.0); }
// env_base.py:448-456
__device__ __forceinline__ double hdiff_norm(const HVec& a, const HVec& b) {
  double x = clip((a.c * b.c + a.s * b.s) / (a.n * b.n + 1e-10), -1.0, 1.0);
  return clip((acos(x) * (180.0 / geo::kPi)) / 180.0, 0.0, 1.0);
}
// env_base.py:434-439
__device__ __forceinline__ double dist_raw(double lat_a, double lon_a, double lat_b, double lon_b) {
  return hypot(lon_b - lon_a, lat_b - lat_a);
}

// ------------------------------------------------------------------------------------- weapons
// ac1.py:144-146
__device__ __forceinline__ bool angle_in_radar_range(double heading, double angle) {
  double c = heading + 60.0;  // sum_angles(heading, 120/2)
  while (c >= 360.0) c -= 360.0;
  while (c < 0.0) c += 360.0;
  double delta = fabs(signed_heading_diff(c, angle));
  return (int)delta <= 60;
}

// ac1.py:72-79 (+ Rocket.__init__, rocket_unit.py:23-30). shooter u (AC1), target index t.
__device__ __forceinline__ void fire_missile(Arena& A, int u, int t) {
  if (!A.hasm[u] && A.mrem[u] > 0) {
    double2 inv = geo::inverse(A.lat[u], A.lon[u], pick4(A.lat, t), pick4(A.lon, t));
    if (inv.x / 1000.0 <= 111.0 && angle_in_radar_range(A.hdg[u], normalize_angle(inv.y))) {
      const int s = u >> 1;
      A.rlat[s] = A.lat[u];
      A.rlon[s] = A.lon[u];
      A.rhdg[s] = A.hdg[u];
      A.rnhdg[s] = A.hdg[u];
      A.ralive[s] = true;
      A.rage[s] = 0;
      A.rtgt[s] = t + 1;
      A.rid[s] = A.next_id;
      A.next_id += 1;
      A.hasm[u] = true;
      A.mrem[u] = A.mrem[u] - 1;
    }
  }
}
// ac1.py:69-70 / ac2.py:65-66
__device__ __forceinline__ void fire_cannon(Arena& A, int u) {
  const int bt = is_ac1(u) ? 5 : 3;
  A.burst[u] = A.crem[u] < bt ? A.crem[u] : bt;
}
__device__ __forceinline__ void set_heading(Arena& A, int u, double h) {
  if (h >= 360.0 || h < 0.0) A.err |= ERR_HEADING;  // the reference raises (ac1.py:58-61)
  A.nhdg[u] = h;
}
__device__ __forceinline__ void set_speed(Arena& A, int u, double s) {
  if (s > max_speed(u) || s < 0.0) A.err |= ERR_SPEED;  // ac1.py:63-67
  A.nspd[u] = s;
}

// nearest live enemy of unit u by normalised flat distance (env_base.py:400-422);
// returns index or -1; d_norm of the winner in dn. Ties -> lower id (stable sort).
__device__ __forceinline__ int nearest_enemy(const Arena& A, const Geom& g, int u, double& dn) {
  const int e0 = u < 2 ? 2 : 0, e1 = e0 + 1;
  double d0 = g.inv_diag * dist_raw(A.lat[u], A.lon[u], A.lat[e0], A.lon[e0]);
  double d1 = g.inv_diag * dist_raw(A.lat[u], A.lon[u], A.lat[e1], A.lon[e1]);
  int best = -1;
  dn = 0.0;
  if (A.alive[e0]) { best = e0; dn = d0; }
  if (A.alive[e1] && (best < 0 || d1 < d0)) { best = e1; dn = d1; }
  return best;
}

// ------------------------------------------------------------------------------------- action phase
// env_base.py:214-238 (mode "LowLevel"); u compile-time, act = 4 ints
template <int MODE>
__device__ __forceinline__ void take_base_action(Arena& A, const Rng& rng, int u, int opp_id,
                                                 const int4 act, double& rew) {
  set_heading(A, u, pymod(A.hdg[u] + (double)((act.x - 6) * 15), 360.0));
  set_speed(A, u, 100.0 + ((max_speed(u) - 100.0) / 8.0) * (double)act.y);
  if (act.z != 0 && A.crem[u] > 0) {
    fire_cannon(A, u);
    if (MODE == 1 && u < 2 && A.crem[u] < 90) rew -= 0.1;
  }
  if (is_ac1(u) && act.w != 0) {
    if (opp_id != 0 && A.mrem[u] > 0 && !A.hasm[u] && A.mwait[u] == 0) {
      fire_missile(A, u, opp_id - 1);
      A.mwait[u] = g_randint(rng, A, 7, 17);
      if (MODE == 1 && u < 2 && A.mrem[u] < 3) rew -= 0.1;
    }
  }
  if (A.mwait[u] > 0 && !A.hasm[u]) A.mwait[u] -= 1;
}

// env_hetero.py:118-123 (and the identical tail of __opp_level2, :132-136)
__device__ __forceinline__ void opp_missile_rule(Arena& A, const Rng& rng, const Geom& g, int u) {
  if (!A.hasm[u] && (A.steps % 40) < 3 && g_randint(rng, A, 0, 1) != 0 && A.mwait[u] == 0 && is_ac1(u)) {
    double dn;
    int t = nearest_enemy(A, g, u, dn);
    if (t >= 0) {
      fire_missile(A, u, t);
      A.mwait[u] = 5;
    }
  }
}
// env_hetero.py:125-136
__device__ __forceinline__ void opp_level2(Arena& A, const Rng& rng, const Geom& g, int u) {
  fire_cannon(A, u);
  bool turn = A.steps <= 5;
  if (!turn) turn = (A.steps % g_randint(rng, A, 35, 45)) <= 5;
  if (turn) {
    int r = g_randint(rng, A, 0, 1);
    set_heading(A, u, pymod(A.hdg[u] + (r ? -90.0 : 90.0), 360.0));
    set_speed(A, u, (double)(100 + g_randint(rng, A, 0, 4) * 75));
  }
  opp_missile_rule(A, rng, g, u);
}
// env_base.py:464-487
__device__ __forceinline__ int correct_angle_sign(double lat_o, double lon_o, double hdg_o, double lat_a,
                                                  double lon_a) {
  double s, c;
  sincos(pymod(hdg_o, 360.0) * (geo::kPi / 180.0), &s, &c);
  double x1 = lon_o + rint(s * 1000.0) / 1000.0;  // round(sin, 3)
  double y1 = lat_o + rint(c * 1000.0) / 1000.0;
  double val = (x1 - lon_o) * (lat_a - lat_o) - (lon_a - lon_o) * (y1 - lat_o);
  return val < 0.0 ? 1 : -1;
}
// env_hetero.py:138-158 with _escaping_opp (:227-245) and _hardcoded_opp (:247-271)
__device__ __forceinline__ void opp_level3(Arena& A, const Rng& rng, const Geom& g, int u) {
  if (A.steps % 60 == 0 && !A.escaping) {
    A.escaping = g_randint(rng, A, 0, 1) != 0;
    if (A.escaping) A.esc_time = (int)g_uniform(rng, A, 20.0, 30.0);
  }
  double heading, speed;
  bool fire = false, fire_m = false;
  int opp = -1;
  if (A.escaping) {
    double y, x;
    rel_pos(g, A.lat[u], A.lon[u], y, x);
    double lo = y < 0.5 ? (x < 0.5 ? 30.0 : 300.0) : (x < 0.5 ? 120.0 : 210.0);
    heading = (double)(int)g_uniform(rng, A, lo, lo + 30.0);
    speed = (double)(int)g_uniform(rng, A, 300.0, 600.0);
    fire = g_randint(rng, A, 0, 1) != 0;
    A.esc_time -= 1;
    if (A.esc_time <= 0) A.escaping = false;
  } else {
    heading = A.hdg[u];
    speed = (double)(int)g_uniform(rng, A, 100.0, 400.0);
    double dn;
    int t = nearest_enemy(A, g, u, dn);
    if (t >= 0) {
      double lat_t = pick4(A.lat, t), lon_t = pick4(A.lon, t);
      int sign = correct_angle_sign(A.lat[u], A.lon[u], A.hdg[u], lat_t, lon_t);
      double r = g_uniform(rng, A, 0.7, 1.3);
      double focus = focus_deg(heading_vec(A.hdg[u]), A.lat[u], A.lon[u], lat_t, lon_t);
      if (dn > 0.008 && focus > 4.0)
        heading = pymod(__dadd_rn(heading, __dmul_rn(__dmul_rn(r, (double)sign), focus)), 360.0);
      if (dn > 0.05)
        speed = focus < 30.0 ? (double)(int)g_uniform(rng, A, 500.0, 800.0)
                             : (double)(int)g_uniform(rng, A, 100.0, 500.0);
      fire = dn < 0.03 && focus < 10.0;
      fire_m = dn < 0.09 && focus < 5.0;
      opp = t;
    }
    if (!is_ac1(u)) speed = clip(speed, 0.0, 600.0);
  }
  set_heading(A, u, heading);
  set_speed(A, u, speed);
  if (fire) fire_cannon(A, u);
  if (fire_m && opp >= 0 && !A.hasm[u] && A.mwait[u] == 0 && is_ac1(u)) {
    fire_missile(A, u, opp);
    A.mwait[u] = 10;
  }
}

// ------------------------------------------------------------------------------------- tick
struct Kills {
  int killer[4];    // 0 = none, else killer id 1..4 (UnitDestroyedEvent.unit_killer)
  bool by_rocket[4];  // event origin is a Rocket (origin.id >= total_num + 1)
};

// Rafale.update / RafaleLong.update (ac1.py:81-133, ac2.py:68-107) for aircraft u.
template <int U>
__device__ __forceinline__ void aircraft_update(Arena& A, const Rng& rng, const Params& P, Kills& K) {
  constexpr double max_deg = is_ac1(U) ? 5.0 : 3.5;
  constexpr double max_kn = is_ac1(U) ? 35.0 : 28.0;
  if (A.hdg[U] != A.nhdg[U]) {
    double delta = signed_heading_diff(A.hdg[U], A.nhdg[U]);
    if (fabs(delta) <= max_deg) {
      A.hdg[U] = A.nhdg[U];
    } else {
      A.hdg[U] = pymod(A.hdg[U] + (delta >= 0.0 ? max_deg : -max_deg), 360.0);
    }
  }
  if (A.spd[U] != A.nspd[U]) {
    double delta = A.nspd[U] - A.spd[U];
    if (fabs(delta) <= max_kn)
      A.spd[U] = A.nspd[U];
    else
      A.spd[U] += delta >= 0.0 ? max_kn : -max_kn;
  }
  if (A.burst[U] > 0) {
    constexpr double range = is_ac1(U) ? 2.0 : 4.5;
    constexpr double half_width = (is_ac1(U) ? 10.0 : 7.0) / 2.0;
    const double p_hit = is_ac1(U) ? 0.75 / (5.0 / 1.0) : 0.9 / (3.0 / 1.0);
    A.burst[U] -= 1;                              // max(burst - tick, 0), burst >= 1 here
    A.crem[U] = A.crem[U] > 0 ? A.crem[U] - 1 : 0;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      if (j == U) continue;
      bool group_ok = P.friendly_kill || ((U < 2) != (j < 2));
      if (A.alive[j] && group_ok) {               // target still in sim.active_units *now*
        double2 inv = geo::inverse(A.lat[U], A.lon[U], A.lat[j], A.lon[j]);
        if (inv.x / 1000.0 < range) {
          double delta = fabs(signed_heading_diff(A.hdg[U], normalize_angle(inv.y)));
          if (delta <= half_width) {
            if (c_random(rng, A) < p_hit) {
              A.alive[j] = false;
              K.killer[j] = U + 1;
              K.by_rocket[j] = false;
            }
          }
        }
      }
    }
  }
  if (is_ac1(U) && A.hasm[U]) {                   // ac1.py:117-128
    constexpr int s = U >> 1;
    if (!A.ralive[s]) {
      A.hasm[U] = false;
    } else {
      double h = clip(__dmul_rn(A.rhdg[s], g_uniform(rng, A, 0.95, 1.05)), 0.0, 359.0);
      A.rnhdg[s] = h;                             // Rocket.set_heading: always inside [0, 360)
    }
  }
  if (A.spd[U] > 0.0) {                           // Unit.update, cmano_simulator.py:65-72
    double2 p = geo::direct(A.lat[U], A.lon[U], A.hdg[U], A.spd[U] * kKnotsToMs * 1.0);
    A.lat[U] = p.x;
    A.lon[U] = p.y;
  }
}

// Rocket.update (rocket_unit.py:37-73) for slot s (source id = 2s+1)
__device__ __forceinline__ void rocket_update(Arena& A, const Params& P, Kills& K, int s) {
  double rlat = pick2(A.rlat, s), rlon = pick2(A.rlon, s);
  int t = pick2(A.rtgt, s) - 1;
  int source = 2 * s + 1;
  double d = geo::inverse(rlat, rlon, pick4(A.lat, t), pick4(A.lon, t)).x / 1000.0;
  if (d < 1.0 && pick4(A.alive, t)) {
    put2(A.ralive, s, false);
    put4(A.alive, t, false);
    put4(K.killer, t, source);
    put4(K.by_rocket, t, true);
    return;
  }
  if (P.friendly_kill) {
    // friendly_id = 1 if source.id == 2 else 2 -> always id 2 here (sources are ids 1 and 3)
    if (A.alive[1]) {
      double df = geo::inverse(rlat, rlon, A.lat[1], A.lon[1]).x / 1000.0;
      if (df < 1.0) {
        put2(A.ralive, s, false);
        A.alive[1] = false;
        K.killer[1] = source;
        K.by_rocket[1] = true;
        return;
      }
    }
  }
  int life = pick2(A.rage, s);
  if (life > 10) {
    put2(A.ralive, s, false);
    return;
  }
  double h = pick2(A.rhdg, s), nh = pick2(A.rnhdg, s);
  if (h != nh) {
    double delta = signed_heading_diff(h, nh);
    if (fabs(delta) <= 10.0)
      h = nh;
    else
      h += delta >= 0.0 ? 10.0 : -10.0;
  }
  double2 p = geo::direct(rlat, rlon, h, rocket_speed(life) * kKnotsToMs * 1.0);
  put2(A.rhdg, s, h);
  put2(A.rlat, s, p.x);
  put2(A.rlon, s, p.y);
  put2(A.rage, s, life + 1);
}

// CmanoSimulator.do_tick (cmano_simulator.py:138-157): snapshot of active units in id order
// (aircraft 1..4, then rockets in launch order); units removed earlier in the tick still update.
__device__ __forceinline__ void do_tick(Arena& A, const Rng& rng, const Params& P, Kills& K) {
  const bool a0 = A.alive[0], a1 = A.alive[1], a2 = A.alive[2], a3 = A.alive[3];
  const bool r0 = A.ralive[0], r1 = A.ralive[1];
  if (a0) aircraft_update<0>(A, rng, P, K);
  if (a1) aircraft_update<1>(A, rng, P, K);
  if (a2) aircraft_update<2>(A, rng, P, K);
  if (a3) aircraft_update<3>(A, rng, P, K);
  const int first = (r0 && r1) ? (A.rid[0] < A.rid[1] ? 0 : 1) : (r0 ? 0 : 1);
#pragma unroll 1
  for (int k = 0; k < 2; ++k) {
    int s = k == 0 ? first : 1 - first;
    bool live = s == 0 ? r0 : r1;
    if (live) rocket_update(A, P, K, s);
  }
}

// ------------------------------------------------------------------------------------- rewards
// _combat_rewards (env_base.py:240-310, mode "LowLevel") + _get_rewards (env_hetero.py:188-225)
template <int MODE>
__device__ __forceinline__ void assemble_rewards(Arena& A, const Params& P, const Geom& g, const Kills& K,
                                                 const double (&opp_focus)[2], const bool (&present)[2],
                                                 double (&rew)[2]) {
  const double s = P.rew_scale;
  double rews[2] = {0.0, 0.0};
  bool destroyed[2] = {false, false};
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    if (A.alive[i] && !in_boundary(g, A.lat[i], A.lon[i])) {
      A.alive[i] = false;
      if (i < 2) {
        rews[i] += -5.0 * s;
        destroyed[i] = true;
        A.alive_ag -= 1;
      } else {
        A.alive_op -= 1;
      }
    }
  }
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const int k = K.killer[j];
    if (k == 0) continue;
    if (k <= 2) {
      const int ki = k - 1;
      if (j >= 2) {
        if (MODE == 0) {
          double r;
          if (K.by_rocket[j])
            r = (1.0 + 0.5 * ((double)A.mrem[0] / (double)A.rmax[0])) * s;  // only agent 1 has missiles
          else
            r = ((0.5 + 0.5 * ((double)(ki == 0 ? A.crem[0] : A.crem[1]) / (double)(ki == 0 ? A.cmax[0] : A.cmax[1]))) +
                 (0.5 + 0.5 * pick2(opp_focus, ki))) * s;
          put2(rews, ki, pick2(rews, ki) + r);
        }
        A.alive_op -= 1;
      } else {
        put2(rews, ki, pick2(rews, ki) + -2.0 * s);
        if (P.friendly_punish) {
          rews[j & 1] += -2.0 * s;
          destroyed[j & 1] = true;
        }
        A.alive_ag -= 1;
      }
    } else {
      if (j < 2) {
        rews[j & 1] += -2.0 * s;
        destroyed[j & 1] = true;
        A.alive_ag -= 1;
      } else {
        A.alive_op -= 1;
      }
    }
  }
  if (MODE == 1 && P.esc_dist_rew) {               // env_hetero.py:198-214
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      if (A.alive[i]) {
        double d2 = dist_raw(A.lat[i], A.lon[i], A.lat[2], A.lon[2]);
        double d3 = dist_raw(A.lat[i], A.lon[i], A.lat[3], A.lon[3]);
        double n2 = g.inv_diag * d2, n3 = g.inv_diag * d3;
        bool has2 = A.alive[2], has3 = A.alive[3];
        bool swap = has2 && has3 && n3 < n2;
        double first = has2 ? (swap ? d3 : d2) : d3;
        double second = swap ? d2 : d3;
        int n = (int)has2 + (int)has3;
#pragma unroll
        for (int j = 1; j <= 2; ++j) {
          if (j > n) break;
          double o2 = j == 1 ? first : second;
          if (o2 < 0.06) {
            rews[i] += -0.02 / j;
            if (A.spd[i] < 200.0) rews[i] += -0.02 / j;
          } else if (o2 > 0.13) {
            rews[i] += 0.02 / j;
            if (A.spd[i] > 500.0) rews[i] += 0.02 / j;
          }
        }
      }
    }
  }
#pragma unroll
  for (int i = 0; i < 2; ++i) {
    if (present[i] && (A.alive[i] || destroyed[i])) {
      if (P.glob_frac > 0.0 && MODE == 0)
        rew[i] += rews[i] + P.glob_frac * rews[1 - i];
      else
        rew[i] += rews[i];
    }
  }
}

// ------------------------------------------------------------------------------------- observations
__device__ __forceinline__ double hdg_feature(double heading) { return clip(pymod(heading, 359.0) / 359.0, 0.0, 1.0); }
__device__ __forceinline__ bool shot_flag(const Arena& A, int u) {
  bool shot = pick4(A.burst, u) > 0;
  if ((u & 1) == 0) shot = shot || pick4(A.hasm, u);
  return shot;
}

// friendly_ac_values (env_base.py:166-183): 5 values
__device__ __forceinline__ void friend_block(const Arena& A, const Geom& g, const HVec (&hv)[4], int self_u,
                                             int fri, float* out) {
  if (pick4(A.alive, fri)) {
    double x, y;
    double lat_s = pick4(A.lat, self_u), lon_s = pick4(A.lon, self_u);
    double lat_f = pick4(A.lat, fri), lon_f = pick4(A.lon, fri);
    rel_pos(g, lat_f, lon_f, x, y);
    out[0] = (float)x;
    out[1] = (float)y;
    out[2] = (float)focus_norm_from_deg(focus_deg(pick4(hv, self_u), lat_s, lon_s, lat_f, lon_f));
    out[3] = (float)focus_norm_from_deg(focus_deg(pick4(hv, fri), lat_f, lon_f, lat_s, lon_s));
    out[4] = (float)(g.inv_diag * dist_raw(lat_s, lon_s, lat_f, lon_f));
  } else {
#pragma unroll
    for (int k = 0; k < 5; ++k) out[k] = 0.0f;
  }
}

// lowlevel_state (env_hetero.py:65-103) for unit U in observation mode OMODE (0 fight, 1 esc).
// Writes obs_len floats to out; updates opp_to_attack[U].
template <int U, int OMODE>
__device__ __forceinline__ void unit_observation(Arena& A, const Geom& g, const HVec (&hv)[4], float* out) {
  constexpr int LEN = OMODE == 0 ? (is_ac1(U) ? OBS_AC1 : OBS_AC2) : (is_ac1(U) ? OBS_ESC_AC1 : OBS_ESC_AC2);
  constexpr int FRI = U ^ 1;  // fri_ac_id: 1<->2, 3<->4
  A.ota[U] = 0;
  double dn;
  int o = A.alive[U] ? nearest_enemy(A, g, U, dn) : -1;
  if (o < 0) {
#pragma unroll
    for (int k = 0; k < LEN; ++k) out[k] = 0.0f;
    return;
  }
  A.ota[U] = o + 1;
  int n = 0;
  double x, y;
  rel_pos(g, A.lat[U], A.lon[U], x, y);
  out[n++] = (float)x;
  out[n++] = (float)y;
  out[n++] = (float)clip(A.spd[U] / max_speed(U), 0.0, 1.0);
  out[n++] = (float)hdg_feature(A.hdg[U]);
  if (OMODE == 0) {  // fight_state_values, env_base.py:111-135
    double lat_o = pick4(A.lat, o), lon_o = pick4(A.lon, o);
    HVec ho = pick4(hv, o);
    double f_so = focus_deg(hv[U], A.lat[U], A.lon[U], lat_o, lon_o);
    double f_os = focus_deg(ho, lat_o, lon_o, A.lat[U], A.lon[U]);
    double hd = hdiff_norm(hv[U], ho);
    out[n++] = (float)focus_norm_from_deg(f_so);
    out[n++] = (float)aspect_from_deg(f_os);
    out[n++] = (float)hd;
    out[n++] = (float)dn;
    out[n++] = (float)clip((double)A.crem[U] / (double)A.cmax[U], 0.0, 1.0);
    if (is_ac1(U)) {
      out[n++] = (float)clip((double)A.mrem[U] / (double)A.rmax[U], 0.0, 1.0);
      out[n++] = A.mwait[U] == 0 ? 1.0f : 0.0f;
      out[n++] = (A.hasm[U] || A.burst[U] > 0) ? 1.0f : 0.0f;
    } else {
      out[n++] = A.burst[U] > 0 ? 1.0f : 0.0f;
    }
    // opp_ac_values("fight"), env_base.py:185-212
    rel_pos(g, lat_o, lon_o, x, y);
    out[n++] = (float)x;
    out[n++] = (float)y;
    out[n++] = (float)clip(pick4(A.spd, o) / ((o & 1) == 0 ? 900.0 : 600.0), 0.0, 1.0);
    out[n++] = (float)hdg_feature(pick4(A.hdg, o));
    out[n++] = (float)hdiff_norm(ho, hv[U]);
    out[n++] = (float)focus_norm_from_deg(f_os);
    out[n++] = (float)aspect_from_deg(f_so);
    out[n++] = (float)dn;
    out[n++] = shot_flag(A, o) ? 1.0f : 0.0f;
  } else {  // esc_state_values, env_base.py:137-164
    out[n++] = (float)clip((double)A.crem[U] / (double)A.cmax[U], 0.0, 1.0);
    if (is_ac1(U)) out[n++] = (float)clip((double)A.mrem[U] / (double)A.rmax[U], 0.0, 1.0);
    out[n++] = shot_flag(A, U) ? 1.0f : 0.0f;
    const int e0 = U < 2 ? 2 : 0;
    const int o2 = (o == e0) ? e0 + 1 : e0;  // the other enemy
    const bool has2 = pick4(A.alive, o2);
#pragma unroll
    for (int k = 0; k < 2; ++k) {
      int q = k == 0 ? o : o2;
      if (k == 1 && !has2) {
#pragma unroll
        for (int z = 0; z < 9; ++z) out[n++] = 0.0f;
        break;
      }
      double lat_q = pick4(A.lat, q), lon_q = pick4(A.lon, q);
      HVec hq = pick4(hv, q);
      double dq = k == 0 ? dn : g.inv_diag * dist_raw(A.lat[U], A.lon[U], lat_q, lon_q);
      rel_pos(g, lat_q, lon_q, x, y);
      out[n++] = (float)x;
      out[n++] = (float)y;
      out[n++] = (float)clip(pick4(A.spd, q) / ((q & 1) == 0 ? 900.0 : 600.0), 0.0, 1.0);
      out[n++] = (float)hdg_feature(pick4(A.hdg, q));
      out[n++] = (float)hdiff_norm(hq, hv[U]);
      out[n++] = (float)focus_norm_from_deg(focus_deg(hv[U], A.lat[U], A.lon[U], lat_q, lon_q));
      out[n++] = (float)focus_norm_from_deg(focus_deg(hq, lat_q, lon_q, A.lat[U], A.lon[U]));
      out[n++] = (float)dq;
      out[n++] = shot_flag(A, q) ? 1.0f : 0.0f;
    }
  }
  friend_block(A, g, hv, U, FRI, out + n);
}

// ------------------------------------------------------------------------------------- reset
// _sample_state (env_base.py:489-549): (lon, lat, heading) for group (0 agent / 1 opp), slot i, side r
__device__ __forceinline__ void sample_state(Arena& A, const Rng& rng, int level, int group, int i, int r,
                                             double& x, double& y, int& a) {
  a = 0;
  const double di = (double)i * 0.1;
  // x-range and heading-range tables; a "near" box (west) and a "far" box (east) per level
  double xw0, xw1, xe0, xe1, y0, y1;
  if (level == 1) { xw0 = 7.12; xw1 = 7.14; xe0 = 7.16; xe1 = 7.17; y0 = 5.1; y1 = 5.11; }
  else if (level == 2) { xw0 = 7.08; xw1 = 7.13; xe0 = 7.18; xe1 = 7.23; y0 = 5.08; y1 = 5.13; }
  else { xw0 = 7.07; xw1 = 7.12; xe0 = 7.18; xe1 = 7.23; y0 = 5.09; y1 = 5.12; }
  const bool west = (group == 0) == (r == 1);  // agents start west when r == 1, opponents east
  x = west ? g_uniform(rng, A, xw0, xw1) : g_uniform(rng, A, xe0, xe1);
  y = g_uniform(rng, A, __dadd_rn(y0, di), __dadd_rn(y1, di));
  if (group == 0) {
    if (level == 1) a = r == 1 ? g_randint(rng, A, 30, 150) : g_randint(rng, A, 200, 330);
    else if (level == 2) a = r == 1 ? g_randint(rng, A, 0, 180) : g_randint(rng, A, 180, 359);
    else a = r == 1 ? g_randint(rng, A, 0, 270) : g_randint(rng, A, 90, 359);
  } else if (level >= 2) {
    a = g_randint(rng, A, 0, 359);
  }
}

// HHMARLBaseEnv.reset + _reset_scenario (env_base.py:62-77, 551-585) + LowLevelEnv.reset
// (env_hetero.py:53-60).  RNG counters persist across episodes.
__device__ __forceinline__ void reset_arena(Arena& A, const Rng& rng, const Params& P) {
  A.steps = 0;
  A.alive_ag = 0;
  A.alive_op = 0;
  A.escaping = false;
  A.esc_time = 0;
  A.next_id = 1;
  const int r = g_randint(rng, A, 1, 2);
#pragma unroll
  for (int u = 0; u < 4; ++u) {
    const int group = u >> 1, i = u & 1;
    double x, y;
    int a;
    sample_state(A, rng, P.level, group, i, r, x, y, a);
    A.lat[u] = y;
    A.lon[u] = x;
    A.hdg[u] = (double)a;
    A.nhdg[u] = (double)a;
    double sp = (P.level <= 2 && group == 1) ? 0.0 : 100.0;
    A.spd[u] = sp;
    A.nspd[u] = sp;
    int cm = 200, mr = is_ac1(u) ? 5 : 0;
    if (P.level <= 4 && group == 1) { cm = 400; if (is_ac1(u)) mr = 8; }
    else if (P.level == 5) { cm = 300; if (is_ac1(u)) mr = 6; }
    A.crem[u] = cm;
    A.cmax[u] = cm;
    A.burst[u] = 0;
    A.mrem[u] = mr;
    A.rmax[u] = mr;
    A.mwait[u] = 0;
    A.alive[u] = true;
    A.hasm[u] = false;
    A.ota[u] = 0;
    A.next_id += 1;
    if (group == 0) A.alive_ag += 1; else A.alive_op += 1;
  }
#pragma unroll
  for (int s = 0; s < 2; ++s) {
    A.rlat[s] = 0.0; A.rlon[s] = 0.0; A.rhdg[s] = 0.0; A.rnhdg[s] = 0.0;
    A.ralive[s] = false; A.rage[s] = 0; A.rtgt[s] = 0; A.rid[s] = 0;
  }
  A.pset = 0;
  A.opp_mode = 0;
  if (P.level == 5 && P.agent_mode == 0) {
    int k = g_randint(rng, A, 3, 5);
    A.pset = k;
    A.opp_mode = k == 5 ? 1 : 0;
  }
}

}  // namespace hh
