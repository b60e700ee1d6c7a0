// hh_core.cuh -- scalar building blocks shared by the kernels: RNG contract, Python-semantics
// arithmetic helpers, map geometry and the flat-plane angle features of env_base.py.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "hh_geodesic.cuh"

namespace hh {

constexpr double kKnotsToMs = 0.514444;  // cmano_simulator.py:21
constexpr int OBS_AC1 = 26, OBS_AC2 = 24, OBS_ESC_AC1 = 30, OBS_ESC_AC2 = 29;

// error bits (where the reference would raise)
enum : int { ERR_HEADING = 1, ERR_SPEED = 2 };

struct Geom {  // per-launch constants of the map (env_base.py:43, map_limits.py)
  double ext_lat, ext_lon, top, right;
  double inv_diag;  // (1 - 0) / (sqrt(2 ms^2) - 0)   (env_base.py:439,458-462)
};
__host__ __device__ __forceinline__ Geom make_geom(double ms) {
  Geom g;
  g.top = 5.0 + ms;
  g.right = 7.0 + ms;
  g.ext_lat = g.top - 5.0;
  g.ext_lon = g.right - 7.0;
  g.inv_diag = 1.0 / sqrt(2.0 * (ms * ms));
  return g;
}

struct Params {
  int n_arenas;
  int level;       // 1..5
  int agent_mode;  // 0 fight, 1 escape
  int horizon;
  int esc_dist_rew, friendly_kill, friendly_punish, autoreset;
  double map_size, rew_scale, glob_frac;
  uint32_t seed_lo, seed_hi;
  uint32_t arena_base;  // global id of local arena 0 (multi-GPU sharding)
  Geom geom;            // make_geom(map_size), filled on the host (IEEE sqrt / division: same bits as on the device)
  int short_moves;      // levels 4/5 finish kernel: 1 = geo::direct_tick (default), 0 = full Karney direct (HH_SHORT_MOVES=0)
};

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
// One block per draw; random() = ((w0 >> 5) * 2^26 + (w1 >> 6)) * 2^-53.
static __device__ __noinline__ double philox_u53(uint32_t k0, uint32_t k1, uint32_t c0, uint32_t c1, uint32_t c2,
                                          uint32_t c3) {
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
// stream G (module-global `random` of the reference) at an explicit draw index
__device__ __forceinline__ double g_random_at(const Rng& r, unsigned long long idx) {
  return philox_u53(r.k0, r.k1, (uint32_t)idx, (uint32_t)(idx >> 32), r.arena, 0u);
}
__device__ __forceinline__ double uniform_from(double a, double b, double rnd) {
  return __dadd_rn(a, __dmul_rn(b - a, rnd));  // a + (b-a)*random(), never fused (CPython semantics)
}
__device__ __forceinline__ int randint_from(int a, int b, double rnd) {
  return a + (int)(rnd * (double)(b - a + 1));
}
// stream C (`sim.rnd_gen`, cannon lottery)
__device__ __forceinline__ double c_random_at(const Rng& r, unsigned int idx) {
  return philox_u53(r.k0, r.k1, idx, 0u, r.arena, 1u);
}

// ------------------------------------------------------------------------------------- scalar helpers
// C fmod(x, m) for the moduli of this code (m = 360 or 359: q m is exact for every quotient below 2^44) and
// |x| < 1e9: r = x - trunc(x / m) m in one fma, which is exact because r is representable; the quotient estimated
// through the reciprocal can be off by one next to a multiple of m, which the compares repair (again exactly).
// libdevice's fmod is a ~60-instruction loop and was 4 % of the step's stall samples (profiles/r1i_*).
__device__ __forceinline__ double fmod_small(double x, double m) {
  if (!(fabs(x) < 1e9)) return m::fmod_(x, m);
  const double q = trunc(x * (1.0 / m));
  double r = fma(-q, m, x);
  if (x >= 0.0) {
    if (r < 0.0) r += m;
    else if (r >= m) r -= m;
  } else {
    if (r > 0.0) r -= m;
    else if (r <= -m) r += m;
  }
  return r;
}
__device__ __forceinline__ double pymod(double x, double m) {  // CPython float %, m > 0
  double r = fmod_small(x, m);
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
// angles.py:10-15
__device__ __forceinline__ double normalize_angle(double a) {
  while (a >= 360.0) a -= 360.0;
  while (a < 0.0) a += 360.0;
  return a;
}
__device__ __forceinline__ bool is_ac1(int u) { return (u & 1) == 0; }
__device__ __forceinline__ double max_speed(int u) { return is_ac1(u) ? 900.0 : 600.0; }

// map_limits.py:37-40
__device__ __forceinline__ void rel_pos(const Geom& g, double lat, double lon, double& lat_rel, double& lon_rel) {
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
  m::sincos_(th, &h.s, &h.c);
  h.n = sqrt(h.c * h.c + h.s * h.s);
  return h;
}
// env_base.py:424-432 -- degrees
__device__ __forceinline__ double focus_deg(const HVec& ha, double lat_a, double lon_a, double lat_b, double lon_b) {
  double v0 = lon_b - lon_a, v1 = lat_b - lat_a;
  double x = clip((ha.c * v0 + ha.s * v1) / (ha.n * sqrt(v0 * v0 + v1 * v1) + 1e-10), -1.0, 1.0);
  return m::acos_(x) * (180.0 / geo::kPi);
}
__device__ __forceinline__ double focus_norm_from_deg(double deg) { return clip(deg / 180.0, 0.0, 1.0); }
// env_base.py:441-446
__device__ __forceinline__ double aspect_from_deg(double deg) { return clip((180.0 - deg) / 180.0, 0.0, 1.0); }
// env_base.py:448-456
__device__ __forceinline__ double hdiff_norm(const HVec& a, const HVec& b) {
  double x = clip((a.c * b.c + a.s * b.s) / (a.n * b.n + 1e-10), -1.0, 1.0);
  return clip((m::acos_(x) * (180.0 / geo::kPi)) / 180.0, 0.0, 1.0);
}
// env_base.py:434-439
__device__ __forceinline__ double dist_raw(double lat_a, double lon_a, double lat_b, double lon_b) {
  const double dx = lon_b - lon_a, dy = lat_b - lat_a;  // math.hypot of two O(0.1) numbers
  return sqrt(dx * dx + dy * dy);
}
__device__ __forceinline__ double hdg_feature(double heading) {
  return clip(pymod(heading, 359.0) / 359.0, 0.0, 1.0);
}

// ac1.py:144-146
__device__ __forceinline__ bool angle_in_radar_range(double heading, double angle) {
  double c = normalize_angle(heading + 60.0);  // sum_angles(heading, 120/2)
  double delta = fabs(signed_heading_diff(c, angle));
  return (int)delta <= 60;
}

// env_base.py:464-487
__device__ __forceinline__ int correct_angle_sign(double lat_o, double lon_o, double hdg_o, double lat_a,
                                                  double lon_a) {
  double s, c;
  m::sincos_(pymod(hdg_o, 360.0) * (geo::kPi / 180.0), &s, &c);
  double x1 = lon_o + rint(s * 1000.0) / 1000.0;  // round(sin, 3)
  double y1 = lat_o + rint(c * 1000.0) / 1000.0;
  double val = (x1 - lon_o) * (lat_a - lat_o) - (lon_a - lon_o) * (y1 - lat_o);
  return val < 0.0 ? 1 : -1;
}

// the two halves of correct_angle_sign: the heading's unit offsets rounded to 3 decimals (depends on the opponent
// only), and the cross product with the line of sight (never fused: CPython evaluates a*b - c*d in three roundings)
__device__ __forceinline__ void angle_sign_offsets(double hdg_o, double& sx, double& cx) {
  double s, c;
  m::sincos_(pymod(hdg_o, 360.0) * (geo::kPi / 180.0), &s, &c);
  sx = rint(s * 1000.0) / 1000.0;
  cx = rint(c * 1000.0) / 1000.0;
}
__device__ __forceinline__ int angle_sign_from(double lat_o, double lon_o, double sx, double cx, double lat_a,
                                               double lon_a) {
  const double x1 = lon_o + sx, y1 = lat_o + cx;
  const double val = __dadd_rn(__dmul_rn(x1 - lon_o, lat_a - lat_o), -__dmul_rn(lon_a - lon_o, y1 - lat_o));
  return val < 0.0 ? 1 : -1;
}

// ------------------------------------------------------------------------------------- range tests
// The reference decides cannon hits and rocket proximity from the WGS84 geodesic distance
// (units_distance_km, cmano_simulator.py:167-169).  On the ellipsoid the metric satisfies
//   ds >= 109.6 km/deg * hypot(dlat, dlon)   for |lat| <= 10 deg
// (meridional arc >= 110.574 km/deg everywhere, parallel arc >= 111.320*cos(10 deg) = 109.63 km/deg),
// so a pair whose flat separation already exceeds range / 109.6 deg is out of range with
// certainty and the ~2.5 k-instruction inverse solve is skipped; everything closer goes through
// the exact solve.  The Boolean outcome is therefore identical to the reference's.
constexpr double kKmPerDegLower = 109.6;

// Decision margins for the local solution (see geo::inverse_local): inside them the exact solver
// decides, outside them the local solution provably agrees with it.
constexpr double kMarginDistM = 1e-3;     // 1 mm  (local error <= 10 um up to 7 km)
constexpr double kMarginAziDeg = 1e-5;    //       (local error <= 3e-8 deg up to 7 km)
constexpr double kMarginAziFarDeg = 1e-3; //       (local error <= 1e-5 deg up to 80 km)

// Cheap superset of the pairs that pass the flat prefilter of unit_in_cannon_range / within_1km (squared distance, a
// hair wider): a pair outside it is out of range with certainty, a pair inside it goes through the full test.
__device__ __forceinline__ bool maybe_within_km(double lat_s, double lon_s, double lat_t, double lon_t, double range_km) {
  const double dx = lon_t - lon_s, dy = lat_t - lat_s;
  const double r = range_km * (1.000001 / kKmPerDegLower);
  return dx * dx + dy * dy < r * r;
}
// ac1.py:135-142 / ac2.py:109-116
__device__ __forceinline__ bool unit_in_cannon_range(double lat_s, double lon_s, double hdg_s, double lat_t,
                                                     double lon_t, double range_km, double half_width) {
  const double dx = lon_t - lon_s, dy = lat_t - lat_s;
  if (sqrt(dx * dx + dy * dy) * kKmPerDegLower >= range_km) return false;
  const double range_m = range_km * 1000.0;
  double2 inv = geo::inverse_local(lat_s, lon_s, lat_t, lon_t);
  bool ambiguous = fabs(inv.x - range_m) < kMarginDistM;
  if (!ambiguous) {
    if (inv.x >= range_m) return false;
    const double delta = fabs(signed_heading_diff(hdg_s, normalize_angle(inv.y)));
    if (fabs(delta - half_width) >= kMarginAziDeg) return delta <= half_width;
  }
  inv = geo::inverse(lat_s, lon_s, lat_t, lon_t);  // razor-thin shell around a threshold: exact
  if (inv.x / 1000.0 < range_km) {
    const double delta = fabs(signed_heading_diff(hdg_s, normalize_angle(inv.y)));
    return delta <= half_width;
  }
  return false;
}
// rocket_unit.py:39,49: units_distance_km(self, x) < 1
__device__ __forceinline__ bool within_1km(double lat_r, double lon_r, double lat_t, double lon_t) {
  const double dx = lon_t - lon_r, dy = lat_t - lat_r;
  if (sqrt(dx * dx + dy * dy) * kKmPerDegLower >= 1.0) return false;
  const double s = geo::inverse_local(lat_r, lon_r, lat_t, lon_t).x;
  if (fabs(s - 1000.0) >= kMarginDistM) return s < 1000.0;
  return geo::inverse(lat_r, lon_r, lat_t, lon_t).x / 1000.0 < 1.0;
}
// Rafale.fire_missile's launch gate (ac1.py:75-76): distance <= 111 km and the skewed radar cone
__device__ __forceinline__ bool launch_gate(double lat_s, double lon_s, double hdg_s, double lat_t, double lon_t) {
  double2 inv = geo::inverse_local(lat_s, lon_s, lat_t, lon_t);
  const double c = normalize_angle(hdg_s + 60.0);  // sum_angles(heading, 120/2), ac1.py:145
  if (inv.x < 100000.0) {                           // certainly <= 111 km
    const double delta = fabs(signed_heading_diff(c, normalize_angle(inv.y)));
    if (fabs(delta - 61.0) >= kMarginAziFarDeg) return delta < 61.0;   // int(delta) <= 60
  }
  inv = geo::inverse(lat_s, lon_s, lat_t, lon_t);
  return inv.x / 1000.0 <= 111.0 && angle_in_radar_range(hdg_s, normalize_angle(inv.y));
}

}  // namespace hh
