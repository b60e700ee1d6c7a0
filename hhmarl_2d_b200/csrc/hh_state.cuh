// hh_state.cuh -- struct-of-arrays arena state in HBM and its register image.
//
// Layout (2-vs-2; aircraft index u = id-1, ids 1,2 = agents AC1/AC2, ids 3,4 = opponents
// AC1/AC2, env_base.py:556-560; rocket slot s = u/2 of the AC1 shooter -- at most one live
// missile per shooter, ac1.py:73):
//
//   f64  lat/lon/hdg/spd/nhdg/nspd [N*4]   aircraft kinematics        6 x 32 B / arena
//   u32x2 acint                    [N*4]   packed aircraft integers       32 B / arena
//   f64  rlat/rlon/rhdg/rnhdg      [N*2]   rocket kinematics          4 x 16 B / arena
//   u32  rint                      [N*2]   packed rocket integers          8 B / arena
//   u32x4 meta                     [N]     arena scalars + C-stream counter 16 B / arena
//   u64  draws_g                   [N]     G-stream draw counter            8 B / arena
//                                                                 total  320 B / arena
// One thread owns one arena; a warp's accesses to every array are contiguous (32 B, 16 B or
// 8 B per lane), so every load/store instruction is fully coalesced.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace hh {

struct StatePtrs {
  double *lat, *lon, *hdg, *spd, *nhdg, *nspd;  // [N*4]
  uint2* acint;                                  // [N*4]
  double *rlat, *rlon, *rhdg, *rnhdg;            // [N*2]
  uint32_t* rint;                                // [N*2]
  uint4* meta;                                   // [N]
  unsigned long long* draws_g;                   // [N]
};

struct Params {
  int n_arenas;
  int level;       // 1..5
  int agent_mode;  // 0 fight, 1 escape
  int horizon;
  int esc_dist_rew, friendly_kill, friendly_punish, autoreset;
  double map_size, rew_scale, glob_frac;
  uint32_t seed_lo, seed_hi;
  uint32_t arena_base;  // global id of local arena 0 (multi-GPU sharding)
};

// error bits (where the reference would raise)
enum : int { ERR_HEADING = 1, ERR_SPEED = 2 };

struct Arena {
  double lat[4], lon[4], hdg[4], spd[4], nhdg[4], nspd[4];
  int crem[4], burst[4], cmax[4], mrem[4], rmax[4], mwait[4];
  bool alive[4], hasm[4];
  double rlat[2], rlon[2], rhdg[2], rnhdg[2];
  bool ralive[2];
  int rage[2], rtgt[2], rid[2];
  int steps, alive_ag, alive_op, esc_time, next_id, pset, opp_mode, err;
  bool escaping;
  int ota[4];  // opp_to_attack: 0 = None, else id 1..4
  unsigned long long dg;
  unsigned int dc;
};

template <typename T>
__device__ __forceinline__ T pick4(const T (&a)[4], int i) {
  return i == 0 ? a[0] : (i == 1 ? a[1] : (i == 2 ? a[2] : a[3]));
}
template <typename T>
__device__ __forceinline__ T pick2(const T (&a)[2], int i) {
  return i == 0 ? a[0] : a[1];
}
template <typename T>
__device__ __forceinline__ void put4(T (&a)[4], int i, T v) {
  a[0] = i == 0 ? v : a[0];
  a[1] = i == 1 ? v : a[1];
  a[2] = i == 2 ? v : a[2];
  a[3] = i == 3 ? v : a[3];
}
template <typename T>
__device__ __forceinline__ void put2(T (&a)[2], int i, T v) {
  a[0] = i == 0 ? v : a[0];
  a[1] = i == 1 ? v : a[1];
}

__device__ __forceinline__ void load4(const double* p, int a, double (&o)[4]) {
  const double2* q = reinterpret_cast<const double2*>(p) + 2 * (size_t)a;
  double2 x = q[0], y = q[1];
  o[0] = x.x; o[1] = x.y; o[2] = y.x; o[3] = y.y;
}
__device__ __forceinline__ void store4(double* p, int a, const double (&o)[4]) {
  double2* q = reinterpret_cast<double2*>(p) + 2 * (size_t)a;
  q[0] = make_double2(o[0], o[1]);
  q[1] = make_double2(o[2], o[3]);
}
__device__ __forceinline__ void load2(const double* p, int a, double (&o)[2]) {
  double2 x = reinterpret_cast<const double2*>(p)[a];
  o[0] = x.x; o[1] = x.y;
}
__device__ __forceinline__ void store2(double* p, int a, const double (&o)[2]) {
  reinterpret_cast<double2*>(p)[a] = make_double2(o[0], o[1]);
}

__device__ __forceinline__ void load_arena(const StatePtrs& s, int a, Arena& A) {
  load4(s.lat, a, A.lat);
  load4(s.lon, a, A.lon);
  load4(s.hdg, a, A.hdg);
  load4(s.spd, a, A.spd);
  load4(s.nhdg, a, A.nhdg);
  load4(s.nspd, a, A.nspd);
  const uint4* ai = reinterpret_cast<const uint4*>(s.acint) + 2 * (size_t)a;
  uint4 p0 = ai[0], p1 = ai[1];
  uint32_t w0[4] = {p0.x, p0.z, p1.x, p1.z};
  uint32_t w1[4] = {p0.y, p0.w, p1.y, p1.w};
#pragma unroll
  for (int u = 0; u < 4; ++u) {
    A.crem[u] = w0[u] & 0xFFFF;
    A.burst[u] = (w0[u] >> 16) & 0xFF;
    A.mrem[u] = (w0[u] >> 24) & 0xFF;
    A.cmax[u] = w1[u] & 0xFFFF;
    A.mwait[u] = (w1[u] >> 16) & 0xFF;
    A.rmax[u] = (w1[u] >> 24) & 0xF;
    A.alive[u] = (w1[u] >> 28) & 1;
    A.hasm[u] = (w1[u] >> 29) & 1;
  }
  load2(s.rlat, a, A.rlat);
  load2(s.rlon, a, A.rlon);
  load2(s.rhdg, a, A.rhdg);
  load2(s.rnhdg, a, A.rnhdg);
  uint2 ri = reinterpret_cast<const uint2*>(s.rint)[a];
  uint32_t rw[2] = {ri.x, ri.y};
#pragma unroll
  for (int r = 0; r < 2; ++r) {
    A.ralive[r] = rw[r] & 1;
    A.rage[r] = (rw[r] >> 1) & 0xF;
    A.rtgt[r] = (rw[r] >> 5) & 0x7;
    A.rid[r] = (rw[r] >> 8) & 0xFFFF;
  }
  uint4 m = s.meta[a];
  A.steps = m.x & 0xFFFF;
  A.alive_ag = (m.x >> 16) & 0xF;
  A.alive_op = (m.x >> 20) & 0xF;
  A.escaping = (m.x >> 24) & 1;
  A.pset = (m.x >> 25) & 0x7;
  A.opp_mode = (m.x >> 28) & 1;
  A.esc_time = m.y & 0xFF;
  A.next_id = (m.y >> 8) & 0xFF;
  A.ota[0] = (m.y >> 24) & 0x3;  // agents attack ids 3/4 -> stored as id-2 (0 = None, 1 -> 3, 2 -> 4)
  A.ota[1] = (m.y >> 26) & 0x3;
  A.ota[2] = (m.y >> 28) & 0x3;  // opponents attack ids 1/2 (0 = None)
  A.ota[3] = (m.y >> 30) & 0x3;
  A.ota[0] = A.ota[0] ? A.ota[0] + 2 : 0;
  A.ota[1] = A.ota[1] ? A.ota[1] + 2 : 0;
  A.dc = m.z;
  A.err = m.w;
  A.dg = s.draws_g[a];
}

__device__ __forceinline__ void store_arena(const StatePtrs& s, int a, const Arena& A) {
  store4(s.lat, a, A.lat);
  store4(s.lon, a, A.lon);
  store4(s.hdg, a, A.hdg);
  store4(s.spd, a, A.spd);
  store4(s.nhdg, a, A.nhdg);
  store4(s.nspd, a, A.nspd);
  uint32_t w0[4], w1[4];
#pragma unroll
  for (int u = 0; u < 4; ++u) {
    w0[u] = (uint32_t)A.crem[u] | ((uint32_t)A.burst[u] << 16) | ((uint32_t)A.mrem[u] << 24);
    w1[u] = (uint32_t)A.cmax[u] | ((uint32_t)A.mwait[u] << 16) | ((uint32_t)A.rmax[u] << 24) |
            ((uint32_t)A.alive[u] << 28) | ((uint32_t)A.hasm[u] << 29);
  }
  uint4* ai = reinterpret_cast<uint4*>(s.acint) + 2 * (size_t)a;
  ai[0] = make_uint4(w0[0], w1[0], w0[1], w1[1]);
  ai[1] = make_uint4(w0[2], w1[2], w0[3], w1[3]);
  store2(s.rlat, a, A.rlat);
  store2(s.rlon, a, A.rlon);
  store2(s.rhdg, a, A.rhdg);
  store2(s.rnhdg, a, A.rnhdg);
  uint32_t rw[2];
#pragma unroll
  for (int r = 0; r < 2; ++r)
    rw[r] = (uint32_t)A.ralive[r] | ((uint32_t)A.rage[r] << 1) | ((uint32_t)A.rtgt[r] << 5) |
            ((uint32_t)A.rid[r] << 8);
  reinterpret_cast<uint2*>(s.rint)[a] = make_uint2(rw[0], rw[1]);
  uint4 m;
  m.x = (uint32_t)A.steps | ((uint32_t)A.alive_ag << 16) | ((uint32_t)A.alive_op << 20) |
        ((uint32_t)A.escaping << 24) | ((uint32_t)A.pset << 25) | ((uint32_t)A.opp_mode << 28);
  m.y = (uint32_t)A.esc_time | ((uint32_t)A.next_id << 8) |
        ((uint32_t)(A.ota[0] ? A.ota[0] - 2 : 0) << 24) | ((uint32_t)(A.ota[1] ? A.ota[1] - 2 : 0) << 26) |
        ((uint32_t)A.ota[2] << 28) | ((uint32_t)A.ota[3] << 30);
  m.z = A.dc;
  m.w = (uint32_t)A.err;
  s.meta[a] = m;
  s.draws_g[a] = A.dg;
}

}  // namespace hh
