// hh_state_pack.h -- host-side conversion between the packed struct-of-arrays words of StatePtrs (hh_quad.cuh:
// load_lane / store_lane) and the plain per-field arrays of hh_state_view (include/hhmarl_b200.h).  Used by
// hh_get_state / hh_set_state (hh_api.cu) and by the CPU emulation harness of the step schedule (tests/emu).
#pragma once
#include <stddef.h>
#include <stdint.h>

#include "../../include/hhmarl_b200.h"

namespace hh {

inline void unpack_state(size_t N, const uint2* ac, const uint32_t* ri, const uint4* meta, const unsigned long long* dg,
                         hh_state_view* o) {
  struct { const uint2* ac; const uint32_t* ri; const uint4* meta; const unsigned long long* dg; } h{ac, ri, meta, dg};
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
}

inline void pack_state(size_t N, const hh_state_view* in, uint2* ac, uint32_t* ri, uint4* meta, unsigned long long* dg) {
  struct { uint2* ac; uint32_t* ri; uint4* meta; unsigned long long* dg; } h{ac, ri, meta, dg};
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
}

}  // namespace hh
