/*
 * oracle/philox.h -- TEST INFRASTRUCTURE (CPU oracle). Not part of the product path.
 *
 * RNG contract shared by the oracle, the Python shim that is injected into the reference
 * modules (oracle/ref_harness.py) and -- independently re-implemented -- the CUDA kernels.
 * The reference itself is not reproducible (global unseeded `random`), so
 * "identical seeds" is defined by this contract:
 *
 *   Philox4x32-10 (Salmon et al., SC'11), key = (seed_lo, seed_hi),
 *   counter = (draw_lo, draw_hi, arena_id, stream)      stream 0 = G (module-global
 *   `random` of env_base.py/env_hetero.py/ac1.py), stream 1 = C (`sim.rnd_gen`, the cannon
 *   lottery of ac1.py:112 / ac2.py:99).  One Philox block per draw, words 0 and 1 used:
 *   random()     = ((w0 >> 5) * 2^26 + (w1 >> 6)) * 2^-53      (CPython's genrand_res53 mapping)
 *   uniform(a,b) = a + (b - a) * random()                      (CPython random.uniform)
 *   randint(a,b) = a + (int)(random() * (b - a + 1))
 *   A draw is consumed only when the reference's control flow reaches the call.
 */
#ifndef HH_ORACLE_PHILOX_H
#define HH_ORACLE_PHILOX_H

#include <stdint.h>

typedef struct {
  uint32_t key[2];
  uint32_t arena;
  uint32_t stream;
  uint64_t draw;
} orc_rng_t;

static inline void orc_philox4x32_10(const uint32_t ctr_in[4], const uint32_t key_in[2],
                                     uint32_t out[4]) {
  uint32_t c0 = ctr_in[0], c1 = ctr_in[1], c2 = ctr_in[2], c3 = ctr_in[3];
  uint32_t k0 = key_in[0], k1 = key_in[1];
  int r;
  for (r = 0; r < 10; ++r) {
    uint64_t p0 = (uint64_t)0xD2511F53u * c0;
    uint64_t p1 = (uint64_t)0xCD9E8D57u * c2;
    uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0;
    uint32_t n1 = (uint32_t)p1;
    uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1;
    uint32_t n3 = (uint32_t)p0;
    c0 = n0; c1 = n1; c2 = n2; c3 = n3;
    k0 += 0x9E3779B9u;
    k1 += 0xBB67AE85u;
  }
  out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

static inline double orc_rng_random(orc_rng_t* s) {
  uint32_t ctr[4], out[4];
  ctr[0] = (uint32_t)s->draw;
  ctr[1] = (uint32_t)(s->draw >> 32);
  ctr[2] = s->arena;
  ctr[3] = s->stream;
  orc_philox4x32_10(ctr, s->key, out);
  s->draw += 1;
  return ((double)(out[0] >> 5) * 67108864.0 + (double)(out[1] >> 6)) * (1.0 / 9007199254740992.0);
}

static inline double orc_rng_uniform(orc_rng_t* s, double a, double b) {
  return a + (b - a) * orc_rng_random(s);
}

static inline int orc_rng_randint(orc_rng_t* s, int a, int b) {
  return a + (int)(orc_rng_random(s) * (double)(b - a + 1));
}

/* random.choices([0, 1], weights=[w0, w1], k=1)[0]: bisect(cum_weights, random() * total) */
static inline int orc_rng_choice2(orc_rng_t* s, double w0, double w1) {
  return orc_rng_random(s) * (w0 + w1) >= w0 ? 1 : 0;
}

#endif
