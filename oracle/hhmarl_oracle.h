/*
 * oracle/hhmarl_oracle.h -- TEST INFRASTRUCTURE (CPU oracle). Not part of the product path.
 *
 * Scalar, single-arena C restatement of the reference's low-level air-combat environment
 * (envs/env_hetero.py LowLevelEnv + envs/env_base.py HHMARLBaseEnv + warsim/simulator/ files).
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
 * may link or call this; the product (hhmarl_2d_b200/) never does.
 *
 * It is pinned against trajectories of the UNMODIFIED reference files run under stubs
 * (oracle/ref_harness.py -> tests/golden/ npz files, tests/test_oracle_vs_reference.py).
 */
#ifndef HH_ORACLE_ENV_H
#define HH_ORACLE_ENV_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ORC_MAX_AC 6     /* aircraft ids 1..total_num (2-vs-2: 4, 3-vs-3: 6) */
#define ORC_MAX_UNITS 96 /* aircraft + every rocket id handed out in one episode */

/* The `args` namespace fields the env reads (config.py:14-56). */
typedef struct {
  int32_t level;           /* 1..5 */
  int32_t agent_mode;      /* 0 = "fight", 1 = "escape" */
  int32_t horizon;
  int32_t num_agents;      /* 2 */
  int32_t num_opps;        /* 2 */
  int32_t esc_dist_rew;    /* bool */
  int32_t friendly_kill;   /* bool (friendly_check of the units) */
  int32_t friendly_punish; /* bool */
  double map_size;         /* 0.3 */
  double rew_scale;        /* 1 */
  double glob_frac;        /* 0 */
  /* HighLevelEnv only (config.py:23,44) */
  int32_t hier_action_assess;   /* bool, default True */
  int32_t hier_opp_fight_ratio; /* percent, default 75 */
} orc_args_t;

typedef struct orc_env orc_env_t;

/* Frozen-policy opponents (env_base.py:349-398): called with the opponent's own observation;
 * must fill `action` (4 ints for ac_type 1, 3 for ac_type 2) with the per-head argmax. */
typedef void (*orc_policy_fn)(void* user, int unit_id, int ac_type, int policy_mode /*0 fight,1 esc*/,
                              int policy_set /* L5: k in 3..5, else 0 */, const float* obs,
                              int obs_len, int32_t* action);

orc_env_t* orc_env_create(const orc_args_t* args, uint64_t seed, uint32_t arena_id);
void orc_env_destroy(orc_env_t* e);
void orc_env_set_policy_fn(orc_env_t* e, orc_policy_fn fn, void* user);

/* env_hetero.py:53-60.  obs1/obs2: float32 observation of agents 1 and 2 (26/24 or 30/29). */
void orc_env_reset(orc_env_t* e, float* obs1, float* obs2);

/* env_base.py:79-109.  actions: int32[2][4] (agent 2 uses 3 entries).
 * rew[2] is 0 for an agent that had no entry in the reference's reward dict. Returns done. */
int orc_env_step(orc_env_t* e, const int32_t* actions, float* obs1, float* obs2, double* rew,
                 int32_t* rew_present);

/* ---- state inspection for parity tests (aircraft ids 1..4 -> index 0..3) */
typedef struct {
  double lat[ORC_MAX_AC], lon[ORC_MAX_AC], heading[ORC_MAX_AC], speed[ORC_MAX_AC];
  double new_heading[ORC_MAX_AC], new_speed[ORC_MAX_AC];
  double cannon_remain[ORC_MAX_AC], cannon_burst[ORC_MAX_AC], cannon_max[ORC_MAX_AC];
  int32_t missile_remain[ORC_MAX_AC], rocket_max[ORC_MAX_AC], missile_wait[ORC_MAX_AC];
  int32_t alive[ORC_MAX_AC], has_missile[ORC_MAX_AC], opp_to_attack[ORC_MAX_AC];
  int32_t ac_type[ORC_MAX_AC];
  /* the live rocket referenced by actual_missile of each aircraft (if any) */
  double r_lat[ORC_MAX_AC], r_lon[ORC_MAX_AC], r_heading[ORC_MAX_AC], r_new_heading[ORC_MAX_AC];
  double r_speed[ORC_MAX_AC];
  int32_t r_alive[ORC_MAX_AC], r_target[ORC_MAX_AC], r_id[ORC_MAX_AC], r_age[ORC_MAX_AC];
  int32_t steps, alive_agents, alive_opps, escaping, escaping_time, next_unit_id, opp_mode;
  int32_t policy_set, error;
  uint64_t draws_g, draws_c;
} orc_state_t;

void orc_env_get_state(const orc_env_t* e, orc_state_t* out);
int orc_obs_len(const orc_args_t* args, int agent_id);

/* ---- HighLevelEnv (envs/env_hier.py): 3-vs-3 commander environment.  obs: float32 [num_agents][34];
 * commander_actions: int32[num_agents] in {0,1,2}; rew: double[num_agents] (every agent has an entry);
 * info[0] = number of low-level sub-steps taken, info[1..total_num] = commander action of unit id (after
 * _action_assess, -1 for None).  The policy callback is used for ALL aircraft (env_hier.py:126-130). */
#define ORC_OBS_HL 34
void orc_hier_reset(orc_env_t* e, float* obs);
int orc_hier_step(orc_env_t* e, const int32_t* commander_actions, float* obs, double* rew, int32_t* info);
/* env_base.py:91-107 (args.eval_info): out[12] = agents_win, opps_win, draw, agent_fight, agent_escape, opp_fight,
 * opp_escape, agent_steps, opp_steps, opp1, opp2, opp3 for the step orc_hier_step just made (call before a reset). */
#define ORC_EVAL_INFO_LEN 12
void orc_hier_eval_info(const orc_env_t* e, int32_t* out);

/* Throughput helper for bench.py's CPU baseline: runs `n_steps` env steps with auto-reset and
 * uniformly random MultiDiscrete actions (own xorshift stream, not the contract RNG). */
uint64_t orc_env_run_random(orc_env_t* e, uint64_t n_steps, uint64_t action_seed);

#ifdef __cplusplus
}
#endif
#endif
