/*
 * include/hhmarl_b200.h -- C ABI of the B200-native hhmarl_2D hot path.
 *
 * The reference (IDSIA/hhmarl_2D) is pure Python and has no FFI; the interface this library
 * replaces is the RLlib MultiAgentEnv surface of envs/env_hetero.py + envs/env_base.py:
 *
 *   hh_create      <- LowLevelEnv.__init__(env_config)            env_hetero.py:20-51
 *   hh_reset       <- LowLevelEnv.reset()                         env_hetero.py:53-60, env_base.py:62-77
 *   hh_step        <- HHMARLBaseEnv.step(action_dict)             env_base.py:79-109
 *                     (LowLevelEnv._take_action env_hetero.py:105-186, CmanoSimulator.do_tick
 *                      cmano_simulator.py:138-157, _get_rewards env_hetero.py:188-225,
 *                      lowlevel_state env_hetero.py:65-103) for N arenas in lock-step
 *   hh_step_host / hh_reset_host  same, host buffers in / out (the call an RLlib-style CPU
 *                     driver makes; copies are part of the call)
 *   hh_get_state / hh_set_state   (no reference counterpart: test / checkpoint access to the
 *                     struct-of-arrays arena state, incl. RNG counters)
 *
 * Conventions: plain pointers and sizes only; every function returns 0 on success or a
 * negative error code and records a message retrievable with hh_last_error(); no C++
 * exception crosses the boundary.  A handle is single-owner and stream-ordered: hh_step
 * enqueues on `stream` (a cudaStream_t passed as void*, NULL = default stream) and never
 * synchronises.  All *_dev pointers are device pointers owned by the caller.
 *
 * Shapes (N = n_arenas, row-major):
 *   actions  int32 [N][2][4]   agent 1: MultiDiscrete[13,9,2,2]; agent 2: [13,9,2] (4th ignored)
 *   obs1     f32   [N][hh_obs_dim(env,1)]   26 (fight) / 30 (escape)
 *   obs2     f32   [N][hh_obs_dim(env,2)]   24 (fight) / 29 (escape)
 *   rew      f32   [N][2]      0 for an agent that has no entry in the reference's reward dict
 *   done     u8    [N]         terminateds["__all__"] (== truncateds["__all__"], env_base.py:89-90)
 */
#ifndef HHMARL_B200_H
#define HHMARL_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct hh_env hh_env;

/* The `args` fields the env reads (config.py:14-56). */
typedef struct {
  int32_t level;           /* 1..3 scripted opponents (hh_step); 4, 5 frozen-policy opponents (hh_step_begin/finish) */
  int32_t agent_mode;      /* 0 = "fight", 1 = "escape" */
  int32_t horizon;         /* config.py:95 {1:150, 2:200, 3:300, 4:350, 5:400} */
  int32_t esc_dist_rew;    /* bool */
  int32_t friendly_kill;   /* bool */
  int32_t friendly_punish; /* bool */
  int32_t autoreset;       /* 1: an arena whose step returns done is reset inside the same launch and
                              obs are the first observation of the new episode (what RLlib's sampler
                              does after a terminal step); 0: caller resets (reference semantics) */
  int32_t reserved;
  double map_size;         /* 0.3 */
  double rew_scale;        /* 1 */
  double glob_frac;        /* 0 */
  uint64_t seed;           /* Philox key */
  uint64_t arena_base;     /* global id of local arena 0: results are invariant to how arenas are
                              partitioned over GPUs */
} hh_config;

/* Host-side, caller-allocated view of the complete state of all arenas (unpacked). */
typedef struct {
  double *lat, *lon, *heading, *speed, *new_heading, *new_speed;           /* [N*4] */
  int32_t *cannon_remain, *cannon_burst, *cannon_max;                      /* [N*4] */
  int32_t *missile_remain, *rocket_max, *missile_wait, *alive, *has_missile; /* [N*4] */
  int32_t *opp_to_attack;                                                  /* [N*4] 0 = None */
  double *r_lat, *r_lon, *r_heading, *r_new_heading;                       /* [N*2] slot = shooter id 1 / 3 */
  int32_t *r_alive, *r_age, *r_target, *r_id;                              /* [N*2] */
  int32_t *steps, *alive_agents, *alive_opps, *escaping, *escaping_time;   /* [N] */
  int32_t *next_unit_id, *policy_set, *opp_mode, *error;                   /* [N] */
  uint64_t *draws_g, *draws_c;                                             /* [N] */
} hh_state_view;

int hh_create(const hh_config* cfg, int32_t n_arenas, int32_t device, hh_env** out);
void hh_destroy(hh_env* env);

int32_t hh_n_arenas(const hh_env* env);
int32_t hh_obs_dim(const hh_env* env, int32_t agent_id /* 1 or 2 */);

/* mask_dev: u8[N] (non-zero = reset that arena) or NULL = all. obs pointers may be NULL. */
int hh_reset(hh_env* env, const uint8_t* mask_dev, float* obs1_dev, float* obs2_dev, void* stream);
int hh_step(hh_env* env, const int32_t* actions_dev, float* obs1_dev, float* obs2_dev, float* rew_dev,
            uint8_t* done_dev, void* stream);
/* The same for arenas [first, first + count) only (levels 1-3; `first` a multiple of 32).  All pointers address the whole batch's
 * arrays.  Arenas are independent (each env of the reference is its own process, train_hetero.py:212), so two ranges may be in
 * flight on two streams: the rollout sampler steps one half of the batch while the other half's policy forward runs. */
int hh_step_range(hh_env* env, int32_t first, int32_t count, const int32_t* actions_dev, float* obs1_dev, float* obs2_dev,
                  float* rew_dev, uint8_t* done_dev, void* stream);
/* hh_step_range that ALSO writes the new observations as the central-critic rows of both policies (central_critic_observer,
 * train_hetero.py:162-181; what hh_pack_central does in a launch of its own): central1[a][7 ..] = [obs1 | obs2],
 * central2[a][7 ..] = [obs2 | obs1], rows `ld` floats apart, the 7 action columns untouched.  obs1_dev / obs2_dev may be NULL. */
int hh_step_range_central(hh_env* env, int32_t first, int32_t count, const int32_t* actions_dev, float* obs1_dev, float* obs2_dev,
                          float* rew_dev, uint8_t* done_dev, float* central1_dev, float* central2_dev, int32_t ld, void* stream);

/* Levels 4/5 (frozen-policy opponents, env_base.py:349-398): the opponents' own observations are needed
 * mid-step, after the agents' fire decisions and before the tick, so the step is split around the
 * (caller-run, batched) opponent networks:
 *   hh_step_begin : agents' _take_base_action, then lowlevel_state(opp_mode, opp) of opponents 3 and 4
 *                   -> opp_obs3 f32[N][30], opp_obs4 f32[N][29] (fight mode fills the first 26 / 24 entries;
 *                   the per-arena mode / policy set are the `opp_mode` / `policy_set` fields of the state)
 *   hh_step_finish: opponents' _take_base_action with opp_actions int32[N][2][4] (per-head argmax of the
 *                   networks, env_base.py:373-382), tick, rewards, observations -- outputs as hh_step. */
int hh_step_begin(hh_env* env, const int32_t* actions_dev, float* opp_obs3_dev, float* opp_obs4_dev,
                  uint8_t* policy_set_dev /* u8[N] out, nullable: k of env_hetero.py:57, 0 below level 5 */, void* stream);
int hh_step_finish(hh_env* env, const int32_t* opp_actions_dev, float* obs1_dev, float* obs2_dev, float* rew_dev,
                   uint8_t* done_dev, void* stream);

/* Host-buffer variants: H2D of the actions, the launch, D2H of obs/rew/done and a stream
 * synchronise all happen inside the call (pinned staging owned by the handle). */
int hh_reset_host(hh_env* env, const uint8_t* mask_host, float* obs1_host, float* obs2_host);
int hh_step_host(hh_env* env, const int32_t* actions_host, float* obs1_host, float* obs2_host,
                 float* rew_host, uint8_t* done_host);

/* Pointers into the handle's own PINNED host slab (valid until hh_destroy): filling `actions` in place and
 * passing exactly these pointers to hh_step_host / hh_reset_host makes the call zero-copy on the host side
 * (one H2D of the actions, one D2H of obs1|obs2|rew|done, one stream synchronise). */
/* The two halves of hh_step_host, for callers that keep several handles in flight (RLlib's asynchronous env
 * interface BaseEnv.send_actions() / poll()): _begin enqueues the step on the handle's private stream and returns,
 * _end waits for it and delivers the results.  At most one step per handle may be pending. */
int hh_step_host_begin(hh_env* env, const int32_t* actions_host);
int hh_step_host_end(hh_env* env, float* obs1_host, float* obs2_host, float* rew_host, uint8_t* done_host);
int hh_host_buffers(hh_env* env, int32_t** actions, float** obs1, float** obs2, float** rew, uint8_t** done);
/* How hh_reset_host / hh_step_host move data: 1 = zero-copy (default: the kernels read the actions from and write the
 * results to the handle's pinned slab through its device mapping), 0 = staged (one H2D of the actions, the launch,
 * one D2H of the results; also selected by HH_HOST_MODE=staged).  Either way the call returns after the results are
 * in host memory. */
int hh_set_host_mode(hh_env* env, int32_t mode);

int hh_get_state(hh_env* env, hh_state_view* out_host);
int hh_set_state(hh_env* env, const hh_state_view* in_host);

/* Number of kernels this library has launched on behalf of `env` since creation. */
uint64_t hh_launch_count(const hh_env* env);

/* Generalised advantage estimation over a rollout fragment (what RLlib's postprocessing does per episode
 * before on_postprocess_trajectory, train_hetero.py:120-160; gamma 0.99, lambda 0.95 at :216).
 *   rew, vf: f32[T][N][2]; last_vf: f32[N][2] (value of the observation after the fragment);
 *   done: u8[T][N] (episode ended at that step: no bootstrap across it); adv, vtarg: f32[T][N][2] out. */
int hh_gae(int32_t T, int32_t n_arenas, const float* rew_dev, const float* vf_dev, const float* last_vf_dev,
           const uint8_t* done_dev, float gamma, float lam, float* adv_dev, float* vtarg_dev, void* stream);
/* same with n_agents learning agents per arena (3 for the commander policy of train_hier.py): [T][N][n_agents] */
int hh_gae_agents(int32_t T, int32_t n_arenas, int32_t n_agents, const float* rew_dev, const float* vf_dev,
                  const float* last_vf_dev, const uint8_t* done_dev, float gamma, float lam, float* adv_dev,
                  float* vtarg_dev, void* stream);

/* The learner's MultiCategorical terms of one policy (RLlib TorchMultiCategorical.logp / entropy / kl, summed over the MultiDiscrete
 * heads of widths[n_heads], e.g. {13, 9, 2, 2}): forward writes logp of the taken actions int32[N][ld_act], the entropy and
 * KL(old || new) per row; backward writes d / d logits f32[N][sum(widths)] (contiguous) from the three upstream gradients.
 * `widths` is a HOST array.  Used by PPOLearner's loss (hhmarl_2d_b200/sampler.py: multicategorical_logp_entropy_kl). */
int hh_multicat_forward(int32_t n_rows, int32_t n_heads, const int32_t* widths, const float* logits_dev, int32_t ld,
                        const float* old_logits_dev, int32_t ld_old, const int32_t* actions_dev, int32_t ld_act, float* logp_dev,
                        float* entropy_dev, float* kl_dev, void* stream);
int hh_multicat_backward(int32_t n_rows, int32_t n_heads, const int32_t* widths, const float* logits_dev, int32_t ld,
                         const float* old_logits_dev, int32_t ld_old, const int32_t* actions_dev, int32_t ld_act,
                         const float* g_logp_dev, const float* g_entropy_dev, const float* g_kl_dev, float* g_logits_dev, void* stream);

/* PPO's loss of one policy from its per-row terms (RLlib 2.4 ppo_torch_policy.loss): clipped surrogate of
 * ratio = exp(logp - old_logp) with advantages adv, kl_coeff (a DEVICE scalar: it adapts between updates) x kl, vf_coeff x the
 * value loss clamp((vf - vtarg)^2, 0, vf_clip), -ent_coeff x entropy.  sums[4] = sum over the rows of {loss, kl, value loss,
 * entropy}; deriv f32[4][N] = d (mean loss) / d {logp, entropy, kl, vf} per row (what PPOLearner's backward multiplies by the
 * upstream gradient). */
int hh_ppo_loss(int32_t n_rows, const float* logp_dev, const float* entropy_dev, const float* kl_dev, const float* vf_dev,
                const float* old_logp_dev, const float* adv_dev, const float* vtarg_dev, const float* kl_coeff_dev, float clip,
                float vf_clip, float vf_coeff, float ent_coeff, float* sums_dev, float* deriv_dev, void* stream);

/* The two ends of a rollout fragment in the sampler's central-critic layout (rows [7 action columns | own obs | other obs] of D
 * floats, flat f32[T][N][D] per policy):
 * hh_fragment_prepare  : zero the action columns of all rows (the critic sees zero actions while sampling) and copy the current
 *                        central observation rows cur1 / cur2 f32[N][D] into tick 0;
 * hh_fragment_writeback: CustomCallback.on_postprocess_trajectory (train_hetero.py:120-160): the action columns get the taken
 *                        actions int32[T][N][2][4], scaled a0 / 12, a1 / 8, a2, a3 (train_hetero.py:143-146). */
int hh_fragment_prepare(int32_t T, int32_t n_arenas, int32_t D, float* flat1_dev, float* flat2_dev, const float* cur1_dev,
                        const float* cur2_dev, void* stream);
int hh_fragment_writeback(int32_t T, int32_t n_arenas, int32_t D, const int32_t* actions_dev, float* flat1_dev, float* flat2_dev,
                          void* stream);

/* Sampler glue (what RLlib's sampler does between the policy forward and env.step):
 * hh_sample_actions: TorchMultiCategorical.sample()/logp() of both policies in one launch.  logits1 f32[N][26]
 *   (heads 13|9|2|2, agent 1), logits2 f32[N][24] (13|9|2, agent 2); counters u32[N][2] (per (arena, agent) draw
 *   counter, advanced by the call); explore = 0 takes the per-head argmax; actions int32[N][2][4]; logp f32[N][2].
 * hh_pack_central: central_critic_observer (train_hetero.py:162-181): writes obs1/obs2 into the observation
 *   columns of flat1 = [act(4)|act(3)|obs1|obs2] and flat2 = [act(3)|act(4)|obs2|obs1] (row length 7+d1+d2). */
int hh_sample_actions(int32_t n_arenas, const float* logits1_dev, const float* logits2_dev, uint64_t seed,
                      uint64_t arena_base, uint32_t* counters_dev, int32_t explore, int32_t* actions_dev, float* logp_dev,
                      void* stream);
int hh_pack_central(int32_t n_arenas, int32_t d1, int32_t d2, const float* obs1_dev, const float* obs2_dev,
                    float* flat1_dev, float* flat2_dev, void* stream);

/* ---- fused forward of the sampler's network chains (models/ac_models_hetero.py:86-103, 256-291, 368-404; the
 * Policy.compute_actions -> TorchModelV2.forward call of the rollout worker).
 * One chain = x [n_rows, d_in] -> tanh(x W1 + b1) [500] -> optional single-token attention block on columns
 * [att_lo, 500) (residual + L2 normalisation) -> tanh(. Ws + bs) with the shared 500x500 layer -> head [n_out].
 * Weight matrices are [in, out], zero-padded -- w1 [k1_pad, 512], watt [att_pad, att_pad], ws [504, 512],
 * wh [504, 32] -- and stored in MMA FRAGMENT ORDER: [in / 8][out / 8][lane = 4 (out % 8) + (in % 4)][(in % 8) / 4],
 * i.e. the (b0, b1) operand pair of mma.m16n8k8 of every lane, block by block.  Biases are plain vectors: b1 [512],
 * batt [att_pad], bs [512], bh [32]; att_n = 0 means no attention block (Esc1 / Esc2).
 * The four chains (policy 1 actor, policy 1 critic, policy 2 actor, policy 2 critic) run in ONE launch.
 * precision 0: 3xTF32 tensor-core products on mma.sync (fp32-equivalent results); 1: plain TF32 (hh_policy.cu).  The
 * tcgen05 / TMEM path (precision 2, hh_policy_tc.cu) is reached through hh_policy_forward_ex below with operand images
 * built by hh_policy_pack; it does not read the fragment-ordered matrices. */
typedef struct {
  const float *x, *w1, *b1, *watt, *batt, *wh, *bh;   /* device pointers */
  float* out;                                         /* [n_rows, ld_out], first n_out columns written */
  int32_t ldx, d_in, k1_pad, att_lo, att_n, att_pad, n_out, ld_out;
} hh_policy_chain;
int hh_policy_forward(int32_t n_rows, const hh_policy_chain* chains /* host array of 4 */, const float* ws_dev,
                      const float* bs_dev, int32_t precision, void* stream);
/* General form: 1..8 chains per launch, each with its own shared-layer weights (frozen opponent policy sets have
 * their own SHARED_LAYER), an optional row gather (rows: local row r reads x[rows[begin + r]] and writes the same global
 * row), an optional {begin, count} pair in DEVICE memory (data-dependent row lists without a host synchronisation;
 * n_rows is then the capacity) and an optional per-head argmax (act_out int32 [.., 4]: what _policy_actions does with
 * the logits, env_base.py:373-382; head[] are the MultiDiscrete split sizes).  Used for the level-4/5 opponents and
 * the low-level policies inside HighLevelEnv. */
typedef struct {
  const float *x, *w1, *b1, *watt, *batt, *ws, *bs, *wh, *bh;
  float* out;
  const int32_t* rows;
  const int32_t* range_dev;
  int32_t* act_out;
  int32_t n_rows, ldx, d_in, k1_pad, att_lo, att_n, att_pad, n_out, ld_out, n_heads, head[4];
  int32_t ld_act;   /* row stride of act_out in units of 4 int32 (0 or 1: dense [.., 4]) */
  /* precision = 2 (tcgen05 path): operand images of the four weight matrices and their scale pairs, written by
   * hh_policy_pack; w1 / watt / ws / wh are not read then */
  const void *img_w1, *img_att, *img_ws, *img_wh;
  const float *us_w1, *us_att, *us_ws, *us_wh;
} hh_policy_chain_ex;
/* precision: 0 = 3xTF32 on mma.sync (fp32-equivalent), 1 = TF32 on mma.sync, 2 = tcgen05 / TMEM path: every fp32 operand as
 * two fp16 halves of a power-of-two multiple, three kind::f16 MMAs per product, fp32 accumulation in tensor memory
 * (fp32-equivalent; inputs x are expected in [-15.99, 15.99] -- observations are Box(0, 1), env_hetero.py:28-36). */
int hh_policy_forward_ex(int32_t n_chains, const hh_policy_chain_ex* chains, int32_t precision, void* stream);
/* Operand image of one weight matrix for precision = 2.  w_dev: fp32 row-major [k_rows][ldw] (K x N), of which the first
 * n_cols columns are packed (zero beyond, up to n_total) in chunks of n_chunk columns (the MMA N: 256 for the 500-wide layers
 * -> n_total 512; the padded width itself for the attention block -- a multiple of 16 -- and 32 for the head), image row
 * k' = w row k' - row_shift (zero outside), ksteps * 16 image rows, kps K steps per 16 KB ring stage (1; 8 for the head).
 * The image is laid out for the kernel hh_policy_tc_pair() selects (1: CTA pairs, each CTA streams half of a chunk's columns).
 * image_dev: hh_policy_image_bytes(ksteps, n_total) bytes; unscale_dev: float[2] = {2^-(12 + s), 2^s}.  Runs on `stream`, no
 * host synchronisation (call it again after the weights changed). */
int64_t hh_policy_image_bytes(int32_t ksteps, int32_t n_total);
int hh_policy_pack(const float* w_dev, int32_t k_rows, int32_t n_cols, int32_t ldw, int32_t n_total, int32_t n_chunk,
                   int32_t row_shift, int32_t ksteps, int32_t kps, void* image_dev, float* unscale_dev, void* stream);
int32_t hh_policy_tc_pair(void);
int32_t hh_policy_tc_mode(void);   /* 0: 64-row tiles, 1: CTA pairs, 2 (default): 128-row tiles, lo halves in tensor memory */
/* Row lists per key, built on the device: rows_dev int32 [n_keys][n], ranges_dev int32 [n_keys][2] = {k n, count_k} for the
 * arenas i with key_dev[i] == keys_host[k] (n_keys <= 4).  Level 5 draws the opponents' policy set per arena and episode
 * (env_hetero.py:55-59); the lists feed hh_policy_chain_ex.rows / range_dev without a host synchronisation. */
int hh_policy_rows_by_key(int32_t n, const uint8_t* key_dev, int32_t n_keys, const int32_t* keys_host, int32_t* rows_dev,
                          int32_t* ranges_dev, void* stream);
const char* hh_policy_last_error(void);

/* Test access to the device WGS84 solvers (replacing geographiclib's Geodesic.WGS84 as used at
 * warsim/utils/geodesics.py:12-24).  in_host: f64[4][n], out_host: f64[2][n].
 *   mode 0: direct  (lat1, lon1, azi1 [deg], s12 [m]) -> (lat2, lon2)
 *   mode 1: inverse (lat1, lon1, lat2, lon2)          -> (s12 [m], azi1 [deg])   exact series
 *   mode 2: inverse, local closed form used for threshold decisions (same outputs)
 *   mode 3 / 4: the short-arc direct solve as the step kernel calls it for aircraft / rockets (same in / out as 0) */
int hh_debug_geodesic(int32_t mode, int32_t n, const double* in_host, double* out_host);

const char* hh_last_error(void);
const char* hh_version(void);

/* ============================================================================================
 * HighLevelEnv (envs/env_hier.py): the 3-vs-3 commander environment (BASELINE config 5).
 *   hh_hier_create  <- HighLevelEnv.__init__(env_config)          env_hier.py:31-42
 *   hh_hier_reset   <- HighLevelEnv.reset()                       env_hier.py:44-47
 *   one HHMARLBaseEnv.step(commander_actions) (env_base.py:79-109 -> _take_action env_hier.py:114-140) is
 *     hh_hier_begin  : _action_assess (env_hier.py:142-190) + low-level observations of the agents
 *     16 x { [agents' networks]  hh_hier_agents : agents' _take_base_action + opponents' low-level observations
 *            [opponents' networks] hh_hier_tick : opponents' _take_base_action, do_tick, _get_rewards,
 *                                                 _surrounding_event, next observations of the agents }
 *     hh_hier_end    : termination (env_base.py:89-90), optional auto-reset, HighLevelEnv.state() (34-d)
 *   Arenas whose sub-step loop has ended idle through the remaining sub-step calls.
 * Shapes: commander_actions int32[N][3] in {0,1,2}; ll_obs f32[N][6][30] (row of unit id-1; fight 26/24,
 * escape 30/29 entries used); ll_info u8[N][6]: bit0 = unit queries its policy now, bit1 = escape policy,
 * bit2 = aircraft type 2, bit3 (as written by hh_hier_begin only) = alive at the start of the commander step;
 * actions int32[N][6][4]; obs f32[N][3][34]; rew f32[N][3]; done u8[N]; substeps i32[N].
 * ============================================================================================ */
typedef struct hh_hier_env hh_hier_env;

typedef struct {
  int32_t horizon;              /* 500 (config.py:98) */
  int32_t level;                /* args.level keeps its default 1 in train_hier.py: opponents start at speed 0 */
  int32_t friendly_kill;        /* bool */
  int32_t hier_action_assess;   /* bool (config.py:44) */
  int32_t hier_opp_fight_ratio; /* percent (config.py:23) */
  int32_t autoreset;
  double map_size;              /* 0.5 */
  double rew_scale, glob_frac;
  uint64_t seed, arena_base;
} hh_hier_config;

/* Per-arena record, device layout == host layout (hh_hier_get_state / hh_hier_set_state copy it verbatim).
 * Index u = aircraft id - 1 (ids 1-3 agents, 4-6 opponents); rocket fields belong to shooter u. */
typedef struct {
  double lat[6], lon[6], hdg[6], spd[6], nhdg[6], nspd[6];
  double rlat[6], rlon[6], rhdg[6], rnhdg[6];
  double ota_dn[6][3];          /* opp_to_attack[i][k][1]: normalised distance at the last commander step */
  double rewards[3];
  uint64_t dg;                  /* G-stream draw counter */
  int32_t crem[6], burst[6], mrem[6], mwait[6], rid[6];
  int32_t steps, alive_ag, alive_op, next_id, sub, kill_event, situation_event, active, err;
  uint32_t dc;                  /* C-stream draw counter */
  int8_t ca[6];                 /* commander actions after _action_assess, -1 = None */
  uint8_t alive[6], hasm[6], actype[6], ralive[6], rage[6], rtgt[6];
  uint8_t ota_n[6], ota_id[6][3]; /* opp_to_attack lists (0-based unit indices) */
} hh_hier_arena;

int hh_hier_create(const hh_hier_config* cfg, int32_t n_arenas, int32_t device, hh_hier_env** out);
void hh_hier_destroy(hh_hier_env* env);
int hh_hier_reset(hh_hier_env* env, const uint8_t* mask_dev, float* obs_dev, void* stream);
int hh_hier_begin(hh_hier_env* env, const int32_t* commander_actions_dev, float* ll_obs_dev, uint8_t* ll_info_dev, void* stream);
int hh_hier_agents(hh_hier_env* env, const int32_t* actions_dev, float* ll_obs_dev, uint8_t* ll_info_dev, void* stream);
int hh_hier_tick(hh_hier_env* env, const int32_t* actions_dev, float* ll_obs_dev, uint8_t* ll_info_dev, void* stream);
int hh_hier_end(hh_hier_env* env, float* obs_dev, float* rew_dev, uint8_t* done_dev, int32_t* substeps_dev, void* stream);
/* HHMARLBaseEnv.step with args.eval_info (env_base.py:91-107; summed by evaluation.py:59-60, reported by
 * postprocess_eval evaluation.py:66-82): info int32[N][12] = agents_win, opps_win, draw, agent_fight, agent_escape,
 * opp_fight, opp_escape, agent_steps, opp_steps, opp1, opp2, opp3 of the commander step in flight.  Call it after
 * the last hh_hier_tick and BEFORE hh_hier_end (whose auto-reset replaces the finished episode). */
#define HH_HIER_EVAL_INFO_LEN 12
/* Row lists of the frozen low-level policies for the commander step that hh_hier_begin opened, built on the device (no host
 * synchronisation): list 2 k + h (k: 0 fight AC1, 1 fight AC2, 2 escape AC1, 3 escape AC2; h: 0 agents = units 1-3, 1 opponents
 * = units 4-6) holds the flat unit indices arena * 6 + unit of the units alive at the start of the step (ll_info bit 3).
 * rows_dev int32 [8][3 N], ranges_dev int32 [8][2] = {list * 3 N, count}: feed hh_policy_chain_ex.rows / range_dev. */
int hh_hier_policy_rows(hh_hier_env* env, const uint8_t* ll_info_dev, int32_t* rows_dev, int32_t* ranges_dev, void* stream);
int hh_hier_eval_info(hh_hier_env* env, int32_t* info_dev, void* stream);
int hh_hier_get_state(hh_hier_env* env, hh_hier_arena* out_host);
int hh_hier_set_state(hh_hier_env* env, const hh_hier_arena* in_host);
uint64_t hh_hier_launch_count(const hh_hier_env* env);
const char* hh_hier_last_error(void);

#ifdef __cplusplus
}
#endif
#endif
