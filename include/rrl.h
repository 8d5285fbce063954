/* rrl.h -- C ABI of librrl.so: the B200 (sm_100a) hot path of Recovery RL.
 *
 * The reference (abalakrishna123/recovery-rl) has NO native/FFI interface: its seam is the
 * Python class API consumed by recovery_rl/experiment.py.  Each entry point below therefore
 * cites the reference Python symbol (file:line) whose arithmetic it replaces; the Python
 * binding a maintainer would add is shown in INTEGRATION.md.
 *
 * Conventions
 *   - every function returns 0 on success, <0 on error; rrl_last_error() gives the message
 *     (thread-local).  Launch-only: functions ENQUEUE work on `stream` (a cudaStream_t passed
 *     as void*) and never synchronise, so every call is CUDA-graph capturable, except the
 *     explicitly named *_host helpers which touch host memory and the rrl_debug_* diagnostics (they synchronise the device).
 *   - all arrays are caller-owned DEVICE pointers (e.g. torch tensors' data_ptr()); nothing
 *     is allocated here.  No global state besides the launch-mode switch (rrl_set_pdl) and the diagnostic sums of
 *     rrl_debug_opt_times; the device is whatever is current (cudaSetDevice).
 *   - layouts: env state fp64 SoA [2][n] (plane 0 = x, plane 1 = y); actions fp32 [n][2];
 *     replay = ring of 32-byte records {s.x,s.y,a.x,a.y,r,s2.x,s2.y,mask} fp32 (one DRAM
 *     sector per transition) + one flag byte per slot for the constraint buffer.
 *   - data-dependent control flow (len(memory) > batch_size, the Q_risk online gate, the
 *     effective batch size) lives in device counters so that a captured graph replays exactly.
 */
#ifndef RRL_H_
#define RRL_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define RRL_VERSION 6

const char* rrl_last_error(void);
int rrl_version(void);
/* Programmatic dependent launch of the vector step's kernels (sampler, tcgen05 forward / backward, optimizer step, acting,
 * env step): each is scheduled while its predecessor in the stream still runs and waits (griddepcontrol.wait, its first
 * statement) for that kernel's completion, which hides the launch latency of the ~20 short dependent kernels of a step.
 * OFF by default (measured slower on B200 inside the captured step: 0.366 vs 0.339 ms); RRL_PDL=1 in the environment or
 * rrl_set_pdl(1) turns it on.  Returns the previous setting
 * (enabled < 0: query only). */
int rrl_set_pdl(int enabled);
/* Diagnostics of the tiled optimizer-step kernels (tcgen05 path) since the previous call, summed over launches, device
 * globaltimer nanoseconds; kept OUTSIDE the counter block (which stays deterministic).  Synchronises the device, resets the sums.
 *   out[0] CTA 0: kernel start -> bias corrections / cross-GPU barrier done   out[1] -> gradients (own + peers') loaded
 *   out[2] -> CTA 0 done          out[3] launches
 *   out[4] time CTA 0 waited for the slowest peer's gradient flag (fused barrier: rank skew + signal latency)   out[5] barriers
 *   out[6] CTA 0 start -> last CTA end (the whole kernel)                     out[7] reserved */
int rrl_debug_opt_times(unsigned long long* out8);

/* ------------------------------------------------------------------ environments ------- */
enum { RRL_ENV_NAV1 = 0, RRL_ENV_NAV2 = 1, RRL_ENV_MAZE = 2 };

/* device counter block (int64), shared by env/replay/agent kernels */
enum {
    RRL_C_TOTAL_NUMSTEPS = 0,  /* experiment.py:429  total_numsteps                       */
    RRL_C_EPISODES       = 1,  /* finished episodes                                        */
    RRL_C_NUM_VIOLS      = 2,  /* experiment.py:455-456                                    */
    RRL_C_NUM_SUCCESSES  = 3,  /* experiment.py:461                                        */
    RRL_C_VIOL_RECOVERY  = 4,  /* experiment.py:457-458                                    */
    RRL_C_VIOL_NO_RECOV  = 5,  /* experiment.py:459-460                                    */
    RRL_C_OFFLINE_VIOLS  = 6,  /* experiment.py:279  num_constraint_violations (host-set) */
    RRL_C_VEC_STEP       = 7,  /* vector-step index (Philox counter)                       */
    RRL_C_SAC_UPDATES    = 8,  /* experiment.py:416  updates                               */
    RRL_C_QRISK_UPDATES  = 9,  /* qrisk.py:163       self.updates                          */
    RRL_C_TASK_POS       = 10, /* replay_memory.py:25 position  (task buffer)              */
    RRL_C_TASK_LEN       = 11, /* len(memory)                                              */
    RRL_C_CONS_POS       = 12, /* replay_memory.py:52 position  (constraint buffer)        */
    RRL_C_CONS_LEN       = 13,
    RRL_C_SAC_ROWS       = 14, /* effective batch rows of the current SAC update (0 = skip)*/
    RRL_C_QRISK_ROWS     = 15, /* effective batch rows of the current Q_risk update        */
    RRL_C_ADAM_T0        = 16, /* Adam step counts: critic, policy, qrisk, recovery        */
    RRL_C_EXT_VIOLS      = 20, /* violations seen on OTHER ranks (multi-GPU gate), host/NCCL-set */
    RRL_C_RETURN_SUM_BITS= 21, /* double bit pattern: sum of finished-episode returns      */
    RRL_C_ERROR          = 22, /* sticky device-side error code (1: sample larger than population, 2: a peer never reached a barrier) */
    RRL_C_ADAM_T_ALPHA   = 23, /* Adam step counts of the scalar multipliers: log_alpha (sac.py:245-247),  */
    RRL_C_ADAM_T_NU      = 24, /*   log_nu (sac.py:259-261),                                                */
    RRL_C_ADAM_T_LAMBDA  = 25, /*   log_lambda_RCPO (sac.py:268-270)                                        */
    RRL_C_TICKET         = 26, /* scratch: CTA arrival ticket of the fused optimizer-step kernel (always 0 between launches) */
    RRL_C_TICKET2        = 27, /* scratch: CTA arrival ticket of the update kernels whose last CTA runs the loss / sample-backward stage */
    RRL_C_GATE_SATISFIED = 28, /* multi-GPU: the violation count of the Q_risk online gate (experiment.py:410) has passed its threshold on
                                  every rank (monotone: counts only grow) -> rrl_peer_sync_gate_counts stops exchanging */
    RRL_NUM_COUNTERS     = 32
};

typedef struct {
    int32_t kind;              /* RRL_ENV_*                                                */
    int32_t horizon;           /* _max_episode_steps: navigation1.py:34,65  maze.py:16,121 */
    int64_t n_envs;
    double  reward_penalty;    /* experiment.py:431-432 constraint_reward_penalty          */
    uint64_t seed;             /* Philox key (production RNG mode)                         */
    int32_t stream_id;         /* rank: second half of the Philox key                      */
    int32_t maze_substeps;     /* maze.py:147 (500)                                        */
    int32_t flags;             /* RRL_ENV_NO_AUTO_RESET: the caller resets explicitly (gym-style
                                  env.reset(), offline-data generators that keep stepping after done) */
    int32_t reserved;
} rrl_env_config_t;
enum { RRL_ENV_NO_AUTO_RESET = 1 };

/* env.reset (navigation1.py:91-97, navigation2.py:90-96, maze.py:184-213 mode 'h').
 * mask: NULL = reset all, else reset env i iff mask[i] != 0.
 * draws: NULL = Philox, else host-supplied fp64 [2][n]: N(0,1) for navigation, U[0,1) for maze. */
int rrl_env_reset(const rrl_env_config_t* cfg, const uint8_t* mask, const double* draws,
                  double* state, int32_t* ep_steps, double* ep_return, const int64_t* counters,
                  void* stream);

/* One vector step of N env copies (navigation1.py:71-110, navigation2.py:70-110,
 * maze.py:139-168, obstacle.py:13-15,44-45) fused with what experiment.py:420-461 does with
 * the result: reward penalty, mask = !done BEFORE the horizon check, push of
 * (s, a_task, r, s', mask) into the task ring and (s, a_real, constraint, s', mask) into the
 * constraint ring (replay_memory.py:21-25,47-52), episode statistics, auto-reset.
 *   action_task/action_real : fp32 [n][2] proposed / executed action (experiment.py:438-445)
 *   recovery                : u8 [n] recovery_used flag (may be NULL)
 *   noise      : NULL = Philox; else fp64 [2][n] N(0,1) dynamics noise (navigation only)
 *   reset_draws: NULL = Philox; else fp64 [2][n] draws used by envs that finish this step
 *   task_ring / cons_ring / cons_flags : may be NULL (no push)
 *   out_*      : per-env results of THIS step (pre-reset), any may be NULL
 *   action_real_f64 : NULL, or fp64 [n][2] executed action used for the DYNAMICS instead of action_real
 *                (the offline-data generators step with float64 actions: navigation1.py:150-152)
 */
int rrl_env_step(const rrl_env_config_t* cfg, const float* action_task, const float* action_real,
                 const uint8_t* recovery, const double* noise, const double* reset_draws,
                 double* state, int32_t* ep_steps, double* ep_return,
                 float* task_ring, int64_t task_capacity,
                 float* cons_ring, uint8_t* cons_flags, int64_t cons_capacity,
                 int64_t* counters,
                 double* out_next_state, double* out_reward, uint8_t* out_done,
                 uint8_t* out_constraint, uint8_t* out_success, const double* action_real_f64, void* stream);

/* After rrl_env_step: position/len of both rings += n, total_numsteps += n, vec_step += 1
 * (replay_memory.py:22-25; experiment.py:429).  push_task / push_cons select the rings. */
int rrl_counters_advance(int64_t* counters, int64_t n, int64_t task_capacity, int64_t cons_capacity,
                         int push_task, int push_cons, void* stream);

/* ------------------------------------------------------------------ replay ------------- */
/* CPython `random.seed(int)` (init_by_array) -> 624-word MT19937 state + index (word 624).
 * Host helper: fills a HOST buffer of 625 uint32 (replay_memory.py:16,41). */
int rrl_mt19937_seed_host(const uint32_t* key_limbs, int n_limbs, uint32_t* state625_host);

/* ReplayMemory.push / ConstraintReplayMemory.push for n caller-supplied transitions (offline
 * demos: experiment.py:277-282).  rec: fp32 [n][8]; flags written when cons_flags != NULL. */
int rrl_replay_push(float* ring, uint8_t* cons_flags, int64_t capacity, const float* rec, int64_t n,
                    int64_t* counters, int is_constraint_buffer, void* stream);

/* Per-chunk positive/negative counts of the constraint flags (np.argwhere over pos_idx,
 * replay_memory.py:58-66).  chunk_counts: int32 [2][n_chunks]. */
int rrl_replay_flag_count(const uint8_t* cons_flags, int64_t capacity, int32_t chunk,
                          int32_t* chunk_counts, void* stream);

typedef struct {
    int64_t capacity;
    int32_t batch_size;        /* requested B                                             */
    int32_t is_constraint;     /* 0: ReplayMemory.sample, 1: ConstraintReplayMemory.sample */
    double  pos_fraction;      /* < 0: None (qrisk.py:77)                                  */
    int32_t gate_mode;         /* 0: always sample min(B, len) rows (pre-training, qrisk.py:100-104);
                                  1: SAC gate   len > B                 (experiment.py:397);
                                  2: Q_risk gate: SAC gate && len > B && (viols)/B > pos_fraction (experiment.py:397,407-410) */
    int32_t chunk;             /* flag chunk size used by rrl_replay_flag_count            */
    double  gate_pos_fraction; /* exp_cfg.pos_fraction as given on the command line        */
} rrl_sample_config_t;

/* random.sample-compatible index draw + gather (replay_memory.py:27-30, 54-72; CPython
 * random.py sample/_randbelow).  Single-CTA kernel; writes the effective row count into
 * counters[rows_counter] (0 = gate closed: the MT stream is NOT advanced).
 *   mt_state : uint32 [625] device (shared by both buffers: one global `random` stream)
 *   out_idx  : int64 [B] slot indices;  out_* : fp32 batch arrays [B][2],[B][2],[B],[B][2],[B] */
int rrl_replay_sample(const rrl_sample_config_t* cfg, const float* ring, const uint8_t* cons_flags,
                      const int32_t* chunk_counts, uint32_t* mt_state, int64_t* counters,
                      int rows_counter, int64_t* out_idx, float* out_s, float* out_a, float* out_r,
                      float* out_s2, float* out_m, void* stream);

/* ------------------------------------------------------------------ agent -------------- */
enum { RRL_NET_CRITIC = 0, RRL_NET_CRITIC_TARGET = 1, RRL_NET_POLICY = 2, RRL_NET_QRISK = 3,
       RRL_NET_QRISK_TARGET = 4, RRL_NET_RECOVERY = 5, RRL_NUM_NETS = 6 };

typedef struct {
    int32_t hidden;            /* 256 (arg_utils.py:89-92); kernels are specialised for 256 */
    int32_t max_batch;         /* rows of update scratch                                   */
    float gamma, alpha, tau;               /* arg_utils.py:55-68                            */
    float gamma_safe, tau_safe, eps_safe;  /* arg_utils.py:113-127                          */
    float lr, beta1, beta2, adam_eps;      /* torch.optim.Adam defaults, sac.py:84          */
    float action_scale[2], action_bias[2]; /* model.py:312-315                              */
    int32_t target_update_interval;        /* arg_utils.py:39-43                            */
    int32_t mf_recovery;                   /* qrisk.py:150                                  */
    float grad_scale;                      /* 1/world_size applied inside Adam              */
    int32_t use_tensor_cores;              /* 0 = fp32 FFMA kernels, 1 = tcgen05 fp16 hi/lo split (3 MMAs) for the acting kernel and the update
                                              GEMMs, 2 = 1 + the per-row update stages (losses, sample backward) fused into the
                                              producing kernels as last-CTA tails */
    /* comparison-algorithm branches of SAC.update_parameters (sac.py:52-72, 95-101, 116-117) */
    int32_t algo_flags;                    /* RRL_ALGO_* bits                               */
    float target_entropy;                  /* -dim(A) (sac.py:97-98)                         */
    double nu, lambda_rcpo;                /* initial multipliers (arg_utils.py:230-255)     */
    double lr64;                           /* args.lr as the python float (0 -> (double)lr): the scalar
                                              multipliers are float64 tensors with Adam lr 0.1*lr (sac.py:64,72) */
} rrl_agent_config_t;
enum {
    RRL_ALGO_DGD = 1,            /* --DGD_constraints: policy loss += nu*(max Q_risk(s,pi) - eps_safe)  sac.py:224-228 */
    RRL_ALGO_UPDATE_NU = 2,      /* --update_nu: Adam on log_nu                                    sac.py:256-262 */
    RRL_ALGO_RCPO = 4,           /* --RCPO: target -= lambda*max Q_risk(s,a); Adam on log_lambda   sac.py:202-205,265-271 */
    RRL_ALGO_AUTO_ALPHA = 8,     /* --automatic_entropy_tuning: Adam on log_alpha                  sac.py:241-250 */
    RRL_ALGO_DETERMINISTIC = 16  /* --policy Deterministic (model.py:447-485): alpha = 0, action = mean + noise */
};
/* scalar block inside the arena (scratch name "scalars", 32 floats, 8-byte aligned).  fp32 slots: */
enum {
    RRL_S_ALPHA = 0,       /* alpha used by the NEXT update (args.alpha, or exp(log_alpha) after a tuning step) */
    RRL_S_NU_ARG = 1,      /* the `nu` argument of update_parameters (experiment.py:406: nu_schedule(i_episode)); host-set */
    RRL_S_LOG_ALPHA = 2, RRL_S_G_LOG_ALPHA = 3, RRL_S_M_ALPHA = 4, RRL_S_V_ALPHA = 5,
    RRL_S_ALPHA_LOSS = 6,
    RRL_S_F64_BASE = 8     /* float64 slots follow (index in doubles from here): */
};
enum {
    RRL_D_G_LOG_NU = 0, RRL_D_G_LOG_LAMBDA = 1,   /* adjacent: the multi-GPU gradient sum of the two */
    RRL_D_LOG_NU = 2, RRL_D_M_NU = 3, RRL_D_V_NU = 4,
    RRL_D_LOG_LAMBDA = 5, RRL_D_M_LAMBDA = 6, RRL_D_V_LAMBDA = 7,
    RRL_D_LAMBDA = 8,      /* lambda_RCPO used by the next update (args value, then exp(log_lambda)) */
    RRL_D_NU_LEARNED = 9,  /* exp(log_nu) (sac.py:262; the reference never feeds it back: nu is always passed) */
    RRL_D_ZERO = 10        /* constant 0: log_std of the Deterministic policy's unit-variance noise term */
};

/* Arena layout (fp32 elements).  Tensors of every net follow torch's parameters() order of the
 * reference module (model.py:49-76,172-199,295-343,489-530), each padded to 4 floats. */
int64_t rrl_agent_arena_floats(const rrl_agent_config_t* cfg);
int rrl_agent_num_tensors(int net);
int rrl_agent_tensor_info(const rrl_agent_config_t* cfg, int net, int tensor, int64_t* offset,
                          int64_t* rows, int64_t* cols);
/* flat gradient block [critic | policy | qrisk | recovery] for the NCCL all-reduce */
int rrl_agent_grad_range(const rrl_agent_config_t* cfg, int net, int64_t* offset, int64_t* count);
/* named scratch regions inside the arena (batch arrays, per-row outputs, losses) */
int rrl_agent_scratch_info(const rrl_agent_config_t* cfg, const char* name, int64_t* offset,
                           int64_t* count);

/* Write the initial values of the scalar block (alpha, nu, lambda, their logs; Adam moments zero) from cfg.
 * Call once after allocating the arena (sac.py:48,56-72,99-101). */
int rrl_agent_init_scalars(const rrl_agent_config_t* cfg, float* arena, void* stream);

/* Rebuild the derived weight images (k-major copies of the ten 256x256 hidden matrices that the GEMM
 * kernels stream) after the host wrote parameters into the arena (e.g. the xavier init of
 * model.py:23-26 or a checkpoint load).  The update kernels keep them fresh themselves. */
int rrl_agent_refresh(const rrl_agent_config_t* cfg, float* arena, void* stream);

/* Rebuild only the fp16 hi/lo operand images the tensor-core acting kernel streams (policy, Q_risk twin,
 * recovery policy); call after an optimizer step when cfg->use_tensor_cores is set. */
int rrl_agent_tc_refresh(const rrl_agent_config_t* cfg, float* arena, void* stream);

/* Stages of the composite action, for callers that overlap acting with the updates (rrl_agent_act_stage): each stage only
 * needs the networks named, so it can be enqueued right after THEIR optimizer step, next to the remaining updates.
 * All three in one call == rrl_agent_act. */
enum {
    RRL_ACT_STAGE_POLICY = 1,    /* task policy (sac.py:133-168) or the start-phase random action -> action_task            */
    RRL_ACT_STAGE_QRISK = 2,     /* Q_risk(s, action_task) (qrisk.py:184-196), threshold (experiment.py:555) -> qrisk_out, recovery */
    RRL_ACT_STAGE_RECOVERY = 4,  /* recovery policy (qrisk.py:198-213), select -> action_real                                */
    RRL_ACT_STAGE_ALL = 7
};
/* Composite action selection for N envs (experiment.py:546-577; sac.py:133-168;
 * qrisk.py:184-213; model.py:317-338, 512-525):
 *   a_task = tanh(mu + sigma*eps_task)*scale + bias   (or mean action when eval != 0;
 *            or U(low,high) while total_numsteps < start_steps, experiment.py:559-560)
 *   recovery = max(Q1,Q2)_risk(s, a_task) > eps_safe   (only if use_recovery)
 *   a_real = recovery ? mu_rec + sigma_rec*eps_rec : a_task
 * state fp64 [2][n]; eps_* fp32 [n][2] or NULL (Philox); rand_u fp32 [n][2] U[0,1) or NULL. */
int rrl_agent_act(const rrl_agent_config_t* cfg, float* arena, int64_t n, const double* state,
                  const float* eps_task, const float* eps_rec, const float* rand_u,
                  int use_recovery, int eval, int64_t start_steps, uint64_t seed, int32_t stream_id,
                  const int64_t* counters, float* action_task, float* action_real,
                  uint8_t* recovery, float* qrisk_out, void* stream);
/* The same, restricted to `stages` (ONE RRL_ACT_STAGE_* or RRL_ACT_STAGE_ALL; single stages on the tcgen05 path only) on at most
 * `max_ctas` SMs (0: all): later stages read action_task / recovery written by the earlier ones. */
int rrl_agent_act_stage(const rrl_agent_config_t* cfg, float* arena, int64_t n, const double* state,
                  const float* eps_task, const float* eps_rec, const float* rand_u,
                  int use_recovery, int eval, int64_t start_steps, uint64_t seed, int32_t stream_id,
                  const int64_t* counters, float* action_task, float* action_real,
                  uint8_t* recovery, float* qrisk_out, int stages, int max_ctas, void* stream);

/* ---- comparison branches of the acting path, vectorised over env copies (csrc/select.cu) ----
 * Every env copy evaluates `samples` candidate actions through the batched forward passes above; `workspace` holds the
 * candidates of one chunk of env copies (rrl_select_workspace_floats(chunk_envs, samples) floats; a workspace smaller than
 * n env copies makes the call loop over chunks, the draws do not depend on the chunking). */
int64_t rrl_select_workspace_floats(int64_t chunk_envs, int32_t samples);
/* SAC.select_action with --use_constraint_sampling (SQRL, sac.py:139-161): per env copy
 *   pi_j, log_pi_j = GaussianPolicy.sample(state), j < samples (100 in the reference);  q_j = max(Q1,Q2)_risk(state, pi_j)
 *   F = {j : q_j <= eps_safe};  F empty: a = pi[argmin q];  else r ~ Categorical(exp(log_pi_F)) and a = pi[r] -- r is the index
 *   INTO F used on the unfiltered list, as sac.py:157-159 is written.
 * The Categorical draw is an inverse-CDF draw over F in sample order with one uniform per env.  While total_numsteps <
 * start_steps the action is U(low, high) as in rrl_agent_act (experiment.py:559-560).  Writes action_task = action_real = a,
 * recovery = 0, qrisk_out = q of the chosen candidate.
 * eps_cand fp32 [n][samples][2] N(0,1) or NULL (Philox); cat_u fp32 [n] U[0,1) or NULL; rand_u fp32 [n][2] or NULL. */
int rrl_sqrl_select_action(const rrl_agent_config_t* cfg, float* arena, int64_t n, int32_t samples, const double* state,
                           const float* eps_cand, const float* cat_u, const float* rand_u, int64_t start_steps,
                           uint64_t seed, int32_t stream_id, const int64_t* counters, float* workspace,
                           int64_t workspace_floats, float* action_task, float* action_real, uint8_t* recovery,
                           float* qrisk_out, void* stream);
/* QRiskWrapper.select_action with --Q_sampling_recovery (qrisk.py:214-225): for every env copy whose recovery flag is set
 * (NULL: all), `samples` (1000 in the reference) uniform actions of the Box, a_real = the one with the smallest
 * max(Q1,Q2)_risk.  cand_u fp32 [n][samples][2] U[0,1) or NULL (Philox). */
int rrl_qsample_recovery_action(const rrl_agent_config_t* cfg, float* arena, int64_t n, int32_t samples,
                                const double* state, const float* cand_u, const uint8_t* recovery, uint64_t seed,
                                int32_t stream_id, const int64_t* counters, float* workspace, int64_t workspace_floats,
                                float* action_real, void* stream);
/* --add_both_transitions (experiment.py:446-448): after rrl_env_step + rrl_counters_advance, every env copy whose recovery
 * policy acted pushes (state, real_action, reward, next_state, mask) into the task ring as well: the rows are appended in
 * env order after the step's n rows and TASK_POS / TASK_LEN advance by their number.  Needs task_capacity >= 2 n. */
int rrl_replay_push_both(float* task_ring, int64_t task_capacity, int64_t n, const uint8_t* recovery,
                         const float* action_real, int64_t* counters, void* stream);

/* SAC.update_parameters (sac.py:170-277, ordering "Variant B" of SURVEY.md §8c) split so that
 * the host can all-reduce the gradient block between the two halves:
 *   rrl_sac_backward : forward passes, losses, grads of critic and policy into the grad block
 *   rrl_sac_apply    : Adam on critic+policy (torch.optim.Adam), Polyak on critic_target
 * batch arrays live in the arena scratch ("sac_s","sac_a","sac_r","sac_s2","sac_m");
 * eps_next/eps_cur fp32 [B][2] or NULL (Philox).  losses: fp32 [8] device:
 *   {qf1_loss, qf2_loss, policy_loss, alpha_loss, alpha (before this update's tuning step)}.
 * cfg->algo_flags selects the comparison branches: they read alpha / nu / lambda from the scalar block, add the
 * Q_risk(s,a) (RCPO) and Q_risk(s,pi) (DGD, update_nu) passes, and rrl_sac_apply also steps the scalar Adams.
 * With RRL_ALGO_DETERMINISTIC eps_next / eps_cur hold the (already scaled and clamped) noise vector of
 * DeterministicPolicy.sample repeated on every row; NULL draws ONE such vector per pass from Philox. */
int rrl_sac_backward(const rrl_agent_config_t* cfg, float* arena, const float* eps_next,
                     const float* eps_cur, uint64_t seed, int32_t stream_id, int64_t* counters,
                     float* losses, void* stream);
int rrl_sac_apply(const rrl_agent_config_t* cfg, float* arena, int64_t* counters, void* stream);

/* QRiskWrapper.update_parameters (qrisk.py:86-182): critic half, then the MF recovery policy
 * on the POST-step critic (qrisk.py:150-158), then Polyak (qrisk.py:160-163).
 * losses: fp32 [8]: {q1_loss, q2_loss, recovery_policy_loss}. */
int rrl_qrisk_backward(const rrl_agent_config_t* cfg, float* arena, const float* eps_next,
                       uint64_t seed, int32_t stream_id, int64_t* counters, float* losses,
                       void* stream);
int rrl_qrisk_apply(const rrl_agent_config_t* cfg, float* arena, int64_t* counters, void* stream);
int rrl_recovery_backward(const rrl_agent_config_t* cfg, float* arena, const float* eps_rec,
                          uint64_t seed, int32_t stream_id, int64_t* counters, float* losses,
                          void* stream);
int rrl_recovery_apply(const rrl_agent_config_t* cfg, float* arena, int64_t* counters, void* stream);
/* rrl_recovery_backward in two calls, for callers that overlap: the recovery policy's own forward pass (qrisk.py:150-151
 * `self.policy.sample(state_batch)`) needs only the sampled batch and the recovery policy, so it can be enqueued on another
 * stream next to the SAC / safety-critic updates; rrl_recovery_backward_rest is everything that needs the POST-step safety
 * critic (qrisk.py:152-158).  rrl_recovery_forward + rrl_recovery_backward_rest == rrl_recovery_backward. */
int rrl_recovery_forward(const rrl_agent_config_t* cfg, float* arena, const float* eps_rec, uint64_t seed, int32_t stream_id,
                         int64_t* counters, void* stream);
int rrl_recovery_backward_rest(const rrl_agent_config_t* cfg, float* arena, int64_t* counters, float* losses, void* stream);

/* ---- multi-GPU: gradient sum fused into the optimizer step over NVLink peer memory --------------------------------
 * With the agent arena of every rank allocated in symmetric (peer-mapped) memory, the NCCL all-reduce between
 * rrl_*_backward and rrl_*_apply is replaced by: rrl_peer_barrier (every rank has finished writing its gradients),
 * then rrl_*_apply_p2p, whose optimizer-step kernel reads the gradient block of EVERY rank through the peer pointers
 * and sums them in rank order (identical result on all ranks) before Adam -- no staging buffer, no collective launch.
 * One barrier per optimizer step suffices: a gradient region is rewritten only after two later barriers of the cycle.
 *   arena[r]  : device pointer to rank r's arena as mapped in THIS process (arena[rank] == the local arena)
 *   signal[r] : device pointer to rank r's uint32 signal pad (>= 2 KB, zero-initialised); words [256, 256 + 32) are used
 *   epoch     : device int64, private to the rank, zero-initialised, NOT restored by snapshots (barrier generation) */
typedef struct {
    int32_t world, rank;
    uint64_t arena[8];
    uint64_t signal[8];
    uint64_t epoch;      /* device pointer to the int64 barrier generation counter of this rank, or 0.  Non-zero: the
                            rrl_*_apply_p2p kernels (tcgen05 path) run the flag barrier THEMSELVES before they load the peers'
                            gradients -- CTA 0 publishes this rank's generation, every CTA waits on the local pad -- and the
                            caller launches no rrl_peer_barrier in front of them */
    uint64_t mc_arena;   /* multicast (NVLS) address of the symmetric arena, or 0.  Non-zero: the optimizer-step kernels read the SUM of
                            all ranks' gradients with multimem.ld_reduce (reduced inside the NVSwitch: 1 x instead of world x gradient
                            bytes per rank) instead of loading every peer's block.  The summation order is the switch's, identical on
                            every rank; like NCCL's it differs from the rank-order sum of the peer loads by rounding above 2 ranks */
} rrl_peers_t;
int rrl_peer_barrier(const rrl_peers_t* peers, int64_t* epoch, int64_t* counters, void* stream);
/* the same barrier, also exchanging the Q_risk gate counts (experiment.py:407-410 must open on every rank in the same
 * step): counters[RRL_C_EXT_VIOLS] = sum over the other ranks of (RRL_C_NUM_VIOLS + RRL_C_OFFLINE_VIOLS).
 * Uses int64 slots at signal words [272, 272 + 2*world).  gate_batch / gate_pos_fraction: the gate's own threshold
 * (total / batch > pos_fraction, experiment.py:410); once it is passed -- on every rank in the same step, and for good, the
 * counts only grow -- RRL_C_GATE_SATISFIED is set and later calls return without touching the other GPUs (gate_batch 0:
 * exchange forever). */
int rrl_peer_sync_gate_counts(const rrl_peers_t* peers, int64_t* epoch, int64_t* counters, int32_t gate_batch,
                              double gate_pos_fraction, void* stream);
int rrl_sac_apply_p2p(const rrl_agent_config_t* cfg, float* arena, int64_t* counters, const rrl_peers_t* peers, void* stream);
int rrl_qrisk_apply_p2p(const rrl_agent_config_t* cfg, float* arena, int64_t* counters, const rrl_peers_t* peers, void* stream);
int rrl_recovery_apply_p2p(const rrl_agent_config_t* cfg, float* arena, int64_t* counters, const rrl_peers_t* peers, void* stream);

/* Twin forward used by QRiskWrapper.get_value / __call__ and the critic (qrisk.py:184-196,
 * 303-307; model.py:65-76,188-199): q1,q2 fp32 [n]; s,a fp32 [n][2]. */
int rrl_twin_q_forward(const rrl_agent_config_t* cfg, const float* arena, int net, int64_t n,
                       const float* s, const float* a, float* q1, float* q2, void* stream);
/* GaussianPolicy.sample / StochasticPolicy.sample on fp32 states (model.py:325-338,522-525). */
int rrl_policy_sample(const rrl_agent_config_t* cfg, const float* arena, int net, int64_t n,
                      const float* s, const float* eps, float* action, float* log_prob,
                      float* mean_action, void* stream);

/* utils.soft_update / hard_update (utils.py:46-54) on whole nets. */
int rrl_hard_update(const rrl_agent_config_t* cfg, float* arena, int dst_net, int src_net, void* stream);
/* recovery_rl/utils.py:46-49 soft_update(target, source, tau): target = target*(1-tau) + source*tau (+ operand images). */
int rrl_soft_update(const rrl_agent_config_t* cfg, float* arena, int dst_net, int src_net, float tau, void* stream);

/* ------------------------------------------------------------------ model-based recovery (PETS / CEM) ---- */
/* BASELINE config 5.  The planner of recovery_rl/MPC.py:322-467 + recovery_rl/optimizers.py:73-124 over the
 * probabilistic ensemble of config/maze.py:23-96 (PtModel; identical class in config/navigation1.py, navigation2.py). */
typedef struct {
    int32_t plan_hor;      /* config/maze.py:110 (15), navigation1.py:110 (5)                      */
    int32_t popsize;       /* config/maze.py:122-127: 400 candidates ...                           */
    int32_t num_elites;    /*   ... 40 elites                                                      */
    int32_t npart;         /* config/default.py:109: 20 particles per candidate (multiple of num_nets, MPC.py:160) */
    int32_t num_nets;      /* config/default.py:91: 5 bootstrap nets                               */
    int32_t max_iters;     /* 5                                                                    */
    double  alpha;         /* 0.1: mean/var smoothing, optimizers.py:115-116                        */
    double  epsilon;       /* 0.001: stop when max(var) <= epsilon, optimizers.py:91                */
    float   ac_lb[2], ac_ub[2];  /* env.action_space.low / high (MPC.py:119-122)                    */
    uint64_t seed;         /* Philox key (production RNG mode)                                      */
    int32_t stream_id;
    int32_t reserved;
} rrl_mpc_config_t;

/* Packed ensemble ("dyn image"): hidden width zero-padded to 256, every matrix k-major.  rrl_dyn_pack builds it
 * from tensors in the reference's layout (lin*_w [nets][in][out], lin*_b [nets][1][out], inputs_mu/sigma [1][4],
 * max/min_logvar [1][2]; all fp32 device pointers). */
int64_t rrl_dyn_image_floats(void);
int rrl_dyn_pack(const float* lin0_w, const float* lin0_b, const float* lin1_w, const float* lin1_b,
                 const float* lin2_w, const float* lin2_b, const float* lin3_w, const float* lin3_b,
                 const float* inputs_mu, const float* inputs_sigma, const float* max_logvar,
                 const float* min_logvar, int hidden, float* image, void* stream);

/* Ensemble training (MPC.train, MPC.py:268-296).  The trainable parameters live in one flat fp32 "train arena" in the
 * reference's layout, in this order: lin0_w [5][4][200], lin0_b [5][1][200], lin1_w [5][200][200], lin1_b, lin2_w,
 * lin2_b, lin3_w [5][200][4], lin3_b [5][1][4], max_logvar [1][2], min_logvar [1][2]  (rrl_dyn_train_floats() floats);
 * adam_m / adam_v mirror it.  wt: 2*5*200*200 floats (transposed lin1_w / lin2_w, kept by the kernel; rebuild with
 * rrl_dyn_train_sync after the host wrote parameters).  partial: 32 floats scratch.  One rrl_dyn_train_step = one
 * mini-batch of MPC.py:275-296 for all five nets: rows = idx[net][col0 .. col0+rows) of inputs [n][4] / targets [n][2]
 * (device arrays, idx int64 [5][n_idx]); loss = NLL + 0.01*(sum max_logvar - sum min_logvar) + weight decays; torch Adam
 * (lr, 0.9, 0.999, 1e-8); *step (device int64) counts the Adam steps; *ticket (device u32) must be 0. */
int64_t rrl_dyn_train_floats(void);
int rrl_dyn_train_sync(const float* params, float* wt, void* stream);
int rrl_dyn_train_step(float* params, float* adam_m, float* adam_v, float* wt, float* partial, const float* mu,
                       const float* sigma, const float* inputs, const float* targets, const int64_t* idx,
                       int64_t n_idx, int64_t col0, int rows, float lr, int64_t* step, uint32_t* ticket,
                       float* loss_out, void* stream);

/* One MPC.act() for n_envs env copies = rrl_mpc_begin, then max_iters x (rrl_mpc_sample, rrl_mpc_rollout,
 * rrl_mpc_update), then rrl_mpc_finish.  All state is caller-owned device memory:
 *   prev_sol, mean, var : fp64 [n_envs][plan_hor*2]      active : i32 [n_envs] (CEM loop still running)
 *   samples : fp32 [n_envs][popsize][plan_hor*2]          row_cost : fp32 [n_envs][popsize][npart]
 * begin  : mean = prev_sol, var = (ub - lb)^2 / 16                                          (MPC.py:179-181,340)
 * sample : active &= iter < max_iters && max(var) > epsilon; candidates = z * sqrt(min(var, (dist/2)^2)) + mean as fp32
 *          z: fp64 [n_envs][popsize][sol] truncated-normal draws in [-2, 2], or NULL (Philox, inverse CDF)  (optimizers.py:91-101)
 * rollout: every (candidate, particle) through its bootstrap net for plan_hor steps, cost = sum_t max(Q1,Q2)_risk(obs_t, a_t),
 *          NaN -> 1e6; eps: fp32 [n_envs][plan_hor][nets][popsize*npart/nets][2] or NULL (Philox)           (MPC.py:374-439)
 *          state: fp64 [2][n_envs] current observations; the safety critic is read from the agent arena
 * update : cost = mean over particles; elites = num_elites lowest; mean/var <- alpha*old + (1-alpha)*elite stats  (optimizers.py:110-116)
 * finish : action[e] = mean[e][0:2] (fp64); prev_sol[e] = shift(mean[e]) for envs with mask[e] != 0 (NULL = all) (MPC.py:341-345) */
int rrl_mpc_begin(const rrl_mpc_config_t* cfg, int64_t n_envs, const double* prev_sol, double* mean, double* var,
                  int32_t* active, void* stream);
int rrl_mpc_sample(const rrl_mpc_config_t* cfg, int64_t n_envs, int iter, const double* mean, const double* var,
                   const double* z, const int64_t* counters, float* samples, int32_t* active, void* stream);
int rrl_mpc_rollout(const rrl_mpc_config_t* cfg, const rrl_agent_config_t* agent_cfg, const float* arena,
                    const float* dyn_image, int64_t n_envs, const double* state, const float* samples,
                    const float* eps, const int32_t* active, int iter, const int64_t* counters, float* row_cost,
                    void* stream);
int rrl_mpc_update(const rrl_mpc_config_t* cfg, int64_t n_envs, int iter, const float* samples,
                   const float* row_cost, const int32_t* active, double* mean, double* var, void* stream);
int rrl_mpc_finish(const rrl_mpc_config_t* cfg, int64_t n_envs, const double* mean, const uint8_t* mask,
                   double* prev_sol, double* action, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* RRL_H_ */
