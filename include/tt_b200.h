/*
 * tt_b200.h -- C ABI of libtt_b200.so: the B200 (sm_100a) batched truck-trailer rollout path.
 *
 * The reference (pain7576/ddpg-trucktrailer) is pure Python and has no FFI; its boundary for this path is
 * Python duck-typing.  Each entry point below replaces one reference method for N
 * environments at once; the host-side mirror in ddpg-trucktrailer_b200/{env,agent,replay}.py binds them
 * with ctypes and keeps the reference's method names (INTEGRATION.md shows the stub).
 *
 * Conventions
 *   - extern "C", plain pointers and sizes; no C++ / torch types.
 *   - every pointer named d_* is a DEVICE pointer owned by the caller; nothing is allocated behind the
 *     caller's back: tt_env_create() is handed a device workspace of tt_env_workspace_bytes(n) bytes.
 *   - every launch goes to the caller's stream (a cudaStream_t passed as void*); no call synchronises it.
 *   - return value: 0 = TT_OK, negative = TT_ERR_*; tt_last_error() gives the text (thread-local).
 *   - thread-compatible: one tt_env per host thread / GPU.
 *   - there is NO CPU fallback: without a CUDA device every compute call returns TT_ERR_CUDA.
 *
 * Observation rows are TT_OBS_DIM (23) float32 with a caller-chosen row stride `ld_obs` >= 23 (in floats).
 */
#ifndef TT_B200_H
#define TT_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define TT_ABI_VERSION 1
#define TT_OBS_DIM 23            /* simv2.py:76 observation_dim */
#define TT_NCOMP 10              /* reward components emitted besides total (see tt_step_info) */

enum {
    TT_OK = 0,
    TT_ERR_INVALID = -1,         /* bad argument (null pointer, size, alignment) */
    TT_ERR_CUDA = -2,            /* CUDA runtime error / no device */
    TT_ERR_WORKSPACE = -3        /* workspace too small or misaligned */
};

/* violation_type codes: reward_functionv1.py:378-419, last writer wins in this order */
enum {
    TT_V_NONE = 0, TT_V_JACKKNIFE = 1, TT_V_JACKKNIFE_WARNING = 2, TT_V_MAJOR_BOUNDARY = 3,
    TT_V_MINOR_BOUNDARY = 4, TT_V_PAST_GOAL = 5, TT_V_MAX_STEP = 6, TT_V_EXCESSIVE_BACKWARD = 7
};

/* termination bits in tt_step_info.d_flags: simv2.py:528-541 */
enum {
    TT_F_JACKKNIFE = 1, TT_F_OUT_OF_MAP = 2, TT_F_MAX_STEPS = 4, TT_F_GOAL_REACHED = 8,
    TT_F_GOAL_PASSED = 16, TT_F_EXCESSIVE_BACKWARD = 32
};

/* Environment constants; defaults are the reference literals (simv2.py:23-101, :331-337). */
typedef struct tt_env_cfg {
    double L1, L2, v1x, dt;                       /* 5, 7, -5.012, 0.08 */
    double map_min, map_max;                      /* -40, 40 */
    double max_hitch;                             /* deg2rad(90) */
    double steer_max;                             /* deg2rad(45) */
    double pos_thr, ori_thr;                      /* 0.5, deg2rad(15) */
    double step_len;                              /* 0.40096 (= |v1x| * dt), simv2.py:265 */
    double start_x_lo, start_x_hi;                /* -27, 27   generate_valid_random_poses */
    double start_y_lo, start_y_hi;                /* 0, 27 */
    double start_yaw_lo, start_yaw_hi;            /* deg2rad(45), deg2rad(120) */
    double goal_x, goal_y, goal_yaw;              /* 0, -30, deg2rad(90) */
} tt_env_cfg;

/* Optional per-step diagnostics = the reference's `info` dict (reward_functionv1.py:488-504) as arrays.
 * Any pointer may be NULL. d_comps is SoA: component c of env i at d_comps[c * n_envs + i], order:
 * distance, progress, heading, orientation, staged_success, safety_penalty, exploration_bonus,
 * final_success_bonus, backward_penalty, smoothness_penalty. */
typedef struct tt_step_info {
    float   *d_comps;       /* [TT_NCOMP, n_envs] */
    uint8_t *d_violation;   /* [n_envs] TT_V_* */
    uint8_t *d_flags;       /* [n_envs] TT_F_* bits */
    uint8_t *d_success;     /* [n_envs] info['success'] */
} tt_step_info;

typedef struct tt_env tt_env;         /* opaque */
/* The replay ring (ReplayBuffer, DDPG/replay_buffer.py:4-12) as raw device arrays: dense float32 state / new_state
 * [mem_size,23], action / reward [mem_size], terminal uint8 [mem_size]; mem_cntr = transitions stored so far.  A batch of n
 * transitions goes to rows (mem_cntr + i) % mem_size (store_transition semantics).  Members a call does not write may be NULL. */
typedef struct tt_replay_ring {
    float   *d_state_mem, *d_action_mem, *d_reward_mem, *d_new_state_mem;
    uint8_t *d_terminal_mem;
    int64_t  mem_size, mem_cntr;
} tt_replay_ring;
typedef void *tt_stream_t;            /* cudaStream_t */

const char *tt_last_error(void);
/* Leave n SMs of the current device free: the persistent rollout kernels (actor, env step) size their grids for the
 * remaining SMs, so that the small dependent kernels of a learner (tt_learn_step) or of NCCL issued on a side stream run
 * next to a rollout kernel instead of waiting for its last CTA.  0 (default) = the rollout kernels fill the GPU. */
int tt_reserve_sms(int32_t n);
int tt_abi_version(void);
int tt_device_count(void);            /* 0 when no CUDA device is visible */
uint64_t tt_launch_count(void);       /* kernels launched by this library since it was loaded */

/* ---- environment: Truck_trailer_Env_2 (truck_trailer_sim/simv2.py:20-545) for n_envs instances ---- */
int    tt_env_default_cfg(tt_env_cfg *cfg);                       /* simv2.py:23-101 literals */
size_t tt_env_workspace_bytes(int64_t n_envs);
/* `global_env_offset` = global id of env 0 on this rank; the Philox streams are keyed by (seed, global id)
 * so a given env's trajectory does not depend on how many GPUs the job is sharded over. */
int tt_env_create(tt_env **out, const tt_env_cfg *cfg, int64_t n_envs, uint64_t seed,
                  uint64_t global_env_offset, void *d_workspace, size_t workspace_bytes);
int tt_env_destroy(tt_env *env);
int tt_env_seed(tt_env *env, uint64_t seed, tt_stream_t stream);  /* reset(seed=...): simv2.py:460-462 */

/* reset(): simv2.py:459-498.  d_mask NULL = all envs, else only envs with d_mask[i] != 0 (driver-side reset
 * on done, trainv2.py:489).  Writes the reset observation rows of the affected envs to d_obs. */
int tt_env_reset(tt_env *env, const uint8_t *d_mask, float *d_obs, int64_t ld_obs, tt_stream_t stream);

/* step(): simv2.py:499-545.  d_action = scaled steering [n_envs] (what trainv2.py:516-520 passes).
 * d_obs receives the post-step observation (the terminal one when done), d_reward float32, d_done 0/1.
 * An env that is already done and has not been reset is frozen: reward 0, done 1, same observation. */
int tt_env_step(tt_env *env, const float *d_action, float *d_obs, int64_t ld_obs, float *d_reward,
                uint8_t *d_done, const tt_step_info *info, tt_stream_t stream);

/* K consecutive steps in ONE launch, state held in registers: d_actions[K, n_envs].  Per-step outputs are
 * optional (NULL) and laid out [K, ...]; with auto_reset != 0 a done env is re-initialised from its Philox
 * stream inside the kernel (its d_obs row for that step is still the terminal observation). */
int tt_env_step_k(tt_env *env, const float *d_actions, int32_t K, int32_t auto_reset, float *d_obs,
                  int64_t ld_obs, float *d_reward, uint8_t *d_done, const tt_step_info *info,
                  tt_stream_t stream);

/* State injection / readback (test.py:96-115, heatmap.py:119, episode_replay_collectorv2.py:258).
 * d_idx NULL = envs 0..n-1.  d_state [n,6] = psi1,psi2,x1,y1,x2,y2; d_start [n,3] = startx,starty,startyaw;
 * d_goal [n,3] = goalx,goaly,goalyaw.  Starts a fresh episode for those envs and writes their obs rows. */
int tt_env_set_state(tt_env *env, const int64_t *d_idx, int64_t n, const double *d_state,
                     const double *d_start, const double *d_goal, float *d_obs, int64_t ld_obs,
                     tt_stream_t stream);
int tt_env_get_state(tt_env *env, double *d_state, double *d_start, double *d_goal, int32_t *d_steps,
                     int32_t *d_max_steps, tt_stream_t stream);

/* The reward function's persistent per-episode state (reward_functionv1.py:99-109, threaded through simv2.py:347-373):
 * closest_distance_to_goal, cumulative_backward_movement, the frozen first steering angle (previous_steering), and the
 * running episode return (trainv2.py:529 `score`).  float32 [n_envs]; any pointer may be NULL. */
int tt_env_get_reward_state(tt_env *env, float *d_closest, float *d_cum_backward, float *d_first_steer, float *d_episode_return,
                            tt_stream_t stream);

/* env.L2 = value per environment (heatmap.py:89 draws the trailer length per trial: `env.L2 = np.random.uniform(5,7)`).
 * d_idx NULL = envs 0..n-1; d_l2 [n] float64 > 0.  The dynamics (simv2.py:291) use the new length from the next step on,
 * reset / set_state place the truck L2 ahead of the trailer (simv2.py:483-484, heatmap.py:116-117).  Does not restart
 * the episode.  tt_env_get_l2 writes all n_envs lengths. */
int tt_env_set_l2(tt_env *env, const int64_t *d_idx, int64_t n, const double *d_l2, tt_stream_t stream);
int tt_env_get_l2(tt_env *env, double *d_l2, tt_stream_t stream);

/* Optional bit-packed copy of `done` for hosts that read the flags back every step (1 bit instead of 1 byte per env over PCIe):
 * while d_bits != NULL every single-step launch (tt_env_step, tt_env_step_store, tt_env_step_reset, tt_rollout_step; not
 * tt_env_step_k with K > 1) also writes word i / 32, bit i % 32 = done of env i (what np.packbits(done, bitorder='little') of the
 * byte array gives, read as little-endian uint32).  d_bits: ceil(n_envs / 32) uint32; NULL switches it off.  The pointer is a
 * by-value launch argument: set it before the launches are captured into a CUDA graph. */
int tt_env_set_done_bits(tt_env *env, uint32_t *d_bits);

/* Iteration counter of the Philox streams (device-resident so that launch sequences stay graph-capturable).
 * tt_env_step does not advance it; a driver that steps and resets by hand calls tt_env_tick once per
 * iteration (tt_rollout_step, tt_env_step_k with auto_reset and a full tt_env_reset do it themselves). */
int tt_env_tick(tt_env *env, uint32_t by, tt_stream_t stream);
const uint32_t *tt_env_iter_ptr(tt_env *env);          /* device pointer */
uint64_t tt_env_seed_value(tt_env *env);
uint64_t tt_env_global_offset(tt_env *env);
int64_t  tt_env_num_envs(tt_env *env);

/* Per-rank rollout statistics accumulated on the device since the last read:
 * [0] env steps, [1] episodes finished, [2] successes, [3] sum of episode returns, [4] sum of squares,
 * [5] sum of rewards, [6..11] termination-reason counts (TT_F_* order), [12..15] reserved. */
#define TT_NSTATS 16
int tt_env_stats_read(tt_env *env, double *d_out16, int32_t clear, tt_stream_t stream);

/* ---- OU noise: OUActionNoise.__call__/reset (DDPG/noise.py:12-20) + DDPG_agent.py:41-43 ---- */
/* x <- x + 0.2 (0 - x) 0.01 + 0.15 * 0.1 * N(0,1) with N from Philox(seed; global id, *d_iter, stream 1);
 * d_action (may be NULL) += x.  d_reset_mask (may be NULL): x is zeroed first where mask != 0. */
int tt_ou_step(float *d_x, float *d_action, const uint8_t *d_reset_mask, int64_t n, uint64_t seed,
               uint64_t global_env_offset, const uint32_t *d_iter, tt_stream_t stream);

/* ---- actor: ActorNetwork.forward (DDPG/networks.py:138-147) ---- */
typedef struct tt_actor tt_actor;     /* opaque: packed device weights */
/* TT_PREC_FP32: warp-level fp32 FMA on the CUDA cores (<= 1e-5 of torch fp32, any layer sizes).
 * TT_PREC_F16: tcgen05 tensor cores (layer sizes 23-400-300), fp16 operands with an exact (split hi/lo) first layer, fp32
 * accumulation: <= 1e-3 even for strongly amplified trained weights -- the tensor-core mode that holds north_star's bar.
 * TT_PREC_AUTO: TT_PREC_F16 once the batch is a real dense contraction (n >= 192 rows and 23-400-300 layers), TT_PREC_FP32
 * below (tt_actor_auto_precision tells which).
 * OUT-OF-BAR modes (faster, plain 16-bit operands in both layers; within 1e-3 on reference-scale weights only, 1.2e-3 /
 * 1e-2 on strongly amplified ones -- the Python classes refuse them without allow_out_of_bar=True):
 * TT_PREC_F16_PLAIN = plain fp16 operands, TT_PREC_BF16 = plain bf16 operands. */
enum { TT_PREC_FP32 = 0, TT_PREC_BF16 = 1, TT_PREC_F16 = 2, TT_PREC_F16_PLAIN = 3, TT_PREC_AUTO = 4 };
size_t tt_actor_workspace_bytes(int32_t in_dim, int32_t h1, int32_t h2);
int tt_actor_create(tt_actor **out, int32_t in_dim, int32_t h1, int32_t h2, void *d_workspace,
                    size_t workspace_bytes);
int tt_actor_destroy(tt_actor *a);
/* Weights in the reference state_dict layout (float32, device pointers): fc1.weight[h1,in] fc1.bias[h1]
 * bn1.weight[h1] bn1.bias[h1] fc2.weight[h2,h1] fc2.bias[h2] bn2.weight[h2] bn2.bias[h2] mu.weight[1,h2]
 * mu.bias[1] -- re-packed into the kernels' layouts (transposed fp32 / bf16 UMMA tiles) on `stream`. */
int tt_actor_load(tt_actor *a, const float *d_fc1_w, const float *d_fc1_b, const float *d_ln1_g,
                  const float *d_ln1_b, const float *d_fc2_w, const float *d_fc2_b, const float *d_ln2_g,
                  const float *d_ln2_b, const float *d_mu_w, const float *d_mu_b, tt_stream_t stream);
/* d_mu[n] = tanh(mu(relu(LN(fc2(relu(LN(fc1(obs)))))))).  If d_scaled != NULL it also receives
 * clip(mu, -1, 1) * float32(pi/4) (trainv2.py:516).  precision: one of TT_PREC_*. */
int tt_actor_forward(tt_actor *a, const float *d_obs, int64_t ld_obs, int64_t n, float *d_mu,
                     int32_t precision, tt_stream_t stream);

/* The concrete precision TT_PREC_AUTO resolves to for a batch of n rows (TT_PREC_FP32 or TT_PREC_F16). */
int tt_actor_auto_precision(tt_actor *a, int64_t n);

/* Agent.choose_action (DDPG/DDPG_agent.py:36-49) in ONE launch: d_action[n] = mu(obs) + OU noise (unless evaluate; the OU
 * state d_ou_x[n] is advanced in place with N(0,1) from Philox(seed; global id, *d_iter, stream 1)) -- the UNCLIPPED action
 * the reference returns -- and, if d_scaled != NULL, clip(action, -1, 1) * float32(pi/4) (trainv2.py:516).  With a ring
 * (may be NULL) the observation rows and the raw actions are also written as the `state` / `action` part of the n
 * transitions (agent.remember, DDPG_agent.py:51-52).  Bit-identical to tt_actor_forward + tt_ou_step + tt_scale_action. */
int tt_actor_choose_action(tt_actor *a, const float *d_obs, int64_t ld_obs, int64_t n, float *d_ou_x, uint64_t seed,
                           uint64_t global_env_offset, const uint32_t *d_iter, int32_t evaluate, float *d_action,
                           float *d_scaled, int32_t precision, const tt_replay_ring *ring, tt_stream_t stream);

/* clip(a, -1, 1) * float32(pi/4): the driver-side scaling at trainv2.py:516 */
int tt_scale_action(const float *d_action, float *d_scaled, int64_t n, tt_stream_t stream);

/* ---- replay: ReplayBuffer.store_transition (DDPG/replay_buffer.py:13-21) for n transitions ---- */
/* Equivalent to n sequential store_transition calls in env order: row (mem_cntr + i) % mem_size, last
 * writer wins.  Ring arrays are dense float32: state/new_state [mem_size,23], action/reward [mem_size],
 * terminal uint8 [mem_size].  The caller advances mem_cntr by n. */
int tt_replay_store(float *d_state_mem, float *d_action_mem, float *d_reward_mem, float *d_new_state_mem,
                    uint8_t *d_terminal_mem, int64_t mem_size, int64_t mem_cntr, const float *d_s,
                    int64_t ld_s, const float *d_a, const float *d_r, const float *d_s2, int64_t ld_s2,
                    const uint8_t *d_done, int64_t n, tt_stream_t stream);
/* sample_buffer (replay_buffer.py:23-34) gather for given row indices */
int tt_replay_gather(const float *d_state_mem, const float *d_action_mem, const float *d_reward_mem,
                     const float *d_new_state_mem, const uint8_t *d_terminal_mem, const int64_t *d_rows,
                     int64_t batch, float *d_s, float *d_a, float *d_r, float *d_s2, uint8_t *d_done,
                     tt_stream_t stream);

/* ---- fused store: the producers write their part of the transition straight into the ring ---- */
/* A batch of n transitions goes to rows (mem_cntr + i) % mem_size exactly as tt_replay_store would put it.  The three
 * calls below are the producers of tt_rollout_step with the store fused in (193 B written per transition, nothing
 * re-read): the actor writes `state` while it reads the observations, the noise kernel writes the raw `action`, the
 * env kernel writes `new_state`, `reward` and `terminal`.  Members that a call does not write may be NULL. */
int tt_actor_forward_store(tt_actor *a, const float *d_obs, int64_t ld_obs, int64_t n, float *d_mu, int32_t precision,
                           const tt_replay_ring *ring, tt_stream_t stream);
int tt_ou_step_store(float *d_x, float *d_action, float *d_scaled, int64_t n, uint64_t seed, uint64_t global_env_offset,
                     const uint32_t *d_iter, int32_t evaluate, const tt_replay_ring *ring, tt_stream_t stream);
int tt_env_step_store(tt_env *env, const float *d_action, float *d_obs, int64_t ld_obs, float *d_reward,
                      uint8_t *d_done, const tt_replay_ring *ring, tt_stream_t stream);
/* env.step (simv2.py:499-545) + the driver's `if done: env.reset(); agent.noise.reset()` (trainv2.py:489-492) + the Philox
 * iteration tick in ONE launch -- the env half of a rollout iteration.  d_obs rows of finished envs receive the RESET
 * observation of their next episode (pose from Philox(seed; global id, iteration, stream 0)), their TERMINAL observation goes
 * to the ring's new_state row (ring may be NULL), their OU state d_ou_x[i] (may be NULL) is zeroed; d_done still reports
 * the finished episode.  Bit-identical to tt_env_step_store + tt_env_reset(mask = done) + zeroing + tt_env_tick. */
int tt_env_step_reset(tt_env *env, const float *d_action, float *d_obs, int64_t ld_obs, float *d_reward, uint8_t *d_done,
                      float *d_ou_x, const tt_replay_ring *ring, tt_stream_t stream);

/* ---- one whole rollout iteration (trainv2.py:511-531 without learn()) as one launch sequence ---- */
typedef struct tt_rollout_bufs {
    float   *d_obs_cur;      /* [n, ld_obs] in: s   */
    float   *d_obs_next;     /* [n, ld_obs] out: s' (rows of finished envs then hold the reset observation) */
    int64_t  ld_obs;
    float   *d_ou_x;         /* [n] OU state */
    float   *d_action;       /* [n] out: raw actor + noise (what agent.remember stores) */
    float   *d_scaled;       /* [n] scratch: clip * pi/4 */
    float   *d_reward;       /* [n] out */
    uint8_t *d_done;         /* [n] out */
    /* replay ring (all NULL = do not store) */
    float   *d_state_mem, *d_action_mem, *d_reward_mem, *d_new_state_mem;
    uint8_t *d_terminal_mem;
    int64_t  mem_size, mem_cntr;
} tt_rollout_bufs;
/* tt_actor_choose_action (actor + OU noise unless evaluate + scaling + store of s, a) -> tt_env_step_reset (env step + store of
 * s', r, done + reset of finished envs + tick): two launches.  Advances the env's iteration counter. */
int tt_rollout_step(tt_env *env, tt_actor *actor, const tt_rollout_bufs *b, int32_t precision,
                    int32_t evaluate, tt_stream_t stream);

/* ---- learner: Agent.learn (DDPG/DDPG_agent.py:72-131) on the device-resident ring ---- */
/* One tt_learner holds the four networks of the reference Agent (actor, target_actor, critic, target_critic:
 * DDPG/networks.py:9-68, :98-147), both Adam states and the batch workspace inside the caller's workspace.  Parameters
 * of a network are ONE flat float32 vector, tensors in this order (shapes as in the reference state_dict):
 *   actor  (net 0, target 1): fc1.weight[h1,in] fc1.bias[h1] bn1.weight[h1] bn1.bias[h1] fc2.weight[h2,h1] fc2.bias[h2]
 *                             bn2.weight[h2] bn2.bias[h2] mu.weight[1,h2] mu.bias[1]
 *   critic (net 2, target 3): the same eight trunk tensors, then action_value.weight[h2,1] action_value.bias[h2]
 *                             q.weight[1,h2] q.bias[1]
 * tt_learner_params returns the device pointer (inside the workspace) the caller reads / writes to exchange weights. */
typedef struct tt_learner tt_learner;
size_t tt_learner_workspace_bytes(int32_t in_dim, int32_t h1, int32_t h2, int32_t batch);
int tt_learner_create(tt_learner **out, int32_t in_dim, int32_t h1, int32_t h2, int32_t batch, float alpha, float beta,
                      float gamma, float tau, float critic_weight_decay, uint64_t seed, void *d_workspace,
                      size_t workspace_bytes);
int tt_learner_destroy(tt_learner *ln);
float  *tt_learner_params(tt_learner *ln, int32_t net);          /* 0 actor, 1 target_actor, 2 critic, 3 target_critic */
int64_t tt_learner_param_count(tt_learner *ln, int32_t net);
float  *tt_learner_grads(tt_learner *ln, int32_t which);         /* gradients of the last step: 0 actor, 1 critic (same layout) */
const float   *tt_learner_last_q(tt_learner *ln);                /* [batch] critic(s, a) of the last step */
const int64_t *tt_learner_last_rows(tt_learner *ln);             /* [batch] ring rows of the last step */
int tt_learner_reset_optimizer(tt_learner *ln, tt_stream_t stream);   /* zero both Adam states and the step counter */
/* One Agent.learn(): sample `batch` rows uniformly with replacement from the filled part of the ring
 * (replay_buffer.py:23-34; Philox(seed; b, step, stream 2) -- or the given d_rows[batch] when not NULL), critic update
 * (MSE on r + gamma Q'(s', pi'(s')) with terminal masking, Adam with weight decay), actor update (ascent on Q(s, pi(s))
 * through the UPDATED critic, Adam), soft update of both targets (tau).  If repack_into != NULL the new actor parameters are
 * then re-packed into that rollout actor (tt_actor_load).  14 kernel launches (+ 2 of the re-pack) on `stream`, no
 * synchronisation, graph-capturable (the step counter lives in device memory), deterministic. */
int tt_learn_step(tt_learner *ln, const tt_replay_ring *ring, const int64_t *d_rows, tt_actor *repack_into,
                  tt_stream_t stream);
/* The same update with the batch sampled from `window_count` ring rows starting at row `window_begin` (wrapping around):
 * a learner that runs on a side stream CONCURRENTLY with a rollout iteration samples only the rows earlier iterations
 * completed, not the ones the running iteration is writing. */
int tt_learn_step_window(tt_learner *ln, const tt_replay_ring *ring, const int64_t *d_rows, tt_actor *repack_into,
                         int64_t window_begin, int64_t window_count, tt_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* TT_B200_H */
