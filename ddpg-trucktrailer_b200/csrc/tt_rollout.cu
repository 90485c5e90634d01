// tt_rollout.cu -- one whole rollout iteration (DDPG/trainv2.py:511-531 without learn()) as ONE launch
// sequence on the caller's stream:  actor -> OU noise + scaling -> env step -> replay store -> reset of
// finished envs (+ OU reset, trainv2.py:489-492) -> iteration tick.
#include "tt_actor.cuh"
#include "tt_common.cuh"

extern "C" {
int tt_env_tick(tt_env *env, uint32_t by, tt_stream_t stream);
const uint32_t *tt_env_iter_ptr(tt_env *env);
uint64_t tt_env_seed_value(tt_env *env);
uint64_t tt_env_global_offset(tt_env *env);
int64_t tt_env_num_envs(tt_env *env);
}
namespace tt {
int replay_store(float *S, float *A, float *R, float *S2, uint8_t *D, int64_t cap, int64_t cntr, const float *s, int64_t ld_s,
                 const float *a, const float *r, const float *s2, int64_t ld_s2, const uint8_t *d, int64_t n, cudaStream_t st);
}

extern "C" int tt_rollout_step(tt_env *env, tt_actor *actor, const tt_rollout_bufs *b, int32_t precision, int32_t evaluate,
                               tt_stream_t stream) {
    TT_REQUIRE(env && actor && b, "NULL argument");
    TT_REQUIRE(b->d_obs_cur && b->d_obs_next && b->d_action && b->d_scaled && b->d_reward && b->d_done, "NULL buffer");
    TT_REQUIRE(evaluate || b->d_ou_x, "d_ou_x is NULL");
    const int64_t n = tt_env_num_envs(env);
    cudaStream_t s = tt::as_stream(stream);
    int rc;
    // DDPG_agent.py:36-49 choose_action
    if ((rc = tt_actor_forward(actor, b->d_obs_cur, b->ld_obs, n, b->d_action, precision, stream)) != TT_OK) return rc;
    if ((rc = tt::launch_noise(b->d_ou_x, b->d_action, b->d_scaled, nullptr, n, tt_env_seed_value(env), tt_env_global_offset(env),
                               tt_env_iter_ptr(env), evaluate, s)) != TT_OK) return rc;
    // simv2.py:499-545 env.step(scaled_action)
    if ((rc = tt_env_step(env, b->d_scaled, b->d_obs_next, b->ld_obs, b->d_reward, b->d_done, nullptr, stream)) != TT_OK) return rc;
    // DDPG_agent.py:51-52 remember(observation, action, reward, observation_, done)
    if (b->d_state_mem) {
        TT_REQUIRE(b->d_action_mem && b->d_reward_mem && b->d_new_state_mem && b->d_terminal_mem && b->mem_size > 0, "bad ring");
        if ((rc = tt::replay_store(b->d_state_mem, b->d_action_mem, b->d_reward_mem, b->d_new_state_mem, b->d_terminal_mem,
                                   b->mem_size, b->mem_cntr, b->d_obs_cur, b->ld_obs, b->d_action, b->d_reward, b->d_obs_next,
                                   b->ld_obs, b->d_done, n, s)) != TT_OK) return rc;
    }
    // trainv2.py:489-492: env.reset() + agent.noise.reset() for finished episodes
    if ((rc = tt_env_reset(env, b->d_done, b->d_obs_next, b->ld_obs, stream)) != TT_OK) return rc;
    if (!evaluate && (rc = tt::launch_ou_zero(b->d_ou_x, b->d_done, n, s)) != TT_OK) return rc;
    return tt_env_tick(env, 1u, stream);
}
