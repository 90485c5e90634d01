// tt_rollout.cu -- one whole rollout iteration (DDPG/trainv2.py:511-531 without learn()) as ONE launch
// sequence on the caller's stream:  actor -> OU noise + scaling -> env step -> reset of finished envs (+ OU
// reset, trainv2.py:489-492) -> iteration tick.  The replay store (agent.remember) is FUSED into the producers:
// the actor kernel writes s while it reads the observations (its DRAM pipe is idle), the noise kernel writes the
// raw action, the env kernel writes s', r and done -- 193 B written per transition and nothing re-read, instead of
// a separate scatter kernel that reads and writes 386 B.  (TT_ROLLOUT_UNFUSED=1 selects the separate kernel.)
#include "tt_actor.cuh"
#include "tt_common.cuh"

#include <stdlib.h>
extern "C" {
int tt_env_reset_ou(tt_env *env, const uint8_t *d_mask, float *d_obs, int64_t ld_obs, float *d_ou_x, tt_stream_t stream);
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
    const bool store = b->d_state_mem != nullptr;
    if (store) TT_REQUIRE(b->d_action_mem && b->d_reward_mem && b->d_new_state_mem && b->d_terminal_mem && b->mem_size > 0, "bad ring");
    static const bool unfused_env = [] { const char *e = getenv("TT_ROLLOUT_UNFUSED"); return e && atoi(e) != 0; }();
    const bool tc = precision == TT_PREC_BF16 || precision == TT_PREC_F16 || precision == TT_PREC_F16_PLAIN;
    const bool fused = store && !unfused_env && (!tc || tt::actor_tc_fuses_ring());
    const TTRingMap m = tt_make_ring_map(store ? b->mem_size : 1, store ? b->mem_cntr : 0, store ? n : 0);
    const TTRingS rs = {b->d_state_mem, m};
    const TTRingA ra = {b->d_action_mem, m};
    const tt_replay_ring ring = {b->d_state_mem, b->d_action_mem, b->d_reward_mem, b->d_new_state_mem, b->d_terminal_mem, b->mem_size, b->mem_cntr};
    // DDPG_agent.py:36-49 choose_action (+ DDPG_agent.py:51-52 remember, fused)
    TT_REQUIRE(actor->loaded, "tt_actor_load has not been called");
    if (tc) rc = tt::actor_forward_tc(actor, b->d_obs_cur, b->ld_obs, n, b->d_action, precision, fused ? &rs : nullptr, s);
    else if (precision == TT_PREC_FP32) rc = tt::actor_forward_fp32(actor, b->d_obs_cur, b->ld_obs, n, b->d_action, fused ? &rs : nullptr, s);
    else { tt::set_error("tt_rollout_step: unknown precision %d", precision); rc = TT_ERR_INVALID; }
    if (rc != TT_OK) return rc;
    if ((rc = tt::launch_noise(b->d_ou_x, b->d_action, b->d_scaled, nullptr, n, tt_env_seed_value(env), tt_env_global_offset(env),
                               tt_env_iter_ptr(env), evaluate, fused ? &ra : nullptr, s)) != TT_OK) return rc;
    // simv2.py:499-545 env.step(scaled_action)
    if (fused) rc = tt_env_step_store(env, b->d_scaled, b->d_obs_next, b->ld_obs, b->d_reward, b->d_done, &ring, stream);
    else rc = tt_env_step(env, b->d_scaled, b->d_obs_next, b->ld_obs, b->d_reward, b->d_done, nullptr, stream);
    if (rc != TT_OK) return rc;
    if (store && !fused) {
        if ((rc = tt::replay_store(b->d_state_mem, b->d_action_mem, b->d_reward_mem, b->d_new_state_mem, b->d_terminal_mem,
                                   b->mem_size, b->mem_cntr, b->d_obs_cur, b->ld_obs, b->d_action, b->d_reward, b->d_obs_next,
                                   b->ld_obs, b->d_done, n, s)) != TT_OK) return rc;
    }
    // trainv2.py:489-492: env.reset() + agent.noise.reset() for finished episodes
    if ((rc = tt_env_reset_ou(env, b->d_done, b->d_obs_next, b->ld_obs, evaluate ? nullptr : b->d_ou_x, stream)) != TT_OK) return rc;
    return tt_env_tick(env, 1u, stream);
}
