// tt_rollout.cu -- one whole rollout iteration (DDPG/trainv2.py:489-531 without learn()) as TWO launches on the caller's
// stream:
//   1. the actor kernel (tcgen05 or fp32 by `precision`) with Agent.choose_action's tail fused into its output stage:
//      mu + OU noise (DDPG_agent.py:41-43, noise.py:12-17), clip * pi / 4 (trainv2.py:516), and agent.remember's `state` and
//      raw `action` rows written straight into the replay ring (DDPG_agent.py:51-52);
//   2. the env kernel: simv2.py:499-545 step + reward_functionv1, the ring's `new_state`, `reward`, `terminal` rows, the
//      driver-side `if done: env.reset(); agent.noise.reset()` of the NEXT loop pass (trainv2.py:489-492) from the env's
//      Philox stream, and the iteration tick (last CTA).
// 193 B are written per transition and nothing is re-read; there is no separate noise, scatter, reset or tick launch.
#include "tt_actor.cuh"
#include "tt_common.cuh"

extern "C" {
const uint32_t *tt_env_iter_ptr(tt_env *env);
uint64_t tt_env_seed_value(tt_env *env);
uint64_t tt_env_global_offset(tt_env *env);
int64_t tt_env_num_envs(tt_env *env);
}

extern "C" int tt_rollout_step(tt_env *env, tt_actor *actor, const tt_rollout_bufs *b, int32_t precision, int32_t evaluate,
                               tt_stream_t stream) {
    TT_REQUIRE(env && actor && b, "NULL argument");
    TT_REQUIRE(b->d_obs_cur && b->d_obs_next && b->d_action && b->d_scaled && b->d_reward && b->d_done, "NULL buffer");
    TT_REQUIRE(evaluate || b->d_ou_x, "d_ou_x is NULL");
    const int64_t n = tt_env_num_envs(env);
    const bool store = b->d_state_mem != nullptr;
    if (store) TT_REQUIRE(b->d_action_mem && b->d_reward_mem && b->d_new_state_mem && b->d_terminal_mem && b->mem_size > 0, "bad ring");
    const tt_replay_ring ring = {b->d_state_mem, b->d_action_mem, b->d_reward_mem, b->d_new_state_mem, b->d_terminal_mem, b->mem_size, b->mem_cntr};
    // DDPG_agent.py:36-49 choose_action + trainv2.py:516 scaling (+ DDPG_agent.py:51-52 remember of s, a)
    int rc = tt_actor_choose_action(actor, b->d_obs_cur, b->ld_obs, n, b->d_ou_x, tt_env_seed_value(env), tt_env_global_offset(env),
                                    tt_env_iter_ptr(env), evaluate, b->d_action, b->d_scaled, precision, store ? &ring : nullptr, stream);
    if (rc != TT_OK) return rc;
    // simv2.py:499-545 env.step(scaled_action) (+ remember of s', r, done) + trainv2.py:489-492 reset of finished episodes + tick
    return tt_env_step_reset(env, b->d_scaled, b->d_obs_next, b->ld_obs, b->d_reward, b->d_done, evaluate ? nullptr : b->d_ou_x,
                             store ? &ring : nullptr, stream);
}
