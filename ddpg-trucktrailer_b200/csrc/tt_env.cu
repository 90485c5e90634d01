// tt_env.cu -- kernel (a): the fused truck-trailer environment step for N independent environments, plus
// reset / state injection / readback / statistics.  Replaces Truck_trailer_Env_2.reset/step
// (truck_trailer_sim/simv2.py:459-545) and RewardFunction (reward_functionv1.py) for the whole batch.
//
// Layout in HBM (struct-of-arrays inside the caller-provided workspace, every array 256 B aligned):
//   st[6][N] f64   psi1 psi2 x1 y1 x2 y2            goal[N] float4   gx gy sin(gyaw) cos(gyaw)
//   rsA[N] float4  closest cum first_steer ep_return rsB[N]  float4   g1 g2 g3 d0   (g = distance decrements)
//   packed[N] u32  steps|emax|rmax-emax|stage bits|finished
//   pose[4][N] f64 startx starty startyaw goalyaw   (written on reset only; host-visible attributes)
//   stats[16] f64, iter u32
// One thread owns one environment: all loads/stores are unit-stride across the warp (8 or 16 B per lane).
// The 23-float observation rows ([N, ld_obs] row-major, the layout the actor and the replay ring consume)
// are transposed through shared memory so that the global stores are coalesced.
#include <new>
#include <stdlib.h>
#include <string.h>
#include "tt_common.cuh"
#include "tt_consts.h"
#include "tt_env_math.cuh"

using namespace ttm;

struct EnvPtrs {
    double *st;          // [6][N]
    float4 *goal, *rsA, *rsB;
    uint32_t *packed;
    double *pose;        // [4][N]
    double *stats;       // [16]
    uint32_t *iter;
    int64_t N;
};

struct tt_env {
    tt_env_cfg cfg;
    StepConsts k;
    EnvPtrs p;
    uint64_t seed;
    uint64_t gid0;
};

namespace {

constexpr int kBlock = 128;
#ifndef TT_ENV_MINBLOCKS_DEFAULT
#define TT_ENV_MINBLOCKS_DEFAULT 4
#endif

__device__ __forceinline__ void load_regs(const EnvPtrs &p, int64_t i, EnvRegs &e) {
    const int64_t N = p.N;
    e.psi1 = p.st[i]; e.psi2 = p.st[N + i]; e.x1 = p.st[2 * N + i]; e.y1 = p.st[3 * N + i];
    e.x2 = p.st[4 * N + i]; e.y2 = p.st[5 * N + i];
    const float4 g = __ldg(&p.goal[i]);
    e.gx = g.x; e.gy = g.y; e.sgy = g.z; e.cgy = g.w;
    const float4 a = p.rsA[i], b = p.rsB[i];
    e.closest = a.x; e.cum = a.y; e.first_steer = a.z; e.ep_ret = a.w;
    e.g1 = b.x; e.g2 = b.y; e.g3 = b.z; e.d0 = b.w;
    e.packed = p.packed[i];
}

__device__ __forceinline__ void store_dyn(const EnvPtrs &p, int64_t i, const EnvRegs &e) {
    const int64_t N = p.N;
    p.st[i] = e.psi1; p.st[N + i] = e.psi2; p.st[2 * N + i] = e.x1; p.st[3 * N + i] = e.y1;
    p.st[4 * N + i] = e.x2; p.st[5 * N + i] = e.y2;
    p.rsA[i] = make_float4(e.closest, e.cum, e.first_steer, e.ep_ret);
    p.rsB[i] = make_float4(e.g1, e.g2, e.g3, e.d0);
    p.packed[i] = e.packed;
}

__device__ __forceinline__ void store_episode_consts(const EnvPtrs &p, int64_t i, const EnvRegs &e, double sx, double sy,
                                                     double syaw, double gyaw) {
    const int64_t N = p.N;
    p.goal[i] = make_float4(e.gx, e.gy, e.sgy, e.cgy);
    p.pose[i] = sx; p.pose[N + i] = sy; p.pose[2 * N + i] = syaw; p.pose[3 * N + i] = gyaw;
}

// coalesced store of a [rows, 23] tile held in shared memory to obs[(row0 + r) * ld + c]
__device__ __forceinline__ void store_obs_tile(const float *tile, float *obs, int64_t ld, int64_t row0, int rows) {
    if (ld == TT_OBS_DIM && rows == kBlock && ((reinterpret_cast<uintptr_t>(obs + row0 * TT_OBS_DIM) & 15) == 0)) {
        float4 *dst = reinterpret_cast<float4 *>(obs + row0 * TT_OBS_DIM);
        const float4 *src = reinterpret_cast<const float4 *>(tile);
#pragma unroll
        for (int v = threadIdx.x; v < kBlock * TT_OBS_DIM / 4; v += kBlock) __stcs(&dst[v], src[v]);
    } else {
        for (int v = threadIdx.x; v < rows * TT_OBS_DIM; v += kBlock) {
            const int r = v / TT_OBS_DIM, c = v - r * TT_OBS_DIM;
            obs[(row0 + r) * ld + c] = tile[v];
        }
    }
}

struct StatAcc {
    float steps, episodes, successes, ret, ret2, rew;
    float fl[6];
};

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// K env steps per launch; state stays in registers across the K steps.
template <bool kInfo, int kMinBlocks>
__global__ void __launch_bounds__(kBlock, kMinBlocks) env_step_kernel(EnvPtrs p, StepConsts k, const float *__restrict__ actions,
                                                          int K, int auto_reset, float *__restrict__ obs, int64_t ld,
                                                          float *__restrict__ reward, uint8_t *__restrict__ done,
                                                          tt_step_info info, uint64_t seed, uint64_t gid0) {
    __shared__ __align__(16) float tile[kBlock * TT_OBS_DIM];
    __shared__ float sstat[12][kBlock / 32];
    const int64_t N = p.N;
    const int64_t row0 = (int64_t)blockIdx.x * kBlock;
    const int64_t i = row0 + threadIdx.x;
    const bool active = i < N;
    const int rows = (int)((N - row0) < kBlock ? (N - row0) : kBlock);
    const uint32_t t0 = *p.iter;

    EnvRegs e;
    if (active) load_regs(p, i, e);
    StatAcc sa;
    sa.steps = sa.episodes = sa.successes = sa.ret = sa.ret2 = sa.rew = 0.f;
#pragma unroll
    for (int f = 0; f < 6; f++) sa.fl[f] = 0.f;

    for (int j = 0; j < K; j++) {
        StepOut o;
        if (active) {
            if (e.packed & PK_FINISHED) {
                // frozen: no reset since `done` -> reward 0, done 1, observation of the frozen state, steering 0
                float s2, c2, s1, c1;
                sincos_f32_of_f64(e.psi2, s2, c2);
                sincos_f32_of_f64(e.psi1, s1, c1);
                pack_obs(k, e, s1, c1, s2, c2, fmaf(s1, c2, -c1 * s2), fmaf(c1, c2, s1 * s2), 0.0f, 1.0f, o.obs);
                o.reward = 0.f; o.done = true; o.success = false; o.flags = 0u; o.viol = 0u;
#pragma unroll
                for (int c = 0; c < TT_NCOMP; c++) o.comps[c] = 0.f;
            } else {
                const float a = __ldcs(&actions[(int64_t)j * N + i]);
                env_step<kInfo>(k, e, a, o);
                sa.steps += 1.f; sa.rew += o.reward;
                if (o.done) {
                    sa.episodes += 1.f; sa.successes += o.success ? 1.f : 0.f;
                    sa.ret += e.ep_ret; sa.ret2 += e.ep_ret * e.ep_ret;
#pragma unroll
                    for (int f = 0; f < 6; f++) sa.fl[f] += (o.flags >> f) & 1u ? 1.f : 0.f;
                }
            }
            const int64_t oi = (int64_t)j * N + i;
            if (reward) __stcs(&reward[oi], o.reward);
            if (done) done[oi] = o.done ? 1 : 0;
            if (kInfo) {
                if (info.d_comps) {
#pragma unroll
                    for (int c = 0; c < TT_NCOMP; c++) info.d_comps[((int64_t)j * TT_NCOMP + c) * N + i] = o.comps[c];
                }
                if (info.d_violation) info.d_violation[oi] = (uint8_t)o.viol;
                if (info.d_flags) info.d_flags[oi] = (uint8_t)o.flags;
                if (info.d_success) info.d_success[oi] = o.success ? 1 : 0;
            }
            if (obs) {
#pragma unroll
                for (int c = 0; c < TT_OBS_DIM; c++) tile[threadIdx.x * TT_OBS_DIM + c] = o.obs[c];
            }
            if (o.done && !(e.packed & PK_FINISHED)) {
                if (auto_reset) {
                    double sx, sy, syaw;
                    rng_pose(k, seed, (uint32_t)(gid0 + i), t0 + (uint32_t)j, sx, sy, syaw);
                    reset_from_pose(k, e, sx, sy, syaw, k.gx, k.gy, k.gyaw, nullptr);
                    store_episode_consts(p, i, e, sx, sy, syaw, k.gyaw);
                } else e.packed |= PK_FINISHED;
            }
        }
        if (obs) {
            __syncthreads();
            store_obs_tile(tile, obs + (int64_t)j * N * ld, ld, row0, rows);
            __syncthreads();
        }
    }
    if (active) store_dyn(p, i, e);

    // block-level statistics -> 12 atomics per block
    float v[12] = {sa.steps, sa.episodes, sa.successes, sa.ret, sa.ret2, sa.rew,
                   sa.fl[0], sa.fl[1], sa.fl[2], sa.fl[3], sa.fl[4], sa.fl[5]};
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
#pragma unroll
    for (int s = 0; s < 12; s++) {
        const float r = warp_sum(v[s]);
        if (lane == 0) sstat[s][wid] = r;
    }
    __syncthreads();
    if (threadIdx.x < 12) {
        double tot = 0.0;
#pragma unroll
        for (int w = 0; w < kBlock / 32; w++) tot += (double)sstat[threadIdx.x][w];
        if (tot != 0.0) atomicAdd(&p.stats[threadIdx.x], tot);
    }
}

// reset(): mask == nullptr -> every env; else only where mask[i] != 0.  Also clears the OU state is NOT done
// here (that is tt_ou_step's reset mask, trainv2.py:492).
__global__ void __launch_bounds__(kBlock) env_reset_kernel(EnvPtrs p, StepConsts k, const uint8_t *__restrict__ mask,
                                                           float *__restrict__ obs, int64_t ld, uint64_t seed,
                                                           uint64_t gid0, uint32_t t_salt) {
    const int64_t i = (int64_t)blockIdx.x * kBlock + threadIdx.x;
    if (i >= p.N) return;
    if (mask && !mask[i]) return;
    EnvRegs e;
    double sx, sy, syaw;
    rng_pose(k, seed, (uint32_t)(gid0 + i), *p.iter + t_salt, sx, sy, syaw);
    float o[TT_OBS_DIM];
    reset_from_pose(k, e, sx, sy, syaw, k.gx, k.gy, k.gyaw, obs ? o : nullptr);
    store_dyn(p, i, e);
    store_episode_consts(p, i, e, sx, sy, syaw, k.gyaw);
    if (obs) {
#pragma unroll
        for (int c = 0; c < TT_OBS_DIM; c++) obs[i * ld + c] = o[c];
    }
}

__global__ void __launch_bounds__(kBlock) env_set_state_kernel(EnvPtrs p, StepConsts k, const int64_t *__restrict__ idx,
                                                               int64_t n, const double *__restrict__ state,
                                                               const double *__restrict__ start,
                                                               const double *__restrict__ goal, float *__restrict__ obs,
                                                               int64_t ld) {
    const int64_t j = (int64_t)blockIdx.x * kBlock + threadIdx.x;
    if (j >= n) return;
    const int64_t i = idx ? idx[j] : j;
    if (i < 0 || i >= p.N) return;
    EnvRegs e;
    e.psi1 = state[6 * j]; e.psi2 = state[6 * j + 1]; e.x1 = state[6 * j + 2]; e.y1 = state[6 * j + 3];
    e.x2 = state[6 * j + 4]; e.y2 = state[6 * j + 5];
    float o[TT_OBS_DIM];
    begin_episode(k, e, start[3 * j], start[3 * j + 1], goal[3 * j], goal[3 * j + 1], goal[3 * j + 2], obs ? o : nullptr);
    store_dyn(p, i, e);
    store_episode_consts(p, i, e, start[3 * j], start[3 * j + 1], start[3 * j + 2], goal[3 * j + 2]);
    if (obs) {
#pragma unroll
        for (int c = 0; c < TT_OBS_DIM; c++) obs[i * ld + c] = o[c];
    }
}

__global__ void __launch_bounds__(kBlock) env_get_state_kernel(EnvPtrs p, double *__restrict__ state,
                                                               double *__restrict__ start, double *__restrict__ goal,
                                                               int32_t *__restrict__ steps, int32_t *__restrict__ max_steps) {
    const int64_t i = (int64_t)blockIdx.x * kBlock + threadIdx.x;
    const int64_t N = p.N;
    if (i >= N) return;
    if (state) {
#pragma unroll
        for (int c = 0; c < 6; c++) state[6 * i + c] = p.st[c * N + i];
    }
    if (start) { start[3 * i] = p.pose[i]; start[3 * i + 1] = p.pose[N + i]; start[3 * i + 2] = p.pose[2 * N + i]; }
    if (goal) { const float4 g = p.goal[i]; goal[3 * i] = g.x; goal[3 * i + 1] = g.y; goal[3 * i + 2] = p.pose[3 * N + i]; }
    const uint32_t pk = p.packed[i];
    if (steps) steps[i] = (int32_t)(pk & PK_STEPS_MASK);
    if (max_steps) max_steps[i] = (int32_t)((pk >> PK_EMAX_SHIFT) & PK_EMAX_MASK);
}

__global__ void tick_kernel(uint32_t *iter, uint32_t by) { *iter += by; }

__global__ void stats_read_kernel(double *stats, double *out, int clear) {
    const int t = threadIdx.x;
    if (t < TT_NSTATS) {
        out[t] = stats[t];
        if (clear) stats[t] = 0.0;
    }
}

int64_t grid_for(int64_t n) { return (n + kBlock - 1) / kBlock; }

}  // namespace

// ------------------------------------------------------------------------------------------------ C ABI ----
extern "C" {

int tt_env_default_cfg(tt_env_cfg *cfg) {
    TT_REQUIRE(cfg, "cfg is NULL");
    tt_fill_default_cfg(cfg);
    return TT_OK;
}

static size_t env_layout(int64_t n, EnvPtrs *p, char *base) {
    size_t off = 0;
    auto take = [&](size_t bytes) { size_t o = off; off = tt::align_up(off + bytes, 256); return o; };
    const size_t o_st = take(sizeof(double) * 6 * n), o_goal = take(sizeof(float4) * n), o_a = take(sizeof(float4) * n),
                 o_b = take(sizeof(float4) * n), o_pk = take(sizeof(uint32_t) * n),
                 o_pose = take(sizeof(double) * 4 * n), o_stats = take(sizeof(double) * TT_NSTATS), o_iter = take(256);
    if (p) {
        p->st = reinterpret_cast<double *>(base + o_st); p->goal = reinterpret_cast<float4 *>(base + o_goal);
        p->rsA = reinterpret_cast<float4 *>(base + o_a); p->rsB = reinterpret_cast<float4 *>(base + o_b);
        p->packed = reinterpret_cast<uint32_t *>(base + o_pk);
        p->pose = reinterpret_cast<double *>(base + o_pose); p->stats = reinterpret_cast<double *>(base + o_stats);
        p->iter = reinterpret_cast<uint32_t *>(base + o_iter); p->N = n;
    }
    return off;
}

size_t tt_env_workspace_bytes(int64_t n_envs) { return n_envs > 0 ? env_layout(n_envs, nullptr, nullptr) : 0; }

int tt_env_create(tt_env **out, const tt_env_cfg *cfg, int64_t n_envs, uint64_t seed, uint64_t global_env_offset,
                  void *d_workspace, size_t workspace_bytes) {
    TT_REQUIRE(out && cfg && d_workspace, "NULL argument");
    TT_REQUIRE(n_envs > 0 && n_envs < (int64_t(1) << 31), "n_envs out of range");
    TT_REQUIRE(global_env_offset + (uint64_t)n_envs <= (uint64_t(1) << 32), "global env id must fit 32 bits");
    if ((reinterpret_cast<uintptr_t>(d_workspace) & 255) != 0 || workspace_bytes < tt_env_workspace_bytes(n_envs)) {
        tt::set_error("tt_env_create: workspace must be 256 B aligned and >= %zu bytes", tt_env_workspace_bytes(n_envs));
        return TT_ERR_WORKSPACE;
    }
    if (tt_device_count() <= 0) { tt::set_error("tt_env_create: no CUDA device (there is no CPU fallback)"); return TT_ERR_CUDA; }
    tt_env *e = new (std::nothrow) tt_env;
    TT_REQUIRE(e, "out of host memory");
    e->cfg = *cfg; e->k = tt_make_consts(*cfg); e->seed = seed; e->gid0 = global_env_offset;
    env_layout(n_envs, &e->p, static_cast<char *>(d_workspace));
    cudaError_t err = cudaMemset(d_workspace, 0, tt_env_workspace_bytes(n_envs));
    if (err != cudaSuccess) { delete e; return tt::cuda_fail(err, "cudaMemset(workspace)"); }
    *out = e;
    return TT_OK;
}

int tt_env_destroy(tt_env *env) { delete env; return TT_OK; }

int tt_env_seed(tt_env *env, uint64_t seed, tt_stream_t stream) {
    TT_REQUIRE(env, "env is NULL");
    env->seed = seed;
    TT_CUDA(cudaMemsetAsync(env->p.iter, 0, sizeof(uint32_t), tt::as_stream(stream)));
    return TT_OK;
}

int tt_env_reset(tt_env *env, const uint8_t *d_mask, float *d_obs, int64_t ld_obs, tt_stream_t stream) {
    TT_REQUIRE(env, "env is NULL");
    TT_REQUIRE(!d_obs || ld_obs >= TT_OBS_DIM, "ld_obs < 23");
    cudaStream_t s = tt::as_stream(stream);
    // a masked reset shares the iteration of the step that finished the episode; a full reset uses the
    // salted counter so it never collides with it, then advances the iteration.
    env_reset_kernel<<<(unsigned)grid_for(env->p.N), kBlock, 0, s>>>(env->p, env->k, d_mask, d_obs, ld_obs, env->seed,
                                                                    env->gid0, d_mask ? 0u : 0x80000000u);
    TT_COUNT_LAUNCH(); TT_LAUNCH_CHECK();
    if (!d_mask) { tick_kernel<<<1, 1, 0, s>>>(env->p.iter, 1u); TT_COUNT_LAUNCH(); TT_LAUNCH_CHECK(); }
    return TT_OK;
}

int tt_env_step_k(tt_env *env, const float *d_actions, int32_t K, int32_t auto_reset, float *d_obs, int64_t ld_obs,
                  float *d_reward, uint8_t *d_done, const tt_step_info *info, tt_stream_t stream) {
    TT_REQUIRE(env && d_actions, "NULL argument");
    TT_REQUIRE(K >= 1, "K < 1");
    TT_REQUIRE(!d_obs || ld_obs >= TT_OBS_DIM, "ld_obs < 23");
    cudaStream_t s = tt::as_stream(stream);
    const unsigned grid = (unsigned)grid_for(env->p.N);
    tt_step_info inf;
    memset(&inf, 0, sizeof inf);
    const bool want = info && (info->d_comps || info->d_violation || info->d_flags || info->d_success);
    // occupancy knob (registers per thread): TT_ENV_MINBLOCKS = 4 (128 regs) | 5 (96) | 6 (80) | 8 (64, spills)
    static const int variant = [] { const char *e = getenv("TT_ENV_MINBLOCKS"); return e ? atoi(e) : TT_ENV_MINBLOCKS_DEFAULT; }();
#define TT_LAUNCH_STEP(INFO, MB)                                                                                          \
    env_step_kernel<INFO, MB><<<grid, kBlock, 0, s>>>(env->p, env->k, d_actions, K, auto_reset, d_obs, ld_obs, d_reward, \
                                                      d_done, inf, env->seed, env->gid0)
    if (want) {
        inf = *info;
        TT_LAUNCH_STEP(true, 4);
    } else if (variant == 5) TT_LAUNCH_STEP(false, 5);
    else if (variant == 6) TT_LAUNCH_STEP(false, 6);
    else if (variant == 8) TT_LAUNCH_STEP(false, 8);
    else TT_LAUNCH_STEP(false, 4);
#undef TT_LAUNCH_STEP
    TT_COUNT_LAUNCH(); TT_LAUNCH_CHECK();
    if (auto_reset) { tick_kernel<<<1, 1, 0, s>>>(env->p.iter, (uint32_t)K); TT_COUNT_LAUNCH(); TT_LAUNCH_CHECK(); }
    return TT_OK;
}

int tt_env_step(tt_env *env, const float *d_action, float *d_obs, int64_t ld_obs, float *d_reward, uint8_t *d_done,
                const tt_step_info *info, tt_stream_t stream) {
    return tt_env_step_k(env, d_action, 1, 0, d_obs, ld_obs, d_reward, d_done, info, stream);
}

int tt_env_tick(tt_env *env, uint32_t by, tt_stream_t stream) {
    TT_REQUIRE(env, "env is NULL");
    tick_kernel<<<1, 1, 0, tt::as_stream(stream)>>>(env->p.iter, by);
    TT_COUNT_LAUNCH(); TT_LAUNCH_CHECK();
    return TT_OK;
}

int tt_env_set_state(tt_env *env, const int64_t *d_idx, int64_t n, const double *d_state, const double *d_start,
                     const double *d_goal, float *d_obs, int64_t ld_obs, tt_stream_t stream) {
    TT_REQUIRE(env && d_state && d_start && d_goal, "NULL argument");
    TT_REQUIRE(n > 0 && n <= env->p.N, "n out of range");
    TT_REQUIRE(!d_obs || ld_obs >= TT_OBS_DIM, "ld_obs < 23");
    env_set_state_kernel<<<(unsigned)grid_for(n), kBlock, 0, tt::as_stream(stream)>>>(env->p, env->k, d_idx, n, d_state,
                                                                                     d_start, d_goal, d_obs, ld_obs);
    TT_COUNT_LAUNCH(); TT_LAUNCH_CHECK();
    return TT_OK;
}

int tt_env_get_state(tt_env *env, double *d_state, double *d_start, double *d_goal, int32_t *d_steps,
                     int32_t *d_max_steps, tt_stream_t stream) {
    TT_REQUIRE(env, "env is NULL");
    env_get_state_kernel<<<(unsigned)grid_for(env->p.N), kBlock, 0, tt::as_stream(stream)>>>(env->p, d_state, d_start,
                                                                                            d_goal, d_steps, d_max_steps);
    TT_COUNT_LAUNCH(); TT_LAUNCH_CHECK();
    return TT_OK;
}

int tt_env_stats_read(tt_env *env, double *d_out16, int32_t clear, tt_stream_t stream) {
    TT_REQUIRE(env && d_out16, "NULL argument");
    stats_read_kernel<<<1, 32, 0, tt::as_stream(stream)>>>(env->p.stats, d_out16, clear);
    TT_COUNT_LAUNCH(); TT_LAUNCH_CHECK();
    return TT_OK;
}

const uint32_t *tt_env_iter_ptr(tt_env *env) { return env ? env->p.iter : nullptr; }
uint64_t tt_env_seed_value(tt_env *env) { return env ? env->seed : 0; }
uint64_t tt_env_global_offset(tt_env *env) { return env ? env->gid0 : 0; }
int64_t tt_env_num_envs(tt_env *env) { return env ? env->p.N : 0; }

}  // extern "C"
