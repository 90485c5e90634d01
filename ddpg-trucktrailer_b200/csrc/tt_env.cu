// tt_env.cu -- kernel (a): the fused truck-trailer environment step for N independent environments, plus
// reset / state injection / readback / statistics.  Replaces Truck_trailer_Env_2.reset/step
// (truck_trailer_sim/simv2.py:459-545) and RewardFunction (reward_functionv1.py) for the whole batch.
//
// Layout in HBM (struct-of-arrays inside the caller-provided workspace, every array 256 B aligned; one 16 B
// vector per env and array, so a warp touches 512 contiguous bytes per access):
//   psi[N]  double2  psi1 psi2 (float64: the unstable hitch dynamics need it, see tt_env_math.cuh)
//   pos[N]  int4     x1 y1 x2 y2 in 2^-25 m fixed point
//   rsA[N]  float4   closest cum first_steer episode_return
//   rsB[N]  float4   g1 g2 g3 d0                (g = per-step distance decrements)
//   packed[N] u32    steps | emax | rmax-emax | stage bits | finished
//   goal[N] int4     gx gy (fixed point) sin/cos(gyaw) bits -- only READ when per-env goals were injected
//   l2v[N]  double2  trailer length L2 and v1x / L2 -- only READ by the step kernel after a per-env injection
//   pose[N][4] f64   startx starty startyaw goalyaw (written on reset only, one 32 B sector; host-visible attributes)
//   stats[16] f64, iter u32
// One thread owns one environment.  The step kernel is persistent: a CTA walks over 128-env tiles and issues
// the loads of its NEXT tile before computing the current one (register double buffer), so the DRAM latency
// is hidden behind ~1.2 k instructions of arithmetic instead of relying on occupancy.  The 23-float
// observation rows ([N, ld_obs] row-major, the layout the actor and the replay ring consume) are transposed
// through shared memory so that the global stores are coalesced 16 B vectors.
#include <new>
#include <stdlib.h>
#include <string.h>
#include "tt_common.cuh"
#include "tt_consts.h"
#include "tt_env_math.cuh"

using namespace ttm;

struct EnvPtrs {
    double2 *psi;
    int4 *pos;
    float4 *rsA, *rsB;
    uint32_t *packed;
    int4 *goal;
    double2 *l2v;        // (L2, v1x / L2) per env (heatmap.py:89 varies the trailer length per trial)
    double *pose;        // [4][N]
    double *stats;       // [16]
    uint32_t *iter;      // [0] Philox iteration counter, [1] CTAs-finished counter of the rollout kernel's tick
    uint32_t *done_list; // [N + slack] scratch of the rollout kernel: overflow of a CTA's shared-memory list of finished envs
    uint32_t *done_bits; // optional (tt_env_set_done_bits): bit i % 32 of word i / 32 = done of env i, written by single-step launches
    int64_t N;
};

struct tt_env {
    tt_env_cfg cfg;
    StepConsts k;
    EnvPtrs p;
    uint64_t seed;
    uint64_t gid0;
    bool goal_injected;  // per-env goals were injected (tt_env_set_state): the step kernel reads goal[] (and l2v[]) until the
                         // next FULL reset puts every env back on the configured goal
    bool l2_injected;    // per-env trailer lengths were injected (tt_env_set_l2; like `env.L2 = ...` they persist across resets)
};

namespace {

constexpr int kBlock = 128;
constexpr int kDoneSmem = 512;                 // finished envs a CTA of the rollout kernel lists in shared memory (the rest: global)
constexpr int64_t kListSlack = 128 * 4096;     // global overflow segments are whole tiles per CTA: <= 4096 CTAs of rounding slack
#ifndef TT_ENV_MINBLOCKS_DEFAULT
#define TT_ENV_MINBLOCKS_DEFAULT 4
#endif

// raw per-env words as they sit in HBM (the prefetch buffer)
struct EnvRaw {
    double2 psi;
    int4 pos;
    float4 a, b;
    int4 goal;
    double2 l2v;
    uint32_t packed;
    float action;
};

template <bool kGoal>
__device__ __forceinline__ void load_raw(const EnvPtrs &p, int64_t i, const float *__restrict__ actions, EnvRaw &r) {
    r.psi = p.psi[i]; r.pos = p.pos[i]; r.a = p.rsA[i]; r.b = p.rsB[i]; r.packed = p.packed[i];
    if (kGoal) { r.goal = __ldg(&p.goal[i]); r.l2v = __ldg(&p.l2v[i]); }
    r.action = __ldcs(&actions[i]);
}

// --- shared-memory prefetch ring (cp.async): each thread stages ITS OWN env's words and later reads only those back,
// so completion is per thread (cp.async.wait_group) and needs no barrier.  Frees the ~22 registers of a register
// double buffer and allows a prefetch distance of kStages - 1 tiles.
struct __align__(16) RawSlab {
    double2 psi[kBlock];
    int4 pos[kBlock];
    float4 a[kBlock], b[kBlock];
    int4 goal[kBlock];
    double2 l2v[kBlock];
    uint32_t packed[kBlock];
    float action[kBlock];
};

__device__ __forceinline__ void cp_async_16(void *smem, const void *g) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((uint32_t)__cvta_generic_to_shared(smem)), "l"(g) : "memory");
}
__device__ __forceinline__ void cp_async_4(void *smem, const void *g) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"((uint32_t)__cvta_generic_to_shared(smem)), "l"(g) : "memory");
}

template <bool kGoal>
__device__ __forceinline__ void stage_raw(const EnvPtrs &p, int64_t i, const float *__restrict__ actions, RawSlab &sl, int t) {
    cp_async_16(&sl.psi[t], &p.psi[i]); cp_async_16(&sl.pos[t], &p.pos[i]);
    cp_async_16(&sl.a[t], &p.rsA[i]); cp_async_16(&sl.b[t], &p.rsB[i]);
    if (kGoal) { cp_async_16(&sl.goal[t], &p.goal[i]); cp_async_16(&sl.l2v[t], &p.l2v[i]); }
    cp_async_4(&sl.packed[t], &p.packed[i]); cp_async_4(&sl.action[t], &actions[i]);
}

__device__ __forceinline__ void read_slab(const RawSlab &sl, int t, EnvRaw &r, bool goal) {
    r.psi = sl.psi[t]; r.pos = sl.pos[t]; r.a = sl.a[t]; r.b = sl.b[t]; r.packed = sl.packed[t]; r.action = sl.action[t];
    if (goal) { r.goal = sl.goal[t]; r.l2v = sl.l2v[t]; }
}

template <bool kGoal>
__device__ __forceinline__ void unpack(const StepConsts &k, const EnvRaw &r, EnvRegs &e) {
    e.psi1 = r.psi.x; e.psi2 = r.psi.y;
    e.x1 = r.pos.x; e.y1 = r.pos.y; e.x2 = r.pos.z; e.y2 = r.pos.w;
    e.closest = r.a.x; e.cum = r.a.y; e.first_steer = r.a.z; e.ep_ret = r.a.w;
    e.g1 = r.b.x; e.g2 = r.b.y; e.g3 = r.b.z; e.d0 = r.b.w;
    e.packed = r.packed;
    if (kGoal) { e.gx = r.goal.x; e.gy = r.goal.y; e.sgy = __int_as_float(r.goal.z); e.cgy = __int_as_float(r.goal.w); e.L2 = r.l2v.x; e.vL2 = r.l2v.y; }
    else { e.gx = k.gx_fix; e.gy = k.gy_fix; e.sgy = k.sgy0; e.cgy = k.cgy0; use_default_l2(k, e); }
}

__device__ __forceinline__ void store_dyn(const EnvPtrs &p, int64_t i, const EnvRegs &e) {
    p.psi[i] = make_double2(e.psi1, e.psi2);
    p.pos[i] = make_int4(e.x1, e.y1, e.x2, e.y2);
    p.rsA[i] = make_float4(e.closest, e.cum, e.first_steer, e.ep_ret);
    p.rsB[i] = make_float4(e.g1, e.g2, e.g3, e.d0);
    p.packed[i] = e.packed;
}

__device__ __forceinline__ void store_episode_consts(const EnvPtrs &p, int64_t i, const EnvRegs &e, double sx, double sy,
                                                     double syaw, double gyaw) {
    p.goal[i] = make_int4(e.gx, e.gy, __float_as_int(e.sgy), __float_as_int(e.cgy));
    double *ps = p.pose + 4 * i;
    ps[0] = sx; ps[1] = sy; ps[2] = syaw; ps[3] = gyaw;
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

// reset(): mask == nullptr -> every env; else only where mask[i] != 0.  Also clears the OU state is NOT done
// here (that is tt_ou_step's reset mask, trainv2.py:492).
// kPerEnv = false: no per-env trailer length / goal has been injected -- the configured L2 is used without reading l2v[] (one
// dependent DRAM round trip less) and goal[] already holds the configured goal.
template <bool kPerEnv>
__device__ __forceinline__ void reset_env(const EnvPtrs &p, const StepConsts &k, int64_t i, float *__restrict__ obs, int64_t ld,
                                          uint64_t seed, uint64_t gid0, uint32_t t, float *__restrict__ ou_x) {
    if (ou_x) ou_x[i] = 0.0f;                                 // agent.noise.reset() for the new episode (trainv2.py:492)
    EnvRegs e;
    if (kPerEnv) { const double2 l = p.l2v[i]; e.L2 = l.x; e.vL2 = l.y; } else use_default_l2(k, e);
    double sx, sy, syaw;
    rng_pose(k, seed, (uint32_t)(gid0 + i), t, sx, sy, syaw);
    float o[TT_OBS_DIM];
    reset_from_pose(k, e, sx, sy, syaw, k.gx, k.gy, k.gyaw, obs ? o : nullptr);
    store_dyn(p, i, e);
    if (kPerEnv) store_episode_consts(p, i, e, sx, sy, syaw, k.gyaw);
    else reinterpret_cast<double4 *>(p.pose)[i] = make_double4(sx, sy, syaw, k.gyaw);
    if (obs) {
#pragma unroll
        for (int c = 0; c < TT_OBS_DIM; c++) obs[i * ld + c] = o[c];
    }
}

// K env steps per launch; state stays in registers across the K steps; persistent over 128-env tiles.
// kMinBlocks == 7 is a memory-traffic-only probe (no arithmetic) used for roofline analysis (developer build).
// kRoll: the rollout's form (K == 1, obs != NULL): the driver-side `if done: env.reset(); agent.noise.reset()`
// (trainv2.py:489-492) happens inside the kernel, and the last CTA to finish advances the Philox iteration counter: no
// separate reset / tick launches.  A finished env's TERMINAL observation goes to the ring's new_state row (what
// trainv2.py:525 stores) with the tile's bulk store, like every other row.  Its reset is DEFERRED to the end of the kernel:
// the main loop only appends the env to the CTA's list (~1.4 % of the envs per step, but 84 % of the 128-env tiles and a
// third of the warps would otherwise run a few hundred extra dependent instructions -- Philox pose, float64 sin / cos,
// observation -- on one or two lanes: +22 % kernel time, measured); after the tile loop the CTA resets its listed envs one per
// thread on full warps and overwrites their rows of `obs` with the reset observation (the bulk stores have completed).
// Measured alternatives (N = 2^22, 3.4 % finished envs per step, profiles/env_roll_bench.py; the step + ring store alone: 267 us):
// reset inside the tile loop 322 us, per-warp 32-env tiles + reset inside the loop 316, per-warp tiles + deferred reset 308,
// per-warp tiles + per-warp batches of 12 inside the loop 322, this version (128-env tiles + deferred reset) 294 us.
template <bool kInfo, bool kGoal, int kMinBlocks, int kStages, bool kRoll>
__global__ void __launch_bounds__(kBlock, kMinBlocks) env_step_kernel(EnvPtrs p, StepConsts k, const float *__restrict__ actions,
                                                                      int K, int auto_reset, float *__restrict__ obs, int64_t ld,
                                                                      float *__restrict__ reward, uint8_t *__restrict__ done,
                                                                      tt_step_info info, uint64_t seed, uint64_t gid0, TTRingOut rpl,
                                                                      float *__restrict__ ou_x) {
    // observation tiles: double buffered; full tiles leave through the bulk-copy engine (cp.async.bulk shared ->
    // global), which drains them while the CTA already computes its next tile
    __shared__ __align__(128) float tiles[2][kBlock * TT_OBS_DIM];
    __shared__ int s_ndone;                               // kRoll: finished envs of this CTA so far ...
    __shared__ uint32_t s_done[kRoll ? kDoneSmem : 1];    // ... the first kDoneSmem of them (typically all: ~1.4 % of ~7 000 envs)
    uint32_t tbuf = 0;
    chain_enter();                                        // (tt_common.cuh: chained launches) before the first global access
    if (kRoll) { if (threadIdx.x == 0) s_ndone = 0; __syncthreads(); }
    const bool bulk_ok = obs != nullptr && ld == TT_OBS_DIM && ((reinterpret_cast<uintptr_t>(obs) & 15) == 0);
    const int64_t N = p.N;
    const int64_t ntiles = (N + kBlock - 1) / kBlock;
    const int64_t seg0 = (int64_t)blockIdx.x * ((ntiles + gridDim.x - 1) / gridDim.x) * kBlock;   // kRoll: this CTA's list segment
    const uint32_t t0 = *p.iter;
    StatAcc sa;
    sa.steps = sa.episodes = sa.successes = sa.ret = sa.ret2 = sa.rew = 0.f;
#pragma unroll
    for (int f = 0; f < 6; f++) sa.fl[f] = 0.f;

    extern __shared__ __align__(16) uint8_t dyn_smem[];
    RawSlab *slabs = reinterpret_cast<RawSlab *>(dyn_smem);       // kStages slabs (kStages == 0: unused)
    EnvRaw nxt;
    const int64_t stride = (int64_t)gridDim.x * kBlock;
    if (kStages == 0) {
        const int64_t i0 = (int64_t)blockIdx.x * kBlock + threadIdx.x;
        if (i0 < N) load_raw<kGoal>(p, i0, actions, nxt);
    } else {
#pragma unroll
        for (int st = 0; st < (kStages > 0 ? kStages - 1 : 0); st++) {    // prologue: kStages - 1 tiles in flight
            const int64_t ip = (int64_t)blockIdx.x * kBlock + threadIdx.x + st * stride;
            if (ip < N) stage_raw<kGoal>(p, ip, actions, slabs[st], threadIdx.x);
            asm volatile("cp.async.commit_group;" ::: "memory");
        }
    }
    int ring = 0;
    for (int64_t tileidx = blockIdx.x; tileidx < ntiles; tileidx += gridDim.x) {
        const int64_t row0 = tileidx * kBlock;
        const int64_t i = row0 + threadIdx.x;
        const bool active = i < N;
        const int rows = (int)((N - row0) < kBlock ? (N - row0) : kBlock);
        EnvRaw cur;
        if (kStages == 0) {
            cur = nxt;
            // prefetch the next tile of this CTA: in flight during the whole computation below
            const int64_t in = i + stride;
            if (in < N) load_raw<kGoal>(p, in, actions, nxt);
        } else {
            // issue the tile kStages - 1 ahead into the slab that was consumed in the previous iteration
            const int64_t ip = i + (int64_t)(kStages - 1) * stride;
            int slot = ring + kStages - 1; if (slot >= kStages) slot -= kStages;
            if (ip < N) stage_raw<kGoal>(p, ip, actions, slabs[slot], threadIdx.x);
            asm volatile("cp.async.commit_group;" ::: "memory");
            asm volatile("cp.async.wait_group %0;" ::"n"(kStages > 0 ? kStages - 1 : 0) : "memory");   // this tile's group landed
            if (active) read_slab(slabs[ring], threadIdx.x, cur, kGoal);
            if (++ring == kStages) ring = 0;
        }
        EnvRegs e;
        if (active) unpack<kGoal>(k, cur, e);

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
                    const float a = j == 0 ? cur.action : __ldcs(&actions[(int64_t)j * N + i]);
#if defined(TT_DEV_VARIANTS)
                    if (kMinBlocks == 7) {
                        e.psi1 += a; e.psi2 += a; e.x1 += 1; e.y1 += 2; e.x2 += 3; e.y2 += e.gx;
                        e.closest += a; e.cum += e.sgy; e.first_steer += e.cgy; e.g1 += a; e.g2 += a; e.g3 += a; e.packed += 1;
#pragma unroll
                        for (int c = 0; c < TT_OBS_DIM; c++) o.obs[c] = (float)e.x1 + (float)c;
                        o.reward = e.closest; o.done = false; o.success = false; o.flags = 0; o.viol = 0;
                    } else
#endif
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
                if (rpl.S2 && i >= rpl.m.first) {                       // fused replay store of (r, done): replay_buffer.py:13-21
                    const int64_t rr = rpl.m.row(i);
                    __stcs(&rpl.R[rr], o.reward); rpl.D[rr] = o.done ? 1 : 0;
                }
                if (kInfo) {
                    if (info.d_comps) {
#pragma unroll
                        for (int c = 0; c < TT_NCOMP; c++) info.d_comps[((int64_t)j * TT_NCOMP + c) * N + i] = o.comps[c];
                    }
                    if (info.d_violation) info.d_violation[oi] = (uint8_t)o.viol;
                    if (info.d_flags) info.d_flags[oi] = (uint8_t)o.flags;
                    if (info.d_success) info.d_success[oi] = o.success ? 1 : 0;
                }
                if (!kRoll && o.done && !(e.packed & PK_FINISHED)) {
                    if (auto_reset) {
                        double sx, sy, syaw;
                        rng_pose(k, seed, (uint32_t)(gid0 + i), t0 + (uint32_t)j, sx, sy, syaw);
                        reset_from_pose(k, e, sx, sy, syaw, k.gx, k.gy, k.gyaw, nullptr);
                        store_episode_consts(p, i, e, sx, sy, syaw, k.gyaw);
                    } else e.packed |= PK_FINISHED;
                }
            }
            if (p.done_bits && K == 1) {                                 // warp-uniform: 32 consecutive envs = one word (a tile starts at a multiple of 128)
                const uint32_t bits = __ballot_sync(0xffffffffu, active && o.done);
                if ((threadIdx.x & 31) == 0 && i < N) p.done_bits[i >> 5] = bits;
            }
            if (kRoll && active && o.done) {                             // reset at the end of the kernel
                const int slot = atomicAdd(&s_ndone, 1);
                if (slot < kDoneSmem) s_done[slot] = (uint32_t)i; else p.done_list[seg0 + slot] = (uint32_t)i;
            }
            if (obs) {
                float *tile = tiles[tbuf];
                __syncthreads();                 // thread 0 has waited for the bulk store that last read this buffer
                if (active) {
#pragma unroll
                    for (int c = 0; c < TT_OBS_DIM; c++) tile[threadIdx.x * TT_OBS_DIM + c] = o.obs[c];
                }
                float *dst = obs + (int64_t)j * N * ld;
                // fused replay store of s': the tile goes to the ring's new_state rows as a second bulk copy when those
                // rows are one aligned contiguous run, else element-wise from the same shared-memory tile
                const int64_t rrow0 = rpl.S2 ? rpl.m.row(row0) : 0;
                const bool ring_bulk = rpl.S2 && rows == kBlock && !rpl.m.many && row0 >= rpl.m.first && rrow0 + kBlock <= rpl.m.cap &&
                                       (rrow0 & 3) == 0 && ((reinterpret_cast<uintptr_t>(rpl.S2) & 15) == 0);
                if (bulk_ok && rows == kBlock) {
                    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                    __syncthreads();
                    if (threadIdx.x == 0) {
                        const uint32_t src = (uint32_t)__cvta_generic_to_shared(tile);
                        asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;"
                                     ::"l"(dst + row0 * TT_OBS_DIM), "r"(src), "r"((uint32_t)(kBlock * TT_OBS_DIM * sizeof(float))) : "memory");
                        if (ring_bulk)
                            asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;"
                                         ::"l"(rpl.S2 + rrow0 * TT_OBS_DIM), "r"(src), "r"((uint32_t)(kBlock * TT_OBS_DIM * sizeof(float))) : "memory");
                        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
                        asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");   // the OTHER buffer is free again
                    }
                } else {
                    __syncthreads();
                    store_obs_tile(tile, dst, ld, row0, rows);
                }
                if (rpl.S2 && !(ring_bulk && bulk_ok)) {                 // (the tile stays valid: it is overwritten two tiles later)
                    for (int v = threadIdx.x; v < rows * TT_OBS_DIM; v += kBlock) {
                        const int r = v / TT_OBS_DIM, c = v - r * TT_OBS_DIM;
                        if (row0 + r >= rpl.m.first) rpl.S2[rpl.m.row(row0 + r) * TT_OBS_DIM + c] = tile[v];
                    }
                }
                tbuf ^= 1u;
            }
        }
        if (active) store_dyn(p, i, e);
    }

    if ((threadIdx.x & 31) == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");      // all bulk stores (of this issuing lane) complete
    if (kRoll) {
        // deferred resets (trainv2.py:489-492: env.reset() + agent.noise.reset()): one listed env per thread, full warps.  Every
        // bulk store of this CTA has completed (wait_group 0 above + the barrier), so the reset observation written here
        // replaces the terminal row in `obs`; the ring keeps the terminal row.
        __syncthreads();
        const int nd = s_ndone;
        for (int j = threadIdx.x; j < nd; j += kBlock)
            reset_env<kGoal>(p, k, (int64_t)(j < kDoneSmem ? s_done[j] : p.done_list[seg0 + j]), obs, ld, seed, gid0, t0, ou_x);
        // iteration tick: every CTA read *p.iter (t0) when it started, and the last one to get here has seen all others finish
        __syncthreads();
        if (threadIdx.x == 0 && atomicAdd(p.iter + 1, 1u) == gridDim.x - 1) { p.iter[1] = 0u; *p.iter = t0 + 1u; }
    }

    // statistics: steps and reward every launch, the episode counters only in warps that finished an episode
    // (one warp-shuffle tree each; one double atomic per warp and statistic).  All lanes of the warp get here.
    const int lane = threadIdx.x & 31;
    {
        const float rs = warp_sum(sa.rew), st = warp_sum(sa.steps);
        if (lane == 0 && st != 0.f) { atomicAdd(&p.stats[0], (double)st); atomicAdd(&p.stats[5], (double)rs); }
        if (__any_sync(0xffffffffu, sa.episodes != 0.f)) {
            float v[10] = {sa.episodes, sa.successes, sa.ret, sa.ret2, sa.fl[0], sa.fl[1], sa.fl[2], sa.fl[3], sa.fl[4], sa.fl[5]};
            const int slot[10] = {1, 2, 3, 4, 6, 7, 8, 9, 10, 11};
#pragma unroll
            for (int s = 0; s < 10; s++) {
                const float r = warp_sum(v[s]);
                if (lane == 0 && r != 0.f) atomicAdd(&p.stats[slot[s]], (double)r);
            }
        }
    }
}

__global__ void __launch_bounds__(kBlock) env_reset_kernel(EnvPtrs p, StepConsts k, const uint8_t *__restrict__ mask,
                                                           float *__restrict__ obs, int64_t ld, uint64_t seed,
                                                           uint64_t gid0, uint32_t t_salt, float *__restrict__ ou_x) {
    const int64_t i = (int64_t)blockIdx.x * kBlock + threadIdx.x;
    if (i >= p.N) return;
    if (mask && !mask[i]) return;
    reset_env<true>(p, k, i, obs, ld, seed, gid0, *p.iter + t_salt, ou_x);
}

// Masked reset for the rollout, where about 1 % of the envs finish per step.  One thread per env spends 20 us at N = 2^22
// on scheduling 32 768 CTAs whose threads read one mask byte each and return.  Here a small persistent grid scans the mask
// 16 envs per 16 B load; every warp compacts the finished envs of its 512-env window into a list (ballot + prefix count in
// shared memory) and then resets them one env per LANE, so the reset arithmetic runs on full-width warps instead of on the
// few scattered lanes that happened to own a finished env.  Needs a 16 B aligned mask.
constexpr int kWarpWindow = 32 * 16;
__global__ void __launch_bounds__(kBlock) env_reset_sparse_kernel(EnvPtrs p, StepConsts k, const uint8_t *__restrict__ mask,
                                                                  float *__restrict__ obs, int64_t ld, uint64_t seed,
                                                                  uint64_t gid0, float *__restrict__ ou_x) {
    __shared__ int32_t list[kBlock / 32][kWarpWindow];
    const int64_t N = p.N;
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    const int64_t nwin = (N + kWarpWindow - 1) / kWarpWindow;
    const int64_t warp0 = (int64_t)blockIdx.x * (kBlock / 32) + wib, nwarps = (int64_t)gridDim.x * (kBlock / 32);
    const uint32_t t = *p.iter;
    for (int64_t w = warp0; w < nwin; w += nwarps) {
        const int64_t i0 = w * kWarpWindow + lane * 16;
        uint32_t m[4] = {0u, 0u, 0u, 0u};
        if (i0 + 16 <= N) {
            const uint4 v = __ldcs(reinterpret_cast<const uint4 *>(mask + i0));
            m[0] = v.x; m[1] = v.y; m[2] = v.z; m[3] = v.w;
        } else {
            for (int b = 0; i0 + b < N; b++) m[b >> 2] |= (uint32_t)mask[i0 + b] << (8 * (b & 3));
        }
        uint32_t bits = 0;                                               // one bit per env of this lane's 16
#pragma unroll
        for (int b = 0; b < 16; b++) bits |= ((m[b >> 2] >> (8 * (b & 3))) & 0xFFu) ? (1u << b) : 0u;
        const int cnt = __popc(bits);
        int pre = cnt;                                                   // inclusive prefix sum over the warp
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const int y = __shfl_up_sync(0xffffffffu, pre, o); if (lane >= o) pre += y; }
        const int total = __shfl_sync(0xffffffffu, pre, 31);
        if (total == 0) continue;
        int at = pre - cnt;
        while (bits) { const int b = __ffs(bits) - 1; bits &= bits - 1; list[wib][at++] = lane * 16 + b; }
        __syncwarp();
        for (int j = lane; j < total; j += 32) reset_env<true>(p, k, w * kWarpWindow + list[wib][j], obs, ld, seed, gid0, t, ou_x);
        __syncwarp();
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
    { const double2 l = p.l2v[i]; e.L2 = l.x; e.vL2 = l.y; }
    e.psi1 = state[6 * j]; e.psi2 = state[6 * j + 1];
    e.x1 = pos_from_double(state[6 * j + 2]); e.y1 = pos_from_double(state[6 * j + 3]);
    e.x2 = pos_from_double(state[6 * j + 4]); e.y2 = pos_from_double(state[6 * j + 5]);
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
        const double2 a = p.psi[i];
        const int4 q = p.pos[i];
        state[6 * i] = a.x; state[6 * i + 1] = a.y;
        state[6 * i + 2] = pos_to_double(q.x); state[6 * i + 3] = pos_to_double(q.y);
        state[6 * i + 4] = pos_to_double(q.z); state[6 * i + 5] = pos_to_double(q.w);
    }
    if (start) { start[3 * i] = p.pose[4 * i]; start[3 * i + 1] = p.pose[4 * i + 1]; start[3 * i + 2] = p.pose[4 * i + 2]; }
    if (goal) { const int4 g = p.goal[i]; goal[3 * i] = pos_to_double(g.x); goal[3 * i + 1] = pos_to_double(g.y); goal[3 * i + 2] = p.pose[4 * i + 3]; }
    const uint32_t pk = p.packed[i];
    if (steps) steps[i] = (int32_t)(pk & PK_STEPS_MASK);
    if (max_steps) max_steps[i] = (int32_t)((pk >> PK_EMAX_SHIFT) & PK_EMAX_MASK);
}

// persistent reward state (reward_functionv1.py:99-109) as float32 arrays; any pointer may be NULL
__global__ void __launch_bounds__(kBlock) env_get_reward_state_kernel(EnvPtrs p, float *__restrict__ closest, float *__restrict__ cum,
                                                                      float *__restrict__ first_steer, float *__restrict__ ep_ret) {
    const int64_t i = (int64_t)blockIdx.x * kBlock + threadIdx.x;
    if (i >= p.N) return;
    const float4 a = p.rsA[i];
    if (closest) closest[i] = a.x;
    if (cum) cum[i] = a.y;
    if (first_steer) first_steer[i] = a.z;
    if (ep_ret) ep_ret[i] = a.w;
}

__global__ void tick_kernel(uint32_t *iter, uint32_t by) { *iter += by; }

__global__ void __launch_bounds__(kBlock) env_init_goal_kernel(EnvPtrs p, StepConsts k) {
    const int64_t i = (int64_t)blockIdx.x * kBlock + threadIdx.x;
    if (i < p.N) {
        p.goal[i] = make_int4(k.gx_fix, k.gy_fix, __float_as_int(k.sgy0), __float_as_int(k.cgy0));
        p.l2v[i] = make_double2(k.L2, k.vL2);
        p.pose[4 * i + 3] = k.gyaw;            // env.goalyaw is valid before the first reset (simv2.py:61-63)
    }
}

// env.L2 = value (heatmap.py:89): per-env trailer length.  Takes effect for the dynamics at once and for the truck
// position of the next reset / set_state.
__global__ void __launch_bounds__(kBlock) env_set_l2_kernel(EnvPtrs p, double v1x, const int64_t *__restrict__ idx, int64_t n,
                                                            const double *__restrict__ l2) {
    const int64_t j = (int64_t)blockIdx.x * kBlock + threadIdx.x;
    if (j >= n) return;
    const int64_t i = idx ? idx[j] : j;
    if (i < 0 || i >= p.N) return;
    p.l2v[i] = make_double2(l2[j], v1x / l2[j]);
}
__global__ void __launch_bounds__(kBlock) env_get_l2_kernel(EnvPtrs p, double *__restrict__ l2) {
    const int64_t i = (int64_t)blockIdx.x * kBlock + threadIdx.x;
    if (i < p.N) l2[i] = p.l2v[i].x;
}

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
    const size_t o_psi = take(sizeof(double2) * n), o_pos = take(sizeof(int4) * n), o_a = take(sizeof(float4) * n),
                 o_b = take(sizeof(float4) * n), o_pk = take(sizeof(uint32_t) * n), o_goal = take(sizeof(int4) * n), o_l2v = take(sizeof(double2) * n),
                 o_pose = take(sizeof(double) * 4 * n), o_stats = take(sizeof(double) * TT_NSTATS), o_iter = take(256),
                 o_list = take(sizeof(uint32_t) * (n + kListSlack));
    if (p) {
        p->psi = reinterpret_cast<double2 *>(base + o_psi); p->pos = reinterpret_cast<int4 *>(base + o_pos);
        p->rsA = reinterpret_cast<float4 *>(base + o_a); p->rsB = reinterpret_cast<float4 *>(base + o_b);
        p->packed = reinterpret_cast<uint32_t *>(base + o_pk); p->goal = reinterpret_cast<int4 *>(base + o_goal); p->l2v = reinterpret_cast<double2 *>(base + o_l2v);
        p->pose = reinterpret_cast<double *>(base + o_pose); p->stats = reinterpret_cast<double *>(base + o_stats);
        p->iter = reinterpret_cast<uint32_t *>(base + o_iter); p->done_list = reinterpret_cast<uint32_t *>(base + o_list); p->done_bits = nullptr; p->N = n;
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
    e->cfg = *cfg; e->k = tt_make_consts(*cfg); e->seed = seed; e->gid0 = global_env_offset; e->goal_injected = false; e->l2_injected = false;
    env_layout(n_envs, &e->p, static_cast<char *>(d_workspace));
    cudaError_t err = cudaMemset(d_workspace, 0, tt_env_workspace_bytes(n_envs));
    if (err != cudaSuccess) { delete e; return tt::cuda_fail(err, "cudaMemset(workspace)"); }
    env_init_goal_kernel<<<(unsigned)grid_for(n_envs), kBlock>>>(e->p, e->k);      // default goal everywhere
    TT_COUNT_LAUNCH();
    err = cudaGetLastError();
    if (err != cudaSuccess) { delete e; return tt::cuda_fail(err, "env_init_goal_kernel"); }
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

static int env_reset_impl(tt_env *env, const uint8_t *d_mask, float *d_obs, int64_t ld_obs, float *d_ou_x, tt_stream_t stream) {
    TT_REQUIRE(env, "env is NULL");
    TT_REQUIRE(!d_obs || ld_obs >= TT_OBS_DIM, "ld_obs < 23");
    cudaStream_t s = tt::as_stream(stream);
    // a masked reset shares the iteration of the step that finished the episode; a full reset uses the
    // salted counter so it never collides with it, then advances the iteration.
    if (d_mask && (reinterpret_cast<uintptr_t>(d_mask) & 15) == 0 && env->p.N >= (int64_t)1 << 16) {
        const int64_t nwin = (env->p.N + kWarpWindow - 1) / kWarpWindow, want = (nwin + kBlock / 32 - 1) / (kBlock / 32);
        const int64_t cap = (int64_t)tt::sm_count() * 16;
        env_reset_sparse_kernel<<<(unsigned)(want < cap ? want : cap), kBlock, 0, s>>>(env->p, env->k, d_mask, d_obs, ld_obs, env->seed,
                                                                                       env->gid0, d_ou_x);
    } else
        env_reset_kernel<<<(unsigned)grid_for(env->p.N), kBlock, 0, s>>>(env->p, env->k, d_mask, d_obs, ld_obs, env->seed,
                                                                        env->gid0, d_mask ? 0u : 0x80000000u, d_ou_x);
    TT_COUNT_LAUNCH(); TT_LAUNCH_CHECK();
    if (!d_mask) {
        tick_kernel<<<1, 1, 0, s>>>(env->p.iter, 1u); TT_COUNT_LAUNCH(); TT_LAUNCH_CHECK();
        env->goal_injected = false;            // every env is back on the configured goal: the step kernel stops reading goal[]
    }
    return TT_OK;
}

int tt_env_reset(tt_env *env, const uint8_t *d_mask, float *d_obs, int64_t ld_obs, tt_stream_t stream) {
    return env_reset_impl(env, d_mask, d_obs, ld_obs, nullptr, stream);
}

// internal (tt_rollout.cu): masked reset that also zeroes the OU state of the reset envs (one launch instead of two)
int tt_env_reset_ou(tt_env *env, const uint8_t *d_mask, float *d_obs, int64_t ld_obs, float *d_ou_x, tt_stream_t stream) {
    return env_reset_impl(env, d_mask, d_obs, ld_obs, d_ou_x, stream);
}

static int env_step_impl(tt_env *env, const float *d_actions, int32_t K, int32_t auto_reset, float *d_obs, int64_t ld_obs,
                         float *d_reward, uint8_t *d_done, const tt_step_info *info, const TTRingOut *ring, bool roll, float *d_ou_x,
                         tt_stream_t stream) {
    TT_REQUIRE(env && d_actions, "NULL argument");
    TT_REQUIRE(!ring || (K == 1 && d_obs), "fused replay store needs K == 1 and an observation buffer");
    TT_REQUIRE(!roll || (K == 1 && d_obs), "fused reset needs K == 1 and an observation buffer");
    TTRingOut ro;
    if (ring) ro = *ring; else { ro.S2 = nullptr; ro.R = nullptr; ro.D = nullptr; ro.m = tt_make_ring_map(1, 0, 0); }
    TT_REQUIRE(K >= 1, "K < 1");
    TT_REQUIRE(!d_obs || ld_obs >= TT_OBS_DIM, "ld_obs < 23");
    cudaStream_t s = tt::as_stream(stream);
    tt_step_info inf;
    memset(&inf, 0, sizeof inf);
    const bool want = info && (info->d_comps || info->d_violation || info->d_flags || info->d_success);
    const int64_t ntiles = grid_for(env->p.N);
    // 4 CTAs / SM (<= 128 registers) with a 2-stage cp.async shared-memory prefetch: measured best of the combinations tried
    // (profiles/env_kernel_bench.py; the others exist in the developer build only, TT_ENV_MINBLOCKS).  Opt-in shared memory
    // and occupancy are per device.
#define TT_LAUNCH_STEP(INFO, GOAL, MB, ST, ROLL)                                                                                \
    do {                                                                                                                        \
        auto kern = env_step_kernel<INFO, GOAL, MB, ST, ROLL>;                                                                  \
        const size_t dsm = (size_t)(ST) * sizeof(RawSlab);                                                                      \
        static int per_sm_of[tt::kMaxDevices] = {};                                                                            \
        int &per_sm = per_sm_of[tt::device_index()];                                                                            \
        if (per_sm == 0) {                                                                                                      \
            if (dsm > 0) TT_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dsm));            \
            if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, kBlock, dsm) != cudaSuccess || per_sm <= 0) per_sm = 4; \
        }                                                                                                                       \
        const int64_t cap = (int64_t)tt::grid_sms() * per_sm;                                                                   \
        TT_CUDA(tt::launch_chained(tt::chain_rollout(env->p.N), kern, dim3((unsigned)(ntiles < cap ? ntiles : cap)), dim3(kBlock), dsm, s, \
                                   env->p, env->k, d_actions, (int)K, (int)auto_reset, d_obs, ld_obs, d_reward, d_done, inf, env->seed,     \
                                   env->gid0, ro, d_ou_x));                                                                               \
    } while (0)
    const bool goal = env->goal_injected || env->l2_injected;
    if (roll) {
        if (goal) TT_LAUNCH_STEP(false, true, 4, 2, true); else TT_LAUNCH_STEP(false, false, 4, 2, true);
    } else if (want) {
        inf = *info;
        if (goal) TT_LAUNCH_STEP(true, true, 4, 2, false); else TT_LAUNCH_STEP(true, false, 4, 2, false);
    } else if (goal) TT_LAUNCH_STEP(false, true, 4, 2, false);
    else {
#if defined(TT_DEV_VARIANTS)
        // tuning knob TT_ENV_MINBLOCKS = <CTAs per SM><prefetch stages>; 40/5/6 = register double buffer; 7 = traffic-only probe.
        // Measured at N = 2^22: 42: 208 us (70.4 % of the HBM roofline), 43: 210, 52/53: 218, 40: 264, 62: 276.
        static const int variant = [] { const char *e = getenv("TT_ENV_MINBLOCKS"); return e ? atoi(e) : 42; }();
        if (variant == 5) TT_LAUNCH_STEP(false, false, 5, 0, false);
        else if (variant == 6) TT_LAUNCH_STEP(false, false, 6, 0, false);
        else if (variant == 7) TT_LAUNCH_STEP(false, false, 7, 0, false);
        else if (variant == 40) TT_LAUNCH_STEP(false, false, 4, 0, false);
        else if (variant == 43) TT_LAUNCH_STEP(false, false, 4, 3, false);
        else if (variant == 52) TT_LAUNCH_STEP(false, false, 5, 2, false);
        else if (variant == 53) TT_LAUNCH_STEP(false, false, 5, 3, false);
        else if (variant == 62) TT_LAUNCH_STEP(false, false, 6, 2, false);
        else
#endif
        TT_LAUNCH_STEP(false, false, 4, 2, false);
    }
#undef TT_LAUNCH_STEP
    TT_COUNT_LAUNCH(); TT_LAUNCH_CHECK();
    if (auto_reset && !roll) { tick_kernel<<<1, 1, 0, s>>>(env->p.iter, (uint32_t)K); TT_COUNT_LAUNCH(); TT_LAUNCH_CHECK(); }
    return TT_OK;
}

int tt_env_step_k(tt_env *env, const float *d_actions, int32_t K, int32_t auto_reset, float *d_obs, int64_t ld_obs,
                  float *d_reward, uint8_t *d_done, const tt_step_info *info, tt_stream_t stream) {
    return env_step_impl(env, d_actions, K, auto_reset, d_obs, ld_obs, d_reward, d_done, info, nullptr, false, nullptr, stream);
}

int tt_env_step(tt_env *env, const float *d_action, float *d_obs, int64_t ld_obs, float *d_reward, uint8_t *d_done,
                const tt_step_info *info, tt_stream_t stream) {
    return env_step_impl(env, d_action, 1, 0, d_obs, ld_obs, d_reward, d_done, info, nullptr, false, nullptr, stream);
}

int tt_env_step_store(tt_env *env, const float *d_action, float *d_obs, int64_t ld_obs, float *d_reward, uint8_t *d_done,
                      const tt_replay_ring *ring, tt_stream_t stream) {
    TT_REQUIRE(env && ring && ring->d_new_state_mem && ring->d_reward_mem && ring->d_terminal_mem && ring->mem_size > 0 &&
               ring->mem_cntr >= 0, "bad ring");
    const TTRingOut ro = {ring->d_new_state_mem, ring->d_reward_mem, ring->d_terminal_mem,
                          tt_make_ring_map(ring->mem_size, ring->mem_cntr, env->p.N)};
    return env_step_impl(env, d_action, 1, 0, d_obs, ld_obs, d_reward, d_done, nullptr, &ro, false, nullptr, stream);
}

// step + `if done: env.reset(); agent.noise.reset()` (trainv2.py:489-492) + iteration tick in ONE launch: the env half of a
// rollout iteration.  d_obs rows of finished envs receive the reset observation, the ring (optional) their terminal one.
int tt_env_step_reset(tt_env *env, const float *d_action, float *d_obs, int64_t ld_obs, float *d_reward, uint8_t *d_done,
                      float *d_ou_x, const tt_replay_ring *ring, tt_stream_t stream) {
    TT_REQUIRE(env && d_action && d_obs, "NULL argument");
    if (!ring) return env_step_impl(env, d_action, 1, 1, d_obs, ld_obs, d_reward, d_done, nullptr, nullptr, true, d_ou_x, stream);
    TT_REQUIRE(ring->d_new_state_mem && ring->d_reward_mem && ring->d_terminal_mem && ring->mem_size > 0 && ring->mem_cntr >= 0, "bad ring");
    const TTRingOut ro = {ring->d_new_state_mem, ring->d_reward_mem, ring->d_terminal_mem,
                          tt_make_ring_map(ring->mem_size, ring->mem_cntr, env->p.N)};
    return env_step_impl(env, d_action, 1, 1, d_obs, ld_obs, d_reward, d_done, nullptr, &ro, true, d_ou_x, stream);
}

int tt_env_set_done_bits(tt_env *env, uint32_t *d_bits) {
    TT_REQUIRE(env, "env is NULL");
    TT_REQUIRE((reinterpret_cast<uintptr_t>(d_bits) & 3) == 0, "d_bits must be 4 B aligned");
    env->p.done_bits = d_bits;
    return TT_OK;
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
    env->goal_injected = true;     // injected goals may differ from the configured one: the step kernel reads goal[]
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

int tt_env_set_l2(tt_env *env, const int64_t *d_idx, int64_t n, const double *d_l2, tt_stream_t stream) {
    TT_REQUIRE(env && d_l2, "NULL argument");
    TT_REQUIRE(n > 0 && n <= env->p.N, "bad n");
    env_set_l2_kernel<<<(unsigned)grid_for(n), kBlock, 0, tt::as_stream(stream)>>>(env->p, env->cfg.v1x, d_idx, n, d_l2);
    TT_COUNT_LAUNCH(); TT_LAUNCH_CHECK();
    env->l2_injected = true;       // the step kernel must read l2v[] (and goal[]) from now on
    return TT_OK;
}

int tt_env_get_l2(tt_env *env, double *d_l2, tt_stream_t stream) {
    TT_REQUIRE(env && d_l2, "NULL argument");
    env_get_l2_kernel<<<(unsigned)grid_for(env->p.N), kBlock, 0, tt::as_stream(stream)>>>(env->p, d_l2);
    TT_COUNT_LAUNCH(); TT_LAUNCH_CHECK();
    return TT_OK;
}

int tt_env_get_reward_state(tt_env *env, float *d_closest, float *d_cum_backward, float *d_first_steer, float *d_episode_return,
                            tt_stream_t stream) {
    TT_REQUIRE(env, "env is NULL");
    env_get_reward_state_kernel<<<(unsigned)grid_for(env->p.N), kBlock, 0, tt::as_stream(stream)>>>(env->p, d_closest, d_cum_backward,
                                                                                                   d_first_steer, d_episode_return);
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
