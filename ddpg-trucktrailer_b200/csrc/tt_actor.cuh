// tt_actor.cuh -- packed device weights of the actor (ActorNetwork, DDPG/networks.py:98-147) shared by the
// fp32 CUDA-core kernel (tt_agent.cu) and the tcgen05 kernel (tt_actor_tc4.cu).
//
// Layouts inside the caller-provided workspace (all 256 B aligned):
//   w1t  f32 [k1p][h1p]   fc1.weight transposed (k-major), zero padded; k1p = in_dim rounded up to 8
//   w2t  f32 [h1p][h2p]   fc2.weight transposed (k-major), zero padded; h*p = h* rounded up to 32
//   b1 g1 be1 [h1p], b2 g2 be2 w3 [h2p], b3 [1]            (padded entries are 0)
//   w1c_{f16,bf16} UMMA image [hi | lo][n1 + 32 rows x 64 B]  centred, LayerNorm-scaled fc1 rows + statistic rows, K-major SWIZZLE_64B
//   w2s_{f16,bf16} UMMA image [sweep][kb2][rows x 64 B]       centred fc2.weight|fc2.bias in k-blocks of 32, same swizzle
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stddef.h>
#include <stdint.h>
#include "tt_common.cuh"
#include "tt_env_math.cuh"

// The v4 kernel streams W2 from L2 in every tile on every SM; TT_W2_REPLICAS identical copies of the image (SM i reads
// copy i % TT_W2_REPLICAS) spread that traffic over more L2 lines / slices.
#ifndef TT_W2_REPLICAS
#define TT_W2_REPLICAS 1
#endif

struct tt_actor_dev {
    int in_dim, h1, h2;
    int k1p, h1p, h2p, kb1;
    float *w1t, *w2t, *b1, *g1, *be1, *b2, *g2, *be2, *w3, *b3;
    void *w2s_f16, *w2s_bf16;                       // v4 layer-2 images: k-blocks grouped by output-column sweep
    void *w1c_f16, *w1c_bf16;                       // v4 layer-1 images: centred, LayerNorm-scaled rows + statistic rows (tt_actor_tc4.cu)
    double *l1c_scratch;                            // [12288]: v4 pack: [1200, 2032) column means / linear column of layer 2 (2 x 416), [2048, 11648) Gram partial sums of layer 1 (16 blocks x 600)
};

struct tt_actor {
    tt_actor_dev dev;
    bool loaded;
};

static inline size_t tt_actor_layout(int in_dim, int h1, int h2, tt_actor_dev *d, char *base) {
    const int k1p = (in_dim + 7) / 8 * 8, h1p = (h1 + 31) / 32 * 32, h2p = (h2 + 31) / 32 * 32, kb1 = 64;
    size_t off = 0;
    auto take = [&](size_t bytes) { size_t o = off; off = (off + bytes + 255) / 256 * 256; return o; };
    const size_t o_w1t = take(sizeof(float) * k1p * h1p), o_w2t = take(sizeof(float) * h1p * h2p);
    const size_t o_b1 = take(sizeof(float) * h1p), o_g1 = take(sizeof(float) * h1p), o_be1 = take(sizeof(float) * h1p);
    const size_t o_b2 = take(sizeof(float) * h2p), o_g2 = take(sizeof(float) * h2p), o_be2 = take(sizeof(float) * h2p);
    const size_t o_w3 = take(sizeof(float) * h2p), o_b3 = take(sizeof(float));
    const size_t n1 = (h1 + 15) / 16 * 16, n2 = (h2 + 15) / 16 * 16, kb2 = (h1 + 1 + 31) / 32;
    const size_t o_w2sh = take(TT_W2_REPLICAS * kb2 * n2 * 64), o_w2sb = take(TT_W2_REPLICAS * kb2 * n2 * 64);
    const size_t o_w1ch = take(2 * (n1 + 32) * 64), o_w1cb = take(2 * (n1 + 32) * 64);
    const size_t o_l1s = take(sizeof(double) * 12288);
    if (d) {
        d->in_dim = in_dim; d->h1 = h1; d->h2 = h2; d->k1p = k1p; d->h1p = h1p; d->h2p = h2p; d->kb1 = kb1;
        auto f = [&](size_t o) { return reinterpret_cast<float *>(base + o); };
        d->w1t = f(o_w1t); d->w2t = f(o_w2t); d->b1 = f(o_b1); d->g1 = f(o_g1); d->be1 = f(o_be1);
        d->b2 = f(o_b2); d->g2 = f(o_g2); d->be2 = f(o_be2); d->w3 = f(o_w3); d->b3 = f(o_b3);
        d->w2s_f16 = base + o_w2sh; d->w2s_bf16 = base + o_w2sb;
        d->w1c_f16 = base + o_w1ch; d->w1c_bf16 = base + o_w1cb;
        d->l1c_scratch = reinterpret_cast<double *>(base + o_l1s);
    }
    return off;
}

// What Agent.choose_action / the training loop do with the actor output (DDPG_agent.py:41-43 `mu + noise`, trainv2.py:516
// `clip(a, -1, 1) * pi / 4`, DDPG_agent.py:51-52 remember), fused into the output stage of the actor kernels so that the
// rollout needs no separate noise kernel.  All members optional.
struct TTActorTail {
    float *ou_x;                 // [n] OU state (DDPG/noise.py:12-17), advanced in place; NULL = no noise (evaluate)
    float *scaled;               // [n] clip(a, -1, 1) * float32(pi / 4); NULL = not wanted
    TTRingA ring;                // raw-action rows of the replay ring (A == NULL: no store)
    ttm::PhiloxKeys keys;        // expanded Philox round keys of the env seed
    uint32_t gid0;               // global id of row 0
    const uint32_t *iter;        // device iteration counter of the Philox streams
};
static inline TTActorTail tt_no_tail() {
    TTActorTail t;
    t.ou_x = nullptr; t.scaled = nullptr; t.ring.A = nullptr; t.ring.m = tt_make_ring_map(1, 0, 0);
    t.keys = ttm::philox_expand_key(0); t.gid0 = 0u; t.iter = nullptr;
    return t;
}

namespace tt {
// `ring` (may be NULL): also store the observation rows as the `state` part of the replay transitions (fused store);
// `tail` (may be NULL): OU noise, scaling and the ring store of the action fused into the output stage
int actor_forward_fp32(const tt_actor *a, const float *d_obs, int64_t ld, int64_t n, float *d_mu, const TTRingS *ring, const TTActorTail *tail, cudaStream_t s);
int actor_forward_tc(const tt_actor *a, const float *d_obs, int64_t ld, int64_t n, float *d_mu, int precision, const TTRingS *ring, const TTActorTail *tail, cudaStream_t s);
int actor_forward_any(const tt_actor *a, const float *d_obs, int64_t ld, int64_t n, float *d_mu, int precision, const TTRingS *ring, const TTActorTail *tail, cudaStream_t s);
int actor_resolve_precision(const tt_actor *a, int precision, int64_t n);
int actor_pack_all(tt_actor *a, const float *fc1_w, const float *fc1_b, const float *g1, const float *be1, const float *fc2_w, const float *fc2_b,
                   const float *g2, const float *be2, const float *mu_w, const float *mu_b, cudaStream_t s);
int launch_noise(float *d_x, float *d_action, float *d_scaled, const uint8_t *d_reset_mask, int64_t n, uint64_t seed,
                 uint64_t gid0, const uint32_t *d_iter, int evaluate, const TTRingA *ring, cudaStream_t s);
// per-device launch facts (one entry per CUDA device; a process may drive several GPUs)
int device_index();
int sm_count();
}  // namespace tt
