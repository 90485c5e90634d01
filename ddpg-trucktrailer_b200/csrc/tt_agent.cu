// tt_agent.cu -- kernels (b) OU noise and (c) the batched actor forward (fp32 CUDA-core path), plus the
// weight re-packing.  Replaces Agent.choose_action (DDPG/DDPG_agent.py:36-49), OUActionNoise
// (DDPG/noise.py:12-20) and ActorNetwork.forward (DDPG/networks.py:138-147) for N observations at once.
// The tcgen05 tensor-core actor lives in tt_actor_tc4.cu and shares tt_actor with this file.
#include <new>
#include "tt_actor.cuh"
#include "tt_common.cuh"
#include "tt_env_math.cuh"

namespace {

constexpr int kThreads = 256;
constexpr float kPiOver4F = 0.78539819f;       // float32(pi/4) == env.action_space.high[0] (simv2.py:86-91)

// ---------------------------------------------------------------------------------------------------------
// (b) OU noise, DDPG/noise.py:12-17 with theta=0.2, sigma=0.15, dt=1e-2, mu=0; optional fused
//     "mu + noise" (DDPG_agent.py:41-43) and "clip(a,-1,1) * pi/4" (trainv2.py:516)
// ---------------------------------------------------------------------------------------------------------
using ttm::ou_advance;

__global__ void __launch_bounds__(kThreads) ou_kernel(float *__restrict__ x, float *__restrict__ action,
                                                      float *__restrict__ scaled, const uint8_t *__restrict__ reset_mask,
                                                      int64_t n, ttm::PhiloxKeys keys, uint64_t gid0,
                                                      const uint32_t *__restrict__ iter, int evaluate, TTRingA ring) {
    const int64_t i = (int64_t)blockIdx.x * kThreads + threadIdx.x;
    if (i >= n) return;
    float a = action ? action[i] : 0.0f;
    if (!evaluate) {
        float xp = x[i];
        if (reset_mask && reset_mask[i]) xp = 0.0f;                      // trainv2.py:492 agent.noise.reset()
        const float xn = ou_advance(xp, ttm::rng_normal_ks(keys, (uint32_t)(gid0 + i), *iter));
        x[i] = xn;
        a += xn;
        if (action) action[i] = a;
    }
    if (scaled) scaled[i] = fminf(fmaxf(a, -1.0f), 1.0f) * kPiOver4F;
    if (ring.A && i >= ring.m.first) ring.A[ring.m.row(i)] = a;          // agent.remember stores the UNCLIPPED action (trainv2.py:525)
}

// The rollout's form of the same kernel: four consecutive envs per thread, 16 B loads / stores (x, action, scaled and the
// ring's action rows), no reset mask (tt_rollout_step zeroes the OU state in the reset kernel).  Needs n % 4 == 0, 16 B
// aligned arrays and a ring position that keeps every group of four rows contiguous and aligned.
__global__ void __launch_bounds__(kThreads) ou_kernel_x4(float4 *__restrict__ x, float4 *__restrict__ action, float4 *__restrict__ scaled,
                                                         int64_t n4, ttm::PhiloxKeys keys, uint32_t gid0, const uint32_t *__restrict__ iter,
                                                         int evaluate, float4 *__restrict__ ring_a, int64_t ring_base4, int64_t ring_cap4) {
    const uint32_t t = *iter;
    for (int64_t v = (int64_t)blockIdx.x * kThreads + threadIdx.x; v < n4; v += (int64_t)gridDim.x * kThreads) {
        float4 a = __ldcs(&action[v]);
        if (!evaluate) {
            float4 xp = x[v];
            const uint32_t g = gid0 + (uint32_t)(4 * v);
            xp.x = ou_advance(xp.x, ttm::rng_normal_ks(keys, g, t));     xp.y = ou_advance(xp.y, ttm::rng_normal_ks(keys, g + 1u, t));
            xp.z = ou_advance(xp.z, ttm::rng_normal_ks(keys, g + 2u, t)); xp.w = ou_advance(xp.w, ttm::rng_normal_ks(keys, g + 3u, t));
            x[v] = xp;
            a.x += xp.x; a.y += xp.y; a.z += xp.z; a.w += xp.w;
            action[v] = a;
        }
        float4 sc;
        sc.x = fminf(fmaxf(a.x, -1.0f), 1.0f) * kPiOver4F; sc.y = fminf(fmaxf(a.y, -1.0f), 1.0f) * kPiOver4F;
        sc.z = fminf(fmaxf(a.z, -1.0f), 1.0f) * kPiOver4F; sc.w = fminf(fmaxf(a.w, -1.0f), 1.0f) * kPiOver4F;
        __stcs(&scaled[v], sc);
        if (ring_a) {                                                    // rows (base + 4 v .. + 3) % cap, wrapping between groups only
            int64_t r4 = ring_base4 + v;
            if (r4 >= ring_cap4) r4 -= ring_cap4;
            __stcs(&ring_a[r4], a);
        }
    }
}

__global__ void __launch_bounds__(kThreads) scale_kernel(const float *__restrict__ a, float *__restrict__ s, int64_t n) {
    const int64_t i = (int64_t)blockIdx.x * kThreads + threadIdx.x;
    if (i < n) s[i] = fminf(fmaxf(a[i], -1.0f), 1.0f) * kPiOver4F;
}

// ---------------------------------------------------------------------------------------------------------
// (c) actor forward, fp32 on the CUDA cores.  One CTA = 64 observation rows per tile, persistent over tiles.
// Warp w owns rows 8w..8w+7; lane l owns output columns l, l+32, ...  Activations live in shared memory
// transposed ([k][row], row stride 68 floats) so that the 8 row operands of one k are two broadcast LDS.128;
// the k-major (pre-transposed) weights stream through a cp.async double buffer in chunks of 16 k.
// ---------------------------------------------------------------------------------------------------------
// per rows-per-warp variant R: row stride of the transposed activations (8 R rows + 4 pad, 16 B aligned) and k-rows per streamed
// weight chunk.  A small batch runs on ONE SM and is bound by the round trips of the chunk pipeline (the weights are 526 KB), so the
// variants with small activation buffers spend their shared memory on 3x larger chunks.
__host__ __device__ constexpr int rs_of(int R) { return 8 * R + 4; }
__host__ __device__ constexpr int kc_of(int R) { return R <= 2 ? 48 : (R == 4 ? 32 : 16); }

__device__ __forceinline__ void cp_async16(void *smem, const void *gmem) {
    const uint32_t s = (uint32_t)__cvta_generic_to_shared(smem);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(s), "l"(gmem));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N> __device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;\n" ::"n"(N)); }

__device__ __forceinline__ float warp_sum_f(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// acc[r][j] += sum_k act[k][row0w + r] * wT[k][lane + 32 j]  for k in [0, K); weights streamed from global.
// R = rows per warp (8 for full 64-row tiles; 4 / 2 / 1 for batches of <= 32 / 16 / 8 rows, where the padding rows of a
// 64-row tile would be 88 ... 98 % of the arithmetic of a kernel that runs on ONE SM)
template <int CJ, bool kGuard, int R>
__device__ __forceinline__ void gemm_stream(float (&acc)[R][CJ], const float *__restrict__ act /*smem [K][RS]*/,
                                            const float *__restrict__ wT /*global [K][WP]*/, int K, int WP,
                                            float *wbuf /*smem 2*KC*WP*/, int warp, int lane) {
    constexpr int RS = rs_of(R), KC = kc_of(R);
    const int tid = threadIdx.x;
    const int nchunks = (K + KC - 1) / KC;
    auto issue = [&](int ch, int buf) {
        const int k0 = ch * KC, kn = min(KC, K - k0);
        const int vec = kn * WP / 4;                         // WP % 32 == 0 -> rows are 16 B multiples
        const float4 *src = reinterpret_cast<const float4 *>(wT + (size_t)k0 * WP);
        float4 *dst = reinterpret_cast<float4 *>(wbuf + (size_t)buf * KC * WP);
        for (int v = tid; v < vec; v += kThreads) cp_async16(dst + v, src + v);
        cp_async_commit();
    };
    issue(0, 0);
    for (int ch = 0; ch < nchunks; ch++) {
        if (ch + 1 < nchunks) { issue(ch + 1, (ch + 1) & 1); cp_async_wait<1>(); } else cp_async_wait<0>();
        __syncthreads();
        const float *wb = wbuf + (size_t)(ch & 1) * KC * WP;
        const int k0 = ch * KC, kn = min(KC, K - k0);
#pragma unroll 4
        for (int kk = 0; kk < kn; kk++) {
            float a[R];
            if (R == 8) {
                const float4 a0 = *reinterpret_cast<const float4 *>(act + (size_t)(k0 + kk) * RS + warp * 8);
                const float4 a1 = *reinterpret_cast<const float4 *>(act + (size_t)(k0 + kk) * RS + warp * 8 + 4);
                const float t[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
#pragma unroll
                for (int r = 0; r < R; r++) a[r] = t[r];
            } else {
#pragma unroll
                for (int r = 0; r < R; r++) a[r] = act[(size_t)(k0 + kk) * RS + warp * R + r];
            }
            float w[CJ];
#pragma unroll
            for (int j = 0; j < CJ; j++) w[j] = (!kGuard || lane + 32 * j < WP) ? wb[kk * WP + lane + 32 * j] : 0.f;
#pragma unroll
            for (int r = 0; r < R; r++)
#pragma unroll
                for (int j = 0; j < CJ; j++) acc[r][j] = fmaf(a[r], w[j], acc[r][j]);
        }
        __syncthreads();
    }
}

// bias + LayerNorm (eps 1e-5, biased variance: torch.nn.LayerNorm) + ReLU on a warp-distributed row block
template <int CJ, bool kGuard, int R>
__device__ __forceinline__ void bias_ln_relu(float (&acc)[R][CJ], const float *__restrict__ bias, const float *__restrict__ g,
                                             const float *__restrict__ be, int H, int HP, int lane) {
    float bj[CJ], gj[CJ], bej[CJ];
#pragma unroll
    for (int j = 0; j < CJ; j++) {
        const int c = lane + 32 * j;
        const bool in = !kGuard || c < HP;
        bj[j] = in ? bias[c] : 0.f; gj[j] = in ? g[c] : 0.f; bej[j] = in ? be[c] : 0.f;
    }
    const float invH = 1.0f / (float)H;
#pragma unroll
    for (int r = 0; r < R; r++) {
        float s = 0.f;
#pragma unroll
        for (int j = 0; j < CJ; j++) { acc[r][j] += bj[j]; s += acc[r][j]; }      // padded columns are exactly 0
        const float mean = warp_sum_f(s) * invH;
        float q = 0.f;
#pragma unroll
        for (int j = 0; j < CJ; j++) { const float d = (lane + 32 * j < H) ? acc[r][j] - mean : 0.f; q = fmaf(d, d, q); }
        const float rstd = rsqrtf(warp_sum_f(q) * invH + 1e-5f);
#pragma unroll
        for (int j = 0; j < CJ; j++) acc[r][j] = fmaxf(fmaf((acc[r][j] - mean) * rstd, gj[j], bej[j]), 0.f);   // pad: g=be=0 -> 0
    }
}

template <int CJ1, int CJ2, bool kGuard, int R>
__global__ void __launch_bounds__(kThreads, 1) actor_fp32_kernel(tt_actor_dev A, const float *__restrict__ obs, int64_t ld,
                                                                int64_t n, float *__restrict__ out, TTRingS ring, TTActorTail tail) {
    constexpr int TM = 8 * R, RS = rs_of(R);            // rows per tile: 8 warps x R rows
    chain_enter();
    extern __shared__ __align__(16) float smem[];
    float *xs = smem;                                   // [k1p][RS]
    float *hs = xs + A.k1p * RS;                        // [h1p][RS]
    float *wbuf = hs + A.h1p * RS;                      // max(k1p*h1p, 2*KC*max(h1p,h2p))
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int64_t ntiles = (n + TM - 1) / TM;
    for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const int64_t row0 = tile * TM;
        const int rows = (int)min((int64_t)TM, n - row0);
        // ---- observation tile -> xs[k][r] (zero-padded) ----
        for (int v = threadIdx.x; v < A.k1p * RS; v += kThreads) xs[v] = 0.f;
        __syncthreads();
        for (int v = threadIdx.x; v < rows * A.in_dim; v += kThreads) {
            const int r = v / A.in_dim, k = v - r * A.in_dim;
            const float x = __ldcs(obs + (row0 + r) * ld + k);
            xs[k * RS + r] = x;
            if (ring.S && row0 + r >= ring.m.first) ring.S[ring.m.row(row0 + r) * A.in_dim + k] = x;     // fused replay store of s
        }
        __syncthreads();
        // ---- layer 1: fc1 -> LN -> ReLU (networks.py:139-141) ----
        {
            float acc[R][CJ1];
#pragma unroll
            for (int r = 0; r < R; r++)
#pragma unroll
                for (int j = 0; j < CJ1; j++) acc[r][j] = 0.f;
            gemm_stream<CJ1, kGuard, R>(acc, xs, A.w1t, A.k1p, A.h1p, wbuf, warp, lane);
            bias_ln_relu<CJ1, kGuard, R>(acc, A.b1, A.g1, A.be1, A.h1, A.h1p, lane);
#pragma unroll
            for (int j = 0; j < CJ1; j++) {
                const int c = lane + 32 * j;
                if (c < A.h1p) {
#pragma unroll
                    for (int r = 0; r < R; r++) hs[(size_t)c * RS + warp * R + r] = acc[r][j];
                }
            }
        }
        __syncthreads();
        // ---- layer 2: fc2 -> LN -> ReLU -> mu -> tanh (networks.py:142-145) ----
        {
            float acc[R][CJ2];
#pragma unroll
            for (int r = 0; r < R; r++)
#pragma unroll
                for (int j = 0; j < CJ2; j++) acc[r][j] = 0.f;
            gemm_stream<CJ2, kGuard, R>(acc, hs, A.w2t, A.h1, A.h2p, wbuf, warp, lane);
            bias_ln_relu<CJ2, kGuard, R>(acc, A.b2, A.g2, A.be2, A.h2, A.h2p, lane);
            float w3[CJ2];
#pragma unroll
            for (int j = 0; j < CJ2; j++) w3[j] = (!kGuard || lane + 32 * j < A.h2p) ? A.w3[lane + 32 * j] : 0.f;
            const float b3 = A.b3[0];
            float srow = 0.f;                                   // lane r keeps the output dot of row 8 warp + r
#pragma unroll
            for (int r = 0; r < R; r++) {
                float s = 0.f;
#pragma unroll
                for (int j = 0; j < CJ2; j++) s = fmaf(acc[r][j], w3[j], s);
                s = warp_sum_f(s);
                if (lane == r) srow = s;
            }
            // output stage, one lane per row: tanh, then what Agent.choose_action and the training loop do with it
            // (DDPG_agent.py:41-43 mu + OU noise, trainv2.py:516 clip * pi / 4, agent.remember of the raw action)
            const int row = warp * R + lane;
            if (lane < R && row < rows) {
                const int64_t gr = row0 + row;
                float a = tanhf(srow + b3);
                if (tail.ou_x) {
                    const float xn = ou_advance(tail.ou_x[gr], ttm::rng_normal_ks(tail.keys, tail.gid0 + (uint32_t)gr, *tail.iter));
                    tail.ou_x[gr] = xn;
                    a += xn;
                }
                out[gr] = a;
                if (tail.scaled) tail.scaled[gr] = fminf(fmaxf(a, -1.0f), 1.0f) * kPiOver4F;
                if (tail.ring.A && gr >= tail.ring.m.first) tail.ring.A[tail.ring.m.row(gr)] = a;
            }
        }
        __syncthreads();
    }
}

// ---------------------------------------------------------------------------------------------------------
// (c') the same fp32 forward for SMALL batches (<= 64 rows), split over a thread-block CLUSTER: 8 CTAs share 8 rows, each CTA
// owns one eighth of the hidden units of both layers.  The kernel above runs such a batch on ONE SM, which has to pull all
// 526 KB of fp32 weights through its chunk pipeline (29 us for one row); here every SM fetches its 66 KB slice of W2 with
// cp.async at kernel start -- in flight under layer 1 -- and the layers meet through distributed shared memory:
//   layer 1 (unit slice) -> per-CTA (mean, M2) of the slice -> all peers -> exact merge (Chan) -> LayerNorm + ReLU ->
//   all-gather of the activation slice into every peer's shared memory -> layer 2 (unit slice, K split over the 4 thread
//   groups) -> (mean, M2) exchange -> LayerNorm + ReLU + partial output dot -> CTA 0 -> tanh + the choose_action tail.
// Four cluster barriers; the arithmetic is fp32 FMA in k order like the big kernel (LayerNorm statistics merged pairwise).
// ---------------------------------------------------------------------------------------------------------
constexpr int kClu = 8, kCluRows = 8, kCluT = 256, kCluMaxU = 64, kCluMaxK1 = 32;
__device__ __forceinline__ uint32_t clu_rank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ void clu_sync() {
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// address of `p` (own shared memory) in CTA `rank` of the cluster
__device__ __forceinline__ uint32_t clu_map(const void *p, uint32_t rank) {
    uint32_t a = (uint32_t)__cvta_generic_to_shared(p), r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(rank));
    return r;
}
__device__ __forceinline__ void clu_st2(uint32_t addr, float x, float y) { asm volatile("st.shared::cluster.v2.f32 [%0], {%1, %2};" ::"r"(addr), "f"(x), "f"(y) : "memory"); }
__device__ __forceinline__ void clu_st1(uint32_t addr, float x) { asm volatile("st.shared::cluster.f32 [%0], %1;" ::"r"(addr), "f"(x) : "memory"); }

// (mean, M2) of `cnt` values merged into the running (n, mean, M2)
__device__ __forceinline__ void chan_merge(float &n, float &mean, float &m2, float cnt, float mean_c, float m2_c) {
    if (cnt <= 0.f) return;
    const float nt = n + cnt, d = mean_c - mean;
    mean += d * (cnt / nt);
    m2 += m2_c + d * d * (n * cnt / nt);
    n = nt;
}

// (mean, M2) over the units of this CTA's slice for the thread's two rows: the 64 threads of a row pair are two warps; two-pass
// (mean, then centred squares) with warp shuffles and one shared-memory hand-over per pass.  Every thread gets the result.
__device__ __forceinline__ void slice_stats(float v0, float v1, bool valid, float cnt, float (*red)[2], float &m0, float &m1, float &q0, float &q1) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, w0 = warp & ~1;
    float s0 = warp_sum_f(valid ? v0 : 0.f), s1 = warp_sum_f(valid ? v1 : 0.f);
    if (lane == 0) { red[warp][0] = s0; red[warp][1] = s1; }
    __syncthreads();
    const float inv = cnt > 0.f ? 1.0f / cnt : 0.f;
    m0 = (red[w0][0] + red[w0 + 1][0]) * inv; m1 = (red[w0][1] + red[w0 + 1][1]) * inv;
    __syncthreads();
    const float d0 = valid ? v0 - m0 : 0.f, d1 = valid ? v1 - m1 : 0.f;
    s0 = warp_sum_f(d0 * d0); s1 = warp_sum_f(d1 * d1);
    if (lane == 0) { red[warp][0] = s0; red[warp][1] = s1; }
    __syncthreads();
    q0 = red[w0][0] + red[w0 + 1][0]; q1 = red[w0][1] + red[w0 + 1][1];
}

__global__ void __cluster_dims__(kClu, 1, 1) __launch_bounds__(kCluT, 1)
actor_fp32_cluster_kernel(tt_actor_dev A, const float *__restrict__ obs, int64_t ld, int64_t n, float *__restrict__ out, TTRingS ring, TTActorTail tail) {
    chain_enter();
    extern __shared__ __align__(16) float w2s[];                 // this CTA's slice of W2^T: [h1p][U2]
    __shared__ __align__(16) float xs[kCluRows][kCluMaxK1];      // the cluster's observation rows
    __shared__ __align__(16) float a1s[512 * kCluRows];          // layer-1 activations of ALL units, [k][row] (filled by every CTA of the cluster)
    __shared__ __align__(16) float part[4][kCluMaxU][kCluRows];  // layer-2 partial sums of the four K groups
    __shared__ __align__(8) float st1[kClu][kCluRows][2], st2[kClu][kCluRows][2];   // per-CTA (mean, M2) of both layers, from every peer
    __shared__ float outp[kClu][kCluRows];                       // CTA 0: the peers' partial output dots
    __shared__ float red[kCluT / 32][2];
    const uint32_t rank = clu_rank();
    const int t = threadIdx.x, u = t & 63, grp = t >> 6;
    const int U1 = A.h1p / kClu, U2 = A.h2p / kClu;
    const int64_t row0 = (int64_t)(blockIdx.x / kClu) * kCluRows;
    const int rows = (int)min((int64_t)kCluRows, n - row0);
    // W2 slice: all of it in flight now, needed after two cluster barriers
    {
        const int vecs = U2 / 4;                                 // U2 % 4 == 0 (h2p is a multiple of 32)
        for (int v = t; v < A.h1p * vecs; v += kCluT) {
            const int k = v / vecs, q = v - k * vecs;
            cp_async16(w2s + (size_t)k * U2 + 4 * q, A.w2t + (size_t)k * A.h2p + rank * U2 + 4 * q);
        }
        cp_async_commit();
    }
    // layer-1 weights of this thread's unit (independent of the observation: issued first) and its parameters
    const int c1 = (int)rank * U1 + u;
    const bool on1 = u < U1;
    float w1[kCluMaxK1];
#pragma unroll
    for (int k = 0; k < kCluMaxK1; k++) w1[k] = (on1 && k < A.in_dim) ? __ldg(A.w1t + (size_t)k * A.h1p + c1) : 0.f;
    const float b1 = on1 ? __ldg(A.b1 + c1) : 0.f, g1 = on1 ? __ldg(A.g1 + c1) : 0.f, be1 = on1 ? __ldg(A.be1 + c1) : 0.f;
    for (int v = t; v < kCluRows * kCluMaxK1; v += kCluT) {
        const int r = v / kCluMaxK1, k = v - r * kCluMaxK1;
        float x = 0.f;
        if (r < rows && k < A.in_dim) {
            x = __ldcs(obs + (row0 + r) * ld + k);
            if (rank == 0 && ring.S && row0 + r >= ring.m.first) ring.S[ring.m.row(row0 + r) * A.in_dim + k] = x;     // fused replay store of s
        }
        xs[r][k] = x;
    }
    __syncthreads();
    // ---- layer 1: this thread = unit u of the slice, rows 2 grp and 2 grp + 1 ----
    float h0 = 0.f, h1v = 0.f;
#pragma unroll
    for (int k = 0; k < kCluMaxK1; k++) { h0 = fmaf(xs[2 * grp][k], w1[k], h0); h1v = fmaf(xs[2 * grp + 1][k], w1[k], h1v); }
    h0 += b1; h1v += b1;
    {
        const float cnt1 = (float)max(0, min(U1, A.h1 - (int)rank * U1));       // real (unpadded) units of this slice
        float m0, m1, q0, q1;
        slice_stats(h0, h1v, on1 && c1 < A.h1, cnt1, red, m0, m1, q0, q1);
        if (u < kClu) {                                          // thread u of each row pair tells peer u
            clu_st2(clu_map(&st1[rank][2 * grp][0], (uint32_t)u), m0, q0);
            clu_st2(clu_map(&st1[rank][2 * grp + 1][0], (uint32_t)u), m1, q1);
        }
    }
    clu_sync();                                                  // (1) every CTA has every slice's statistics
    {
        float mean[2], rstd[2];
#pragma unroll
        for (int i = 0; i < 2; i++) {
            float nn = 0.f, mm = 0.f, m2 = 0.f;
            for (int p = 0; p < kClu; p++) chan_merge(nn, mm, m2, (float)max(0, min(U1, A.h1 - p * U1)), st1[p][2 * grp + i][0], st1[p][2 * grp + i][1]);
            mean[i] = mm; rstd[i] = rsqrtf(m2 / (float)A.h1 + 1e-5f);
        }
        if (on1) {
            const bool real = c1 < A.h1;
            const float y0 = real ? fmaxf(fmaf((h0 - mean[0]) * rstd[0], g1, be1), 0.f) : 0.f;
            const float y1 = real ? fmaxf(fmaf((h1v - mean[1]) * rstd[1], g1, be1), 0.f) : 0.f;
            for (uint32_t p = 0; p < kClu; p++) clu_st2(clu_map(&a1s[(size_t)c1 * kCluRows + 2 * grp], p), y0, y1);
        }
    }
    cp_async_wait<0>();
    clu_sync();                                                  // (2) the full activation block is in every CTA; own W2 slice has landed
    __syncthreads();
    // ---- layer 2: unit u of the slice, K group grp, all 8 rows ----
    const int c2 = (int)rank * U2 + u;
    const bool on2 = u < U2;
    {
        float acc[kCluRows];
#pragma unroll
        for (int r = 0; r < kCluRows; r++) acc[r] = 0.f;
        const int kq = A.h1p / 4, k0 = grp * kq;
        if (on2) {
#pragma unroll 4
            for (int k = k0; k < k0 + kq; k++) {
                const float w = w2s[(size_t)k * U2 + u];
                const float4 a0 = *reinterpret_cast<const float4 *>(a1s + (size_t)k * kCluRows), a1 = *reinterpret_cast<const float4 *>(a1s + (size_t)k * kCluRows + 4);
                acc[0] = fmaf(a0.x, w, acc[0]); acc[1] = fmaf(a0.y, w, acc[1]); acc[2] = fmaf(a0.z, w, acc[2]); acc[3] = fmaf(a0.w, w, acc[3]);
                acc[4] = fmaf(a1.x, w, acc[4]); acc[5] = fmaf(a1.y, w, acc[5]); acc[6] = fmaf(a1.z, w, acc[6]); acc[7] = fmaf(a1.w, w, acc[7]);
            }
#pragma unroll
            for (int r = 0; r < kCluRows; r++) part[grp][u][r] = acc[r];
        }
    }
    __syncthreads();
    // thread = (unit u, rows 2 grp, 2 grp + 1) again: sum the K groups in order
    float z0 = 0.f, z1 = 0.f;
    if (on2) {
        const float b2 = __ldg(A.b2 + c2);
        z0 = ((part[0][u][2 * grp] + part[1][u][2 * grp]) + part[2][u][2 * grp]) + part[3][u][2 * grp] + b2;
        z1 = ((part[0][u][2 * grp + 1] + part[1][u][2 * grp + 1]) + part[2][u][2 * grp + 1]) + part[3][u][2 * grp + 1] + b2;
    }
    {
        const float cnt2 = (float)max(0, min(U2, A.h2 - (int)rank * U2));
        float m0, m1, q0, q1;
        slice_stats(z0, z1, on2 && c2 < A.h2, cnt2, red, m0, m1, q0, q1);
        if (u < kClu) {
            clu_st2(clu_map(&st2[rank][2 * grp][0], (uint32_t)u), m0, q0);
            clu_st2(clu_map(&st2[rank][2 * grp + 1][0], (uint32_t)u), m1, q1);
        }
    }
    clu_sync();                                                  // (3)
    {
        float y[2] = {0.f, 0.f};
        if (on2 && c2 < A.h2) {
            const float g2 = __ldg(A.g2 + c2), be2 = __ldg(A.be2 + c2), w3 = __ldg(A.w3 + c2);
#pragma unroll
            for (int i = 0; i < 2; i++) {
                float nn = 0.f, mm = 0.f, m2 = 0.f;
                for (int p = 0; p < kClu; p++) chan_merge(nn, mm, m2, (float)max(0, min(U2, A.h2 - p * U2)), st2[p][2 * grp + i][0], st2[p][2 * grp + i][1]);
                const float rstd = rsqrtf(m2 / (float)A.h2 + 1e-5f);
                y[i] = fmaxf(fmaf(((i ? z1 : z0) - mm) * rstd, g2, be2), 0.f) * w3;
            }
        }
        // partial output dot of this slice: sum over the 64 threads of a row pair (two warps), through shared memory
        y[0] = warp_sum_f(y[0]); y[1] = warp_sum_f(y[1]);
        if ((t & 31) == 0) { part[0][t >> 5][0] = y[0]; part[0][t >> 5][1] = y[1]; }
    }
    __syncthreads();
    if (t < kCluRows) {
        const int g = t >> 1, i = t & 1;                         // row t = rows 2 g + i: warps 2 g and 2 g + 1
        clu_st1(clu_map(&outp[rank][t], 0), part[0][2 * g][i] + part[0][2 * g + 1][i]);
    }
    clu_sync();                                                  // (4) CTA 0 has every slice's partial dot
    if (rank == 0 && t < rows) {
        float d = 0.f;
        for (int p = 0; p < kClu; p++) d += outp[p][t];
        const int64_t gr = row0 + t;
        float a = tanhf(d + __ldg(A.b3));
        if (tail.ou_x) {
            const float xn = ou_advance(tail.ou_x[gr], ttm::rng_normal_ks(tail.keys, tail.gid0 + (uint32_t)gr, *tail.iter));
            tail.ou_x[gr] = xn;
            a += xn;
        }
        out[gr] = a;
        if (tail.scaled) tail.scaled[gr] = fminf(fmaxf(a, -1.0f), 1.0f) * kPiOver4F;
        if (tail.ring.A && gr >= tail.ring.m.first) tail.ring.A[tail.ring.m.row(gr)] = a;
    }
}

size_t actor_fp32_smem(const tt_actor_dev &A, int R) {
    const int wmax = A.h1p > A.h2p ? A.h1p : A.h2p;
    size_t wb = (size_t)2 * kc_of(R) * wmax;
    return sizeof(float) * ((size_t)A.k1p * rs_of(R) + (size_t)A.h1p * rs_of(R) + wb);
}

}  // namespace

namespace tt {

template <int CJ1, int CJ2, bool kGuard, int R>
static int launch_fp32(const tt_actor_dev &A, const float *d_obs, int64_t ld, int64_t n, float *d_mu, const TTRingS &rs, const TTActorTail &tl,
                       cudaStream_t s) {
    auto kern = actor_fp32_kernel<CJ1, CJ2, kGuard, R>;
    const size_t smem = actor_fp32_smem(A, R);
    TT_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const int64_t ntiles = (n + 8 * R - 1) / (8 * R);
    const int grid = (int)(ntiles < tt::grid_sms() ? ntiles : tt::grid_sms());
    TT_CUDA(tt::launch_chained(tt::chain_rollout(n), kern, dim3((unsigned)grid), dim3(kThreads), smem, s, A, d_obs, ld, n, d_mu, rs, tl));
    TT_COUNT_LAUNCH(); TT_LAUNCH_CHECK();
    return TT_OK;
}

int actor_forward_fp32(const tt_actor *a, const float *d_obs, int64_t ld, int64_t n, float *d_mu, const TTRingS *ring, const TTActorTail *tail,
                       cudaStream_t s) {
    TTRingS rs;
    if (ring) rs = *ring; else { rs.S = nullptr; rs.m = tt_make_ring_map(1, 0, 0); }
    const TTActorTail tl = tail ? *tail : tt_no_tail();
    const tt_actor_dev &A = a->dev;
    const int cj1 = A.h1p / 32, cj2 = A.h2p / 32;
    // small batches: 8-CTA clusters, each CTA one eighth of the hidden units (see actor_fp32_cluster_kernel)
    if (n <= 256 && A.in_dim <= kCluMaxK1 && A.h1p <= 512 && A.h2p <= 512) {
        const size_t smem = sizeof(float) * (size_t)A.h1p * (A.h2p / kClu);
        static int ok_of[kMaxDevices] = {};               // per device: 0 = not probed, > 0 = co-resident clusters of 8 at this size, -1 = none
        static size_t smem_of[kMaxDevices] = {};
        static bool attr_of[kMaxDevices] = {};
        const int dev = device_index();
        int &ok = ok_of[dev];
        const unsigned grid = (unsigned)(kClu * ((n + kCluRows - 1) / kCluRows));
        if (ok == 0 || smem_of[dev] != smem) {
            ok = -1; smem_of[dev] = smem;
            if (!attr_of[dev]) attr_of[dev] = cudaFuncSetAttribute(actor_fp32_cluster_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                                                   (int)(sizeof(float) * 512 * (512 / kClu))) == cudaSuccess;
            if (attr_of[dev]) {
                cudaLaunchConfig_t q{};
                q.gridDim = dim3(kClu); q.blockDim = dim3(kCluT); q.dynamicSmemBytes = smem;
                int ncl = 0;
                if (cudaOccupancyMaxActiveClusters(&ncl, actor_fp32_cluster_kernel, &q) == cudaSuccess && ncl >= 1) ok = ncl;
            }
            (void)cudaGetLastError();
        }
        if (ok > 0 && grid <= (unsigned)(kClu * ok)) {          // one wave of clusters (37 at 400 / 300: every batch of the AUTO fp32 range)
            TT_CUDA(launch_chained(chain_rollout(n), actor_fp32_cluster_kernel, dim3(grid), dim3(kCluT), smem, s, A, d_obs, ld, n, d_mu, rs, tl));
            TT_COUNT_LAUNCH(); TT_LAUNCH_CHECK();
            return TT_OK;
        }
    }
    if (cj1 == 13 && cj2 == 10) {                        // the reference's 400 / 300: rows per warp by batch size
        if (n <= 8) return launch_fp32<13, 10, false, 1>(A, d_obs, ld, n, d_mu, rs, tl, s);
        if (n <= 16) return launch_fp32<13, 10, false, 2>(A, d_obs, ld, n, d_mu, rs, tl, s);
        if (n <= 32) return launch_fp32<13, 10, false, 4>(A, d_obs, ld, n, d_mu, rs, tl, s);
        return launch_fp32<13, 10, false, 8>(A, d_obs, ld, n, d_mu, rs, tl, s);
    }
    if (cj1 <= 16 && cj2 <= 16) return launch_fp32<16, 16, true, 8>(A, d_obs, ld, n, d_mu, rs, tl, s);
    set_error("actor: hidden sizes above 512 are not supported (h1=%d h2=%d)", A.h1, A.h2);
    return TT_ERR_INVALID;
}

int launch_noise(float *d_x, float *d_action, float *d_scaled, const uint8_t *d_reset_mask, int64_t n, uint64_t seed,
                 uint64_t gid0, const uint32_t *d_iter, int evaluate, const TTRingA *ring, cudaStream_t s) {
    TTRingA ra;
    if (ring) ra = *ring; else { ra.A = nullptr; ra.m = tt_make_ring_map(1, 0, 0); }
    const ttm::PhiloxKeys keys = ttm::philox_expand_key(seed);
    auto al16 = [](const void *p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; };
    const bool ring_ok = !ra.A || (!ra.m.many && ra.m.first == 0 && ra.m.base % 4 == 0 && ra.m.cap % 4 == 0 && al16(ra.A));
    if (!d_reset_mask && d_action && d_scaled && d_x && n % 4 == 0 && n >= 4096 && al16(d_x) && al16(d_action) && al16(d_scaled) && ring_ok) {
        const int64_t n4 = n / 4, want = (n4 + kThreads - 1) / kThreads, cap = (int64_t)sm_count() * 16;
        ou_kernel_x4<<<(unsigned)(want < cap ? want : cap), kThreads, 0, s>>>(
            reinterpret_cast<float4 *>(d_x), reinterpret_cast<float4 *>(d_action), reinterpret_cast<float4 *>(d_scaled), n4, keys, (uint32_t)gid0, d_iter,
            evaluate, reinterpret_cast<float4 *>(ra.A), ra.m.base / 4, ra.m.cap / 4);
    } else
        ou_kernel<<<(unsigned)((n + kThreads - 1) / kThreads), kThreads, 0, s>>>(d_x, d_action, d_scaled, d_reset_mask, n, keys, gid0, d_iter,
                                                                                evaluate, ra);
    TT_COUNT_LAUNCH(); TT_LAUNCH_CHECK();
    return TT_OK;
}

// TT_PREC_AUTO (north_star (c)): tcgen05 tiles only once the batch is a real dense contraction, warp-level fp32 FMA below.
// Measured on a B200 (profiles/r02_actor_auto_sweep.md, per launch inside a CUDA graph): the tensor-core kernel 8.7-9.1 us from 1 to
// 4 096 rows; the fp32 cluster kernel 13 us up to 64 rows, 17 / 19 us at 128 / 192.  There is no latency crossover -- the threshold
// is an accuracy policy: below it the batch is less than 1.5 tiles and the fp32 kernel is the more accurate one (1e-5 vs 1e-3).
constexpr int64_t kAutoTcMinRows = 192;
int actor_resolve_precision(const tt_actor *a, int precision, int64_t n) {
    if (precision != TT_PREC_AUTO) return precision;
    const tt_actor_dev &A = a->dev;
    const bool tc_ok = A.in_dim == 23 && A.h1 == 400 && A.h2 == 300;
    return tc_ok && n >= kAutoTcMinRows ? TT_PREC_F16 : TT_PREC_FP32;
}

int actor_forward_any(const tt_actor *a, const float *d_obs, int64_t ld, int64_t n, float *d_mu, int precision, const TTRingS *ring,
                      const TTActorTail *tail, cudaStream_t s) {
    precision = actor_resolve_precision(a, precision, n);
    if (precision == TT_PREC_FP32) return actor_forward_fp32(a, d_obs, ld, n, d_mu, ring, tail, s);
    if (precision == TT_PREC_BF16 || precision == TT_PREC_F16 || precision == TT_PREC_F16_PLAIN) return actor_forward_tc(a, d_obs, ld, n, d_mu, precision, ring, tail, s);
    set_error("unknown actor precision %d", precision);
    return TT_ERR_INVALID;
}

}  // namespace tt

extern "C" {

int tt_ou_step(float *d_x, float *d_action, const uint8_t *d_reset_mask, int64_t n, uint64_t seed,
               uint64_t global_env_offset, const uint32_t *d_iter, tt_stream_t stream) {
    TT_REQUIRE(d_x && d_iter && n > 0, "bad argument");
    return tt::launch_noise(d_x, d_action, nullptr, d_reset_mask, n, seed, global_env_offset, d_iter, 0, nullptr, tt::as_stream(stream));
}

int tt_ou_step_store(float *d_x, float *d_action, float *d_scaled, int64_t n, uint64_t seed, uint64_t global_env_offset,
                     const uint32_t *d_iter, int32_t evaluate, const tt_replay_ring *ring, tt_stream_t stream) {
    TT_REQUIRE(d_action && d_iter && n > 0 && (evaluate || d_x), "bad argument");
    TT_REQUIRE(ring && ring->d_action_mem && ring->mem_size > 0 && ring->mem_cntr >= 0, "bad ring");
    const TTRingA ra = {ring->d_action_mem, tt_make_ring_map(ring->mem_size, ring->mem_cntr, n)};
    return tt::launch_noise(d_x, d_action, d_scaled, nullptr, n, seed, global_env_offset, d_iter, evaluate, &ra, tt::as_stream(stream));
}

int tt_scale_action(const float *d_action, float *d_scaled, int64_t n, tt_stream_t stream) {
    TT_REQUIRE(d_action && d_scaled && n > 0, "bad argument");
    scale_kernel<<<(unsigned)((n + kThreads - 1) / kThreads), kThreads, 0, tt::as_stream(stream)>>>(d_action, d_scaled, n);
    TT_COUNT_LAUNCH(); TT_LAUNCH_CHECK();
    return TT_OK;
}

size_t tt_actor_workspace_bytes(int32_t in_dim, int32_t h1, int32_t h2) {
    if (in_dim <= 0 || h1 <= 0 || h2 <= 0) return 0;
    return tt_actor_layout(in_dim, h1, h2, nullptr, nullptr);
}

int tt_actor_create(tt_actor **out, int32_t in_dim, int32_t h1, int32_t h2, void *d_workspace, size_t workspace_bytes) {
    TT_REQUIRE(out && d_workspace, "NULL argument");
    TT_REQUIRE(in_dim > 0 && in_dim <= 32 && h1 > 0 && h1 <= 512 && h2 > 0 && h2 <= 512, "unsupported layer sizes");
    if ((reinterpret_cast<uintptr_t>(d_workspace) & 255) != 0 || workspace_bytes < tt_actor_workspace_bytes(in_dim, h1, h2)) {
        tt::set_error("tt_actor_create: workspace must be 256 B aligned and >= %zu bytes", tt_actor_workspace_bytes(in_dim, h1, h2));
        return TT_ERR_WORKSPACE;
    }
    if (tt_device_count() <= 0) { tt::set_error("tt_actor_create: no CUDA device (there is no CPU fallback)"); return TT_ERR_CUDA; }
    tt_actor *a = new (std::nothrow) tt_actor;
    TT_REQUIRE(a, "out of host memory");
    tt_actor_layout(in_dim, h1, h2, &a->dev, static_cast<char *>(d_workspace));
    a->loaded = false;
    *out = a;
    return TT_OK;
}

int tt_actor_destroy(tt_actor *a) { delete a; return TT_OK; }

int tt_actor_load(tt_actor *a, const float *d_fc1_w, const float *d_fc1_b, const float *d_ln1_g, const float *d_ln1_b,
                  const float *d_fc2_w, const float *d_fc2_b, const float *d_ln2_g, const float *d_ln2_b,
                  const float *d_mu_w, const float *d_mu_b, tt_stream_t stream) {
    TT_REQUIRE(a && d_fc1_w && d_fc1_b && d_ln1_g && d_ln1_b && d_fc2_w && d_fc2_b && d_ln2_g && d_ln2_b && d_mu_w && d_mu_b,
               "NULL argument");
    // fp32 images + the tensor-core operand images (tt_actor_tc4.cu): three launches
    int rc = tt::actor_pack_all(a, d_fc1_w, d_fc1_b, d_ln1_g, d_ln1_b, d_fc2_w, d_fc2_b, d_ln2_g, d_ln2_b, d_mu_w, d_mu_b, tt::as_stream(stream));
    if (rc != TT_OK) return rc;
    a->loaded = true;
    return TT_OK;
}

int tt_actor_forward(tt_actor *a, const float *d_obs, int64_t ld_obs, int64_t n, float *d_mu, int32_t precision,
                     tt_stream_t stream) {
    TT_REQUIRE(a && d_obs && d_mu, "NULL argument");
    TT_REQUIRE(a->loaded, "tt_actor_load has not been called");
    TT_REQUIRE(n > 0 && ld_obs >= a->dev.in_dim, "bad n / ld_obs");
    return tt::actor_forward_any(a, d_obs, ld_obs, n, d_mu, precision, nullptr, nullptr, tt::as_stream(stream));
}

int tt_actor_auto_precision(tt_actor *a, int64_t n) { return a ? tt::actor_resolve_precision(a, TT_PREC_AUTO, n) : TT_ERR_INVALID; }

int tt_actor_forward_store(tt_actor *a, const float *d_obs, int64_t ld_obs, int64_t n, float *d_mu, int32_t precision,
                           const tt_replay_ring *ring, tt_stream_t stream) {
    TT_REQUIRE(a && d_obs && d_mu, "NULL argument");
    TT_REQUIRE(a->loaded, "tt_actor_load has not been called");
    TT_REQUIRE(n > 0 && ld_obs >= a->dev.in_dim, "bad n / ld_obs");
    TT_REQUIRE(ring && ring->d_state_mem && ring->mem_size > 0 && ring->mem_cntr >= 0, "bad ring");
    const TTRingS rs = {ring->d_state_mem, tt_make_ring_map(ring->mem_size, ring->mem_cntr, n)};
    return tt::actor_forward_any(a, d_obs, ld_obs, n, d_mu, precision, &rs, nullptr, tt::as_stream(stream));
}

int tt_actor_choose_action(tt_actor *a, const float *d_obs, int64_t ld_obs, int64_t n, float *d_ou_x, uint64_t seed,
                           uint64_t global_env_offset, const uint32_t *d_iter, int32_t evaluate, float *d_action, float *d_scaled,
                           int32_t precision, const tt_replay_ring *ring, tt_stream_t stream) {
    TT_REQUIRE(a && d_obs && d_action, "NULL argument");
    TT_REQUIRE(a->loaded, "tt_actor_load has not been called");
    TT_REQUIRE(n > 0 && ld_obs >= a->dev.in_dim, "bad n / ld_obs");
    TT_REQUIRE(evaluate || (d_ou_x && d_iter), "noise needs d_ou_x and d_iter");
    TT_REQUIRE(!ring || (ring->d_state_mem && ring->d_action_mem && ring->mem_size > 0 && ring->mem_cntr >= 0), "bad ring");
    TTActorTail tl = tt_no_tail();
    tl.ou_x = evaluate ? nullptr : d_ou_x; tl.scaled = d_scaled; tl.keys = ttm::philox_expand_key(seed);
    tl.gid0 = (uint32_t)global_env_offset; tl.iter = d_iter;
    TTRingS rs; rs.S = nullptr; rs.m = tt_make_ring_map(1, 0, 0);
    if (ring) { rs.S = ring->d_state_mem; rs.m = tt_make_ring_map(ring->mem_size, ring->mem_cntr, n); tl.ring.A = ring->d_action_mem; tl.ring.m = rs.m; }
    return tt::actor_forward_any(a, d_obs, ld_obs, n, d_action, precision, ring ? &rs : nullptr, &tl, tt::as_stream(stream));
}

}  // extern "C"
