// tt_learn.cu -- row f1: one DDPG update (Agent.learn, DDPG/DDPG_agent.py:72-131; CriticNetwork DDPG/networks.py:9-68,
// ActorNetwork :98-147; ReplayBuffer.sample_buffer DDPG/replay_buffer.py:23-34) as 15 small hand-written kernels on the
// caller's stream, on the device-resident replay ring, followed by the re-pack of the new policy into the rollout actor's
// operand images.  The whole sequence is capturable into one CUDA graph (nothing synchronises, the Adam step counter and
// the sampling counter live in device memory).
//
// The step is latency-bound (batch 64 x 23-400-300 networks = 150 MFLOP): what matters is the length of the dependency
// chain, not the arithmetic.  Stages (one launch each):
//   K0  sample 64 ring rows (Philox) and gather s, a, r, s', done
//   K1  fc1 + LayerNorm 1 + ReLU of target_actor(s'), target_critic(s'), critic(s), actor(s)   -- one launch, 4 jobs
//   K2  fc2 of the same four (grouped GEMM launch, cp.async double-buffered operands)
//   K3  critic head (8 CTAs, warp = batch row; the last CTA sums the columns): a' = target_actor head, y = r + gamma Q'(s', a') (1 - done), q = Q(s, a),
//       dL/dq = 2 (q - y) / B, back through q / action_value / LayerNorm 2 -> d h2; column sums = their parameter gradients
//   K4  critic fc2 backward: dW2 = dh2^T a1 (grouped with) da1 = dh2 W2
//   K5  critic LayerNorm 1 backward (warp = row), K5b fc1 backward dW1 = dh1^T s (a dW job)
//   K6  Adam (weight decay 0.01) on the critic + soft update of target_critic
//   K7  fc1 + LayerNorm 1 + ReLU, K8 fc2 of the UPDATED critic on s
//   K9  actor head: a = actor head, dL/da = -(1/B) dQ/da through relu / action_value, back through tanh / mu /
//       LayerNorm 2 -> d h2 of the actor
//   K10 actor fc2 backward, K11 actor LayerNorm 1 + fc1 backward, K12 Adam on the actor + soft update of target_actor
// Every reduction has a fixed order (no atomics): the step is deterministic.
#include <new>
#include "tt_actor.cuh"
#include "tt_common.cuh"

extern "C" int tt_actor_load(tt_actor *a, const float *d_fc1_w, const float *d_fc1_b, const float *d_ln1_g, const float *d_ln1_b,
                             const float *d_fc2_w, const float *d_fc2_b, const float *d_ln2_g, const float *d_ln2_b,
                             const float *d_mu_w, const float *d_mu_b, tt_stream_t stream);

namespace {

constexpr int kMaxB = 64;            // batch rows per update (the reference default, trainv2.py:408)
constexpr int kMaxH = 512;           // hidden width limit (16 values per lane when a warp holds one row)
constexpr float kLnEps = 1e-5f;      // torch.nn.LayerNorm default
constexpr int kT = 256;              // threads of the grouped GEMM kernel

// ---- flat parameter layout of one network (floats); the first eight tensors are shared by actor and critic ----
struct Layout {
    int in, h1, h2;
    __host__ __device__ int w1() const { return 0; }
    __host__ __device__ int b1() const { return h1 * in; }
    __host__ __device__ int g1() const { return b1() + h1; }
    __host__ __device__ int be1() const { return g1() + h1; }
    __host__ __device__ int w2() const { return be1() + h1; }
    __host__ __device__ int b2() const { return w2() + h2 * h1; }
    __host__ __device__ int g2() const { return b2() + h2; }
    __host__ __device__ int be2() const { return g2() + h2; }
    __host__ __device__ int tail() const { return be2() + h2; }
    // actor: mu.weight[h2] mu.bias[1]          critic: action_value.weight[h2] action_value.bias[h2] q.weight[h2] q.bias[1]
    __host__ __device__ int actor_count() const { return tail() + h2 + 1; }
    __host__ __device__ int critic_count() const { return tail() + 3 * h2 + 1; }
};

enum { NET_ACTOR = 0, NET_TARGET_ACTOR = 1, NET_CRITIC = 2, NET_TARGET_CRITIC = 3 };
enum { JOB_TA = 0, JOB_TC = 1, JOB_C = 2, JOB_A = 3, NJOBS = 4 };     // forward passes whose activations are kept

// ---- grouped GEMM jobs (plain fp32 operands, cp.async staging) ----
enum { G_FWD = 0, G_WGRAD = 1, G_XGRAD = 2 };
struct Job {
    int type, ctas;                   // CTAs this job occupies in the launch
    int B, N, K;
    // G_FWD   : Y[b][n] = bias[n] + sum_k X[b][k] W[n][k]                       X [B][K], W [N][K]
    // G_WGRAD : dW[n][k] = sum_b D[b][n] X[b][k],  db[n] = sum_b D[b][n]          D [B][N], X [B][K]
    // G_XGRAD : dX[b][k] = sum_n D[b][n] W[n][k]                                 D [B][N], W [N][K]
    const float *X; int ldx;
    const float *W; int ldw;
    const float *bias;
    const float *D; int ldd;
    float *Y; int ldy;                // output: Y / dW / dX
    float *db;
};
constexpr int kMaxJobs = 4;
struct JobList { Job j[kMaxJobs]; int n; };

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// global -> shared staging of a [rows x 32] float block (row r at dst + r * kP, source row r at src + r * ld, columns
// [c0, c0 + 32) clipped to `cols`; rows >= nrows and columns >= cols are zero-filled).  16 B cp.async when the source rows are
// 16 B aligned (ld % 4 == 0 and cols % 4 == 0: the production sizes), 4 B cp.async otherwise.  No registers, no per-element
// arithmetic: the LayerNorm / ReLU of the operands is materialised by the producer kernels.
constexpr int kP = 36;                                                          // shared-memory row pitch (floats): 16 B aligned, LDS.128 conflict-free
__device__ __forceinline__ void cp16(float *dst, const float *src, bool ok) {
    const uint32_t d = (uint32_t)__cvta_generic_to_shared(dst);
    const int sz = ok ? 16 : 0;                                                 // src-size 0: zero fill
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(d), "l"(src), "r"(sz) : "memory");
}
__device__ __forceinline__ void cp4(float *dst, const float *src, bool ok) {
    const uint32_t d = (uint32_t)__cvta_generic_to_shared(dst);
    const int sz = ok ? 4 : 0;
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;" ::"r"(d), "l"(src), "r"(sz) : "memory");
}
__device__ __forceinline__ void stage_block(float *dst, const float *__restrict__ src, int ld, int rows, int nrows, int c0, int cols, bool vec) {
    if (vec) {
        for (int e = threadIdx.x; e < rows * 8; e += kT) {
            const int r = e >> 3, q = e & 7, c = c0 + 4 * q;
            const bool ok = r < nrows && c < cols;
            cp16(dst + r * kP + 4 * q, ok ? src + (size_t)r * ld + c : src, ok);
        }
    } else {
        for (int e = threadIdx.x; e < rows * 32; e += kT) {
            const int r = e >> 5, q = e & 31, c = c0 + q;
            const bool ok = r < nrows && c < cols;
            cp4(dst + r * kP + q, ok ? src + (size_t)r * ld + c : src, ok);
        }
    }
}
__device__ __forceinline__ void cp_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void cp_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }
__device__ __forceinline__ bool vec_ok(const float *p, int ld, int cols) { return (ld & 3) == 0 && (cols & 3) == 0 && ((uintptr_t)p & 15) == 0; }

// Y tile: all B rows x 16 columns [n0, n0 + 16); K in double-buffered chunks of 32.  Thread = 4 rows x 1 column; per 4 k: four
// 16 B loads of X and one of W for 16 FMAs.
__device__ void job_fwd(const Job &J, int cta, float *smem) {
    float *Xs = smem, *Ws = smem + 2 * 64 * kP;                                 // [2][64][kP], [2][16][kP]
    const int tid = threadIdx.x, tn = tid & 15, tb = tid >> 4, n0 = cta * 16;
    const bool vx = vec_ok(J.X, J.ldx, J.K), vw = vec_ok(J.W, J.ldw, J.K);
    const int nch = (J.K + 31) / 32;
    auto stage = [&](int ch, int buf) {
        stage_block(Xs + buf * 64 * kP, J.X, J.ldx, 64, J.B, ch * 32, J.K, vx);
        stage_block(Ws + buf * 16 * kP, J.W + (size_t)n0 * J.ldw, J.ldw, 16, J.N - n0, ch * 32, J.K, vw);
        cp_commit();
    };
    stage(0, 0);
    float acc[4] = {0.f, 0.f, 0.f, 0.f};
    for (int ch = 0; ch < nch; ch++) {
        if (ch + 1 < nch) { stage(ch + 1, (ch + 1) & 1); cp_wait<1>(); } else cp_wait<0>();
        __syncthreads();
        const float *xs = Xs + (ch & 1) * 64 * kP + 4 * tb * kP, *ws = Ws + (ch & 1) * 16 * kP + tn * kP;
#pragma unroll
        for (int q = 0; q < 8; q++) {
            const float4 w = *reinterpret_cast<const float4 *>(ws + 4 * q);
#pragma unroll
            for (int i = 0; i < 4; i++) {
                const float4 x = *reinterpret_cast<const float4 *>(xs + i * kP + 4 * q);
                acc[i] = fmaf(x.x, w.x, acc[i]); acc[i] = fmaf(x.y, w.y, acc[i]); acc[i] = fmaf(x.z, w.z, acc[i]); acc[i] = fmaf(x.w, w.w, acc[i]);
            }
        }
        __syncthreads();
    }
    if (n0 + tn < J.N) {
        const float bias = J.bias ? J.bias[n0 + tn] : 0.f;
#pragma unroll
        for (int i = 0; i < 4; i++) {
            const int b = 4 * tb + i;
            if (b < J.B) J.Y[(size_t)b * J.ldy + n0 + tn] = acc[i] + bias;
        }
    }
}

// dW tile: 16 rows n x 32 columns k; the reduction runs over the batch (B <= 64) in one stage.  Thread = 1 n x 2 k.
__device__ void job_wgrad(const Job &J, int cta, float *smem) {
    float *Xs = smem, *Ds = smem + 64 * kP;                                     // X: [64 b][kP] (32 k), D: [64 b][kP] (16 n used)
    const int ktiles = (J.K + 31) / 32;
    const int nt = cta / ktiles, kt = cta - nt * ktiles, n0 = nt * 16, k0 = kt * 32;
    const int tid = threadIdx.x, tk = tid & 15, tn = tid >> 4;
    stage_block(Xs, J.X, J.ldx, 64, J.B, k0, J.K, vec_ok(J.X, J.ldx, J.K));
    // D columns [n0, n0 + 16): a 32-wide block whose upper half is clipped away
    stage_block(Ds, J.D, J.ldd, 64, J.B, n0, min(J.N, n0 + 16), vec_ok(J.D, J.ldd, J.N) && ((min(J.N, n0 + 16) & 3) == 0));
    cp_commit(); cp_wait<0>();
    __syncthreads();
    float a0 = 0.f, a1 = 0.f, ab = 0.f;
#pragma unroll 8
    for (int b = 0; b < 64; b++) {
        const float d = Ds[b * kP + tn];
        const float2 x = *reinterpret_cast<const float2 *>(Xs + b * kP + 2 * tk);
        ab += d; a0 = fmaf(d, x.x, a0); a1 = fmaf(d, x.y, a1);
    }
    if (n0 + tn < J.N) {
        const int k = k0 + 2 * tk;
        if (k < J.K) J.Y[(size_t)(n0 + tn) * J.ldy + k] = a0;
        if (k + 1 < J.K) J.Y[(size_t)(n0 + tn) * J.ldy + k + 1] = a1;
        if (J.db && kt == 0 && tk == 0) J.db[n0 + tn] = ab;
    }
}

// dX tile: 32 batch rows x 32 columns [k0, k0 + 32); N in double-buffered chunks of 32.  Thread = 1 row x 4 columns; per 4 n:
// one 16 B load of D and four of W for 16 FMAs.  (26 CTAs for 64 x 400: the reduction over 300 n is the long chain of the
// backward pass, so it gets the small tile.)
__device__ void job_xgrad(const Job &J, int cta, float *smem) {
    float *Ds = smem, *Ws = smem + 2 * 32 * kP;                                 // [2][32 b][kP] (32 n), [2][32 n][kP] (32 k)
    const int ktiles = (J.K + 31) / 32;
    const int bt = cta / ktiles, kt = cta - bt * ktiles, b0 = bt * 32, k0 = kt * 32;
    const int tid = threadIdx.x, tk = tid & 7, tb = tid >> 3;
    const bool vd = vec_ok(J.D, J.ldd, J.N), vw = vec_ok(J.W, J.ldw, J.K);
    const int nch = (J.N + 31) / 32;
    auto stage = [&](int ch, int buf) {
        stage_block(Ds + buf * 32 * kP, J.D + (size_t)b0 * J.ldd, J.ldd, 32, J.B - b0, ch * 32, J.N, vd);
        stage_block(Ws + buf * 32 * kP, J.W + (size_t)ch * 32 * J.ldw, J.ldw, 32, J.N - ch * 32, k0, J.K, vw);
        cp_commit();
    };
    stage(0, 0);
    float acc[4] = {0.f, 0.f, 0.f, 0.f};
    for (int ch = 0; ch < nch; ch++) {
        if (ch + 1 < nch) { stage(ch + 1, (ch + 1) & 1); cp_wait<1>(); } else cp_wait<0>();
        __syncthreads();
        const float *ds = Ds + (ch & 1) * 32 * kP + tb * kP, *ws = Ws + (ch & 1) * 32 * kP + 4 * tk;
#pragma unroll
        for (int q = 0; q < 8; q++) {
            const float4 d = *reinterpret_cast<const float4 *>(ds + 4 * q);
            const float dd[4] = {d.x, d.y, d.z, d.w};
#pragma unroll
            for (int j = 0; j < 4; j++) {
                const float4 w = *reinterpret_cast<const float4 *>(ws + (4 * q + j) * kP);
                acc[0] = fmaf(dd[j], w.x, acc[0]); acc[1] = fmaf(dd[j], w.y, acc[1]);
                acc[2] = fmaf(dd[j], w.z, acc[2]); acc[3] = fmaf(dd[j], w.w, acc[3]);
            }
        }
        __syncthreads();
    }
    const int b = b0 + tb;
    if (b < J.B) {
#pragma unroll
        for (int j = 0; j < 4; j++) { const int k = k0 + 4 * tk + j; if (k < J.K) J.Y[(size_t)b * J.ldy + k] = acc[j]; }
    }
}

__global__ void __launch_bounds__(kT) learn_gemm_kernel(JobList L) {
    chain_enter();
    __shared__ __align__(16) float smem[2 * 64 * kP + 2 * 32 * kP];            // 27 648 B: the dX layout is the largest
    int cta = blockIdx.x;
    for (int i = 0; i < L.n; i++) {
        if (cta < L.j[i].ctas) {
            const Job &J = L.j[i];
            if (J.type == G_FWD) job_fwd(J, cta, smem);
            else if (J.type == G_WGRAD) job_wgrad(J, cta, smem);
            else job_xgrad(J, cta, smem);
            return;
        }
        cta -= L.j[i].ctas;
    }
}

// ---- K0: sampling + gather (replay_buffer.py:23-34: uniform with replacement over the filled part of the ring) ----
struct Batch {
    float *s, *s2, *a, *r, *d;        // [B][in], [B][in], [B], [B], [B] (done as 0 / 1)
    int64_t *rows;                    // [B]
};
__global__ void learn_gather_kernel(tt_replay_ring ring, int64_t win_begin, int64_t max_mem, const int64_t *__restrict__ given_rows, Batch bt, int B,
                                    int in, uint64_t seed, int *__restrict__ step) {
    chain_enter();
    __shared__ int64_t rows[kMaxB];
    const int tid = threadIdx.x;
    const int t = *step;                               // updates done so far = the sampling counter of this one
    __syncthreads();
    if (tid == 0) *step = t + 1;                       // the optimizer kernels of this update read t + 1 (Adam's step number)
    if (tid < B) {
        int64_t row;
        if (given_rows) row = given_rows[tid];
        else {
            uint32_t w[4];
            ttm::philox4x32_10((uint32_t)tid, (uint32_t)t, 2u, 0u, (uint32_t)seed, (uint32_t)(seed >> 32), w);
            const double u = ((double)w[0] * 4294967296.0 + (double)w[1]) * (1.0 / 18446744073709551616.0);       // [0, 1)
            row = (int64_t)(u * (double)max_mem);
            if (row >= max_mem) row = max_mem - 1;
            row = (win_begin + row) % ring.mem_size;                  // the sampling window may wrap around the ring
        }
        rows[tid] = row; bt.rows[tid] = row;
        bt.a[tid] = ring.d_action_mem[row]; bt.r[tid] = ring.d_reward_mem[row]; bt.d[tid] = ring.d_terminal_mem[row] ? 1.f : 0.f;
    }
    __syncthreads();
    // 1024 threads: at most two elements each, all loads issued before the first store (one DRAM latency)
    float vs[2], vs2[2];
#pragma unroll
    for (int i = 0; i < 2; i++) {
        const int v = tid + (int)blockDim.x * i;
        if (v < B * in) { const int b = v / in, c = v - b * in; vs[i] = ring.d_state_mem[rows[b] * in + c]; vs2[i] = ring.d_new_state_mem[rows[b] * in + c]; }
    }
#pragma unroll
    for (int i = 0; i < 2; i++) {
        const int v = tid + (int)blockDim.x * i;
        if (v < B * in) { bt.s[v] = vs[i]; bt.s2[v] = vs2[i]; }
    }
    for (int v = tid + 2 * (int)blockDim.x; v < B * in; v += blockDim.x) {       // (only for blocks smaller than 1024 threads)
        const int b = v / in, c = v - b * in;
        bt.s[v] = ring.d_state_mem[rows[b] * in + c];
        bt.s2[v] = ring.d_new_state_mem[rows[b] * in + c];
    }
}

// ---- row-wise stages: 8 CTAs x 8 warps, one warp per batch row with the row in registers (one warp per row keeps the
//      dependent chain of a row -- load -> statistics -> head -> backward -- at one memory latency per network instead of
//      serialising rows); the LAST CTA to finish (a ticket counter) then takes the column sums over the batch (= the parameter
//      gradients), one column per thread, rows in order (deterministic).  No cluster and no grid barrier: the CTAs need not be
//      co-resident, so the stage also runs on the two or four SMs that rollout.AsyncTrainer leaves free for the learner.
constexpr int kRowT = 256, kPerLane = kMaxH / 32;
struct Row { float v[kPerLane]; };

// true in exactly one CTA of the launch: the one that arrives last, after every other CTA's global writes are visible
__device__ __forceinline__ bool last_cta_done(int *ticket) {
    __shared__ int s_last;
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) {
        const int t = atomicAdd(ticket, 1);
        s_last = t == (int)gridDim.x - 1;
        if (s_last) *ticket = 0;                            // ready for the next row-wise stage (stream order)
    }
    __syncthreads();
    if (s_last) __threadfence();
    return s_last != 0;
}

__device__ __forceinline__ void load_row(Row &r, const float *__restrict__ p, int H, int lane) {
#pragma unroll
    for (int i = 0; i < kPerLane; i++) { const int j = lane + 32 * i; r.v[i] = j < H ? p[j] : 0.f; }
}
// in place: x -> xhat = (x - mean) rstd  (padding lanes stay 0); returns rstd
__device__ __forceinline__ float normalize_row(Row &r, int H, int lane) {
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < kPerLane; i++) s += r.v[i];
    const float mean = warp_sum(s) / (float)H;
    float q = 0.f;
#pragma unroll
    for (int i = 0; i < kPerLane; i++) { const int j = lane + 32 * i; const float d = j < H ? r.v[i] - mean : 0.f; r.v[i] = d; q = fmaf(d, d, q); }
    const float rstd = rsqrtf(warp_sum(q) / (float)H + kLnEps);
#pragma unroll
    for (int i = 0; i < kPerLane; i++) r.v[i] *= rstd;
    return rstd;
}
// LayerNorm backward for one row: given do (gradient w.r.t. the LayerNorm output), xhat, g: dx = rstd (t - mean(t) - xhat mean(t xhat)), t = do g
__device__ __forceinline__ void ln_backward_row(const Row &dout, const Row &xhat, const float *__restrict__ g, float rstd, int H, int lane, Row &dx) {
    float c1 = 0.f, c2 = 0.f;
#pragma unroll
    for (int i = 0; i < kPerLane; i++) {
        const int j = lane + 32 * i;
        const float t = j < H ? dout.v[i] * g[j] : 0.f;
        dx.v[i] = t; c2 += t; c1 = fmaf(t, xhat.v[i], c1);
    }
    c1 = warp_sum(c1) / (float)H; c2 = warp_sum(c2) / (float)H;
#pragma unroll
    for (int i = 0; i < kPerLane; i++) dx.v[i] = rstd * (dx.v[i] - c2 - xhat.v[i] * c1);
}

struct HeadArgs {
    int B, H1, H2;
    float gamma;
    // activations (pre-LayerNorm outputs of fc2) of the four forward jobs, [B][H2]
    const float *h2[NJOBS];
    // parameters
    const float *ta_g2, *ta_be2, *ta_w3, *ta_b3;                         // target actor head
    const float *tc_g2, *tc_be2, *tc_wa, *tc_ba, *tc_wq, *tc_bq;          // target critic head
    const float *c_g2, *c_be2, *c_wa, *c_ba, *c_wq, *c_bq;                // critic head
    const float *a_g2, *a_be2, *a_w3, *a_b3;                              // actor head
    const float *act, *rew, *done;                                        // batch
    float *dh2;                                                           // out: gradient w.r.t. the fc2 output [B][H2]
    float *sc0, *sc1, *sc2;                                               // scratch [B][H2]
    float *dv;                                                            // scratch [B]: dL/dq (critic) or dL/d(pre-tanh) (actor) per row
    int *ticket;                                                          // last-CTA counter (0 between launches)
    // gradients (flat-layout pointers)
    float *g_g2, *g_be2, *g_t0, *g_t1, *g_t2, *g_t3;                      // critic: wa, ba, wq, bq | actor: w3, b3, -, -
    float *q_out, *y_out, *a_out;                                         // diagnostics / hand-over: Q(s,a), target, actor(s)
};

// K1 / K7: fc1 + LayerNorm 1 + ReLU of up to four networks in one launch.  A CTA takes 8 batch rows of one network (warp = row,
// the whole fc1 output row -- <= 512 values -- in registers), W1 passes through shared memory in blocks of 128 output columns.
// Writes h1 (pre-LayerNorm, for the backward pass) and a1 = relu(LN1(h1)) (the operand of fc2 and of its weight gradient).
struct Fc1Job { const float *x, *w1, *b1, *g1, *be1; float *h1, *a1; };
struct Fc1Args { Fc1Job j[NJOBS]; int njobs, B, IN, H1; };
constexpr int kW1P = 25;                                   // shared-memory pitch of a W1 row (IN <= 24 + 1; odd: conflict-free for lane = column)
__global__ void __launch_bounds__(256) learn_fc1_kernel(Fc1Args A) {
    chain_enter();
    extern __shared__ float ws[];                          // W1 [H1][kW1P]: the whole matrix in ONE round of independent loads
    __shared__ float xs[8 * 32];                           // the CTA's 8 input rows
    const int per = (A.B + 7) / 8;
    const int job = blockIdx.x / per, b0 = (blockIdx.x - job * per) * 8;
    const Fc1Job J = A.j[job];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, b = b0 + warp, H = A.H1, IN = A.IN;
    // this lane's columns of the three parameter vectors: loaded up front, together with W1 (one memory latency for everything)
    float pb[kPerLane], pg[kPerLane], pe[kPerLane];
#pragma unroll
    for (int i = 0; i < kPerLane; i++) {
        const int j = lane + 32 * i;
        pb[i] = j < H ? J.b1[j] : 0.f; pg[i] = j < H ? J.g1[j] : 0.f; pe[i] = j < H ? J.be1[j] : 0.f;
    }
    for (int v = threadIdx.x; v < 8 * IN; v += 256) { const int r = v / IN, k = v - r * IN; xs[r * 32 + k] = b0 + r < A.B ? J.x[(size_t)(b0 + r) * IN + k] : 0.f; }
    for (int v = threadIdx.x; v < H * IN; v += 256) { const int c = v / IN, k = v - c * IN; ws[c * kW1P + k] = J.w1[v]; }
    __syncthreads();
    Row h;
#pragma unroll
    for (int i = 0; i < kPerLane; i++) {
        const int c = lane + 32 * i;
        float acc = 0.f;
        if (c < H) {
            for (int k = 0; k < IN; k++) acc = fmaf(xs[warp * 32 + k], ws[c * kW1P + k], acc);
            acc += pb[i];
        }
        h.v[i] = acc;
    }
    if (b < A.B) {
#pragma unroll
        for (int i = 0; i < kPerLane; i++) { const int j = lane + 32 * i; if (j < H) J.h1[(size_t)b * H + j] = h.v[i]; }
        normalize_row(h, H, lane);
#pragma unroll
        for (int i = 0; i < kPerLane; i++) { const int j = lane + 32 * i; if (j < H) J.a1[(size_t)b * H + j] = fmaxf(fmaf(h.v[i], pg[i], pe[i]), 0.f); }
    }
}

// The head / LayerNorm-backward kernels read a dozen parameter vectors element by element inside dependent arithmetic; from
// global memory the compiler issues those loads just in time, one L2 latency after the other (measured: 27 us for 150 k
// instructions).  All vectors are copied to shared memory first: NV loads per thread in flight at once, one latency in total.
template <int NV>
__device__ __forceinline__ void stage_vectors(float (*dst)[kMaxH], const float *const (&src)[NV], int H) {
    for (int j = threadIdx.x; j < H; j += kRowT) {
        float t[NV];
#pragma unroll
        for (int v = 0; v < NV; v++) t[v] = src[v][j];
#pragma unroll
        for (int v = 0; v < NV; v++) dst[v][j] = t[v];
    }
}

// K3: DDPG_agent.py:84-97
__global__ void __launch_bounds__(kRowT) learn_critic_head_kernel(HeadArgs A) {
    chain_enter();
    __shared__ float s_dq[kMaxB], s_act[kMaxB];
    __shared__ float sp[13][kMaxH];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, H = A.H2;
    const int b = (int)blockIdx.x * (kRowT / 32) + warp;
    {
        const float *const src[13] = {A.ta_g2, A.ta_be2, A.ta_w3, A.tc_g2, A.tc_be2, A.tc_wa, A.tc_ba, A.tc_wq, A.c_g2, A.c_be2, A.c_wa, A.c_ba, A.c_wq};
        stage_vectors<13>(sp, src, H);
    }
    const float *ta_g2 = sp[0], *ta_be2 = sp[1], *ta_w3 = sp[2], *tc_g2 = sp[3], *tc_be2 = sp[4], *tc_wa = sp[5], *tc_ba = sp[6], *tc_wq = sp[7],
                *c_g2 = sp[8], *c_be2 = sp[9], *c_wa = sp[10], *c_ba = sp[11], *c_wq = sp[12];
    __syncthreads();
    if (b < A.B) {
        Row x, xt;
        load_row(xt, A.h2[JOB_TA] + (size_t)b * H, H, lane);             // the three rows' loads are independent: issue them together
        Row xc;
        load_row(xc, A.h2[JOB_TC] + (size_t)b * H, H, lane);
        load_row(x, A.h2[JOB_C] + (size_t)b * H, H, lane);
        // a' = target_actor(s') head: tanh(mu(relu(LN2(h2))))            (networks.py:142-145)
        normalize_row(xt, H, lane);
        float p = 0.f;
#pragma unroll
        for (int i = 0; i < kPerLane; i++) { const int j = lane + 32 * i; if (j < H) p = fmaf(fmaxf(fmaf(xt.v[i], ta_g2[j], ta_be2[j]), 0.f), ta_w3[j], p); }
        const float a2 = tanhf(warp_sum(p) + A.ta_b3[0]);
        // Q'(s', a') = q(relu(LN2(h2) + action_value(a')))                 (networks.py:53-68)
        normalize_row(xc, H, lane);
        float q2 = 0.f;
#pragma unroll
        for (int i = 0; i < kPerLane; i++) {
            const int j = lane + 32 * i;
            if (j < H) q2 = fmaf(fmaxf(fmaf(xc.v[i], tc_g2[j], tc_be2[j]) + fmaf(a2, tc_wa[j], tc_ba[j]), 0.f), tc_wq[j], q2);
        }
        q2 = warp_sum(q2) + A.tc_bq[0];
        const float y = A.rew[b] + A.gamma * (A.done[b] != 0.f ? 0.f : q2);       // DDPG_agent.py:90-93
        // Q(s, a) and its backward
        const float rstd = normalize_row(x, H, lane);
        const float act = A.act[b];
        Row z;
        float q = 0.f;
#pragma unroll
        for (int i = 0; i < kPerLane; i++) {
            const int j = lane + 32 * i;
            z.v[i] = j < H ? fmaf(x.v[i], c_g2[j], c_be2[j]) + fmaf(act, c_wa[j], c_ba[j]) : 0.f;
            if (j < H) q = fmaf(fmaxf(z.v[i], 0.f), c_wq[j], q);
        }
        q = warp_sum(q) + A.c_bq[0];
        const float dq = 2.0f * (q - y) / (float)A.B;                    // d mse_loss(target, q) / dq
        Row dz, dx;
#pragma unroll
        for (int i = 0; i < kPerLane; i++) { const int j = lane + 32 * i; dz.v[i] = (j < H && z.v[i] > 0.f) ? dq * c_wq[j] : 0.f; }
        ln_backward_row(dz, x, c_g2, rstd, H, lane, dx);
#pragma unroll
        for (int i = 0; i < kPerLane; i++) {
            const int j = lane + 32 * i;
            if (j < H) {
                const size_t o = (size_t)b * H + j;
                A.dh2[o] = dx.v[i]; A.sc0[o] = dz.v[i]; A.sc1[o] = dz.v[i] * x.v[i]; A.sc2[o] = dq * fmaxf(z.v[i], 0.f);
            }
        }
        if (lane == 0) { A.dv[b] = dq; if (A.q_out) A.q_out[b] = q; if (A.y_out) A.y_out[b] = y; }
    }
    if (!last_cta_done(A.ticket)) return;
    if (threadIdx.x < A.B) { s_dq[threadIdx.x] = A.dv[threadIdx.x]; s_act[threadIdx.x] = A.act[threadIdx.x]; }
    __syncthreads();
    // parameter gradients = column sums over the batch, in row order
    const int gt = threadIdx.x;
    for (int j = gt; j < H; j += kRowT) {
        float gba = 0.f, gwa = 0.f, gg2 = 0.f, gwq = 0.f;
#pragma unroll 32
        for (int bb = 0; bb < A.B; bb++) {
            const size_t o = (size_t)bb * H + j;
            const float dz = A.sc0[o];
            gba += dz; gwa = fmaf(dz, s_act[bb], gwa); gg2 += A.sc1[o]; gwq += A.sc2[o];
        }
        A.g_be2[j] = gba; A.g_g2[j] = gg2; A.g_t0[j] = gwa; A.g_t1[j] = gba; A.g_t2[j] = gwq;
    }
    if (gt == 0) { float sum = 0.f; for (int bb = 0; bb < A.B; bb++) sum += s_dq[bb]; A.g_t3[0] = sum; }
}

// K9: DDPG_agent.py:99-103  actor_loss = -mean(critic(states, actor(states)))
__global__ void __launch_bounds__(kRowT) learn_actor_head_kernel(HeadArgs A) {
    chain_enter();
    __shared__ float s_dp[kMaxB];
    __shared__ float sp[8][kMaxH];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, H = A.H2;
    const int b = (int)blockIdx.x * (kRowT / 32) + warp;
    {
        const float *const src[8] = {A.a_g2, A.a_be2, A.a_w3, A.c_g2, A.c_be2, A.c_wa, A.c_ba, A.c_wq};
        stage_vectors<8>(sp, src, H);
    }
    const float *a_g2 = sp[0], *a_be2 = sp[1], *a_w3 = sp[2], *c_g2 = sp[3], *c_be2 = sp[4], *c_wa = sp[5], *c_ba = sp[6], *c_wq = sp[7];
    __syncthreads();
    if (b < A.B) {
        Row x, o2, c;
        load_row(x, A.h2[JOB_A] + (size_t)b * H, H, lane);
        load_row(c, A.h2[JOB_C] + (size_t)b * H, H, lane);
        const float rstd = normalize_row(x, H, lane);
        float p = 0.f;
#pragma unroll
        for (int i = 0; i < kPerLane; i++) {
            const int j = lane + 32 * i;
            o2.v[i] = j < H ? fmaf(x.v[i], a_g2[j], a_be2[j]) : 0.f;
            if (j < H) p = fmaf(fmaxf(o2.v[i], 0.f), a_w3[j], p);
        }
        const float a = tanhf(warp_sum(p) + A.a_b3[0]);
        // dQ/da through the UPDATED critic: z = LN2(h2') + action_value(a); dq = -1 / B
        normalize_row(c, H, lane);
        float da = 0.f;
        const float dq = -1.0f / (float)A.B;
#pragma unroll
        for (int i = 0; i < kPerLane; i++) {
            const int j = lane + 32 * i;
            if (j < H) {
                const float z = fmaf(c.v[i], c_g2[j], c_be2[j]) + fmaf(a, c_wa[j], c_ba[j]);
                if (z > 0.f) da = fmaf(dq * c_wq[j], c_wa[j], da);
            }
        }
        da = warp_sum(da);
        const float dp = da * (1.0f - a * a);                             // through tanh
        Row dout, dx;
#pragma unroll
        for (int i = 0; i < kPerLane; i++) { const int j = lane + 32 * i; dout.v[i] = (j < H && o2.v[i] > 0.f) ? dp * a_w3[j] : 0.f; }
        ln_backward_row(dout, x, a_g2, rstd, H, lane, dx);
#pragma unroll
        for (int i = 0; i < kPerLane; i++) {
            const int j = lane + 32 * i;
            if (j < H) {
                const size_t o = (size_t)b * H + j;
                A.dh2[o] = dx.v[i]; A.sc0[o] = dout.v[i]; A.sc1[o] = dout.v[i] * x.v[i]; A.sc2[o] = dp * fmaxf(o2.v[i], 0.f);
            }
        }
        if (lane == 0) { A.dv[b] = dp; if (A.a_out) A.a_out[b] = a; }
    }
    if (!last_cta_done(A.ticket)) return;
    if (threadIdx.x < A.B) s_dp[threadIdx.x] = A.dv[threadIdx.x];
    __syncthreads();
    const int gt = threadIdx.x;
    for (int j = gt; j < H; j += kRowT) {
        float gbe = 0.f, gg = 0.f, gw3 = 0.f;
#pragma unroll 32
        for (int bb = 0; bb < A.B; bb++) { const size_t o = (size_t)bb * H + j; gbe += A.sc0[o]; gg += A.sc1[o]; gw3 += A.sc2[o]; }
        A.g_be2[j] = gbe; A.g_g2[j] = gg; A.g_t0[j] = gw3;
    }
    if (gt == 0) { float sum = 0.f; for (int bb = 0; bb < A.B; bb++) sum += s_dp[bb]; A.g_t1[0] = sum; }
}

// K5 / K11: relu + LayerNorm 1 backward (warp = row) -> dh1, and the LayerNorm parameter gradients dg1, dbe1 (last CTA).  The
// fc1 weight / bias gradients are a dW job of the next grouped launch (dW1 = dh1^T x, db1 = column sums of dh1).
struct L1Args {
    int B, H1;
    const float *h1, *da1;           // fc1 output (pre-LayerNorm) and the gradient w.r.t. relu(LN1(h1)), [B][H1]
    const float *g1, *be1;
    float *dh1;                      // out [B][H1]
    float *sc0, *sc1;                // scratch [B][H1]
    float *g_g1, *g_be1;
    int *ticket;
};
__global__ void __launch_bounds__(kRowT) learn_l1_backward_kernel(L1Args A) {
    chain_enter();
    __shared__ float sp[2][kMaxH];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, H = A.H1;
    {
        const float *const src[2] = {A.g1, A.be1};
        stage_vectors<2>(sp, src, H);
    }
    const float *g1 = sp[0], *be1 = sp[1];
    __syncthreads();
    const int b = (int)blockIdx.x * (kRowT / 32) + warp;
    if (b < A.B) {
        Row x, dout, dx;
        load_row(x, A.h1 + (size_t)b * H, H, lane);
        load_row(dout, A.da1 + (size_t)b * H, H, lane);
        const float rstd = normalize_row(x, H, lane);
#pragma unroll
        for (int i = 0; i < kPerLane; i++) {
            const int j = lane + 32 * i;
            if (!(j < H && fmaf(x.v[i], g1[j], be1[j]) > 0.f)) dout.v[i] = 0.f;      // relu mask
        }
        ln_backward_row(dout, x, g1, rstd, H, lane, dx);
#pragma unroll
        for (int i = 0; i < kPerLane; i++) {
            const int j = lane + 32 * i;
            if (j < H) { const size_t o = (size_t)b * H + j; A.dh1[o] = dx.v[i]; A.sc0[o] = dout.v[i]; A.sc1[o] = dout.v[i] * x.v[i]; }
        }
    }
    if (!last_cta_done(A.ticket)) return;
    for (int j = threadIdx.x; j < H; j += kRowT) {
        float gbe = 0.f, gg = 0.f;
#pragma unroll 32
        for (int bb = 0; bb < A.B; bb++) { const size_t o = (size_t)bb * H + j; gbe += A.sc0[o]; gg += A.sc1[o]; }
        A.g_be1[j] = gbe; A.g_g1[j] = gg;
    }
}

// K6 / K12: torch.optim.Adam step (single-tensor form: weight decay added to the gradient, lerp first moment) on every
// parameter of one network + the soft update of its target (DDPG_agent.py:108-131: tau p + (1 - tau) target)
struct AdamArgs {
    float *p, *m, *v, *target;
    const float *g;
    int n;
    float lr, wd, tau, omt;
    const int *step;                  // device counter of updates, already advanced by the gather kernel of this update
};
__global__ void __launch_bounds__(256) learn_adam_kernel(AdamArgs A) {
    chain_enter();
    __shared__ float s_step_size, s_bc2_sqrt;
    const int t = *A.step;
    if (threadIdx.x == 0) {
        const double bc1 = 1.0 - pow(0.9, (double)t), bc2 = 1.0 - pow(0.999, (double)t);
        s_step_size = (float)((double)A.lr / bc1); s_bc2_sqrt = (float)sqrt(bc2);
    }
    __syncthreads();
    const float step_size = s_step_size, bc2_sqrt = s_bc2_sqrt;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < A.n; i += gridDim.x * blockDim.x) {
        float p = A.p[i];
        const float g = A.wd != 0.f ? __fmaf_rn(A.wd, p, A.g[i]) : A.g[i];
        float m = A.m[i], v = A.v[i];
        m = __fmaf_rn(0.1f, g - m, m);                                   // exp_avg.lerp_(grad, 1 - beta1)
        v = __fmaf_rn(0.001f * g, g, __fmul_rn(v, 0.999f));              // exp_avg_sq.mul_(beta2).addcmul_(grad, grad, value = 1 - beta2)
        const float denom = __fadd_rn(__fdiv_rn(__fsqrt_rn(v), bc2_sqrt), 1e-8f);
        p = __fmaf_rn(-step_size, __fdiv_rn(m, denom), p);               // param.addcdiv_(exp_avg, denom, value = -step_size)
        A.p[i] = p; A.m[i] = m; A.v[i] = v;
        A.target[i] = __fadd_rn(__fmul_rn(A.tau, p), __fmul_rn(A.omt, A.target[i]));
    }
}

}  // namespace

struct tt_learner {
    Layout L;
    int B;
    float alpha, beta, gamma, tau, wd;
    uint64_t seed;
    int np[4];                        // parameter count per network
    float *p[4];                      // flat parameters
    float *m[2], *v[2], *g[2];        // Adam moments / gradients: [0] actor, [1] critic
    Batch bt;
    float *h1[NJOBS], *a1[NJOBS], *h2[NJOBS];
    float *dh2, *da1, *dh1, *sc0, *sc1, *sc2;
    float *q, *y, *aout, *dv;
    int *step;                        // number of updates done (Adam's step count and the sampling counter)
};

namespace {

size_t learner_layout(const Layout &L, int B, tt_learner *ln, char *base) {
    size_t off = 0;
    auto take = [&](size_t bytes) { size_t o = off; off = tt::align_up(off + bytes, 256); return o; };
    const int hm = L.h1 > L.h2 ? L.h1 : L.h2;
    const size_t pa = sizeof(float) * L.actor_count(), pc = sizeof(float) * L.critic_count();
    size_t o_p[4] = {take(pa), take(pa), take(pc), take(pc)};
    size_t o_m[2] = {take(pa), take(pc)}, o_v[2] = {take(pa), take(pc)}, o_g[2] = {take(pa), take(pc)};
    size_t o_s = take(sizeof(float) * B * L.in), o_s2 = take(sizeof(float) * B * L.in), o_a = take(sizeof(float) * B),
           o_r = take(sizeof(float) * B), o_d = take(sizeof(float) * B), o_rows = take(sizeof(int64_t) * B);
    size_t o_h1[NJOBS], o_h2[NJOBS], o_a1[NJOBS];
    for (int j = 0; j < NJOBS; j++) { o_h1[j] = take(sizeof(float) * B * L.h1); o_h2[j] = take(sizeof(float) * B * L.h2); o_a1[j] = take(sizeof(float) * B * L.h1); }
    size_t o_dh2 = take(sizeof(float) * B * L.h2), o_da1 = take(sizeof(float) * B * L.h1), o_dh1 = take(sizeof(float) * B * L.h1);
    size_t o_sc[3] = {take(sizeof(float) * B * hm), take(sizeof(float) * B * hm), take(sizeof(float) * B * hm)};
    size_t o_q = take(sizeof(float) * B), o_y = take(sizeof(float) * B), o_ao = take(sizeof(float) * B), o_dv = take(sizeof(float) * B), o_step = take(256);
    if (ln) {
        auto f = [&](size_t o) { return reinterpret_cast<float *>(base + o); };
        for (int i = 0; i < 4; i++) ln->p[i] = f(o_p[i]);
        for (int i = 0; i < 2; i++) { ln->m[i] = f(o_m[i]); ln->v[i] = f(o_v[i]); ln->g[i] = f(o_g[i]); }
        ln->bt.s = f(o_s); ln->bt.s2 = f(o_s2); ln->bt.a = f(o_a); ln->bt.r = f(o_r); ln->bt.d = f(o_d);
        ln->bt.rows = reinterpret_cast<int64_t *>(base + o_rows);
        for (int j = 0; j < NJOBS; j++) { ln->h1[j] = f(o_h1[j]); ln->h2[j] = f(o_h2[j]); ln->a1[j] = f(o_a1[j]); }
        ln->dh2 = f(o_dh2); ln->da1 = f(o_da1); ln->dh1 = f(o_dh1); ln->sc0 = f(o_sc[0]); ln->sc1 = f(o_sc[1]); ln->sc2 = f(o_sc[2]);
        ln->q = f(o_q); ln->y = f(o_y); ln->aout = f(o_ao); ln->dv = f(o_dv); ln->step = reinterpret_cast<int *>(base + o_step);
    }
    return off;
}

Job fwd_job(int B, int N, int K, const float *X, const float *W, const float *bias, float *Y) {
    Job j{};
    j.type = G_FWD; j.ctas = (N + 15) / 16; j.B = B; j.N = N; j.K = K; j.X = X; j.ldx = K; j.W = W; j.ldw = K; j.bias = bias; j.Y = Y; j.ldy = N;
    return j;
}

// a row-wise stage: ceil(B / 8) CTAs (one warp per batch row), the last one to finish sums the columns
template <typename Kern, typename Args>
int launch_rows(Kern kern, const Args &args, int B, cudaStream_t s) {
    TT_CUDA(tt::launch_chained(true, kern, dim3((B + kRowT / 32 - 1) / (kRowT / 32)), dim3(kRowT), 0, s, args));
    TT_COUNT_LAUNCH(); TT_LAUNCH_CHECK();
    return TT_OK;
}

int launch_jobs(const JobList &L, cudaStream_t s) {
    int ctas = 0;
    for (int i = 0; i < L.n; i++) ctas += L.j[i].ctas;
    TT_CUDA(tt::launch_chained(true, learn_gemm_kernel, dim3(ctas), dim3(kT), 0, s, L));
    TT_COUNT_LAUNCH(); TT_LAUNCH_CHECK();
    return TT_OK;
}

}  // namespace

extern "C" {

size_t tt_learner_workspace_bytes(int32_t in_dim, int32_t h1, int32_t h2, int32_t batch) {
    if (in_dim <= 0 || h1 <= 0 || h2 <= 0 || batch <= 0) return 0;
    const Layout L = {in_dim, h1, h2};
    return learner_layout(L, batch, nullptr, nullptr);
}

int tt_learner_create(tt_learner **out, int32_t in_dim, int32_t h1, int32_t h2, int32_t batch, float alpha, float beta, float gamma,
                      float tau, float critic_weight_decay, uint64_t seed, void *d_workspace, size_t workspace_bytes) {
    TT_REQUIRE(out && d_workspace, "NULL argument");
    TT_REQUIRE(in_dim == TT_OBS_DIM, "the replay ring holds 23-float observation rows: in_dim must be 23");
    TT_REQUIRE(h1 > 0 && h1 <= kMaxH && h2 > 0 && h2 <= kMaxH, "hidden sizes must be in 1..512");
    TT_REQUIRE(batch > 0 && batch <= kMaxB, "batch must be in 1..64");
    if ((reinterpret_cast<uintptr_t>(d_workspace) & 255) != 0 || workspace_bytes < tt_learner_workspace_bytes(in_dim, h1, h2, batch)) {
        tt::set_error("tt_learner_create: workspace must be 256 B aligned and >= %zu bytes", tt_learner_workspace_bytes(in_dim, h1, h2, batch));
        return TT_ERR_WORKSPACE;
    }
    if (tt_device_count() <= 0) { tt::set_error("tt_learner_create: no CUDA device (there is no CPU fallback)"); return TT_ERR_CUDA; }
    tt_learner *ln = new (std::nothrow) tt_learner;
    TT_REQUIRE(ln, "out of host memory");
    ln->L = Layout{in_dim, h1, h2}; ln->B = batch;
    ln->alpha = alpha; ln->beta = beta; ln->gamma = gamma; ln->tau = tau; ln->wd = critic_weight_decay; ln->seed = seed;
    ln->np[0] = ln->np[1] = ln->L.actor_count(); ln->np[2] = ln->np[3] = ln->L.critic_count();
    learner_layout(ln->L, batch, ln, static_cast<char *>(d_workspace));
    const cudaError_t err = cudaMemset(d_workspace, 0, tt_learner_workspace_bytes(in_dim, h1, h2, batch));
    if (err != cudaSuccess) { delete ln; return tt::cuda_fail(err, "cudaMemset(workspace)"); }
    *out = ln;
    return TT_OK;
}

int tt_learner_destroy(tt_learner *ln) { delete ln; return TT_OK; }

float *tt_learner_params(tt_learner *ln, int32_t net) { return ln && net >= 0 && net < 4 ? ln->p[net] : nullptr; }
int64_t tt_learner_param_count(tt_learner *ln, int32_t net) { return ln && net >= 0 && net < 4 ? ln->np[net] : 0; }
float *tt_learner_grads(tt_learner *ln, int32_t which) { return ln && which >= 0 && which < 2 ? ln->g[which] : nullptr; }
const float *tt_learner_last_q(tt_learner *ln) { return ln ? ln->q : nullptr; }
const int64_t *tt_learner_last_rows(tt_learner *ln) { return ln ? ln->bt.rows : nullptr; }

int tt_learner_reset_optimizer(tt_learner *ln, tt_stream_t stream) {
    TT_REQUIRE(ln, "learner is NULL");
    cudaStream_t s = tt::as_stream(stream);
    for (int i = 0; i < 2; i++) {
        TT_CUDA(cudaMemsetAsync(ln->m[i], 0, sizeof(float) * ln->np[2 * i], s));
        TT_CUDA(cudaMemsetAsync(ln->v[i], 0, sizeof(float) * ln->np[2 * i], s));
    }
    TT_CUDA(cudaMemsetAsync(ln->step, 0, 256, s));
    return TT_OK;
}

int tt_learn_step(tt_learner *ln, const tt_replay_ring *ring, const int64_t *d_rows, tt_actor *repack_into, tt_stream_t stream) {
    TT_REQUIRE(ln && ring, "NULL argument");
    const int64_t max_mem = ring->mem_cntr < ring->mem_size ? ring->mem_cntr : ring->mem_size;
    return tt_learn_step_window(ln, ring, d_rows, repack_into, 0, max_mem, stream);
}

int tt_learn_step_window(tt_learner *ln, const tt_replay_ring *ring, const int64_t *d_rows, tt_actor *repack_into, int64_t window_begin,
                         int64_t window_count, tt_stream_t stream) {
    TT_REQUIRE(ln && ring, "NULL argument");
    TT_REQUIRE(ring->d_state_mem && ring->d_action_mem && ring->d_reward_mem && ring->d_new_state_mem && ring->d_terminal_mem &&
               ring->mem_size > 0, "bad ring");
    TT_REQUIRE(window_begin >= 0 && window_begin < ring->mem_size && window_count >= 0 && window_count <= ring->mem_size, "bad sampling window");
    const int64_t win_begin = window_begin, max_mem = window_count;
    TT_REQUIRE(max_mem >= ln->B || d_rows, "fewer transitions in the sampling window than the batch size (DDPG_agent.py:73-74)");
    cudaStream_t s = tt::as_stream(stream);
    const Layout &L = ln->L;
    const int B = ln->B, IN = L.in, H1 = L.h1, H2 = L.h2;
    float *pa = ln->p[NET_ACTOR], *pta = ln->p[NET_TARGET_ACTOR], *pc = ln->p[NET_CRITIC], *ptc = ln->p[NET_TARGET_CRITIC];
    float *ga = ln->g[0], *gc = ln->g[1];
    const int T = L.tail();

    // K0
    TT_CUDA(tt::launch_chained(true, learn_gather_kernel, dim3(1), dim3(1024), 0, s, *ring, win_begin, max_mem, d_rows, ln->bt, B, IN, ln->seed, ln->step));
    TT_COUNT_LAUNCH(); TT_LAUNCH_CHECK();
    int rc;
    // K1: fc1 + LayerNorm 1 + ReLU of the four forward passes
    const float *netp[NJOBS] = {pta, ptc, pc, pa};
    const float *netx[NJOBS] = {ln->bt.s2, ln->bt.s2, ln->bt.s, ln->bt.s};
    auto fc1 = [&](const int *jobs, int njobs) -> int {
        Fc1Args A{};
        A.njobs = njobs; A.B = B; A.IN = IN; A.H1 = H1;
        for (int i = 0; i < njobs; i++) {
            const int j = jobs[i];
            A.j[i] = Fc1Job{netx[j], netp[j] + L.w1(), netp[j] + L.b1(), netp[j] + L.g1(), netp[j] + L.be1(), ln->h1[j], ln->a1[j]};
        }
        const size_t dsm = sizeof(float) * (size_t)H1 * kW1P;
        static bool attr_of[tt::kMaxDevices] = {};
        if (dsm > 48 * 1024 && !attr_of[tt::device_index()]) {
            TT_CUDA(cudaFuncSetAttribute(learn_fc1_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(sizeof(float) * kMaxH * kW1P)));
            attr_of[tt::device_index()] = true;
        }
        TT_CUDA(tt::launch_chained(true, learn_fc1_kernel, dim3(njobs * ((B + 7) / 8)), dim3(256), dsm, s, A));
        TT_COUNT_LAUNCH(); TT_LAUNCH_CHECK();
        return TT_OK;
    };
    {
        const int all[NJOBS] = {JOB_TA, JOB_TC, JOB_C, JOB_A};
        if ((rc = fc1(all, NJOBS)) != TT_OK) return rc;
    }
    // K2: fc2
    {
        JobList J{}; J.n = NJOBS;
        for (int j = 0; j < NJOBS; j++) J.j[j] = fwd_job(B, H2, H1, ln->a1[j], netp[j] + L.w2(), netp[j] + L.b2(), ln->h2[j]);
        if ((rc = launch_jobs(J, s)) != TT_OK) return rc;
    }
    // K3: critic head
    HeadArgs H{};
    H.B = B; H.H1 = H1; H.H2 = H2; H.gamma = ln->gamma;
    for (int j = 0; j < NJOBS; j++) H.h2[j] = ln->h2[j];
    H.ta_g2 = pta + L.g2(); H.ta_be2 = pta + L.be2(); H.ta_w3 = pta + T; H.ta_b3 = pta + T + H2;
    H.tc_g2 = ptc + L.g2(); H.tc_be2 = ptc + L.be2(); H.tc_wa = ptc + T; H.tc_ba = ptc + T + H2; H.tc_wq = ptc + T + 2 * H2; H.tc_bq = ptc + T + 3 * H2;
    H.c_g2 = pc + L.g2(); H.c_be2 = pc + L.be2(); H.c_wa = pc + T; H.c_ba = pc + T + H2; H.c_wq = pc + T + 2 * H2; H.c_bq = pc + T + 3 * H2;
    H.a_g2 = pa + L.g2(); H.a_be2 = pa + L.be2(); H.a_w3 = pa + T; H.a_b3 = pa + T + H2;
    H.act = ln->bt.a; H.rew = ln->bt.r; H.done = ln->bt.d;
    H.dh2 = ln->dh2; H.sc0 = ln->sc0; H.sc1 = ln->sc1; H.sc2 = ln->sc2;
    H.q_out = ln->q; H.y_out = ln->y; H.a_out = ln->aout; H.dv = ln->dv; H.ticket = ln->step + 1;
    {
        HeadArgs C = H;
        C.g_g2 = gc + L.g2(); C.g_be2 = gc + L.be2(); C.g_t0 = gc + T; C.g_t1 = gc + T + H2; C.g_t2 = gc + T + 2 * H2; C.g_t3 = gc + T + 3 * H2;
        if ((rc = launch_rows(learn_critic_head_kernel, C, B, s)) != TT_OK) return rc;
    }
    // backward through fc2 / LayerNorm 1 / fc1 of one network whose dh2 is in ln->dh2 and whose forward job is `job`
    auto trunk_backward = [&](int job, const float *p, float *g, const float *x) -> int {
        JobList J{}; J.n = 2;
        Job w{};
        w.type = G_WGRAD; w.B = B; w.N = H2; w.K = H1; w.ctas = ((H2 + 15) / 16) * ((H1 + 31) / 32);
        w.X = ln->a1[job]; w.ldx = H1;
        w.D = ln->dh2; w.ldd = H2; w.Y = g + L.w2(); w.ldy = H1; w.db = g + L.b2();
        Job xg{};
        xg.type = G_XGRAD; xg.B = B; xg.N = H2; xg.K = H1; xg.ctas = ((B + 31) / 32) * ((H1 + 31) / 32);
        xg.D = ln->dh2; xg.ldd = H2; xg.W = p + L.w2(); xg.ldw = H1; xg.Y = ln->da1; xg.ldy = H1;
        J.j[0] = w; J.j[1] = xg;
        int r = launch_jobs(J, s);
        if (r != TT_OK) return r;
        L1Args A{};
        A.B = B; A.H1 = H1; A.h1 = ln->h1[job]; A.da1 = ln->da1; A.g1 = p + L.g1(); A.be1 = p + L.be1();
        A.dh1 = ln->dh1; A.sc0 = ln->sc0; A.sc1 = ln->sc1; A.g_g1 = g + L.g1(); A.g_be1 = g + L.be1(); A.ticket = ln->step + 1;
        if ((r = launch_rows(learn_l1_backward_kernel, A, B, s)) != TT_OK) return r;
        JobList J1{}; J1.n = 1;
        Job w1{};
        w1.type = G_WGRAD; w1.B = B; w1.N = H1; w1.K = IN; w1.ctas = ((H1 + 15) / 16) * ((IN + 31) / 32);
        w1.X = x; w1.ldx = IN; w1.D = ln->dh1; w1.ldd = H1; w1.Y = g + L.w1(); w1.ldy = IN; w1.db = g + L.b1();
        J1.j[0] = w1;
        return launch_jobs(J1, s);
    };
    auto adam = [&](float *p, float *m, float *v, float *target, const float *g, int n, float lr, float wd) -> int {
        AdamArgs A{};
        A.p = p; A.m = m; A.v = v; A.target = target; A.g = g; A.n = n; A.lr = lr; A.wd = wd; A.tau = ln->tau; A.omt = (float)(1.0 - (double)ln->tau);
        A.step = ln->step;
        int blocks = (n + 255) / 256;
        const int cap = tt::sm_count() * 2;
        if (blocks > cap) blocks = cap;
        TT_CUDA(tt::launch_chained(true, learn_adam_kernel, dim3(blocks), dim3(256), 0, s, A));
        TT_COUNT_LAUNCH(); TT_LAUNCH_CHECK();
        return TT_OK;
    };
    // K4, K5, K6: critic backward + optimizer (DDPG_agent.py:95-98)
    if ((rc = trunk_backward(JOB_C, pc, gc, ln->bt.s)) != TT_OK) return rc;
    if ((rc = adam(pc, ln->m[1], ln->v[1], ptc, gc, ln->np[NET_CRITIC], ln->beta, ln->wd)) != TT_OK) return rc;
    // K7, K8: the updated critic's trunk on s (into the JOB_C buffers)
    {
        const int one[1] = {JOB_C};
        if ((rc = fc1(one, 1)) != TT_OK) return rc;
        JobList J{}; J.n = 1;
        J.j[0] = fwd_job(B, H2, H1, ln->a1[JOB_C], pc + L.w2(), pc + L.b2(), ln->h2[JOB_C]);
        if ((rc = launch_jobs(J, s)) != TT_OK) return rc;
    }
    // K9: actor head
    {
        HeadArgs A = H;
        A.g_g2 = ga + L.g2(); A.g_be2 = ga + L.be2(); A.g_t0 = ga + T; A.g_t1 = ga + T + H2; A.g_t2 = nullptr; A.g_t3 = nullptr;
        if ((rc = launch_rows(learn_actor_head_kernel, A, B, s)) != TT_OK) return rc;
    }
    // K10, K11, K12 (DDPG_agent.py:99-106)
    if ((rc = trunk_backward(JOB_A, pa, ga, ln->bt.s)) != TT_OK) return rc;
    if ((rc = adam(pa, ln->m[0], ln->v[0], pta, ga, ln->np[NET_ACTOR], ln->alpha, 0.f)) != TT_OK) return rc;
    // hand the new policy to the rollout actor (re-pack into the fp32 / tensor-core operand images)
    if (repack_into) {
        const float *a = pa;
        return tt_actor_load(repack_into, a + L.w1(), a + L.b1(), a + L.g1(), a + L.be1(), a + L.w2(), a + L.b2(), a + L.g2(), a + L.be2(),
                             a + T, a + T + H2, stream);
    }
    return TT_OK;
}

}  // extern "C"
