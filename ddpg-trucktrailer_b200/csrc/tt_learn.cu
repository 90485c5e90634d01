// tt_learn.cu -- row f1: one DDPG update (Agent.learn, DDPG/DDPG_agent.py:72-131; CriticNetwork DDPG/networks.py:9-68,
// ActorNetwork :98-147; ReplayBuffer.sample_buffer DDPG/replay_buffer.py:23-34) as 13 small hand-written kernels on the
// caller's stream, on the device-resident replay ring, followed by the re-pack of the new policy into the rollout actor's
// operand images.  The whole sequence is capturable into one CUDA graph (nothing synchronises, the Adam step counter and
// the sampling counter live in device memory).
//
// The step is latency-bound (batch 64 x 23-400-300 networks = 150 MFLOP): what matters is the length of the dependency
// chain, not the arithmetic.  Stages (one launch each):
//   K0  sample 64 ring rows (Philox) and gather s, a, r, s', done
//   K1  fc1 of target_actor(s'), target_critic(s'), critic(s), actor(s)              -- one grouped launch, 4 jobs
//   K2  LayerNorm 1 + ReLU (in the GEMM prologue) + fc2 of the same four
//   K3  critic head (one CTA, warp = batch row): a' = target_actor head, y = r + gamma Q'(s', a') (1 - done), q = Q(s, a),
//       dL/dq = 2 (q - y) / B, back through q / action_value / LayerNorm 2 -> d h2; column sums = their parameter gradients
//   K4  critic fc2 backward: dW2 = dh2^T a1 (grouped with) da1 = dh2 W2
//   K5  critic LayerNorm 1 backward (one CTA, warp = row) + fc1 backward dW1 = dh1^T s
//   K6  Adam (weight decay 0.01) on the critic + soft update of target_critic
//   K7  fc1, K8 fc2 of the UPDATED critic on s
//   K9  actor head (one CTA): a = actor head, dL/da = -(1/B) dQ/da through relu / action_value, back through tanh / mu /
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

// ---- grouped GEMM jobs ----
enum { G_FWD = 0, G_WGRAD = 1, G_XGRAD = 2 };
struct Job {
    int type, ctas;                   // CTAs this job occupies in the launch
    int B, N, K;
    // G_FWD   : Y[b][n] = bias[n] + sum_k f(X)[b][k] W[n][k];  f = identity, or relu(LayerNorm(X; g, be)) when g != NULL
    //           (row statistics computed in the prologue, written to `stats` [B][2] by the job's first CTA)
    // G_WGRAD : dW[n][k] = sum_b D[b][n] f(X)[b][k], db[n] = sum_b D[b][n];  f = identity, or relu(LayerNorm) from `stats`
    // G_XGRAD : dX[b][k] = sum_n D[b][n] W[n][k]
    const float *X; int ldx;
    const float *W; int ldw;
    const float *bias;
    const float *g, *be;
    float *stats;
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

// mean / rstd of `rows` rows of width K <= 512 (two-pass, like torch); out[b] = (mean, rstd).  A warp takes FOUR rows at a
// time with all their values in registers: one round of independent loads (one memory latency per four rows instead of two
// per row -- the step is latency-bound), the second pass runs from registers.
__device__ void row_stats(const float *__restrict__ X, int ldx, int rows, int K, float2 *out) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
    for (int b0 = 4 * warp; b0 < rows; b0 += 4 * nw) {
        float v[4][kMaxH / 32];
#pragma unroll
        for (int r = 0; r < 4; r++)
#pragma unroll
            for (int i = 0; i < kMaxH / 32; i++) {
                const int k = lane + 32 * i;
                v[r][i] = (b0 + r < rows && k < K) ? X[(size_t)(b0 + r) * ldx + k] : 0.f;
            }
        float mean[4];
#pragma unroll
        for (int r = 0; r < 4; r++) {
            float sum = 0.f;
#pragma unroll
            for (int i = 0; i < kMaxH / 32; i++) sum += v[r][i];
            mean[r] = warp_sum(sum) / (float)K;
        }
#pragma unroll
        for (int r = 0; r < 4; r++) {
            float q = 0.f;
#pragma unroll
            for (int i = 0; i < kMaxH / 32; i++) { const float d = (lane + 32 * i < K) ? v[r][i] - mean[r] : 0.f; q = fmaf(d, d, q); }
            const float rstd = rsqrtf(warp_sum(q) / (float)K + kLnEps);
            if (lane == 0 && b0 + r < rows) out[b0 + r] = make_float2(mean[r], rstd);
        }
    }
}

__device__ __forceinline__ float ln_relu(float x, float2 st, float g, float be) { return fmaxf(fmaf((x - st.x) * st.y, g, be), 0.f); }

// Y tile: all B rows x 32 columns [n0, n0 + 32); reduction over K in chunks of 32.  A thread owns 4 rows x 2 columns (one
// 16 B and one 8 B shared-memory load per 8 FMAs: a 1 x 4 tile is shared-memory-bandwidth bound, 5 loads per 4 FMAs).  The next
// chunk's operands are fetched into registers while the current one is multiplied (the global-load latency of every chunk
// would otherwise be exposed).
constexpr int kXS = 68, kWS = 34;                                                // padded row strides (floats), 16 B / 8 B aligned
__device__ void job_fwd(const Job &J, int cta, float *smem) {
    float *Xs = smem;                                                           // [32 k][kXS]  (b)
    float *Ws = smem + 32 * kXS;                                                // [32 k][kWS]  (n)
    float2 *st = reinterpret_cast<float2 *>(smem + 32 * kXS + 32 * kWS);        // [64]
    const int tid = threadIdx.x, tn = tid & 15, tb = tid >> 4, n0 = cta * 32;
    const bool ln = J.g != nullptr;
    float xr[8], wr[4];
    auto fetch = [&](int k0) {                                                  // raw operands of chunk k0 -> registers
#pragma unroll
        for (int i = 0; i < 8; i++) {
            const int e = tid + kT * i, b = e >> 5, k = k0 + (e & 31);
            xr[i] = (b < J.B && k < J.K) ? J.X[(size_t)b * J.ldx + k] : 0.f;
        }
#pragma unroll
        for (int i = 0; i < 4; i++) {
            const int e = tid + kT * i, n = e >> 5, k = k0 + (e & 31);
            wr[i] = (n0 + n < J.N && k < J.K) ? J.W[(size_t)(n0 + n) * J.ldw + k] : 0.f;
        }
    };
    fetch(0);
    if (ln) {
        row_stats(J.X, J.ldx, J.B, J.K, st);
        __syncthreads();
        if (cta == 0 && J.stats && tid < J.B) reinterpret_cast<float2 *>(J.stats)[tid] = st[tid];
    }
    float acc[4][2] = {};
    for (int k0 = 0; k0 < J.K; k0 += 32) {
#pragma unroll
        for (int i = 0; i < 8; i++) {
            const int e = tid + kT * i, b = e >> 5, kk = e & 31, k = k0 + kk;
            float x = xr[i];
            if (ln) x = (b < J.B && k < J.K) ? ln_relu(x, st[b], J.g[k], J.be[k]) : 0.f;
            Xs[kk * kXS + b] = x;
        }
#pragma unroll
        for (int i = 0; i < 4; i++) { const int e = tid + kT * i; Ws[(e & 31) * kWS + (e >> 5)] = wr[i]; }
        __syncthreads();
        if (k0 + 32 < J.K) fetch(k0 + 32);
#pragma unroll
        for (int kk = 0; kk < 32; kk++) {
            const float4 x = *reinterpret_cast<const float4 *>(Xs + kk * kXS + 4 * tb);
            const float2 w = *reinterpret_cast<const float2 *>(Ws + kk * kWS + 2 * tn);
            acc[0][0] = fmaf(x.x, w.x, acc[0][0]); acc[0][1] = fmaf(x.x, w.y, acc[0][1]);
            acc[1][0] = fmaf(x.y, w.x, acc[1][0]); acc[1][1] = fmaf(x.y, w.y, acc[1][1]);
            acc[2][0] = fmaf(x.z, w.x, acc[2][0]); acc[2][1] = fmaf(x.z, w.y, acc[2][1]);
            acc[3][0] = fmaf(x.w, w.x, acc[3][0]); acc[3][1] = fmaf(x.w, w.y, acc[3][1]);
        }
        __syncthreads();
    }
#pragma unroll
    for (int j = 0; j < 2; j++) {
        const int n = n0 + 2 * tn + j;
        if (n < J.N) {
            const float bias = J.bias ? J.bias[n] : 0.f;
#pragma unroll
            for (int i = 0; i < 4; i++) {
                const int b = 4 * tb + i;
                if (b < J.B) J.Y[(size_t)b * J.ldy + n] = acc[i][j] + bias;
            }
        }
    }
}

// dW tile: 16 rows n x 64 columns k; reduction over the batch (B <= 64) in one pass
__device__ void job_wgrad(const Job &J, int cta, float *smem) {
    float (*Xs)[65] = reinterpret_cast<float (*)[65]>(smem);                    // [64 b][65]  (k)
    float (*Ds)[17] = reinterpret_cast<float (*)[17]>(smem + 64 * 65);          // [64 b][17]  (n)
    const int ktiles = (J.K + 63) / 64;
    const int nt = cta / ktiles, kt = cta - nt * ktiles, n0 = nt * 16, k0 = kt * 64;
    const int tid = threadIdx.x, tn = tid & 15, tk = tid >> 4;
    const bool ln = J.g != nullptr;
    const float2 *st = reinterpret_cast<const float2 *>(J.stats);
#pragma unroll
    for (int i = 0; i < 16; i++) {
        const int e = tid + kT * i, b = e >> 6, kk = e & 63, k = k0 + kk;
        float x = 0.f;
        if (b < J.B && k < J.K) {
            x = J.X[(size_t)b * J.ldx + k];
            if (ln) x = ln_relu(x, st[b], J.g[k], J.be[k]);
        }
        Xs[b][kk] = x;
    }
#pragma unroll
    for (int i = 0; i < 4; i++) {
        const int e = tid + kT * i, b = e >> 4, n = e & 15;
        Ds[b][n] = (b < J.B && n0 + n < J.N) ? J.D[(size_t)b * J.ldd + n0 + n] : 0.f;
    }
    __syncthreads();
    float acc[4] = {0.f, 0.f, 0.f, 0.f}, accb = 0.f;
    for (int b = 0; b < J.B; b++) {
        const float d = Ds[b][tn];
        accb += d;
#pragma unroll
        for (int i = 0; i < 4; i++) acc[i] = fmaf(d, Xs[b][tk + 16 * i], acc[i]);
    }
    if (n0 + tn < J.N) {
#pragma unroll
        for (int i = 0; i < 4; i++) {
            const int k = k0 + tk + 16 * i;
            if (k < J.K) J.Y[(size_t)(n0 + tn) * J.ldy + k] = acc[i];
        }
        if (J.db && kt == 0 && tk == 0) J.db[n0 + tn] = accb;
    }
}

// dX tile: all B rows x 32 columns [k0, k0 + 32); reduction over N in chunks of 32 (same register tile and prefetch as job_fwd)
__device__ void job_xgrad(const Job &J, int cta, float *smem) {
    float *Ds = smem;                                                           // [32 n][kXS]  (b)
    float *Ws = smem + 32 * kXS;                                                // [32 n][kWS]  (k)
    const int tid = threadIdx.x, tk = tid & 15, tb = tid >> 4, k0 = cta * 32;
    float dr[8], wr[4];
    auto fetch = [&](int nb) {
#pragma unroll
        for (int i = 0; i < 8; i++) {
            const int e = tid + kT * i, b = e >> 5, nn = e & 31;
            dr[i] = (b < J.B && nb + nn < J.N) ? J.D[(size_t)b * J.ldd + nb + nn] : 0.f;
        }
#pragma unroll
        for (int i = 0; i < 4; i++) {
            const int e = tid + kT * i, nn = e >> 5, kk = e & 31;
            wr[i] = (nb + nn < J.N && k0 + kk < J.K) ? J.W[(size_t)(nb + nn) * J.ldw + k0 + kk] : 0.f;
        }
    };
    fetch(0);
    float acc[4][2] = {};
    for (int nb = 0; nb < J.N; nb += 32) {
#pragma unroll
        for (int i = 0; i < 8; i++) { const int e = tid + kT * i; Ds[(e & 31) * kXS + (e >> 5)] = dr[i]; }
#pragma unroll
        for (int i = 0; i < 4; i++) { const int e = tid + kT * i; Ws[(e >> 5) * kWS + (e & 31)] = wr[i]; }
        __syncthreads();
        if (nb + 32 < J.N) fetch(nb + 32);
#pragma unroll
        for (int nn = 0; nn < 32; nn++) {
            const float4 x = *reinterpret_cast<const float4 *>(Ds + nn * kXS + 4 * tb);
            const float2 w = *reinterpret_cast<const float2 *>(Ws + nn * kWS + 2 * tk);
            acc[0][0] = fmaf(x.x, w.x, acc[0][0]); acc[0][1] = fmaf(x.x, w.y, acc[0][1]);
            acc[1][0] = fmaf(x.y, w.x, acc[1][0]); acc[1][1] = fmaf(x.y, w.y, acc[1][1]);
            acc[2][0] = fmaf(x.z, w.x, acc[2][0]); acc[2][1] = fmaf(x.z, w.y, acc[2][1]);
            acc[3][0] = fmaf(x.w, w.x, acc[3][0]); acc[3][1] = fmaf(x.w, w.y, acc[3][1]);
        }
        __syncthreads();
    }
#pragma unroll
    for (int j = 0; j < 2; j++) {
        const int k = k0 + 2 * tk + j;
        if (k < J.K) {
#pragma unroll
            for (int i = 0; i < 4; i++) {
                const int b = 4 * tb + i;
                if (b < J.B) J.Y[(size_t)b * J.ldy + k] = acc[i][j];
            }
        }
    }
}

__global__ void __launch_bounds__(kT) learn_gemm_kernel(JobList L) {
    __shared__ __align__(16) float smem[64 * 65 + 64 * 17];       // >= 32 * kXS + 32 * kWS + 128 (forward / dX) and the dW layout
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
__global__ void learn_gather_kernel(tt_replay_ring ring, int64_t max_mem, const int64_t *__restrict__ given_rows, Batch bt, int B, int in,
                                    uint64_t seed, int *__restrict__ step) {
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

// ---- row-wise stages: a CLUSTER of 8 CTAs x 8 warps, one warp per batch row with the row in registers; after a cluster
//      barrier the same 2 048 threads take the column sums over the batch (= the parameter gradients), one column each, rows
//      in order (deterministic).  One warp per row keeps the dependent chain of a row (load -> statistics -> head -> backward)
//      at one memory latency per network instead of serialising rows.
constexpr int kRowCtas = 8, kRowT = 256, kRowThreads = kRowCtas * kRowT, kPerLane = kMaxH / 32;
struct Row { float v[kPerLane]; };

__device__ __forceinline__ uint32_t cluster_rank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ void cluster_barrier() {        // release / acquire at cluster scope: the rows' global writes are visible
    __threadfence();
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
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
    // gradients (flat-layout pointers)
    float *g_g2, *g_be2, *g_t0, *g_t1, *g_t2, *g_t3;                      // critic: wa, ba, wq, bq | actor: w3, b3, -, -
    float *q_out, *y_out, *a_out;                                         // diagnostics / hand-over: Q(s,a), target, actor(s)
};

// K3: DDPG_agent.py:84-97
__global__ void __launch_bounds__(kRowT) learn_critic_head_kernel(HeadArgs A) {
    __shared__ float s_dq[kMaxB], s_act[kMaxB];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, H = A.H2;
    const int b = (int)cluster_rank() * (kRowT / 32) + warp;
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
        for (int i = 0; i < kPerLane; i++) { const int j = lane + 32 * i; if (j < H) p = fmaf(fmaxf(fmaf(xt.v[i], A.ta_g2[j], A.ta_be2[j]), 0.f), A.ta_w3[j], p); }
        const float a2 = tanhf(warp_sum(p) + A.ta_b3[0]);
        // Q'(s', a') = q(relu(LN2(h2) + action_value(a')))                 (networks.py:53-68)
        normalize_row(xc, H, lane);
        float q2 = 0.f;
#pragma unroll
        for (int i = 0; i < kPerLane; i++) {
            const int j = lane + 32 * i;
            if (j < H) q2 = fmaf(fmaxf(fmaf(xc.v[i], A.tc_g2[j], A.tc_be2[j]) + fmaf(a2, A.tc_wa[j], A.tc_ba[j]), 0.f), A.tc_wq[j], q2);
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
            z.v[i] = j < H ? fmaf(x.v[i], A.c_g2[j], A.c_be2[j]) + fmaf(act, A.c_wa[j], A.c_ba[j]) : 0.f;
            if (j < H) q = fmaf(fmaxf(z.v[i], 0.f), A.c_wq[j], q);
        }
        q = warp_sum(q) + A.c_bq[0];
        const float dq = 2.0f * (q - y) / (float)A.B;                    // d mse_loss(target, q) / dq
        Row dz, dx;
#pragma unroll
        for (int i = 0; i < kPerLane; i++) { const int j = lane + 32 * i; dz.v[i] = (j < H && z.v[i] > 0.f) ? dq * A.c_wq[j] : 0.f; }
        ln_backward_row(dz, x, A.c_g2, rstd, H, lane, dx);
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
    cluster_barrier();
    if (threadIdx.x < A.B) { s_dq[threadIdx.x] = A.dv[threadIdx.x]; s_act[threadIdx.x] = A.act[threadIdx.x]; }
    __syncthreads();
    // parameter gradients = column sums over the batch, in row order
    const int gt = (int)cluster_rank() * kRowT + threadIdx.x;
    for (int j = gt; j < H; j += kRowThreads) {
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
    __shared__ float s_dp[kMaxB];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, H = A.H2;
    const int b = (int)cluster_rank() * (kRowT / 32) + warp;
    if (b < A.B) {
        Row x, o2, c;
        load_row(x, A.h2[JOB_A] + (size_t)b * H, H, lane);
        load_row(c, A.h2[JOB_C] + (size_t)b * H, H, lane);
        const float rstd = normalize_row(x, H, lane);
        float p = 0.f;
#pragma unroll
        for (int i = 0; i < kPerLane; i++) {
            const int j = lane + 32 * i;
            o2.v[i] = j < H ? fmaf(x.v[i], A.a_g2[j], A.a_be2[j]) : 0.f;
            if (j < H) p = fmaf(fmaxf(o2.v[i], 0.f), A.a_w3[j], p);
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
                const float z = fmaf(c.v[i], A.c_g2[j], A.c_be2[j]) + fmaf(a, A.c_wa[j], A.c_ba[j]);
                if (z > 0.f) da = fmaf(dq * A.c_wq[j], A.c_wa[j], da);
            }
        }
        da = warp_sum(da);
        const float dp = da * (1.0f - a * a);                             // through tanh
        Row dout, dx;
#pragma unroll
        for (int i = 0; i < kPerLane; i++) { const int j = lane + 32 * i; dout.v[i] = (j < H && o2.v[i] > 0.f) ? dp * A.a_w3[j] : 0.f; }
        ln_backward_row(dout, x, A.a_g2, rstd, H, lane, dx);
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
    cluster_barrier();
    if (threadIdx.x < A.B) s_dp[threadIdx.x] = A.dv[threadIdx.x];
    __syncthreads();
    const int gt = (int)cluster_rank() * kRowT + threadIdx.x;
    for (int j = gt; j < H; j += kRowThreads) {
        float gbe = 0.f, gg = 0.f, gw3 = 0.f;
#pragma unroll 32
        for (int bb = 0; bb < A.B; bb++) { const size_t o = (size_t)bb * H + j; gbe += A.sc0[o]; gg += A.sc1[o]; gw3 += A.sc2[o]; }
        A.g_be2[j] = gbe; A.g_g2[j] = gg; A.g_t0[j] = gw3;
    }
    if (gt == 0) { float sum = 0.f; for (int bb = 0; bb < A.B; bb++) sum += s_dp[bb]; A.g_t1[0] = sum; }
}

// K5 / K11: relu + LayerNorm 1 backward (warp = row), then the fc1 backward: dW1[n][k] = sum_b dh1[b][n] x[b][k], db1, dg1, dbe1
struct L1Args {
    int B, IN, H1;
    const float *h1, *da1;           // fc1 output (pre-LayerNorm) and the gradient w.r.t. relu(LN1(h1)), [B][H1]
    const float *g1, *be1;
    const float *x;                  // network input [B][IN]
    float *dh1;                      // scratch [B][H1]
    float *sc0, *sc1;                // scratch [B][H1]
    float *g_w1, *g_b1, *g_g1, *g_be1;
};
__global__ void __launch_bounds__(kRowT) learn_l1_backward_kernel(L1Args A) {
    __shared__ float xs[kMaxB * 32];
    __shared__ float ds[kMaxB * (kMaxH / kRowCtas)];      // this CTA's column slice of dh1: [B][H1 / 8]
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, H = A.H1;
    for (int v = threadIdx.x; v < A.B * A.IN; v += kRowT) xs[v] = A.x[v];
    const int b = (int)cluster_rank() * (kRowT / 32) + warp;
    if (b < A.B) {
        Row x, dout, dx;
        load_row(x, A.h1 + (size_t)b * H, H, lane);
        load_row(dout, A.da1 + (size_t)b * H, H, lane);
        const float rstd = normalize_row(x, H, lane);
#pragma unroll
        for (int i = 0; i < kPerLane; i++) {
            const int j = lane + 32 * i;
            if (!(j < H && fmaf(x.v[i], A.g1[j], A.be1[j]) > 0.f)) dout.v[i] = 0.f;      // relu mask
        }
        ln_backward_row(dout, x, A.g1, rstd, H, lane, dx);
#pragma unroll
        for (int i = 0; i < kPerLane; i++) {
            const int j = lane + 32 * i;
            if (j < H) { const size_t o = (size_t)b * H + j; A.dh1[o] = dx.v[i]; A.sc0[o] = dout.v[i]; A.sc1[o] = dout.v[i] * x.v[i]; }
        }
    }
    cluster_barrier();
    __syncthreads();
    const int gt = (int)cluster_rank() * kRowT + threadIdx.x;
    for (int j = gt; j < H; j += kRowThreads) {
        float gbe = 0.f, gg = 0.f, gb = 0.f;
#pragma unroll 32
        for (int bb = 0; bb < A.B; bb++) { const size_t o = (size_t)bb * H + j; gbe += A.sc0[o]; gg += A.sc1[o]; gb += A.dh1[o]; }
        A.g_be1[j] = gbe; A.g_g1[j] = gg; A.g_b1[j] = gb;
    }
    // dW1[n][k] = sum_b dh1[b][n] x[b][k]: CTA r owns the columns n of its slice of H1, staged in shared memory with one round
    // of independent loads (reading dh1 from L2 inside the 64-deep dot product would expose the L2 latency 64 times)
    const int per = (H + kRowCtas - 1) / kRowCtas, nlo = (int)cluster_rank() * per, nhi = min(H, nlo + per), nw = nhi - nlo;
    for (int v = threadIdx.x; v < A.B * per; v += kRowT) {
        const int bb = v / per, c = v - bb * per;
        ds[v] = c < nw ? A.dh1[(size_t)bb * H + nlo + c] : 0.f;
    }
    __syncthreads();
    for (int e = threadIdx.x; e < nw * A.IN; e += kRowT) {
        const int c = e / A.IN, k = e - c * A.IN;
        float sum = 0.f;
#pragma unroll 16
        for (int bb = 0; bb < A.B; bb++) sum = fmaf(ds[bb * per + c], xs[bb * A.IN + k], sum);
        A.g_w1[(size_t)(nlo + c) * A.IN + k] = sum;
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
    float *h1[NJOBS], *h2[NJOBS], *st1[NJOBS];
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
    size_t o_h1[NJOBS], o_h2[NJOBS], o_st[NJOBS];
    for (int j = 0; j < NJOBS; j++) { o_h1[j] = take(sizeof(float) * B * L.h1); o_h2[j] = take(sizeof(float) * B * L.h2); o_st[j] = take(sizeof(float) * 2 * B); }
    size_t o_dh2 = take(sizeof(float) * B * L.h2), o_da1 = take(sizeof(float) * B * L.h1), o_dh1 = take(sizeof(float) * B * L.h1);
    size_t o_sc[3] = {take(sizeof(float) * B * hm), take(sizeof(float) * B * hm), take(sizeof(float) * B * hm)};
    size_t o_q = take(sizeof(float) * B), o_y = take(sizeof(float) * B), o_ao = take(sizeof(float) * B), o_dv = take(sizeof(float) * B), o_step = take(256);
    if (ln) {
        auto f = [&](size_t o) { return reinterpret_cast<float *>(base + o); };
        for (int i = 0; i < 4; i++) ln->p[i] = f(o_p[i]);
        for (int i = 0; i < 2; i++) { ln->m[i] = f(o_m[i]); ln->v[i] = f(o_v[i]); ln->g[i] = f(o_g[i]); }
        ln->bt.s = f(o_s); ln->bt.s2 = f(o_s2); ln->bt.a = f(o_a); ln->bt.r = f(o_r); ln->bt.d = f(o_d);
        ln->bt.rows = reinterpret_cast<int64_t *>(base + o_rows);
        for (int j = 0; j < NJOBS; j++) { ln->h1[j] = f(o_h1[j]); ln->h2[j] = f(o_h2[j]); ln->st1[j] = f(o_st[j]); }
        ln->dh2 = f(o_dh2); ln->da1 = f(o_da1); ln->dh1 = f(o_dh1); ln->sc0 = f(o_sc[0]); ln->sc1 = f(o_sc[1]); ln->sc2 = f(o_sc[2]);
        ln->q = f(o_q); ln->y = f(o_y); ln->aout = f(o_ao); ln->dv = f(o_dv); ln->step = reinterpret_cast<int *>(base + o_step);
    }
    return off;
}

Job fwd_job(int B, int N, int K, const float *X, int ldx, const float *W, const float *bias, const float *g, const float *be, float *stats, float *Y) {
    Job j{};
    j.type = G_FWD; j.ctas = (N + 31) / 32; j.B = B; j.N = N; j.K = K; j.X = X; j.ldx = ldx; j.W = W; j.ldw = K; j.bias = bias; j.g = g; j.be = be;
    j.stats = stats; j.Y = Y; j.ldy = N;
    return j;
}

// a row-wise stage: one cluster of kRowCtas CTAs
template <typename Kern, typename Args>
int launch_rows(Kern kern, const Args &args, cudaStream_t s) {
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(kRowCtas); cfg.blockDim = dim3(kRowT); cfg.dynamicSmemBytes = 0; cfg.stream = s;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension; at[0].val.clusterDim.x = kRowCtas; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    cfg.attrs = at; cfg.numAttrs = 1;
    TT_CUDA(cudaLaunchKernelEx(&cfg, kern, args));
    TT_COUNT_LAUNCH(); TT_LAUNCH_CHECK();
    return TT_OK;
}

int launch_jobs(const JobList &L, cudaStream_t s) {
    int ctas = 0;
    for (int i = 0; i < L.n; i++) ctas += L.j[i].ctas;
    learn_gemm_kernel<<<ctas, kT, 0, s>>>(L);
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
    TT_REQUIRE(ring->d_state_mem && ring->d_action_mem && ring->d_reward_mem && ring->d_new_state_mem && ring->d_terminal_mem &&
               ring->mem_size > 0, "bad ring");
    const int64_t max_mem = ring->mem_cntr < ring->mem_size ? ring->mem_cntr : ring->mem_size;
    TT_REQUIRE(max_mem >= ln->B || d_rows, "fewer transitions in the ring than the batch size (DDPG_agent.py:73-74)");
    cudaStream_t s = tt::as_stream(stream);
    const Layout &L = ln->L;
    const int B = ln->B, IN = L.in, H1 = L.h1, H2 = L.h2;
    float *pa = ln->p[NET_ACTOR], *pta = ln->p[NET_TARGET_ACTOR], *pc = ln->p[NET_CRITIC], *ptc = ln->p[NET_TARGET_CRITIC];
    float *ga = ln->g[0], *gc = ln->g[1];
    const int T = L.tail();

    // K0
    learn_gather_kernel<<<1, 1024, 0, s>>>(*ring, max_mem, d_rows, ln->bt, B, IN, ln->seed, ln->step);
    TT_COUNT_LAUNCH(); TT_LAUNCH_CHECK();
    int rc;
    // K1: fc1 of the four forward passes
    const float *netp[NJOBS] = {pta, ptc, pc, pa};
    const float *netx[NJOBS] = {ln->bt.s2, ln->bt.s2, ln->bt.s, ln->bt.s};
    {
        JobList J{}; J.n = NJOBS;
        for (int j = 0; j < NJOBS; j++) J.j[j] = fwd_job(B, H1, IN, netx[j], IN, netp[j] + L.w1(), netp[j] + L.b1(), nullptr, nullptr, nullptr, ln->h1[j]);
        if ((rc = launch_jobs(J, s)) != TT_OK) return rc;
    }
    // K2: LayerNorm 1 + ReLU + fc2
    {
        JobList J{}; J.n = NJOBS;
        for (int j = 0; j < NJOBS; j++)
            J.j[j] = fwd_job(B, H2, H1, ln->h1[j], H1, netp[j] + L.w2(), netp[j] + L.b2(), netp[j] + L.g1(), netp[j] + L.be1(), ln->st1[j], ln->h2[j]);
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
    H.q_out = ln->q; H.y_out = ln->y; H.a_out = ln->aout; H.dv = ln->dv;
    {
        HeadArgs C = H;
        C.g_g2 = gc + L.g2(); C.g_be2 = gc + L.be2(); C.g_t0 = gc + T; C.g_t1 = gc + T + H2; C.g_t2 = gc + T + 2 * H2; C.g_t3 = gc + T + 3 * H2;
        if ((rc = launch_rows(learn_critic_head_kernel, C, s)) != TT_OK) return rc;
    }
    // backward through fc2 / LayerNorm 1 / fc1 of one network whose dh2 is in ln->dh2 and whose forward job is `job`
    auto trunk_backward = [&](int job, const float *p, float *g, const float *x) -> int {
        JobList J{}; J.n = 2;
        Job w{};
        w.type = G_WGRAD; w.B = B; w.N = H2; w.K = H1; w.ctas = ((H2 + 15) / 16) * ((H1 + 63) / 64);
        w.X = ln->h1[job]; w.ldx = H1; w.g = p + L.g1(); w.be = p + L.be1(); w.stats = ln->st1[job];
        w.D = ln->dh2; w.ldd = H2; w.Y = g + L.w2(); w.ldy = H1; w.db = g + L.b2();
        Job xg{};
        xg.type = G_XGRAD; xg.B = B; xg.N = H2; xg.K = H1; xg.ctas = (H1 + 31) / 32;
        xg.D = ln->dh2; xg.ldd = H2; xg.W = p + L.w2(); xg.ldw = H1; xg.Y = ln->da1; xg.ldy = H1;
        J.j[0] = w; J.j[1] = xg;
        int r = launch_jobs(J, s);
        if (r != TT_OK) return r;
        L1Args A{};
        A.B = B; A.IN = IN; A.H1 = H1; A.h1 = ln->h1[job]; A.da1 = ln->da1; A.g1 = p + L.g1(); A.be1 = p + L.be1(); A.x = x;
        A.dh1 = ln->dh1; A.sc0 = ln->sc0; A.sc1 = ln->sc1;
        A.g_w1 = g + L.w1(); A.g_b1 = g + L.b1(); A.g_g1 = g + L.g1(); A.g_be1 = g + L.be1();
        return launch_rows(learn_l1_backward_kernel, A, s);
    };
    auto adam = [&](float *p, float *m, float *v, float *target, const float *g, int n, float lr, float wd) -> int {
        AdamArgs A{};
        A.p = p; A.m = m; A.v = v; A.target = target; A.g = g; A.n = n; A.lr = lr; A.wd = wd; A.tau = ln->tau; A.omt = (float)(1.0 - (double)ln->tau);
        A.step = ln->step;
        int blocks = (n + 255) / 256;
        const int cap = tt::sm_count() * 2;
        if (blocks > cap) blocks = cap;
        learn_adam_kernel<<<blocks, 256, 0, s>>>(A);
        TT_COUNT_LAUNCH(); TT_LAUNCH_CHECK();
        return TT_OK;
    };
    // K4, K5, K6: critic backward + optimizer (DDPG_agent.py:95-98)
    if ((rc = trunk_backward(JOB_C, pc, gc, ln->bt.s)) != TT_OK) return rc;
    if ((rc = adam(pc, ln->m[1], ln->v[1], ptc, gc, ln->np[NET_CRITIC], ln->beta, ln->wd)) != TT_OK) return rc;
    // K7, K8: the updated critic's trunk on s (into the JOB_C buffers)
    {
        JobList J{}; J.n = 1;
        J.j[0] = fwd_job(B, H1, IN, ln->bt.s, IN, pc + L.w1(), pc + L.b1(), nullptr, nullptr, nullptr, ln->h1[JOB_C]);
        if ((rc = launch_jobs(J, s)) != TT_OK) return rc;
        J.j[0] = fwd_job(B, H2, H1, ln->h1[JOB_C], H1, pc + L.w2(), pc + L.b2(), pc + L.g1(), pc + L.be1(), ln->st1[JOB_C], ln->h2[JOB_C]);
        if ((rc = launch_jobs(J, s)) != TT_OK) return rc;
    }
    // K9: actor head
    {
        HeadArgs A = H;
        A.g_g2 = ga + L.g2(); A.g_be2 = ga + L.be2(); A.g_t0 = ga + T; A.g_t1 = ga + T + H2; A.g_t2 = nullptr; A.g_t3 = nullptr;
        if ((rc = launch_rows(learn_actor_head_kernel, A, s)) != TT_OK) return rc;
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
