// tt_learn.cu -- row f1: one DDPG update (Agent.learn, DDPG/DDPG_agent.py:72-131; CriticNetwork DDPG/networks.py:9-68,
// ActorNetwork :98-147; ReplayBuffer.sample_buffer DDPG/replay_buffer.py:23-34) as 14 small hand-written kernels on the
// caller's stream, on the device-resident replay ring, followed by the re-pack of the new policy into the rollout actor's
// operand images.  The whole sequence is capturable into one CUDA graph (nothing synchronises, the Adam step counter and
// the sampling counter live in device memory).
//
// The step is latency-bound (batch 64 x 23-400-300 networks = 150 MFLOP): what matters is the length of the dependency
// chain, not the arithmetic.  Stages (one launch each):
//   K1  sample 64 ring rows (Philox), gather s, a, r, s', done and fc1 + LayerNorm 1 + ReLU of target_actor(s'), target_critic(s'),
//       critic(s), actor(s)   -- one launch, 4 jobs, every CTA samples the rows it needs
//   K2  fc2 of the same four (grouped GEMM launch, the whole reduction in flight, split over the warps)
//   K3  critic head (8 CTAs, warp = batch row): a' = target_actor head, y = r + gamma Q'(s', a') (1 - done), q = Q(s, a),
//       dL/dq = 2 (q - y) / B, back through q / action_value / LayerNorm 2 -> d h2 and the per-row terms of the parameter gradients
//   K4  critic fc2 backward: dW2 = dh2^T a1 (grouped with) da1 = dh2 W2 (and with) the column sums of K3's per-row terms
//   K5  critic LayerNorm 1 backward (warp = row), K5b fc1 backward dW1 = dh1^T s (a dW job) + the column sums of K5
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
// Column sums over the batch (= the gradients of bias-like parameters), riding in the grouped launch that FOLLOWS the row-wise
// kernel which produced the rows:  dst[v][j] (and dst2[v][j]) = sum_b src[v][b H + j] * (scale[v] ? scale[v][b] : 1),  v < nv;
// tot_dst[0] = sum_b tot_src[b].  (The row-wise kernels used to end with a ticket: __threadfence, an atomic, and the LAST CTA
// summing the columns alone -- 4 us of pure latency per kernel on the critical path; here the sums run beside the GEMM tiles.)
struct ColSumJob {
    const float *src[4], *scale[4];
    float *dst[4], *dst2[4];
    const float *tot_src; float *tot_dst;
    int nv, B, H, ctas;                 // ctas = ceil(H / 32), 0 = no such job in this launch
};
constexpr int kMaxJobs = 4;
struct JobList { Job j[kMaxJobs]; int n; ColSumJob cs; };

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
// at most `n` of the committed groups still pending (run-time n: every chunk of a tile is in flight at once, see job_fwd)
__device__ __forceinline__ void cp_wait_n(int n) {
    switch (n) {
    case 0: cp_wait<0>(); break;   case 1: cp_wait<1>(); break;   case 2: cp_wait<2>(); break;   case 3: cp_wait<3>(); break;
    case 4: cp_wait<4>(); break;   case 5: cp_wait<5>(); break;   case 6: cp_wait<6>(); break;   case 7: cp_wait<7>(); break;
    case 8: cp_wait<8>(); break;   case 9: cp_wait<9>(); break;   case 10: cp_wait<10>(); break; case 11: cp_wait<11>(); break;
    case 12: cp_wait<12>(); break; case 13: cp_wait<13>(); break; case 14: cp_wait<14>(); break; default: cp_wait<15>(); break;
    }
}
__device__ __forceinline__ bool vec_ok(const float *p, int ld, int cols) { return (ld & 3) == 0 && (cols & 3) == 0 && ((uintptr_t)p & 15) == 0; }

// The two long GEMM tiles (forward: 64 batch rows x 16 outputs over K <= 512; dX: 32 batch rows x 32 inputs over N <= 512) are
// latency problems, not throughput problems: 1 024 outputs per CTA and a reduction of 300-400 terms.  Both run the same scheme:
//  * the reduction dimension comes in as chunks of 32, ALL of them in flight at once (one cp.async group per chunk, <= 16 chunks
//    = 184 KB of shared memory): one memory round trip plus the transfer instead of one round trip per chunk;
//  * the 8 warps SPLIT the reduction: warp w takes the w-th group of four terms of every chunk, and each lane holds an 8 x 4
//    register tile of the CTA's outputs: 12 conflict-free 16 B shared-memory loads per 128 FMAs on 32 independent accumulators
//    (the first version -- thread = 4 outputs, 5 loads per 16 FMAs on 4 dependent chains at 34 registers -- took 20 800 cycles
//    per tile at 8 warps per SM: pure instruction latency);
//  * the eight partial tiles meet in shared memory ([warp][register][lane]: conflict-free both ways) and are summed in warp order.
constexpr int kMaxChunks = 16;                                                  // kMaxH / 32
constexpr int kRedFloats = 8 * 32 * 32;                                         // the partial tiles of the 8 warps

// sum the 8 warps' partial tiles; thread tid gets outputs idx = tid + 256 q (q < 4) as out[q]; idx = reg * 32 + lane
__device__ __forceinline__ void reduce_partials(const float (&acc)[8][4], float *red, float (&out)[4]) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    __syncthreads();                                                            // every warp is done with the staging area `red` aliases
#pragma unroll
    for (int i = 0; i < 8; i++)
#pragma unroll
        for (int j2 = 0; j2 < 4; j2++) red[(warp * 32 + i * 4 + j2) * 32 + lane] = acc[i][j2];
    __syncthreads();
#pragma unroll
    for (int q = 0; q < 4; q++) {
        float t = 0.f;
#pragma unroll
        for (int w = 0; w < 8; w++) t += red[w * 1024 + threadIdx.x + 256 * q];
        out[q] = t;
    }
}

// Y tile: all B rows x 16 columns [n0, n0 + 16).  Lane (rg = lane / 4, cg = lane % 4) holds rows 8 i + rg, columns 4 j + cg.
__device__ void job_fwd(const Job &J, int cta, float *smem) {
    const int nch = (J.K + 31) / 32;
    float *Xs = smem, *Ws = smem + nch * 64 * kP;                               // [nch][64][kP], [nch][16][kP]
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, rg = lane >> 2, cg = lane & 3, n0 = cta * 16;
    const bool vx = vec_ok(J.X, J.ldx, J.K), vw = vec_ok(J.W, J.ldw, J.K);
    if (vx && vw) {
        // production sizes: every thread's source / destination of a chunk is the previous chunk's plus a constant (the generic
        // stage_block spends ~40 instructions of index arithmetic per 16 B copy: a third of this kernel's instructions)
        const int q = threadIdx.x & 7, r = threadIdx.x >> 3;                    // 16 B column group, row 0..31
        const bool rx0 = r < J.B, rx1 = r + 32 < J.B, rw = r < 16 && n0 + r < J.N;
        const float *x0 = J.X + (size_t)(rx0 ? r : 0) * J.ldx + 4 * q, *x1 = J.X + (size_t)(rx1 ? r + 32 : 0) * J.ldx + 4 * q;
        const float *w0 = J.W + (size_t)(rw ? n0 + r : 0) * J.ldw + 4 * q;
        float *dx = Xs + r * kP + 4 * q, *dw = Ws + r * kP + 4 * q;
        for (int ch = 0; ch < nch; ch++) {
            const bool in = 32 * ch + 4 * q < J.K;
            cp16(dx, in ? x0 : J.X, rx0 && in); cp16(dx + 32 * kP, in ? x1 : J.X, rx1 && in);
            if (r < 16) cp16(dw, in ? w0 : J.W, rw && in);
            cp_commit();
            x0 += 32; x1 += 32; w0 += 32; dx += 64 * kP; dw += 16 * kP;
        }
    } else {
        for (int ch = 0; ch < nch; ch++) {
            stage_block(Xs + ch * 64 * kP, J.X, J.ldx, 64, J.B, ch * 32, J.K, vx);
            stage_block(Ws + ch * 16 * kP, J.W + (size_t)n0 * J.ldw, J.ldw, 16, J.N - n0, ch * 32, J.K, vw);
            cp_commit();
        }
    }
    float acc[8][4];
#pragma unroll
    for (int i = 0; i < 8; i++)
#pragma unroll
        for (int j2 = 0; j2 < 4; j2++) acc[i][j2] = 0.f;
    for (int ch = 0; ch < nch; ch++) {
        cp_wait_n(nch - 1 - ch);
        __syncthreads();
        const float *xs = Xs + ch * 64 * kP + rg * kP + 4 * warp, *ws = Ws + ch * 16 * kP + cg * kP + 4 * warp;
        float4 b[4];
#pragma unroll
        for (int j2 = 0; j2 < 4; j2++) b[j2] = *reinterpret_cast<const float4 *>(ws + 4 * j2 * kP);
#pragma unroll
        for (int i = 0; i < 8; i++) {
            const float4 a = *reinterpret_cast<const float4 *>(xs + 8 * i * kP);
#pragma unroll
            for (int j2 = 0; j2 < 4; j2++) {
                acc[i][j2] = fmaf(a.x, b[j2].x, acc[i][j2]); acc[i][j2] = fmaf(a.y, b[j2].y, acc[i][j2]);
                acc[i][j2] = fmaf(a.z, b[j2].z, acc[i][j2]); acc[i][j2] = fmaf(a.w, b[j2].w, acc[i][j2]);
            }
        }
    }
    float out[4];
    reduce_partials(acc, smem, out);
#pragma unroll
    for (int q = 0; q < 4; q++) {
        const int idx = threadIdx.x + 256 * q, reg = idx >> 5, ln = idx & 31;
        const int b_ = 8 * (reg >> 2) + (ln >> 2), n = n0 + 4 * (reg & 3) + (ln & 3);
        if (b_ < J.B && n < J.N) J.Y[(size_t)b_ * J.ldy + n] = out[q] + (J.bias ? J.bias[n] : 0.f);
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

// dX tile: 32 batch rows x 32 columns [k0, k0 + 32), reduction over N (see above).  Lane (rg = lane / 8, cg = lane % 8) holds
// rows 4 i + rg, columns 4 cg .. 4 cg + 3.  (26 CTAs for 64 x 400: the reduction over 300 n is the long chain of the backward
// pass, so it gets the small tile.)
__device__ void job_xgrad(const Job &J, int cta, float *smem) {
    const int nch = (J.N + 31) / 32;
    float *Ds = smem, *Ws = smem + nch * 32 * kP;                               // [nch][32 b][kP] (32 n), [nch][32 n][kP] (32 k)
    const int ktiles = (J.K + 31) / 32;
    const int bt = cta / ktiles, kt = cta - bt * ktiles, b0 = bt * 32, k0 = kt * 32;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, rg = lane >> 3, cg = lane & 7;
    const bool vd = vec_ok(J.D, J.ldd, J.N), vw = vec_ok(J.W, J.ldw, J.K);
    if (vd && vw) {
        const int q = threadIdx.x & 7, r = threadIdx.x >> 3;                    // (see job_fwd) one copy of D and one of W per chunk
        const bool rd = b0 + r < J.B, cw = k0 + 4 * q < J.K;
        const float *d0 = J.D + (size_t)(rd ? b0 + r : 0) * J.ldd + 4 * q, *w0 = J.W + (size_t)r * J.ldw + (cw ? k0 + 4 * q : 0);
        float *dd = Ds + r * kP + 4 * q, *dw = Ws + r * kP + 4 * q;
        for (int ch = 0; ch < nch; ch++) {
            const bool cd = 32 * ch + 4 * q < J.N, rw = 32 * ch + r < J.N;
            cp16(dd, cd ? d0 : J.D, rd && cd);
            cp16(dw, rw ? w0 : J.W, rw && cw);
            cp_commit();
            d0 += 32; w0 += (size_t)32 * J.ldw; dd += 32 * kP; dw += 32 * kP;
        }
    } else {
        for (int ch = 0; ch < nch; ch++) {
            stage_block(Ds + ch * 32 * kP, J.D + (size_t)b0 * J.ldd, J.ldd, 32, J.B - b0, ch * 32, J.N, vd);
            stage_block(Ws + ch * 32 * kP, J.W + (size_t)ch * 32 * J.ldw, J.ldw, 32, J.N - ch * 32, k0, J.K, vw);
            cp_commit();
        }
    }
    float acc[8][4];
#pragma unroll
    for (int i = 0; i < 8; i++)
#pragma unroll
        for (int j2 = 0; j2 < 4; j2++) acc[i][j2] = 0.f;
    for (int ch = 0; ch < nch; ch++) {
        cp_wait_n(nch - 1 - ch);
        __syncthreads();
        const float *ds = Ds + ch * 32 * kP + rg * kP + 4 * warp, *ws = Ws + ch * 32 * kP + 4 * warp * kP + 4 * cg;
        float4 w[4];                                                            // W[n = 4 warp + e][k0 + 4 cg .. + 3]
#pragma unroll
        for (int e = 0; e < 4; e++) w[e] = *reinterpret_cast<const float4 *>(ws + e * kP);
#pragma unroll
        for (int i = 0; i < 8; i++) {
            const float4 d = *reinterpret_cast<const float4 *>(ds + 4 * i * kP);  // D[b = 4 i + rg][n = 4 warp .. + 3]
            acc[i][0] = fmaf(d.x, w[0].x, acc[i][0]); acc[i][1] = fmaf(d.x, w[0].y, acc[i][1]); acc[i][2] = fmaf(d.x, w[0].z, acc[i][2]); acc[i][3] = fmaf(d.x, w[0].w, acc[i][3]);
            acc[i][0] = fmaf(d.y, w[1].x, acc[i][0]); acc[i][1] = fmaf(d.y, w[1].y, acc[i][1]); acc[i][2] = fmaf(d.y, w[1].z, acc[i][2]); acc[i][3] = fmaf(d.y, w[1].w, acc[i][3]);
            acc[i][0] = fmaf(d.z, w[2].x, acc[i][0]); acc[i][1] = fmaf(d.z, w[2].y, acc[i][1]); acc[i][2] = fmaf(d.z, w[2].z, acc[i][2]); acc[i][3] = fmaf(d.z, w[2].w, acc[i][3]);
            acc[i][0] = fmaf(d.w, w[3].x, acc[i][0]); acc[i][1] = fmaf(d.w, w[3].y, acc[i][1]); acc[i][2] = fmaf(d.w, w[3].z, acc[i][2]); acc[i][3] = fmaf(d.w, w[3].w, acc[i][3]);
        }
    }
    float out[4];
    reduce_partials(acc, smem, out);
#pragma unroll
    for (int q = 0; q < 4; q++) {
        const int idx = threadIdx.x + 256 * q, reg = idx >> 5, ln = idx & 31;
        const int b_ = b0 + 4 * (reg >> 2) + (ln >> 3), k = k0 + 4 * (ln & 7) + (reg & 3);
        if (b_ < J.B && k < J.K) J.Y[(size_t)b_ * J.ldy + k] = out[q];
    }
}

// 32 columns x 8 groups of 8 batch rows per CTA: 8 nv independent loads per thread (one memory round trip), the groups meet in
// shared memory and are added in order (deterministic)
__device__ void job_colsum(const ColSumJob &C, int cta, float *smem) {
    const int cl = threadIdx.x & 31, rg = threadIdx.x >> 5, j = cta * 32 + cl;
    // every load first (nvcc otherwise keeps load -> use -> load order: 8 nv dependent L2 round trips), then the arithmetic
    float x[4][8], sc[4][8];
#pragma unroll
    for (int v = 0; v < 4; v++)
#pragma unroll
        for (int i = 0; i < 8; i++) {
            const int b = rg * 8 + i;
            const bool in = v < C.nv && b < C.B && j < C.H;
            x[v][i] = in ? __ldcg(C.src[v] + (size_t)b * C.H + j) : 0.f;
            sc[v][i] = (in && C.scale[v]) ? __ldcg(C.scale[v] + b) : 1.f;
        }
    float tot = 0.f;
    const bool do_tot = cta == 0 && rg == 7 && C.tot_dst;                       // warp 7 of the first CTA: sum of tot_src (B <= 64)
    if (do_tot) tot = (cl < C.B ? __ldcg(C.tot_src + cl) : 0.f) + (cl + 32 < C.B ? __ldcg(C.tot_src + cl + 32) : 0.f);
#pragma unroll
    for (int v = 0; v < 4; v++) {
        float part = 0.f;
#pragma unroll
        for (int i = 0; i < 8; i++) part = C.scale[v] ? fmaf(x[v][i], sc[v][i], part) : part + x[v][i];
        smem[(v * 8 + rg) * 32 + cl] = part;
    }
    __syncthreads();
    if (rg < C.nv && j < C.H) {
        float t = 0.f;
#pragma unroll
        for (int g = 0; g < 8; g++) t += smem[(rg * 8 + g) * 32 + cl];
        C.dst[rg][j] = t;
        if (C.dst2[rg]) C.dst2[rg][j] = t;
    }
    if (do_tot) { tot = warp_sum(tot); if (cl == 0) C.tot_dst[0] = tot; }
}

// shared memory (floats) a job's CTAs need
__host__ __device__ inline int job_smem_floats(const Job &J) {
    int f = (64 + 64) * kP;
    if (J.type == G_FWD) f = ((J.K + 31) / 32) * (64 + 16) * kP;
    if (J.type == G_XGRAD) f = ((J.N + 31) / 32) * (32 + 32) * kP;
    return J.type != G_WGRAD && f < kRedFloats ? kRedFloats : f;
}

__global__ void __launch_bounds__(kT) learn_gemm_kernel(JobList L) {
    chain_enter();
    extern __shared__ __align__(16) float smem[];                              // max over the launch's jobs of job_smem_floats()
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
    if (cta < L.cs.ctas) job_colsum(L.cs, cta, smem);
}

// ---- K0 (inside K1): sampling + gather (replay_buffer.py:23-34: uniform with replacement over the filled part of the ring) ----
struct Batch {
    float *s, *s2, *a, *r, *d;        // [B][in], [B][in], [B], [B], [B] (done as 0 / 1)
    int64_t *rows;                    // [B]
};
// The first fc1 launch of an update samples its own rows: every CTA derives the ring rows of its 8 batch rows from the update
// counter (Philox stream 2) and reads its inputs straight from the ring; the CTAs of the critic(s) job also write the batch
// (s, a, r, done, rows) that the later stages read, those of the target-critic job s'.  A separate gather launch in front cost
// 4 us of the chain.  The counter is advanced by the critic-head kernel, after every CTA here has read it.
struct Gather {
    tt_replay_ring ring;
    int64_t win_begin, max_mem;       // sampling window: `max_mem` rows starting at `win_begin` (wrapping)
    const int64_t *given_rows;        // explicit rows instead of sampling (tests), or NULL
    Batch bt;
    uint64_t seed;
    const int *step;                  // updates done so far = the sampling counter of this one
    int on;
};
__device__ __forceinline__ int64_t sample_row(const Gather &G, int b, int t) {
    if (G.given_rows) return G.given_rows[b];
    uint32_t w[4];
    ttm::philox4x32_10((uint32_t)b, (uint32_t)t, 2u, 0u, (uint32_t)G.seed, (uint32_t)(G.seed >> 32), w);
    const double u = ((double)w[0] * 4294967296.0 + (double)w[1]) * (1.0 / 18446744073709551616.0);       // [0, 1)
    int64_t row = (int64_t)(u * (double)G.max_mem);
    if (row >= G.max_mem) row = G.max_mem - 1;
    return (G.win_begin + row) % G.ring.mem_size;                  // the sampling window may wrap around the ring
}

// ---- row-wise stages: 8 CTAs x 8 warps, one warp per batch row with the row in registers (one warp per row keeps the
//      dependent chain of a row -- load -> statistics -> head -> backward -- at one memory latency per network instead of
//      serialising rows).  The column sums over the batch (= the parameter gradients) are ColSumJobs of the grouped launch that
//      follows.  No cluster and no grid barrier: the CTAs need not be co-resident, so the stage also runs on the two or four
//      SMs that rollout.AsyncTrainer leaves free for the learner.
constexpr int kRowT = 256, kPerLane = kMaxH / 32;
struct Row { float v[kPerLane]; };

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
    int *step;                                                            // update counter (advanced by the critic head)
    float *q_out, *y_out, *a_out;                                         // diagnostics / hand-over: Q(s,a), target, actor(s)
};

// K1 / K7: fc1 + LayerNorm 1 + ReLU of up to four networks in one launch.  A CTA takes 8 batch rows of one network; a THREAD owns
// one or two output columns for all 8 rows: W1 comes in as one flat 16 B cp.async copy (no index arithmetic, every element read
// from shared memory by exactly one thread of the CTA), the 8 input rows are broadcast reads, and the LayerNorm statistics are
// two block-wide reductions.  (The first version gave every WARP a row: each warp then read the whole of W1 from shared
// memory -- 3 300 instructions per warp and 16 us for 0.6 MFLOP.)  Writes h1 (pre-LayerNorm, for the backward pass) and
// a1 = relu(LN1(h1)) (the operand of fc2 and of its weight gradient).
struct Fc1Job { const float *x, *w1, *b1, *g1, *be1; float *h1, *a1; int from_s2, writes_batch; };   // (the last two: gather mode)
struct Fc1Args { Fc1Job j[NJOBS]; int njobs, B, IN, H1; Gather G; };
constexpr int kFc1T = 256, kFc1Cols = kMaxH / kFc1T, kFc1Rows = 8, kXP = 24;    // kXP: pitch of an input row (IN <= 24)
__global__ void __launch_bounds__(kFc1T, 1) learn_fc1_kernel(Fc1Args A) {
    chain_enter();
    extern __shared__ __align__(16) float ws[];            // W1 flat [H1 * IN] (row c at c * IN: an odd IN is conflict-free for lane = column)
    __shared__ __align__(16) float xs[kFc1Rows * kXP];     // the CTA's 8 input rows, zero padded
    __shared__ float red[2][kFc1T / 32][kFc1Rows];
    const int per = (A.B + kFc1Rows - 1) / kFc1Rows;
    const int job = blockIdx.x / per, b0 = (blockIdx.x - job * per) * kFc1Rows;
    const Fc1Job J = A.j[job];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, H = A.H1, IN = A.IN, nw = H * IN;
    if ((nw & 3) == 0 && (reinterpret_cast<uintptr_t>(J.w1) & 15) == 0) {
        for (int v = tid; v < nw / 4; v += kFc1T) cp16(ws + 4 * v, J.w1 + 4 * v, true);
    } else {
        for (int v = tid; v < nw; v += kFc1T) cp4(ws + v, J.w1 + v, true);
    }
    cp_commit();
    // this thread's columns of the three parameter vectors and the input rows (gather mode: found by sampling): in flight together with W1
    float pb[kFc1Cols], pg[kFc1Cols], pe[kFc1Cols];
#pragma unroll
    for (int i = 0; i < kFc1Cols; i++) {
        const int c = tid + kFc1T * i;
        pb[i] = c < H ? J.b1[c] : 0.f; pg[i] = c < H ? J.g1[c] : 0.f; pe[i] = c < H ? J.be1[c] : 0.f;
    }
    if (tid < kFc1Rows * kXP) {
        const int r = tid / kXP, k = tid - r * kXP, b = b0 + r;
        float x = 0.f;
        if (A.G.on) {
            if (b < A.B) {
                const int64_t row = sample_row(A.G, b, *A.G.step);
                const float *src = J.from_s2 ? A.G.ring.d_new_state_mem : A.G.ring.d_state_mem;
                if (k < IN) x = src[row * IN + k];
                if (J.writes_batch) {
                    if (k < IN) (J.from_s2 ? A.G.bt.s2 : A.G.bt.s)[(size_t)b * IN + k] = x;
                    if (k == 0 && !J.from_s2) {
                        A.G.bt.rows[b] = row; A.G.bt.a[b] = A.G.ring.d_action_mem[row]; A.G.bt.r[b] = A.G.ring.d_reward_mem[row];
                        A.G.bt.d[b] = A.G.ring.d_terminal_mem[row] ? 1.f : 0.f;
                    }
                }
            }
        } else if (k < IN && b < A.B) x = J.x[(size_t)b * IN + k];
        xs[tid] = x;
    }
    cp_wait<0>();
    __syncthreads();
    float h[kFc1Cols][kFc1Rows];
#pragma unroll
    for (int i = 0; i < kFc1Cols; i++) {
        const int c = tid + kFc1T * i;
#pragma unroll
        for (int r = 0; r < kFc1Rows; r++) h[i][r] = 0.f;
        if (c < H) {
            const float *w = ws + c * IN;
#pragma unroll
            for (int k4 = 0; k4 < kXP / 4; k4++) {
                float wk[4];
#pragma unroll
                for (int e = 0; e < 4; e++) wk[e] = 4 * k4 + e < IN ? w[4 * k4 + e] : 0.f;
#pragma unroll
                for (int r = 0; r < kFc1Rows; r++) {
                    const float4 x = *reinterpret_cast<const float4 *>(xs + r * kXP + 4 * k4);
                    h[i][r] = fmaf(x.x, wk[0], h[i][r]); h[i][r] = fmaf(x.y, wk[1], h[i][r]);
                    h[i][r] = fmaf(x.z, wk[2], h[i][r]); h[i][r] = fmaf(x.w, wk[3], h[i][r]);
                }
            }
#pragma unroll
            for (int r = 0; r < kFc1Rows; r++) h[i][r] += pb[i];
        }
    }
    // LayerNorm statistics of the 8 rows: mean, then the centred sum of squares (two passes, like normalize_row)
    float stat[kFc1Rows];
#pragma unroll
    for (int pass = 0; pass < 2; pass++) {
#pragma unroll
        for (int r = 0; r < kFc1Rows; r++) {
            float v = 0.f;
#pragma unroll
            for (int i = 0; i < kFc1Cols; i++) {
                const bool in = tid + kFc1T * i < H;
                if (pass == 0) v += in ? h[i][r] : 0.f;
                else { const float d = in ? h[i][r] - stat[r] : 0.f; v = fmaf(d, d, v); }
            }
            v = warp_sum(v);
            if (lane == 0) red[pass][warp][r] = v;
        }
        __syncthreads();
#pragma unroll
        for (int r = 0; r < kFc1Rows; r++) {
            float t = 0.f;
#pragma unroll
            for (int w8 = 0; w8 < kFc1T / 32; w8++) t += red[pass][w8][r];
            if (pass == 0) stat[r] = t / (float)H;                                   // mean
            else {
                const float mean = stat[r], rstd = rsqrtf(t / (float)H + kLnEps);
#pragma unroll
                for (int i = 0; i < kFc1Cols; i++) {
                    const int c = tid + kFc1T * i;
                    if (c < H && b0 + r < A.B) {
                        J.h1[(size_t)(b0 + r) * H + c] = h[i][r];
                        J.a1[(size_t)(b0 + r) * H + c] = fmaxf(fmaf((h[i][r] - mean) * rstd, pg[i], pe[i]), 0.f);
                    }
                }
            }
        }
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
    __shared__ float sp[13][kMaxH];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, H = A.H2;
    const int b = (int)blockIdx.x * (kRowT / 32) + warp;
    Row x, xt, xc;                                                       // the three rows' loads and the 13 vectors: one memory latency together
    if (b < A.B) {
        load_row(xt, A.h2[JOB_TA] + (size_t)b * H, H, lane);
        load_row(xc, A.h2[JOB_TC] + (size_t)b * H, H, lane);
        load_row(x, A.h2[JOB_C] + (size_t)b * H, H, lane);
    }
    {
        const float *const src[13] = {A.ta_g2, A.ta_be2, A.ta_w3, A.tc_g2, A.tc_be2, A.tc_wa, A.tc_ba, A.tc_wq, A.c_g2, A.c_be2, A.c_wa, A.c_ba, A.c_wq};
        stage_vectors<13>(sp, src, H);
    }
    const float *ta_g2 = sp[0], *ta_be2 = sp[1], *ta_w3 = sp[2], *tc_g2 = sp[3], *tc_be2 = sp[4], *tc_wa = sp[5], *tc_ba = sp[6], *tc_wq = sp[7],
                *c_g2 = sp[8], *c_be2 = sp[9], *c_wa = sp[10], *c_ba = sp[11], *c_wq = sp[12];
    __syncthreads();
    if (b < A.B) {
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
    if (blockIdx.x == 0 && threadIdx.x == 0) *A.step = *A.step + 1;     // this update's rows are sampled (learn_fc1): Adam reads t + 1
}

// K9: DDPG_agent.py:99-103  actor_loss = -mean(critic(states, actor(states)))
__global__ void __launch_bounds__(kRowT) learn_actor_head_kernel(HeadArgs A) {
    chain_enter();
    __shared__ float sp[8][kMaxH];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, H = A.H2;
    const int b = (int)blockIdx.x * (kRowT / 32) + warp;
    Row x, c;
    if (b < A.B) {
        load_row(x, A.h2[JOB_A] + (size_t)b * H, H, lane);
        load_row(c, A.h2[JOB_C] + (size_t)b * H, H, lane);
    }
    {
        const float *const src[8] = {A.a_g2, A.a_be2, A.a_w3, A.c_g2, A.c_be2, A.c_wa, A.c_ba, A.c_wq};
        stage_vectors<8>(sp, src, H);
    }
    const float *a_g2 = sp[0], *a_be2 = sp[1], *a_w3 = sp[2], *c_g2 = sp[3], *c_be2 = sp[4], *c_wa = sp[5], *c_ba = sp[6], *c_wq = sp[7];
    __syncthreads();
    if (b < A.B) {
        Row o2;
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
}

// K5 / K11: relu + LayerNorm 1 backward (warp = row) -> dh1.  The fc1 weight / bias gradients (dW1 = dh1^T x, db1 = column sums of
// dh1) and the LayerNorm parameter gradients dg1, dbe1 (column sums of sc1, sc0) are jobs of the next grouped launch.
struct L1Args {
    int B, H1;
    const float *h1, *da1;           // fc1 output (pre-LayerNorm) and the gradient w.r.t. relu(LN1(h1)), [B][H1]
    const float *g1, *be1;
    float *dh1;                      // out [B][H1]
    float *sc0, *sc1;                // scratch [B][H1]
};
__global__ void __launch_bounds__(kRowT) learn_l1_backward_kernel(L1Args A) {
    chain_enter();
    __shared__ float sp[2][kMaxH];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, H = A.H1;
    const int b = (int)blockIdx.x * (kRowT / 32) + warp;
    Row x, dout;
    if (b < A.B) {
        load_row(x, A.h1 + (size_t)b * H, H, lane);
        load_row(dout, A.da1 + (size_t)b * H, H, lane);
    }
    {
        const float *const src[2] = {A.g1, A.be1};
        stage_vectors<2>(sp, src, H);
    }
    const float *g1 = sp[0], *be1 = sp[1];
    __syncthreads();
    if (b < A.B) {
        Row dx;
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
    int ctas = 0, fl = 0;
    for (int i = 0; i < L.n; i++) { ctas += L.j[i].ctas; const int f = job_smem_floats(L.j[i]); if (f > fl) fl = f; }
    ctas += L.cs.ctas;
    if (L.cs.ctas && fl < 4 * 8 * 32) fl = 4 * 8 * 32;
    static bool attr_of[tt::kMaxDevices] = {};
    if (!attr_of[tt::device_index()]) {
        TT_CUDA(cudaFuncSetAttribute(learn_gemm_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(sizeof(float) * kMaxChunks * (64 + 16) * kP)));
        attr_of[tt::device_index()] = true;
    }
    TT_CUDA(tt::launch_chained(true, learn_gemm_kernel, dim3(ctas), dim3(kT), sizeof(float) * (size_t)fl, s, L));
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
    // developer build (-DTT_LEARN_PROFILE, profiles/learner_stage_times.py): TT_LEARN_STOP=k ends the sequence after k launches
#if defined(TT_LEARN_PROFILE)
    const char *e_ = getenv("TT_LEARN_STOP"); const int stop_at = e_ ? atoi(e_) : 0; int nl_ = 0;
#define TT_STAGE_END(ret) do { if (++nl_ == stop_at) return ret; } while (0)
#else
#define TT_STAGE_END(ret) do { } while (0)
#endif
    const Layout &L = ln->L;
    const int B = ln->B, IN = L.in, H1 = L.h1, H2 = L.h2;
    float *pa = ln->p[NET_ACTOR], *pta = ln->p[NET_TARGET_ACTOR], *pc = ln->p[NET_CRITIC], *ptc = ln->p[NET_TARGET_CRITIC];
    float *ga = ln->g[0], *gc = ln->g[1];
    const int T = L.tail();

    int rc;
    // K0 + K1: sampling + gather inside fc1 + LayerNorm 1 + ReLU of the four forward passes
    const float *netp[NJOBS] = {pta, ptc, pc, pa};
    const float *netx[NJOBS] = {ln->bt.s2, ln->bt.s2, ln->bt.s, ln->bt.s};
    auto fc1 = [&](const int *jobs, int njobs, bool gather) -> int {
        Fc1Args A{};
        A.njobs = njobs; A.B = B; A.IN = IN; A.H1 = H1;
        for (int i = 0; i < njobs; i++) {
            const int j = jobs[i];
            A.j[i] = Fc1Job{netx[j], netp[j] + L.w1(), netp[j] + L.b1(), netp[j] + L.g1(), netp[j] + L.be1(), ln->h1[j], ln->a1[j],
                            j == JOB_TA || j == JOB_TC, j == JOB_C || j == JOB_TC};
        }
        if (gather) A.G = Gather{*ring, win_begin, max_mem, d_rows, ln->bt, ln->seed, ln->step, 1};
        const size_t dsm = sizeof(float) * (size_t)H1 * IN + 16;
        static bool attr_of[tt::kMaxDevices] = {};
        if (dsm > 48 * 1024 && !attr_of[tt::device_index()]) {
            TT_CUDA(cudaFuncSetAttribute(learn_fc1_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(sizeof(float) * kMaxH * kXP + 16)));
            attr_of[tt::device_index()] = true;
        }
        TT_CUDA(tt::launch_chained(true, learn_fc1_kernel, dim3(njobs * ((B + kFc1Rows - 1) / kFc1Rows)), dim3(kFc1T), dsm, s, A));
        TT_COUNT_LAUNCH(); TT_LAUNCH_CHECK();
        return TT_OK;
    };
    {
        const int all[NJOBS] = {JOB_TA, JOB_TC, JOB_C, JOB_A};
        if ((rc = fc1(all, NJOBS, true)) != TT_OK) return rc;
        TT_STAGE_END(TT_OK);
    }
    // K2: fc2
    {
        JobList J{}; J.n = NJOBS;
        for (int j = 0; j < NJOBS; j++) J.j[j] = fwd_job(B, H2, H1, ln->a1[j], netp[j] + L.w2(), netp[j] + L.b2(), ln->h2[j]);
        if ((rc = launch_jobs(J, s)) != TT_OK) return rc;
        TT_STAGE_END(TT_OK);
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
    H.q_out = ln->q; H.y_out = ln->y; H.a_out = ln->aout; H.dv = ln->dv; H.step = ln->step;
    {
        if ((rc = launch_rows(learn_critic_head_kernel, H, B, s)) != TT_OK) return rc;
        TT_STAGE_END(TT_OK);
    }
    // backward through fc2 / LayerNorm 1 / fc1 of one network whose dh2 is in ln->dh2 and whose forward job is `job`
    // (`head`: the column sums that turn the head kernel's per-row terms sc0 / sc1 / sc2 / dv into its parameter gradients)
    auto trunk_backward = [&](int job, const float *p, float *g, const float *x, const ColSumJob &head) -> int {
        JobList J{}; J.n = 2; J.cs = head;
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
        TT_STAGE_END(77);
        L1Args A{};
        A.B = B; A.H1 = H1; A.h1 = ln->h1[job]; A.da1 = ln->da1; A.g1 = p + L.g1(); A.be1 = p + L.be1();
        A.dh1 = ln->dh1; A.sc0 = ln->sc0; A.sc1 = ln->sc1;
        if ((r = launch_rows(learn_l1_backward_kernel, A, B, s)) != TT_OK) return r;
        TT_STAGE_END(77);
        JobList J1{}; J1.n = 1;
        Job w1{};
        w1.type = G_WGRAD; w1.B = B; w1.N = H1; w1.K = IN; w1.ctas = ((H1 + 15) / 16) * ((IN + 31) / 32);
        w1.X = x; w1.ldx = IN; w1.D = ln->dh1; w1.ldd = H1; w1.Y = g + L.w1(); w1.ldy = IN; w1.db = g + L.b1();
        J1.j[0] = w1;
        ColSumJob c1{};                                     // LayerNorm 1: d be1 = column sums of sc0, d g1 = of sc1
        c1.nv = 2; c1.B = B; c1.H = H1; c1.ctas = (H1 + 31) / 32;
        c1.src[0] = ln->sc0; c1.dst[0] = g + L.be1(); c1.src[1] = ln->sc1; c1.dst[1] = g + L.g1();
        J1.cs = c1;
        r = launch_jobs(J1, s); if (r != TT_OK) return r;
        TT_STAGE_END(77);
        return TT_OK;
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
    {
        ColSumJob c{};                                      // critic head: d be2 = d action_value.bias = sum dz, d action_value.weight = sum dz a,
        c.nv = 4; c.B = B; c.H = H2; c.ctas = (H2 + 31) / 32;  // d g2 = sum dz xhat, d q.weight = sum dq relu(z), d q.bias = sum dq
        c.src[0] = ln->sc0; c.dst[0] = gc + L.be2(); c.dst2[0] = gc + T + H2;
        c.src[1] = ln->sc0; c.scale[1] = ln->bt.a; c.dst[1] = gc + T;
        c.src[2] = ln->sc1; c.dst[2] = gc + L.g2();
        c.src[3] = ln->sc2; c.dst[3] = gc + T + 2 * H2;
        c.tot_src = ln->dv; c.tot_dst = gc + T + 3 * H2;
        if ((rc = trunk_backward(JOB_C, pc, gc, ln->bt.s, c)) != TT_OK) return rc == 77 ? TT_OK : rc;
    }
    if ((rc = adam(pc, ln->m[1], ln->v[1], ptc, gc, ln->np[NET_CRITIC], ln->beta, ln->wd)) != TT_OK) return rc;
    TT_STAGE_END(TT_OK);
    // K7, K8: the updated critic's trunk on s (into the JOB_C buffers)
    {
        const int one[1] = {JOB_C};
        if ((rc = fc1(one, 1, false)) != TT_OK) return rc;
        TT_STAGE_END(TT_OK);
        JobList J{}; J.n = 1;
        J.j[0] = fwd_job(B, H2, H1, ln->a1[JOB_C], pc + L.w2(), pc + L.b2(), ln->h2[JOB_C]);
        if ((rc = launch_jobs(J, s)) != TT_OK) return rc;
        TT_STAGE_END(TT_OK);
    }
    // K9: actor head
    {
        if ((rc = launch_rows(learn_actor_head_kernel, H, B, s)) != TT_OK) return rc;
        TT_STAGE_END(TT_OK);
    }
    // K10, K11, K12 (DDPG_agent.py:99-106)
    {
        ColSumJob c{};                                      // actor head: d be2 = sum dout, d g2 = sum dout xhat, d mu.weight = sum dp relu(o2), d mu.bias = sum dp
        c.nv = 3; c.B = B; c.H = H2; c.ctas = (H2 + 31) / 32;
        c.src[0] = ln->sc0; c.dst[0] = ga + L.be2();
        c.src[1] = ln->sc1; c.dst[1] = ga + L.g2();
        c.src[2] = ln->sc2; c.dst[2] = ga + T;
        c.tot_src = ln->dv; c.tot_dst = ga + T + H2;
        if ((rc = trunk_backward(JOB_A, pa, ga, ln->bt.s, c)) != TT_OK) return rc == 77 ? TT_OK : rc;
    }
    if ((rc = adam(pa, ln->m[0], ln->v[0], pta, ga, ln->np[NET_ACTOR], ln->alpha, 0.f)) != TT_OK) return rc;
    // hand the new policy to the rollout actor (re-pack into the fp32 / tensor-core operand images)
    if (repack_into) {
        const float *a = pa;
        return tt_actor_load(repack_into, a + L.w1(), a + L.b1(), a + L.g1(), a + L.be1(), a + L.w2(), a + L.b2(), a + L.g2(), a + L.be2(),
                             a + T, a + T + H2, stream);
    }
    return TT_OK;
#undef TT_STAGE_END
}

}  // extern "C"
