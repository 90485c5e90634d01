// tt_actor_tc4.cu -- kernel (c), tensor-core path, 4th version ("v4") of the batched actor forward
// (ActorNetwork.forward, DDPG/networks.py:138-147) on tcgen05 / TMEM.
//
// What changed against v3 (tt_actor_tc.cu) and why.  Measured on v3 (profiles/): a tile took 14 300 cycles of which the
// tensor pipe was busy 5 000; the rest was the CUDA-core side running strictly after / before the MMAs, and every
// tcgen05.mma with N <= 64 costs 52 cycles whatever N (profiles/mma_probe.cu), N = 256 + 48 therefore 185 per k-step.
//  * LayerNorm 1 is folded into the operands.  W1' = diag(g1) (W1 - 1 m^T) with m = column mean of [fc1.weight | fc1.bias]
//    makes the tensor core emit t' = g1 * (h - mean(h)) directly, so epilogue 1 is ONE packed FMA per column pair
//    (y = relu(t' * rstd + be1)).
//  * The LayerNorm variance comes from the tensor core too: var(h) = x^T Gc x / 400 with Gc the Gram matrix of the
//    centred weights; with the Cholesky factor Gc = L L^T, x^T Gc x = |L^T x|^2, and L^T x is 24 extra output columns
//    of the layer-1 GEMM (32 TMEM columns in front of part 0).  Every epilogue thread reads them and squares them: no
//    Gram-matrix phase on the CUDA cores (2 600 cycles per tile in v3), no cross-warp reduction, no barrier.
//  * Layer 1 = 3 parts (N = 128 | 160 | 144: statistics + 96, 160, 144 columns) through one 160-column TMEM window; an
//    epilogue thread pulls its 8 columns of every 32-column A2 block of the part into registers at once and frees the
//    window before it does the math, so the next part's MMAs run under it.
//  * Layer 2 is split by OUTPUT columns into two accumulator halves (N = 160 and N = 144; 84 + 76 cycles per k-step instead of
//    185) with separate full/free barriers, their k-blocks interleaved (half A leads by kLead blocks, tt_tc4_layout.cuh): the
//    statistics pass over half A runs under the rest of half B -- placed where the epilogue would otherwise wait for the
//    layer-1 MMAs of part 1 --, the next tile's half A starts as soon as pass 2 has read half A, and the A2 blocks are
//    released to the next tile's epilogue 1 from step kLead + 1 on.  W2 is streamed in half-size k-blocks through a 6-slot
//    ring (5 KB multicast halves per CTA of a pair).
//  * LayerNorm 2 and the output layer are folded into the GEMM as far as they are linear (see the pack kernels): W2 is
//    centred over its 300 outputs, so the accumulators are h - mean(h) and the variance is a plain sum of squares; and with
//    relu(y) = (y + |y|) / 2 the y / 2 half of the final dot product is two extra output columns (hi / lo) in the padding.
//    Epilogue 2 is then ONE FFMA2 per column pair in pass 1 and TWO in pass 2 (|x rstd + be / g| (w3 |g| / 2), the |.| is a
//    source modifier), with two parameter vectors instead of three.
//  * The observation tile (layer 1's A operand) lives in tensor memory: every epilogue thread converts 8 inputs of its own
//    row and writes them with one tcgen05.st (two in split mode); the 16 KB of shared memory this frees are two more ring
//    slots.  Epilogue 2: contiguous balanced column ranges (40 + 36 columns per thread).
//  * An mbarrier wait costs ~90 cycles even when the phase is already complete: epilogue 1 waits for A2 blocks 4 times per
//    tile (they are released in order), not 13 times, and fetches its be1 vectors one chunk ahead of the wait.
//
// TMEM columns: [0,160) H2 half A | [160,304) H2 half B | [304,464) layer-1 window (part 0: 32 statistic columns first) |
// [464,496) observation tile, hi and lo halves.
// Warp roles (640 threads): warps 0-15 epilogue (warp w: TMEM lanes 32 (w % 4), column group w / 4), warp 16 lane 0
// layer-2 MMA issuer, warp 17 lane 0 bulk-copy producer (W2 k-blocks streamed from L2), warp 18 lane 0 layer-1 MMA
// issuer, warp 19 fused replay-ring store of the observation rows.
#include <stdlib.h>
#include "tt_actor.cuh"
#include "tt_common.cuh"
#include "tt_tc_ptx.cuh"
#include "tt_tc4_layout.cuh"

// Timing-only ablations (WRONG results; profiles/build_variants.sh): which stage pins the tile time?
//   1: no W2 stream (the layer-2 MMAs read whatever the ring holds)   2: epilogue 2 without pass 2's math
//   4: epilogue 1 without the A2 conversion / stores                 8: pass 1 without its math
//  16: the layer-2 MMAs do not wait for epilogue 2 to have read the previous tile's accumulators
#ifndef TT_TC4_P1A
#define TT_TC4_P1A (TT_TC4_LEAD >= 9 ? 0 : TT_TC4_LEAD >= 6 ? 1 : 2)
#endif
// The timing-only ablations and the per-phase cycle counters exist only in the developer build (-DTT_DEV_VARIANTS,
// profiles/build_variants.sh); the product library carries neither.
#if !defined(TT_DEV_VARIANTS)
#undef TT_ABLATE
#endif
#ifndef TT_ABLATE
#define TT_ABLATE 0
#endif

namespace {

constexpr int kP1A = TT_TC4_P1A;
// A2 blocks the epilogue waits on (bit ch = block ch ends a run of blocks released together; bit 12 must be set): the
// layer-2 issuer commits exactly these, epilogue 1 waits for a run's last block before it writes the run's first.
#ifndef TT_TC4_A2WAITS
#define TT_TC4_A2WAITS 0x1484      // blocks 2, 7, 10, 12
#endif
constexpr uint32_t kA2Waits = TT_TC4_A2WAITS;
static_assert((kA2Waits >> (KB2 - 1)) == 1u, "the last A2 block must end the last run");
__host__ __device__ constexpr bool a2_run_start(int ch) { return ch == 0 || ((kA2Waits >> (ch - 1)) & 1u); }
__host__ __device__ constexpr int a2_run_end(int ch) { return ((kA2Waits >> ch) & 1u) ? ch : a2_run_end(ch + 1); }

// ---- v4 layer-1 image, built at tt_actor_load time ----
// rows 0..23   : L[k][j] (row j): lower Cholesky factor of Gc = sum_c (Wf[c] - m)(Wf[c] - m)^T, so that
//                sum_j (row_j . x)^2 = x^T Gc x = sum_c (h_c - mean(h))^2;   rows 24..31: 0
// rows 32..431 : g1[c] * (Wf[c][k] - m[k]),  Wf = [fc1.weight | fc1.bias], m = column means
// Written as [hi | lo] f16 blocks (lo = rounding residual) and as one bf16 block.
// Two launches (the policy is re-packed after every learner step, so this is on the end-to-end path):
//   A. partial sums  S_ij = sum_c Wf[c][i] Wf[c][j],  s_i = sum_c Wf[c][i]  over a slice of c per block, each block into its OWN
//      slot of the scratch (no atomics, nothing to zero, summed in block order: deterministic)
//   C. every image block: column means m = s / 400 -> rows 32..431;  block 0 alone: Gc = S - s s^T / 400, Cholesky (ONE warp,
//      lane = row held in registers, no block-wide barrier in the 24 dependent pivot steps) -> rows 0..31.
//      (A separate 576-thread Cholesky kernel between A and C cost 10 us of the 19 us re-pack: 48 block-wide barriers and a
//      float64 sqrt + division per pivot, and a launch.)
constexpr int kGramBlocks = 16;
constexpr int kGramOff = 2048;                      // l1c_scratch: [1200, 2032) layer-2 column statistics, [2048, 2048 + 16 * 600) Gram partials
__device__ __forceinline__ double wfull(const float *__restrict__ fc1_w, const float *__restrict__ fc1_b, int c, int k) {
    return (double)(k < IN ? fc1_w[c * IN + k] : fc1_b[c]);
}
__device__ __forceinline__ void pack_l1c_gram(int bid, double *__restrict__ part, const float *__restrict__ fc1_w, const float *__restrict__ fc1_b) {
    const int t = threadIdx.x, i = t / 24, j = t - i * 24;
    if (t >= 576) return;
    const int c0 = bid * (H1 / kGramBlocks), c1 = c0 + H1 / kGramBlocks;
    double si = 0.0, sij = 0.0;
#pragma unroll 5
    for (int c = c0; c < c1; c++) { const double a = wfull(fc1_w, fc1_b, c, i), b = wfull(fc1_w, fc1_b, c, j); si += a; sij += a * b; }
    part[bid * 600 + t] = sij;
    if (j == 0) part[bid * 600 + 576 + i] = si;
}
static_assert(H1 % kGramBlocks == 0, "gram slices");

// sum over the Gram blocks of element e (< 600), in block order; the 16 loads are independent
__device__ __forceinline__ double gram_total(const double *__restrict__ part, int e) {
    double v[kGramBlocks];
#pragma unroll
    for (int b = 0; b < kGramBlocks; b++) v[b] = __ldcg(part + b * 600 + e);
    double t = 0.0;
#pragma unroll
    for (int b = 0; b < kGramBlocks; b++) t += v[b];
    return t;
}

// block 0 of stage C (256 threads): the statistic rows 0..31 of the three layer-1 images
__device__ void pack_l1c_stat_rows(char *__restrict__ img_f16, char *__restrict__ img_bf16, const double *__restrict__ part) {
    __shared__ double sG[24][25], sL[24][25], sS[24];
    const int t = threadIdx.x;
    if (t < 24) sS[t] = gram_total(part, 576 + t);
    double g[3];
#pragma unroll
    for (int q = 0; q < 3; q++) { const int e = t + 256 * q; g[q] = e < 576 ? gram_total(part, e) : 0.0; }
    __syncthreads();
#pragma unroll
    for (int q = 0; q < 3; q++) {
        const int e = t + 256 * q;
        if (e < 576) { const int i = e / 24, j = e - i * 24; sG[i][j] = g[q] - sS[i] * sS[j] / H1; sL[i][j] = 0.0; }
    }
    __syncthreads();
    if (t < 32) {
        // right-looking Cholesky of the lower triangle by ONE warp: lane i owns row i (registers) and its diagonal element as a
        // scalar; column k of L travels through shared memory (broadcast reads).  The dependent chain of a step is only
        // shuffle(diagonal) -> rsqrt -> L[i][k] -> own diagonal update; the column broadcast runs beside the next pivot's rsqrt.
        // Tiny pivots -> zero column.
        const int i = t < 24 ? t : 23;
        double row[24], gm = 0.0;
#pragma unroll
        for (int j = 0; j < 24; j++) { row[j] = sG[i][j]; gm = fmax(gm, sG[j][j]); }
        double diag = sG[i][i];
        const double tiny = 1e-13 * gm;
#pragma unroll
        for (int k = 0; k < 24; k++) {
            const double d = __shfl_sync(0xffffffffu, diag, k);                   // G[k][k] as updated so far
            const bool ok = d > tiny && d > 0.0;
            double pinv = 0.0;
            if (ok) {                                                              // 1 / sqrt(d): float seed + two Newton steps (1e-15)
                const double h = 0.5 * d;
                pinv = (double)rsqrtf((float)d);
                pinv = pinv * fma(-h * pinv, pinv, 1.5);
                pinv = pinv * fma(-h * pinv, pinv, 1.5);
            }
            const double lik = i == k ? d * pinv : row[k] * pinv;                  // L[i][k], i >= k
            if (i > k) diag = fma(-lik, lik, diag);
            if (t < 24 && i >= k) sL[i][k] = lik;
            __syncwarp();
#pragma unroll
            for (int j = k + 1; j < 24; j++) { const double ljk = sL[j][k]; if (j < i) row[j] = fma(-lik, ljk, row[j]); }
        }
    }
    __syncthreads();
    for (int v = t; v < kStatRows * 32; v += 256) {
        const int rowi = v >> 5, k = v & 31;
        const float xf = (rowi < 24 && k < 24) ? (float)sL[k][rowi] : 0.f;        // row j of the statistic block = column j of L
        const __half hi = __float2half_rn(xf);
        *reinterpret_cast<__half *>(img_f16 + sw64_off(rowi, k)) = hi;
        *reinterpret_cast<__half *>(img_f16 + N1I * kRowB + sw64_off(rowi, k)) = __float2half_rn(xf - __half2float(hi));
        *reinterpret_cast<__nv_bfloat16 *>(img_bf16 + sw64_off(rowi, k)) = __float2bfloat16_rn(xf);
    }
}

// blocks 1 .. nblk - 1 of the image part of stage C: rows 32..431 = g1[c] (Wf[c][k] - m[k])
__device__ __forceinline__ void pack_l1c_image(int bid, int nblk, char *__restrict__ img_f16, char *__restrict__ img_bf16, const double *__restrict__ part,
                                               const float *__restrict__ fc1_w, const float *__restrict__ fc1_b, const float *__restrict__ g1) {
    __shared__ double m[24];
    if (threadIdx.x < 24) m[threadIdx.x] = gram_total(part, 576 + threadIdx.x) / H1;
    __syncthreads();
    for (int v = kStatRows * 32 + (bid - 1) * blockDim.x + threadIdx.x; v < N1I * 32; v += (nblk - 1) * blockDim.x) {
        const int row = v >> 5, k = v & 31;
        const double x = k < 24 ? (double)g1[row - kStatRows] * (wfull(fc1_w, fc1_b, row - kStatRows, k) - m[k]) : 0.0;
        const float xf = (float)x;
        const __half hi = __float2half_rn(xf);
        *reinterpret_cast<__half *>(img_f16 + sw64_off(row, k)) = hi;
        *reinterpret_cast<__half *>(img_f16 + N1I * kRowB + sw64_off(row, k)) = __float2half_rn(xf - __half2float(hi));
        *reinterpret_cast<__nv_bfloat16 *>(img_bf16 + sw64_off(row, k)) = __float2bfloat16_rn(xf);
    }
}

// ---- v4 layer-2 image: [sweep A: kb 0..12, 160 rows x 64 B][sweep B: kb 0..12, 144 rows x 64 B] ----
// row n of sweep A = output column n, of sweep B = output column 160 + n.  With Wf2 = [fc2.weight | fc2.bias] (k == H1: bias):
//   columns 0..299 : Wf2[n][k] - m2[k], m2 = mean over the 300 outputs: the GEMM delivers h - mean(h), i.e. LayerNorm 2's
//                    centring costs nothing and its variance is a plain sum of squares;
//   columns 300,301: hi / lo halves of l2[k] = 1/2 sum_n (Wf2[n][k] - m2[k]) g2[n] w3[n].  relu(y) = (y + |y|) / 2, and
//                    sum_n w3[n] y[n] / 2 is linear in the layer-2 input: rstd * (a2 . l2) + 1/2 sum be2 w3 -- the tensor core
//                    computes it in two of the four padding columns and the epilogue only accumulates |y[n]| w3[n] / 2.
__device__ __forceinline__ float wf2(const float *__restrict__ fc2_w, const float *__restrict__ fc2_b, int n, int k) {
    return k < H1 ? fc2_w[n * H1 + k] : (k == H1 ? fc2_b[n] : 0.f);
}
constexpr int kStatSlices = 32;
__device__ __forceinline__ void pack_w2_colstats(int bid, double *__restrict__ out /* m2[K2P], l2[K2P] */, const float *__restrict__ fc2_w,
                                                 const float *__restrict__ fc2_b, const float *__restrict__ g2, const float *__restrict__ w3) {
    __shared__ double r1[kStatSlices][33], r2[kStatSlices][33], rg[kStatSlices][33];
    const int kk = threadIdx.x & 31, sl = threadIdx.x >> 5, k = bid * 32 + kk;
    double s1 = 0.0, s2 = 0.0, sg = 0.0;
    constexpr int kIt = (H2 + kStatSlices - 1) / kStatSlices;
    float wv[kIt], gv[kIt], qv[kIt];                                  // every load first: one memory round trip, not kIt dependent ones
#pragma unroll
    for (int it = 0; it < kIt; it++) {
        const int n = sl + it * kStatSlices;
        const bool in = n < H2;
        wv[it] = in ? wf2(fc2_w, fc2_b, n, k) : 0.f; gv[it] = in ? g2[n] : 0.f; qv[it] = in ? w3[n] : 0.f;
    }
#pragma unroll
    for (int it = 0; it < kIt; it++) {
        const double w = (double)wv[it], gw = (double)gv[it] * (double)qv[it];
        s1 += w; s2 += w * gw; sg += gw;
    }
    r1[sl][kk] = s1; r2[sl][kk] = s2; rg[sl][kk] = sg;
    __syncthreads();
    if (sl == 0) {
#pragma unroll 8
        for (int i = 1; i < kStatSlices; i++) { s1 += r1[i][kk]; s2 += r2[i][kk]; sg += rg[i][kk]; }
        const double m = s1 / H2;
        out[k] = m;
        out[K2P + k] = 0.5 * (s2 - m * sg);
    }
}
template <typename OpT>
__device__ __forceinline__ void pack_w2s(int bid, int nblk, char *__restrict__ img, const double *__restrict__ st, const float *__restrict__ fc2_w,
                                         const float *__restrict__ fc2_b) {
    const int tid = bid * blockDim.x + threadIdx.x, nth = nblk * blockDim.x;
    for (int v = tid; v < KB2 * N2 * 32; v += nth) {
        const int kb = v / (N2 * 32), rem = v - kb * N2 * 32, col = rem / 32, kk = rem - col * 32, k = kb * 32 + kk;
        float x = 0.f;
        if (col < H2) x = (float)((double)wf2(fc2_w, fc2_b, col, k) - st[k]);
        else if (col == H2) x = (float)st[K2P + k];
        else if (col == H2 + 1) { const float l = (float)st[K2P + k]; x = l - op_to_float(to_op<OpT>(l)); }
        const size_t off = col < kNA ? (size_t)kb * kNA * kRowB + sw64_off(col, kk)
                                     : kW2SweepB + (size_t)kb * kNB * kRowB + sw64_off(col - kNA, kk);
        const OpT o = to_op<OpT>(x);
#pragma unroll
        for (int rep = 0; rep < TT_W2_REPLICAS; rep++) *reinterpret_cast<OpT *>(img + (size_t)rep * kW2ImageB + off) = o;
    }
}

// fp32 image of the CUDA-core actor kernel (tt_actor.cuh): transposed, zero-padded weights + parameter vectors
__device__ __forceinline__ void pack_fp32(int bid, int nblk, const tt_actor_dev &A, const float *fc1_w, const float *fc1_b, const float *g1, const float *be1,
                                          const float *fc2_w, const float *fc2_b, const float *g2, const float *be2, const float *mu_w, const float *mu_b) {
    const int tid = bid * blockDim.x + threadIdx.x, nth = nblk * blockDim.x;
    for (int v = tid; v < A.k1p * A.h1p; v += nth) {                    // W1T[k][c] = fc1.weight[c][k]
        const int k = v / A.h1p, c = v - k * A.h1p;
        A.w1t[v] = (k < A.in_dim && c < A.h1) ? fc1_w[c * A.in_dim + k] : 0.0f;
    }
    for (int v = tid; v < A.h1p * A.h2p; v += nth) {                    // W2T[k][c] = fc2.weight[c][k]
        const int k = v / A.h2p, c = v - k * A.h2p;
        A.w2t[v] = (k < A.h1 && c < A.h2) ? fc2_w[c * A.h1 + k] : 0.0f;
    }
    for (int c = tid; c < A.h1p; c += nth) {
        const bool in = c < A.h1;
        A.b1[c] = in ? fc1_b[c] : 0.0f; A.g1[c] = in ? g1[c] : 0.0f; A.be1[c] = in ? be1[c] : 0.0f;
    }
    for (int c = tid; c < A.h2p; c += nth) {
        const bool in = c < A.h2;
        A.b2[c] = in ? fc2_b[c] : 0.0f; A.g2[c] = in ? g2[c] : 0.0f; A.be2[c] = in ? be2[c] : 0.0f;
        A.w3[c] = in ? mu_w[c] : 0.0f;
    }
    if (tid == 0) A.b3[0] = mu_b[0];
}

// The re-pack of a policy (tt_actor_load) is on the end-to-end path: a learner hands over new weights every iteration.  Two
// launches instead of seven: stage A = everything that only reads the raw weights (Gram partial sums of layer 1, column
// statistics of layer 2, the fp32 images), stage C = the three tensor-core operand images (block 0: the Cholesky factor).
struct PackSrc { const float *fc1_w, *fc1_b, *g1, *be1, *fc2_w, *fc2_b, *g2, *be2, *mu_w, *mu_b; };
constexpr int kPackFp32Blocks = 16;
__global__ void __launch_bounds__(1024) pack_stage_a_kernel(tt_actor_dev A, PackSrc w, int tc) {
    chain_enter();
    int bid = blockIdx.x;
    if (tc) {
        if (bid < kGramBlocks) { pack_l1c_gram(bid, A.l1c_scratch + kGramOff, w.fc1_w, w.fc1_b); return; }
        bid -= kGramBlocks;
        if (bid < KB2) { pack_w2_colstats(bid, A.l1c_scratch + 1200, w.fc2_w, w.fc2_b, w.g2, w.mu_w); return; }
        bid -= KB2;
    }
    pack_fp32(bid, kPackFp32Blocks, A, w.fc1_w, w.fc1_b, w.g1, w.be1, w.fc2_w, w.fc2_b, w.g2, w.be2, w.mu_w, w.mu_b);
}
constexpr int kImgBlocks = 54, kW2sBlocks = 128;
__global__ void __launch_bounds__(256) pack_stage_c_kernel(tt_actor_dev A, PackSrc w) {
    chain_enter();
    int bid = blockIdx.x;
    if (bid == 0) { pack_l1c_stat_rows(reinterpret_cast<char *>(A.w1c_f16), reinterpret_cast<char *>(A.w1c_bf16), A.l1c_scratch + kGramOff); return; }
    if (bid < kImgBlocks) {
        pack_l1c_image(bid, kImgBlocks, reinterpret_cast<char *>(A.w1c_f16), reinterpret_cast<char *>(A.w1c_bf16), A.l1c_scratch + kGramOff, w.fc1_w, w.fc1_b, w.g1);
        return;
    }
    bid -= kImgBlocks;
    if (bid < kW2sBlocks) pack_w2s<__half>(bid, kW2sBlocks, reinterpret_cast<char *>(A.w2s_f16), A.l1c_scratch + 1200, w.fc2_w, w.fc2_b);
    else pack_w2s<__nv_bfloat16>(bid - kW2sBlocks, kW2sBlocks, reinterpret_cast<char *>(A.w2s_bf16), A.l1c_scratch + 1200, w.fc2_w, w.fc2_b);
}

template <bool kSplit>
struct Plan4 {
    static constexpr int kXBlocks = kSplit ? 2 : 1;                               // operand blocks of layer 1: hi [+ lo residual]
    static constexpr int kSlots = 6;
    static constexpr uint32_t kW2Slot = kW2SlotB;
    static constexpr uint32_t w1 = 0;                                             // (the observation tile, layer 1's A operand, lives in TMEM)
    static constexpr uint32_t a2 = w1 + kXBlocks * N1I * kRowB;
    static constexpr uint32_t w2 = a2 + KB2 * kTileM * kRowB;
    static constexpr uint32_t par = w2 + kSlots * kW2Slot;
    static constexpr uint32_t npar = K2P + 2 * H2P;                               // be1 | be2 / g2, w3 |g2| / 2
    static constexpr uint32_t red = par + npar * 4;
    static constexpr uint32_t bars = red + 2 * 4 * kTileM * 4;
    static constexpr uint32_t nbars = 40;
    static constexpr uint32_t tmem_slot = bars + nbars * 8;
    static constexpr uint32_t total = tmem_slot + 16;                             // the dynamic shared memory window is 1024 B aligned (checked)
};

enum { D_W1 = 0, D_XFULL, D_WFULL, D_WFREE, D_A2FULL, D_H2AFULL, D_H2BFULL, D_H2AFREE, D_H2BFREE,
       D_W2FULL, D_W2EMPTY = D_W2FULL + 6, D_A2FREE = D_W2EMPTY + 6, D_NOISE = D_A2FREE + KB2, D_COUNT = D_NOISE + 4 };
static_assert(D_COUNT <= 40, "barrier table");

// kProf: per-phase cycle counters of block 0 (profiles/tc_phase_profile.py); compiled out of the production instantiation
// (they cost ~25 registers per thread).
// kCluster: CTAs per cluster (1 or 2).  With 2, the two CTAs of a pair work on different row tiles but share the W2 stream:
// each fetches half of every k-block and multicasts it into both shared memories (half the L2 -> SM traffic, which bounds the
// sweeps); a ring slot is recycled when BOTH CTAs' MMAs have read it (multicast tcgen05.commit, mbarrier count 2).  The pair
// runs the same number of ring iterations; a CTA whose last tile does not exist streams and releases without computing.
template <typename OpT, bool kSplit, bool kProf, int kCluster>
__global__ void __launch_bounds__(640, 1) actor_tc4_kernel(const char *__restrict__ w1img, const char *__restrict__ w2img,
                                                           tt_actor_dev A, const float *__restrict__ obs, int64_t ld, int64_t n,
                                                           float *__restrict__ out, TTRingS ring, TTActorTail tail,
                                                           unsigned long long *__restrict__ dbg) {
    using P = Plan4<kSplit>;
    constexpr int kGroups = 4, kEpiThreads = 512, kThreads = 640;
    constexpr int kM2Warp = 16, kProdWarp = 17, kM1Warp = 18, kCopyWarp = 19;
    constexpr uint32_t kFmt = std::is_same<OpT, __nv_bfloat16>::value ? 1u : 0u;
    extern __shared__ uint8_t smem_raw[];
    const uint32_t base = smem_u32(smem_raw);
    if (base & 1023u) __trap();                      // SWIZZLE_64B operand blocks need 1024 B alignment; there is no static shared memory in front
    uint8_t *sm = smem_raw;
    const uint32_t sW1 = base + P::w1, sA2 = base + P::a2, sW2 = base + P::w2, sBar = base + P::bars;
    float *par = reinterpret_cast<float *>(sm + P::par);
    float *pbe1 = par, *pbe2 = par + K2P, *pw3 = pbe2 + H2P;
    float *red1 = reinterpret_cast<float *>(sm + P::red);          // per (column group, row): sum of squares | output dot
    float *red3 = red1 + 4 * kTileM;
    volatile uint32_t *tmem_slot = reinterpret_cast<volatile uint32_t *>(sm + P::tmem_slot);
    volatile uint32_t *staged = tmem_slot + 2;           // number of observation tiles the epilogue has staged so far
    auto bar = [&](int i) { return sBar + 8u * (uint32_t)i; };

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int64_t ntiles = (n + kTileM - 1) / kTileM;
    // ring iterations of this CTA (pair): the lower-ranked CTA of a pair never has fewer tiles than the other
    const int64_t pair_first = kCluster > 1 ? (int64_t)(blockIdx.x - cluster_ctarank()) : (int64_t)blockIdx.x;
    const int64_t niter = pair_first < ntiles ? (ntiles - pair_first + gridDim.x - 1) / gridDim.x : 0;
    constexpr uint16_t kMask = (uint16_t)((1u << kCluster) - 1u);

    // ---------------- one-time setup ----------------
    if (threadIdx.x == 0) {
        *staged = 0u;
        mbar_init(bar(D_W1), 1);
        mbar_init(bar(D_XFULL), kEpiThreads); mbar_init(bar(D_WFULL), 1); mbar_init(bar(D_WFREE), kEpiThreads);
        mbar_init(bar(D_A2FULL), kEpiThreads);
        mbar_init(bar(D_H2AFULL), 1); mbar_init(bar(D_H2BFULL), 1); mbar_init(bar(D_H2AFREE), kEpiThreads); mbar_init(bar(D_H2BFREE), kEpiThreads);
        for (int i = 0; i < P::kSlots; i++) { mbar_init(bar(D_W2FULL + i), 1); mbar_init(bar(D_W2EMPTY + i), kCluster); }
        for (int i = 0; i < KB2; i++) mbar_init(bar(D_A2FREE + i), 1);
        for (int i = 0; i < 4; i++) mbar_init(bar(D_NOISE + i), 32);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == kM2Warp) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(base + P::tmem_slot), "r"(kTmemCols) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    chain_wait();                                    // (tt_common.cuh: chained launches) the setup above touched no global memory
    chain_trigger();
    // The first tile's observation loads (epilogue threads: 8 inputs of the thread's own row) and the loads of the output constant
    // are issued HERE, in front of the parameter loops and the set-up barrier, so that their memory round trips overlap with
    // everything below instead of following it (a batch of one tile per CTA is all latency: 9.3 -> 8.8 us per launch in a CUDA graph).
    float xreg0[8];
    if (warp < 16) {
        const int grp0 = warp >> 2, r0 = (warp & 3) * 32 + lane, kk0 = 8 * grp0;
        const int64_t gr = (int64_t)blockIdx.x * kTileM + r0;
        const bool ok = gr < n && grp0 < 3;
        const float *p = obs + gr * ld + kk0;
#pragma unroll
        for (int i = 0; i < 8; i++) xreg0[i] = (kk0 + i < IN) ? (ok ? __ldg(p + i) : 0.f) : (kk0 + i == IN ? 1.0f : 0.f);
    }
    // constant part of the output: mu.bias + 1/2 sum be2 w3 (every warp computes it, once).  All 30 loads of a lane are issued
    // before the first use: nvcc otherwise keeps load -> FMA order, ten dependent L2 round trips (2.6 us) in front of the first tile.
    float b3 = 0.f;
    {
        constexpr int kIt = (H2 + 31) / 32;
        float be[kIt], w[kIt], g[kIt];
#pragma unroll
        for (int i = 0; i < kIt; i++) {
            const int c = lane + 32 * i;
            const bool in = c < H2;
            be[i] = in ? __ldg(A.be2 + c) : 0.f; w[i] = in ? __ldg(A.w3 + c) : 0.f; g[i] = in ? __ldg(A.g2 + c) : 1.f;
        }
        const float mub = __ldg(A.b3);
#pragma unroll
        for (int i = 0; i < kIt; i++) b3 = fmaf(fabsf(g[i]) > 1e-30f ? 0.5f * be[i] : fmaxf(be[i], 0.f), w[i], b3);
#pragma unroll
        for (int o = 16; o; o >>= 1) b3 += __shfl_xor_sync(0xffffffffu, b3, o);
        b3 += mub;
    }

    for (int c = threadIdx.x; c < K2P; c += kThreads) pbe1[c] = c < H1 ? A.be1[c] : 0.f;
    for (int c = threadIdx.x; c < H2P; c += kThreads) {
        // LayerNorm 2 + ReLU + mu:  w3 relu(g z + be), z = (h - mean) rstd.  relu(y) = (y + |y|) / 2; the y / 2 half is linear in
        // the layer-2 input and comes out of the GEMM (see the pack); |g z + be| w3 / 2 = |z + be / g| (w3 |g| / 2).  A column
        // with g == 0 is the constant relu(be) w3: part of b3 below.
        const bool in = c < H2 && fabsf(A.g2[c]) > 1e-30f;
        pbe2[c] = in ? A.be2[c] / A.g2[c] : 0.f; pw3[c] = in ? 0.5f * A.w3[c] * fabsf(A.g2[c]) : 0.f;
    }
    // X blocks: zero once (k = 24..31 stay zero).  A2 block 12, columns 400..415: constant (1, 0, ..., 0) -- column 400
    // carries the fc2 bias -- written once; the epilogue only ever rewrites columns 384..399 of that block.
    if (threadIdx.x < kTileM) {
        const int r = threadIdx.x;
        const uint32_t xsw = ((uint32_t)r >> 1) & 3u;
        uint8_t *blk = sm + P::a2 + (size_t)12 * kTileM * kRowB + (size_t)r * kRowB;
        uint4 one = make_uint4(0u, 0u, 0u, 0u), zero = make_uint4(0u, 0u, 0u, 0u);
        const OpT o = to_op<OpT>(1.0f);
        one.x = (uint32_t)(*reinterpret_cast<const uint16_t *>(&o));
        *reinterpret_cast<uint4 *>(blk + ((2u ^ xsw) << 4)) = one;
        *reinterpret_cast<uint4 *>(blk + ((3u ^ xsw) << 4)) = zero;
    }
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    if (kCluster > 1) cluster_sync_all();            // the partner's mbarriers are initialised before anything multicasts into them
    tc_fence_after();
    const uint32_t tmem = *tmem_slot;
    if (warp == kProdWarp) {
        // ================= bulk-copy producer: W1 once, then per tile 13 half k-blocks of sweep A and 13 of sweep B =================
        // (warp-uniform control flow; one elected lane issues the copies)
        constexpr uint32_t w1bytes = (uint32_t)P::kXBlocks * N1I * kRowB;
        if (elect_one()) {
            mbar_expect_tx(bar(D_W1), w1bytes);
            bulk_g2s(sW1, w1img, w1bytes, bar(D_W1));
        }
        __syncwarp();
        const char *w2rep = w2img + (size_t)(blockIdx.x % TT_W2_REPLICAS) * kW2ImageB;     // this SM's replica
        const uint32_t crank = kCluster > 1 ? cluster_ctarank() : 0u;
        uint32_t tp = 0;
        for (int64_t itn = 0; itn < ((TT_ABLATE & 1) ? 0 : niter); itn++, tp ^= 1u) {
#pragma unroll
            for (int step = 0; step < kSteps; step++) {
                const int sweep = w2_step_sweep(step), kb = w2_step_kb(step), slot = w2_slot(step, P::kSlots);
                const uint32_t bytes = (uint32_t)(sweep ? kNB : kNA) * kRowB;
                mbar_wait(bar(D_W2EMPTY + slot), w2_parity(step, P::kSlots, tp) ^ 1u);     // every CTA of the cluster has read the slot
                if (elect_one()) {
                    mbar_expect_tx(bar(D_W2FULL + slot), bytes);
                    const char *src = w2rep + (sweep ? kW2SweepB : 0) + (size_t)kb * bytes;
                    if (kCluster == 1) bulk_g2s(sW2 + slot * P::kW2Slot, src, bytes, bar(D_W2FULL + slot));
                    else {                                                                  // my share of the block, to every CTA
                        const uint32_t part = bytes / kCluster, off = crank * part;
                        bulk_g2s_mc(sW2 + slot * P::kW2Slot + off, src + off, part, bar(D_W2FULL + slot), kMask);
                    }
                }
                __syncwarp();
            }
        }
    } else if (warp == kCopyWarp) {
        // ================= fused replay store of s: observation rows -> ring `state` rows =================
        // It follows the epilogue's staging counter, so its reads of a tile come right after the epilogue's own prefetch of
        // the same rows and hit L2 (unsynchronised, this warp ran far ahead and every observation row was fetched from
        // DRAM twice: 743 MB instead of 386 MB per launch at N = 2^22).
        // ... and the Ornstein-Uhlenbeck noise of the tile's rows (DDPG/noise.py:12-17): x <- x + theta (0 - x) dt + sigma
        // sqrt(dt) N(0, 1), four rows per lane, written back to the OU state array; the output stage of the epilogue (group 0)
        // picks x up from there (D_NOISE: one of four mbarriers per tile, this warp is never more than three tiles ahead of
        // the output stage because it follows the staging counter) and adds it to mu (DDPG_agent.py:41-43).  The Philox
        // rounds run on this otherwise idle warp instead of on the epilogue's critical path.
        if (ring.S || tail.ou_x) {
            uint32_t j = 0;
            const uint32_t titer = tail.ou_x ? *tail.iter : 0u;
            for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x, j++) {
                while (*staged <= j) __nanosleep(200);
                const int64_t row0 = tile * kTileM;
                const int rows = (int)((n - row0) < kTileM ? (n - row0) : kTileM);
                if (tail.ou_x) {
                    float xv[4];
#pragma unroll
                    for (int i = 0; i < 4; i++) { const int rr = lane + 32 * i; xv[i] = rr < rows ? __ldcg(tail.ou_x + row0 + rr) : 0.f; }
#pragma unroll
                    for (int i = 0; i < 4; i++) {
                        const int rr = lane + 32 * i;
                        if (rr < rows) __stcg(tail.ou_x + row0 + rr, ttm::ou_advance(xv[i], ttm::rng_normal_ks(tail.keys, tail.gid0 + (uint32_t)(row0 + rr), titer)));
                    }
                    mbar_arrive(bar(D_NOISE + (j & 3u)));              // release: this lane's stores are visible to the waiting epilogue threads
                }
                if (!ring.S) continue;
                const int64_t rrow0 = ring.m.row(row0);
                const float *src = obs + row0 * ld;
                float *dst = ring.S + rrow0 * IN;
                const bool flat = ld == IN && rows == kTileM && !ring.m.many && row0 >= ring.m.first && rrow0 + kTileM <= ring.m.cap;
                if (flat && (((uintptr_t)src | (uintptr_t)dst) & 15) == 0) {
                    // the tile (128 x 23 floats = 23 x 32 float4) in two rounds of 12 / 11 loads in flight per lane: two memory round trips.
                    // (4 loads in flight per lane made this warp latency-bound at ~17 k cycles per tile -- slower than
                    // the GEMMs it is supposed to hide under.)
                    const float4 *s4 = reinterpret_cast<const float4 *>(src);
                    float4 *d4 = reinterpret_cast<float4 *>(dst);
                    constexpr int kV = kTileM * IN / 4 / 32, kHalf = (kV + 1) / 2;      // 23 float4 per lane, in two rounds of 12 + 11
                    float4 buf[kHalf];
#pragma unroll
                    for (int h = 0; h < 2; h++) {
#pragma unroll
                        for (int i = 0; i < kHalf; i++) if (h * kHalf + i < kV) buf[i] = __ldcs(&s4[lane + 32 * (h * kHalf + i)]);
#pragma unroll
                        for (int i = 0; i < kHalf; i++) if (h * kHalf + i < kV) __stcs(&d4[lane + 32 * (h * kHalf + i)], buf[i]);
                    }
                } else if (flat) {
#pragma unroll 4
                    for (int v = lane; v < kTileM * IN; v += 32) __stcs(&dst[v], __ldcs(&src[v]));
                } else {
                    for (int v = lane; v < rows * IN; v += 32) {
                        const int rr = v / IN, k = v - rr * IN;
                        if (row0 + rr >= ring.m.first) ring.S[ring.m.row(row0 + rr) * IN + k] = __ldcs(obs + (row0 + rr) * ld + k);
                    }
                }
            }
        }
    } else if (warp == kM1Warp) {
        // ================= layer-1 MMA issuer: 3 parts through the window =================
        {
            const uint32_t idp[kParts] = {make_idesc(128, kFmt), make_idesc(160, kFmt), make_idesc(144, kFmt)};
            constexpr uint32_t rowp[kParts] = {0u, 128u, 288u};            // first image row of each part
            constexpr int npairs = kSplit ? 3 : 1;
            const uint64_t dW1 = make_desc(sW1);                          // descriptor address field is in 16 B units
            uint32_t c1 = 0, use = 0;
            mbar_wait(bar(D_W1), 0);
            for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x, c1++) {
                mbar_wait(bar(D_XFULL), c1 & 1u);
#pragma unroll
                for (int p = 0; p < kParts; p++, use++) {
                    mbar_wait(bar(D_WFREE), (use & 1u) ^ 1u);              // the epilogue has pulled the previous part out of the window
                    tc_fence_after();
                    if (elect_one()) {
#pragma unroll
                        for (int pr = 0; pr < npairs; pr++) {
                            const uint32_t xa = tmem + kXCol0 + (pr == 1 ? 16u : 0u);                 // hi . W1hi + lo . W1hi + hi . W1lo
                            const uint64_t wb = dW1 + (uint64_t)((pr == 2 ? N1I * kRowB : 0) + rowp[p] * kRowB) / 16;
                            umma_ts(tmem + kWin0, xa, wb, idp[p], pr ? 1u : 0u);
                            umma_ts(tmem + kWin0, xa + 8u, wb + 2, idp[p], 1u);                      // next 16 inputs: 8 TMEM columns
                        }
                        umma_commit(bar(D_WFULL));
                    }
                    __syncwarp();
                }
            }
        }
    } else if (warp == kM2Warp) {
        // ================= layer-2 MMA issuer: sweep A (columns 0..159), then sweep B (160..303) =================
        {
            const uint32_t idA = make_idesc(kNA, kFmt), idB = make_idesc(kNB, kFmt);
            const uint64_t dA2 = make_desc(sA2), dW2 = make_desc(sW2);    // descriptor address field is in 16 B units
            constexpr bool prof = kProf;
            uint32_t c2 = 0;
            long long t_a2 = 0, t_w2 = 0, t_h2 = 0, t0 = 0;
            for (int64_t itn = 0; itn < niter; itn++, c2++) {
                const uint32_t ph = c2 & 1u;
                if (kCluster > 1 && (int64_t)blockIdx.x + itn * gridDim.x >= ntiles) {
                    // no tile of its own in the pair's last iteration: keep the shared W2 ring moving
#pragma unroll
                    for (int step = 0; step < kSteps; step++) {
                        const int slot = w2_slot(step, P::kSlots);
                        if (!(TT_ABLATE & 1)) mbar_wait(bar(D_W2FULL + slot), w2_parity(step, P::kSlots, ph));
                        if (elect_one()) umma_commit_mc(bar(D_W2EMPTY + slot), kMask);
                        __syncwarp();
                    }
                    continue;
                }
                if (prof) t0 = clock64();
                mbar_wait(bar(D_A2FULL), ph);
                if (prof) t_a2 += clock64() - t0;
#pragma unroll
                for (int step = 0; step < kSteps; step++) {
                    const int sweep = w2_step_sweep(step), kb = w2_step_kb(step), slot = w2_slot(step, P::kSlots);
                    if (kb == 0) {
                        if (prof) t0 = clock64();
                        if (!(TT_ABLATE & 16)) mbar_wait(bar(sweep ? D_H2BFREE : D_H2AFREE), ph ^ 1u);   // pass 2 of the previous tile has read this half
                        if (prof) t_h2 += clock64() - t0;
                    }
                    if (prof) t0 = clock64();
                    if (!(TT_ABLATE & 1)) mbar_wait(bar(D_W2FULL + slot), w2_parity(step, P::kSlots, ph));
                    if (prof) t_w2 += clock64() - t0;
                    tc_fence_after();
                    if (elect_one()) {
                        const uint64_t a = dA2 + (uint64_t)(kb * (kTileM * kRowB / 16)), b = dW2 + (uint64_t)(slot * (P::kW2Slot / 16));
                        const uint32_t d = tmem + (sweep ? kNA : 0);
                        umma(d, a, b, sweep ? idB : idA, kb ? 1u : 0u);
                        umma(d, a + 2, b + 2, sweep ? idB : idA, 1u);
                        if (kCluster == 1) umma_commit(bar(D_W2EMPTY + slot)); else umma_commit_mc(bar(D_W2EMPTY + slot), kMask);
                        if (sweep && ((kA2Waits >> kb) & 1u)) umma_commit(bar(D_A2FREE + kb));   // A2 blocks <= kb may be overwritten (epilogue 1: a2wait)
                        if (kb == KB2 - 1) umma_commit(bar(sweep ? D_H2BFULL : D_H2AFULL));
                    }
                    __syncwarp();
                }
            }
            if (prof && blockIdx.x == 0 && lane == 0) { dbg[0] = t_h2; dbg[1] = t_a2; dbg[2] = t_w2; dbg[3] = c2; }
        }
    } else {
        // ================= epilogue warps: thread = (row, column group) =================
        const int grp = warp >> 2;
        const int r = (warp & 3) * 32 + lane;
        const int et = threadIdx.x;
        const uint32_t trow = tmem + ((uint32_t)((warp & 3) * 32) << 16);
        const uint32_t xsw = (((uint32_t)r >> 1) & 3u);
        // X staging role: this thread's own row (its TMEM lane), inputs 8 grp .. 8 grp + 7; k = 23 is the constant-1 bias column,
        // k = 24..31 (group 3) are zero
        const int k0 = 8 * grp;
        float xreg[8];
        auto load_x = [&](int64_t t) {
            const int64_t gr = t * kTileM + r;
            const bool ok = gr < n && grp < 3;
            const float *p = obs + gr * ld + k0;
#pragma unroll
            for (int i = 0; i < 8; i++) xreg[i] = (k0 + i < IN) ? (ok ? __ldg(p + i) : 0.f) : (k0 + i == IN ? 1.0f : 0.f);
        };
        constexpr bool prof = kProf;
        long long e_wf = 0, e_af = 0, e_b1 = 0, e_b2 = 0, tw = 0, e_st = 0, e_1 = 0, e_wa = 0, e_pa = 0, e_wb = 0, e_2 = 0, e_2a = 0, e_2b = 0, e_2c = 0, e_2d = 0, t0 = 0, t1 = 0, t2 = 0;
        const long long t_begin = clock64();
        uint32_t wuse = 0;

        // stage the observation tile held in xreg as the layer-1 A operand (hi [+ lo residual]) and release it
        auto stage = [&]() {
            if (prof) t0 = clock64();
            uint32_t hi[4], lo[4];
#pragma unroll
            for (int i = 0; i < 4; i++) {
                const float a = xreg[2 * i], b = xreg[2 * i + 1];
                const OpT ah = to_op<OpT>(a), bh = to_op<OpT>(b);
                hi[i] = (uint32_t)(*reinterpret_cast<const uint16_t *>(&ah)) | ((uint32_t)(*reinterpret_cast<const uint16_t *>(&bh)) << 16);
                lo[i] = pack2<OpT>(a - op_to_float(ah), b - op_to_float(bh));
            }
            tmem_st4(trow + (uint32_t)(kXCol0 + 4 * grp), hi);             // TMEM column c of the block = inputs 2 c, 2 c + 1 of this lane's row
            if (kSplit) tmem_st4(trow + (uint32_t)(kXCol0 + 16 + 4 * grp), lo);
            tmem_wait_st();
            tc_fence_before();
            mbar_arrive(bar(D_XFULL));
            if (et == 0) *staged = *staged + 1u;                           // progress signal for the replay-store warp
            if (prof) { t1 = clock64(); e_st += t1 - t0; }
        };

        // layer-2 side of one tile: LayerNorm + ReLU over H2, dot with mu.weight, tanh.  Column group g owns columns
        // [40 g, 40 g + 40) of half A and [160 + 36 g, 160 + 36 g + 36) of half B.
        const int ca = 40 * grp, cbb = kNA + 36 * grp;
        float2 q2;                                                         // the accumulators are centred (see the pack): variance = sum of squares
        auto acc = [&](const uint32_t *v, int cnt) {
            if (TT_ABLATE & 8) return;
#pragma unroll
            for (int j = 0; j < cnt / 2; j++) {
                const float2 x = make_float2(__uint_as_float(v[2 * j]), __uint_as_float(v[2 * j + 1]));
                q2 = __ffma2_rn(x, x, q2);
            }
        };
        auto pass1a = [&](uint32_t c2) {                                   // statistics of half A (runs under sweep B)
            if (prof) t0 = clock64();
            mbar_wait(bar(D_H2AFULL), c2 & 1u);
            if (prof) { t1 = clock64(); e_wa += t1 - t0; t0 = t1; }
            tc_fence_after();
            uint32_t va[32], vt[8];
            tmem_ld32_async(trow + (uint32_t)ca, va);
            tmem_ld8_async(trow + (uint32_t)(ca + 32), vt);
            tmem_wait();
            q2 = make_float2(0.f, 0.f);
            acc(va, 32); acc(vt, 8);
            if (prof) { t1 = clock64(); e_pa += t1 - t0; t0 = t1; }
        };

        // layer-1 side of one tile: statistics -> rstd, then the 3 parts -> A2.  c1 = layer-1 tile counter
        // `with_p1a`: run the statistics pass over half A of the layer-2 tile c2 between part 0 and part 1 -- it fills the
        // wait for the layer-1 MMAs of part 1, which queue behind the sweep-B MMAs.
        auto layer1 = [&](uint32_t c1, int p1a_pos, uint32_t c2) {
            const uint32_t ph = c1 & 1u;
            if (prof) t0 = clock64();
            float2 rstd2 = make_float2(0.f, 0.f);
            // A2 blocks are released in order (sweep B's commits), and an mbarrier wait costs ~90 cycles even when the phase is
            // already complete: wait once per run of blocks (kA2Waits), on the last block of the run.
            auto a2wait = [&](int ch) {
                if (prof) tw = clock64();
                mbar_wait(bar(D_A2FREE + ch), ph ^ 1u);                    // sweep B of the previous tile has read blocks <= ch
                if (prof) e_af += clock64() - tw;
            };
            float4 E0[2], E1[2];                                           // be1 of the next chunk, fetched one chunk ahead
            auto lde = [&](int slot, int ch) {
                if (TT_ABLATE & 4) return;
                E0[slot] = lds128_early(pbe1 + ch * 32 + grp * 8); E1[slot] = lds128_early(pbe1 + ch * 32 + grp * 8 + 4);
            };
            auto emit = [&](const uint32_t (&v)[8], int ch, int slot) {    // 8 columns of A2 block ch: relu(t' rstd + be1)
                if (TT_ABLATE & 4) return;
                const float4 e0 = E0[slot], e1 = E1[slot];
                const float2 ee[4] = {make_float2(e0.x, e0.y), make_float2(e0.z, e0.w), make_float2(e1.x, e1.y), make_float2(e1.z, e1.w)};
                uint32_t pk4[4];
#pragma unroll
                for (int j = 0; j < 4; j++) {
                    const float2 x = make_float2(__uint_as_float(v[2 * j]), __uint_as_float(v[2 * j + 1]));
                    const float2 y = __ffma2_rn(x, rstd2, ee[j]);
                    pk4[j] = pack2_relu<OpT>(y.x, y.y);
                }
                uint4 pk;
                pk.x = pk4[0]; pk.y = pk4[1]; pk.z = pk4[2]; pk.w = pk4[3];
                *reinterpret_cast<uint4 *>(sm + P::a2 + (size_t)ch * kTileM * kRowB + (size_t)r * kRowB + (((uint32_t)grp ^ xsw) << 4)) = pk;
            };
            const uint32_t wcol = trow + (uint32_t)(kWin0 + grp * 8);
            uint32_t v[5][8];
            {   // part 0: 32 statistic columns, then A2 blocks 0..2
                if (prof) tw = clock64();
                mbar_wait(bar(D_WFULL), wuse & 1u); wuse++;
                if (prof) e_wf += clock64() - tw;
                tc_fence_after();
                uint32_t sv[32];
                tmem_ld32_async(trow + kWin0, sv);
#pragma unroll
                for (int c = 0; c < 3; c++) tmem_ld8_async(wcol + 32 + 32 * c, v[c]);
                lde(0, 0);
                tmem_wait();
                tc_fence_before();
                mbar_arrive(bar(D_WFREE));                                 // values are in registers: the window may be refilled
                float2 q2 = make_float2(0.f, 0.f);
#pragma unroll
                for (int j = 0; j < 12; j++) {                             // columns 24..31 are zero rows of the image
                    const float2 x = make_float2(__uint_as_float(sv[2 * j]), __uint_as_float(sv[2 * j + 1]));
                    q2 = __ffma2_rn(x, x, q2);
                }
                const float rstd = rsqrtf((q2.x + q2.y) * (1.0f / H1) + 1e-5f);
                rstd2 = make_float2(rstd, rstd);
#pragma unroll
                for (int c = 0; c < 3; c++) {
                    if (c < 2) lde((c + 1) & 1, c + 1);
                    if (a2_run_start(c)) a2wait(a2_run_end(c));
                    emit(v[c], c, c & 1);
                }
            }
            if (p1a_pos == 0) { if (prof) { t1 = clock64(); e_1 += t1 - t0; } pass1a(c2); }
            {   // part 1: A2 blocks 3..7
                if (prof) tw = clock64();
                mbar_wait(bar(D_WFULL), wuse & 1u); wuse++;
                if (prof) e_wf += clock64() - tw;
                tc_fence_after();
#pragma unroll
                for (int c = 0; c < 5; c++) tmem_ld8_async(wcol + 32 * c, v[c]);
                lde(0, 3);
                tmem_wait();
                tc_fence_before();
                mbar_arrive(bar(D_WFREE));
#pragma unroll
                for (int c = 0; c < 5; c++) {
                    if (c < 4) lde((c + 1) & 1, 4 + c);
                    if (a2_run_start(3 + c)) a2wait(a2_run_end(3 + c));
                    emit(v[c], 3 + c, c & 1);
                }
            }
            if (p1a_pos == 1) { if (prof) { t1 = clock64(); e_1 += t1 - t0; } pass1a(c2); }
            {   // part 2: A2 blocks 8..11 and the 16 real columns of block 12 (groups 0, 1)
                if (prof) tw = clock64();
                mbar_wait(bar(D_WFULL), wuse & 1u); wuse++;
                if (prof) e_wf += clock64() - tw;
                tc_fence_after();
#pragma unroll
                for (int c = 0; c < 4; c++) tmem_ld8_async(wcol + 32 * c, v[c]);
                if (grp < 2) tmem_ld8_async(wcol + 128, v[4]);
                lde(0, 8);
                tmem_wait();
                tc_fence_before();
                mbar_arrive(bar(D_WFREE));
#pragma unroll
                for (int c = 0; c < 4; c++) {
                    if (c < 3) lde((c + 1) & 1, 9 + c); else if (grp < 2) lde(0, 12);
                    if (a2_run_start(8 + c)) a2wait(a2_run_end(8 + c));
                    emit(v[c], 8 + c, c & 1);
                }
                if (a2_run_start(12)) a2wait(12);                          // (all groups wait: A2FULL below covers block 12 as well)
                if (grp < 2) emit(v[4], 12, 0);
            }
            fence_proxy_async();
            mbar_arrive(bar(D_A2FULL));
            if (prof) { t1 = clock64(); e_1 += t1 - t0; t0 = t1; }
        };

        auto layer2 = [&](int64_t tile, uint32_t c2) {
            const uint32_t ph = c2 & 1u;
            const int64_t row0 = tile * kTileM;
            const int rows = (int)((n - row0) < kTileM ? (n - row0) : kTileM);
            if (prof) t0 = clock64();
            mbar_wait(bar(D_H2BFULL), ph);
            if (prof) { t1 = clock64(); e_wb += t1 - t0; t0 = t1; }
            tc_fence_after();
            uint32_t va[32], vt[8], vq[4];
            tmem_ld32_async(trow + (uint32_t)cbb, va);
            tmem_ld4_async(trow + (uint32_t)(cbb + 32), vq);
            tmem_wait();
            float lin = 0.f;                                               // columns 300, 301: the linear half of the output dot (hi + lo)
            if (grp == kGroups - 1) { lin = __uint_as_float(vq[0]) + __uint_as_float(vq[1]); vq[0] = 0u; vq[1] = 0u; }
            acc(va, 32); acc(vq, 4);
            if (prof) { t2 = clock64(); e_2a += t2 - t0; }
            red1[grp * kTileM + r] = q2.x + q2.y;
            // pass 2 re-reads the accumulators; half A's loads fly while the statistics are exchanged
            tmem_ld32_async(trow + (uint32_t)ca, va);
            tmem_ld8_async(trow + (uint32_t)(ca + 32), vt);
            if (prof) tw = clock64();
            named_bar_sync(1, kEpiThreads);
            if (prof) e_b1 += clock64() - tw;
            float sq = 0.f;
#pragma unroll
            for (int g = 0; g < kGroups; g++) sq += red1[g * kTileM + r];
            const float rstd = rsqrtf(sq * (1.0f / H2) + 1e-5f);
            const float2 rstd2 = make_float2(rstd, rstd);
            float2 dot2 = make_float2(rstd * lin, 0.f);
            // Pass 2 over this thread's 19 column quads (10 of half A, 9 of half B).  Parameters are fetched two quads ahead
            // (lds128_early) into a 2-slot register ring; half B's TMEM load is issued as soon as va is dead and flies under
            // the last two quads of half A.
            constexpr int kQA = 10, kQ = 19;
            float4 E[2], W[2];
            auto qcol = [&](int qi) { return qi < kQA ? ca + 4 * qi : cbb + 4 * (qi - kQA); };
            auto loadp = [&](int slot, int qi) {
                const int c = qcol(qi);
                if (TT_ABLATE & 2) return;
                E[slot] = lds128_early(pbe2 + c); W[slot] = lds128_early(pw3 + c);
            };
            auto quad = [&](int slot, uint32_t x0, uint32_t x1, uint32_t x2, uint32_t x3) {
                if (TT_ABLATE & 2) return;
                const float2 xa = make_float2(__uint_as_float(x0), __uint_as_float(x1)), xb = make_float2(__uint_as_float(x2), __uint_as_float(x3));
                // |x rstd + be / g| (w3 |g| / 2): two FFMA2 per column pair, the second with an |.| source modifier
                const float2 ya = __ffma2_rn(xa, rstd2, make_float2(E[slot].x, E[slot].y));
                const float2 yb = __ffma2_rn(xb, rstd2, make_float2(E[slot].z, E[slot].w));
                dot2 = __ffma2_rn(make_float2(fabsf(ya.x), fabsf(ya.y)), make_float2(W[slot].x, W[slot].y), dot2);
                dot2 = __ffma2_rn(make_float2(fabsf(yb.x), fabsf(yb.y)), make_float2(W[slot].z, W[slot].w), dot2);
            };
            loadp(0, 0); loadp(1, 1);
            tmem_wait();
            if (prof) { t1 = clock64(); e_2b += t1 - t2; t2 = t1; }
            tc_fence_before();
            mbar_arrive(bar(D_H2AFREE));                                   // half A is in registers: sweep A of the next tile may start
#pragma unroll
            for (int qi = 0; qi < 8; qi++) {                               // half A, columns ca .. ca + 31
                quad(qi & 1, va[4 * qi], va[4 * qi + 1], va[4 * qi + 2], va[4 * qi + 3]);
                loadp(qi & 1, qi + 2);
            }
            tmem_ld32_async(trow + (uint32_t)cbb, va);                     // half B flies under the last two quads of half A
            tmem_ld4_async(trow + (uint32_t)(cbb + 32), vq);
#pragma unroll
            for (int qi = 8; qi < kQA; qi++) {                             // half A, columns ca + 32 .. ca + 39
                quad(qi & 1, vt[4 * (qi - 8)], vt[4 * (qi - 8) + 1], vt[4 * (qi - 8) + 2], vt[4 * (qi - 8) + 3]);
                loadp(qi & 1, qi + 2);
            }
            tmem_wait();
            tc_fence_before();
            mbar_arrive(bar(D_H2BFREE));
            if (prof) { t1 = clock64(); e_2c += t1 - t2; t2 = t1; }
#pragma unroll
            for (int qi = kQA; qi < kQ - 1; qi++) {                        // half B, columns cbb .. cbb + 31
                quad(qi & 1, va[4 * (qi - kQA)], va[4 * (qi - kQA) + 1], va[4 * (qi - kQA) + 2], va[4 * (qi - kQA) + 3]);
                if (qi + 2 < kQ) loadp(qi & 1, qi + 2);
            }
            quad((kQ - 1) & 1, vq[0], vq[1], vq[2], vq[3]);                // columns cbb + 32 .. cbb + 35
            if (prof) { t1 = clock64(); e_2d += t1 - t2; t2 = t1; }
            red3[grp * kTileM + r] = dot2.x + dot2.y;
            float noise = 0.f;                                             // this row's OU state, advanced by the copy warp (see there)
            if (grp == 0 && tail.ou_x) {
                mbar_wait(bar(D_NOISE + (c2 & 3u)), (c2 >> 2) & 1u);
                if (r < rows) noise = __ldcg(tail.ou_x + row0 + r);
            }
            if (prof) tw = clock64();
            named_bar_sync(1, kEpiThreads);
            if (prof) e_b2 += clock64() - tw;
            if (grp == 0 && r < rows) {
                float d = b3;
#pragma unroll
                for (int g = 0; g < kGroups; g++) d += red3[g * kTileM + r];
                const float a = tanhf(d) + noise;                          // DDPG_agent.py:41-43: mu + noise, unclipped
                const int64_t row = row0 + r;
                out[row] = a;
                if (tail.scaled) tail.scaled[row] = fminf(fmaxf(a, -1.0f), 1.0f) * 0.78539819f;    // trainv2.py:516
                if (tail.ring.A && row >= tail.ring.m.first) tail.ring.A[tail.ring.m.row(row)] = a;  // agent.remember keeps the raw action
            }
            if (prof) { t1 = clock64(); e_2 += t1 - t0; t0 = t1; }
        };


        const int64_t first = blockIdx.x, G = gridDim.x;
        uint32_t c1 = 0, c2 = 0;
        if (first < ntiles) {
#pragma unroll
            for (int i = 0; i < 8; i++) xreg[i] = xreg0[i];                // tile `first`: loaded during the set-up
            stage(); load_x(first + G);                                    // X(0); registers <- tile 1
            layer1(c1++, -1, 0u);
            if (first + G < ntiles) { stage(); load_x(first + 2 * G); }    // X(1)
            int64_t prev = first;
            for (;;) {
                const int64_t next = prev + G;
                const bool has_next = next < ntiles;
                if (has_next) {
                    // runs under the layer-2 MMAs of `prev`, block by block as they release A2.  The statistics pass over half A
                    // of `prev` (complete at step 26 - kLead) goes where the epilogue would otherwise wait for the next
                    // layer-1 part: after part 0 (kP1A = 0), after part 1 (1), or after the whole of epilogue 1 (2).
                    layer1(c1++, kP1A, c2);
                    if (next + G < ntiles) { stage(); load_x(next + 2 * G); }   // X of the tile after: ready long before it is needed
                    if (kP1A == 2) pass1a(c2);
                } else pass1a(c2);
                layer2(prev, c2++);
                if (!has_next) break;
                prev = next;
            }
        }
        if (prof && blockIdx.x == 0 && threadIdx.x == 0) {
            dbg[4] = e_wa; dbg[5] = e_pa; dbg[6] = e_1; dbg[7] = e_wb; dbg[8] = e_2; dbg[9] = clock64() - t_begin;
            dbg[10] = e_st; dbg[11] = e_2a; dbg[12] = e_2b; dbg[13] = e_2c; dbg[14] = e_2d;
            dbg[15] = e_wf; dbg[16] = e_af; dbg[17] = e_b1; dbg[18] = e_b2;
        }
    }
    // ---------------- teardown ----------------
    __syncwarp();
    tc_fence_before();
    __syncthreads();
    if (kCluster > 1) cluster_sync_all();            // nobody leaves while the partner may still multicast into / arrive on this CTA
    if (warp == kM2Warp) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(kTmemCols) : "memory");
    }
}

template <typename OpT, bool kSplit>
int launch_tc4(const char *w1img, const char *w2img, const tt_actor_dev &A, const float *d_obs, int64_t ld, int64_t n, float *d_mu,
               const TTRingS *ring, const TTActorTail *tail, unsigned long long *dbg, cudaStream_t st) {
    TTRingS rs;
    if (ring) rs = *ring; else { rs.S = nullptr; rs.m = tt_make_ring_map(1, 0, 0); }
    const TTActorTail tl = tail ? *tail : tt_no_tail();
    using P = Plan4<kSplit>;
    static_assert(P::total <= 232448u, "shared-memory plan exceeds 227 KB");
    // per device: 2 = CTA pairs with multicast W2 loads (default when 74 pairs are co-resident), 1 = single CTAs, 0 = not probed yet
    static int cluster_of[tt::kMaxDevices] = {};
    int &cluster = cluster_of[tt::device_index()];
    if (cluster == 0) {
#if defined(TT_DEV_VARIANTS)
        TT_CUDA(cudaFuncSetAttribute(actor_tc4_kernel<OpT, kSplit, true, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)P::total));
        TT_CUDA(cudaFuncSetAttribute(actor_tc4_kernel<OpT, kSplit, true, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)P::total));
#endif
        TT_CUDA(cudaFuncSetAttribute(actor_tc4_kernel<OpT, kSplit, false, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)P::total));
        TT_CUDA(cudaFuncSetAttribute(actor_tc4_kernel<OpT, kSplit, false, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)P::total));
        const char *e = getenv("TT_TC_CLUSTER");
        int want = e ? atoi(e) : 2;
        if (want == 2) {
            cudaLaunchConfig_t q{};
            q.gridDim = dim3((unsigned)(tt::sm_count() & ~1)); q.blockDim = dim3(640); q.dynamicSmemBytes = P::total;
            cudaLaunchAttribute at[1];
            at[0].id = cudaLaunchAttributeClusterDimension; at[0].val.clusterDim.x = 2; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
            q.attrs = at; q.numAttrs = 1;
            int nclusters = 0;
            if (cudaOccupancyMaxActiveClusters(&nclusters, actor_tc4_kernel<OpT, kSplit, false, 2>, &q) != cudaSuccess || 2 * nclusters < (tt::sm_count() & ~1)) {
                (void)cudaGetLastError();
                want = 1;
            }
        }
        cluster = want == 2 ? 2 : 1;
    }
    const int64_t ntiles = (n + kTileM - 1) / kTileM;
    const int sms = tt::grid_sms();
    int grid = (int)(ntiles < sms ? ntiles : sms);
    const bool chained = tt::chain_rollout(n);
#if !defined(TT_DEV_VARIANTS)
    dbg = nullptr;
#endif
    // CTA pairs (multicast W2 stream) pay off only when every SM runs many tiles: measured per rollout iteration, pairs / single
    // CTAs: 2^12 envs 20.0 / 18.0 us, 2^16 41.6 / 39.0, 2^18 98.5 / 96.6, 2^20 354 / 357, 2^22 - 1.7 % with pairs (the pair's set-up --
    // cluster barrier, remote mbarrier arrivals -- is pure latency for a CTA that sees one or a few tiles)
    if (cluster == 2 && ntiles > 2048) {
        grid = (grid + 1) & ~1;                                  // whole pairs; a CTA without tiles only serves the pair's W2 ring
        if (grid > (sms & ~1)) grid = sms & ~1;
        cudaLaunchConfig_t cfg{};
        cfg.gridDim = dim3((unsigned)grid); cfg.blockDim = dim3(640); cfg.dynamicSmemBytes = P::total; cfg.stream = st;
        cudaLaunchAttribute at[2];
        at[0].id = cudaLaunchAttributeClusterDimension; at[0].val.clusterDim.x = 2; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
        at[1].id = cudaLaunchAttributeProgrammaticStreamSerialization; at[1].val.programmaticStreamSerializationAllowed = 1;
        cfg.attrs = at; cfg.numAttrs = chained ? 2 : 1;
#if defined(TT_DEV_VARIANTS)
        if (dbg) TT_CUDA(cudaLaunchKernelEx(&cfg, actor_tc4_kernel<OpT, kSplit, true, 2>, w1img, w2img, A, d_obs, ld, n, d_mu, rs, tl, dbg));
        else
#endif
        TT_CUDA(cudaLaunchKernelEx(&cfg, actor_tc4_kernel<OpT, kSplit, false, 2>, w1img, w2img, A, d_obs, ld, n, d_mu, rs, tl, dbg));
    } else {
#if defined(TT_DEV_VARIANTS)
        if (dbg) actor_tc4_kernel<OpT, kSplit, true, 1><<<grid, 640, P::total, st>>>(w1img, w2img, A, d_obs, ld, n, d_mu, rs, tl, dbg);
        else
#endif
        TT_CUDA(tt::launch_chained(chained, actor_tc4_kernel<OpT, kSplit, false, 1>, dim3((unsigned)grid), dim3(640), P::total, st, w1img, w2img, A, d_obs, ld, n,
                                   d_mu, rs, tl, dbg));
    }
    TT_COUNT_LAUNCH(); TT_LAUNCH_CHECK();
    return TT_OK;
}

unsigned long long *g_tc_dbg = nullptr;

}  // namespace

namespace tt {

bool actor_tc_supported(const tt_actor_dev &A) { return A.in_dim == IN && A.h1 == H1 && A.h2 == H2; }

int actor_pack_all(tt_actor *a, const float *fc1_w, const float *fc1_b, const float *g1, const float *be1, const float *fc2_w, const float *fc2_b,
                   const float *g2, const float *be2, const float *mu_w, const float *mu_b, cudaStream_t s) {
    const tt_actor_dev &A = a->dev;
    const PackSrc w = {fc1_w, fc1_b, g1, be1, fc2_w, fc2_b, g2, be2, mu_w, mu_b};
    const int tc = actor_tc_supported(A) ? 1 : 0;       // the tensor-core path is specialised to 23-400-300; forward() refuses otherwise
    TT_CUDA(tt::launch_chained(true, pack_stage_a_kernel, dim3((tc ? kGramBlocks + KB2 : 0) + kPackFp32Blocks), dim3(1024), 0, s, A, w, tc));
    TT_COUNT_LAUNCH(); TT_LAUNCH_CHECK();
    if (!tc) return TT_OK;
#if defined(TT_LEARN_PROFILE)
    { const char *e_ = getenv("TT_PACK_STOP"); if (e_ && atoi(e_) == 1) return TT_OK; }
#endif
    TT_CUDA(tt::launch_chained(true, pack_stage_c_kernel, dim3(kImgBlocks + 2 * kW2sBlocks), dim3(256), 0, s, A, w));
    TT_COUNT_LAUNCH(); TT_LAUNCH_CHECK();
    return TT_OK;
}

int actor_forward_tc(const tt_actor *a, const float *d_obs, int64_t ld, int64_t n, float *d_mu, int precision, const TTRingS *ring,
                     const TTActorTail *tail, cudaStream_t st) {
    const tt_actor_dev &A = a->dev;
    if (!actor_tc_supported(A)) {
        set_error("tensor-core actor is specialised to layer sizes 23-400-300 (got %d-%d-%d); use TT_PREC_FP32", A.in_dim, A.h1, A.h2);
        return TT_ERR_INVALID;
    }
    if (precision == TT_PREC_BF16)
        return launch_tc4<__nv_bfloat16, false>(reinterpret_cast<const char *>(A.w1c_bf16), reinterpret_cast<const char *>(A.w2s_bf16), A, d_obs, ld, n, d_mu, ring, tail, g_tc_dbg, st);
    if (precision == TT_PREC_F16_PLAIN)          // plain fp16 first layer: only the hi block of the image is loaded
        return launch_tc4<__half, false>(reinterpret_cast<const char *>(A.w1c_f16), reinterpret_cast<const char *>(A.w2s_f16), A, d_obs, ld, n, d_mu, ring, tail, g_tc_dbg, st);
    return launch_tc4<__half, true>(reinterpret_cast<const char *>(A.w1c_f16), reinterpret_cast<const char *>(A.w2s_f16), A, d_obs, ld, n, d_mu, ring, tail, g_tc_dbg, st);
}

}  // namespace tt

#if defined(TT_DEV_VARIANTS)
// Developer build only (not part of the public header): per-phase cycle counters of block 0 of the tensor-core actor.
// d_buf: device buffer of >= 24 uint64, or NULL to switch profiling off.
extern "C" void tt_debug_set_tc_profile(unsigned long long *d_buf) { g_tc_dbg = d_buf; }
#endif
