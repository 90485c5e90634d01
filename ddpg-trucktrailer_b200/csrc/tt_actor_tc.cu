// tt_actor_tc.cu -- kernel (c), tensor-core path of the batched actor forward (ActorNetwork.forward,
// DDPG/networks.py:138-147): dispatch between the kernel generations and the PREVIOUS generation ("v3": pipelined, LayerNorm-1
// statistics from the Gram matrix of W1 on the CUDA cores, layer 2 as one N = 256 + 48 sweep).  The default kernel ("v4") lives
// in tt_actor_tc4.cu; v3 stays selectable with TT_TC_VARIANT=3 as a cross-check.
//
//   X[128 x 32] 16-bit (23 obs + a constant-1 column that carries the fc1 bias)       shared, SWIZZLE_64B
//   H1 = X . W1^T       tcgen05.mma kind::f16, N = 400 in two halves through a 208-column TMEM window, K = 32
//   A2 = relu(LN(H1))   tcgen05.ld -> registers -> 16-bit -> shared (13 k-blocks of 32; column 400 = 1 carries
//                       the fc2 bias), written directly in the UMMA K-major SWIZZLE_64B image
//   H2 = A2 . W2^T      N = 304 (300 padded) as 256 + 48, K = 416; W2 is streamed per tile from L2 in 13
//                       k-blocks of 19 KB by cp.async.bulk into a ring                   TMEM cols [0, 304)
//   mu = tanh(w3 . relu(LN(H2)) + b3)                                                 tcgen05.ld epilogue
//
// The weight images are pre-swizzled at tt_actor_load time so a k-block is one contiguous bulk copy.
#include <cuda_fp16.h>
#include <stdlib.h>
#include <type_traits>
#include "tt_actor.cuh"
#include "tt_common.cuh"

#ifndef TT_TC_VARIANT_DEFAULT
#define TT_TC_VARIANT_DEFAULT 4   // 4 = v4 (tt_actor_tc4.cu, default), 3 = v3 (this file; TT_TC_VARIANT overrides)
#endif

#include "tt_tc_ptx.cuh"

namespace {

// ---- weight images (device global), built at load time ----
//   w1 image: [hi block | lo block], each N1 rows x 64 B (k < IN: fc1.weight, k == IN: fc1.bias, else 0);
//             lo = the 16-bit rounding residual of hi (used by the split layer-1 product)
//   w2 image: KB2 blocks x (N2 rows x 64 B)   (k < H1: fc2.weight, k == H1: fc2.bias, else 0)
template <typename OpT>
__global__ void pack_tc_kernel(char *__restrict__ w1, char *__restrict__ w2, const float *__restrict__ fc1_w,
                               const float *__restrict__ fc1_b, const float *__restrict__ fc2_w, const float *__restrict__ fc2_b) {
    const int tid = blockIdx.x * blockDim.x + threadIdx.x, nth = gridDim.x * blockDim.x;
    for (int v = tid; v < N1 * 32; v += nth) {
        const int n = v / 32, k = v - n * 32;
        float x = 0.f;
        if (n < H1) x = k < IN ? fc1_w[n * IN + k] : (k == IN ? fc1_b[n] : 0.f);
        const OpT hi = to_op<OpT>(x);
        *reinterpret_cast<OpT *>(w1 + sw64_off(n, k)) = hi;
        *reinterpret_cast<OpT *>(w1 + N1 * kRowB + sw64_off(n, k)) = to_op<OpT>(x - op_to_float(hi));
    }
    for (int v = tid; v < KB2 * N2 * 32; v += nth) {
        const int kb = v / (N2 * 32), rem = v - kb * N2 * 32, n = rem / 32, kk = rem - n * 32, k = kb * 32 + kk;
        float x = 0.f;
        if (n < H2) x = k < H1 ? fc2_w[n * H1 + k] : (k == H1 ? fc2_b[n] : 0.f);
        *reinterpret_cast<OpT *>(w2 + (size_t)kb * N2 * kRowB + sw64_off(n, kk)) = to_op<OpT>(x);
    }
}

// Gram matrix of the (operand-rounded) first layer incl. its bias column: G[i][j] = sum_c W[c][i] W[c][j], wbar[i] =
// sum_c W[c][i] (i, j < 24).  With them the LayerNorm statistics of H1 = W x follow from x alone:
// sum_c h_c = wbar . x,  sum_c h_c^2 = x^T G x  -- so epilogue 1 needs ONE pass over TMEM instead of two.
// Stored as the upper-triangular form U (U_ii = G_ii, U_ij = 2 G_ij for j > i, 0 below) so that x^T G x = sum_i x_i
// sum_{j>=i} U_ij x_j needs half the multiply-adds.  layout: U row-major [24][24], then wbar[24].
// kRound: round W to the operand type first (plain bf16 mode).
// (Keeping these per-column parameters in __constant__ memory instead of shared memory was tried and measured 30 %
// SLOWER: the 12 KB working set thrashes the small constant cache.)
template <typename OpT, bool kRound>
__global__ void pack_gram_kernel(float *__restrict__ gram, const float *__restrict__ fc1_w, const float *__restrict__ fc1_b) {
    const int t = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;     // one warp per output element
    if (t >= 25 * 24) return;
    const int i = t / 24, j = t - i * 24;
    auto w = [&](int c, int k) {
        const float x = k < IN ? fc1_w[c * IN + k] : fc1_b[c];
        return (double)(kRound ? op_to_float(to_op<OpT>(x)) : x);
    };
    double acc = 0.0;
    if (i < 24) { if (j >= i) for (int c = lane; c < H1; c += 32) acc += w(c, i) * w(c, j); }
    else for (int c = lane; c < H1; c += 32) acc += w(c, j);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if (lane == 0) gram[t] = (float)((i < 24 && j > i) ? 2.0 * acc : acc);
}

unsigned long long *g_tc_dbg = nullptr;      // optional device buffer for the per-phase cycle counters (tt_debug_set_tc_profile)


// =====================================================================================================================
// v3: pipelined variant.  TMEM columns [0,304) hold H2 of tile t while a 208-column window [304,512) carries H1 of
// tile t+1 in two halves (192 + 208 columns); the LayerNorm statistics of H1 come from the Gram matrix of W1 (one TMEM
// pass), and epilogue 1 of tile t+1 writes A2 block kb as soon as the layer-2 MMAs of tile t have consumed it.  So
// epilogue 1 (t+1) and both layer-1 MMAs run in the shadow of the layer-2 MMAs of tile t; two MMA-issuer threads
// (layer 1 / layer 2) feed the one in-order tensor pipe.
// =====================================================================================================================
template <bool kSplit>
struct Plan3 {
    static constexpr int kXBlocks = kSplit ? 2 : 1;
    static constexpr int kSlots = kSplit ? 2 : 3;
    static constexpr uint32_t kW2Slot = ((uint32_t)N2 * kRowB + 1023u) / 1024u * 1024u;
    static constexpr uint32_t x = 0;
    static constexpr uint32_t w1 = x + kXBlocks * kTileM * kRowB;
    static constexpr uint32_t a2 = w1 + kXBlocks * N1 * kRowB;
    static constexpr uint32_t w2 = a2 + KB2 * kTileM * kRowB;
    static constexpr uint32_t par = w2 + kSlots * kW2Slot;
    static constexpr uint32_t npar = 2 * K2P + 3 * H2P + 25 * 24;                // + Gram matrix and wbar
    static constexpr uint32_t red = par + npar * 4;
    static constexpr uint32_t bars = red + 4 * kTileM * 8 + 4 * kTileM * 4;
    static constexpr uint32_t nbars = 32;
    static constexpr uint32_t tmem_slot = bars + nbars * 8;
    static constexpr uint32_t total = tmem_slot + 16 + 1024;
};

enum { C_W1 = 0, C_XFULL, C_H1AFULL, C_H1BFULL, C_WINFREE_A, C_WINFREE_B, C_A2FULL, C_H2FULL, C_H2FREE, C_W2FULL,
       C_W2EMPTY = C_W2FULL + 3, C_A2FREE = C_W2EMPTY + 3, C_COUNT = C_A2FREE + KB2 };
static_assert(C_COUNT <= 32, "barrier table");

constexpr int kWinCol = 304;       // first TMEM column of the H1 window
constexpr int kN1A = 192, kN1B = N1 - kN1A;                    // layer-1 column halves (6 and 6.5 k-blocks of A2)
constexpr int kChA = kN1A / 32;                                // 6 chunks in half a; half b = chunks 6..12

template <typename OpT, bool kSplit>
__global__ void __launch_bounds__(640, 1) actor_tc3_kernel(const char *__restrict__ w1img, const char *__restrict__ w2img,
                                                           const float *__restrict__ gram, tt_actor_dev A,
                                                           const float *__restrict__ obs, int64_t ld, int64_t n,
                                                           float *__restrict__ out, TTRingS ring,
                                                           unsigned long long *__restrict__ dbg) {
    using P = Plan3<kSplit>;
    constexpr int kGroups = 4, kEpiThreads = 512, kThreads = 640;
    constexpr int kM2Warp = 16, kProdWarp = 17, kM1Warp = 18, kCopyWarp = 19;
    constexpr uint32_t kFmt = std::is_same<OpT, __nv_bfloat16>::value ? 1u : 0u;
    extern __shared__ uint8_t smem_raw[];
    const uint32_t raw = smem_u32(smem_raw);
    const uint32_t base = (raw + 1023u) & ~1023u;
    uint8_t *sm = smem_raw + (base - raw);
    const uint32_t sX = base + P::x, sW1 = base + P::w1, sA2 = base + P::a2, sW2 = base + P::w2, sBar = base + P::bars;
    float *par = reinterpret_cast<float *>(sm + P::par);
    float *pg1 = par, *pbe1 = par + K2P, *pg2 = par + 2 * K2P, *pbe2 = pg2 + H2P, *pw3 = pbe2 + H2P, *pgram = pw3 + H2P;
    float2 *red1 = reinterpret_cast<float2 *>(sm + P::red);
    float *red3 = reinterpret_cast<float *>(red1 + 4 * kTileM);
    volatile uint32_t *tmem_slot = reinterpret_cast<volatile uint32_t *>(sm + P::tmem_slot);
    auto bar = [&](int i) { return sBar + 8u * (uint32_t)i; };

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int64_t ntiles = (n + kTileM - 1) / kTileM;

    // ---------------- one-time setup ----------------
    if (threadIdx.x == 0) {
        mbar_init(bar(C_W1), 1);
        mbar_init(bar(C_XFULL), kEpiThreads); mbar_init(bar(C_H1AFULL), 1); mbar_init(bar(C_H1BFULL), 1);
        mbar_init(bar(C_WINFREE_A), kEpiThreads); mbar_init(bar(C_WINFREE_B), kEpiThreads); mbar_init(bar(C_A2FULL), kEpiThreads);
        mbar_init(bar(C_H2FULL), 1); mbar_init(bar(C_H2FREE), kEpiThreads);
        for (int i = 0; i < P::kSlots; i++) { mbar_init(bar(C_W2FULL + i), 1); mbar_init(bar(C_W2EMPTY + i), 1); }
        for (int i = 0; i < KB2; i++) mbar_init(bar(C_A2FREE + i), 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == kM2Warp) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(base + P::tmem_slot), "r"(kTmemCols) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    for (int c = threadIdx.x; c < K2P; c += kThreads) {
        pg1[c] = c < H1 ? A.g1[c] : 0.f;
        pbe1[c] = c < H1 ? A.be1[c] : (c == H1 ? 1.f : 0.f);
    }
    for (int c = threadIdx.x; c < H2P; c += kThreads) {
        const bool in = c < H2;
        pg2[c] = in ? A.g2[c] : 0.f; pbe2[c] = in ? A.be2[c] : 0.f; pw3[c] = in ? A.w3[c] : 0.f;
    }
    for (int c = threadIdx.x; c < 25 * 24; c += kThreads) pgram[c] = gram[c];
    for (int v = threadIdx.x; v < P::kXBlocks * kTileM * kRowB / 4; v += kThreads) reinterpret_cast<uint32_t *>(sm + P::x)[v] = 0u;
    __syncthreads();
    if (threadIdx.x < kTileM) *reinterpret_cast<OpT *>(sm + P::x + sw64_off(threadIdx.x, IN)) = to_op<OpT>(1.0f);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *tmem_slot;
    const float b3 = A.b3[0];

    if (warp == kProdWarp) {
        // ================= bulk-copy producer =================
        if (lane == 0) {
            constexpr uint32_t w1bytes = (uint32_t)P::kXBlocks * N1 * kRowB;
            mbar_expect_tx(bar(C_W1), w1bytes);
            bulk_g2s(sW1, w1img, w1bytes, bar(C_W1));
            uint32_t it = 0;
            for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
                for (int kb = 0; kb < KB2; kb++, it++) {
                    const uint32_t slot = it % P::kSlots, ph = (it / P::kSlots) & 1u;
                    mbar_wait(bar(C_W2EMPTY + slot), ph ^ 1u);
                    mbar_expect_tx(bar(C_W2FULL + slot), (uint32_t)N2 * kRowB);
                    bulk_g2s(sW2 + slot * P::kW2Slot, w2img + (size_t)kb * N2 * kRowB, (uint32_t)N2 * kRowB, bar(C_W2FULL + slot));
                }
            }
        }
    } else if (warp == kCopyWarp) {
        // ================= fused replay store of s: observation rows -> ring `state` rows =================
        // A dedicated warp, so the copy stays off the epilogue's critical path; DRAM is idle in this kernel (2.5 %).
        if (ring.S) {
            for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
                const int64_t row0 = tile * kTileM;
                const int rows = (int)((n - row0) < kTileM ? (n - row0) : kTileM);
                const int64_t rrow0 = ring.m.row(row0);
                const float *src = obs + row0 * ld;
                float *dst = ring.S + rrow0 * IN;
                const bool flat = ld == IN && rows == kTileM && !ring.m.many && row0 >= ring.m.first && rrow0 + kTileM <= ring.m.cap;
                if (flat && (((uintptr_t)src | (uintptr_t)dst) & 15) == 0) {
                    const float4 *s4 = reinterpret_cast<const float4 *>(src);
                    float4 *d4 = reinterpret_cast<float4 *>(dst);
#pragma unroll 4
                    for (int v = lane; v < kTileM * IN / 4; v += 32) __stcs(&d4[v], __ldcs(&s4[v]));
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
        // ================= layer-1 MMA issuer =================
        if (lane == 0) {
            const uint32_t ida = make_idesc(kN1A, kFmt), idb = make_idesc(kN1B, kFmt);
            constexpr int npairs = kSplit ? 3 : 1;
            uint32_t c1 = 0;
            mbar_wait(bar(C_W1), 0);
            for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x, c1++) {
                const uint32_t ph = c1 & 1u;
                mbar_wait(bar(C_XFULL), ph);
                mbar_wait(bar(C_WINFREE_B), ph ^ 1u);                     // epilogue 1b of the previous tile has drained the window
                tc_fence_after();
#pragma unroll
                for (int half = 0; half < 2; half++) {
                    if (half == 1) { mbar_wait(bar(C_WINFREE_A), ph); tc_fence_after(); }   // epilogue 1a has drained the window
#pragma unroll
                    for (int pr = 0; pr < npairs; pr++) {
                        const uint32_t xb = sX + (pr == 1 ? kTileM * kRowB : 0);
                        const uint32_t wb = sW1 + (pr == 2 ? N1 * kRowB : 0) + (half ? kN1A * kRowB : 0);
#pragma unroll
                        for (int ks = 0; ks < 2; ks++)
                            umma(tmem + kWinCol, make_desc(xb + ks * 32), make_desc(wb + ks * 32), half ? idb : ida, (pr | ks) ? 1u : 0u);
                    }
                    umma_commit(bar(half ? C_H1BFULL : C_H1AFULL));
                }
            }
        }
    } else if (warp == kM2Warp) {
        // ================= layer-2 MMA issuer =================
        if (lane == 0) {
            constexpr int mA = 256, mB = N2 - 256;
            const uint32_t id2a = make_idesc(mA, kFmt), id2b = make_idesc(mB, kFmt);
            uint32_t it = 0, c2 = 0;
            long long t_a2 = 0, t_w2 = 0, t_h2 = 0, t0;
            for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x, c2++) {
                const uint32_t ph = c2 & 1u;
                t0 = clock64();
                mbar_wait(bar(C_A2FULL), ph);
                t_a2 += clock64() - t0; t0 = clock64();
                mbar_wait(bar(C_H2FREE), ph ^ 1u);                        // epilogue 2 of the previous tile has drained H2
                t_h2 += clock64() - t0;
                tc_fence_after();
                for (int kb = 0; kb < KB2; kb++, it++) {
                    const uint32_t slot = it % P::kSlots, wph = (it / P::kSlots) & 1u;
                    t0 = clock64();
                    mbar_wait(bar(C_W2FULL + slot), wph);
                    t_w2 += clock64() - t0;
                    tc_fence_after();
                    const uint32_t wb = sW2 + slot * P::kW2Slot, ab = sA2 + (uint32_t)kb * kTileM * kRowB;
#pragma unroll
                    for (int ks = 0; ks < 2; ks++) {
                        const uint64_t a = make_desc(ab + ks * 32);
                        const uint32_t acc = (kb | ks) ? 1u : 0u;
                        umma(tmem, a, make_desc(wb + ks * 32), id2a, acc);
                        umma(tmem + 256, a, make_desc(wb + 256 * kRowB + ks * 32), id2b, acc);
                    }
                    umma_commit(bar(C_W2EMPTY + slot));
                    umma_commit(bar(C_A2FREE + kb));                      // A2 block kb may be overwritten for the next tile
                }
                umma_commit(bar(C_H2FULL));
            }
            if (dbg && blockIdx.x == 0) { dbg[0] = t_h2; dbg[1] = t_a2; dbg[2] = t_w2; dbg[3] = c2; }
        }
    } else {
        // ================= epilogue warps: thread = (row, column group) =================
        const int grp = warp >> 2;
        const int r = (warp & 3) * 32 + lane;
        const int et = threadIdx.x;
        const uint32_t trow = tmem + ((uint32_t)((warp & 3) * 32) << 16);
        const uint32_t xsw = (((uint32_t)r >> 1) & 3u);
        constexpr int kXPer = (kTileM * IN + kEpiThreads - 1) / kEpiThreads;
        float xreg[kXPer];
        auto load_x = [&](int64_t t) {
            const int64_t r0 = t * kTileM;
#pragma unroll
            for (int i = 0; i < kXPer; i++) {
                const int v = et + i * kEpiThreads;
                const int rr = v / IN, k = v - rr * IN;
                xreg[i] = (v < kTileM * IN && r0 + rr < n) ? __ldg(obs + (r0 + rr) * ld + k) : 0.f;
            }
        };
        long long e_1 = 0, e_w2 = 0, e_2 = 0, e_wa = 0, e_s0 = 0, e_s1 = 0, e_s2 = 0, t0, t1;
        const long long t_begin = clock64();
        uint32_t v[32];

        // layer-1 side of one tile: stage X, Gram statistics, epilogue 1a / 1b -> A2.  c1 = layer-1 tile counter
        auto layer1 = [&](int64_t tile, uint32_t c1) {
            const uint32_t ph = c1 & 1u;
            const int64_t row0 = tile * kTileM;
            t0 = clock64();
            // X may be overwritten: both layer-1 MMAs of the previous tile completed before its H1BFULL (waited on below)
#pragma unroll
            for (int i = 0; i < kXPer; i++) {
                const int vv = et + i * kEpiThreads;
                if (vv < kTileM * IN) {
                    const int rr = vv / IN, k = vv - rr * IN;
                    const float x = xreg[i];
                    const OpT hi = to_op<OpT>(x);
                    *reinterpret_cast<OpT *>(sm + P::x + sw64_off(rr, k)) = hi;

                    if (kSplit) *reinterpret_cast<OpT *>(sm + P::x + kTileM * kRowB + sw64_off(rr, k)) = to_op<OpT>(x - op_to_float(hi));
                }
            }
            fence_proxy_async();
            mbar_arrive(bar(C_XFULL));
            load_x(tile + gridDim.x);
            // LayerNorm statistics of H1 from the Gram matrix.  x_r is read back from the staged operand tile (hi [+ lo]):
            // exactly the values the tensor core multiplies.  This thread covers rows i = grp, grp+4, ... of G.
            t1 = clock64(); e_s0 += t1 - t0; t0 = t1;
            named_bar_sync(1, kEpiThreads);                               // the whole X tile is staged
            t1 = clock64(); e_s1 += t1 - t0; t0 = t1;
            float xr[24];
            {
                const uint8_t *xrow = sm + P::x + (size_t)r * kRowB;
#pragma unroll
                for (int q = 0; q < 3; q++) {                              // 3 x 16 B = 24 operands of row r
                    const uint4 hq = *reinterpret_cast<const uint4 *>(xrow + (((uint32_t)q ^ xsw) << 4));
                    const OpT *hp = reinterpret_cast<const OpT *>(&hq);
#pragma unroll
                    for (int j = 0; j < 8; j++) xr[q * 8 + j] = op_to_float(hp[j]);
                    if (kSplit) {
                        const uint4 lq = *reinterpret_cast<const uint4 *>(xrow + kTileM * kRowB + (((uint32_t)q ^ xsw) << 4));
                        const OpT *lp = reinterpret_cast<const OpT *>(&lq);
#pragma unroll
                        for (int j = 0; j < 8; j++) xr[q * 8 + j] += op_to_float(lp[j]);
                    }
                }
            }
            float ms = 0.f, qs = 0.f;
#pragma unroll
            for (int ii = 0; ii < 6; ii++) {
                const int i = grp + 4 * ii;                               // rows grp, grp+4, ...: entries j >= 4*ii suffice (U is upper)
                float2 in2 = make_float2(0.f, 0.f);
#pragma unroll
                for (int c = ii; c < 6; c++) {
                    const float4 g = *reinterpret_cast<const float4 *>(pgram + i * 24 + 4 * c);
                    in2 = __ffma2_rn(make_float2(g.x, g.y), make_float2(xr[4 * c], xr[4 * c + 1]), in2);
                    in2 = __ffma2_rn(make_float2(g.z, g.w), make_float2(xr[4 * c + 2], xr[4 * c + 3]), in2);
                }
                // x_i for this thread's rows of U: grp is warp-uniform, so this is a uniform 4-way select
                const float xi = grp == 0 ? xr[4 * ii] : grp == 1 ? xr[4 * ii + 1] : grp == 2 ? xr[4 * ii + 2] : xr[4 * ii + 3];
                qs = fmaf(xi, in2.x + in2.y, qs);
                ms = fmaf(pgram[24 * 24 + i], xi, ms);
            }
            red1[grp * kTileM + r] = make_float2(ms, qs);
            t1 = clock64(); e_s2 += t1 - t0; t0 = t1;
            named_bar_sync(1, kEpiThreads);
            float sum = 0.f, sq = 0.f;
#pragma unroll
            for (int g = 0; g < kGroups; g++) { const float2 t = red1[g * kTileM + r]; sum += t.x; sq += t.y; }
            const float mean = sum * (1.0f / H1);
            const float rstd = rsqrtf(fmaxf(sq * (1.0f / H1) - mean * mean, 0.f) + 1e-5f);
            const float nmr = -mean * rstd;
            const float2 rstd2 = make_float2(rstd, rstd), nmr2 = make_float2(nmr, nmr);
            t1 = clock64(); e_wa += t1 - t0; t0 = t1;
#pragma unroll
            for (int half = 0; half < 2; half++) {
                mbar_wait(bar(half ? C_H1BFULL : C_H1AFULL), ph);
                tc_fence_after();
#pragma unroll
                for (int ch = 0; ch < NCH1; ch++) {
                    if ((ch < kChA) != (half == 0) || ch % kGroups != grp) continue;
                    mbar_wait(bar(C_A2FREE + ch), ph ^ 1u);               // layer-2 MMAs of the previous tile have read this block
                    const int lc = half ? ch - kChA : ch;                 // chunk index inside the window
                    // last chunk: 16 accumulator columns (window columns 192..207) + the constant columns (read as 0)
                    if (ch == NCH1 - 1) tmem_ld16_async(trow + (uint32_t)(kWinCol + lc * 32), v);
                    else tmem_ld32_async(trow + (uint32_t)(kWinCol + lc * 32), v);
                    tmem_wait();
                    uint8_t *blk = sm + P::a2 + (size_t)ch * kTileM * kRowB + (size_t)r * kRowB;
#pragma unroll
                    for (int q = 0; q < 4; q++) {
                        const float4 g0 = *reinterpret_cast<const float4 *>(pg1 + ch * 32 + q * 8), g1 = *reinterpret_cast<const float4 *>(pg1 + ch * 32 + q * 8 + 4);
                        const float4 e0 = *reinterpret_cast<const float4 *>(pbe1 + ch * 32 + q * 8), e1 = *reinterpret_cast<const float4 *>(pbe1 + ch * 32 + q * 8 + 4);
                        const float2 gg[4] = {make_float2(g0.x, g0.y), make_float2(g0.z, g0.w), make_float2(g1.x, g1.y), make_float2(g1.z, g1.w)};
                        const float2 ee[4] = {make_float2(e0.x, e0.y), make_float2(e0.z, e0.w), make_float2(e1.x, e1.y), make_float2(e1.z, e1.w)};
                        uint32_t pk4[4];
#pragma unroll
                        for (int j = 0; j < 4; j++) {
                            const float2 x = make_float2(__uint_as_float(v[q * 8 + 2 * j]), __uint_as_float(v[q * 8 + 2 * j + 1]));
                            const float2 y = __ffma2_rn(__ffma2_rn(x, rstd2, nmr2), gg[j], ee[j]);
                            pk4[j] = pack2_relu<OpT>(y.x, y.y);
                        }
                        uint4 pk;
                        pk.x = pk4[0]; pk.y = pk4[1]; pk.z = pk4[2]; pk.w = pk4[3];
                        *reinterpret_cast<uint4 *>(blk + (((uint32_t)q ^ xsw) << 4)) = pk;
                    }
                }
                tc_fence_before();
                if (half == 0) mbar_arrive(bar(C_WINFREE_A));
                else { fence_proxy_async(); mbar_arrive(bar(C_WINFREE_B)); mbar_arrive(bar(C_A2FULL)); }
            }
            t1 = clock64(); e_1 += t1 - t0; t0 = t1;
        };

        // layer-2 side of one tile: LayerNorm + ReLU over H2, dot with mu.weight, tanh.  c2 = layer-2 tile counter
        auto layer2 = [&](int64_t tile, uint32_t c2) {
            const uint32_t ph = c2 & 1u;
            const int64_t row0 = tile * kTileM;
            const int rows = (int)((n - row0) < kTileM ? (n - row0) : kTileM);
            t0 = clock64();
            mbar_wait(bar(C_H2FULL), ph);
            t1 = clock64(); e_w2 += t1 - t0; t0 = t1;
            tc_fence_after();
            float2 s2 = make_float2(0.f, 0.f), q2 = make_float2(0.f, 0.f);
#pragma unroll
            for (int ch = 0; ch < NCH2; ch++) {
                if (ch % kGroups != grp) continue;
                tmem_ld_chunk<N2>(trow, ch, v);
                tmem_wait();
#pragma unroll
                for (int j = 0; j < 16; j++) {
                    const float2 x = make_float2(__uint_as_float(v[2 * j]), __uint_as_float(v[2 * j + 1]));
                    s2 = __fadd2_rn(s2, x); q2 = __ffma2_rn(x, x, q2);
                }
            }
            red1[grp * kTileM + r] = make_float2(s2.x + s2.y, q2.x + q2.y);
            named_bar_sync(1, kEpiThreads);
            float sum = 0.f, sq = 0.f;
#pragma unroll
            for (int g = 0; g < kGroups; g++) { const float2 t = red1[g * kTileM + r]; sum += t.x; sq += t.y; }
            const float mean = sum * (1.0f / H2);
            const float rstd = rsqrtf(fmaxf(sq * (1.0f / H2) - mean * mean, 0.f) + 1e-5f);
            const float nmr = -mean * rstd;
            const float2 rstd2 = make_float2(rstd, rstd), nmr2 = make_float2(nmr, nmr);
            float2 dot2 = make_float2(0.f, 0.f);
#pragma unroll
            for (int ch = 0; ch < NCH2; ch++) {
                if (ch % kGroups != grp) continue;
                tmem_ld_chunk<N2>(trow, ch, v);
                tmem_wait();
#pragma unroll
                for (int q = 0; q < 8; q++) {
                    const float4 g0 = *reinterpret_cast<const float4 *>(pg2 + ch * 32 + q * 4), e0 = *reinterpret_cast<const float4 *>(pbe2 + ch * 32 + q * 4),
                                 w0 = *reinterpret_cast<const float4 *>(pw3 + ch * 32 + q * 4);
                    const float2 xa = make_float2(__uint_as_float(v[q * 4 + 0]), __uint_as_float(v[q * 4 + 1]));
                    const float2 xb = make_float2(__uint_as_float(v[q * 4 + 2]), __uint_as_float(v[q * 4 + 3]));
                    float2 ya = __ffma2_rn(__ffma2_rn(xa, rstd2, nmr2), make_float2(g0.x, g0.y), make_float2(e0.x, e0.y));
                    float2 yb = __ffma2_rn(__ffma2_rn(xb, rstd2, nmr2), make_float2(g0.z, g0.w), make_float2(e0.z, e0.w));
                    ya.x = fmaxf(ya.x, 0.f); ya.y = fmaxf(ya.y, 0.f); yb.x = fmaxf(yb.x, 0.f); yb.y = fmaxf(yb.y, 0.f);
                    dot2 = __ffma2_rn(ya, make_float2(w0.x, w0.y), dot2);
                    dot2 = __ffma2_rn(yb, make_float2(w0.z, w0.w), dot2);
                }
            }
            tc_fence_before();
            mbar_arrive(bar(C_H2FREE));
            red3[grp * kTileM + r] = dot2.x + dot2.y;
            named_bar_sync(1, kEpiThreads);
            if (grp == 0 && r < rows) {
                float d = b3;
#pragma unroll
                for (int g = 0; g < kGroups; g++) d += red3[g * kTileM + r];
                out[row0 + r] = tanhf(d);
            }
            t1 = clock64(); e_2 += t1 - t0; t0 = t1;
        };

        load_x(blockIdx.x);
        uint32_t c1 = 0, c2 = 0;
        int64_t prev = blockIdx.x;
        if (prev < ntiles) {
            layer1(prev, c1++);
            for (;;) {
                const int64_t next = prev + gridDim.x;
                const bool has_next = next < ntiles;
                if (has_next) layer1(next, c1++);        // overlaps the layer-2 MMAs of `prev`
                layer2(prev, c2++);
                if (!has_next) break;
                prev = next;
            }
        }
        if (dbg && blockIdx.x == 0 && threadIdx.x == 0) {
            dbg[4] = e_wa; dbg[5] = 0; dbg[6] = e_1; dbg[7] = e_w2; dbg[8] = e_2; dbg[9] = clock64() - t_begin;
            dbg[10] = e_s0; dbg[11] = e_s1; dbg[12] = e_s2;
        }
    }
    // ---------------- teardown ----------------
    __syncwarp();
    tc_fence_before();
    __syncthreads();
    if (warp == kM2Warp) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(kTmemCols) : "memory");
    }
}

template <typename OpT, bool kSplit>
int launch_tc3(const char *w1img, const char *w2img, const float *gram, const tt_actor_dev &A, const float *d_obs, int64_t ld, int64_t n,
               float *d_mu, const TTRingS *ring, cudaStream_t st) {
    TTRingS rs;
    if (ring) rs = *ring; else { rs.S = nullptr; rs.m = tt_make_ring_map(1, 0, 0); }
    using P = Plan3<kSplit>;
    static_assert(P::total <= 232448u, "shared-memory plan exceeds 227 KB");
    auto kern = actor_tc3_kernel<OpT, kSplit>;
    static bool attr_set = false;
    if (!attr_set) {
        TT_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)P::total));
        attr_set = true;
    }
    const int64_t ntiles = (n + kTileM - 1) / kTileM;
    const int grid = (int)(ntiles < tt::sm_count() ? ntiles : tt::sm_count());
    kern<<<grid, 640, P::total, st>>>(w1img, w2img, gram, A, d_obs, ld, n, d_mu, rs, g_tc_dbg);
    TT_COUNT_LAUNCH(); TT_LAUNCH_CHECK();
    return TT_OK;
}

}  // namespace

namespace tt {

void set_tc_profile_buffer(unsigned long long *d) { g_tc_dbg = d; }

bool actor_tc_supported(const tt_actor_dev &A) { return A.in_dim == IN && A.h1 == H1 && A.h2 == H2; }

static int tc_variant();

int actor_pack_tc_full(tt_actor *a, const float *fc1_w, const float *fc1_b, const float *fc2_w, const float *fc2_b, cudaStream_t s) {
    const tt_actor_dev &A = a->dev;
    if (!actor_tc_supported(A)) return TT_OK;       // tensor-core path is specialised to 23-400-300; forward() will refuse
    // only the images of the kernel variant in use (the policy is re-packed after every learner step)
    if (tc_variant() >= 4) return actor_pack_tc4(a, fc1_w, fc1_b, A.g1, fc2_w, fc2_b, s);
    pack_tc_kernel<__half><<<128, 256, 0, s>>>(reinterpret_cast<char *>(A.w1_f16), reinterpret_cast<char *>(A.w2_f16), fc1_w, fc1_b, fc2_w, fc2_b);
    TT_COUNT_LAUNCH(); TT_LAUNCH_CHECK();
    pack_tc_kernel<__nv_bfloat16><<<128, 256, 0, s>>>(reinterpret_cast<char *>(A.w1_bf16), reinterpret_cast<char *>(A.w2_bf16), fc1_w, fc1_b, fc2_w, fc2_b);
    TT_COUNT_LAUNCH(); TT_LAUNCH_CHECK();
    pack_gram_kernel<__half, false><<<75, 256, 0, s>>>(A.gram_f16, fc1_w, fc1_b);
    TT_COUNT_LAUNCH(); TT_LAUNCH_CHECK();
    pack_gram_kernel<__nv_bfloat16, true><<<75, 256, 0, s>>>(A.gram_bf16, fc1_w, fc1_b);
    TT_COUNT_LAUNCH(); TT_LAUNCH_CHECK();
    return TT_OK;
}

static int tc_variant() {
    static const int variant = [] { const char *e = getenv("TT_TC_VARIANT"); return e ? atoi(e) : TT_TC_VARIANT_DEFAULT; }();
    return variant;
}
bool actor_tc_fuses_ring() { return tc_variant() >= 3; }

int actor_forward_tc(const tt_actor *a, const float *d_obs, int64_t ld, int64_t n, float *d_mu, int precision, const TTRingS *ring,
                     cudaStream_t st) {
    const tt_actor_dev &A = a->dev;
    if (!actor_tc_supported(A)) {
        set_error("tensor-core actor is specialised to layer sizes 23-400-300 (got %d-%d-%d); use TT_PREC_FP32", A.in_dim, A.h1, A.h2);
        return TT_ERR_INVALID;
    }
    const int variant = tc_variant();
    if (variant < 4 && precision == TT_PREC_F16_PLAIN) precision = TT_PREC_F16;      // the previous kernel has no plain-fp16 mode
    if (variant >= 4) return actor_forward_tc4(a, d_obs, ld, n, d_mu, precision, ring, g_tc_dbg, st);
    if (variant == 3) {
        if (precision == TT_PREC_BF16)
            return launch_tc3<__nv_bfloat16, false>(reinterpret_cast<const char *>(A.w1_bf16), reinterpret_cast<const char *>(A.w2_bf16), A.gram_bf16, A, d_obs, ld, n, d_mu, ring, st);
        return launch_tc3<__half, true>(reinterpret_cast<const char *>(A.w1_f16), reinterpret_cast<const char *>(A.w2_f16), A.gram_f16, A, d_obs, ld, n, d_mu, ring, st);
    }
    set_error("unknown tensor-core kernel variant %d (TT_TC_VARIANT: 4 = default, 3 = previous pipelined kernel)", variant);
    return TT_ERR_INVALID;
}

}  // namespace tt

// Debug hook (not part of the public header): per-phase cycle counters of block 0 of the tensor-core actor.
// d_buf: device buffer of >= 16 uint64, or NULL to switch profiling off.
extern "C" void tt_debug_set_tc_profile(unsigned long long *d_buf) { tt::set_tc_profile_buffer(d_buf); }
