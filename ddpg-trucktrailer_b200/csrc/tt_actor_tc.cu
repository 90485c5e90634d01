// tt_actor_tc.cu -- kernel (c), tensor-core path: the batched actor forward (ActorNetwork.forward,
// DDPG/networks.py:138-147) on the 5th-generation tensor cores of sm_100a: tcgen05.mma with bf16 operands
// staged in shared memory, fp32 accumulators in TMEM, tcgen05.ld for the LayerNorm/ReLU/tanh epilogues.
//
// One persistent CTA per SM; one tile = 128 observation rows (TMEM lane = row).
//
//   X[128 x 32] bf16  (23 obs + a constant-1 column that carries the fc1 bias)       shared, SWIZZLE_64B
//   H1 = X . W1^T       tcgen05.mma kind::f16, N = 400 as 256 + 144, K = 32           TMEM cols [0, 400)
//   A2 = relu(LN(H1))   tcgen05.ld -> registers -> bf16 -> shared (13 k-blocks of 32; column 400 = 1 carries
//                       the fc2 bias), written directly in the UMMA K-major SWIZZLE_64B image
//   H2 = A2 . W2^T      N = 304 (300 padded) as 256 + 48, K = 416; W2 is streamed per tile from L2 in 13
//                       k-blocks of 19 KB by cp.async.bulk into a 3-slot ring           TMEM cols [0, 304)
//   mu = tanh(w3 . relu(LN(H2)) + b3)                                                 tcgen05.ld epilogue
//
// Warp roles (192 threads): warps 0-3 epilogue (warp w owns TMEM lanes 32w..32w+31 = rows), warp 4 lane 0
// issues every tcgen05.mma / commit, warp 5 lane 0 is the bulk-copy producer.  All hand-offs are mbarriers.
// The weight images are pre-swizzled once at tt_actor_load time so a k-block is one contiguous bulk copy.
#include "tt_actor.cuh"
#include "tt_common.cuh"

namespace {

constexpr int kTileM = 128;
constexpr int kThreads = 192;
constexpr int kSlots = 3;          // W2 ring depth
constexpr int kRowB = 64;          // bytes per operand row in a k-block (32 bf16), SWIZZLE_64B
constexpr uint32_t kTmemCols = 512;

// ---- shapes derived from (in_dim, h1, h2) ----
struct TcShape {
    int in_dim, h1, h2;
    int n1;      // layer-1 MMA N  (h1 rounded up to 16)
    int n2;      // layer-2 MMA N  (h2 rounded up to 16)
    int kb2;     // layer-2 k-blocks of 32: ceil((h1 + 1) / 32)
    int k2p;     // kb2 * 32
    int h2p32;   // h2 rounded to 32 (parameter array stride)
};

__host__ __device__ inline TcShape make_shape(int in_dim, int h1, int h2) {
    TcShape s;
    s.in_dim = in_dim; s.h1 = h1; s.h2 = h2;
    s.n1 = (h1 + 15) / 16 * 16; s.n2 = (h2 + 15) / 16 * 16;
    s.kb2 = (h1 + 1 + 31) / 32; s.k2p = s.kb2 * 32;
    s.h2p32 = (h2 + 31) / 32 * 32;
    return s;
}

// byte offset of element (row, k) inside one SWIZZLE_64B k-block (rows of 64 B, 16 B chunks XOR (row>>1)&3)
__host__ __device__ inline uint32_t sw64_off(int row, int k) {
    return (uint32_t)row * kRowB + ((((uint32_t)k >> 3) ^ (((uint32_t)row >> 1) & 3u)) << 4) + (((uint32_t)k & 7u) << 1);
}

// ---- weight images (device global), built at load time ----
//   w1img: n1 rows x 64 B                  (k < in_dim: fc1.weight, k == in_dim: fc1.bias, else 0)
//   w2img: kb2 blocks x (n2 rows x 64 B)   (k < h1: fc2.weight, k == h1: fc2.bias, else 0)
__global__ void pack_tc_kernel(tt_actor_dev A, TcShape s, const float *__restrict__ fc1_w, const float *__restrict__ fc1_b,
                               const float *__restrict__ fc2_w, const float *__restrict__ fc2_b) {
    const int tid = blockIdx.x * blockDim.x + threadIdx.x, nth = gridDim.x * blockDim.x;
    char *w1 = reinterpret_cast<char *>(A.w1b), *w2 = reinterpret_cast<char *>(A.w2b);
    for (int v = tid; v < s.n1 * 32; v += nth) {
        const int n = v / 32, k = v - n * 32;
        float x = 0.f;
        if (n < s.h1) x = k < s.in_dim ? fc1_w[n * s.in_dim + k] : (k == s.in_dim ? fc1_b[n] : 0.f);
        *reinterpret_cast<__nv_bfloat16 *>(w1 + sw64_off(n, k)) = __float2bfloat16_rn(x);
    }
    for (int v = tid; v < s.kb2 * s.n2 * 32; v += nth) {
        const int kb = v / (s.n2 * 32), rem = v - kb * s.n2 * 32, n = rem / 32, kk = rem - n * 32, k = kb * 32 + kk;
        float x = 0.f;
        if (n < s.h2) x = k < s.h1 ? fc2_w[n * s.h1 + k] : (k == s.h1 ? fc2_b[n] : 0.f);
        *reinterpret_cast<__nv_bfloat16 *>(w2 + (size_t)kb * s.n2 * kRowB + sw64_off(n, kk)) = __float2bfloat16_rn(x);
    }
}

// ---- PTX wrappers ----
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    uint32_t done;
    do {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(done) : "r"(bar), "r"(parity) : "memory");
    } while (!done);
}
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void *src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// shared-memory matrix descriptor: K-major, SWIZZLE_64B, 8-row groups 512 B apart (SBO), LBO unused
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr & 0x3FFFFu) >> 4);          // start address
    d |= (uint64_t)(512u >> 4) << 32;                  // stride byte offset
    d |= (uint64_t)1 << 46;                            // descriptor version (sm_100)
    d |= (uint64_t)4 << 61;                            // layout type: SWIZZLE_64B
    return d;
}
// instruction descriptor: kind::f16, A = B = bf16 (K-major), D = fp32, M = 128, N = n
__device__ __forceinline__ uint32_t make_idesc(int n) {
    return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(kTileM >> 4) << 24);
}
__device__ __forceinline__ void umma(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
                 ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}

__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float (&v)[32]) {
    uint32_t r[32];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, %17, %18, "
        "%19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
          "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]),
          "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]),
          "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 32; i++) v[i] = __uint_as_float(r[i]);
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float (&v)[32]) {
    uint32_t r[16];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
          "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 16; i++) v[i] = __uint_as_float(r[i]);
#pragma unroll
    for (int i = 16; i < 32; i++) v[i] = 0.f;
}
// load `cnt` (32 or 16) accumulator columns starting at column c0 of this thread's row
__device__ __forceinline__ void tmem_ld_cols(uint32_t taddr_row, int c0, int cnt, float (&v)[32]) {
    if (cnt >= 32) tmem_ld32(taddr_row + (uint32_t)c0, v); else tmem_ld16(taddr_row + (uint32_t)c0, v);
}

__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi) {
    __nv_bfloat162 p = __floats2bfloat162_rn(lo, hi);
    return *reinterpret_cast<uint32_t *>(&p);
}

struct SmemPlan {
    uint32_t x, w1, a2, w2, par, bars, tmem_slot, total;
};
__host__ __device__ inline SmemPlan make_plan(const TcShape &s) {
    SmemPlan p;
    uint32_t off = 0;
    auto take = [&](uint32_t bytes, uint32_t align) { off = (off + align - 1) / align * align; uint32_t o = off; off += bytes; return o; };
    p.x = take(kTileM * kRowB, 1024);
    p.w1 = take((uint32_t)s.n1 * kRowB, 1024);
    p.a2 = take((uint32_t)s.kb2 * kTileM * kRowB, 1024);
    p.w2 = take((uint32_t)kSlots * (((uint32_t)s.n2 * kRowB + 1023) / 1024 * 1024), 1024);
    p.par = take((uint32_t)(2 * s.k2p + 3 * s.h2p32 + 4) * 4, 16);
    p.bars = take(16 * 8, 8);
    p.tmem_slot = take(16, 16);
    p.total = off + 1024;        // slack for the manual 1024 B alignment of the dynamic smem base
    return p;
}

enum { B_W1 = 0, B_XFULL, B_H1FULL, B_A2FULL, B_H2FULL, B_TMEMFREE, B_W2FULL, B_W2EMPTY = B_W2FULL + kSlots, B_COUNT = B_W2EMPTY + kSlots };

__global__ void __launch_bounds__(kThreads, 1) actor_tc_kernel(tt_actor_dev A, TcShape s, const float *__restrict__ obs, int64_t ld,
                                                              int64_t n, float *__restrict__ out) {
    extern __shared__ uint8_t smem_raw[];
    const uint32_t raw = smem_u32(smem_raw);
    const uint32_t base = (raw + 1023u) & ~1023u;
    uint8_t *sm = smem_raw + (base - raw);
    const SmemPlan P = make_plan(s);
    const uint32_t sX = base + P.x, sW1 = base + P.w1, sA2 = base + P.a2, sW2 = base + P.w2, sBar = base + P.bars;
    const uint32_t w2_slot_bytes = ((uint32_t)s.n2 * kRowB + 1023u) / 1024u * 1024u;
    float *par = reinterpret_cast<float *>(sm + P.par);
    float *pg1 = par, *pbe1 = par + s.k2p, *pg2 = par + 2 * s.k2p, *pbe2 = pg2 + s.h2p32, *pw3 = pbe2 + s.h2p32;
    volatile uint32_t *tmem_slot = reinterpret_cast<volatile uint32_t *>(sm + P.tmem_slot);
    auto bar = [&](int i) { return sBar + 8u * (uint32_t)i; };

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int64_t ntiles = (n + kTileM - 1) / kTileM;

    // ---------------- one-time setup ----------------
    if (threadIdx.x == 0) {
        mbar_init(bar(B_W1), 1);
        mbar_init(bar(B_XFULL), 128); mbar_init(bar(B_H1FULL), 1); mbar_init(bar(B_A2FULL), 128);
        mbar_init(bar(B_H2FULL), 1); mbar_init(bar(B_TMEMFREE), 128);
        for (int i = 0; i < kSlots; i++) { mbar_init(bar(B_W2FULL + i), 1); mbar_init(bar(B_W2EMPTY + i), 1); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 4) {     // TMEM allocation (whole warp), 512 columns: one CTA per SM
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(base + P.tmem_slot), "r"(kTmemCols) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    // per-column epilogue parameters -> shared (g = 0 / be = 1 at column h1 makes A2[:, h1] == 1: the fc2-bias column)
    for (int c = threadIdx.x; c < s.k2p; c += kThreads) {
        pg1[c] = c < s.h1 ? A.g1[c] : 0.f;
        pbe1[c] = c < s.h1 ? A.be1[c] : (c == s.h1 ? 1.f : 0.f);
    }
    for (int c = threadIdx.x; c < s.h2p32; c += kThreads) {
        const bool in = c < s.h2;
        pg2[c] = in ? A.g2[c] : 0.f; pbe2[c] = in ? A.be2[c] : 0.f; pw3[c] = in ? A.w3[c] : 0.f;
    }
    // X tile: zero everything once, constant-1 bias column at k = in_dim (both are never overwritten)
    for (int v = threadIdx.x; v < kTileM * kRowB / 4; v += kThreads) reinterpret_cast<uint32_t *>(sm + P.x)[v] = 0u;
    __syncthreads();
    if (threadIdx.x < kTileM)
        *reinterpret_cast<__nv_bfloat16 *>(sm + P.x + sw64_off(threadIdx.x, s.in_dim)) = __float2bfloat16_rn(1.0f);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *tmem_slot;
    const float b3 = A.b3[0];

    if (warp == 5) {
        // ================= bulk-copy producer =================
        if (lane == 0) {
            mbar_expect_tx(bar(B_W1), (uint32_t)s.n1 * kRowB);
            bulk_g2s(sW1, A.w1b, (uint32_t)s.n1 * kRowB, bar(B_W1));
            uint32_t it = 0;
            for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
                for (int kb = 0; kb < s.kb2; kb++, it++) {
                    const uint32_t slot = it % kSlots, ph = (it / kSlots) & 1u;
                    mbar_wait(bar(B_W2EMPTY + slot), ph ^ 1u);          // first pass through the ring passes immediately
                    mbar_expect_tx(bar(B_W2FULL + slot), (uint32_t)s.n2 * kRowB);
                    bulk_g2s(sW2 + slot * w2_slot_bytes, reinterpret_cast<const char *>(A.w2b) + (size_t)kb * s.n2 * kRowB,
                             (uint32_t)s.n2 * kRowB, bar(B_W2FULL + slot));
                }
            }
        }
    } else if (warp == 4) {
        // ================= MMA issuer (one thread) =================
        if (lane == 0) {
            const int nA = s.n1 > 256 ? 256 : s.n1, nB = s.n1 - nA;      // layer-1 N split
            const int mA = s.n2 > 256 ? 256 : s.n2, mB = s.n2 - mA;      // layer-2 N split
            uint32_t it = 0, tcount = 0;
            mbar_wait(bar(B_W1), 0);
            for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x, tcount++) {
                const uint32_t ph = tcount & 1u;
                mbar_wait(bar(B_XFULL), ph);
                mbar_wait(bar(B_TMEMFREE), ph ^ 1u);                      // previous tile's accumulators drained
                tc_fence_after();
#pragma unroll
                for (int ks = 0; ks < 2; ks++) {                          // K = 32 = 2 x UMMA_K(16)
                    const uint64_t a = make_desc(sX + ks * 32);
                    umma(tmem, a, make_desc(sW1 + ks * 32), make_idesc(nA), ks);
                    if (nB > 0) umma(tmem + 256, a, make_desc(sW1 + 256 * kRowB + ks * 32), make_idesc(nB), ks);
                }
                umma_commit(bar(B_H1FULL));
                mbar_wait(bar(B_A2FULL), ph);
                tc_fence_after();
                for (int kb = 0; kb < s.kb2; kb++, it++) {
                    const uint32_t slot = it % kSlots, wph = (it / kSlots) & 1u;
                    mbar_wait(bar(B_W2FULL + slot), wph);
                    tc_fence_after();
                    const uint32_t wb = sW2 + slot * w2_slot_bytes, ab = sA2 + (uint32_t)kb * kTileM * kRowB;
#pragma unroll
                    for (int ks = 0; ks < 2; ks++) {
                        const uint64_t a = make_desc(ab + ks * 32);
                        const uint32_t acc = (kb | ks) ? 1u : 0u;
                        umma(tmem, a, make_desc(wb + ks * 32), make_idesc(mA), acc);
                        if (mB > 0) umma(tmem + 256, a, make_desc(wb + 256 * kRowB + ks * 32), make_idesc(mB), acc);
                    }
                    umma_commit(bar(B_W2EMPTY + slot));                   // frees the W2 slot once these MMAs retire
                }
                umma_commit(bar(B_H2FULL));
            }
        }
    } else {
        // ================= epilogue warps: thread = row =================
        const int r = threadIdx.x;                                        // 0..127, TMEM lane
        const uint32_t trow = tmem + ((uint32_t)(warp * 32) << 16);
        const uint32_t xsw = (((uint32_t)r >> 1) & 3u);
        uint32_t tcount = 0;
        for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x, tcount++) {
            const uint32_t ph = tcount & 1u;
            const int64_t row0 = tile * kTileM;
            const int rows = (int)((n - row0) < kTileM ? (n - row0) : kTileM);
            // ---- observation tile -> bf16, swizzled (coalesced global reads, 2 B shared stores) ----
            if (tcount > 0) mbar_wait(bar(B_H1FULL), ph ^ 1u);            // previous tile's layer-1 MMAs have read X
            for (int v = r; v < kTileM * s.in_dim; v += kTileM) {
                const int rr = v / s.in_dim, k = v - rr * s.in_dim;
                const float x = rr < rows ? __ldcs(obs + (row0 + rr) * ld + k) : 0.f;
                *reinterpret_cast<__nv_bfloat16 *>(sm + P.x + sw64_off(rr, k)) = __float2bfloat16_rn(x);
            }
            fence_proxy_async();
            mbar_arrive(bar(B_XFULL));
            // ---- epilogue 1: LayerNorm + ReLU over h1 columns -> A2 (bf16, UMMA image) ----
            mbar_wait(bar(B_H1FULL), ph);
            tc_fence_after();
            float v[32];
            float sum = 0.f, sq = 0.f;
            for (int c0 = 0; c0 < s.n1; c0 += 32) {
                const int cnt = s.n1 - c0 >= 32 ? 32 : 16;
                tmem_ld_cols(trow, c0, cnt, v);
#pragma unroll
                for (int j = 0; j < 32; j++) { const float x = (c0 + j < s.h1) ? v[j] : 0.f; sum += x; sq = fmaf(x, x, sq); }
            }
            float mean = sum / (float)s.h1;
            float rstd = rsqrtf(fmaxf(sq / (float)s.h1 - mean * mean, 0.f) + 1e-5f);
            for (int c0 = 0; c0 < s.k2p; c0 += 32) {
                const int avail = s.n1 - c0;
                if (avail >= 32) tmem_ld32(trow + (uint32_t)c0, v);
                else if (avail >= 16) tmem_ld16(trow + (uint32_t)c0, v);
                else {
#pragma unroll
                    for (int j = 0; j < 32; j++) v[j] = 0.f;
                }
                uint8_t *blk = sm + P.a2 + (size_t)(c0 >> 5) * kTileM * kRowB + (size_t)r * kRowB;
#pragma unroll
                for (int q = 0; q < 4; q++) {                             // 4 chunks of 8 columns = 16 B each
                    float y[8];
#pragma unroll
                    for (int j = 0; j < 8; j++) {
                        const int c = c0 + q * 8 + j;
                        const float a = rstd * pg1[c];
                        y[j] = fmaxf(fmaf(v[q * 8 + j], a, fmaf(-mean, a, pbe1[c])), 0.f);
                    }
                    uint4 pk;
                    pk.x = pack_bf16x2(y[0], y[1]); pk.y = pack_bf16x2(y[2], y[3]); pk.z = pack_bf16x2(y[4], y[5]); pk.w = pack_bf16x2(y[6], y[7]);
                    *reinterpret_cast<uint4 *>(blk + (((uint32_t)q ^ xsw) << 4)) = pk;
                }
            }
            fence_proxy_async();
            tc_fence_before();
            mbar_arrive(bar(B_A2FULL));
            // ---- epilogue 2: LayerNorm + ReLU over h2 columns, dot with mu.weight, tanh ----
            mbar_wait(bar(B_H2FULL), ph);
            tc_fence_after();
            sum = 0.f; sq = 0.f;
            for (int c0 = 0; c0 < s.n2; c0 += 32) {
                const int cnt = s.n2 - c0 >= 32 ? 32 : 16;
                tmem_ld_cols(trow, c0, cnt, v);
#pragma unroll
                for (int j = 0; j < 32; j++) { const float x = (c0 + j < s.h2) ? v[j] : 0.f; sum += x; sq = fmaf(x, x, sq); }
            }
            mean = sum / (float)s.h2;
            rstd = rsqrtf(fmaxf(sq / (float)s.h2 - mean * mean, 0.f) + 1e-5f);
            float dot = 0.f;
            for (int c0 = 0; c0 < s.n2; c0 += 32) {
                const int cnt = s.n2 - c0 >= 32 ? 32 : 16;
                tmem_ld_cols(trow, c0, cnt, v);
#pragma unroll
                for (int j = 0; j < 32; j++) {
                    const int c = c0 + j;                                 // c < h2p32: padded parameters are 0
                    const float a = rstd * pg2[c];
                    dot = fmaf(fmaxf(fmaf(v[j], a, fmaf(-mean, a, pbe2[c])), 0.f), pw3[c], dot);
                }
            }
            tc_fence_before();
            mbar_arrive(bar(B_TMEMFREE));
            if (r < rows) out[row0 + r] = tanhf(dot + b3);
        }
    }
    // ---------------- teardown ----------------
    __syncwarp();
    tc_fence_before();
    __syncthreads();
    if (warp == 4) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(kTmemCols) : "memory");
    }
}

}  // namespace

namespace tt {

int actor_pack_tc_full(tt_actor *a, const float *fc1_w, const float *fc1_b, const float *fc2_w, const float *fc2_b, cudaStream_t s) {
    const tt_actor_dev &A = a->dev;
    if (A.in_dim > 31) { set_error("tcgen05 actor: in_dim must be <= 31"); return TT_ERR_INVALID; }
    const TcShape sh = make_shape(A.in_dim, A.h1, A.h2);
    pack_tc_kernel<<<128, 256, 0, s>>>(A, sh, fc1_w, fc1_b, fc2_w, fc2_b);
    TT_COUNT_LAUNCH(); TT_LAUNCH_CHECK();
    return TT_OK;
}

int actor_forward_tc(const tt_actor *a, const float *d_obs, int64_t ld, int64_t n, float *d_mu, cudaStream_t st) {
    const tt_actor_dev &A = a->dev;
    const TcShape sh = make_shape(A.in_dim, A.h1, A.h2);
    const SmemPlan P = make_plan(sh);
    if (P.total > 232448u) { set_error("tcgen05 actor: layer sizes need %u B of shared memory (> 227 KB)", P.total); return TT_ERR_INVALID; }
    static bool attr_set = false;
    if (!attr_set) {
        TT_CUDA(cudaFuncSetAttribute(actor_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 232448));
        attr_set = true;
    }
    const int64_t ntiles = (n + kTileM - 1) / kTileM;
    const int grid = (int)(ntiles < sm_count() ? ntiles : sm_count());
    actor_tc_kernel<<<grid, kThreads, P.total, st>>>(A, sh, d_obs, ld, n, d_mu);
    TT_COUNT_LAUNCH(); TT_LAUNCH_CHECK();
    return TT_OK;
}

}  // namespace tt
