// tt_tc_ptx.cuh -- shared pieces of the tcgen05 actor kernels (tt_actor_tc.cu, tt_actor_tc4.cu): layer-size constants,
// the SWIZZLE_64B operand addressing, operand-type conversions and the inline-PTX wrappers (mbarrier, bulk copy,
// tcgen05.mma / commit / ld, descriptors).
#pragma once
#include <cuda_fp16.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <type_traits>

namespace {

constexpr int kTileM = 128;
constexpr int kRowB = 64;          // bytes per operand row in a k-block (32 x 16-bit), SWIZZLE_64B
constexpr uint32_t kTmemCols = 512;

// ---- the layer sizes this kernel is specialised to (the reference's defaults, DDPG/trainv2.py:404-408) ----
constexpr int IN = 23, H1 = 400, H2 = 300;
constexpr int N1 = 400;            // layer-1 MMA N (H1 rounded up to 16)
constexpr int N2 = 304;            // layer-2 MMA N (H2 rounded up to 16)
constexpr int KB2 = 13;            // layer-2 k-blocks of 32: ceil((H1 + 1) / 32); column H1 carries the fc2 bias
constexpr int K2P = KB2 * 32;      // 416
constexpr int NCH1 = K2P / 32;     // 13 column chunks of A2 (chunk 12 = 16 accumulator columns + constants)
constexpr int NCH2 = (N2 + 31) / 32;   // 10 column chunks of H2 (chunk 9 = 16 columns)
constexpr int H2P = 320;           // parameter array length for layer 2

// byte offset of element (row, k) inside one SWIZZLE_64B k-block (rows of 64 B, 16 B chunks XOR (row>>1)&3)
__host__ __device__ inline uint32_t sw64_off(int row, int k) {
    return (uint32_t)row * kRowB + ((((uint32_t)k >> 3) ^ (((uint32_t)row >> 1) & 3u)) << 4) + (((uint32_t)k & 7u) << 1);
}

template <typename T> __device__ __forceinline__ T to_op(float x);
template <> __device__ __forceinline__ __half to_op<__half>(float x) { return __float2half_rn(x); }
template <> __device__ __forceinline__ __nv_bfloat16 to_op<__nv_bfloat16>(float x) { return __float2bfloat16_rn(x); }
__device__ __forceinline__ float op_to_float(__half x) { return __half2float(x); }
__device__ __forceinline__ float op_to_float(__nv_bfloat16 x) { return __bfloat162float(x); }

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
// try_wait parks the thread until the phase completes or the suspend-time hint expires (it is woken at once when the phase
// completes).  Without a hint the limit is short and a waiting warp keeps re-issuing SYNCS + BRA, taking issue slots from the
// epilogue warps on its scheduler; 10 M ns is what CUTLASS passes.
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    uint32_t done;
    do {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(done) : "r"(bar), "r"(parity), "r"(0x989680u) : "memory");
    } while (!done);
}
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void *src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}
// One elected lane of a fully converged warp.  The MMA / bulk-copy issuing warps run their loops warp-uniformly and wrap
// only the tcgen05 / cp.async.bulk instructions in `if (elect_one())`: the compiler then keeps descriptors and addresses
// in uniform registers.  Inside an `if (lane == 0)` region it cannot, and wraps EVERY UTCHMMA / UTCBAR / UBLKCP in an
// R2UR + ELECT + BRA.U.ANY loop -- measured 165-210 cycles of issue time per MMA (profiles/mma_probe.cu).
__device__ __forceinline__ bool elect_one() {
    uint32_t pred = 0;
    asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
    return pred != 0;
}
// 16 B shared-memory load the compiler may not sink towards its first use (volatile): the epilogues issue their parameter
// loads two column quads ahead of the math; nvcc otherwise re-schedules them right in front of the consumer to save
// registers and exposes the full LDS latency on every quad (4 warps per scheduler cannot hide it).
__device__ __forceinline__ float4 lds128_early(const float *p) {
    float4 v;
    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(smem_u32(p)));
    return v;
}
// ---- thread-block clusters: W2 k-blocks are fetched once per CTA pair and multicast into both CTAs' shared memory ----
__device__ __forceinline__ uint32_t cluster_ctarank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// bulk copy global -> the SAME shared-memory offset of every CTA in `mask`; completes `bytes` on each CTA's mbarrier at `bar`
__device__ __forceinline__ void bulk_g2s_mc(uint32_t dst, const void *src, uint32_t bytes, uint32_t bar, uint16_t mask) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1], %2, [%3], %4;"
                 ::"r"(dst), "l"(src), "r"(bytes), "r"(bar), "h"(mask) : "memory");
}
// tcgen05.commit that arrives on the mbarrier at the same offset in every CTA of `mask`
__device__ __forceinline__ void umma_commit_mc(uint32_t bar, uint16_t mask) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(bar), "h"(mask) : "memory");
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void named_bar_sync(int id, int nthreads) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory"); }

// shared-memory matrix descriptor: K-major, SWIZZLE_64B, 8-row groups 512 B apart (SBO), LBO unused
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr & 0x3FFFFu) >> 4);          // start address
    d |= (uint64_t)(512u >> 4) << 32;                  // stride byte offset
    d |= (uint64_t)1 << 46;                            // descriptor version (sm_100)
    d |= (uint64_t)4 << 61;                            // layout type: SWIZZLE_64B
    return d;
}
// instruction descriptor: kind::f16, A and B 16-bit K-major (fmt 0 = f16, 1 = bf16), D = fp32, M = 128, N = n
__device__ __forceinline__ uint32_t make_idesc(int n, uint32_t fmt) {
    return (1u << 4) | (fmt << 7) | (fmt << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(kTileM >> 4) << 24);
}
__device__ __forceinline__ void umma(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
                 ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}

// asynchronous TMEM loads of this thread's row: 32 / 16 consecutive accumulator columns; tmem_wait() completes them
__device__ __forceinline__ void tmem_ld32_async(uint32_t taddr, uint32_t (&r)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, %17, %18, "
        "%19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
          "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]),
          "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]),
          "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld16_async(uint32_t taddr, uint32_t (&r)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
          "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr));
#pragma unroll
    for (int i = 16; i < 32; i++) r[i] = 0u;
}
__device__ __forceinline__ void tmem_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// chunk `ch` (32 columns) of an accumulator that is `ncols` wide (ncols % 16 == 0)
template <int NCOLS>
__device__ __forceinline__ void tmem_ld_chunk(uint32_t trow, int ch, uint32_t (&r)[32]) {
    const int c0 = ch * 32;
    if (NCOLS - c0 >= 32) tmem_ld32_async(trow + (uint32_t)c0, r);
    else if (NCOLS - c0 >= 16) tmem_ld16_async(trow + (uint32_t)c0, r);
    else {
#pragma unroll
        for (int i = 0; i < 32; i++) r[i] = 0u;
    }
}

template <typename OpT> __device__ __forceinline__ uint32_t pack2(float lo, float hi);
template <> __device__ __forceinline__ uint32_t pack2<__half>(float lo, float hi) {
    __half2 p = __floats2half2_rn(lo, hi);
    return *reinterpret_cast<uint32_t *>(&p);
}
template <> __device__ __forceinline__ uint32_t pack2<__nv_bfloat16>(float lo, float hi) {
    __nv_bfloat162 p = __floats2bfloat162_rn(lo, hi);
    return *reinterpret_cast<uint32_t *>(&p);
}

// ReLU and round two floats to the packed operand type in ONE instruction (F2FP.RELU...PACK_AB; NaN stays NaN like torch's relu)
template <typename OpT> __device__ __forceinline__ uint32_t pack2_relu(float lo, float hi);
template <> __device__ __forceinline__ uint32_t pack2_relu<__half>(float lo, float hi) {
    uint32_t r;
    asm("cvt.rn.relu.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
    return r;
}
template <> __device__ __forceinline__ uint32_t pack2_relu<__nv_bfloat16>(float lo, float hi) {
    uint32_t r;
    asm("cvt.rn.relu.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
    return r;
}

}  // namespace
