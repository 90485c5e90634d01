// tt_tc4_layout.cuh -- operand-image layout, TMEM column map and W2 ring schedule of the v4 tensor-core actor kernel
// (tt_actor_tc4.cu packs the images and consumes them).
#pragma once
#include "tt_tc_ptx.cuh"

namespace {

constexpr int kStatRows = 32;                 // statistic rows appended to the layer-1 B image (24 used)
constexpr int N1I = N1 + kStatRows;           // rows per hi / lo block of the v4 layer-1 image (432)
constexpr int kParts = 3;                     // layer-1 parts: [32 statistics + 96] | 160 | 144 columns
constexpr int kWin0 = 304;                    // first TMEM column of the layer-1 window (160 columns)
constexpr int kXCol0 = kWin0 + 160;             // observation tile (layer-1 A operand): 16 columns hi [464, 480) + 16 columns lo [480, 496)
constexpr int kNA = 160, kNB = N2 - kNA;      // layer-2 output halves (sweep A / sweep B)
constexpr uint32_t kW2SlotB = kNA * kRowB;    // ring slot: one half k-block (10 240 B; sweep B uses 9 216 of it)
constexpr size_t kW2SweepB = (size_t)KB2 * kNA * kRowB;   // byte offset of sweep B inside the v4 W2 image
constexpr size_t kW2ImageB = (size_t)KB2 * N2 * kRowB;     // one replica of the image

__device__ __forceinline__ void tmem_ld4_async(uint32_t taddr, uint32_t (&r)[4]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0, %1, %2, %3}, [%4];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(taddr));
}
__device__ __forceinline__ void tmem_st4(uint32_t taddr, const uint32_t (&r)[4]) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x4.b32 [%0], {%1, %2, %3, %4};" ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]) : "memory");
}
__device__ __forceinline__ void tmem_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
// D[tmem] (+)= A[tmem] . B[smem]: the A operand (128 lanes x 8 columns of packed 16-bit pairs per K = 16 step) read from tensor memory
__device__ __forceinline__ void umma_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
                 ::"r"(d_tmem), "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(accumulate) : "memory");
}
// 16 lanes x 256 bit: thread t gets, per 8-column octet, (row t / 4, columns 2 (t % 4), + 1) and (row t / 4 + 8, same columns)
__device__ __forceinline__ void tmem_ld16x256_x1(uint32_t taddr, uint32_t (&r)[4]) {
    asm volatile("tcgen05.ld.sync.aligned.16x256b.x1.b32 {%0, %1, %2, %3}, [%4];" : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld16x256_x4(uint32_t taddr, uint32_t (&r)[16]) {
    asm volatile("tcgen05.ld.sync.aligned.16x256b.x4.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
                   "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
                 : "r"(taddr));
}
__device__ __forceinline__ float2 lds64_early(const float *p) {        // like lds128_early (tt_tc_ptx.cuh): a load the compiler may not sink
    float2 v;
    asm volatile("ld.shared.v2.f32 {%0, %1}, [%2];" : "=f"(v.x), "=f"(v.y) : "r"(smem_u32(p)));
    return v;
}
template <int NR>
__device__ __forceinline__ void tmem_ld8_async(uint32_t taddr, uint32_t (&r)[NR]) {      // 8 columns into r[0..7], rest zero
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
                 : "r"(taddr));
#pragma unroll
    for (int i = 8; i < NR; i++) r[i] = 0u;
}

// W2 ring schedule: one tile = 26 steps (13 k-blocks of half A, 13 of half B, order below); step -> slot is a fixed compile-time
// pattern (round-robin, the last 26 % kSlots steps reuse slots 0, 1, ...) so that the fully unrolled issue loops carry no
// address or phase arithmetic.  A slot is used kSteps / kSlots (+ 1 for the first 26 % kSlots slots) times per tile.
constexpr int kSteps = 2 * KB2;
// Step order within a tile.  The two output halves are separate accumulators (so that the statistics of half A can be taken
// while half B is still being computed), and their k-blocks are INTERLEAVED: half A leads by kLead k-blocks, then the steps
// alternate B(j), A(j + kLead), and half B finishes alone.  A2 block j is free for the next tile's epilogue 1 after B(j),
// i.e. from step kLead + 2 j + 1 on, and half A is complete at step 26 - kLead.  kLead trades the two: a small lead releases
// the A2 blocks early, a large one completes half A early enough for its statistics pass to fill the epilogue's wait for the
// layer-1 MMAs of part 1 (kP1A in tt_actor_tc4.cu).  Measured (profiles/build_variants.sh): kLead 9..13 with the pass in that
// gap are 3 % faster than kLead 2..8 with the pass at the end; 13 = two back-to-back sweeps.
#ifndef TT_TC4_LEAD
#define TT_TC4_LEAD 10
#endif
constexpr int kLead = TT_TC4_LEAD;
static_assert(kLead >= 1 && kLead <= KB2, "sweep A leads by 1..13 k-blocks");
__host__ __device__ constexpr int w2_step_sweep(int step) {
    return step < kLead ? 0 : step >= kSteps - kLead ? 1 : ((step - kLead) & 1) ? 0 : 1;
}
__host__ __device__ constexpr int w2_step_kb(int step) {
    return step < kLead ? step : step >= kSteps - kLead ? step - KB2 : ((step - kLead) & 1) ? kLead + (step - kLead) / 2 : (step - kLead) / 2;
}
__host__ __device__ constexpr int w2_slot(int step, int nslots) { return step < (kSteps / nslots) * nslots ? step % nslots : step - (kSteps / nslots) * nslots; }
__host__ __device__ constexpr int w2_uses_per_tile(int slot, int nslots) { return kSteps / nslots + (slot < kSteps % nslots ? 1 : 0); }
// phase parity of the `step`-th ring use of tile number `tile_parity` (0/1)
__device__ __forceinline__ uint32_t w2_parity(int step, int nslots, uint32_t tile_parity) {
    return ((w2_uses_per_tile(w2_slot(step, nslots), nslots) & 1) ? tile_parity : 0u) ^ (uint32_t)((step / nslots) & 1);
}

}  // namespace
