// cta2_probe.cu -- ground work for the next actor kernel: one tcgen05.mma.cta_group::2 (CTA pair, M = 256 = 2 x 128 rows) with
// known operands, checked on the host.  Pins down the semantics the kernel will rely on:
//   * TMEM allocation with .cta_group::2 issued by one warp of EACH CTA of the pair,
//   * A: every CTA supplies its own 128 rows from ITS shared memory at the descriptor's offset,
//   * B: every CTA supplies HALF of the N rows (rank 0: rows 0 .. N/2-1, rank 1: rows N/2 .. N-1),
//   * D: every CTA finds its 128 rows x N columns in ITS tensor memory,
//   * only the leader issues the MMA; tcgen05.commit.cta_group::2 ... multicast::cluster arrives on both CTAs' mbarriers.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I../ddpg-trucktrailer_b200/csrc -o bin/cta2_probe cta2_probe.cu
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <cmath>
#include <vector>
#include <cuda_runtime.h>
#include <cuda_fp16.h>
#include "tt_tc_ptx.cuh"

constexpr int kN = 32;          // total N of the pair's MMA; each CTA holds kN / 2 rows of B

__host__ __device__ inline float a_val(int grow, int k) { return (float)((grow * 3 + k * 5) % 17 - 8) / 8.0f; }
__host__ __device__ inline float b_val(int n, int k) { return (float)((n * 7 + k * 11) % 13 - 6) / 4.0f; }

__device__ __forceinline__ uint32_t make_idesc_m(int m, int n) {      // kind::f16, f16 operands, fp32 accumulate, K-major A and B
    return (1u << 4) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(m >> 4) << 24);
}

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(128, 1) probe(float *out) {
    extern __shared__ __align__(1024) uint8_t sm[];
    __shared__ uint32_t slot;
    __shared__ uint64_t bar;
    const uint32_t rank = cluster_ctarank();
    const int warp = threadIdx.x >> 5;
    uint8_t *sA = sm, *sB = sm + 8192;                    // A: 128 rows x 64 B; B half: kN/2 rows x 64 B (SWIZZLE_64B k-blocks of 32)
    for (int v = threadIdx.x; v < 128 * 32; v += blockDim.x) {
        const int r = v >> 5, k = v & 31;
        *reinterpret_cast<__half *>(sA + sw64_off(r, k)) = __float2half_rn(a_val((int)rank * 128 + r, k));
    }
    for (int v = threadIdx.x; v < (kN / 2) * 32; v += blockDim.x) {
        const int n = v >> 5, k = v & 31;
        *reinterpret_cast<__half *>(sB + sw64_off(n, k)) = __float2half_rn(b_val((int)rank * (kN / 2) + n, k));
    }
    if (threadIdx.x == 0) { mbar_init(smem_u32(&bar), 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&slot)), "r"(32u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    }
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    cluster_sync_all();
    tc_fence_after();
    const uint32_t tmem = slot;
    if (rank == 0 && warp == 1) {
        if (elect_one()) {
            const uint32_t id = make_idesc_m(256, kN);
            const uint64_t dA = make_desc(smem_u32(sA)), dB = make_desc(smem_u32(sB));
            for (int ks = 0; ks < 2; ks++)
                asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
                             ::"r"(tmem), "l"(dA + (uint64_t)(2 * ks)), "l"(dB + (uint64_t)(2 * ks)), "r"(id), "r"((uint32_t)ks) : "memory");
            asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                         ::"r"(smem_u32(&bar)), "h"((uint16_t)3) : "memory");
        }
        __syncwarp();
    }
    mbar_wait(smem_u32(&bar), 0);
    tc_fence_after();
    uint32_t v[32];
    tmem_ld32_async(tmem + ((uint32_t)(warp * 32) << 16), v);
    tmem_wait();
    const int row = (int)rank * 128 + threadIdx.x;
    for (int c = 0; c < kN; c++) out[row * kN + c] = __uint_as_float(v[c]);
    if (threadIdx.x == 0) out[256 * kN + rank] = (float)tmem;
    tc_fence_before();
    __syncthreads();
    cluster_sync_all();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(32u) : "memory");
}

int main() {
    float *d_out;
    cudaMalloc(&d_out, (256 * kN + 2) * sizeof(float));
    cudaMemset(d_out, 0xff, (256 * kN + 2) * sizeof(float));
    cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 16384);
    probe<<<2, 128, 16384>>>(d_out);
    cudaError_t e = cudaDeviceSynchronize();
    printf("launch: %s\n", cudaGetErrorString(e));
    std::vector<float> h(256 * kN + 2);
    cudaMemcpy(h.data(), d_out, h.size() * sizeof(float), cudaMemcpyDeviceToHost);
    double worst = 0.0; int bad = 0;
    for (int m = 0; m < 256; m++)
        for (int n = 0; n < kN; n++) {
            double ref = 0.0;
            for (int k = 0; k < 32; k++) ref += (double)__half2float(__float2half_rn(a_val(m, k))) * (double)__half2float(__float2half_rn(b_val(n, k)));
            const double err = fabs(ref - h[m * kN + n]);
            if (err > worst) worst = err;
            if (err > 1e-3 && bad++ < 6) printf("  D[%d][%d] = %f, expected %f\n", m, n, h[m * kN + n], ref);
        }
    printf("tmem base: leader %.0f peer %.0f;  max |err| = %.3e over 256 x %d  -> %s\n", h[256 * kN], h[256 * kN + 1], worst, kN, bad ? "MISMATCH" : "OK");
    return 0;
}
