// mma_probe.cu -- micro-benchmark: cycles per tcgen05.mma (cta_group::1, kind::f16, M = 128, K = 16, operands in shared
// memory, SWIZZLE_64B K-major) as a function of N, alone and with concurrent shared-memory traffic from 16 other warps.
// Tells whether the actor kernel's small-N MMAs (N = 16/32/48/64) carry a fixed per-instruction cost and how much the
// tensor core's operand fetch competes with LDS/STS for the shared-memory port.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I../ddpg-trucktrailer_b200/csrc -o bin/mma_probe mma_probe.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include "tt_tc_ptx.cuh"

// mode 0: MMAs only.  mode 1: + 16 warps doing broadcast LDS.128.  mode 2: + 16 warps doing conflict-free STS.128.
__global__ void __launch_bounds__(640, 1) probe(int nmma, int N, int reps, int mode, int commit_every, unsigned long long *out) {
    extern __shared__ uint8_t smem_raw[];
    const uint32_t raw = smem_u32(smem_raw);
    const uint32_t base = (raw + 1023u) & ~1023u;
    uint8_t *sm = smem_raw + (base - raw);
    __shared__ uint32_t slot;
    __shared__ uint64_t bars[4];
    __shared__ volatile int stop;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int i = threadIdx.x; i < (180 * 1024) / 4; i += blockDim.x) reinterpret_cast<uint32_t *>(sm)[i] = 0u;
    if (threadIdx.x == 0) { mbar_init(smem_u32(&bars[0]), 1); mbar_init(smem_u32(&bars[1]), 1); mbar_init(smem_u32(&bars[2]), 1); mbar_init(smem_u32(&bars[3]), 1); stop = 0; asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
    if (warp == 16) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&slot)), "r"(512u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = slot;
    if (warp == 16) {      // warp-uniform issue loop, one elected lane issues (see tt_tc_ptx.cuh: elect_one)
        const uint32_t id = make_idesc(N, 0u);
        const uint32_t sA = base, sB = base + 13 * 8192;       // A: 13 blocks of 128 rows x 64 B, B: slots of 10 KB
        const uint64_t dA = make_desc(sA), dB = make_desc(sB);
        uint32_t ph = 0;
        long long best = 1ll << 60;
        for (int rep = 0; rep < reps; rep++) {
            const long long t0 = clock64();
            if (commit_every >= 0) {                 // a commit (nobody waits on it) after every `commit_every` (power of two) MMAs
                for (int i = 0; i < nmma; i++) {
                    if (elect_one()) {
                        umma(tmem, dA + (uint64_t)((i & 1) * 2), dB + (uint64_t)((i & 1) * 2), id, 1u);
                        if (commit_every && (i & (commit_every - 1)) == commit_every - 1) umma_commit(smem_u32(&bars[1]));
                    }
                    __syncwarp();
                }
            } else {                                  // the actor kernel's k-block step: wait on a (complete) barrier, fence, 2 MMAs on
                for (int i = 0; i < nmma; i += 2) {   // moving operand addresses, 1 or 2 commits
                    mbar_wait(smem_u32(&bars[2]), 1u);                                 // fresh barrier, parity 1: passes at once
                    tc_fence_after();
                    const uint64_t a = dA + (uint64_t)(((i >> 1) % 13) * 512), b = dB + (uint64_t)(((i >> 1) & 3) * 640);
                    if (elect_one()) {
                        umma(tmem, a, b, id, 1u);
                        umma(tmem, a + 2, b + 2, id, 1u);
                        umma_commit(smem_u32(&bars[1]));
                        if (commit_every == -2) umma_commit(smem_u32(&bars[3]));
                    }
                    __syncwarp();
                }
            }
            if (elect_one()) umma_commit(smem_u32(&bars[0]));
            __syncwarp();
            mbar_wait(smem_u32(&bars[0]), ph); ph ^= 1u;
            const long long t1 = clock64();
            if (t1 - t0 < best) best = t1 - t0;
        }
        if (blockIdx.x == 0 && lane == 0) out[0] = (unsigned long long)best;
        if (lane == 0) stop = 1;
    } else if (warp < 16 && mode != 0) {
        float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
        const float4 *p = reinterpret_cast<const float4 *>(sm + 160 * 1024);
        float4 *q = reinterpret_cast<float4 *>(sm + 160 * 1024 + 4096) + threadIdx.x;
        int i = 0;
        while (!stop) {
            if (mode == 1) {
#pragma unroll
                for (int j = 0; j < 16; j++) { const float4 v = p[(i + j) & 63]; acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w; }
            } else {
#pragma unroll
                for (int j = 0; j < 16; j++) { q[0] = acc; acc.x += 1.f; }
            }
            i += 16;
        }
        if (acc.x == 123.456f) out[1] = 1;
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 16) { tc_fence_after(); asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512u) : "memory"); }
}

int main() {
    unsigned long long *d_out;
    cudaMalloc(&d_out, 64);
    cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    const int Ns[] = {16, 32, 48, 64, 96, 128, 144, 160, 192, 208, 256};
    for (int mode = 0; mode < 1; mode++) {
        for (int N : Ns) {
            for (int nmma : {8, 64}) {
                probe<<<148, 640, 200 * 1024>>>(nmma, N, 20, mode, 0, d_out);
                cudaDeviceSynchronize();
                unsigned long long cyc = 0;
                cudaMemcpy(&cyc, d_out, 8, cudaMemcpyDeviceToHost);
                printf("mode %d N %3d nmma %2d : %6llu cycles  %.1f per MMA  (math floor %d)  %s\n", mode, N, nmma, cyc, (double)cyc / nmma, N / 2,
                       cudaGetErrorString(cudaGetLastError()));
            }
        }
    }
    for (int ce : {0, 1, 2, 4, 8, -1, -2}) {      // cost of a tcgen05.commit after every `ce` MMAs (N = 160, 64 MMAs)
        probe<<<148, 640, 200 * 1024>>>(64, 160, 20, 0, ce, d_out);
        cudaDeviceSynchronize();
        unsigned long long cyc = 0;
        cudaMemcpy(&cyc, d_out, 8, cudaMemcpyDeviceToHost);
        printf("N 160 nmma 64 commit_every %d : %6llu cycles  %.1f per MMA\n", ce, cyc, (double)cyc / 64);
    }
    return 0;
}
