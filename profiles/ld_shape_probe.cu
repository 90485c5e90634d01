// ld_shape_probe.cu -- which (lane, column) does every register of tcgen05.ld.16x256b.xN return?  TMEM is filled through
// 32x32b stores with value = 1000 * lane + column, then read back with the 16x256b shape; prints the mapping of warp 1.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o bin/ld_shape_probe ld_shape_probe.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__global__ void probe(uint32_t *out) {
    __shared__ uint32_t slot;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"((uint32_t)__cvta_generic_to_shared(&slot)), "r"(64u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tm = slot + ((uint32_t)(warp * 32) << 16);
    for (int c0 = 0; c0 < 32; c0 += 8) {
        uint32_t v[8];
        for (int j = 0; j < 8; j++) v[j] = 1000u * (uint32_t)(warp * 32 + lane) + (uint32_t)(c0 + j);
        asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"r"(tm + c0), "r"(v[0]), "r"(v[1]), "r"(v[2]),
                     "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]) : "memory");
    }
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    uint32_t r[16];
    // x1 at lane offset 0, x1 at lane offset 16, x2 (16 columns) at lane offset 0 starting at column 8
    asm volatile("tcgen05.ld.sync.aligned.16x256b.x1.b32 {%0,%1,%2,%3}, [%4];" : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(tm));
    asm volatile("tcgen05.ld.sync.aligned.16x256b.x1.b32 {%0,%1,%2,%3}, [%4];" : "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]) : "r"(tm + (16u << 16)));
    asm volatile("tcgen05.ld.sync.aligned.16x256b.x2.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]) : "r"(tm + 8u));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    if (warp == 1) for (int j = 0; j < 16; j++) out[lane * 16 + j] = r[j];
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(slot), "r"(64u) : "memory");
}

int main() {
    uint32_t *d, h[32 * 16];
    cudaMalloc(&d, sizeof(h));
    probe<<<1, 128>>>(d);
    printf("%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
    cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
    for (int t = 0; t < 32; t++) {
        printf("thread %2d:", t);
        for (int j = 0; j < 16; j++) printf(" %s(%3u,%2u)", j == 4 || j == 8 ? "| " : "", h[t * 16 + j] / 1000, h[t * 16 + j] % 1000);
        printf("\n");
    }
    return 0;
}
