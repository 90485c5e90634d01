// tmem_probe.cu -- micro-benchmark of the TMEM read port (tcgen05.ld) on sm_100a.
// Answers: how many bytes/cycle/SM can the epilogue warps of the actor kernel pull out of TMEM, as a function of the
// number of warps per SM sub-partition, the load shape (x16/x32/x64) and whether each load is waited on immediately
// (serial) or one load ahead (pipelined).  Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tmem_probe tmem_probe.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

template <int NREG> __device__ __forceinline__ void tld(uint32_t a, uint32_t *r);

template <> __device__ __forceinline__ void tld<16>(uint32_t a, uint32_t *r) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
                   "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
                 : "r"(a));
}
template <> __device__ __forceinline__ void tld<32>(uint32_t a, uint32_t *r) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,"
                 "%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
                   "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]),
                   "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]),
                   "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
                 : "r"(a));
}
__device__ __forceinline__ void twait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// kWarpsPerQ warps per lane quarter (= per SM sub-partition); each warp reads `iters` chunks of NREG columns.
// mode 0: ld, wait, consume;  mode 1: two buffers, next load issued before the previous one is consumed.
template <int NREG, int MODE>
__global__ void probe(int iters, unsigned long long *out, uint32_t *sink) {
    __shared__ uint32_t slot;
    const int warp = threadIdx.x >> 5;
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"((uint32_t)__cvta_generic_to_shared(&slot)), "r"(512u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tm = slot + ((uint32_t)((warp & 3) * 32) << 16);
    uint32_t a[NREG], b[NREG], acc = 0;
    __syncthreads();
    const long long t0 = clock64();
    if (MODE == 0) {
        for (int i = 0; i < iters; i++) {
            tld<NREG>(tm + (uint32_t)((i * NREG) & (512 - NREG)), a);
            twait();
#pragma unroll
            for (int j = 0; j < NREG; j++) acc ^= a[j];
        }
    } else {
        tld<NREG>(tm, a);
        for (int i = 0; i < iters; i += 2) {
            tld<NREG>(tm + (uint32_t)(((i + 1) * NREG) & (512 - NREG)), b);
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");   // waits for both (no per-load wait exists)
#pragma unroll
            for (int j = 0; j < NREG; j++) acc ^= a[j];
            tld<NREG>(tm + (uint32_t)(((i + 2) * NREG) & (512 - NREG)), a);
#pragma unroll
            for (int j = 0; j < NREG; j++) acc ^= b[j];
            twait();
        }
    }
    __syncthreads();
    const long long t1 = clock64();
    if (threadIdx.x == 0 && blockIdx.x == 0) out[0] = (unsigned long long)(t1 - t0);
    if (acc == 0x12345u) sink[threadIdx.x] = acc;
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(slot), "r"(512u) : "memory");
}

template <int NREG, int MODE>
void run(int warps_per_q, unsigned long long *d_out, uint32_t *d_sink) {
    const int iters = 4096, threads = 128 * warps_per_q;
    probe<NREG, MODE><<<148, threads>>>(iters, d_out, d_sink);
    probe<NREG, MODE><<<148, threads>>>(iters, d_out, d_sink);
    cudaDeviceSynchronize();
    unsigned long long cyc;
    cudaMemcpy(&cyc, d_out, 8, cudaMemcpyDeviceToHost);
    const double bytes = (double)iters * (MODE ? 1.0 : 1.0) * NREG * 4.0 * 32.0 * (threads / 32);
    printf("x%-3d mode %d warps/quarter %d : %8llu cycles  %.1f B/clk/SM  (%.1f cycles per ld per warp)  err=%s\n", NREG, MODE, warps_per_q, cyc,
           bytes / (double)cyc, (double)cyc / iters, cudaGetErrorString(cudaGetLastError()));
}

int main() {
    unsigned long long *d_out; uint32_t *d_sink;
    cudaMalloc(&d_out, 64); cudaMalloc(&d_sink, 4096 * 4);
    for (int w = 1; w <= 4; w *= 2) { run<16, 0>(w, d_out, d_sink); run<32, 0>(w, d_out, d_sink); run<16, 1>(w, d_out, d_sink); run<32, 1>(w, d_out, d_sink); }
    return 0;
}
