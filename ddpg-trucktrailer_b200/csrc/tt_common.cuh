// tt_common.cuh -- shared host-side plumbing of libtt_b200.so (error reporting, launch geometry).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include "../../include/tt_b200.h"

// Where transition i of the current batch lands in the replay ring (ReplayBuffer.store_transition semantics,
// DDPG/replay_buffer.py:13-21: row (mem_cntr + i) % mem_size, last writer wins).  Used by the kernels that write
// their part of the transition straight into the ring (fused store) and by the stand-alone scatter kernel.
struct TTRingMap {
    int64_t cap;      // mem_size
    int64_t base;     // mem_cntr % mem_size
    int64_t first;    // max(0, n - cap): earlier transitions of this batch are overwritten by later ones -> skipped
    int many;         // n > cap: (base + i) may exceed 2 * cap, use a real modulo
    __host__ __device__ __forceinline__ int64_t row(int64_t i) const {
        int64_t r = base + i;
        if (many) r %= cap; else if (r >= cap) r -= cap;
        return r;
    }
};
static inline TTRingMap tt_make_ring_map(int64_t cap, int64_t cntr, int64_t n) {
    TTRingMap m;
    m.cap = cap; m.base = cntr % cap; m.first = n > cap ? n - cap : 0; m.many = n > cap ? 1 : 0;
    return m;
}
struct TTRingS { float *S; TTRingMap m; };                                   // s  rows   (written by the actor kernel)
struct TTRingA { float *A; TTRingMap m; };                                   // a         (written by the noise kernel)
struct TTRingOut { float *S2; float *R; uint8_t *D; TTRingMap m; };          // s', r, d  (written by the env kernel)

namespace tt {

void set_error(const char *fmt, ...);
int cuda_fail(cudaError_t e, const char *what);

#define TT_CUDA(call)                                              \
    do {                                                           \
        cudaError_t _e = (call);                                   \
        if (_e != cudaSuccess) return ::tt::cuda_fail(_e, #call);  \
    } while (0)

#define TT_REQUIRE(cond, msg)                                      \
    do {                                                           \
        if (!(cond)) { ::tt::set_error("%s: %s", __func__, msg); return TT_ERR_INVALID; } \
    } while (0)

// launch check that does NOT synchronise the stream
#define TT_LAUNCH_CHECK()                                          \
    do {                                                           \
        cudaError_t _e = cudaGetLastError();                       \
        if (_e != cudaSuccess) return ::tt::cuda_fail(_e, "kernel launch"); \
    } while (0)

inline cudaStream_t as_stream(tt_stream_t s) { return reinterpret_cast<cudaStream_t>(s); }
inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }
constexpr int kMaxDevices = 64;
int device_index();          // the current CUDA device, as an index into per-device caches
int sm_count();              // SM count of the current device
int grid_sms();              // SMs the persistent rollout kernels may occupy: sm_count() minus tt_reserve_sms()

// number of kernels launched by this library since load (bench.py reports it as gpu_launches)
extern unsigned long long g_launches;
#define TT_COUNT_LAUNCH() (++::tt::g_launches)

// ---- programmatic dependent launch (chains of short kernels: the learner's 15 stages, a rollout iteration at small N) ----
// A kernel launched through launch_chained() may be SCHEDULED while its predecessor in the stream still runs; its first
// statement must be chain_enter(): griddepcontrol.wait blocks until the predecessor grid has completed and its writes are
// visible (so the stream order of all memory accesses is unchanged), launch_dependents then lets the successor be scheduled
// behind this grid.  Every CTA waits unconditionally and before any global access -- a grid that finished without waiting
// would release its successor ahead of its own predecessor.  What is hidden is the launch latency between dependent kernels
// of an EAGER stream: the learner step 153 -> 126 us, a 4 096-env rollout iteration 22.6 -> 20.6 us, i.e. what a CUDA graph
// of the same launches takes (a graph gains nothing more: 127 -> 127, 20.5 -> 20.1 us).  Triggering BEFORE the wait -- the
// whole chain resident at once -- was slower (learner graph 127 -> 144 us).  `chained = false` (and TT_NO_PDL=1 in the
// environment) launches normally.
bool chained_launches_enabled();
// The two kernels of a rollout iteration are chained only where the launch gap is a visible share of the iteration; a large
// batch gains nothing (0.3 %), and early-scheduled CTAs of the next kernel would sit on the SMs tt_reserve_sms() keeps free.
inline bool chain_rollout(int64_t n_envs) { return n_envs <= 262144 && chained_launches_enabled(); }
template <typename... KArgs, typename... Args>
inline cudaError_t launch_chained(bool chained, void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t s, Args &&...args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = s;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at; cfg.numAttrs = chained && chained_launches_enabled() ? 1 : 0;
    return cudaLaunchKernelEx(&cfg, kern, static_cast<KArgs>(args)...);
}

}  // namespace tt

#ifdef __CUDACC__
__device__ __forceinline__ void chain_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void chain_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void chain_enter() { chain_wait(); chain_trigger(); }
#endif
