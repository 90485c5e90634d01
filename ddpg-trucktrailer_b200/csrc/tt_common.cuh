// tt_common.cuh -- shared host-side plumbing of libtt_b200.so (error reporting, launch geometry).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include "../../include/tt_b200.h"

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
int sm_count();

// number of kernels launched by this library since load (bench.py reports it as gpu_launches)
extern unsigned long long g_launches;
#define TT_COUNT_LAUNCH() (++::tt::g_launches)

}  // namespace tt
