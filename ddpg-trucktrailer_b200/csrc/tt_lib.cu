// tt_lib.cu -- library-wide plumbing of libtt_b200.so: thread-local error text, device probing.
#include <stdarg.h>
#include "tt_common.cuh"

namespace tt {

static thread_local char g_err[512] = "";
unsigned long long g_launches = 0;

void set_error(const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof g_err, fmt, ap);
    va_end(ap);
}

int cuda_fail(cudaError_t e, const char *what) {
    set_error("CUDA error %d (%s) at %s", (int)e, cudaGetErrorString(e), what);
    return TT_ERR_CUDA;
}

int sm_count() {
    static int n = 0;
    if (n == 0) {
        int dev = 0;
        if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0)
            n = 148;
    }
    return n;
}

}  // namespace tt

extern "C" {

const char *tt_last_error(void) { return tt::g_err; }
int tt_abi_version(void) { return TT_ABI_VERSION; }

int tt_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
    return n;
}

uint64_t tt_launch_count(void) { return tt::g_launches; }

}  // extern "C"
