// tt_lib.cu -- library-wide plumbing of libtt_b200.so: thread-local error text, device probing.
#include <stdarg.h>
#include <stdlib.h>
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

bool chained_launches_enabled() {
    static const bool on = [] { const char *e = getenv("TT_NO_PDL"); return !(e && e[0] && e[0] != '0'); }();
    return on;
}

// Launch facts are cached PER DEVICE: a process may drive several GPUs (every Python class takes a device argument), and
// the opt-in shared-memory attributes, occupancy figures and SM counts belong to the device that is current at the launch.
int device_index() {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) { (void)cudaGetLastError(); dev = 0; }
    return dev >= 0 && dev < kMaxDevices ? dev : 0;
}

int sm_count() {
    static int n[kMaxDevices] = {};
    const int dev = device_index();
    if (n[dev] == 0) {
        if (cudaDeviceGetAttribute(&n[dev], cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n[dev] <= 0) n[dev] = 148;
    }
    return n[dev];
}

// SMs the rollout kernels leave free (per device): their persistent grids are sized for grid_sms() instead of sm_count(), so
// that the small dependent kernels of a learner / of NCCL on a side stream find room while a rollout kernel runs.
static int g_reserved[kMaxDevices] = {};
int grid_sms() {
    const int n = sm_count() - g_reserved[device_index()];
    return n > 0 ? n : 1;
}

}  // namespace tt

extern "C" {

int tt_reserve_sms(int32_t n) {
    if (n < 0 || n > tt::sm_count() / 2) { tt::set_error("tt_reserve_sms: n must be in 0 .. %d", tt::sm_count() / 2); return TT_ERR_INVALID; }
    tt::g_reserved[tt::device_index()] = n;
    return TT_OK;
}

const char *tt_last_error(void) { return tt::g_err; }
int tt_abi_version(void) { return TT_ABI_VERSION; }

int tt_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
    return n;
}

uint64_t tt_launch_count(void) { return tt::g_launches; }

}  // extern "C"
