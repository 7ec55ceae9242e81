// Error channel, launch counter and device queries shared by every entry point of libclipgp.so.
#include <stdarg.h>
#include <atomic>

#include "common.cuh"

namespace clipgp {

static thread_local char g_err[512] = "";
static std::atomic<int64_t> g_launches{0};

void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

int check_launch(const char* what) {
    g_launches.fetch_add(1, std::memory_order_relaxed);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) {
        set_error("%s: kernel launch failed: %s", what, cudaGetErrorString(e));
        return CLIPGP_ERR_CUDA;
    }
    return CLIPGP_OK;
}

int num_sms() {
    static int sms = 0;
    if (sms == 0) {
        int dev = 0;
        if (cudaGetDevice(&dev) != cudaSuccess) return 148;
        if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || sms <= 0) sms = 148;
    }
    return sms;
}

}  // namespace clipgp

extern "C" {
const char* clipgp_last_error(void) { return clipgp::g_err; }
int clipgp_version(void) { return 100; }
int64_t clipgp_launch_count(void) { return clipgp::g_launches.load(); }
}
