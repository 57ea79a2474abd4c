// Misc C-ABI entry points: error string, version, device count, launch counter.
#include "common.cuh"

#include <cstdarg>

namespace cb {

static thread_local char g_err[1024] = "";
thread_local int64_t g_launches = 0;

void set_error(const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

}  // namespace cb

extern "C" {

const char *cb_last_error(void) { return cb::g_err; }

int cb_abi_version(void) { return 1; }

int cb_device_count(int *n) {
    CB_REQUIRE(n != nullptr, "cb_device_count: null out pointer");
    *n = 0;
    cudaError_t e = cudaGetDeviceCount(n);
    if (e != cudaSuccess) {
        cudaGetLastError();
        *n = 0;
        cb::set_error("cudaGetDeviceCount: %s", cudaGetErrorString(e));
        return CB_ERR_NOGPU;
    }
    return CB_OK;
}

int64_t cb_launch_count(int reset) {
    int64_t v = cb::g_launches;
    if (reset) cb::g_launches = 0;
    return v;
}

}  // extern "C"
