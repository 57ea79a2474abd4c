// Shared helpers for libclipb200 (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda_fp16.h>
#include <stdint.h>
#include <stdio.h>
#include <string>

#include "../../../include/clipb200.h"

namespace cb {

void set_error(const char *fmt, ...);
extern thread_local int64_t g_launches;

#define CB_CUDA(expr)                                                              \
    do {                                                                           \
        cudaError_t _e = (expr);                                                   \
        if (_e != cudaSuccess) {                                                   \
            cb::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e),  \
                          __FILE__, __LINE__);                                     \
            return _e == cudaErrorMemoryAllocation ? CB_ERR_OOM : CB_ERR_CUDA;     \
        }                                                                          \
    } while (0)

#define CB_LAUNCH_CHECK()                                                          \
    do {                                                                           \
        cb::g_launches++;                                                          \
        cudaError_t _e = cudaGetLastError();                                       \
        if (_e != cudaSuccess) {                                                   \
            cb::set_error("kernel launch failed: %s (%s:%d)",                      \
                          cudaGetErrorString(_e), __FILE__, __LINE__);             \
            return CB_ERR_CUDA;                                                    \
        }                                                                          \
    } while (0)

#define CB_REQUIRE(cond, ...)                                                      \
    do {                                                                           \
        if (!(cond)) {                                                             \
            cb::set_error(__VA_ARGS__);                                            \
            return CB_ERR_INVALID;                                                 \
        }                                                                          \
    } while (0)

constexpr int kNumSMs = 148;

// RAII device guard
struct DeviceGuard {
    int prev = -1;
    bool ok = true;
    explicit DeviceGuard(int dev) {
        if (cudaGetDevice(&prev) != cudaSuccess) { ok = false; return; }
        if (prev != dev && cudaSetDevice(dev) != cudaSuccess) ok = false;
        cur = dev;
    }
    ~DeviceGuard() {
        if (prev >= 0 && prev != cur) cudaSetDevice(prev);
    }
    int cur = -1;
};

// ---- device helpers -------------------------------------------------------

// streaming 128-bit load: read-only path, do not allocate in L1
__device__ __forceinline__ uint4 ld_stream_v4(const void *p) {
    uint4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w)
                 : "l"(p));
    return r;
}

// monotone float -> uint32 key (larger float <=> larger key); -0.0 folded to +0.0
__device__ __forceinline__ uint32_t f2key(float f) {
    f += 0.0f;
    uint32_t b = __float_as_uint(f);
    return (b & 0x80000000u) ? ~b : (b | 0x80000000u);
}
__device__ __forceinline__ float key2f(uint32_t k) {
    uint32_t b = (k & 0x80000000u) ? (k & 0x7fffffffu) : ~k;
    return __uint_as_float(b);
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

}  // namespace cb
