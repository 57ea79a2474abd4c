// Shared helpers for libclipb200 (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda_fp16.h>
#include <stdint.h>
#include <stdio.h>
#include <string>
#include <utility>

#include "../../../include/clipb200.h"

namespace cb {

void set_error(const char *fmt, ...);
extern thread_local int64_t g_launches;

// Process-wide tuning knobs (cb_tuning_set / cb_tuning_get, clipb200.h).  Each is read from its
// CLIPB200_* environment variable ONCE, when the library is loaded; -1 = unset (built-in default).
// Launch paths read a relaxed atomic, never the environment.
enum Tune : int {
    T_GEMM_BN = 0,        // CLIPB200_GEMM_BN            force the GEMM N tile (64 / 128 / 192 / 256)
    T_GEMM_NCTA,          // CLIPB200_GEMM_NCTA          force the GEMM cluster size (1 / 2)
    T_GEMM_STAGES,        // CLIPB200_GEMM_STAGES        cap the TMA ring depth
    T_GEMM_RASTER,        // CLIPB200_GEMM_RASTER        tile rasterisation (0 n-fastest, 1 m-fastest, g>=2 grouped)
    T_BATCH_MIN_NQ,       // CLIPB200_BATCH_MIN_NQ       smallest query batch served by the tensor-core search
    T_NO_GRAPH,           // CLIPB200_NO_GRAPH           1: never replay small forward passes as CUDA graphs
    T_LN_BLOCKS_PER_SM,   // CLIPB200_LN_BLOCKS_PER_SM   LayerNorm grid size
    T_LN_FOLD,            // CLIPB200_LN_FOLD            0: keep ln_1 / ln_2 as separate launches; 1: force the fold;
                          //                             unset: fold, checked by the calibration pass of cb_clip_finalize
    T_GEMM_DEBUG,         // CLIPB200_GEMM_DEBUG         (only in -DCLIPB200_EXPERIMENTS builds) result-corrupting probes
    T_SKIP,               // CLIPB200_SKIP               (only in -DCLIPB200_EXPERIMENTS builds)
    T_PDL,                // CLIPB200_PDL                0: no programmatic dependent launch between the towers' kernels
    T_GEMM_SKINNY,        // CLIPB200_GEMM_SKINNY        0: single-row-block GEMMs (M <= 128) use the general kernel; n > 0: force split n
    T_GEMM_RESID_STAGES,  // CLIPB200_GEMM_RESID_STAGES  ring depth of the residual-epilogue GEMMs (their staging boxes leave less room)
    T_ATTN_TC,            // CLIPB200_ATTN_TC            0: vision-tower attention on the mma.sync kernel instead of the tcgen05 pair kernel
    T_COUNT
};
int64_t tune(Tune t);

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

// ---- programmatic dependent launch -------------------------------------------------------
// The towers are chains of ~70 short kernels on one stream.  Launched with the programmatic-serialization
// attribute, a kernel's CTAs may be scheduled while its predecessor is still draining: their prologue
// (barrier init, TMEM allocation, descriptor prefetch, LUTs) overlaps the predecessor's tail, and the
// launch latency disappears from the critical path.  Every such kernel calls pdl_wait() before it touches
// global memory (it returns once the predecessor grid has completed and its writes are visible) and
// pdl_launch_dependents() to let its own successor be scheduled.  Both are no-ops in a plain launch.
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

// launch with optional cluster dimension and (unless the pdl knob is 0) the programmatic-serialization attribute
template <typename... KArgs, typename... Args>
cudaError_t launch_ex(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t s, int cluster_x,
                      bool pdl, Args &&...args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = smem;
    cfg.stream = s;
    cudaLaunchAttribute attr[2];
    int na = 0;
    if (cluster_x > 1) {
        attr[na].id = cudaLaunchAttributeClusterDimension;
        attr[na].val.clusterDim.x = cluster_x;
        attr[na].val.clusterDim.y = 1;
        attr[na].val.clusterDim.z = 1;
        na++;
    }
    if (pdl && tune(T_PDL) != 0) {
        // the attribute is kept inside CUDA-graph captures too (programmatic edges: single-query encode_image
        // 0.486 -> 0.432 ms, encode_text 0.406 -> 0.334 ms, profiles/r02_latency_pdl_in_graph.txt); pdl = 2 drops
        // it there
        cudaStreamCaptureStatus cs = cudaStreamCaptureStatusNone;
        if (tune(T_PDL) == 2 &&
            cudaStreamIsCapturing(s, &cs) != cudaSuccess) { cudaGetLastError(); cs = cudaStreamCaptureStatusActive; }
        if (cs == cudaStreamCaptureStatusNone) {
            attr[na].id = cudaLaunchAttributeProgrammaticStreamSerialization;
            attr[na].val.programmaticStreamSerializationAllowed = 1;
            na++;
        }
    }
    cfg.attrs = attr;
    cfg.numAttrs = na;
    return cudaLaunchKernelEx(&cfg, kernel, std::forward<Args>(args)...);
}

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
