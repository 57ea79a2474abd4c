// Misc C-ABI entry points: error string, version, device count, launch counter.
#include "common.cuh"

#include <atomic>
#include <cstdarg>
#include <cstdlib>
#include <cstring>

namespace cb {

static thread_local char g_err[1024] = "";
thread_local int64_t g_launches = 0;

void set_error(const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

// ---- tuning knobs ---------------------------------------------------------------------
static const char *const kTuneNames[T_COUNT] = {
    "gemm_bn", "gemm_ncta", "gemm_stages", "gemm_raster", "batch_min_nq", "no_graph",
    "ln_blocks_per_sm", "ln_fold", "gemm_debug", "skip", "pdl", "gemm_skinny", "gemm_resid_stages",
    "attn_tc"};
static std::atomic<int64_t> g_tune[T_COUNT];

// environment -> table, once, at load time (CLIPB200_<NAME IN CAPITALS>)
static struct TuneInit {
    TuneInit() {
        for (int i = 0; i < T_COUNT; i++) {
            char env[64] = "CLIPB200_";
            size_t n = strlen(env);
            for (const char *c = kTuneNames[i]; *c && n + 1 < sizeof(env); c++)
                env[n++] = (char)((*c >= 'a' && *c <= 'z') ? *c - 32 : *c);
            env[n] = 0;
            const char *e = getenv(env);
            g_tune[i].store(e && *e ? atoll(e) : -1, std::memory_order_relaxed);
        }
        // historical spelling: CLIPB200_NO_LN_FOLD=1 == CLIPB200_LN_FOLD=0
        if (const char *e = getenv("CLIPB200_NO_LN_FOLD"))
            if (atoi(e)) g_tune[T_LN_FOLD].store(0, std::memory_order_relaxed);
    }
} g_tune_init;

int64_t tune(Tune t) { return g_tune[t].load(std::memory_order_relaxed); }

}  // namespace cb

extern "C" {

int cb_tuning_set(const char *name, int64_t value) {
    CB_REQUIRE(name != nullptr, "cb_tuning_set: null name");
    for (int i = 0; i < cb::T_COUNT; i++)
        if (!strcmp(name, cb::kTuneNames[i])) {
#ifndef CLIPB200_EXPERIMENTS
            CB_REQUIRE(i != cb::T_GEMM_DEBUG && i != cb::T_SKIP,
                       "cb_tuning_set: %s exists only in -DCLIPB200_EXPERIMENTS builds", name);
#endif
            cb::g_tune[i].store(value, std::memory_order_relaxed);
            return CB_OK;
        }
    cb::set_error("cb_tuning_set: unknown knob %s", name);
    return CB_ERR_INVALID;
}

int cb_tuning_get(const char *name, int64_t *value) {
    CB_REQUIRE(name && value, "cb_tuning_get: null argument");
    for (int i = 0; i < cb::T_COUNT; i++)
        if (!strcmp(name, cb::kTuneNames[i])) { *value = cb::tune((cb::Tune)i); return CB_OK; }
    cb::set_error("cb_tuning_get: unknown knob %s", name);
    return CB_ERR_INVALID;
}

const char *cb_last_error(void) { return cb::g_err; }

int cb_abi_version(void) { return 2; }

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
