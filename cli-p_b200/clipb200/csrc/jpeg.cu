// JPEG files -> uint8 [n,224,224,3] pixels in HBM, ready for cb_clip_submit_image_u8_device.
//
// The index-time caller of hot path A (reference: Image.open + transform at
// /root/reference/build-index.py:47-48, one file at a time on one core).  Decoding is LIBRARY work
// (nvjpeg, like cuBLAS for a plain GEMM) - what this file adds is the plumbing that keeps it off the
// critical path: N host threads, each with its own nvjpeg decoder state, pinned/device buffers and
// CUDA stream, pull files from a shared counter and run nvjpeg's decoupled phases
//     read file -> parse -> Huffman decode (host) -> transfer -> IDCT + colour (device, RGB interleaved)
// straight into the caller's batch slot when the image is 224 x 224, or into a per-thread scratch
// image followed by the Pillow-exact resize kernel (resize.cu) otherwise.  No Python in the loop, no
// GIL, one C call per batch.  nvjpeg is loaded with dlopen so libclipb200.so itself never depends on it.
#include "common.cuh"

#include <nvjpeg.h>

#include <dlfcn.h>

#include <atomic>
#include <cstdio>
#include <mutex>
#include <thread>
#include <vector>

extern "C" int cb_resize224_u8_device(const uint8_t *src_hwc, int h, int w, uint8_t *dst, void *stream);

namespace cb {
namespace {

#define NVJ_FUNCS(X)                                                                                      \
    X(nvjpegCreateSimple) X(nvjpegDestroy) X(nvjpegDecoderCreate) X(nvjpegDecoderDestroy)                 \
    X(nvjpegDecoderStateCreate) X(nvjpegJpegStateDestroy) X(nvjpegBufferPinnedCreate)                     \
    X(nvjpegBufferPinnedDestroy) X(nvjpegBufferDeviceCreate) X(nvjpegBufferDeviceDestroy)                 \
    X(nvjpegStateAttachPinnedBuffer) X(nvjpegStateAttachDeviceBuffer) X(nvjpegJpegStreamCreate)           \
    X(nvjpegJpegStreamDestroy) X(nvjpegJpegStreamParse) X(nvjpegJpegStreamGetFrameDimensions)             \
    X(nvjpegDecodeParamsCreate) X(nvjpegDecodeParamsDestroy) X(nvjpegDecodeParamsSetOutputFormat)         \
    X(nvjpegDecodeJpegHost) X(nvjpegDecodeJpegTransferToDevice) X(nvjpegDecodeJpegDevice)

struct NvjApi {
    void *so = nullptr;
#define X(name) decltype(&::name) name = nullptr;
    NVJ_FUNCS(X)
#undef X
};

int load_api(NvjApi **out) {
    static NvjApi api;
    static std::mutex mu;
    static int state = 0;      // 0 not tried, 1 ok, -1 failed
    std::lock_guard<std::mutex> lk(mu);
    if (state == 0) {
        const char *cands[] = {getenv("CLIPB200_NVJPEG_LIB"), "libnvjpeg.so.12", "libnvjpeg.so",
                               "/usr/local/cuda/lib64/libnvjpeg.so.12"};
        for (const char *c : cands) {
            if (!c || !*c) continue;
            api.so = dlopen(c, RTLD_NOW | RTLD_LOCAL);
            if (api.so) break;
        }
        state = api.so ? 1 : -1;
        if (api.so) {
#define X(name)                                                            \
    api.name = reinterpret_cast<decltype(&::name)>(dlsym(api.so, #name)); \
    if (!api.name) state = -1;
            NVJ_FUNCS(X)
#undef X
        }
    }
    if (state != 1) {
        set_error("nvjpeg is not available (dlopen libnvjpeg.so.12 failed; set CLIPB200_NVJPEG_LIB)");
        return CB_ERR_INVALID;
    }
    *out = &api;
    return CB_OK;
}

constexpr int OUT = 224;
constexpr size_t OUT_BYTES = (size_t)OUT * OUT * 3;

// Two nvjpeg back ends per worker, picked per image by its pixel count (profiles/r01_jpeg_sizes.txt):
// [0] NVJPEG_BACKEND_HYBRID      Huffman decoding on the worker's host core: 55 k images/s at 224 px on 16
//                                threads, but 377/s at 12 Mpixel
// [1] NVJPEG_BACKEND_GPU_HYBRID  Huffman decoding on the GPU: 940 images/s (11 Gpixel/s) at 12 Mpixel, but
//                                4x slower than [0] on thumbnails
constexpr int64_t kGpuHuffmanMinPixels = 4 << 20;

struct Worker {
    nvjpegJpegDecoder_t decoder[2] = {nullptr, nullptr};
    nvjpegJpegState_t state[2] = {nullptr, nullptr};
    nvjpegBufferPinned_t pinned = nullptr;
    nvjpegBufferDevice_t devbuf = nullptr;
    nvjpegJpegStream_t jstream = nullptr;
    nvjpegDecodeParams_t params = nullptr;
    cudaStream_t stream = nullptr;
    uint8_t *scratch = nullptr;        // decoded image when it is not 224 x 224
    size_t scratch_bytes = 0;
    std::vector<uint8_t> file;
};

}  // namespace
}  // namespace cb

struct cb_jpeg {
    int device = 0;
    cb::NvjApi *api = nullptr;
    nvjpegHandle_t handle = nullptr;
    std::vector<cb::Worker> workers;
    int force_backend = -1;            // CLIPB200_NVJPEG_BACKEND=1|2 pins one back end (experiments)
};

using namespace cb;

namespace {

// status codes written per image: 0 ok, 1 file unreadable, 2 not a decodable JPEG, 3 CUDA error,
// 4 unsupported by this path (e.g. CMYK) - the caller falls back to its CPU decoder for those
int decode_one(cb_jpeg *j, Worker &w, const uint8_t *data, size_t len, uint8_t *dst) {
    NvjApi &a = *j->api;
    if (a.nvjpegJpegStreamParse(j->handle, data, len, 0, 0, w.jstream) != NVJPEG_STATUS_SUCCESS) return 2;
    unsigned int W = 0, H = 0;
    if (a.nvjpegJpegStreamGetFrameDimensions(w.jstream, &W, &H) != NVJPEG_STATUS_SUCCESS || W == 0 || H == 0) return 2;
    if (W > 65535 || H > 65535) return 4;
    const bool direct = W == OUT && H == OUT;
    uint8_t *target = dst;
    if (!direct) {
        const size_t need = (size_t)W * H * 3;
        if (need > w.scratch_bytes) {
            if (w.scratch) cudaFree(w.scratch);
            w.scratch = nullptr;
            w.scratch_bytes = 0;
            if (cudaMalloc(&w.scratch, need) != cudaSuccess) { cudaGetLastError(); return 3; }
            w.scratch_bytes = need;
        }
        target = w.scratch;
    }
    nvjpegImage_t img = {};
    img.channel[0] = target;
    img.pitch[0] = (size_t)W * 3;
    int be = (int64_t)W * H >= kGpuHuffmanMinPixels ? 1 : 0;
    if (j->force_backend >= 0) be = j->force_backend;
    if (!w.decoder[be]) be ^= 1;
    nvjpegStatus_t st = a.nvjpegDecodeJpegHost(j->handle, w.decoder[be], w.state[be], w.params, w.jstream);
    if (st == NVJPEG_STATUS_SUCCESS) st = a.nvjpegDecodeJpegTransferToDevice(j->handle, w.decoder[be], w.state[be], w.jstream, w.stream);
    if (st == NVJPEG_STATUS_SUCCESS) st = a.nvjpegDecodeJpegDevice(j->handle, w.decoder[be], w.state[be], &img, w.stream);
    if (st != NVJPEG_STATUS_SUCCESS && be == 1 && w.decoder[0]) {
        // the GPU Huffman decoder takes baseline files only: progressive and the like go to the host decoder
        cudaStreamSynchronize(w.stream);
        st = a.nvjpegDecodeJpegHost(j->handle, w.decoder[0], w.state[0], w.params, w.jstream);
        if (st == NVJPEG_STATUS_SUCCESS) st = a.nvjpegDecodeJpegTransferToDevice(j->handle, w.decoder[0], w.state[0], w.jstream, w.stream);
        if (st == NVJPEG_STATUS_SUCCESS) st = a.nvjpegDecodeJpegDevice(j->handle, w.decoder[0], w.state[0], &img, w.stream);
    }
    if (st != NVJPEG_STATUS_SUCCESS) {
        cudaStreamSynchronize(w.stream);
        return st == NVJPEG_STATUS_JPEG_NOT_SUPPORTED ? 4 : 2;
    }
    int rc = 0;
    if (!direct && cb_resize224_u8_device(target, (int)H, (int)W, dst, w.stream) != CB_OK) rc = 3;
    // the pinned/device buffers of this worker and its scratch image are reused by its next file
    if (cudaStreamSynchronize(w.stream) != cudaSuccess) { cudaGetLastError(); rc = 3; }
    return rc;
}

bool read_file(const char *path, std::vector<uint8_t> &buf) {
    FILE *f = fopen(path, "rb");
    if (!f) return false;
    bool ok = false;
    if (fseek(f, 0, SEEK_END) == 0) {
        const long n = ftell(f);
        if (n > 0 && fseek(f, 0, SEEK_SET) == 0) {
            buf.resize((size_t)n);
            ok = fread(buf.data(), 1, (size_t)n, f) == (size_t)n;
        }
    }
    fclose(f);
    return ok;
}

void free_worker(cb_jpeg *j, Worker &w) {
    NvjApi &a = *j->api;
    if (w.stream) cudaStreamSynchronize(w.stream);
    if (w.params) a.nvjpegDecodeParamsDestroy(w.params);
    if (w.jstream) a.nvjpegJpegStreamDestroy(w.jstream);
    for (int b = 0; b < 2; b++)
        if (w.state[b]) a.nvjpegJpegStateDestroy(w.state[b]);
    if (w.devbuf) a.nvjpegBufferDeviceDestroy(w.devbuf);
    if (w.pinned) a.nvjpegBufferPinnedDestroy(w.pinned);
    for (int b = 0; b < 2; b++)
        if (w.decoder[b]) a.nvjpegDecoderDestroy(w.decoder[b]);
    if (w.scratch) cudaFree(w.scratch);
    if (w.stream) cudaStreamDestroy(w.stream);
    w = Worker();
}

template <typename GetData>
int decode_batch(cb_jpeg *j, int64_t n, uint8_t *out_dev, int32_t *status, GetData &&get) {
    DeviceGuard guard(j->device);          // worker 0 is the calling thread: leave its current device as it was
    std::atomic<int64_t> next(0);
    const int nt = (int)std::min<int64_t>((int64_t)j->workers.size(), n);
    auto body = [&](int t) {
        cudaSetDevice(j->device);
        Worker &w = j->workers[t];
        for (;;) {
            const int64_t i = next.fetch_add(1);
            if (i >= n) break;
            const uint8_t *data = nullptr;
            size_t len = 0;
            if (!get(i, w, &data, &len)) { status[i] = 1; continue; }
            status[i] = decode_one(j, w, data, len, out_dev + (size_t)i * OUT_BYTES);
        }
    };
    std::vector<std::thread> th;
    for (int t = 1; t < nt; t++) th.emplace_back(body, t);
    body(0);
    for (auto &x : th) x.join();
    return CB_OK;
}

}  // namespace

extern "C" {

int cb_jpeg_create(int device, int threads, cb_jpeg **out) {
    CB_REQUIRE(out != nullptr, "cb_jpeg_create: out is null");
    *out = nullptr;
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
        cudaGetLastError();
        set_error("cb_jpeg_create: no CUDA device (this library has no CPU fallback)");
        return CB_ERR_NOGPU;
    }
    CB_REQUIRE(device >= 0 && device < ndev, "cb_jpeg_create: device %d out of range", device);
    NvjApi *api = nullptr;
    int rc = load_api(&api);
    if (rc) return rc;
    DeviceGuard g(device);
    if (threads <= 0) threads = (int)std::min(16u, std::max(1u, std::thread::hardware_concurrency()));
    threads = std::min(threads, 64);
    cb_jpeg *j = new (std::nothrow) cb_jpeg();
    if (!j) { set_error("out of host memory"); return CB_ERR_OOM; }
    j->device = device;
    j->api = api;
    bool ok = api->nvjpegCreateSimple(&j->handle) == NVJPEG_STATUS_SUCCESS;
    if (const char *e = getenv("CLIPB200_NVJPEG_BACKEND")) {
        const int v = atoi(e);
        if (v == 1 || v == 2) j->force_backend = v - 1;
    }
    j->workers.resize(ok ? threads : 0);
    for (Worker &w : j->workers) {
        ok = ok && api->nvjpegBufferPinnedCreate(j->handle, nullptr, &w.pinned) == NVJPEG_STATUS_SUCCESS;
        ok = ok && api->nvjpegBufferDeviceCreate(j->handle, nullptr, &w.devbuf) == NVJPEG_STATUS_SUCCESS;
        const nvjpegBackend_t kinds[2] = {NVJPEG_BACKEND_HYBRID, NVJPEG_BACKEND_GPU_HYBRID};
        for (int b = 0; b < 2 && ok; b++) {
            // the GPU Huffman back end is optional (the host one decodes everything); both share the buffers
            bool got = api->nvjpegDecoderCreate(j->handle, kinds[b], &w.decoder[b]) == NVJPEG_STATUS_SUCCESS &&
                       api->nvjpegDecoderStateCreate(j->handle, w.decoder[b], &w.state[b]) == NVJPEG_STATUS_SUCCESS &&
                       api->nvjpegStateAttachPinnedBuffer(w.state[b], w.pinned) == NVJPEG_STATUS_SUCCESS &&
                       api->nvjpegStateAttachDeviceBuffer(w.state[b], w.devbuf) == NVJPEG_STATUS_SUCCESS;
            if (!got && b == 0) ok = false;
            if (!got && b == 1) {
                if (w.state[1]) api->nvjpegJpegStateDestroy(w.state[1]);
                if (w.decoder[1]) api->nvjpegDecoderDestroy(w.decoder[1]);
                w.state[1] = nullptr;
                w.decoder[1] = nullptr;
            }
        }
        ok = ok && api->nvjpegJpegStreamCreate(j->handle, &w.jstream) == NVJPEG_STATUS_SUCCESS;
        ok = ok && api->nvjpegDecodeParamsCreate(j->handle, &w.params) == NVJPEG_STATUS_SUCCESS;
        ok = ok && api->nvjpegDecodeParamsSetOutputFormat(w.params, NVJPEG_OUTPUT_RGBI) == NVJPEG_STATUS_SUCCESS;
        ok = ok && cudaStreamCreateWithFlags(&w.stream, cudaStreamNonBlocking) == cudaSuccess;
        if (!ok) break;
    }
    if (!ok) {
        cudaGetLastError();
        set_error("cb_jpeg_create: nvjpeg initialisation failed");
        cb_jpeg_free(j);
        return CB_ERR_CUDA;
    }
    *out = j;
    return CB_OK;
}

void cb_jpeg_free(cb_jpeg *j) {
    if (!j) return;
    DeviceGuard g(j->device);
    for (Worker &w : j->workers) free_worker(j, w);
    if (j->handle) j->api->nvjpegDestroy(j->handle);
    delete j;
}

int cb_jpeg_threads(const cb_jpeg *j) { return j ? (int)j->workers.size() : 0; }

int cb_jpeg_decode_files(cb_jpeg *j, int64_t n, const char *const *paths, uint8_t *out_dev, int32_t *status) {
    CB_REQUIRE(j != nullptr, "cb_jpeg_decode_files: null handle");
    CB_REQUIRE(n >= 0, "cb_jpeg_decode_files: n < 0");
    if (n == 0) return CB_OK;
    CB_REQUIRE(paths && out_dev && status, "cb_jpeg_decode_files: null buffer");
    return decode_batch(j, n, out_dev, status, [&](int64_t i, Worker &w, const uint8_t **data, size_t *len) {
        if (!paths[i] || !read_file(paths[i], w.file)) return false;
        *data = w.file.data();
        *len = w.file.size();
        return true;
    });
}

int cb_jpeg_decode_memory(cb_jpeg *j, int64_t n, const uint8_t *const *data, const int64_t *sizes, uint8_t *out_dev,
                          int32_t *status) {
    CB_REQUIRE(j != nullptr, "cb_jpeg_decode_memory: null handle");
    CB_REQUIRE(n >= 0, "cb_jpeg_decode_memory: n < 0");
    if (n == 0) return CB_OK;
    CB_REQUIRE(data && sizes && out_dev && status, "cb_jpeg_decode_memory: null buffer");
    return decode_batch(j, n, out_dev, status, [&](int64_t i, Worker &, const uint8_t **d, size_t *len) {
        if (!data[i] || sizes[i] <= 0) return false;
        *d = data[i];
        *len = (size_t)sizes[i];
        return true;
    });
}

}  // extern "C"
