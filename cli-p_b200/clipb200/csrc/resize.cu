// PIL-exact bicubic Resize(224, shorter side) + CenterCrop(224) on the GPU, uint8 RGB.
//
// Replaces the first two steps of clip._transform that the reference runs on the CPU through
// Pillow at /root/reference/build-index.py:48 (torchvision Resize(224, BICUBIC) -> PIL
// Image.resize -> ImagingResample; then CenterCrop).  Pillow's 8-bit resample is
// deterministic integer arithmetic: per output pixel a window of bicubic (a = -0.5) weights
// (support 2 x max(scale, 1)), normalised in double, rounded to 22-bit fixed point, summed in
// int32 from 1 << 21, shifted and clamped to uint8 -- horizontal pass first, then vertical on the
// uint8 intermediate.  The weights are computed on the host with the same double arithmetic
// and cached per (in size, out size, crop start); the two passes below reproduce Pillow
// bit for bit (tests/test_resize.py pins the CPU restatement against Pillow, the GPU test pins
// this kernel against both).  Only the 224 cropped columns / the rows the crop needs are computed.
#include "common.cuh"

#include <cmath>
#include <map>
#include <mutex>
#include <tuple>
#include <vector>

namespace cb {
namespace {

constexpr int PRECISION_BITS = 32 - 8 - 2;
constexpr int OUT = 224;

struct Table {
    int *bounds = nullptr;   // [OUT][2] (first input index, tap count) for the cropped outputs
    int *kk = nullptr;       // [OUT][ksize] fixed-point weights
    int ksize = 0;
    int first = 0, last = 0; // input index range [first, last) touched by the cropped outputs
};

double bicubic(double x) {
    const double a = -0.5;
    if (x < 0.0) x = -x;
    if (x < 1.0) return ((a + 2.0) * x - (a + 3.0)) * x * x + 1;
    if (x < 2.0) return (((x - 5) * x + 8) * x - 4) * a;
    return 0.0;
}

std::mutex g_mu;
std::map<std::tuple<int, int, int, int>, Table> g_tables;   // (device, in, out, start)

int get_table(int device, int in_size, int out_size, int start, Table *out) {
    std::lock_guard<std::mutex> lk(g_mu);
    auto key = std::make_tuple(device, in_size, out_size, start);
    auto it = g_tables.find(key);
    if (it != g_tables.end()) { *out = it->second; return CB_OK; }
    if (g_tables.size() >= 2048) {
        // bound the cache (a photo library has few distinct sizes; a hostile one does not): wait for
        // kernels that may still read the old tables, then drop them all
        CB_CUDA(cudaDeviceSynchronize());
        for (auto &kv : g_tables) { cudaFree(kv.second.bounds); cudaFree(kv.second.kk); }
        g_tables.clear();
    }
    const double scale = (double)in_size / out_size;
    double filterscale = scale < 1.0 ? 1.0 : scale;
    const double support = 2.0 * filterscale;
    const int ksize = (int)std::ceil(support) * 2 + 1;
    const double ss = 1.0 / filterscale;
    std::vector<int> bounds(OUT * 2), kk((size_t)OUT * ksize, 0);
    std::vector<double> w(ksize);
    int first = in_size, last = 0;
    for (int o = 0; o < OUT; o++) {
        const int xx = start + o;
        const double center = (xx + 0.5) * scale;
        int xmin = (int)(center - support + 0.5);
        if (xmin < 0) xmin = 0;
        int xmax = (int)(center + support + 0.5);
        if (xmax > in_size) xmax = in_size;
        xmax -= xmin;
        double ww = 0.0;
        for (int x = 0; x < xmax; x++) { w[x] = bicubic((x + xmin - center + 0.5) * ss); ww += w[x]; }
        for (int x = 0; x < xmax; x++) {
            const double v = ww != 0.0 ? w[x] / ww : w[x];
            kk[(size_t)o * ksize + x] = v < 0 ? (int)(-0.5 + v * (1 << PRECISION_BITS)) : (int)(0.5 + v * (1 << PRECISION_BITS));
        }
        bounds[o * 2] = xmin;
        bounds[o * 2 + 1] = xmax;
        first = std::min(first, xmin);
        last = std::max(last, xmin + xmax);
    }
    Table t;
    t.ksize = ksize; t.first = first; t.last = last;
    CB_CUDA(cudaMalloc(&t.bounds, bounds.size() * 4));
    CB_CUDA(cudaMalloc(&t.kk, kk.size() * 4));
    CB_CUDA(cudaMemcpy(t.bounds, bounds.data(), bounds.size() * 4, cudaMemcpyHostToDevice));
    CB_CUDA(cudaMemcpy(t.kk, kk.data(), kk.size() * 4, cudaMemcpyHostToDevice));
    g_tables[key] = t;
    *out = t;
    return CB_OK;
}

__device__ __forceinline__ uint8_t clip8(int v) {
    v >>= PRECISION_BITS;
    return (uint8_t)(v < 0 ? 0 : (v > 255 ? 255 : v));
}

// horizontal pass: src rows [row0, row0 + nrows) x w -> dst [nrows][224][3]
__global__ void resize_h_kernel(const uint8_t *__restrict__ src, int w, int row0, int nrows,
                                const int *__restrict__ bounds, const int *__restrict__ kk, int ksize,
                                uint8_t *__restrict__ dst) {
    const int o = blockIdx.x * blockDim.x + threadIdx.x;     // output column within the crop
    const int r = blockIdx.y;
    if (o >= OUT || r >= nrows) return;
    const int lo = bounds[o * 2], n = bounds[o * 2 + 1];
    const uint8_t *p = src + ((size_t)(row0 + r) * w + lo) * 3;
    const int *k = kk + (size_t)o * ksize;
    int s0 = 1 << (PRECISION_BITS - 1), s1 = s0, s2 = s0;
    for (int x = 0; x < n; x++) {
        const int c = k[x];
        s0 += p[x * 3] * c; s1 += p[x * 3 + 1] * c; s2 += p[x * 3 + 2] * c;
    }
    uint8_t *d = dst + ((size_t)r * OUT + o) * 3;
    d[0] = clip8(s0); d[1] = clip8(s1); d[2] = clip8(s2);
}

// vertical pass: src [.., row stride, col offset] -> dst [224][224][3]
__global__ void resize_v_kernel(const uint8_t *__restrict__ src, int src_row_pixels, int src_col0, int src_row0,
                                const int *__restrict__ bounds, const int *__restrict__ kk, int ksize,
                                uint8_t *__restrict__ dst) {
    const int x = blockIdx.x * blockDim.x + threadIdx.x;
    const int o = blockIdx.y;                                 // output row within the crop
    if (x >= OUT) return;
    const int lo = bounds[o * 2], n = bounds[o * 2 + 1];
    const int *k = kk + (size_t)o * ksize;
    const uint8_t *p = src + ((size_t)(lo - src_row0) * src_row_pixels + src_col0 + x) * 3;
    int s0 = 1 << (PRECISION_BITS - 1), s1 = s0, s2 = s0;
    for (int y = 0; y < n; y++) {
        const int c = k[y];
        const uint8_t *q = p + (size_t)y * src_row_pixels * 3;
        s0 += q[0] * c; s1 += q[1] * c; s2 += q[2] * c;
    }
    uint8_t *d = dst + ((size_t)o * OUT + x) * 3;
    d[0] = clip8(s0); d[1] = clip8(s1); d[2] = clip8(s2);
}

// plain crop copy (one of the dimensions already has the target size)
__global__ void crop_kernel(const uint8_t *__restrict__ src, int src_row_pixels, int col0, int row0,
                            uint8_t *__restrict__ dst) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= OUT * OUT * 3) return;
    const int r = i / (OUT * 3), c = i % (OUT * 3);
    dst[i] = src[((size_t)(row0 + r) * src_row_pixels + col0) * 3 + c];
}

}  // namespace
}  // namespace cb

using namespace cb;

extern "C" int cb_resize224_u8_device(const uint8_t *src_hwc, int h, int w, uint8_t *dst, void *stream) {
    CB_REQUIRE(src_hwc && dst, "cb_resize224_u8_device: null buffer");
    CB_REQUIRE(h > 0 && w > 0 && h <= 65535 && w <= 65535, "cb_resize224_u8_device: bad image size %dx%d", w, h);
    cudaStream_t s = (cudaStream_t)stream;
    int device = 0;
    CB_CUDA(cudaGetDevice(&device));
    // torchvision Resize(int): shorter side -> 224, longer side truncated; CenterCrop rounds half the margin
    int nw, nh;
    if (w <= h) { nw = OUT; nh = (int)((double)OUT * h / w); }
    else { nw = (int)((double)OUT * w / h); nh = OUT; }
    const int left = (int)std::nearbyint((nw - OUT) / 2.0), top = (int)std::nearbyint((nh - OUT) / 2.0);
    const bool need_h = nw != w, need_v = nh != h;
    if (!need_h && !need_v) {
        crop_kernel<<<(OUT * OUT * 3 + 255) / 256, 256, 0, s>>>(src_hwc, w, left, top, dst);
        CB_LAUNCH_CHECK();
        return CB_OK;
    }
    Table th, tv;
    int rc;
    if (need_h && (rc = get_table(device, w, nw, left, &th))) return rc;
    if (need_v && (rc = get_table(device, h, nh, top, &tv))) return rc;
    if (need_h && !need_v) {
        // rows [top, top+224) straight to the output
        resize_h_kernel<<<dim3((OUT + 63) / 64, OUT), 64, 0, s>>>(src_hwc, w, top, OUT, th.bounds, th.kk, th.ksize, dst);
        CB_LAUNCH_CHECK();
        return CB_OK;
    }
    if (!need_h && need_v) {
        resize_v_kernel<<<dim3((OUT + 63) / 64, OUT), 64, 0, s>>>(src_hwc, w, left, 0, tv.bounds, tv.kk, tv.ksize, dst);
        CB_LAUNCH_CHECK();
        return CB_OK;
    }
    // both: horizontal pass on exactly the input rows the cropped vertical pass will read
    const int row0 = tv.first, nrows = tv.last - tv.first;
    // intermediate rows: a grow-only buffer per (device, stream) - calls on one stream are ordered, so the
    // buffer is free again by the time the next call's first kernel runs.  (The stream-ordered allocator
    // cost ~0.4 ms per call once 16 decoder threads hit it at the same time.)
    const size_t need = (size_t)nrows * OUT * 3;
    uint8_t *tmp = nullptr;
    {
        static std::mutex mu;
        static std::map<std::pair<int, cudaStream_t>, std::pair<uint8_t *, size_t>> scratch;
        std::lock_guard<std::mutex> lk(mu);
        if (scratch.size() >= 256 && !scratch.count({device, s})) {     // streams come and go: start over
            CB_CUDA(cudaDeviceSynchronize());
            for (auto &kv : scratch) cudaFree(kv.second.first);
            scratch.clear();
        }
        auto &slot = scratch[{device, s}];
        if (slot.second < need) {
            if (slot.first) { CB_CUDA(cudaStreamSynchronize(s)); cudaFree(slot.first); slot = {nullptr, 0}; }
            CB_CUDA(cudaMalloc(&slot.first, need));
            slot.second = need;
        }
        tmp = slot.first;
    }

    resize_h_kernel<<<dim3((OUT + 63) / 64, nrows), 64, 0, s>>>(src_hwc, w, row0, nrows, th.bounds, th.kk, th.ksize, tmp);
    CB_LAUNCH_CHECK();
    resize_v_kernel<<<dim3((OUT + 63) / 64, OUT), 64, 0, s>>>(tmp, OUT, 0, row0, tv.bounds, tv.kk, tv.ksize, dst);
    CB_LAUNCH_CHECK();
    return CB_OK;
}
