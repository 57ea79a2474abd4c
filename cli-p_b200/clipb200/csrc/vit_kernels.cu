// HBM/L2-bound helper kernels of the CLIP towers: preprocess (clip._transform's
// ToTensor+Normalize, written straight into patch-major im2col layout), LayerNorm
// (fp32 statistics, as openai/CLIP's LayerNorm subclass does), row L2-normalise
// (/root/reference/build-index.py:50), token-embedding gather (CLIP.encode_text).
// All vectorised to 128-bit accesses, one warp per row, warp-shuffle reductions.
#include "common.cuh"
#include "vit_kernels.cuh"

namespace cb {
namespace {

#define K_MEAN0 0.48145466f
#define K_MEAN1 0.4578275f
#define K_MEAN2 0.40821073f
#define K_STD0 0.26862954f
#define K_STD1 0.26130258f
#define K_STD2 0.27577711f

__device__ __forceinline__ uint32_t pack2(float a, float b) {
    __half2 h = __floats2half2_rn(a, b);
    return *reinterpret_cast<uint32_t *>(&h);
}

// one thread = 8 consecutive bytes of one image row (672 B = 84 chunks); 12 consecutive
// threads cover one 32-pixel patch row = 96 contiguous output halfs.  A pixel byte has only
// 256 x 3 possible results, so each block first builds the table fp16((v/255 - mean_c)/std_c)
// with exact IEEE divisions (bit-identical to torch's ToTensor + Normalize) and the streaming
// loop is pure load / look-up / store.
__global__ void __launch_bounds__(256) preprocess_u8_kernel(const uint8_t *__restrict__ img,
                                                           __half *__restrict__ patches, int B) {
    __shared__ __half lut[3][256];
    for (int i = threadIdx.x; i < 768; i += blockDim.x) {
        const int c = i >> 8, v = i & 255;
        const float mean = c == 0 ? K_MEAN0 : (c == 1 ? K_MEAN1 : K_MEAN2);
        const float sd = c == 0 ? K_STD0 : (c == 1 ? K_STD1 : K_STD2);
        lut[c][v] = __float2half_rn(__fdiv_rn(__fdiv_rn((float)v, 255.0f) - mean, sd));
    }
    __syncthreads();
    pdl_wait();                     // the LUT above needed no global memory (common.cuh)
    const int64_t total = (int64_t)B * 224 * 84;
    for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < total;
         t += (int64_t)gridDim.x * blockDim.x) {
        const int xb = (int)(t % 84);
        const int y = (int)((t / 84) % 224);
        const int b = (int)(t / (84 * 224));
        const uint2 raw = __ldg(reinterpret_cast<const uint2 *>(img + ((size_t)(b * 224 + y) * 224) * 3) + xb);
        const uint8_t *px = reinterpret_cast<const uint8_t *>(&raw);
        __half h[8];
        int c = (xb * 8) % 3;
#pragma unroll
        for (int e = 0; e < 8; e++) {
            h[e] = lut[c][px[e]];
            c = c == 2 ? 0 : c + 1;
        }
        const int gy = y >> 5, py = y & 31, gx = xb / 12, j = xb - gx * 12;
        __half *dst = patches + ((size_t)(b * 49 + gy * 7 + gx)) * 3072 + py * 96 + j * 8;
        *reinterpret_cast<uint4 *>(dst) = *reinterpret_cast<const uint4 *>(h);
    }
    pdl_launch_dependents();
}

// one thread = 4 pixels x 3 channels of an NCHW fp32 image
__global__ void __launch_bounds__(256) preprocess_f32_kernel(const float *__restrict__ img,
                                                            __half *__restrict__ patches, int B) {
    pdl_wait();
    const int64_t total = (int64_t)B * 224 * 56;
    for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < total;
         t += (int64_t)gridDim.x * blockDim.x) {
        const int x4 = (int)(t % 56);
        const int y = (int)((t / 56) % 224);
        const int b = (int)(t / (56 * 224));
        float4 ch[3];
#pragma unroll
        for (int c = 0; c < 3; c++)
            ch[c] = __ldg(reinterpret_cast<const float4 *>(img + (((size_t)b * 3 + c) * 224 + y) * 224) + x4);
        const int gy = y >> 5, py = y & 31, gx = (x4 * 4) >> 5, px0 = (x4 * 4) & 31;
        __half *dst = patches + ((size_t)(b * 49 + gy * 7 + gx)) * 3072 + py * 96 + px0 * 3;
        uint2 *d2 = reinterpret_cast<uint2 *>(dst);
        d2[0] = make_uint2(pack2(ch[0].x, ch[1].x), pack2(ch[2].x, ch[0].y));
        d2[1] = make_uint2(pack2(ch[1].y, ch[2].y), pack2(ch[0].z, ch[1].z));
        d2[2] = make_uint2(pack2(ch[2].z, ch[0].w), pack2(ch[1].w, ch[2].w));
    }
}

// persistent warps: each warp walks rows with a stride and keeps the next row's loads in
// flight while it reduces / writes the current one
template <int W>
__global__ void __launch_bounds__(256) layernorm_kernel(const __half *__restrict__ in, __half *__restrict__ out,
                                                       const float *__restrict__ gamma, const float *__restrict__ beta,
                                                       int rows, int in_row_stride, const int *__restrict__ gather,
                                                       const float *__restrict__ cls_fill, int cls_period,
                                                       float *__restrict__ stats_out) {
    constexpr int NV = W / 256;     // uint4 (8 halfs) per lane
    const int lane = threadIdx.x & 31;
    const int stride = gridDim.x * (blockDim.x >> 5);
    int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    pdl_wait();                     // programmatic dependent launch (common.cuh): inputs come from the previous kernel
    pdl_launch_dependents();
    if (row >= rows) return;

    auto src_row = [&](int r) -> int64_t { return gather ? (int64_t)gather[r] : (int64_t)r * in_row_stride; };
    auto load = [&](uint4 (&u)[NV], int64_t in_row) {
        const uint4 *p = reinterpret_cast<const uint4 *>(in + in_row * W);
#pragma unroll
        for (int c = 0; c < NV; c++) u[c] = p[lane + 32 * c];
    };
    uint4 cur[NV], nxt[NV];
    int64_t cur_in = src_row(row);
    load(cur, cur_in);
    for (; row < rows; row += stride) {
        const int nrow = row + stride;
        int64_t nxt_in = 0;
        if (nrow < rows) { nxt_in = src_row(nrow); load(nxt, nxt_in); }
        float v[NV * 8];
        if (cls_fill != nullptr && (cur_in % cls_period) == 0) {
#pragma unroll
            for (int c = 0; c < NV; c++) {
                const float4 *p = reinterpret_cast<const float4 *>(cls_fill + (lane + 32 * c) * 8);
                float4 a = __ldg(p), b = __ldg(p + 1);
                v[c * 8 + 0] = a.x; v[c * 8 + 1] = a.y; v[c * 8 + 2] = a.z; v[c * 8 + 3] = a.w;
                v[c * 8 + 4] = b.x; v[c * 8 + 5] = b.y; v[c * 8 + 6] = b.z; v[c * 8 + 7] = b.w;
            }
        } else {
#pragma unroll
            for (int c = 0; c < NV; c++) {
                const __half2 *h = reinterpret_cast<const __half2 *>(&cur[c]);
#pragma unroll
                for (int e = 0; e < 4; e++) {
                    float2 f = __half22float2(h[e]);
                    v[c * 8 + 2 * e] = f.x;
                    v[c * 8 + 2 * e + 1] = f.y;
                }
            }
        }
        float s = 0.f;
#pragma unroll
        for (int i = 0; i < NV * 8; i++) s += v[i];
        const float mean = warp_sum(s) * (1.0f / W);
        float q = 0.f;
#pragma unroll
        for (int i = 0; i < NV * 8; i++) { float d = v[i] - mean; q = fmaf(d, d, q); }
        const float rstd = rsqrtf(warp_sum(q) * (1.0f / W) + 1e-5f);
        uint4 *o = reinterpret_cast<uint4 *>(out + (size_t)row * W);
        float os = 0.f, oq = 0.f;
#pragma unroll
        for (int c = 0; c < NV; c++) {
            const int col = (lane + 32 * c) * 8;     // gamma/beta: 6 KB, L1-resident after the first row
            const float4 g0 = __ldg(reinterpret_cast<const float4 *>(gamma + col)), g1 = __ldg(reinterpret_cast<const float4 *>(gamma + col) + 1);
            const float4 b0 = __ldg(reinterpret_cast<const float4 *>(beta + col)), b1 = __ldg(reinterpret_cast<const float4 *>(beta + col) + 1);
            const float g[8] = {g0.x, g0.y, g0.z, g0.w, g1.x, g1.y, g1.z, g1.w};
            const float bt[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
            float r[8];
#pragma unroll
            for (int e = 0; e < 8; e++) r[e] = fmaf((v[c * 8 + e] - mean) * rstd, g[e], bt[e]);
            o[lane + 32 * c] = make_uint4(pack2(r[0], r[1]), pack2(r[2], r[3]), pack2(r[4], r[5]), pack2(r[6], r[7]));
            if (stats_out) {
#pragma unroll
                for (int e = 0; e < 8; e++) {
                    const float h = __half2float(__float2half_rn(r[e]));
                    os += h;
                    oq = fmaf(h, h, oq);
                }
            }
        }
        if (stats_out) {
            os = warp_sum(os);
            oq = warp_sum(oq);
            if (lane == 0) reinterpret_cast<float2 *>(stats_out)[row] = make_float2(os, oq);
        }
#pragma unroll
        for (int c = 0; c < NV; c++) cur[c] = nxt[c];
        cur_in = nxt_in;
    }
}

__global__ void __launch_bounds__(256) l2norm_kernel(const float *__restrict__ in, float *__restrict__ out, int rows,
                                                    int width) {
    const int lane = threadIdx.x & 31;
    const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    pdl_wait();
    pdl_launch_dependents();
    if (row >= rows) return;
    const float4 *p = reinterpret_cast<const float4 *>(in + (size_t)row * width);
    float4 *o = reinterpret_cast<float4 *>(out + (size_t)row * width);
    const int n4 = width / 4;
    float s = 0.f;
    for (int i = lane; i < n4; i += 32) {
        float4 f = p[i];
        s = fmaf(f.x, f.x, s); s = fmaf(f.y, f.y, s); s = fmaf(f.z, f.z, s); s = fmaf(f.w, f.w, s);
    }
    const float inv = 1.0f / sqrtf(warp_sum(s));
    for (int i = lane; i < n4; i += 32) {
        float4 f = p[i];
        o[i] = make_float4(f.x * inv, f.y * inv, f.z * inv, f.w * inv);
    }
}

// one block per text row (77 tokens), one warp per token in turn
__global__ void __launch_bounds__(256) text_embed_kernel(const int32_t *__restrict__ ids, const float *__restrict__ tok,
                                                        const float *__restrict__ pos, __half *__restrict__ x,
                                                        int *__restrict__ eot_row, int ctx, int width, int vocab,
                                                        float *__restrict__ stats_out) {
    const int b = blockIdx.x;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = blockDim.x >> 5;
    pdl_wait();
    pdl_launch_dependents();
    const int32_t *row = ids + (size_t)b * ctx;
    if (warp == 0) {
        int best = INT_MIN, best_t = 0;
        for (int t = lane; t < ctx; t += 32) {
            int v = row[t];
            if (v > best) { best = v; best_t = t; }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            int ov = __shfl_xor_sync(0xffffffffu, best, o), ot = __shfl_xor_sync(0xffffffffu, best_t, o);
            if (ov > best || (ov == best && ot < best_t)) { best = ov; best_t = ot; }
        }
        if (lane == 0) eot_row[b] = b * ctx + best_t;
    }
    const int n4 = width / 4;
    for (int t = warp; t < ctx; t += nwarp) {
        int id = row[t];
        id = id < 0 ? 0 : (id >= vocab ? vocab - 1 : id);
        const float4 *e = reinterpret_cast<const float4 *>(tok + (size_t)id * width);
        const float4 *p = reinterpret_cast<const float4 *>(pos + (size_t)t * width);
        uint2 *o = reinterpret_cast<uint2 *>(x + ((size_t)b * ctx + t) * width);
        float os = 0.f, oq = 0.f;
        for (int i = lane; i < n4; i += 32) {
            float4 a = __ldg(e + i), c = __ldg(p + i);
            const float v[4] = {a.x + c.x, a.y + c.y, a.z + c.z, a.w + c.w};
            o[i] = make_uint2(pack2(v[0], v[1]), pack2(v[2], v[3]));
            if (stats_out) {
#pragma unroll
                for (int k = 0; k < 4; k++) {
                    const float h = __half2float(__float2half_rn(v[k]));
                    os += h;
                    oq = fmaf(h, h, oq);
                }
            }
        }
        if (stats_out) {
            os = warp_sum(os);
            oq = warp_sum(oq);
            if (lane == 0) reinterpret_cast<float2 *>(stats_out)[(size_t)b * ctx + t] = make_float2(os, oq);
        }
    }
}

inline int grid_for(int64_t threads, int block) {
    return (int)std::min<int64_t>((threads + block - 1) / block, (int64_t)kNumSMs * 16);
}

}  // namespace

int preprocess_u8(const uint8_t *img, __half *patches, int B, cudaStream_t s) {
    CB_CUDA(launch_ex(preprocess_u8_kernel, dim3(grid_for((int64_t)B * 224 * 84, 256)), dim3(256), 0, s, 1, true, img, patches, B));
    CB_LAUNCH_CHECK();
    return CB_OK;
}
int preprocess_f32(const float *img, __half *patches, int B, cudaStream_t s) {
    CB_CUDA(launch_ex(preprocess_f32_kernel, dim3(grid_for((int64_t)B * 224 * 56, 256)), dim3(256), 0, s, 1, true, img, patches, B));
    CB_LAUNCH_CHECK();
    return CB_OK;
}
int layernorm_f16(const __half *in, __half *out, const float *gamma, const float *beta, int rows, int width,
                  int in_row_stride, const int *gather, const float *cls_fill, int cls_period, cudaStream_t s,
                  float *stats_out) {
    CB_REQUIRE(width == 768 || width == 512, "layernorm_f16: width %d not supported", width);
    if (rows == 0) return CB_OK;
    int per_sm = 6;
    if (tune(T_LN_BLOCKS_PER_SM) > 0) per_sm = (int)tune(T_LN_BLOCKS_PER_SM);
    const int grid = std::min((rows + 7) / 8, kNumSMs * per_sm);
    if (cls_period <= 0) cls_period = 1;
    if (width == 768)
        CB_CUDA(launch_ex(layernorm_kernel<768>, dim3(grid), dim3(256), 0, s, 1, true, in, out, gamma, beta, rows, in_row_stride,
                          gather, cls_fill, cls_period, stats_out));
    else
        CB_CUDA(launch_ex(layernorm_kernel<512>, dim3(grid), dim3(256), 0, s, 1, true, in, out, gamma, beta, rows, in_row_stride,
                          gather, cls_fill, cls_period, stats_out));
    CB_LAUNCH_CHECK();
    return CB_OK;
}
int l2norm_rows_f32(const float *in, float *out, int rows, int width, cudaStream_t s) {
    CB_REQUIRE(width % 4 == 0, "l2norm_rows_f32: width must be a multiple of 4");
    if (rows == 0) return CB_OK;
    CB_CUDA(launch_ex(l2norm_kernel, dim3((rows + 7) / 8), dim3(256), 0, s, 1, true, in, out, rows, width));
    CB_LAUNCH_CHECK();
    return CB_OK;
}
int text_embed(const int32_t *ids, const float *tok_emb, const float *pos_emb, __half *x, int *eot_row, int B,
               int ctx, int width, int vocab, cudaStream_t s, float *stats_out) {
    if (B == 0) return CB_OK;
    CB_CUDA(launch_ex(text_embed_kernel, dim3(B), dim3(256), 0, s, 1, true, ids, tok_emb, pos_emb, x, eot_row, ctx, width, vocab,
                      stats_out));
    CB_LAUNCH_CHECK();
    return CB_OK;
}

}  // namespace cb
