// Fused multi-head attention for CLIP's tiny sequences (L = 50 vision tokens, no mask;
// L = 77 text tokens, additive causal mask), d_head = 64.
//
// Replaces nn.MultiheadAttention's softmax(q k^T / sqrt(64)) v inside openai/CLIP's
// ResidualAttentionBlock (reference call sites
// /root/reference/build-index.py:49, query-index.py:108).  ~1 % of the model FLOPs: one
// CTA per (image, head), Q/K/V of that head staged once in shared memory, one warp per
// 16 query rows, S = QK^T and O = PV on mma.sync m16n8k16 (legacy tensor path is ample
// here; the tcgen05 budget goes to the GEMMs), softmax in fp32 registers, P re-used as
// the A operand straight from the accumulator fragments.
#include "common.cuh"
#include "vit_kernels.cuh"

namespace cb {
namespace {

constexpr int HD = 64;          // head dim
constexpr int LDS = 72;         // smem row stride in halfs (144 B: conflict-free ldmatrix)

__device__ __forceinline__ void ldsm_x4(uint32_t (&r)[4], const __half *p) {
    uint32_t a = (uint32_t)__cvta_generic_to_shared(p);
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(a));
}
__device__ __forceinline__ void ldsm_x4_t(uint32_t (&r)[4], const __half *p) {
    uint32_t a = (uint32_t)__cvta_generic_to_shared(p);
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(a));
}
__device__ __forceinline__ void mma16816(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ uint32_t pack2(float a, float b) {
    __half2 h = __floats2half2_rn(a, b);
    return *reinterpret_cast<uint32_t *>(&h);
}

template <int L, bool CAUSAL>
__global__ void __launch_bounds__(((L + 15) / 16) * 32, L <= 64 ? 8 : 6)
attention_kernel(const __half *__restrict__ qkv, __half *__restrict__ out, int heads) {
    constexpr int KT = (L + 15) / 16;   // 16-key steps (= 16-row query tiles = warps)
    constexpr int LP = KT * 16;         // padded sequence
    constexpr int NT = 2 * KT;          // 8-key score tiles
    // Shared memory is kept small (23.6 KB at L = 50: 9 CTAs per SM).  Q and K hold only the L real rows; the
    // padded tile rows of Q (>= L) read on into K and those of K read on into V: garbage there only
    // reaches query rows that are never stored / keys that are masked.  V keeps LP rows, the padded
    // ones zero, because P (= 0 there) still multiplies them.
    __shared__ __align__(16) __half sbuf[(2 * L + LP) * LDS];
    __half *sQ = sbuf, *sK = sbuf + L * LDS, *sV = sbuf + 2 * L * LDS;

    const int b = blockIdx.x / heads, h = blockIdx.x % heads;
    const int W = heads * HD;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    pdl_wait();                     // qkv comes from the previous kernel of the stream (common.cuh)

    // stage q|k|v of this head: 3 x LP rows x 8 chunks of 16 B, global -> shared with cp.async (no register
    // round trip); rows >= L of V are zero.  blockDim = 8 chunks x (4 KT) rows, so iteration `it` of the
    // fully unrolled loop covers rows (it & 3) * 4 KT .. of matrix it >> 2: no index arithmetic per chunk
    {
        const int r8 = threadIdx.x >> 3, ch = threadIdx.x & 7;
        const __half *src0 = qkv + ((size_t)(b * L + r8) * 3) * W + h * HD + ch * 8;
        __half *dst0 = sbuf + r8 * LDS + ch * 8;
#pragma unroll
        for (int it = 0; it < 12; it++) {
            const int mat = it >> 2, rb = (it & 3) * (KT * 4);        // compile-time after unrolling
            const int row = r8 + rb;
            __half *dst = dst0 + ((mat == 0 ? 0 : (mat == 1 ? L : 2 * L)) + rb) * LDS;
            if (row < L) {
                const __half *src = src0 + ((size_t)rb * 3 + mat) * W;
                asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((uint32_t)__cvta_generic_to_shared(dst)), "l"(src) : "memory");
            } else if (mat == 2) {
                *reinterpret_cast<uint4 *>(dst) = make_uint4(0, 0, 0, 0);
            }
        }
    }
    asm volatile("cp.async.commit_group;\ncp.async.wait_group 0;" ::: "memory");
    __syncthreads();

    // Q fragments of this warp's 16 rows
    uint32_t qa[4][4];
#pragma unroll
    for (int ks = 0; ks < 4; ks++)
        ldsm_x4(qa[ks], sQ + (warp * 16 + (lane & 15)) * LDS + ks * 16 + (lane >> 4) * 8);

    float s[NT][4];
#pragma unroll
    for (int nt = 0; nt < NT; nt++) { s[nt][0] = s[nt][1] = s[nt][2] = s[nt][3] = 0.f; }
#pragma unroll
    for (int nt = 0; nt < NT; nt += 2) {
#pragma unroll
        for (int ks = 0; ks < 4; ks++) {
            uint32_t kb[4];
            ldsm_x4(kb, sK + (nt * 8 + (lane & 7) + (lane >> 4) * 8) * LDS + ks * 16 + ((lane >> 3) & 1) * 8);
            mma16816(s[nt], qa[ks], kb[0], kb[1]);
            mma16816(s[nt + 1], qa[ks], kb[2], kb[3]);
        }
    }

    // softmax over keys, fp32; 1/sqrt(64) folded into the exp2 argument
    const float sl2 = 0.125f * 1.4426950408889634f;
    const int r0 = warp * 16 + (lane >> 2), r1 = r0 + 8;
    float m0 = -INFINITY, m1 = -INFINITY;
#pragma unroll
    for (int nt = 0; nt < NT; nt++) {
#pragma unroll
        for (int e = 0; e < 4; e++) {
            const int key = nt * 8 + (lane & 3) * 2 + (e & 1);
            const int row = e < 2 ? r0 : r1;
            const bool masked = key >= L || (CAUSAL && key > row);
            const float v = masked ? -INFINITY : s[nt][e];        // raw score; the scale rides in the FMA below
            s[nt][e] = v;
            if (e < 2) m0 = fmaxf(m0, v); else m1 = fmaxf(m1, v);
        }
    }
    m0 = fmaxf(m0, __shfl_xor_sync(0xffffffffu, m0, 1)); m0 = fmaxf(m0, __shfl_xor_sync(0xffffffffu, m0, 2));
    m1 = fmaxf(m1, __shfl_xor_sync(0xffffffffu, m1, 1)); m1 = fmaxf(m1, __shfl_xor_sync(0xffffffffu, m1, 2));
    float sum0 = 0.f, sum1 = 0.f;
    const float c0 = -m0 * sl2, c1 = -m1 * sl2;                   // every row has an unmasked key: m is finite
#pragma unroll
    for (int nt = 0; nt < NT; nt++) {
#pragma unroll
        for (int e = 0; e < 4; e++) {
            float p;                                               // ex2.approx(-inf) = 0 for masked keys
            asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(p) : "f"(fmaf(s[nt][e], sl2, e < 2 ? c0 : c1)));
            s[nt][e] = p;
            if (e < 2) sum0 += p; else sum1 += p;
        }
    }
    sum0 += __shfl_xor_sync(0xffffffffu, sum0, 1); sum0 += __shfl_xor_sync(0xffffffffu, sum0, 2);
    sum1 += __shfl_xor_sync(0xffffffffu, sum1, 1); sum1 += __shfl_xor_sync(0xffffffffu, sum1, 2);

    // O = P V
    float o[8][4];
#pragma unroll
    for (int dt = 0; dt < 8; dt++) { o[dt][0] = o[dt][1] = o[dt][2] = o[dt][3] = 0.f; }
#pragma unroll
    for (int kt = 0; kt < KT; kt++) {
        uint32_t pa[4];
        pa[0] = pack2(s[2 * kt][0], s[2 * kt][1]);
        pa[1] = pack2(s[2 * kt][2], s[2 * kt][3]);
        pa[2] = pack2(s[2 * kt + 1][0], s[2 * kt + 1][1]);
        pa[3] = pack2(s[2 * kt + 1][2], s[2 * kt + 1][3]);
#pragma unroll
        for (int dt = 0; dt < 8; dt += 2) {
            uint32_t vb[4];
            ldsm_x4_t(vb, sV + (kt * 16 + (lane & 7) + ((lane >> 3) & 1) * 8) * LDS + dt * 8 + (lane >> 4) * 8);
            mma16816(o[dt], pa, vb[0], vb[1]);
            mma16816(o[dt + 1], pa, vb[2], vb[3]);
        }
    }
    const float i0 = 1.0f / sum0, i1 = 1.0f / sum1;
#pragma unroll
    for (int dt = 0; dt < 8; dt++) {
        const int col = h * HD + dt * 8 + (lane & 3) * 2;
        if (r0 < L) *reinterpret_cast<uint32_t *>(out + (size_t)(b * L + r0) * W + col) = pack2(o[dt][0] * i0, o[dt][1] * i0);
        if (r1 < L) *reinterpret_cast<uint32_t *>(out + (size_t)(b * L + r1) * W + col) = pack2(o[dt][2] * i1, o[dt][3] * i1);
    }
    pdl_launch_dependents();
}

}  // namespace

int attention_f16(const __half *qkv, __half *out, int B, int L, int heads, bool causal, cudaStream_t s) {
    if (B == 0) return CB_OK;
    if (L == 50 && !causal) {
        CB_CUDA(launch_ex(attention_kernel<50, false>, dim3(B * heads), dim3(4 * 32), 0, s, 1, true, qkv, out, heads));
    } else if (L == 77 && causal) {
        CB_CUDA(launch_ex(attention_kernel<77, true>, dim3(B * heads), dim3(5 * 32), 0, s, 1, true, qkv, out, heads));
    } else if (L == 77 && !causal) {
        CB_CUDA(launch_ex(attention_kernel<77, false>, dim3(B * heads), dim3(5 * 32), 0, s, 1, true, qkv, out, heads));
    } else if (L == 50 && causal) {
        CB_CUDA(launch_ex(attention_kernel<50, true>, dim3(B * heads), dim3(4 * 32), 0, s, 1, true, qkv, out, heads));
    } else {
        set_error("attention_f16: sequence length %d not supported (50 or 77)", L);
        return CB_ERR_INVALID;
    }
    CB_LAUNCH_CHECK();
    return CB_OK;
}

}  // namespace cb
