// Fused multi-head attention for CLIP's tiny sequences (L = 50 vision tokens, no mask;
// L = 77 text tokens, additive causal mask), d_head = 64.
//
// Replaces nn.MultiheadAttention's softmax(q k^T / sqrt(64)) v inside openai/CLIP's
// ResidualAttentionBlock (reference call sites
// /root/reference/build-index.py:49, query-index.py:108).  ~1 % of the model FLOPs: one
// CTA per (image, head), Q/K/V of that head staged once in shared memory, one warp per
// 16 query rows, S = QK^T and O = PV on mma.sync m16n8k16, softmax in fp32 registers, P re-used as
// the A operand straight from the accumulator fragments (attention_kernel: the text tower and the
// fallback for other shapes).
//
// The vision tower at batch size (L = 50, no mask) runs attention_pair_kernel instead: TWO images of one
// head share a 128-row tcgen05 tile (rows 0..49 and 64..113), Q/K/V arrive by TMA, S and O live in tensor
// memory, softmax is one row per thread with no shuffles.  The mma.sync kernel was bound by the latency of
// its dependent chain at 50 % occupancy (4.3 us per CTA alone, ~500 instructions per warp for 16 rows); the
// pair kernel issues ~4x fewer instructions per row.
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <mutex>

#include "common.cuh"
#include "tc_ptx.cuh"
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
attention_kernel(const __half *__restrict__ qkv, __half *__restrict__ out, int heads, int pairs) {
    constexpr int KT = (L + 15) / 16;   // 16-key steps (= 16-row query tiles = warps)
    constexpr int LP = KT * 16;         // padded sequence
    constexpr int NT = 2 * KT;          // 8-key score tiles
    constexpr int NTV = (L + 7) / 8;    // score tiles that contain a real key (7 of 8 at L = 50): the rest is never computed
    // Shared memory is kept small (23.6 KB at L = 50: 9 CTAs per SM).  Q and K hold only the L real rows; the
    // padded tile rows of Q (>= L) read on into K and those of K read on into V: garbage there only
    // reaches query rows that are never stored / keys that are masked.  V keeps LP rows, the padded
    // ones zero, because P (= 0 there) still multiplies them.
    __shared__ __align__(16) __half sbuf[(2 * L + LP) * LDS];
    __half *sQ = sbuf, *sK = sbuf + L * LDS, *sV = sbuf + 2 * L * LDS;

    const int W = heads * HD;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    pdl_wait();                     // qkv comes from the previous kernel of the stream (common.cuh)
    // the grid is capped at what is resident at once; a CTA walks over (sequence, head) pairs
    for (int pairi = blockIdx.x; pairi < pairs; pairi += gridDim.x) {
    const int b = pairi / heads, h = pairi % heads;
    if (pairi != (int)blockIdx.x) __syncthreads();                    // the previous pair's tiles are consumed

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
        if (nt >= NTV) continue;                                  // both tiles are padding keys
#pragma unroll
        for (int ks = 0; ks < 4; ks++) {
            uint32_t kb[4];
            ldsm_x4(kb, sK + (nt * 8 + (lane & 7) + (lane >> 4) * 8) * LDS + ks * 16 + ((lane >> 3) & 1) * 8);
            mma16816(s[nt], qa[ks], kb[0], kb[1]);
            if (nt + 1 < NTV) mma16816(s[nt + 1], qa[ks], kb[2], kb[3]);
        }
    }

    // softmax over keys, fp32; 1/sqrt(64) folded into the exp2 argument
    const float sl2 = 0.125f * 1.4426950408889634f;
    const int r0 = warp * 16 + (lane >> 2), r1 = r0 + 8;
    float m0 = -INFINITY, m1 = -INFINITY;
#pragma unroll
    for (int nt = 0; nt < NTV; nt++) {
#pragma unroll
        for (int e = 0; e < 4; e++) {
            const int key = nt * 8 + (lane & 3) * 2 + (e & 1);
            const int row = e < 2 ? r0 : r1;
            const bool masked = ((nt + 1) * 8 > L && key >= L) || (CAUSAL && key > row);
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
    for (int nt = 0; nt < NTV; nt++) {
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
    }
    pdl_launch_dependents();
}

// ---- vision tower: two images per CTA on tcgen05 ------------------------------------------------------
// Tile rows: image A tokens at rows 0..49, image B tokens at rows 64..113 (warps 0-1 / 2-3: the TMEM
// column window a warp reads is warp-uniform).  S[128 x 128] = Q K^T with keys of A in columns 0..63 and
// keys of B in 64..127: a row reads only its own image's half.  P goes back into tensor memory as fp16,
// block-diagonal over the 128 keys of both images, and is the A operand of the second MMA from there (with
// P in shared memory the 128 B/clk shared-memory port was ~80 % busy: a quarter of its traffic was P).
// O[128 x 64] = P x [V_A ; V_B], V an MN-major operand straight from the row-major TMA image.  The cross
// terms of S and the zero blocks of P are wasted tensor work (about half of the ~900 tensor-pipe clocks per
// unit; what bounds the kernel is the length of the dependent chain S -> softmax -> PV -> output of each of
// the four units an SM can hold, not a pipe).  O overwrites S in tensor memory; the fp16 output tile leaves
// by TMA store.  CTAs are persistent, one per SM, with four independent groups (all 512 tensor-memory
// columns); per group a fifth warp's elected thread issues every TMA and MMA and fetches the next unit's
// Q, K as soon as S exists and its V as soon as the output tile has left; the row threads and that thread
// hand over through mbarriers only.
namespace pair {
using namespace tc;
constexpr int L = 50;
constexpr int kGroups = 4;                // independent pipelines per CTA: each owns 128 tensor-memory columns and a Q/K/V stage
constexpr int kRowWarps = 4;              // per group: one tile row per thread
constexpr int kThreads = kGroups * (kRowWarps + 1) * 32;   // + per group the warp whose elected thread issues TMA and MMA
constexpr int kTile = 128 * 128;             // bytes of a 128-row x 64-half tile (also two 64-row tiles)
constexpr int kHalf = kTile / 2;             // image B's rows start here
constexpr int kImgBytes = L * HD * 2;        // one TMA box: 50 rows x 128 B
constexpr int kStage = 3 * kTile;         // Q, K, V of one group (the output tile reuses V): 48 KB
constexpr int kSmem = kGroups * kStage + 1024 /*alignment*/ + kGroups * 64 + 64 /*barriers + tmem slot*/;

__device__ __forceinline__ void tmem_ld_x16(uint32_t taddr, uint32_t *v) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
          "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
        : "r"(taddr)
        : "memory");
}
__device__ __forceinline__ void tmem_ld_x2(uint32_t taddr, uint32_t *v) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x2.b32 {%0, %1}, [%2];" : "=r"(v[0]), "=r"(v[1]) : "r"(taddr) : "memory");
}
__device__ __forceinline__ void tmem_st_x32(uint32_t taddr, const uint32_t *v) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
        "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
        "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};"
        ::"r"(taddr), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]),
          "r"(v[8]), "r"(v[9]), "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15]),
          "r"(v[16]), "r"(v[17]), "r"(v[18]), "r"(v[19]), "r"(v[20]), "r"(v[21]), "r"(v[22]), "r"(v[23]),
          "r"(v[24]), "r"(v[25]), "r"(v[26]), "r"(v[27]), "r"(v[28]), "r"(v[29]), "r"(v[30]), "r"(v[31])
        : "memory");
}
// tight poll (mbarrier.test_wait): the hand-overs of this kernel are a few hundred clocks apart, so the
// suspend time of try_wait would be a visible part of every round trip
__device__ __forceinline__ void mbar_spin(uint64_t *bar, uint32_t parity) {
    const uint32_t addr = smem_u32(bar);
    uint32_t ok = 0, spins = 0;
    while (true) {
        asm volatile(
            "{\n"
            ".reg .pred p;\n"
            "mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n"
            "selp.u32 %0, 1, 0, p;\n"
            "}\n"
            : "=r"(ok)
            : "r"(addr), "r"(parity)
            : "memory");
        if (ok) break;
        if (++spins > kSpinLimit) __trap();
    }
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
// D[tmem] (+)= A[tmem] * B[smem]: A is fp16, two K elements per 32-bit column, one row per lane
__device__ __forceinline__ void umma_f16_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "setp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n"
        "}\n" ::"r"(tmem_d), "r"(tmem_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
        : "memory");
}
// MN-major operand, 128B swizzle: 8-element (16 B) chunks of the N dimension contiguous, rows of 128 B along K,
// 8-row groups SBO = 1024 B apart, the second 64 columns of N LBO = 8192 B away (canonical layout
// ((8,n),(8,k)):((1,LBO),(8,SBO)) in 16-byte units)
__device__ __forceinline__ uint64_t make_smem_desc_mn(uint32_t smem_addr) {
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr >> 4) & 0x3fff);
    d |= (uint64_t)(kHalf >> 4) << 16;
    d |= (uint64_t)(1024 >> 4) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)2 << 61;
    return d;
}
constexpr uint32_t kIdescS = make_idesc(128, 128);                 // Q, K both K-major
constexpr uint32_t kIdescO = make_idesc(128, 64) | (1u << 16);     // P K-major (tensor memory), V MN-major
constexpr uint32_t kPCol = 64;                                     // P: tensor-memory columns 64..127 (fp16 pairs of 128 keys)

__global__ void __launch_bounds__(kThreads, 1)
attention_pair_kernel(const __grid_constant__ CUtensorMap tm_qkv, const __grid_constant__ CUtensorMap tm_out, int B,
                      int heads, int units) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t *smem = reinterpret_cast<uint8_t *>(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    const int tid = threadIdx.x, warp = tid >> 5;
    const int W = heads * HD;
    // warps 0 .. 4 kGroups - 1: row warps, four per group (warp % 4 = the TMEM lane quadrant a warp may touch);
    // warps 4 kGroups ..: one issuing warp per group
    const bool is_row = warp < kGroups * kRowWarps;
    const int grp = is_row ? warp / kRowWarps : warp - kGroups * kRowWarps;
    uint8_t *sQ = smem + grp * kStage, *sK = sQ + kTile, *sV = sQ + 2 * kTile;
    uint8_t *sPO = sV;                // the fp16 output tile is written over V once O exists
    uint64_t *bars = reinterpret_cast<uint64_t *>(smem + kGroups * kStage) + grp * 8;
    uint64_t *full_qk = bars + 0;     // Q, K landed (TMA bytes)
    uint64_t *full_v = bars + 1;      // V landed
    uint64_t *s_done = bars + 2;      // S in tensor memory (MMA commit)
    uint64_t *o_done = bars + 3;      // O in tensor memory (MMA commit): P and V are consumed
    uint64_t *p_ready = bars + 4;     // 128 arrivals: P in tensor memory, S read
    uint64_t *o_ready = bars + 5;     // 128 arrivals: output tile written, O read
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(smem + kGroups * kStage + kGroups * 64);

    if (!is_row) {
        if (elect_one()) {
            for (int i = 0; i < 4; i++) mbar_init(bars + i, 1);
            mbar_init(p_ready, kRowWarps * 32);
            mbar_init(o_ready, kRowWarps * 32);
            fence_barrier_init();
            if (grp == 0) {
                tma_prefetch_desc(&tm_qkv);
                tma_prefetch_desc(&tm_out);
            }
        }
    } else if (warp == 0) {
        tmem_alloc<1>(tmem_slot, 512);
    }
    // Rows of V the loads never write (tokens 50..63 of either image; the whole second tile when the batch is
    // odd) must be finite: P is exactly 0 there, and 0 x NaN is not.  Zeroed once; the loads come after the barrier.
    for (int g = 0; g < kGroups; g++)
        for (int i = tid; i < kTile / 16; i += kThreads) reinterpret_cast<uint4 *>(smem + g * kStage + 2 * kTile)[i] = make_uint4(0, 0, 0, 0);
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *tmem_slot + (uint32_t)(grp * 128);        // this group's 128 columns
    // Dependents are released as soon as every CTA of this grid is resident (none is left to schedule, so a
    // dependent that parks on an SM with its shared and tensor memory cannot starve this grid); they wait for
    // this grid's completion in their own griddepcontrol.wait before touching global memory.
    pdl_launch_dependents();
    pdl_wait();                     // qkv comes from the previous kernel of the stream

    // unit u = (pair of images, head); group g of CTA c takes u = g gridDim.x + c, + kGroups gridDim.x, ... (the
    // left-over units of the last round spread over CTAs, not over the groups of a few)
    const int u0 = grp * gridDim.x + blockIdx.x, ustride = gridDim.x * kGroups;
    if (!is_row) {
        // ---- one thread: TMA loads, both MMAs, the output store -------------------------------------
        if (elect_one()) {
            auto load_qk = [&](int u) {
                const int pr = u / heads, h = u % heads, imgA = 2 * pr;
                const bool hasB = imgA + 1 < B;
                mbar_expect_tx(full_qk, (hasB ? 4 : 2) * kImgBytes);
                tma_load_2d(sQ, &tm_qkv, h * HD, imgA * L, full_qk);
                tma_load_2d(sK, &tm_qkv, W + h * HD, imgA * L, full_qk);
                if (hasB) {
                    tma_load_2d(sQ + kHalf, &tm_qkv, h * HD, (imgA + 1) * L, full_qk);
                    tma_load_2d(sK + kHalf, &tm_qkv, W + h * HD, (imgA + 1) * L, full_qk);
                }
            };
            auto load_v = [&](int u) {
                const int pr = u / heads, h = u % heads, imgA = 2 * pr;
                const bool hasB = imgA + 1 < B;
                mbar_expect_tx(full_v, (hasB ? 2 : 1) * kImgBytes);
                tma_load_2d(sV, &tm_qkv, 2 * W + h * HD, imgA * L, full_v);
                if (hasB) tma_load_2d(sV + kHalf, &tm_qkv, 2 * W + h * HD, (imgA + 1) * L, full_v);
            };
            auto issue_s = [&](uint32_t ph) {                         // S = Q K^T once Q, K have landed
                mbar_spin(full_qk, ph);
                tc_fence_after();
#pragma unroll
                for (int k = 0; k < HD / 16; k++)
                    umma_f16<1>(tmem, make_smem_desc(smem_u32(sQ) + k * 32), make_smem_desc(smem_u32(sK) + k * 32), kIdescS, k > 0);
                umma_commit<1>(s_done);
            };
            if (u0 < units) {
                load_qk(u0);
                load_v(u0);
                issue_s(0);
            }
            uint32_t ph = 0;
            for (int u = u0; u < units; u += ustride, ph ^= 1) {
                const int un = u + ustride;                           // this group's next unit
                const int pr = u / heads, h = u % heads, imgA = 2 * pr;
                mbar_spin(s_done, ph);                                // Q, K consumed: fetch the next unit's
                if (un < units) load_qk(un);
                mbar_spin(p_ready, ph);                               // P in shared memory, S read by every row
                mbar_spin(full_v, ph);
                tc_fence_after();
#pragma unroll
                for (int k = 0; k < 128 / 16; k++)
                    umma_f16_ts(tmem, tmem + kPCol + k * 8, make_smem_desc_mn(smem_u32(sV) + k * 2048), kIdescO, k > 0);
                umma_commit<1>(o_done);
                mbar_spin(o_ready, ph);                               // output tile (over V) complete, O read by every row
                if (un < units) issue_s(ph ^ 1);                      // the next S overwrites O
                else tc_fence_after();
                asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];"
                             ::"l"(reinterpret_cast<uint64_t>(&tm_out)), "r"(smem_u32(sPO)), "r"(h * HD), "r"(imgA * L) : "memory");
                if (imgA + 1 < B)
                    asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];"
                                 ::"l"(reinterpret_cast<uint64_t>(&tm_out)), "r"(smem_u32(sPO) + kHalf), "r"(h * HD), "r"((imgA + 1) * L) : "memory");
                asm volatile("cp.async.bulk.commit_group;" ::: "memory");
                asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
                if (un < units) load_v(un);                           // the tile is read: V's buffer is free
            }
            asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
        }
        __syncwarp();
    } else {
        // ---- 128 threads: one row of the tile each ---------------------------------------------------
        const int row = tid & 127;                                    // row of the group's tile
        const int img = row >> 6, tok = row & 63;
        const uint32_t tlane = tmem + ((uint32_t)((warp & 3) * 32) << 16);
        const uint32_t trow = tlane + (uint32_t)(img * 64);          // S: the own image's keys
        const uint32_t prow = smem_u32(sPO) + row * 128;              // this thread's row of the output tile
        const float sl2 = 0.125f * 1.4426950408889634f;              // 1/sqrt(64) folded into the exp2 argument
        uint32_t ph = 0;
        for (int u = u0; u < units; u += ustride, ph ^= 1) {
            const bool hasB = 2 * (u / heads) + 1 < B;
            const bool valid = tok < L && (img == 0 || hasB);
            mbar_spin(s_done, ph);
            tc_fence_after();
            uint32_t sr[52];
            {
                uint32_t (&a)[32] = *reinterpret_cast<uint32_t (*)[32]>(&sr[0]);
                tmem_ld_32x32(trow, a);
                tmem_ld_x16(trow + 32, &sr[32]);
                tmem_ld_x2(trow + 48, &sr[48]);
                tmem_ld_wait();
            }
            float m4[4] = {__uint_as_float(sr[0]), __uint_as_float(sr[1]), __uint_as_float(sr[2]), __uint_as_float(sr[3])};
#pragma unroll
            for (int j = 4; j < L; j++) m4[j & 3] = fmaxf(m4[j & 3], __uint_as_float(sr[j]));   // four chains, not one of 49
            const float m = fmaxf(fmaxf(m4[0], m4[1]), fmaxf(m4[2], m4[3]));
            const float c = -m * sl2;
            float s4[4] = {0.f, 0.f, 0.f, 0.f};
            uint32_t pk[32];                                          // 64 keys as fp16 pairs; keys 50..63 are 0
#pragma unroll
            for (int j = 0; j < L; j += 2) {
                float p0, p1;
                asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(p0) : "f"(fmaf(__uint_as_float(sr[j]), sl2, c)));
                asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(p1) : "f"(fmaf(__uint_as_float(sr[j + 1]), sl2, c)));
                if (!valid) p0 = p1 = 0.f;                            // padding rows: S is garbage there
                s4[(j >> 1) & 3] += p0 + p1;
                pk[j >> 1] = pack_h2(p0, p1);
            }
            const float sum = (s4[0] + s4[1]) + (s4[2] + s4[3]);
            // P of this row, block-diagonal over the 128 keys of both images: the own image's 32 pairs (keys
            // 50..63 zero), zeros for the other image's keys.  It overwrites columns of S only this thread
            // reads (its own lane), and lies outside the 64 columns O is written to.
#pragma unroll
            for (int j = 25; j < 32; j++) pk[j] = 0u;
            uint32_t zz[32];
#pragma unroll
            for (int j = 0; j < 32; j++) zz[j] = 0u;
            tmem_st_x32(tlane + kPCol + img * 32, pk);
            tmem_st_x32(tlane + kPCol + (img ^ 1) * 32, zz);
            tmem_st_wait();
            tc_fence_before();
            mbar_arrive(p_ready);

            mbar_spin(o_done, ph);
            tc_fence_after();
            const float inv = valid ? 1.0f / sum : 0.f;               // padding rows stay zero: the tile lies over V, whose
                                                                      // padding rows must stay finite for the next unit
#pragma unroll
            for (int half = 0; half < 2; half++) {                    // this row of O, normalised, fp16
                uint32_t o[32];
                tmem_ld_32x32(tlane + half * 32, o);
                tmem_ld_wait();
#pragma unroll
                for (int ch = 0; ch < 4; ch++) {
                    uint4 v;
                    v.x = pack_h2(__uint_as_float(o[8 * ch + 0]) * inv, __uint_as_float(o[8 * ch + 1]) * inv);
                    v.y = pack_h2(__uint_as_float(o[8 * ch + 2]) * inv, __uint_as_float(o[8 * ch + 3]) * inv);
                    v.z = pack_h2(__uint_as_float(o[8 * ch + 4]) * inv, __uint_as_float(o[8 * ch + 5]) * inv);
                    v.w = pack_h2(__uint_as_float(o[8 * ch + 6]) * inv, __uint_as_float(o[8 * ch + 7]) * inv);
                    sts128(prow + (((half * 4 + ch) ^ (row & 7)) << 4), v);
                }
            }
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            tc_fence_before();
            mbar_arrive(o_ready);
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) {
        tc_fence_after();
        tmem_dealloc<1>(*tmem_slot, 512);
    }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                  const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
// fp16 [rows, cols] row-major, box = 50 rows x 64 columns (one head of one image), 128B swizzle
int make_head_map(CUtensorMap *m, const void *base, uint64_t rows, uint64_t cols) {
    static EncodeTiledFn fn = nullptr;
    if (!fn) {
        void *p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess ||
            q != cudaDriverEntryPointSuccess) {
            set_error("cuTensorMapEncodeTiled entry point not available");
            return CB_ERR_CUDA;
        }
        fn = reinterpret_cast<EncodeTiledFn>(p);
    }
    cuuint64_t gdim[2] = {cols, rows};
    cuuint64_t gstride[1] = {cols * 2};
    cuuint32_t box[2] = {(cuuint32_t)HD, (cuuint32_t)L};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = fn(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, const_cast<void *>(base), gdim, gstride, box, estr,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled (attention) failed (%d)", (int)r); return CB_ERR_CUDA; }
    return CB_OK;
}

int launch(const __half *qkv, __half *out, int B, int heads, cudaStream_t s) {
    CUtensorMap tq, to;
    int rc;
    if ((rc = make_head_map(&tq, qkv, (uint64_t)B * L, (uint64_t)3 * heads * HD))) return rc;
    if ((rc = make_head_map(&to, out, (uint64_t)B * L, (uint64_t)heads * HD))) return rc;
    int dev = 0;
    CB_CUDA(cudaGetDevice(&dev));
    static std::once_flag once[64];        // function attributes are per device
    cudaError_t err = cudaSuccess;
    std::call_once(once[dev & 63], [&] {
        err = cudaFuncSetAttribute(attention_pair_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmem);
        // 193 KB of the 228: ask for the largest shared-memory carve-out
        if (err == cudaSuccess)
            err = cudaFuncSetAttribute(attention_pair_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
    });
    CB_CUDA(err);
    // persistent, one CTA per SM (measured: with one CTA per unit, the CTAs of an SM started ~1200 clocks apart, and
    // each paid barrier set-up and a tensor-memory allocation)
    const int units = ((B + 1) / 2) * heads;
    const int grid = std::min((units + kGroups - 1) / kGroups, kNumSMs);
    CB_CUDA(launch_ex(attention_pair_kernel, dim3((unsigned)grid), dim3(kThreads), (size_t)kSmem, s, 1, true, tq, to, B, heads, units));
    return CB_OK;
}
}  // namespace pair

}  // namespace

int attention_f16(const __half *qkv, __half *out, int B, int L, int heads, bool causal, cudaStream_t s) {
    if (B == 0) return CB_OK;
    // TMA needs 16-byte aligned tensors; a lone image would leave half of every tile empty
    if (L == 50 && !causal && B >= 2 && tune(T_ATTN_TC) != 0 && (((uintptr_t)qkv | (uintptr_t)out) & 15) == 0) {
        int rc = pair::launch(qkv, out, B, heads, s);
        if (rc) return rc;
    } else if (L == 50 && !causal) {
        CB_CUDA(launch_ex(attention_kernel<50, false>, dim3(std::min(B * heads, 8 * kNumSMs)), dim3(4 * 32), 0, s, 1, true, qkv, out, heads, B * heads));
    } else if (L == 77 && causal) {
        CB_CUDA(launch_ex(attention_kernel<77, true>, dim3(std::min(B * heads, 6 * kNumSMs)), dim3(5 * 32), 0, s, 1, true, qkv, out, heads, B * heads));
    } else if (L == 77 && !causal) {
        CB_CUDA(launch_ex(attention_kernel<77, false>, dim3(std::min(B * heads, 6 * kNumSMs)), dim3(5 * 32), 0, s, 1, true, qkv, out, heads, B * heads));
    } else if (L == 50 && causal) {
        CB_CUDA(launch_ex(attention_kernel<50, true>, dim3(std::min(B * heads, 8 * kNumSMs)), dim3(4 * 32), 0, s, 1, true, qkv, out, heads, B * heads));
    } else {
        set_error("attention_f16: sequence length %d not supported (50 or 77)", L);
        return CB_ERR_INVALID;
    }
    CB_LAUNCH_CHECK();
    return CB_OK;
}

}  // namespace cb
