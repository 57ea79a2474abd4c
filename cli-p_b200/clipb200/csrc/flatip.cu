// Exact inner-product top-k over an N x 512 shard resident in HBM.
//
// Replaces faiss.IndexFlatIP.search as called at /root/reference/query-index.py:111
// (index built at /root/reference/build-index.py:80-81,99,107).
//
// Small-nq path (this file): HBM-bound.  One pass over the shard with coalesced
// 128-bit streaming loads, fp32 FMA on CUDA cores (free when HBM-bound), a
// transposed warp-shuffle reduction, scores written once (4 B/row = 0.4 % of the
// 1024 B/row read) and an exact radix select over the fp32 scores:
//
//   K1 scan      : scores[q][i] = <q, x_i>; fused histogram of the top 11 key bits
//   K2 refine x2 : histogram of the next 11 / last 10 key bits inside the selected bin
//                  (reads only the 4 B/row scores, L2-resident; skipped once the bin is
//                  taken whole)
//   K3 collect   : gather the k winners, order ties by id, bitonic sort, write D/I
//
// The "find the bin that holds the k-th key" step of every kernel runs in its last
// block to retire (threadfence + atomic ticket), so a search is 4 launches, no host
// round trip, any k from 1 to ntotal.  Order is the total order (-score, id): results
// do not depend on the launch geometry, so 1-GPU and sharded results are identical.
#include "common.cuh"

#include <algorithm>
#include <cstdarg>
#include <cstring>
#include <new>

namespace cb {

// tensor-core batch path (flatip_batch.cu)
struct BatchWs;
BatchWs *batch_ws_new();
void batch_ws_delete(BatchWs *w);
int flatip_search_batch(BatchWs *w, const void *rows_f16, int64_t n, int device, int64_t nq, const float *q_dev,
                        int64_t k, float *D_dev, int64_t *I_dev, int64_t id_base, cudaStream_t s,
                        bool *overflowed);

constexpr int kD = 512;
constexpr int kBins0 = 2048;           // 11 + 11 + 10 bit digits
constexpr int kMaxNQ = 4;              // queries sharing one pass over the shard (register budget)
constexpr int kSortSmem = 4096;        // composite keys sorted in shared memory
constexpr int kScanThreads = 256;
constexpr int kPostThreads = 256;

struct SelState {
    uint32_t prefix;     // selected key prefix, right aligned, `bits` wide
    uint32_t bits;       // prefix bits fixed so far: 0, 11, 22, 32
    uint32_t k_rem;      // winners still to take from inside the prefix bin
    uint32_t cnt_bin;    // elements inside the prefix bin
    uint32_t done;       // 1: cnt_bin == k_rem -> whole bin wins, stop refining
    uint32_t n_cand;     // collect cursor
    uint32_t ticket[4];  // block-retire counters, one per kernel
    uint32_t pad[6];
};
static_assert(sizeof(SelState) == 64, "SelState layout");

// workspace per query: 3 histograms + state
struct QueryWs {
    uint32_t hist[3][kBins0];
    SelState st;
};

// ---------------------------------------------------------------------------------
// find the bin that holds the k_rem-th largest key; one warp, bins scanned from the top
__device__ void find_bin(const uint32_t *hist, int nbins, int digit_bits, SelState *st) {
    const int lane = threadIdx.x & 31;
    const int per = nbins / 32;
    const int hi = nbins - 1 - lane * per;
    uint32_t s = 0;
    for (int i = 0; i < per; i++) s += __ldcg(hist + hi - i);
    uint32_t incl = s;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        uint32_t t = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += t;
    }
    const uint32_t excl = incl - s;
    const uint32_t k_rem = st->k_rem;
    if (excl < k_rem && incl >= k_rem) {
        uint32_t acc = excl;
        for (int i = 0; i < per; i++) {
            uint32_t c = __ldcg(hist + hi - i);
            if (acc + c >= k_rem) {
                st->prefix = (st->prefix << digit_bits) | (uint32_t)(hi - i);
                st->bits += digit_bits;
                st->k_rem = k_rem - acc;
                st->cnt_bin = c;
                st->done = (c == k_rem - acc) ? 1u : 0u;
                break;
            }
            acc += c;
        }
    }
    __syncwarp();
}

// block-retire ticket: returns true in every thread of the last block to arrive
__device__ bool last_block(uint32_t *ticket, uint32_t nblocks) {
    __shared__ bool s_last;
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) s_last = (atomicAdd(ticket, 1u) == nblocks - 1);
    __syncthreads();
    if (s_last) __threadfence();
    return s_last;
}

// ---------------------------------------------------------------------------------
// transposed reduction: R per-lane partial sums -> lane l holds the full sum of row
// (l >> (5 - log2 R)); 1 + R shuffles instead of 5 R.
template <int R>
__device__ __forceinline__ float reduce_rows(float (&v)[R], int lane) {
    int off = 16;
#pragma unroll
    for (int c = R; c > 1; c >>= 1) {
        const bool upper = (lane & off) != 0;
#pragma unroll
        for (int i = 0; i < c / 2; i++) {
            float keep = upper ? v[i + c / 2] : v[i];
            float send = upper ? v[i] : v[i + c / 2];
            v[i] = keep + __shfl_xor_sync(0xffffffffu, send, off);
        }
        off >>= 1;
    }
    float t = v[0];
    for (; off > 0; off >>= 1) t += __shfl_xor_sync(0xffffffffu, t, off);
    return t;
}

__device__ __forceinline__ float dot8_h(const uint4 &u, const float *q) {
    const __half2 *h = reinterpret_cast<const __half2 *>(&u);
    float2 a = __half22float2(h[0]), b = __half22float2(h[1]);
    float2 c = __half22float2(h[2]), d = __half22float2(h[3]);
    float s = a.x * q[0];
    s = fmaf(a.y, q[1], s);
    s = fmaf(b.x, q[2], s);
    s = fmaf(b.y, q[3], s);
    s = fmaf(c.x, q[4], s);
    s = fmaf(c.y, q[5], s);
    s = fmaf(d.x, q[6], s);
    s = fmaf(d.y, q[7], s);
    return s;
}
__device__ __forceinline__ float dot4_f(const uint4 &u, const float *q) {
    float s = __uint_as_float(u.x) * q[0];
    s = fmaf(__uint_as_float(u.y), q[1], s);
    s = fmaf(__uint_as_float(u.z), q[2], s);
    s = fmaf(__uint_as_float(u.w), q[3], s);
    return s;
}

// K1: one pass over the shard.  Row = 512 elements: fp16 -> 64 x 16 B chunks (lane
// takes chunks lane, lane+32), fp32 -> 128 chunks (lane, +32, +64, +96); either way a
// lane owns 16 elements of every row and keeps the matching 16 query values per query
// in registers.  R rows in flight per warp iteration (16 x 128-bit loads per lane).
template <int NQ, bool F16>
__global__ void __launch_bounds__(kScanThreads)
flatip_scan_kernel(const uint4 *__restrict__ xb, int64_t n, const float *__restrict__ xq,
                   int nq_valid, float *__restrict__ scores, int64_t stride, QueryWs *ws,
                   uint32_t k_eff) {
    constexpr int R = F16 ? 8 : 4;          // rows per warp iteration
    constexpr int CH = F16 ? 2 : 4;         // chunks per lane per row
    constexpr int EPC = F16 ? 8 : 4;        // elements per chunk
    constexpr int ROW_V4 = F16 ? 64 : 128;  // uint4 per row
    extern __shared__ uint32_t s_hist[];    // [NQ][kBins0]

    const int lane = threadIdx.x & 31;
    for (int i = threadIdx.x; i < NQ * kBins0; i += blockDim.x) s_hist[i] = 0;

    float qreg[NQ][CH * EPC];
#pragma unroll
    for (int q = 0; q < NQ; q++) {
        const float *qp = xq + (size_t)min(q, nq_valid - 1) * kD;
#pragma unroll
        for (int c = 0; c < CH; c++)
#pragma unroll
            for (int e = 0; e < EPC; e++) qreg[q][c * EPC + e] = qp[(lane + 32 * c) * EPC + e];
    }
    __syncthreads();

    const int64_t warps = (int64_t)gridDim.x * (blockDim.x >> 5);
    const int64_t gw = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int64_t groups = (n + R - 1) / R;
    constexpr int SH = F16 ? 2 : 3;         // lanes sharing a row after reduce = 1 << SH
    const int my_row = lane >> SH;

    for (int64_t g = gw; g < groups; g += warps) {
        const int64_t row0 = g * R;
        uint4 v[R][CH];
#pragma unroll
        for (int r = 0; r < R; r++) {
            const int64_t row = min(row0 + r, n - 1);
            const uint4 *p = xb + row * ROW_V4 + lane;
#pragma unroll
            for (int c = 0; c < CH; c++) v[r][c] = ld_stream_v4(p + 32 * c);
        }
#pragma unroll
        for (int q = 0; q < NQ; q++) {
            float acc[R];
#pragma unroll
            for (int r = 0; r < R; r++) {
                float s = 0.f;
#pragma unroll
                for (int c = 0; c < CH; c++) {
                    if (F16) s += dot8_h(v[r][c], &qreg[q][c * EPC]);
                    else     s += dot4_f(v[r][c], &qreg[q][c * EPC]);
                }
                acc[r] = s;
            }
            float tot = reduce_rows<R>(acc, lane);
            const int64_t row = row0 + my_row;
            if ((lane & ((1 << SH) - 1)) == 0 && row < n && q < nq_valid) {
                tot += 0.0f;
                scores[(size_t)q * stride + row] = tot;
                atomicAdd(&s_hist[q * kBins0 + (f2key(tot) >> 21)], 1u);
            }
        }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < NQ * kBins0; i += blockDim.x) {
        uint32_t c = s_hist[i];
        int q = i / kBins0;
        if (c && q < nq_valid) atomicAdd(&ws[q].hist[0][i - q * kBins0], c);
    }
    if (last_block(&ws[0].st.ticket[0], gridDim.x)) {
        const int w = threadIdx.x >> 5;
        for (int q = w; q < nq_valid; q += (blockDim.x >> 5)) {
            if (lane == 0) ws[q].st.k_rem = k_eff;
            __syncwarp();
            find_bin(ws[q].hist[0], kBins0, 11, &ws[q].st);
        }
    }
}

// K2: refine the selected bin by the next digit.  grid = (blocks, nq).
__global__ void __launch_bounds__(kPostThreads)
flatip_refine_kernel(const float *__restrict__ scores, int64_t n, int64_t stride, QueryWs *ws,
                     int pass /*1 or 2*/) {
    QueryWs *w = ws + blockIdx.y;
    const SelState st = w->st;
    if (st.done) return;                       // uniform over the whole grid
    const int digit_bits = pass == 1 ? 11 : 10;
    const int shift_prev = 32 - (int)st.bits;  // st.bits is 11 or 22 here
    const int shift = shift_prev - digit_bits;
    const uint32_t mask = (1u << digit_bits) - 1;
    __shared__ uint32_t s_h[kBins0];
    for (int i = threadIdx.x; i < kBins0; i += blockDim.x) s_h[i] = 0;
    __syncthreads();
    const float4 *s4 = reinterpret_cast<const float4 *>(scores + (size_t)blockIdx.y * stride);
    const int64_t n4 = (n + 3) / 4;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4;
         i += (int64_t)gridDim.x * blockDim.x) {
        float4 f = __ldcg(s4 + i);
        const float e[4] = {f.x, f.y, f.z, f.w};
#pragma unroll
        for (int j = 0; j < 4; j++) {
            if (i * 4 + j < n) {
                uint32_t key = f2key(e[j]);
                if ((key >> shift_prev) == st.prefix) atomicAdd(&s_h[(key >> shift) & mask], 1u);
            }
        }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < (1 << digit_bits); i += blockDim.x)
        if (s_h[i]) atomicAdd(&w->hist[pass][i], s_h[i]);
    if (last_block(&w->st.ticket[pass], gridDim.x)) {
        if (threadIdx.x < 32) find_bin(w->hist[pass], 1 << digit_bits, digit_bits, &w->st);
    }
}

// descending bitonic sort of n_pow2 composite keys by the whole block
__device__ void bitonic_desc(uint64_t *a, uint32_t n_pow2) {
    for (uint32_t k = 2; k <= n_pow2; k <<= 1) {
        for (uint32_t j = k >> 1; j > 0; j >>= 1) {
            for (uint32_t t = threadIdx.x; t < n_pow2 / 2; t += blockDim.x) {
                uint32_t i = ((t & ~(j - 1)) << 1) | (t & (j - 1));
                uint32_t l = i | j;
                bool desc = (i & k) == 0;
                uint64_t x = a[i], y = a[l];
                if ((x < y) == desc) { a[i] = y; a[l] = x; }
            }
            __syncthreads();
        }
    }
}

__device__ __forceinline__ uint64_t make_comp(uint32_t key, uint32_t id) {
    return ((uint64_t)key << 32) | (uint64_t)(0xffffffffu - id);
}

// write one query's sorted winners + faiss padding
__device__ void emit_sorted(const uint64_t *a, uint32_t count, int64_t k, int64_t id_base,
                            float *D, int64_t *I) {
    for (int64_t j = threadIdx.x; j < k; j += blockDim.x) {
        if (j < count) {
            uint64_t c = a[j];
            D[j] = key2f((uint32_t)(c >> 32));
            I[j] = id_base + (int64_t)(0xffffffffu - (uint32_t)c);
        } else {
            D[j] = -3.4028234663852886e38f;
            I[j] = -1;
        }
    }
}

// K3: gather winners; the last block orders exact ties by id, sorts and writes D/I.
__global__ void __launch_bounds__(kPostThreads)
flatip_collect_kernel(const float *__restrict__ scores_all, int64_t n, int64_t stride, QueryWs *ws,
                      uint64_t *cand_all, uint32_t cand_cap, int64_t k, int64_t id_base,
                      float *D_all, int64_t *I_all) {
    QueryWs *w = ws + blockIdx.y;
    const SelState st = w->st;
    const float *scores = scores_all + (size_t)blockIdx.y * stride;
    uint64_t *cand = cand_all + (size_t)blockIdx.y * cand_cap;
    const int sh = 32 - (int)st.bits;   // 0 when all 32 bits are fixed
    const float4 *s4 = reinterpret_cast<const float4 *>(scores);
    const int64_t n4 = (n + 3) / 4;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4;
         i += (int64_t)gridDim.x * blockDim.x) {
        float4 f = __ldcg(s4 + i);
        const float e[4] = {f.x, f.y, f.z, f.w};
#pragma unroll
        for (int j = 0; j < 4; j++) {
            const int64_t row = i * 4 + j;
            if (row < n) {
                uint32_t key = f2key(e[j]);
                uint32_t kp = sh ? (key >> sh) : key;
                if (kp > st.prefix || (kp == st.prefix && st.done)) {
                    uint32_t pos = atomicAdd(&w->st.n_cand, 1u);
                    if (pos < cand_cap) cand[pos] = make_comp(key, (uint32_t)row);
                }
            }
        }
    }
    if (!last_block(&w->st.ticket[3], gridDim.x)) return;

    __shared__ uint64_t s_sort[kSortSmem];
    __shared__ uint32_t s_warp[kPostThreads / 32];
    __shared__ uint32_t s_filled;
    uint32_t count = *((volatile uint32_t *)&w->st.n_cand);
    if (!st.done) {
        // exact ties at the k-th key (all 32 bits fixed, more equal keys than slots):
        // take the k_rem lowest ids in row order.
        if (threadIdx.x == 0) s_filled = 0;
        __syncthreads();
        const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
        for (int64_t base = 0; base < n; base += blockDim.x) {
            const int64_t row = base + threadIdx.x;
            bool hit = row < n && f2key(__ldcg(scores + row)) == st.prefix;
            uint32_t bal = __ballot_sync(0xffffffffu, hit);
            if (lane == 0) s_warp[wid] = __popc(bal);
            __syncthreads();
            uint32_t before = s_filled;
            for (int x = 0; x < wid; x++) before += s_warp[x];
            uint32_t rank = before + __popc(bal & ((1u << lane) - 1));
            if (hit && rank < st.k_rem) cand[count + rank] = make_comp(st.prefix, (uint32_t)row);
            __syncthreads();
            if (threadIdx.x == 0) {
                uint32_t tot = 0;
                for (int x = 0; x < (int)(blockDim.x >> 5); x++) tot += s_warp[x];
                s_filled += tot;
            }
            __syncthreads();
            if (s_filled >= st.k_rem) break;
        }
        count += st.k_rem;
        __threadfence_block();
        __syncthreads();
    }
    uint32_t p2 = 1;
    while (p2 < count) p2 <<= 1;
    float *D = D_all + (size_t)blockIdx.y * k;
    int64_t *I = I_all + (size_t)blockIdx.y * k;
    if (p2 <= (uint32_t)kSortSmem) {
        for (uint32_t i = threadIdx.x; i < p2; i += blockDim.x) s_sort[i] = i < count ? __ldcg(cand + i) : 0ull;
        __syncthreads();
        bitonic_desc(s_sort, p2);
        emit_sorted(s_sort, count, k, id_base, D, I);
    } else {
        for (uint32_t i = count + threadIdx.x; i < p2; i += blockDim.x) cand[i] = 0ull;
        __syncthreads();
        bitonic_desc(cand, p2);   // global-memory sort for very large k (REPL paging)
        emit_sorted(cand, count, k, id_base, D, I);
    }
}

__global__ void fill_empty_kernel(float *D, int64_t *I, int64_t total) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total;
         i += (int64_t)gridDim.x * blockDim.x) {
        D[i] = -3.4028234663852886e38f;
        I[i] = -1;
    }
}

// merge R sorted per-shard lists: one block per query
__global__ void __launch_bounds__(kPostThreads)
topk_merge_kernel(int R, int64_t nq, int64_t k, const float *__restrict__ D_in,
                  const int64_t *__restrict__ I_in, int64_t shard_stride_D, int64_t shard_stride_I,
                  float *D_out, int64_t *I_out, float *g_s, int64_t *g_i, uint32_t p2) {
    // scores/ids are sorted as (key, ~id) pairs; ids are global (up to 2^63), so keep
    // them beside the key instead of packing them
    extern __shared__ unsigned char s_raw[];
    const int64_t q = blockIdx.x;
    const uint32_t tot = (uint32_t)(R * k);
    float *ss;
    int64_t *si;
    if (g_s) { ss = g_s + (size_t)q * p2; si = g_i + (size_t)q * p2; }
    else { si = reinterpret_cast<int64_t *>(s_raw); ss = reinterpret_cast<float *>(si + p2); }
    for (uint32_t i = threadIdx.x; i < p2; i += blockDim.x) {
        if (i < tot) {
            uint32_t r = i / (uint32_t)k, j = i % (uint32_t)k;
            size_t src = (size_t)q * k + j;
            ss[i] = D_in[(size_t)r * shard_stride_D + src];
            si[i] = I_in[(size_t)r * shard_stride_I + src];
        } else { ss[i] = 0.f; si[i] = -1; }
    }
    __syncthreads();
    // a "before" b: valid first, then higher score, then lower id
    auto before = [](float sa, int64_t ia, float sb, int64_t ib) {
        if ((ia < 0) != (ib < 0)) return ib < 0;
        if (ia < 0) return false;
        if (sa != sb) return sa > sb;
        return ia < ib;
    };
    for (uint32_t kk = 2; kk <= p2; kk <<= 1) {
        for (uint32_t j = kk >> 1; j > 0; j >>= 1) {
            for (uint32_t t = threadIdx.x; t < p2 / 2; t += blockDim.x) {
                uint32_t i = ((t & ~(j - 1)) << 1) | (t & (j - 1));
                uint32_t l = i | j;
                bool fwd = (i & kk) == 0;
                float sa = ss[i], sb = ss[l];
                int64_t ia = si[i], ib = si[l];
                bool swap = fwd ? before(sb, ib, sa, ia) : before(sa, ia, sb, ib);
                if (swap) { ss[i] = sb; ss[l] = sa; si[i] = ib; si[l] = ia; }
            }
            __syncthreads();
        }
    }
    for (int64_t j = threadIdx.x; j < k; j += blockDim.x) {
        bool valid = (uint32_t)j < tot && si[j] >= 0;
        D_out[q * k + j] = valid ? ss[j] : -3.4028234663852886e38f;
        I_out[q * k + j] = valid ? si[j] : -1;
    }
}

__global__ void f32_to_f16_kernel(const float4 *__restrict__ src, uint2 *__restrict__ dst, int64_t n4) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4;
         i += (int64_t)gridDim.x * blockDim.x) {
        float4 f = src[i];
        __half2 a = __floats2half2_rn(f.x, f.y), b = __floats2half2_rn(f.z, f.w);
        uint2 o;
        o.x = *reinterpret_cast<uint32_t *>(&a);
        o.y = *reinterpret_cast<uint32_t *>(&b);
        dst[i] = o;
    }
}
__global__ void f16_to_f32_kernel(const __half *__restrict__ src, float *__restrict__ dst, int64_t n) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n;
         i += (int64_t)gridDim.x * blockDim.x)
        dst[i] = __half2float(src[i]);
}

}  // namespace cb

// =====================================================================================
// host side
// =====================================================================================
using namespace cb;

struct cb_index {
    int d = kD;
    int dtype = CB_F16;
    int device = 0;
    int64_t ntotal = 0;
    int64_t capacity = 0;
    void *rows = nullptr;            // [capacity][512] of dtype
    cudaStream_t stream = nullptr;   // used by the host-pointer entry points
    // search workspace
    float *scores = nullptr;         // [kMaxNQ][score_stride]
    int64_t score_stride = 0;
    QueryWs *ws = nullptr;           // [kMaxNQ]
    uint64_t *cand = nullptr;        // [kMaxNQ][cand_cap]
    uint32_t cand_cap = 0;
    // pinned + device staging for the host-pointer entry points
    float *h_q = nullptr, *d_q = nullptr;
    int64_t q_cap = 0;               // queries
    float *h_D = nullptr, *d_D = nullptr;
    int64_t *h_I = nullptr, *d_I = nullptr;
    int64_t out_cap = 0;             // nq*k elements
    void *d_stage = nullptr;         // add() staging
    void *h_stage = nullptr;
    int64_t stage_bytes = 0;
    int scan_blocks_per_sm[2][3] = {{0}};
    cb::BatchWs *bws = nullptr;      // workspace of the tensor-core batch path
    int64_t n_batch_searches = 0, n_batch_overflows = 0;
    // optional live timing of the scan kernel (bench.py roofline): event pairs on the
    // launching stream, resolved lazily by cb_flatip_timing_read
    bool timing = false;
    static constexpr int kEv = 256;
    cudaEvent_t ev0[kEv] = {nullptr}, ev1[kEv] = {nullptr};
    int ev_n = 0;

    size_t row_bytes() const { return (size_t)d * (dtype == CB_F16 ? 2 : 4); }
};

static int grow_rows(cb_index *ix, int64_t need, cudaStream_t s) {
    if (need <= ix->capacity) return CB_OK;
    int64_t cap = std::max<int64_t>(need, ix->capacity + ix->capacity / 2);
    cap = (cap + 63) / 64 * 64;
    void *p = nullptr;
    CB_CUDA(cudaMalloc(&p, (size_t)cap * ix->row_bytes()));
    if (ix->rows && ix->ntotal)
        CB_CUDA(cudaMemcpyAsync(p, ix->rows, (size_t)ix->ntotal * ix->row_bytes(),
                                cudaMemcpyDeviceToDevice, s));
    CB_CUDA(cudaStreamSynchronize(s));
    if (ix->rows) CB_CUDA(cudaFree(ix->rows));
    ix->rows = p;
    ix->capacity = cap;
    return CB_OK;
}

static int ensure_ws(cb_index *ix, int64_t k_eff) {
    const int64_t stride = (ix->ntotal + 63) / 64 * 64;
    if (stride > ix->score_stride) {
        if (ix->scores) CB_CUDA(cudaFree(ix->scores));
        ix->scores = nullptr;
        int64_t s = std::max<int64_t>(stride, (ix->capacity + 63) / 64 * 64);
        CB_CUDA(cudaMalloc(&ix->scores, (size_t)kMaxNQ * s * sizeof(float)));
        ix->score_stride = s;
    }
    if (!ix->ws) CB_CUDA(cudaMalloc(&ix->ws, sizeof(QueryWs) * kMaxNQ));
    uint32_t p2 = 256;
    while ((int64_t)p2 < k_eff) p2 <<= 1;
    if (p2 > ix->cand_cap) {
        if (ix->cand) CB_CUDA(cudaFree(ix->cand));
        ix->cand = nullptr;
        CB_CUDA(cudaMalloc(&ix->cand, (size_t)kMaxNQ * p2 * sizeof(uint64_t)));
        ix->cand_cap = p2;
    }
    return CB_OK;
}

template <int NQ, bool F16>
static int launch_scan(cb_index *ix, const float *q_dev, int nq_valid, uint32_t k_eff, cudaStream_t s) {
    auto kern = flatip_scan_kernel<NQ, F16>;
    const size_t smem = (size_t)NQ * kBins0 * sizeof(uint32_t);
    constexpr int slot = NQ == 1 ? 0 : NQ == 2 ? 1 : 2;
    int &bps = ix->scan_blocks_per_sm[F16 ? 1 : 0][slot];
    if (bps == 0) {
        CB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        CB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&bps, kern, kScanThreads, smem));
        if (bps < 1) bps = 1;
    }
    int sms = kNumSMs;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, ix->device);
    constexpr int R = F16 ? 8 : 4;
    int64_t groups = (ix->ntotal + R - 1) / R;
    int64_t want = (groups + (kScanThreads / 32) - 1) / (kScanThreads / 32);
    int grid = (int)std::min<int64_t>((int64_t)sms * bps, std::max<int64_t>(want, 1));
    const bool timed = ix->timing && ix->ev_n < cb_index::kEv;
    if (timed) {
        if (!ix->ev0[ix->ev_n]) {
            CB_CUDA(cudaEventCreate(&ix->ev0[ix->ev_n]));
            CB_CUDA(cudaEventCreate(&ix->ev1[ix->ev_n]));
        }
        CB_CUDA(cudaEventRecord(ix->ev0[ix->ev_n], s));
    }
    kern<<<grid, kScanThreads, smem, s>>>((const uint4 *)ix->rows, ix->ntotal, q_dev, nq_valid,
                                           ix->scores, ix->score_stride, ix->ws, k_eff);
    CB_LAUNCH_CHECK();
    if (timed) {
        CB_CUDA(cudaEventRecord(ix->ev1[ix->ev_n], s));
        ix->ev_n++;
    }
    return CB_OK;
}

static int search_tile(cb_index *ix, int nq, const float *q_dev, int64_t k, float *D_dev,
                       int64_t *I_dev, int64_t id_base, cudaStream_t s) {
    const int64_t n = ix->ntotal;
    const int64_t k_eff = std::min<int64_t>(k, n);
    // reset histograms + state (k_rem is seeded by the scan kernel's last block)
    CB_CUDA(cudaMemsetAsync(ix->ws, 0, sizeof(QueryWs) * nq, s));
    const bool f16 = ix->dtype == CB_F16;
    int rc;
    if (nq == 1) rc = f16 ? launch_scan<1, true>(ix, q_dev, nq, (uint32_t)k_eff, s) : launch_scan<1, false>(ix, q_dev, nq, (uint32_t)k_eff, s);
    else if (nq == 2) rc = f16 ? launch_scan<2, true>(ix, q_dev, nq, (uint32_t)k_eff, s) : launch_scan<2, false>(ix, q_dev, nq, (uint32_t)k_eff, s);
    else rc = f16 ? launch_scan<4, true>(ix, q_dev, nq, (uint32_t)k_eff, s) : launch_scan<4, false>(ix, q_dev, nq, (uint32_t)k_eff, s);
    if (rc) return rc;

    int sms = kNumSMs;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, ix->device);
    const int64_t n4 = (n + 3) / 4;
    int gx = (int)std::min<int64_t>((int64_t)sms * 4, std::max<int64_t>((n4 + kPostThreads - 1) / kPostThreads, 1));
    dim3 grid(gx, nq);
    for (int pass = 1; pass <= 2; pass++) {
        flatip_refine_kernel<<<grid, kPostThreads, 0, s>>>(ix->scores, n, ix->score_stride, ix->ws, pass);
        CB_LAUNCH_CHECK();
    }
    flatip_collect_kernel<<<grid, kPostThreads, 0, s>>>(ix->scores, n, ix->score_stride, ix->ws, ix->cand,
                                                        ix->cand_cap, k, id_base, D_dev, I_dev);
    CB_LAUNCH_CHECK();
    return CB_OK;
}

extern "C" {

int cb_flatip_create(int d, int storage_dtype, int device, cb_index **out) {
    CB_REQUIRE(out != nullptr, "cb_flatip_create: out is null");
    *out = nullptr;
    CB_REQUIRE(d == kD, "cb_flatip_create: d must be %d (got %d)", kD, d);
    CB_REQUIRE(storage_dtype == CB_F16 || storage_dtype == CB_F32,
               "cb_flatip_create: storage dtype must be CB_F32 or CB_F16");
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
        cudaGetLastError();
        set_error("cb_flatip_create: no CUDA device (this library has no CPU fallback)");
        return CB_ERR_NOGPU;
    }
    CB_REQUIRE(device >= 0 && device < ndev, "cb_flatip_create: device %d out of range (%d devices)", device, ndev);
    DeviceGuard g(device);
    cb_index *ix = new (std::nothrow) cb_index();
    if (!ix) { set_error("out of host memory"); return CB_ERR_OOM; }
    ix->d = d;
    ix->dtype = storage_dtype;
    ix->device = device;
    cudaError_t e = cudaStreamCreateWithFlags(&ix->stream, cudaStreamNonBlocking);
    if (e != cudaSuccess) {
        set_error("cudaStreamCreate failed: %s", cudaGetErrorString(e));
        delete ix;
        return CB_ERR_CUDA;
    }
    *out = ix;
    return CB_OK;
}

void cb_flatip_free(cb_index *ix) {
    if (!ix) return;
    DeviceGuard g(ix->device);
    if (ix->stream) cudaStreamSynchronize(ix->stream);
    cudaFree(ix->rows); cudaFree(ix->scores); cudaFree(ix->ws); cudaFree(ix->cand);
    cudaFree(ix->d_q); cudaFree(ix->d_D); cudaFree(ix->d_I); cudaFree(ix->d_stage);
    cudaFreeHost(ix->h_q); cudaFreeHost(ix->h_D); cudaFreeHost(ix->h_I); cudaFreeHost(ix->h_stage);
    cb::batch_ws_delete(ix->bws);
    for (int i = 0; i < cb_index::kEv; i++) {
        if (ix->ev0[i]) cudaEventDestroy(ix->ev0[i]);
        if (ix->ev1[i]) cudaEventDestroy(ix->ev1[i]);
    }
    if (ix->stream) cudaStreamDestroy(ix->stream);
    delete ix;
}

int64_t cb_flatip_ntotal(const cb_index *ix) { return ix ? ix->ntotal : -1; }
int cb_flatip_dim(const cb_index *ix) { return ix ? ix->d : -1; }
int cb_flatip_storage_dtype(const cb_index *ix) { return ix ? ix->dtype : -1; }
const void *cb_flatip_device_rows(const cb_index *ix) { return ix ? ix->rows : nullptr; }

int cb_flatip_reserve(cb_index *ix, int64_t n_rows) {
    CB_REQUIRE(ix != nullptr, "cb_flatip_reserve: null index");
    CB_REQUIRE(n_rows >= 0 && n_rows < (1ll << 32), "cb_flatip_reserve: a shard holds < 2^32 rows");
    DeviceGuard g(ix->device);
    return grow_rows(ix, n_rows, ix->stream);
}

int cb_flatip_reset(cb_index *ix) {
    CB_REQUIRE(ix != nullptr, "cb_flatip_reset: null index");
    ix->ntotal = 0;
    return CB_OK;
}

int cb_flatip_add_device(cb_index *ix, int64_t n, const void *x_dev, int src_dtype, void *stream) {
    CB_REQUIRE(ix != nullptr, "cb_flatip_add_device: null index");
    CB_REQUIRE(n >= 0, "cb_flatip_add_device: n < 0");
    CB_REQUIRE(src_dtype == CB_F16 || src_dtype == CB_F32, "cb_flatip_add_device: bad src dtype");
    if (n == 0) return CB_OK;
    CB_REQUIRE(x_dev != nullptr, "cb_flatip_add_device: null rows");
    CB_REQUIRE(ix->ntotal + n < (1ll << 32), "cb_flatip_add_device: a shard holds < 2^32 rows");
    DeviceGuard g(ix->device);
    cudaStream_t s = (cudaStream_t)stream;
    int rc = grow_rows(ix, ix->ntotal + n, s);
    if (rc) return rc;
    char *dst = (char *)ix->rows + (size_t)ix->ntotal * ix->row_bytes();
    const int64_t elems = n * ix->d;
    if (src_dtype == ix->dtype) {
        CB_CUDA(cudaMemcpyAsync(dst, x_dev, (size_t)n * ix->row_bytes(), cudaMemcpyDeviceToDevice, s));
    } else if (src_dtype == CB_F32) {
        int64_t n4 = elems / 4;
        int grid = (int)std::min<int64_t>(kNumSMs * 8, (n4 + 255) / 256);
        f32_to_f16_kernel<<<grid, 256, 0, s>>>((const float4 *)x_dev, (uint2 *)dst, n4);
        CB_LAUNCH_CHECK();
    } else {
        int grid = (int)std::min<int64_t>(kNumSMs * 8, (elems + 255) / 256);
        f16_to_f32_kernel<<<grid, 256, 0, s>>>((const __half *)x_dev, (float *)dst, elems);
        CB_LAUNCH_CHECK();
    }
    ix->ntotal += n;
    return CB_OK;
}

int cb_flatip_add(cb_index *ix, int64_t n, const float *x_host) {
    CB_REQUIRE(ix != nullptr, "cb_flatip_add: null index");
    CB_REQUIRE(n >= 0, "cb_flatip_add: n < 0");
    if (n == 0) return CB_OK;
    CB_REQUIRE(x_host != nullptr, "cb_flatip_add: null rows");
    DeviceGuard g(ix->device);
    int rc = grow_rows(ix, ix->ntotal + n, ix->stream);
    if (rc) return rc;
    // stream the rows through a bounded staging buffer (64 MiB of fp32)
    const int64_t chunk_rows = (64ll << 20) / (ix->d * 4);
    if (!ix->d_stage) {
        ix->stage_bytes = chunk_rows * ix->d * 4;
        CB_CUDA(cudaMalloc(&ix->d_stage, ix->stage_bytes));
    }
    for (int64_t lo = 0; lo < n; lo += chunk_rows) {
        int64_t m = std::min(chunk_rows, n - lo);
        CB_CUDA(cudaMemcpyAsync(ix->d_stage, x_host + lo * ix->d, (size_t)m * ix->d * 4,
                                cudaMemcpyHostToDevice, ix->stream));
        rc = cb_flatip_add_device(ix, m, ix->d_stage, CB_F32, ix->stream);
        if (rc) return rc;
        CB_CUDA(cudaStreamSynchronize(ix->stream));
    }
    return CB_OK;
}

int cb_flatip_search_device(cb_index *ix, int64_t nq, const float *q_dev, int64_t k, float *D_dev,
                            int64_t *I_dev, int64_t id_base, void *stream) {
    CB_REQUIRE(ix != nullptr, "cb_flatip_search_device: null index");
    CB_REQUIRE(nq >= 0, "cb_flatip_search_device: nq < 0");
    CB_REQUIRE(k > 0, "cb_flatip_search_device: k must be > 0 (got %lld)", (long long)k);
    CB_REQUIRE(k < (1ll << 31), "cb_flatip_search_device: k too large");
    if (nq == 0) return CB_OK;
    CB_REQUIRE(q_dev && D_dev && I_dev, "cb_flatip_search_device: null buffer");
    DeviceGuard g(ix->device);
    cudaStream_t s = (cudaStream_t)stream;
    if (ix->ntotal == 0) {
        int64_t tot = nq * k;
        fill_empty_kernel<<<(int)std::min<int64_t>(1024, (tot + 255) / 256), 256, 0, s>>>(D_dev, I_dev, tot);
        CB_LAUNCH_CHECK();
        return CB_OK;
    }
    // query batches: tensor-core GEMM + fused threshold filter (fp16 shards, k <= 1024).
    // The GEMM pass costs ~3.3 ms over 10M rows whatever nq <= 256 is (one 256-query tile column); the
    // streaming scan costs 1.5 ms for one query, 2.9 ms for four and another pass per four after that,
    // so the crossover is at five queries (profiles/r01_batch_crossover.txt).  Small shards keep the old
    // threshold: there both paths are bounded by their launch counts, not by the pass over the rows.
    int64_t batch_min = ix->ntotal >= (1ll << 20) ? 5 : 16;
    if (const char *e = getenv("CLIPB200_BATCH_MIN_NQ")) batch_min = atoll(e);
    if (nq >= batch_min && ix->dtype == CB_F16 && k <= 1024 && ix->ntotal >= 8192) {
        if (!ix->bws) ix->bws = batch_ws_new();
        bool overflowed = false;
        int brc = flatip_search_batch(ix->bws, ix->rows, ix->ntotal, ix->device, nq, q_dev, k, D_dev, I_dev,
                                      id_base, s, &overflowed);
        if (brc) return brc;
        ix->n_batch_searches++;
        if (!overflowed) return CB_OK;
        ix->n_batch_overflows++;      // adversarial row order: redo exactly with the scan path
    }
    int rc = ensure_ws(ix, std::min<int64_t>(k, ix->ntotal));
    if (rc) return rc;
    for (int64_t q0 = 0; q0 < nq; q0 += kMaxNQ) {
        int t = (int)std::min<int64_t>(kMaxNQ, nq - q0);
        rc = search_tile(ix, t, q_dev + q0 * ix->d, k, D_dev + q0 * k, I_dev + q0 * k, id_base, s);
        if (rc) return rc;
    }
    return CB_OK;
}

int cb_flatip_search(cb_index *ix, int64_t nq, const float *q_host, int64_t k, float *D_host,
                     int64_t *I_host) {
    CB_REQUIRE(ix != nullptr, "cb_flatip_search: null index");
    CB_REQUIRE(nq >= 0, "cb_flatip_search: nq < 0");
    CB_REQUIRE(k > 0, "cb_flatip_search: k must be > 0 (got %lld)", (long long)k);
    if (nq == 0) return CB_OK;
    CB_REQUIRE(q_host && D_host && I_host, "cb_flatip_search: null buffer");
    DeviceGuard g(ix->device);
    if (nq > ix->q_cap) {
        cudaFreeHost(ix->h_q); cudaFree(ix->d_q);
        ix->h_q = nullptr; ix->d_q = nullptr;
        CB_CUDA(cudaMallocHost(&ix->h_q, (size_t)nq * ix->d * 4));
        CB_CUDA(cudaMalloc(&ix->d_q, (size_t)nq * ix->d * 4));
        ix->q_cap = nq;
    }
    if (nq * k > ix->out_cap) {
        cudaFreeHost(ix->h_D); cudaFreeHost(ix->h_I); cudaFree(ix->d_D); cudaFree(ix->d_I);
        ix->h_D = nullptr; ix->h_I = nullptr; ix->d_D = nullptr; ix->d_I = nullptr;
        CB_CUDA(cudaMallocHost(&ix->h_D, (size_t)nq * k * 4));
        CB_CUDA(cudaMallocHost(&ix->h_I, (size_t)nq * k * 8));
        CB_CUDA(cudaMalloc(&ix->d_D, (size_t)nq * k * 4));
        CB_CUDA(cudaMalloc(&ix->d_I, (size_t)nq * k * 8));
        ix->out_cap = nq * k;
    }
    memcpy(ix->h_q, q_host, (size_t)nq * ix->d * 4);
    CB_CUDA(cudaMemcpyAsync(ix->d_q, ix->h_q, (size_t)nq * ix->d * 4, cudaMemcpyHostToDevice, ix->stream));
    int rc = cb_flatip_search_device(ix, nq, ix->d_q, k, ix->d_D, ix->d_I, 0, ix->stream);
    if (rc) return rc;
    CB_CUDA(cudaMemcpyAsync(ix->h_D, ix->d_D, (size_t)nq * k * 4, cudaMemcpyDeviceToHost, ix->stream));
    CB_CUDA(cudaMemcpyAsync(ix->h_I, ix->d_I, (size_t)nq * k * 8, cudaMemcpyDeviceToHost, ix->stream));
    CB_CUDA(cudaStreamSynchronize(ix->stream));
    memcpy(D_host, ix->h_D, (size_t)nq * k * 4);
    memcpy(I_host, ix->h_I, (size_t)nq * k * 8);
    return CB_OK;
}

int cb_topk_merge_device(int R, int64_t nq, int64_t k, const float *D_in, const int64_t *I_in,
                         int64_t shard_stride_D, int64_t shard_stride_I, float *D_out,
                         int64_t *I_out, void *stream) {
    CB_REQUIRE(R > 0 && nq >= 0 && k > 0, "cb_topk_merge_device: bad shape");
    if (nq == 0) return CB_OK;
    CB_REQUIRE(D_in && I_in && D_out && I_out, "cb_topk_merge_device: null buffer");
    CB_REQUIRE((int64_t)R * k < (1ll << 30), "cb_topk_merge_device: R*k too large");
    if (shard_stride_D <= 0) shard_stride_D = nq * k;
    if (shard_stride_I <= 0) shard_stride_I = nq * k;
    cudaStream_t s = (cudaStream_t)stream;
    uint32_t p2 = 2;
    while ((int64_t)p2 < (int64_t)R * k) p2 <<= 1;
    size_t smem = (size_t)p2 * 12;
    float *g_s = nullptr;
    int64_t *g_i = nullptr;
    if (smem > 96 * 1024) {
        // very large k: sort in global scratch (allocated per call; rare REPL paging case)
        CB_CUDA(cudaMallocAsync(&g_i, (size_t)nq * p2 * 8, s));
        CB_CUDA(cudaMallocAsync(&g_s, (size_t)nq * p2 * 4, s));
        smem = 0;
    } else if (smem > 48 * 1024) {
        CB_CUDA(cudaFuncSetAttribute(topk_merge_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024));
    }
    topk_merge_kernel<<<(unsigned)nq, kPostThreads, smem, s>>>(R, nq, k, D_in, I_in, shard_stride_D, shard_stride_I,
                                                               D_out, I_out, g_s, g_i, p2);
    CB_LAUNCH_CHECK();
    if (g_s) { CB_CUDA(cudaFreeAsync(g_s, s)); CB_CUDA(cudaFreeAsync(g_i, s)); }
    return CB_OK;
}

int cb_flatip_batch_stats(cb_index *ix, int64_t *n_batch_searches, int64_t *n_overflow_fallbacks) {
    CB_REQUIRE(ix && n_batch_searches && n_overflow_fallbacks, "cb_flatip_batch_stats: null argument");
    *n_batch_searches = ix->n_batch_searches;
    *n_overflow_fallbacks = ix->n_batch_overflows;
    return CB_OK;
}

int cb_flatip_timing(cb_index *ix, int enable) {
    CB_REQUIRE(ix != nullptr, "cb_flatip_timing: null index");
    ix->timing = enable != 0;
    ix->ev_n = 0;
    return CB_OK;
}

int cb_flatip_timing_read(cb_index *ix, double *scan_ms_total, int *n_scans) {
    CB_REQUIRE(ix && scan_ms_total && n_scans, "cb_flatip_timing_read: null argument");
    DeviceGuard g(ix->device);
    double tot = 0;
    for (int i = 0; i < ix->ev_n; i++) {
        CB_CUDA(cudaEventSynchronize(ix->ev1[i]));
        float ms = 0;
        CB_CUDA(cudaEventElapsedTime(&ms, ix->ev0[i], ix->ev1[i]));
        tot += ms;
    }
    *scan_ms_total = tot;
    *n_scans = ix->ev_n;
    ix->ev_n = 0;
    return CB_OK;
}

int cb_flatip_get_rows(cb_index *ix, int64_t start, int64_t n, float *out_host) {
    CB_REQUIRE(ix != nullptr, "cb_flatip_get_rows: null index");
    CB_REQUIRE(start >= 0 && n >= 0 && start + n <= ix->ntotal, "cb_flatip_get_rows: range out of bounds");
    if (n == 0) return CB_OK;
    CB_REQUIRE(out_host != nullptr, "cb_flatip_get_rows: null buffer");
    DeviceGuard g(ix->device);
    const char *src = (const char *)ix->rows + (size_t)start * ix->row_bytes();
    if (ix->dtype == CB_F32) {
        CB_CUDA(cudaMemcpyAsync(out_host, src, (size_t)n * ix->d * 4, cudaMemcpyDeviceToHost, ix->stream));
        CB_CUDA(cudaStreamSynchronize(ix->stream));
        return CB_OK;
    }
    const int64_t chunk_rows = (64ll << 20) / (ix->d * 4);
    if (!ix->d_stage) {
        ix->stage_bytes = chunk_rows * ix->d * 4;
        CB_CUDA(cudaMalloc(&ix->d_stage, ix->stage_bytes));
    }
    for (int64_t lo = 0; lo < n; lo += chunk_rows) {
        int64_t m = std::min(chunk_rows, n - lo);
        int64_t elems = m * ix->d;
        int grid = (int)std::min<int64_t>(kNumSMs * 8, (elems + 255) / 256);
        f16_to_f32_kernel<<<grid, 256, 0, ix->stream>>>((const __half *)(src + (size_t)lo * ix->row_bytes()),
                                                        (float *)ix->d_stage, elems);
        CB_LAUNCH_CHECK();
        CB_CUDA(cudaMemcpyAsync(out_host + lo * ix->d, ix->d_stage, (size_t)elems * 4, cudaMemcpyDeviceToHost, ix->stream));
        CB_CUDA(cudaStreamSynchronize(ix->stream));
    }
    return CB_OK;
}

}  // extern "C"
