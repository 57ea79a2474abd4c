// Exact inner-product top-k over an N x 512 shard resident in HBM.
//
// Replaces faiss.IndexFlatIP.search as called at /root/reference/query-index.py:111
// (index built at /root/reference/build-index.py:80-81,99,107).
//
// Small-nq path (this file): HBM-bound.  ONE cooperative launch per search
// (flatip_search_kernel): a pass over the shard with coalesced 128-bit streaming loads,
// fp32 FMA on CUDA cores (free when HBM-bound), a transposed warp-shuffle reduction,
// scores written once (4 B/row = 0.4 % of the 1024 B/row read) with a fused histogram
// over 2048 linear score bins; a grid barrier; then the bin holding the k-th score is
// known, the few rows at or above it are gathered from the (L2-resident) scores, and the
// last block to retire sorts them, orders ties by id and writes D / I.  Degenerate score
// distributions (ties, constant rows) and very large k fall back, inside the same launch,
// to an exact radix select over the 32 key bits.  No host round trip, any k from 1 to
// ntotal.  Order is the total order (-score, id): results do not depend on the launch
// geometry, so 1-GPU and sharded results are identical.
#include "common.cuh"
#include "flatip.cuh"

#include <algorithm>
#include <cstdarg>
#include <cstring>
#include <new>
#include <vector>

namespace cb {

// tensor-core batch path (flatip_batch.cu)
struct BatchWs;
BatchWs *batch_ws_new();
void batch_ws_delete(BatchWs *w);
int flatip_search_batch(BatchWs *w, const void *rows_f16, int64_t n, int device, int64_t nq, const float *q_dev,
                        int64_t k, float *D_dev, int64_t *I_dev, const IdMap &ids, const PeerOut &po,
                        const float *max_norm2, cudaStream_t s);
int batch_ws_stats(BatchWs *w, int64_t *rescued, cudaStream_t s);

constexpr int kBins0 = 2048;           // bins per histogram level
constexpr int kMaxNQ = 4;              // queries sharing one pass over the shard (register budget)
constexpr int kSortSmem = 4096;        // composite keys sorted in shared memory
constexpr int kScanThreads = 256;
constexpr int kPostThreads = 256;

struct SelState {
    uint32_t prefix;     // radix levels: selected key prefix, right aligned, `bits` wide
    uint32_t bits;       // prefix bits fixed so far: 0, 11, 22, 32
    uint32_t k_rem;      // winners still to take from inside the selected bin
    uint32_t cnt_bin;    // elements inside the selected bin
    uint32_t done;       // 1: cnt_bin == k_rem -> whole bin wins, stop refining
    uint32_t n_cand;     // collect cursor
    uint32_t ticket;     // block-retire counter of the collect phase
    uint32_t bar;        // grid barrier counter (ws[0] only)
    uint32_t pad[8];
};
static_assert(sizeof(SelState) == 64, "SelState layout");

// workspace per query: level 0 = linear score bins, levels 1..3 = radix digits (11 + 11 + 10 bits) of
// the float key inside the level-0 bin; + state.  All zero between searches.
struct QueryWs {
    uint32_t hist[4][kBins0];
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
    __syncwarp();
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

// the same over a histogram already in shared memory (no dependent L2 round trips)
__device__ void find_bin_smem(const uint32_t *hist, int nbins, int digit_bits, SelState *st) {
    const int lane = threadIdx.x & 31;
    const int per = nbins / 32;
    const int hi = nbins - 1 - lane * per;
    uint32_t s = 0;
    for (int i = 0; i < per; i++) s += hist[hi - i];
    uint32_t incl = s;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        uint32_t t = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += t;
    }
    const uint32_t excl = incl - s;
    const uint32_t k_rem = st->k_rem;
    __syncwarp();
    if (excl < k_rem && incl >= k_rem) {
        uint32_t acc = excl;
        for (int i = 0; i < per; i++) {
            uint32_t c = hist[hi - i];
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

// Grid-wide barrier for a kernel launched COOPERATIVELY (every block resident).  `ctr` counts
// arrivals since the workspace was last zeroed; `epoch` = barriers passed so far in this launch.
__device__ __forceinline__ void grid_barrier(uint32_t *ctr, uint32_t &epoch) {
    __syncthreads();
    epoch++;
    if (threadIdx.x == 0) {
        __threadfence();
        atomicAdd(ctr, 1u);
        const uint32_t want = epoch * gridDim.x;
        uint32_t v;
        do {
            asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(ctr) : "memory");
        } while (v < want);
    }
    __syncthreads();
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

// selection state as other blocks left it (written between grid barriers): read through L2
__device__ __forceinline__ SelState load_state(const SelState *p) {
    SelState st;
    const uint4 *src = reinterpret_cast<const uint4 *>(p);
    uint4 *dst = reinterpret_cast<uint4 *>(&st);
#pragma unroll
    for (int i = 0; i < 4; i++) dst[i] = __ldcg(src + i);
    return st;
}

// level-0 bin of a score: linear over [-B, B] with B >= |q| max|x| (so every score is in range),
// 2048 bins.  Monotone non-decreasing in s, hence a higher bin always means a higher score.
__device__ __forceinline__ int lin_bin(float s, float scale) {
    const int b = (int)floorf(fmaf(s, scale, 1024.f));
    return min(kBins0 - 1, max(0, b));
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

// write one query's sorted winners + faiss padding
__device__ void emit_sorted(const uint64_t *a, uint32_t count, int64_t k, const IdMap &ids,
                            float *D, int64_t *I) {
    for (int64_t j = threadIdx.x; j < k; j += blockDim.x) {
        if (j < count) {
            uint64_t c = a[j];
            D[j] = key2f((uint32_t)(c >> 32));
            I[j] = map_id(ids, 0xffffffffu - (uint32_t)c);
        } else {
            D[j] = -3.4028234663852886e38f;
            I[j] = -1;
        }
    }
}

// The whole small-nq search in ONE cooperative launch.
//
// Phase 1 (HBM-bound, >= 96 % of the time): one pass over the shard.  Row = 512 elements: fp16 ->
//   64 x 16 B chunks (lane takes chunks lane, lane+32), fp32 -> 128 chunks (lane, +32, +64, +96);
//   a lane owns 16 elements of every row and keeps the matching 16 query values per query in
//   registers.  R rows in flight per warp iteration (16 x 128-bit loads per lane), fp32 FMA, the
//   transposed shuffle reduction, scores written once (4 B/row), and a histogram of the scores over
//   2048 LINEAR bins spanning [-|q| max|x|, +|q| max|x|].
// -- grid barrier --
// Phase 2: every block finds the bin b0 that holds the k-th best score.  With 2048 linear bins the
//   rows in bins >= b0 are usually only a little more than k ("short list"): phase 3 takes them all.
//   Otherwise (k too large for shared memory, or scores piled into one bin: ties, constant rows) an
//   exact radix select over the 32 key bits runs inside bin b0 (3 more histogram levels, two grid
//   barriers each) -- same result, more passes over the 4 B/row scores.
// Phase 3: every block gathers its share of the winners as 64-bit (key, ~id) composites; the last
//   block to retire orders exact ties by id, sorts, writes D / I (possibly straight into the root
//   GPU's mailbox, see PeerOut) and zeroes the workspace for the next search.
// Order is the total order (-score, id): the result does not depend on the launch geometry, so
// 1-GPU and sharded answers are identical.
template <int NQ, bool F16>
__global__ void __launch_bounds__(kScanThreads)
flatip_search_kernel(const uint4 *__restrict__ xb, int64_t n, const float *__restrict__ xq,
                     int nq_valid, float *__restrict__ scores, int64_t stride, QueryWs *ws,
                     uint32_t k_eff, int64_t k, uint64_t *cand_all, uint32_t cand_cap,
                     const float *__restrict__ max_norm2, const IdMap ids, const PeerOut po,
                     float *D_all, int64_t *I_all, unsigned long long *phase_t) {
    constexpr int R = F16 ? 8 : 4;          // rows per warp iteration
    // optional phase stamps (cb_flatip_timing): block 0 at entry / after the scan's barrier / before the
    // gather; the last block of the last query at exit
    auto stamp = [&](int i) {
        if (phase_t != nullptr && threadIdx.x == 0) {
            unsigned long long t;
            asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
            phase_t[i] = t;
        }
    };
    if (blockIdx.x == 0) stamp(0);
    constexpr int CH = F16 ? 2 : 4;         // chunks per lane per row
    constexpr int EPC = F16 ? 8 : 4;        // elements per chunk
    constexpr int ROW_V4 = F16 ? 64 : 128;  // uint4 per row
    extern __shared__ __align__(16) uint32_t s_dyn[];   // phase 1: [NQ][kBins0] histograms; later: scratch / sort
    __shared__ SelState s_l0[NQ];           // this block's copy of the level-0 decision per query
    __shared__ uint32_t s_warp[kScanThreads / 32];
    __shared__ uint32_t s_filled;
    __shared__ float s_scale[NQ];
    __shared__ uint32_t s_slot_free;
    uint32_t *s_hist = s_dyn;

    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    for (int i = threadIdx.x; i < NQ * kBins0; i += blockDim.x) s_hist[i] = 0;
    // any block may turn out to be the one that writes the result: probe the mailbox slot now
    if (threadIdx.x == 32) s_slot_free = peer_slot_probe(po) ? 1u : 0u;

    float qreg[NQ][CH * EPC];
    float qscale[NQ];
    {
        const float mx = sqrtf(__ldg(max_norm2));
#pragma unroll
        for (int q = 0; q < NQ; q++) {
            const float *qp = xq + (size_t)min(q, nq_valid - 1) * kD;
            float n2 = 0.f;
#pragma unroll
            for (int c = 0; c < CH; c++)
#pragma unroll
                for (int e = 0; e < EPC; e++) {
                    const float v = qp[(lane + 32 * c) * EPC + e];
                    qreg[q][c * EPC + e] = v;
                    n2 = fmaf(v, v, n2);
                }
            n2 = warp_sum(n2);
            const float bound = sqrtf(n2) * mx * 1.001f + 1e-30f;     // |<q, x>| <= |q| |x| (+ rounding slack)
            qscale[q] = 1024.f / bound;                                // inf / 0 for degenerate input: one bin, radix path
            if (threadIdx.x == 0) s_scale[q] = qscale[q];
        }
    }
    __syncthreads();

    const int64_t warps = (int64_t)gridDim.x * (blockDim.x >> 5);
    const int64_t gw = (int64_t)blockIdx.x * (blockDim.x >> 5) + wid;
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
                atomicAdd(&s_hist[q * kBins0 + lin_bin(tot, qscale[q])], 1u);
            }
        }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < NQ * kBins0; i += blockDim.x) {
        uint32_t c = s_hist[i];
        int q = i / kBins0;
        if (c && q < nq_valid) atomicAdd(&ws[q].hist[0][i - q * kBins0], c);
    }
    uint32_t epoch = 0;
    uint32_t *bar = &ws[0].st.bar;
    grid_barrier(bar, epoch);
    if (blockIdx.x == 0) stamp(1);

    // ---- phase 2: level-0 decision, computed by every block for itself: the 2048 counters per query
    // come through L2 in one coalesced round trip into shared memory, then one warp per query scans them
    for (int i = threadIdx.x; i < nq_valid * kBins0; i += blockDim.x) {
        const int q = i / kBins0;
        s_hist[i] = __ldcg(&ws[q].hist[0][i - q * kBins0]);
    }
    __syncthreads();
    if (wid < nq_valid && wid < NQ) {
        SelState *l0 = &s_l0[wid];
        if (lane == 0) { l0->prefix = 0; l0->bits = 0; l0->k_rem = k_eff; l0->cnt_bin = 0; l0->done = 0; }
        __syncwarp();
        find_bin_smem(s_hist + wid * kBins0, kBins0, 11, l0);
    }
    __syncthreads();
    const uint32_t short_cap = min(cand_cap, (uint32_t)kSortSmem);
    bool need_radix = false;                 // uniform over the grid: derived from the global histograms
    for (int q = 0; q < nq_valid; q++)
        need_radix |= (k_eff - s_l0[q].k_rem + s_l0[q].cnt_bin) > short_cap;

    const int64_t n4 = (n + 3) / 4;
    if (need_radix) {
        // exact radix select over the key bits of the rows inside bin b0 (per query that needs it)
        for (int pass = 1; pass <= 3; pass++) {
            const int digit_bits = pass == 3 ? 10 : 11;
            bool any = false;
            for (int q = 0; q < nq_valid; q++) {
                const SelState l0 = s_l0[q];
                if ((k_eff - l0.k_rem + l0.cnt_bin) <= short_cap) continue;
                const SelState st = load_state(&ws[q].st);         // radix state (block 0 writes it between barriers)
                if (pass > 1 && st.done) continue;
                any = true;
                const int shift_prev = 32 - (int)st.bits;          // 32, 21, 10
                const int shift = shift_prev - digit_bits;
                const uint32_t mask = (1u << digit_bits) - 1;
                for (int i = threadIdx.x; i < kBins0; i += blockDim.x) s_dyn[i] = 0;
                __syncthreads();
                const float4 *s4 = reinterpret_cast<const float4 *>(scores + (size_t)q * stride);
                for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (int64_t)gridDim.x * blockDim.x) {
                    float4 f = __ldcg(s4 + i);
                    const float e[4] = {f.x, f.y, f.z, f.w};
#pragma unroll
                    for (int j = 0; j < 4; j++) {
                        if (i * 4 + j < n && lin_bin(e[j], s_scale[q]) == (int)l0.prefix) {
                            const uint32_t key = f2key(e[j]);
                            if (pass == 1 || (key >> shift_prev) == st.prefix) atomicAdd(&s_dyn[(key >> shift) & mask], 1u);
                        }
                    }
                }
                __syncthreads();
                for (int i = threadIdx.x; i < (1 << digit_bits); i += blockDim.x)
                    if (s_dyn[i]) atomicAdd(&ws[q].hist[pass][i], s_dyn[i]);
                __syncthreads();
            }
            if (!any) break;                                       // uniform
            grid_barrier(bar, epoch);
            if (blockIdx.x == 0 && wid < nq_valid && wid < NQ) {
                const SelState l0 = s_l0[wid];
                if ((k_eff - l0.k_rem + l0.cnt_bin) > short_cap && !(pass > 1 && load_state(&ws[wid].st).done)) {
                    if (pass == 1 && lane == 0) ws[wid].st.k_rem = l0.k_rem;
                    __syncwarp();
                    __threadfence();
                    find_bin(ws[wid].hist[pass], 1 << digit_bits, digit_bits, &ws[wid].st);
                }
            }
            grid_barrier(bar, epoch);
        }
    }

    // ---- phase 3: gather the winners; the last block per query sorts and writes
    if (blockIdx.x == 0) stamp(2);
    uint64_t *s_sort = reinterpret_cast<uint64_t *>(s_dyn);
    for (int q = 0; q < nq_valid; q++) {
        QueryWs *w = ws + q;
        const SelState l0 = s_l0[q];
        const bool is_short = (k_eff - l0.k_rem + l0.cnt_bin) <= short_cap;
        const SelState st = load_state(&w->st);
        const float *sc = scores + (size_t)q * stride;
        uint64_t *cand = cand_all + (size_t)q * cand_cap;
        const int sh = 32 - (int)st.bits;   // 0 when all 32 bits are fixed
        const float4 *s4 = reinterpret_cast<const float4 *>(sc);
        for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (int64_t)gridDim.x * blockDim.x) {
            float4 f = __ldcg(s4 + i);
            const float e[4] = {f.x, f.y, f.z, f.w};
#pragma unroll
            for (int j = 0; j < 4; j++) {
                const int64_t row = i * 4 + j;
                if (row >= n) continue;
                const int b = lin_bin(e[j], s_scale[q]);
                bool take = b > (int)l0.prefix;
                uint32_t key = 0;
                if (b >= (int)l0.prefix) key = f2key(e[j]);
                if (b == (int)l0.prefix) {
                    if (is_short) take = true;
                    else {
                        const uint32_t kp = sh ? (sh < 32 ? (key >> sh) : 0u) : key;
                        take = kp > st.prefix || (kp == st.prefix && st.done);
                    }
                }
                if (take) {
                    uint32_t pos = atomicAdd(&w->st.n_cand, 1u);
                    if (pos < cand_cap) cand[pos] = make_comp(key, (uint32_t)row);
                }
            }
        }
        if (!last_block(&w->st.ticket, gridDim.x)) continue;

        uint32_t count = *((volatile uint32_t *)&w->st.n_cand);
        if (!is_short && !st.done) {
            // exact ties at the k-th key (all 32 bits fixed, more equal keys than slots):
            // take the k_rem lowest ids in row order.
            if (threadIdx.x == 0) s_filled = 0;
            __syncthreads();
            for (int64_t base = 0; base < n; base += blockDim.x) {
                const int64_t row = base + threadIdx.x;
                bool hit = row < n && f2key(__ldcg(sc + row)) == st.prefix;
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
        float *D = D_all + (size_t)q * k;
        int64_t *I = I_all + (size_t)q * k;
        if (p2 <= (uint32_t)kSortSmem) {
            for (uint32_t i = threadIdx.x; i < p2; i += blockDim.x) s_sort[i] = i < count ? __ldcg(cand + i) : 0ull;
            __syncthreads();
            bitonic_desc(s_sort, p2);
            peer_wait_slot(po, s_slot_free != 0);
            emit_sorted(s_sort, min(count, k_eff), k, ids, D, I);
        } else {
            for (uint32_t i = count + threadIdx.x; i < p2; i += blockDim.x) cand[i] = 0ull;
            __syncthreads();
            bitonic_desc(cand, p2);   // global-memory sort for very large k (REPL paging)
            peer_wait_slot(po, s_slot_free != 0);
            emit_sorted(cand, min(count, k_eff), k, ids, D, I);
        }
        peer_signal(po, 1u);
        // every other block is done with this query: reset histograms + selection state for the next search
        // (ws[0] also carries the grid barrier counter: no barrier follows phase 2)
        __syncthreads();
        uint32_t *wz = reinterpret_cast<uint32_t *>(w);
        for (uint32_t i = threadIdx.x; i < sizeof(QueryWs) / 4; i += blockDim.x) wz[i] = 0u;
        __syncthreads();
        if (q == nq_valid - 1) stamp(3);
    }
}

// empty shard: every slot is padding.  One block, so it can take part in the peer protocol.
__global__ void fill_empty_kernel(float *D, int64_t *I, int64_t total, const PeerOut po, uint32_t nq) {
    peer_wait_slot(po);
    for (int64_t i = threadIdx.x; i < total; i += blockDim.x) {
        D[i] = -3.4028234663852886e38f;
        I[i] = -1;
    }
    peer_signal(po, nq);
}

// what the root's merge waits for / announces (null pointers: plain merge of resident lists)
struct MergeSync {
    const uint32_t *done = nullptr;   // [R] cumulative queries delivered per rank
    uint32_t need_done = 0;           // every rank has delivered this search once done[r] >= need_done
    uint32_t *consumed = nullptr;     // += 1 per merged query (frees the slot for the search after next)
    uint32_t *error = nullptr;
};

// merge R sorted per-shard lists: one block per query
__global__ void __launch_bounds__(kPostThreads)
topk_merge_kernel(int R, int64_t nq, int64_t k, const float *D_in,
                  const int64_t *I_in, int64_t shard_stride_D, int64_t shard_stride_I,
                  float *D_out, int64_t *I_out, float *g_s, int64_t *g_i, uint32_t p2, const MergeSync ms) {
    if (ms.done) {
        // the lists arrive over NVLink: acquire every rank's delivery counter before reading them
        if (threadIdx.x < (unsigned)R) spin_until(ms.done + threadIdx.x, ms.need_done, ms.error);
        __syncthreads();
    }
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
            ss[i] = __ldcg(D_in + (size_t)r * shard_stride_D + src);    // L2: peers wrote these lines
            si[i] = __ldcg(I_in + (size_t)r * shard_stride_I + src);
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
    if (ms.consumed) {
        // this query's slot lines have been read (they sit in ss / si): let the peers reuse them
        __syncthreads();
        if (threadIdx.x == 0) {
            __threadfence_system();
            red_release_sys_add(ms.consumed, 1u);
        }
    }
}

// The same merge without a sort, for R k <= 4096: the lists are already sorted, so an entry's final
// position is its own index plus, for every other list, the number of entries that come before it
// there (a binary search).  No barriers between the load and the store; ~50 shared-memory probes per entry.
__global__ void __launch_bounds__(kPostThreads)
topk_merge_rank_kernel(int R, int64_t nq, int64_t k, const float *D_in, const int64_t *I_in,
                       int64_t shard_stride_D, int64_t shard_stride_I, float *D_out, int64_t *I_out,
                       const MergeSync ms) {
    extern __shared__ unsigned char s_raw[];
    if (ms.done) {
        if (threadIdx.x < (unsigned)R) spin_until(ms.done + threadIdx.x, ms.need_done, ms.error);
        __syncthreads();
    }
    const int64_t q = blockIdx.x;
    const uint32_t kk = (uint32_t)k, tot = (uint32_t)R * kk;
    int64_t *si = reinterpret_cast<int64_t *>(s_raw);
    float *ss = reinterpret_cast<float *>(si + tot);
    for (uint32_t i = threadIdx.x; i < tot; i += blockDim.x) {
        const uint32_t r = i / kk, j = i - r * kk;
        const size_t src = (size_t)q * k + j;
        ss[i] = __ldcg(D_in + (size_t)r * shard_stride_D + src);    // L2: peers wrote these lines
        si[i] = __ldcg(I_in + (size_t)r * shard_stride_I + src);
    }
    for (int64_t j = threadIdx.x; j < k; j += blockDim.x) {         // padding first; winners overwrite it
        D_out[q * k + j] = -3.4028234663852886e38f;
        I_out[q * k + j] = -1;
    }
    __syncthreads();
    for (uint32_t i = threadIdx.x; i < tot; i += blockDim.x) {
        const float s = ss[i];
        const int64_t id = si[i];
        if (id < 0) continue;                                       // padding of a short shard
        const uint32_t r = i / kk;
        uint32_t rank = i - r * kk;                                 // everything before it in its own list is valid
        for (uint32_t o = 0; o < (uint32_t)R; o++) {
            if (o == r) continue;
            // entries of list o that come before (s, id): valid, and higher score or equal score with a lower id
            uint32_t lo = 0, hi = kk;
            while (lo < hi) {
                const uint32_t mid = (lo + hi) >> 1;
                const float so = ss[o * kk + mid];
                const int64_t io = si[o * kk + mid];
                const bool before = io >= 0 && (so > s || (so == s && io < id));
                if (before) lo = mid + 1; else hi = mid;
            }
            rank += lo;
        }
        if (rank < kk) {
            D_out[q * k + rank] = s;
            I_out[q * k + rank] = id;
        }
    }
    if (ms.consumed) {
        // this query's slot lines have been read (they sit in ss / si): let the peers reuse them
        __syncthreads();
        if (threadIdx.x == 0) {
            __threadfence_system();
            red_release_sys_add(ms.consumed, 1u);
        }
    }
}

// max over rows of ||x||^2 (as stored), kept per shard: bounds |<q - fp16(q), x>| in the batch path
template <bool F16>
__global__ void rownorm_max_kernel(const uint4 *__restrict__ rows, int64_t n, float *max_norm2) {
    constexpr int ROW_V4 = F16 ? 64 : 128;
    const int lane = threadIdx.x & 31;
    const int64_t warps = (int64_t)gridDim.x * (blockDim.x >> 5);
    float best = 0.f;
    for (int64_t r = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); r < n; r += warps) {
        float s = 0.f;
        for (int c = lane; c < ROW_V4; c += 32) {
            const uint4 u = ld_stream_v4(rows + r * ROW_V4 + c);
            if (F16) {
                const __half2 *h = reinterpret_cast<const __half2 *>(&u);
#pragma unroll
                for (int e = 0; e < 4; e++) { float2 f = __half22float2(h[e]); s = fmaf(f.x, f.x, s); s = fmaf(f.y, f.y, s); }
            } else {
                const float f[4] = {__uint_as_float(u.x), __uint_as_float(u.y), __uint_as_float(u.z), __uint_as_float(u.w)};
#pragma unroll
                for (int e = 0; e < 4; e++) s = fmaf(f[e], f[e], s);
            }
        }
        s = warp_sum(s);
        best = fmaxf(best, s);
    }
    // norms are >= 0: the float bit pattern orders like an unsigned integer
    if (lane == 0 && best > 0.f) atomicMax(reinterpret_cast<unsigned int *>(max_norm2), __float_as_uint(best));
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

namespace {

// Root mailbox (device memory of rank 0, mapped into every rank): a 256-byte header followed by
// kSlots x world result slots of `cap` elements each (D block fp32, then I block int64).
constexpr int kSlots = 2;                 // a rank may run one search ahead of the root's merge
constexpr size_t kMailHeader = 256;
constexpr size_t kOffDone = 0;            // uint32 done[kSlots][kMaxRanks]: queries delivered into a slot, per rank
constexpr size_t kOffConsumed = 128;      // uint32 consumed[kSlots]: queries of a slot merged by the root
constexpr size_t kOffError = 192;         // uint32 error

struct P2P {
    bool attached = false;
    int rank = 0, world = 1;
    int64_t cap = 0;                      // elements per (slot, rank)
    char *own = nullptr;                  // allocation owned by this index (rank 0 only)
    char *root = nullptr;                 // the root's mailbox as addressable from this device
    bool root_is_ipc = false;             // opened with cudaIpcOpenMemHandle (close on free)
    uint32_t seq = 0;                     // searches issued so far (slot = seq % kSlots = the search lane)
    uint32_t ring[kSlots] = {0};          // queries issued into each slot so far (cumulative, wraps)
    // root only: merge scratch for very large k
    size_t slot_bytes() const { return (size_t)cap * 12; }
    float *slot_D(int slot, int r) const { return reinterpret_cast<float *>(root + kMailHeader + ((size_t)slot * world + r) * slot_bytes()); }
    int64_t *slot_I(int slot, int r) const { return reinterpret_cast<int64_t *>(reinterpret_cast<char *>(slot_D(slot, r)) + (size_t)cap * 4); }
    // counters are per slot: the two slots belong to two search lanes whose kernels may finish out of order
    uint32_t *done(int slot) const { return reinterpret_cast<uint32_t *>(root + kOffDone) + slot * kMaxRanks; }
    uint32_t *consumed(int slot) const { return reinterpret_cast<uint32_t *>(root + kOffConsumed) + slot; }
    uint32_t *error() const { return reinterpret_cast<uint32_t *>(root + kOffError); }
    size_t bytes() const { return kMailHeader + (size_t)kSlots * world * slot_bytes(); }
};

}  // namespace

struct cb_index {
    int d = kD;
    int dtype = CB_F16;
    int device = 0;
    int64_t ntotal = 0;
    int64_t capacity = 0;
    void *rows = nullptr;            // [capacity][512] of dtype
    cudaStream_t stream = nullptr;   // used by the host-pointer entry points
    cudaEvent_t add_ev = nullptr;    // orders ix->stream after add_device() calls on foreign streams
    float *max_norm2 = nullptr;      // device scalar: max ||row||^2 over the shard (as stored)
    // search workspace: two lanes.  The plain entry points use lane 0 on the caller's stream; the pipelined
    // submit / join pair alternates lanes (own stream + workspace each), so the selection + exchange tail of
    // one query overlaps the pass over the shard of the next.
    struct Lane {
        float *scores = nullptr;         // [kMaxNQ][score_stride]
        int64_t score_stride = 0;
        QueryWs *ws = nullptr;           // [kMaxNQ]; zero between searches (the search kernel re-zeroes it)
        bool ws_dirty = true;
        uint64_t *cand = nullptr;        // [kMaxNQ][cand_cap]
        uint32_t cand_cap = 0;
        cb::BatchWs *bws = nullptr;      // workspace of the tensor-core batch path
        cudaStream_t stream = nullptr;   // lane stream of the submit API
        cudaEvent_t done = nullptr, fence = nullptr;
        bool used = false;
    } lanes[2];
    int next_lane = 0;
    bool half_grid = false;          // submit API: two searches in flight share the SMs
    // global-id segments of a shard inside a multi-shard index (cb_sharded): device copies
    int64_t *seg_dev = nullptr;      // [2][seg_cap]: local starts, then global starts
    int seg_cap = 0, nseg = 0;
    // pinned + device staging for the host-pointer entry points
    float *h_q = nullptr, *d_q = nullptr;
    int64_t q_cap = 0;               // queries
    float *h_D = nullptr, *d_D = nullptr;
    int64_t *h_I = nullptr, *d_I = nullptr;
    int64_t out_cap = 0;             // nq*k elements
    void *d_stage = nullptr;         // add() staging
    int64_t stage_bytes = 0;
    int scan_blocks_per_sm[2][3] = {{0}};
    int sms = 0;
    int64_t n_batch_searches = 0;
    P2P p2p;
    // optional live timing of the scan kernel (bench.py roofline): event pairs on the
    // launching stream, resolved lazily by cb_flatip_timing_read
    bool timing = false;
    static constexpr int kEv = 256;
    cudaEvent_t ev0[kEv] = {nullptr}, ev1[kEv] = {nullptr};
    int ev_n = 0;
    unsigned long long *phase_t = nullptr;   // device: 4 %globaltimer stamps of the most recent timed launch

    size_t row_bytes() const { return (size_t)d * (dtype == CB_F16 ? 2 : 4); }
    IdMap idmap(int64_t id_base) const {
        IdMap m;
        if (nseg > 0) { m.seg_local = seg_dev; m.seg_global = seg_dev + seg_cap; m.nseg = nseg; }
        m.id_base = id_base;
        return m;
    }
};

static int grow_rows(cb_index *ix, int64_t need, cudaStream_t s) {
    if (need <= ix->capacity) return CB_OK;
    int64_t cap = std::max<int64_t>(need, ix->capacity + ix->capacity / 2);
    cap = (cap + 63) / 64 * 64;
    void *p = nullptr;
    CB_CUDA(cudaMalloc(&p, (size_t)cap * ix->row_bytes()));
    // the old buffer may still be read by searches queued on the index's own stream
    if (ix->stream && ix->stream != s) CB_CUDA(cudaStreamSynchronize(ix->stream));
    if (ix->rows && ix->ntotal)
        CB_CUDA(cudaMemcpyAsync(p, ix->rows, (size_t)ix->ntotal * ix->row_bytes(),
                                cudaMemcpyDeviceToDevice, s));
    CB_CUDA(cudaStreamSynchronize(s));
    if (ix->rows) CB_CUDA(cudaFree(ix->rows));
    ix->rows = p;
    ix->capacity = cap;
    return CB_OK;
}

static int ensure_ws(cb_index *ix, cb_index::Lane &L, int64_t k_eff) {
    const int64_t stride = (ix->ntotal + 63) / 64 * 64;
    if (stride > L.score_stride) {
        if (L.scores) CB_CUDA(cudaFree(L.scores));
        L.scores = nullptr;
        int64_t s = std::max<int64_t>(stride, (ix->capacity + 63) / 64 * 64);
        CB_CUDA(cudaMalloc(&L.scores, (size_t)kMaxNQ * s * sizeof(float)));
        L.score_stride = s;
    }
    if (!L.ws) {
        CB_CUDA(cudaMalloc(&L.ws, sizeof(QueryWs) * kMaxNQ));
        L.ws_dirty = true;
    }
    // room for the short list (everything in the bins at or above the k-th score's bin: k plus a bin's worth
    // of rows) as long as it can still be sorted in shared memory; never less than k
    uint32_t p2 = 256;
    while ((int64_t)p2 < std::min<int64_t>(2 * k_eff, kSortSmem) || (int64_t)p2 < k_eff) p2 <<= 1;
    if (p2 > L.cand_cap) {
        if (L.cand) CB_CUDA(cudaFree(L.cand));
        L.cand = nullptr;
        CB_CUDA(cudaMalloc(&L.cand, (size_t)kMaxNQ * p2 * sizeof(uint64_t)));
        L.cand_cap = p2;
    }
    return CB_OK;
}

template <int NQ, bool F16>
static int launch_search(cb_index *ix, cb_index::Lane &L, const float *q_dev, int nq_valid, uint32_t k_eff, int64_t k,
                         float *D_dev, int64_t *I_dev, const IdMap &ids, const PeerOut &po, cudaStream_t s) {
    auto kern = flatip_search_kernel<NQ, F16>;
    // phase 1 histograms, later reused as the sort buffer of the last block
    const size_t smem = std::max<size_t>((size_t)NQ * kBins0 * sizeof(uint32_t), (size_t)kSortSmem * sizeof(uint64_t));
    constexpr int slot = NQ == 1 ? 0 : NQ == 2 ? 1 : 2;
    int &bps = ix->scan_blocks_per_sm[F16 ? 1 : 0][slot];
    if (bps == 0) {
        CB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        CB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&bps, kern, kScanThreads, smem));
        if (bps < 1) bps = 1;
    }
    constexpr int R = F16 ? 8 : 4;
    int64_t groups = (ix->ntotal + R - 1) / R;
    int64_t want = (groups + (kScanThreads / 32) - 1) / (kScanThreads / 32);
    // cooperative launch: every block is resident (the kernel has grid-wide barriers).  Two searches in flight
    // (submit API) take half of the residency each, so that both are resident at once.
    const int per_sm = ix->half_grid ? std::max(1, bps / 2) : bps;
    int grid = (int)std::min<int64_t>((int64_t)ix->sms * per_sm, std::max<int64_t>(want, 1));
    const bool timed = ix->timing && ix->ev_n < cb_index::kEv;
    if (timed) {
        if (!ix->ev0[ix->ev_n]) {
            CB_CUDA(cudaEventCreate(&ix->ev0[ix->ev_n]));
            CB_CUDA(cudaEventCreate(&ix->ev1[ix->ev_n]));
        }
        CB_CUDA(cudaEventRecord(ix->ev0[ix->ev_n], s));
    }
    const uint4 *rows = (const uint4 *)ix->rows;
    int64_t n = ix->ntotal, stride = L.score_stride;
    float *scores = L.scores;
    QueryWs *ws = L.ws;
    uint64_t *cand = L.cand;
    uint32_t cand_cap = L.cand_cap;
    const float *mx = ix->max_norm2;
    IdMap ids_v = ids;
    PeerOut po_v = po;
    unsigned long long *phase_t = nullptr;
    if (ix->timing) {
        if (!ix->phase_t) {
            CB_CUDA(cudaMalloc(&ix->phase_t, 64));
            CB_CUDA(cudaMemset(ix->phase_t, 0, 64));
        }
        phase_t = ix->phase_t;
    }
    void *args[] = {&rows, &n, &q_dev, &nq_valid, &scores, &stride, &ws, &k_eff, &k, &cand, &cand_cap, &mx,
                    &ids_v, &po_v, &D_dev, &I_dev, &phase_t};
    CB_CUDA(cudaLaunchCooperativeKernel((const void *)kern, dim3(grid), dim3(kScanThreads), args, smem, s));
    CB_LAUNCH_CHECK();
    if (timed) {
        CB_CUDA(cudaEventRecord(ix->ev1[ix->ev_n], s));
        ix->ev_n++;
    }
    return CB_OK;
}

// the streaming path over <= kMaxNQ queries: ONE cooperative launch (scan + select + write)
static int search_tile(cb_index *ix, cb_index::Lane &L, int nq, const float *q_dev, int64_t k, float *D_dev,
                       int64_t *I_dev, const IdMap &ids, const PeerOut &po, cudaStream_t s) {
    const int64_t n = ix->ntotal;
    const uint32_t k_eff = (uint32_t)std::min<int64_t>(k, n);
    // histograms + state are zero between searches: the kernel's last block re-zeroes them.  Only a search
    // that failed half way needs a memset.
    if (L.ws_dirty) CB_CUDA(cudaMemsetAsync(L.ws, 0, sizeof(QueryWs) * kMaxNQ, s));
    L.ws_dirty = true;
    const bool f16 = ix->dtype == CB_F16;
    int rc;
    if (nq == 1) rc = f16 ? launch_search<1, true>(ix, L, q_dev, nq, k_eff, k, D_dev, I_dev, ids, po, s)
                          : launch_search<1, false>(ix, L, q_dev, nq, k_eff, k, D_dev, I_dev, ids, po, s);
    else if (nq == 2) rc = f16 ? launch_search<2, true>(ix, L, q_dev, nq, k_eff, k, D_dev, I_dev, ids, po, s)
                               : launch_search<2, false>(ix, L, q_dev, nq, k_eff, k, D_dev, I_dev, ids, po, s);
    else rc = f16 ? launch_search<4, true>(ix, L, q_dev, nq, k_eff, k, D_dev, I_dev, ids, po, s)
                  : launch_search<4, false>(ix, L, q_dev, nq, k_eff, k, D_dev, I_dev, ids, po, s);
    if (rc) return rc;
    L.ws_dirty = false;
    return CB_OK;
}

// Dispatch of one search over this shard.  nq < batch_min (or fp32 storage, or k > 1024, or a tiny
// shard) streams the shard once per 4 queries (HBM-bound scan); larger batches on fp16 shards run the
// tcgen05 GEMM with a fused per-query threshold filter and exact fp32 re-scoring (tensor-bound).
// Nothing here synchronises.  With `po` set, every query's list is delivered to the root's mailbox.
static int search_core(cb_index *ix, int64_t nq, const float *q_dev, int64_t k, float *D_dev, int64_t *I_dev,
                       const IdMap &ids, const PeerOut &po, cudaStream_t s, int lane = 0) {
    cb_index::Lane &L = ix->lanes[lane];
    if (ix->ntotal == 0) {
        fill_empty_kernel<<<1, 256, 0, s>>>(D_dev, I_dev, nq * k, po, (uint32_t)nq);
        CB_LAUNCH_CHECK();
        return CB_OK;
    }
    // The GEMM pass over 10M rows costs 1.7-1.9 ms whatever nq <= 128 is (HBM-bound, profiles/r02_search_probe.txt),
    // the streaming scan 1.5 ms for one query, 1.9 ms for two, 2.6 ms for four and another pass per four after
    // that.  Small shards keep a higher threshold: there both paths are bounded by their launch counts, not
    // by the pass over the rows.
    int64_t batch_min = ix->ntotal >= (1ll << 20) ? 2 : 16;
    if (tune(T_BATCH_MIN_NQ) > 0) batch_min = tune(T_BATCH_MIN_NQ);
    if (nq >= batch_min && ix->dtype == CB_F16 && k <= 1024 && ix->ntotal >= 8192) {
        if (!L.bws) L.bws = batch_ws_new();
        int brc = flatip_search_batch(L.bws, ix->rows, ix->ntotal, ix->device, nq, q_dev, k, D_dev, I_dev, ids, po,
                                      ix->max_norm2, s);
        if (brc) return brc;
        ix->n_batch_searches++;
        return CB_OK;
    }
    int rc = ensure_ws(ix, L, std::min<int64_t>(k, ix->ntotal));
    if (rc) return rc;
    for (int64_t q0 = 0; q0 < nq; q0 += kMaxNQ) {
        int t = (int)std::min<int64_t>(kMaxNQ, nq - q0);
        rc = search_tile(ix, L, t, q_dev + q0 * ix->d, k, D_dev + q0 * k, I_dev + q0 * k, ids, po, s);
        if (rc) return rc;
    }
    return CB_OK;
}

static int launch_merge(int R, int64_t nq, int64_t k, const float *D_in, const int64_t *I_in, int64_t shard_stride_D,
                        int64_t shard_stride_I, float *D_out, int64_t *I_out, const MergeSync &ms, cudaStream_t s) {
    if ((int64_t)R * k <= 4096) {
        // sorted lists, few entries: place every entry by rank (no sort, no barriers)
        topk_merge_rank_kernel<<<(unsigned)nq, kPostThreads, (size_t)R * k * 12, s>>>(
            R, nq, k, D_in, I_in, shard_stride_D, shard_stride_I, D_out, I_out, ms);
        CB_LAUNCH_CHECK();
        return CB_OK;
    }
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
                                                               D_out, I_out, g_s, g_i, p2, ms);
    CB_LAUNCH_CHECK();
    if (g_s) { CB_CUDA(cudaFreeAsync(g_s, s)); CB_CUDA(cudaFreeAsync(g_i, s)); }
    return CB_OK;
}

static int ensure_q(cb_index *ix, int64_t nq) {
    if (nq <= ix->q_cap) return CB_OK;
    cudaFreeHost(ix->h_q); cudaFree(ix->d_q);
    ix->h_q = nullptr; ix->d_q = nullptr; ix->q_cap = 0;
    CB_CUDA(cudaMallocHost(&ix->h_q, (size_t)nq * ix->d * 4));
    CB_CUDA(cudaMalloc(&ix->d_q, (size_t)nq * ix->d * 4));
    ix->q_cap = nq;
    return CB_OK;
}

static int ensure_out(cb_index *ix, int64_t elems) {
    if (elems <= ix->out_cap) return CB_OK;
    cudaFreeHost(ix->h_D); cudaFreeHost(ix->h_I); cudaFree(ix->d_D); cudaFree(ix->d_I);
    ix->h_D = nullptr; ix->h_I = nullptr; ix->d_D = nullptr; ix->d_I = nullptr; ix->out_cap = 0;
    CB_CUDA(cudaMallocHost(&ix->h_D, (size_t)elems * 4));
    CB_CUDA(cudaMallocHost(&ix->h_I, (size_t)elems * 8));
    CB_CUDA(cudaMalloc(&ix->d_D, (size_t)elems * 4));
    CB_CUDA(cudaMalloc(&ix->d_I, (size_t)elems * 8));
    ix->out_cap = elems;
    return CB_OK;
}

// ---- peer delivery: attach / one search ---------------------------------------------------
static void p2p_detach(cb_index *ix) {
    P2P &p = ix->p2p;
    if (p.root_is_ipc && p.root) cudaIpcCloseMemHandle(p.root);
    if (p.own) cudaFree(p.own);
    p = P2P();
}

static int p2p_alloc_root(cb_index *ix, int world, int64_t max_elems) {
    P2P &p = ix->p2p;
    p.rank = 0;
    p.world = world;
    p.cap = (std::max<int64_t>(max_elems, 1024) + 1) / 2 * 2;
    CB_CUDA(cudaMalloc(&p.own, p.bytes()));
    CB_CUDA(cudaMemset(p.own, 0, kMailHeader));
    p.root = p.own;
    p.attached = true;
    return CB_OK;
}

// The local part of a sharded search: this rank's top-k goes into the root's mailbox slot.
// Returns the slot and the cumulative query count after this search.
static int p2p_local_search(cb_index *ix, int64_t nq, const float *q_dev, int64_t k, int64_t id_base, cudaStream_t s,
                            int *slot_out, uint32_t *total_out, int lane = -1) {
    P2P &p = ix->p2p;
    // slot = lane of the submit API; the plain entry points alternate slots on one stream (every rank issues
    // the same sequence of calls, so every rank picks the same slot)
    const int slot = lane >= 0 ? lane : (int)(p.seq % kSlots);
    PeerOut po;
    po.done = p.done(slot) + p.rank;
    po.consumed = p.consumed(slot);
    po.need_consumed = p.ring[slot];          // the search that last used this slot must be merged
    po.error = p.error();
    int rc = search_core(ix, nq, q_dev, k, p.slot_D(slot, p.rank), p.slot_I(slot, p.rank), ix->idmap(id_base), po, s,
                         std::max(lane, 0));
    if (rc) return rc;
    p.ring[slot] += (uint32_t)nq;
    p.seq++;
    *slot_out = slot;
    *total_out = p.ring[slot];
    return CB_OK;
}

static int p2p_root_merge(cb_index *ix, int slot, uint32_t total, int64_t nq, int64_t k, float *D_dev, int64_t *I_dev,
                          cudaStream_t s) {
    P2P &p = ix->p2p;
    MergeSync ms;
    ms.done = p.done(slot);
    ms.need_done = total;
    ms.consumed = p.consumed(slot);
    ms.error = p.error();
    const int64_t stride_D = (int64_t)(p.slot_bytes() / 4), stride_I = (int64_t)(p.slot_bytes() / 8);
    return launch_merge(p.world, nq, k, p.slot_D(slot, 0), p.slot_I(slot, 0), stride_D, stride_I, D_dev, I_dev, ms, s);
}

static int set_segments(cb_index *ix, const std::vector<int64_t> &local, const std::vector<int64_t> &global) {
    const int n = (int)local.size();
    if (n <= 1) {            // a single run of ids is just an offset
        ix->nseg = 0;
        return CB_OK;
    }
    DeviceGuard g(ix->device);
    if (n > ix->seg_cap) {
        if (ix->seg_dev) CB_CUDA(cudaFree(ix->seg_dev));
        ix->seg_dev = nullptr;
        const int cap = std::max(64, n * 2);
        CB_CUDA(cudaMalloc(&ix->seg_dev, (size_t)cap * 2 * sizeof(int64_t)));
        ix->seg_cap = cap;
    }
    CB_CUDA(cudaStreamSynchronize(ix->stream));
    CB_CUDA(cudaMemcpy(ix->seg_dev, local.data(), (size_t)n * 8, cudaMemcpyHostToDevice));
    CB_CUDA(cudaMemcpy(ix->seg_dev + ix->seg_cap, global.data(), (size_t)n * 8, cudaMemcpyHostToDevice));
    ix->nseg = n;
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
    ix->sms = kNumSMs;
    cudaDeviceGetAttribute(&ix->sms, cudaDevAttrMultiProcessorCount, device);
    cudaError_t e = cudaStreamCreateWithFlags(&ix->stream, cudaStreamNonBlocking);
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&ix->add_ev, cudaEventDisableTiming);
    if (e == cudaSuccess) e = cudaMalloc(&ix->max_norm2, 256);
    if (e == cudaSuccess) e = cudaMemset(ix->max_norm2, 0, 256);
    if (e != cudaSuccess) {
        set_error("cb_flatip_create: %s", cudaGetErrorString(e));
        cb_flatip_free(ix);
        return CB_ERR_CUDA;
    }
    *out = ix;
    return CB_OK;
}

void cb_flatip_free(cb_index *ix) {
    if (!ix) return;
    DeviceGuard g(ix->device);
    if (ix->stream) cudaStreamSynchronize(ix->stream);
    p2p_detach(ix);
    cudaFree(ix->rows);
    for (cb_index::Lane &L : ix->lanes) {
        if (L.stream) cudaStreamSynchronize(L.stream);
        cudaFree(L.scores); cudaFree(L.ws); cudaFree(L.cand);
        cb::batch_ws_delete(L.bws);
        if (L.done) cudaEventDestroy(L.done);
        if (L.fence) cudaEventDestroy(L.fence);
        if (L.stream) cudaStreamDestroy(L.stream);
    }
    cudaFree(ix->d_q); cudaFree(ix->d_D); cudaFree(ix->d_I); cudaFree(ix->d_stage);
    cudaFree(ix->max_norm2); cudaFree(ix->seg_dev); cudaFree(ix->phase_t);
    cudaFreeHost(ix->h_q); cudaFreeHost(ix->h_D); cudaFreeHost(ix->h_I);
    for (int i = 0; i < cb_index::kEv; i++) {
        if (ix->ev0[i]) cudaEventDestroy(ix->ev0[i]);
        if (ix->ev1[i]) cudaEventDestroy(ix->ev1[i]);
    }
    if (ix->add_ev) cudaEventDestroy(ix->add_ev);
    if (ix->stream) cudaStreamDestroy(ix->stream);
    delete ix;
}

int64_t cb_flatip_ntotal(const cb_index *ix) { return ix ? ix->ntotal : -1; }
int cb_flatip_dim(const cb_index *ix) { return ix ? ix->d : -1; }
int cb_flatip_storage_dtype(const cb_index *ix) { return ix ? ix->dtype : -1; }
int cb_flatip_device(const cb_index *ix) { return ix ? ix->device : -1; }
const void *cb_flatip_device_rows(const cb_index *ix) { return ix ? ix->rows : nullptr; }

int cb_flatip_reserve(cb_index *ix, int64_t n_rows) {
    CB_REQUIRE(ix != nullptr, "cb_flatip_reserve: null index");
    CB_REQUIRE(n_rows >= 0 && n_rows < (1ll << 32), "cb_flatip_reserve: a shard holds < 2^32 rows");
    DeviceGuard g(ix->device);
    return grow_rows(ix, n_rows, ix->stream);
}

int cb_flatip_reset(cb_index *ix) {
    CB_REQUIRE(ix != nullptr, "cb_flatip_reset: null index");
    DeviceGuard g(ix->device);
    CB_CUDA(cudaStreamSynchronize(ix->stream));
    CB_CUDA(cudaMemset(ix->max_norm2, 0, 4));
    ix->ntotal = 0;
    ix->nseg = 0;
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
    {
        const int grid = (int)std::min<int64_t>((int64_t)ix->sms * 8, (n + 7) / 8);
        if (ix->dtype == CB_F16) rownorm_max_kernel<true><<<grid, 256, 0, s>>>((const uint4 *)dst, n, ix->max_norm2);
        else rownorm_max_kernel<false><<<grid, 256, 0, s>>>((const uint4 *)dst, n, ix->max_norm2);
        CB_LAUNCH_CHECK();
    }
    // the host-pointer entry points (search, get_rows) run on the index's own stream: order it after
    // this add, whatever stream the caller used
    if (s != ix->stream) {
        CB_CUDA(cudaEventRecord(ix->add_ev, s));
        CB_CUDA(cudaStreamWaitEvent(ix->stream, ix->add_ev, 0));
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
    return search_core(ix, nq, q_dev, k, D_dev, I_dev, ix->idmap(id_base), PeerOut(), (cudaStream_t)stream);
}

int cb_flatip_search(cb_index *ix, int64_t nq, const float *q_host, int64_t k, float *D_host,
                     int64_t *I_host) {
    CB_REQUIRE(ix != nullptr, "cb_flatip_search: null index");
    CB_REQUIRE(nq >= 0, "cb_flatip_search: nq < 0");
    CB_REQUIRE(k > 0, "cb_flatip_search: k must be > 0 (got %lld)", (long long)k);
    if (nq == 0) return CB_OK;
    CB_REQUIRE(q_host && D_host && I_host, "cb_flatip_search: null buffer");
    DeviceGuard g(ix->device);
    int rc;
    if ((rc = ensure_q(ix, nq))) return rc;
    if ((rc = ensure_out(ix, nq * k))) return rc;
    memcpy(ix->h_q, q_host, (size_t)nq * ix->d * 4);
    CB_CUDA(cudaMemcpyAsync(ix->d_q, ix->h_q, (size_t)nq * ix->d * 4, cudaMemcpyHostToDevice, ix->stream));
    rc = cb_flatip_search_device(ix, nq, ix->d_q, k, ix->d_D, ix->d_I, 0, ix->stream);
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
    return launch_merge(R, nq, k, D_in, I_in, shard_stride_D, shard_stride_I, D_out, I_out, MergeSync(),
                        (cudaStream_t)stream);
}

// ---- sharded search, one process per GPU: NVLink peer delivery instead of a collective --------
int cb_flatip_p2p_init(cb_index *ix, int rank, int world, int64_t max_elems, void *ipc_handle_out64) {
    CB_REQUIRE(ix && ipc_handle_out64, "cb_flatip_p2p_init: null argument");
    CB_REQUIRE(world >= 1 && world <= kMaxRanks && rank >= 0 && rank < world,
               "cb_flatip_p2p_init: rank %d / world %d out of range (at most %d ranks)", rank, world, kMaxRanks);
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
    DeviceGuard g(ix->device);
    CB_CUDA(cudaStreamSynchronize(ix->stream));
    p2p_detach(ix);
    memset(ipc_handle_out64, 0, 64);
    if (rank == 0) {
        int rc = p2p_alloc_root(ix, world, max_elems);
        if (rc) return rc;
        if (world > 1) {
            cudaIpcMemHandle_t h;
            CB_CUDA(cudaIpcGetMemHandle(&h, ix->p2p.own));
            memcpy(ipc_handle_out64, &h, 64);
        }
    } else {
        ix->p2p.rank = rank;
        ix->p2p.world = world;
        ix->p2p.cap = (std::max<int64_t>(max_elems, 1024) + 1) / 2 * 2;
    }
    return CB_OK;
}

int cb_flatip_p2p_connect(cb_index *ix, const void *root_ipc_handle64) {
    CB_REQUIRE(ix != nullptr, "cb_flatip_p2p_connect: null index");
    P2P &p = ix->p2p;
    if (p.rank == 0) {
        CB_REQUIRE(p.attached, "cb_flatip_p2p_connect: call cb_flatip_p2p_init first");
        return CB_OK;
    }
    CB_REQUIRE(root_ipc_handle64 != nullptr, "cb_flatip_p2p_connect: null handle");
    CB_REQUIRE(p.cap > 0, "cb_flatip_p2p_connect: call cb_flatip_p2p_init first");
    DeviceGuard g(ix->device);
    cudaIpcMemHandle_t h;
    memcpy(&h, root_ipc_handle64, 64);
    void *ptr = nullptr;
    CB_CUDA(cudaIpcOpenMemHandle(&ptr, h, cudaIpcMemLazyEnablePeerAccess));
    p.root = (char *)ptr;
    p.root_is_ipc = true;
    p.attached = true;
    return CB_OK;
}

int cb_flatip_search_p2p_device(cb_index *ix, int64_t nq, const float *q_dev, int64_t k, float *D_dev,
                                int64_t *I_dev, int64_t id_base, void *stream) {
    CB_REQUIRE(ix != nullptr, "cb_flatip_search_p2p_device: null index");
    P2P &p = ix->p2p;
    CB_REQUIRE(p.attached, "cb_flatip_search_p2p_device: not attached (cb_flatip_p2p_init / _connect)");
    CB_REQUIRE(nq >= 0, "cb_flatip_search_p2p_device: nq < 0");
    CB_REQUIRE(k > 0 && k <= p.cap, "cb_flatip_search_p2p_device: k = %lld does not fit a mailbox slot of %lld elements",
               (long long)k, (long long)p.cap);
    if (nq == 0) return CB_OK;
    CB_REQUIRE(q_dev != nullptr, "cb_flatip_search_p2p_device: null query");
    CB_REQUIRE(p.rank != 0 || (D_dev && I_dev), "cb_flatip_search_p2p_device: the root needs output buffers");
    DeviceGuard g(ix->device);
    cudaStream_t s = (cudaStream_t)stream;
    const int64_t chunk = std::max<int64_t>(1, p.cap / k);       // queries per mailbox slot
    for (int64_t q0 = 0; q0 < nq; q0 += chunk) {
        const int64_t m = std::min(chunk, nq - q0);
        int slot = 0;
        uint32_t total = 0;
        int rc = p2p_local_search(ix, m, q_dev + q0 * ix->d, k, id_base, s, &slot, &total);
        if (rc) return rc;
        if (p.rank == 0 && (rc = p2p_root_merge(ix, slot, total, m, k, D_dev + q0 * k, I_dev + q0 * k, s))) return rc;
    }
    return CB_OK;
}

// Host-buffer form of the sharded search (the call a host program makes per query): pinned staging, H2D of
// the query, the rank's kernel chain, and on rank 0 the merge, D2H and one synchronisation.
int cb_flatip_search_p2p(cb_index *ix, int64_t nq, const float *q_host, int64_t k, float *D_host, int64_t *I_host,
                         int64_t id_base) {
    CB_REQUIRE(ix != nullptr, "cb_flatip_search_p2p: null index");
    CB_REQUIRE(ix->p2p.attached, "cb_flatip_search_p2p: not attached (cb_flatip_p2p_init / _connect)");
    CB_REQUIRE(nq > 0 && k > 0, "cb_flatip_search_p2p: nq and k must be > 0");
    CB_REQUIRE(q_host != nullptr, "cb_flatip_search_p2p: null query");
    const bool root = ix->p2p.rank == 0;
    CB_REQUIRE(!root || (D_host && I_host), "cb_flatip_search_p2p: the root needs output buffers");
    DeviceGuard g(ix->device);
    int rc;
    if ((rc = ensure_q(ix, nq))) return rc;
    if (root && (rc = ensure_out(ix, nq * k))) return rc;
    memcpy(ix->h_q, q_host, (size_t)nq * ix->d * 4);
    CB_CUDA(cudaMemcpyAsync(ix->d_q, ix->h_q, (size_t)nq * ix->d * 4, cudaMemcpyHostToDevice, ix->stream));
    rc = cb_flatip_search_p2p_device(ix, nq, ix->d_q, k, ix->d_D, ix->d_I, id_base, ix->stream);
    if (rc) return rc;
    if (root) {
        CB_CUDA(cudaMemcpyAsync(ix->h_D, ix->d_D, (size_t)nq * k * 4, cudaMemcpyDeviceToHost, ix->stream));
        CB_CUDA(cudaMemcpyAsync(ix->h_I, ix->d_I, (size_t)nq * k * 8, cudaMemcpyDeviceToHost, ix->stream));
    }
    CB_CUDA(cudaStreamSynchronize(ix->stream));      // the staged query must not be overwritten by the next call
    if (root) {
        memcpy(D_host, ix->h_D, (size_t)nq * k * 4);
        memcpy(I_host, ix->h_I, (size_t)nq * k * 8);
    }
    return CB_OK;
}

// Pipelined form: queue one search and return.  Searches alternate between two lanes (own stream, workspace
// and mailbox slot); each lane's kernels take half of the SMs' residency, so the selection / exchange / merge
// tail of one query overlaps the pass over the shard of the next.  Works with or without an attached mailbox.
int cb_flatip_submit_search_device(cb_index *ix, int64_t nq, const float *q_dev, int64_t k, float *D_dev,
                                   int64_t *I_dev, int64_t id_base, void *after_stream) {
    CB_REQUIRE(ix != nullptr, "cb_flatip_submit_search_device: null index");
    CB_REQUIRE(nq > 0 && k > 0 && k < (1ll << 31), "cb_flatip_submit_search_device: nq and k must be > 0");
    CB_REQUIRE(q_dev != nullptr, "cb_flatip_submit_search_device: null query");
    P2P &p = ix->p2p;
    const bool peer = p.attached && p.world > 1;
    CB_REQUIRE((peer && p.rank != 0) || (D_dev && I_dev), "cb_flatip_submit_search_device: null output buffer");
    CB_REQUIRE(!peer || nq * k <= p.cap, "cb_flatip_submit_search_device: nq * k = %lld exceeds a mailbox slot (%lld)",
               (long long)(nq * k), (long long)p.cap);
    DeviceGuard g(ix->device);
    const int lane = ix->next_lane;
    ix->next_lane ^= 1;
    cb_index::Lane &L = ix->lanes[lane];
    if (!L.stream) {
        CB_CUDA(cudaStreamCreateWithFlags(&L.stream, cudaStreamNonBlocking));
        CB_CUDA(cudaEventCreateWithFlags(&L.done, cudaEventDisableTiming));
        CB_CUDA(cudaEventCreateWithFlags(&L.fence, cudaEventDisableTiming));
    }
    CB_CUDA(cudaEventRecord(L.fence, (cudaStream_t)after_stream));
    CB_CUDA(cudaStreamWaitEvent(L.stream, L.fence, 0));
    ix->half_grid = true;
    int rc;
    if (peer) {
        int slot = 0;
        uint32_t total = 0;
        rc = p2p_local_search(ix, nq, q_dev, k, id_base, L.stream, &slot, &total, lane);
        if (!rc && p.rank == 0) rc = p2p_root_merge(ix, slot, total, nq, k, D_dev, I_dev, L.stream);
    } else {
        rc = search_core(ix, nq, q_dev, k, D_dev, I_dev, ix->idmap(id_base), PeerOut(), L.stream, lane);
    }
    ix->half_grid = false;
    if (rc) return rc;
    CB_CUDA(cudaEventRecord(L.done, L.stream));
    L.used = true;
    return CB_OK;
}

// make `stream` wait for every submitted search (no host synchronisation)
int cb_flatip_join(cb_index *ix, void *stream) {
    CB_REQUIRE(ix != nullptr, "cb_flatip_join: null index");
    DeviceGuard g(ix->device);
    for (cb_index::Lane &L : ix->lanes)
        if (L.used) CB_CUDA(cudaStreamWaitEvent((cudaStream_t)stream, L.done, 0));
    return CB_OK;
}

int cb_flatip_p2p_status(cb_index *ix, int *error) {
    CB_REQUIRE(ix && error, "cb_flatip_p2p_status: null argument");
    *error = 0;
    if (!ix->p2p.attached) return CB_OK;
    DeviceGuard g(ix->device);
    uint32_t e = 0;
    CB_CUDA(cudaMemcpy(&e, ix->p2p.error(), 4, cudaMemcpyDeviceToHost));
    *error = (int)e;
    return CB_OK;
}

int cb_flatip_batch_stats(cb_index *ix, int64_t *n_batch_searches, int64_t *n_rescued) {
    CB_REQUIRE(ix && n_batch_searches && n_rescued, "cb_flatip_batch_stats: null argument");
    *n_batch_searches = ix->n_batch_searches;
    *n_rescued = 0;
    DeviceGuard g(ix->device);
    for (cb_index::Lane &L : ix->lanes)
        if (L.bws) {
            int64_t r = 0;
            int rc = batch_ws_stats(L.bws, &r, ix->stream);
            if (rc) return rc;
            *n_rescued += r;
        }
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

int cb_flatip_phase_times(cb_index *ix, double *ms3) {
    CB_REQUIRE(ix && ms3, "cb_flatip_phase_times: null argument");
    ms3[0] = ms3[1] = ms3[2] = 0;
    if (!ix->phase_t) return CB_OK;
    DeviceGuard g(ix->device);
    unsigned long long t[4];
    CB_CUDA(cudaMemcpy(t, ix->phase_t, sizeof(t), cudaMemcpyDeviceToHost));
    if (t[3] >= t[2] && t[2] >= t[1] && t[1] >= t[0]) {
        ms3[0] = (double)(t[1] - t[0]) * 1e-6;
        ms3[1] = (double)(t[2] - t[1]) * 1e-6;
        ms3[2] = (double)(t[3] - t[2]) * 1e-6;
    }
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

// =====================================================================================
// cb_sharded: one index over several GPUs driven by ONE process (the REPL of query-index.py)
// =====================================================================================
struct cb_sharded {
    int d = kD, dtype = CB_F16;
    std::vector<cb_index *> sh;
    std::vector<int> devices;
    // per shard: runs of consecutive global ids (local start, global start, count)
    struct Seg { int64_t local, global, count; };
    std::vector<std::vector<Seg>> segs;
    int64_t ntotal = 0;
    std::vector<cudaEvent_t> ev_done;     // shard r's local search has been queued / finished
    cudaEvent_t ev_q = nullptr;           // the query is ready on the root stream
    int64_t mail_cap = 0;
};

namespace {

int sharded_sync_segments(cb_sharded *S, int r) {
    std::vector<int64_t> l, g;
    for (const auto &sg : S->segs[r]) { l.push_back(sg.local); g.push_back(sg.global); }
    return set_segments(S->sh[r], l, g);
}

void sharded_note_rows(cb_sharded *S, int r, int64_t local0, int64_t global0, int64_t count) {
    auto &v = S->segs[r];
    if (!v.empty() && v.back().local + v.back().count == local0 && v.back().global + v.back().count == global0)
        v.back().count += count;          // still one run
    else
        v.push_back({local0, global0, count});
}

int64_t shard_id_base(const cb_sharded *S, int r) {
    const auto &v = S->segs[r];
    return v.size() == 1 ? v[0].global - v[0].local : 0;
}

// (re)allocate the root mailbox so that nq x k elements fit one slot
int sharded_ensure_mailbox(cb_sharded *S, int64_t elems) {
    const int R = (int)S->sh.size();
    if (elems <= S->mail_cap && S->sh[0]->p2p.attached) return CB_OK;
    const int64_t cap = std::max<int64_t>(elems, 128 * 1024);
    for (cb_index *ix : S->sh) {
        DeviceGuard g(ix->device);
        CB_CUDA(cudaStreamSynchronize(ix->stream));
        p2p_detach(ix);
    }
    {
        DeviceGuard g(S->sh[0]->device);
        int rc = p2p_alloc_root(S->sh[0], R, cap);
        if (rc) return rc;
    }
    for (int r = 1; r < R; r++) {
        P2P &p = S->sh[r]->p2p;
        p.rank = r;
        p.world = R;
        p.cap = S->sh[0]->p2p.cap;
        p.root = S->sh[0]->p2p.root;      // same process: peer access makes the pointer valid on every device
        p.attached = true;
    }
    S->mail_cap = S->sh[0]->p2p.cap;
    return CB_OK;
}

// q: device pointer on devices[0] (q_host == nullptr) or pinned host memory; outputs on devices[0];
// s0: stream of devices[0] that the outputs are ordered on
int sharded_search_impl(cb_sharded *S, int64_t nq, const float *q_dev0, const float *q_host, int64_t k,
                        float *D_dev0, int64_t *I_dev0, cudaStream_t s0) {
    const int R = (int)S->sh.size();
    cb_index *root = S->sh[0];
    if (R == 1) {
        DeviceGuard g(root->device);
        const float *q = q_dev0;
        if (q_host) {
            CB_CUDA(cudaMemcpyAsync(root->d_q, q_host, (size_t)nq * kD * 4, cudaMemcpyHostToDevice, s0));
            q = root->d_q;
        }
        return search_core(root, nq, q, k, D_dev0, I_dev0, root->idmap(shard_id_base(S, 0)), PeerOut(), s0);
    }
    int rc = sharded_ensure_mailbox(S, std::min<int64_t>(nq, 1024) * k);
    if (rc) return rc;
    const int64_t chunk = std::max<int64_t>(1, S->mail_cap / k);
    CB_REQUIRE(k <= S->mail_cap, "cb_sharded_search: k too large for the mailbox");
    if (!q_host) {
        DeviceGuard g(root->device);
        CB_CUDA(cudaEventRecord(S->ev_q, s0));
    }
    for (int64_t q0 = 0; q0 < nq; q0 += chunk) {
        const int64_t m = std::min(chunk, nq - q0);
        int slot0 = 0;
        uint32_t total0 = 0;
        for (int r = 0; r < R; r++) {
            cb_index *ix = S->sh[r];
            DeviceGuard g(ix->device);
            cudaStream_t sr = r == 0 ? s0 : ix->stream;
            const float *q = nullptr;
            if (q_host) {
                CB_CUDA(cudaMemcpyAsync(ix->d_q, q_host + q0 * kD, (size_t)m * kD * 4, cudaMemcpyHostToDevice, sr));
                q = ix->d_q;
            } else if (r == 0) {
                q = q_dev0 + q0 * kD;
            } else {
                if (q0 == 0) CB_CUDA(cudaStreamWaitEvent(sr, S->ev_q, 0));
                CB_CUDA(cudaMemcpyPeerAsync(ix->d_q, ix->device, q_dev0 + q0 * kD, root->device, (size_t)m * kD * 4, sr));
                q = ix->d_q;
            }
            int slot = 0;
            uint32_t total = 0;
            rc = p2p_local_search(ix, m, q, k, shard_id_base(S, r), sr, &slot, &total);
            if (rc) return rc;
            if (r == 0) { slot0 = slot; total0 = total; }
            else CB_CUDA(cudaEventRecord(S->ev_done[r], sr));
        }
        DeviceGuard g(root->device);
        // one process: order the merge after every shard's kernels with events as well, so the merge
        // blocks never spin beside kernels that still wait for an SM (shards may share a device)
        for (int r = 1; r < R; r++) CB_CUDA(cudaStreamWaitEvent(s0, S->ev_done[r], 0));
        if ((rc = p2p_root_merge(root, slot0, total0, m, k, D_dev0 + q0 * k, I_dev0 + q0 * k, s0))) return rc;
    }
    return CB_OK;
}

}  // namespace

extern "C" {

int cb_sharded_create(int d, int storage_dtype, int ndev, const int *devices, cb_sharded **out) {
    CB_REQUIRE(out != nullptr, "cb_sharded_create: out is null");
    *out = nullptr;
    CB_REQUIRE(ndev >= 1 && ndev <= kMaxRanks && devices, "cb_sharded_create: 1..%d devices", kMaxRanks);
    cb_sharded *S = new (std::nothrow) cb_sharded();
    if (!S) { set_error("out of host memory"); return CB_ERR_OOM; }
    S->d = d;
    S->dtype = storage_dtype;
    for (int r = 0; r < ndev; r++) {
        cb_index *ix = nullptr;
        int rc = cb_flatip_create(d, storage_dtype, devices[r], &ix);
        if (rc) { cb_sharded_free(S); return rc; }
        S->sh.push_back(ix);
        S->devices.push_back(devices[r]);
    }
    S->segs.resize(ndev);
    S->ev_done.assign(ndev, nullptr);
    // every shard's device must be able to store into the root device's memory
    for (int r = 0; r < ndev; r++) {
        DeviceGuard g(devices[r]);
        cudaError_t e = cudaEventCreateWithFlags(&S->ev_done[r], cudaEventDisableTiming);
        if (e == cudaSuccess && r == 0) e = cudaEventCreateWithFlags(&S->ev_q, cudaEventDisableTiming);
        if (e == cudaSuccess && devices[r] != devices[0]) {
            int can = 0;
            e = cudaDeviceCanAccessPeer(&can, devices[r], devices[0]);
            if (e == cudaSuccess && !can) {
                set_error("cb_sharded_create: device %d cannot access device %d's memory (no NVLink/P2P path)", devices[r], devices[0]);
                cb_sharded_free(S);
                return CB_ERR_CUDA;
            }
            if (e == cudaSuccess) {
                e = cudaDeviceEnablePeerAccess(devices[0], 0);
                if (e == cudaErrorPeerAccessAlreadyEnabled) { cudaGetLastError(); e = cudaSuccess; }
            }
        }
        if (e != cudaSuccess) {
            set_error("cb_sharded_create: %s", cudaGetErrorString(e));
            cb_sharded_free(S);
            return CB_ERR_CUDA;
        }
    }
    *out = S;
    return CB_OK;
}

void cb_sharded_free(cb_sharded *S) {
    if (!S) return;
    // shards 1.. point into the root's mailbox: drop those references before the root frees it
    for (size_t r = 1; r < S->sh.size(); r++)
        if (S->sh[r]) { DeviceGuard g(S->sh[r]->device); cudaStreamSynchronize(S->sh[r]->stream); S->sh[r]->p2p = P2P(); }
    for (size_t r = 0; r < S->sh.size(); r++) {
        if (r < S->ev_done.size() && S->ev_done[r]) { DeviceGuard g(S->devices[r]); cudaEventDestroy(S->ev_done[r]); }
    }
    if (S->ev_q) { DeviceGuard g(S->devices[0]); cudaEventDestroy(S->ev_q); }
    for (cb_index *ix : S->sh) cb_flatip_free(ix);
    delete S;
}

int64_t cb_sharded_ntotal(const cb_sharded *S) { return S ? S->ntotal : -1; }
int cb_sharded_num_shards(const cb_sharded *S) { return S ? (int)S->sh.size() : -1; }
cb_index *cb_sharded_shard(cb_sharded *S, int r) { return (S && r >= 0 && r < (int)S->sh.size()) ? S->sh[r] : nullptr; }

int cb_sharded_reserve(cb_sharded *S, int64_t n_rows_total) {
    CB_REQUIRE(S != nullptr, "cb_sharded_reserve: null index");
    const int64_t R = (int64_t)S->sh.size(), per = (n_rows_total + R - 1) / R;
    for (cb_index *ix : S->sh) {
        int rc = cb_flatip_reserve(ix, per);
        if (rc) return rc;
    }
    return CB_OK;
}

int cb_sharded_reset(cb_sharded *S) {
    CB_REQUIRE(S != nullptr, "cb_sharded_reset: null index");
    for (size_t r = 0; r < S->sh.size(); r++) {
        int rc = cb_flatip_reset(S->sh[r]);
        if (rc) return rc;
        S->segs[r].clear();
    }
    S->ntotal = 0;
    return CB_OK;
}

int cb_sharded_add(cb_sharded *S, int64_t n, const float *x_host) {
    CB_REQUIRE(S != nullptr, "cb_sharded_add: null index");
    CB_REQUIRE(n >= 0, "cb_sharded_add: n < 0");
    if (n == 0) return CB_OK;
    CB_REQUIRE(x_host != nullptr, "cb_sharded_add: null rows");
    const int64_t R = (int64_t)S->sh.size(), per = (n + R - 1) / R;
    for (int r = 0; r < (int)R; r++) {
        const int64_t lo = std::min<int64_t>(r * per, n), hi = std::min<int64_t>((r + 1) * per, n);
        if (hi <= lo) continue;
        const int64_t local0 = S->sh[r]->ntotal;
        int rc = cb_flatip_add(S->sh[r], hi - lo, x_host + lo * kD);
        if (rc) return rc;
        sharded_note_rows(S, r, local0, S->ntotal + lo, hi - lo);
        if ((rc = sharded_sync_segments(S, r))) return rc;
    }
    S->ntotal += n;
    return CB_OK;
}

int cb_sharded_add_device(cb_sharded *S, int shard, int64_t n, const void *x_dev, int src_dtype, void *stream) {
    CB_REQUIRE(S != nullptr, "cb_sharded_add_device: null index");
    CB_REQUIRE(shard >= 0 && shard < (int)S->sh.size(), "cb_sharded_add_device: shard %d out of range", shard);
    if (n == 0) return CB_OK;
    const int64_t local0 = S->sh[shard]->ntotal;
    int rc = cb_flatip_add_device(S->sh[shard], n, x_dev, src_dtype, stream);
    if (rc) return rc;
    sharded_note_rows(S, shard, local0, S->ntotal, n);
    if ((rc = sharded_sync_segments(S, shard))) return rc;
    S->ntotal += n;
    return CB_OK;
}

int cb_sharded_search_device(cb_sharded *S, int64_t nq, const float *q_dev, int64_t k, float *D_dev,
                             int64_t *I_dev, void *stream) {
    CB_REQUIRE(S != nullptr, "cb_sharded_search_device: null index");
    CB_REQUIRE(nq >= 0, "cb_sharded_search_device: nq < 0");
    CB_REQUIRE(k > 0 && k < (1ll << 31), "cb_sharded_search_device: k must be in [1, 2^31) (got %lld)", (long long)k);
    if (nq == 0) return CB_OK;
    CB_REQUIRE(q_dev && D_dev && I_dev, "cb_sharded_search_device: null buffer");
    for (size_t r = 1; r < S->sh.size(); r++) {
        DeviceGuard g(S->sh[r]->device);
        int rc = ensure_q(S->sh[r], std::min<int64_t>(nq, 1024));
        if (rc) return rc;
    }
    return sharded_search_impl(S, nq, q_dev, nullptr, k, D_dev, I_dev, (cudaStream_t)stream);
}

int cb_sharded_search(cb_sharded *S, int64_t nq, const float *q_host, int64_t k, float *D_host, int64_t *I_host) {
    CB_REQUIRE(S != nullptr, "cb_sharded_search: null index");
    CB_REQUIRE(nq >= 0, "cb_sharded_search: nq < 0");
    CB_REQUIRE(k > 0 && k < (1ll << 31), "cb_sharded_search: k must be > 0 (got %lld)", (long long)k);
    if (nq == 0) return CB_OK;
    CB_REQUIRE(q_host && D_host && I_host, "cb_sharded_search: null buffer");
    cb_index *root = S->sh[0];
    if (S->sh.size() == 1) return cb_flatip_search(root, nq, q_host, k, D_host, I_host);
    int rc;
    for (cb_index *ix : S->sh) {
        DeviceGuard g(ix->device);
        if ((rc = ensure_q(ix, nq))) return rc;
    }
    DeviceGuard g(root->device);
    if ((rc = ensure_out(root, nq * k))) return rc;
    memcpy(root->h_q, q_host, (size_t)nq * kD * 4);           // one pinned copy feeds every shard's H2D
    rc = sharded_search_impl(S, nq, nullptr, root->h_q, k, root->d_D, root->d_I, root->stream);
    if (rc) return rc;
    CB_CUDA(cudaMemcpyAsync(root->h_D, root->d_D, (size_t)nq * k * 4, cudaMemcpyDeviceToHost, root->stream));
    CB_CUDA(cudaMemcpyAsync(root->h_I, root->d_I, (size_t)nq * k * 8, cudaMemcpyDeviceToHost, root->stream));
    CB_CUDA(cudaStreamSynchronize(root->stream));
    memcpy(D_host, root->h_D, (size_t)nq * k * 4);
    memcpy(I_host, root->h_I, (size_t)nq * k * 8);
    return CB_OK;
}

int cb_sharded_get_rows(cb_sharded *S, int64_t start, int64_t n, float *out_host) {
    CB_REQUIRE(S != nullptr, "cb_sharded_get_rows: null index");
    CB_REQUIRE(start >= 0 && n >= 0 && start + n <= S->ntotal, "cb_sharded_get_rows: range out of bounds");
    if (n == 0) return CB_OK;
    CB_REQUIRE(out_host != nullptr, "cb_sharded_get_rows: null buffer");
    for (size_t r = 0; r < S->sh.size(); r++)
        for (const auto &sg : S->segs[r]) {
            const int64_t lo = std::max(sg.global, start), hi = std::min(sg.global + sg.count, start + n);
            if (hi <= lo) continue;
            int rc = cb_flatip_get_rows(S->sh[r], sg.local + (lo - sg.global), hi - lo, out_host + (lo - start) * kD);
            if (rc) return rc;
        }
    return CB_OK;
}

}  // extern "C"
