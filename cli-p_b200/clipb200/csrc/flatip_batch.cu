// Exact inner-product top-k for QUERY BATCHES on the tensor cores (sm_100a).
//
// Replaces faiss.IndexFlatIP.search for nq > 2 -- the path faiss serves with blocked sgemm
// (reference call site /root/reference/query-index.py:111, BASELINE
// configs[2] "batch-1024 throughput" and configs[4]).
//
//   FILTER   S~[q, i] = <fp16(q), x_i>  as a tcgen05 GEMM: A = the query block rounded to fp16,
//            RESIDENT in shared memory for the whole launch (128 queries x 512 = 128 KB per CTA),
//            B = fp16 database rows streamed by TMA, fp32 accumulators in TMEM.  One MMA pass:
//            1024 FLOP per (query, row) issued, the algorithmic count.  The score matrix
//            (nq x N, 40 GB for 1024 x 10M) is never materialised: the epilogue compares each
//            S~ with thr[q] - margin[q] and appends survivors to a per-query candidate list.
//            thr is the exact k-th best so far; margin = (2^-11 + 2^-16) |q| max|x| + 1e-6 bounds
//            |<q - fp16(q), x>| plus the tensor core's accumulation error, so no row whose exact
//            score beats thr is ever dropped.
//   RE-SCORE between row ranges (they grow 1K, x16, x16 ... for k ~ 100, slower for larger k), one
//            block per query first orders its list by score LOWER bounds, drops what cannot reach
//            the top k any more, and computes the EXACT fp32 score of the ~k contenders
//            with the same summation order as the streaming scan (flatip.cuh): batch and scan
//            answers are bit-identical, scores included.  It then sorts by (score desc, id asc),
//            keeps k and raises thr.  Ranges ascend in id, so ties resolve exactly.
//   RESCUE   if a candidate list overflowed in a range (adversarial row order), the same block
//            re-reads that range and selects exactly on its own.  No host round trip anywhere: the
//            whole search is stream-ordered launches (CUDA-graph capturable).
//
// Kernel shapes: nq > 128: CTA pairs (cta_group::2), 256 queries x 256 rows per tile, each CTA
// streams half of the row tile; nq <= 128: single CTAs (cta_group::1), 128 x 256.  K = 512 in
// 8 k-blocks of 4 MMAs, TMA ring of 6 x 16 KB / 3 x 32 KB, double-buffered TMEM accumulators,
// 8 epilogue warps.  A cluster keeps ONE query block for the whole launch; clusters that hold
// different query blocks walk the row tiles in lockstep, so a row tile is fetched from HBM once
// and served to the other query blocks by the L2.
#include "common.cuh"
#include "flatip.cuh"
#include "tc_ptx.cuh"

#include <algorithm>
#include <mutex>

namespace cb {
namespace {

using namespace tc;

constexpr int QM = 128;                 // queries per CTA
constexpr int RN = 256;                 // database rows per tile
constexpr int BK = 64;
constexpr int KB = 512 / BK;
constexpr int kEpiWarps = 8;
constexpr int kThreads = 64 + kEpiWarps * 32;
constexpr int kCap = 8192;              // candidate slots per query between compactions
constexpr int kMaxGroup = 1024;         // queries per pass over the shard (4 query blocks of 256)

template <int NCTA>
struct BCfg {
    static constexpr int A_KB_BYTES = QM * BK * 2;             // one k-block of the query block: 16 KB
    static constexpr int A_BYTES = KB * A_KB_BYTES;            // 128 KB, resident
    static constexpr int B_ROWS = RN / NCTA;
    static constexpr int B_BYTES = B_ROWS * BK * 2;            // 16 KB (pair) / 32 KB (single CTA)
    static constexpr int STAGES = NCTA == 2 ? 6 : 3;
    static constexpr int SMEM_BYTES = A_BYTES + STAGES * B_BYTES + 256 + 1024;
    static_assert(SMEM_BYTES <= 232448, "shared memory budget");
};

template <int NCTA>
__global__ void __launch_bounds__(kThreads, 1)
flatip_batch_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmX,
                    const int64_t r0, const int64_t r1, const int nq, const int q_blocks,
                    const float *__restrict__ thr_eff, uint32_t *__restrict__ cnt, uint64_t *__restrict__ cand) {
    using C = BCfg<NCTA>;
    extern __shared__ uint8_t smem_raw[];
    uint8_t *smem = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    uint8_t *smem_b = smem + C::A_BYTES;
    uint64_t *full = reinterpret_cast<uint64_t *>(smem_b + C::STAGES * C::B_BYTES);
    uint64_t *empty = full + C::STAGES;
    uint64_t *tfull = empty + C::STAGES;
    uint64_t *tempty = tfull + 2;
    uint64_t *afull = tempty + 2;
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(afull + 1);

    const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;   // warp-uniform
    const uint32_t cta_rank = NCTA == 1 ? 0u : cluster_ctarank();
    const bool leader = cta_rank == 0;
    const int cluster = blockIdx.x / NCTA, num_clusters = gridDim.x / NCTA;
    // this cluster's query block and its share of the row tiles
    const int qb = cluster % q_blocks;
    const int grp = cluster / q_blocks, groups = num_clusters / q_blocks;
    const int64_t tiles = (r1 - r0 + RN - 1) / RN;
    const int qrow = qb * NCTA * QM + (int)cta_rank * QM;

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tmQ);
        tma_prefetch_desc(&tmX);
    }
    if (warp == 1) {
        if (lane == 0) {
            for (int i = 0; i < C::STAGES; i++) { mbar_init(&full[i], 1); mbar_init(&empty[i], 1); }
            for (int i = 0; i < 2; i++) { mbar_init(&tfull[i], 1); mbar_init(&tempty[i], kEpiWarps * NCTA); }
            mbar_init(afull, 1);
            fence_barrier_init();
        }
        __syncwarp();
        tmem_alloc<NCTA>(tmem_slot, 512);
    }
    tc_fence_before();
    __syncthreads();
    if (NCTA > 1) cluster_sync_all();
    tc_fence_after();
    const uint32_t tmem_base = *reinterpret_cast<volatile uint32_t *>(tmem_slot);

    // the TMA-issue and MMA-issue roles are warp-uniform loops with the issuing instructions behind
    // elect.sync (as in gemm.cu): descriptors and coordinates stay in uniform registers
    if (warp == 0) {
        // the query block, once: 8 k-blocks of [128 queries x 64]
        if (elect_one()) {
            if (leader) mbar_expect_tx(afull, NCTA * C::A_BYTES);
            for (int kb = 0; kb < KB; kb++) {
                if (NCTA == 1) tma_load_2d(smem + kb * C::A_KB_BYTES, &tmQ, kb * BK, qrow, afull);
                else tma_load_2d_2sm(smem + kb * C::A_KB_BYTES, &tmQ, kb * BK, qrow, afull);
            }
        }
        __syncwarp();
        uint32_t stage = 0, phase = 0;
        for (int64_t t = grp; t < tiles; t += groups) {
            const int64_t n0 = r0 + t * RN;
            for (int kb = 0; kb < KB; kb++) {
                mbar_wait(&empty[stage], phase ^ 1);
                uint8_t *s = smem_b + stage * C::B_BYTES;
                if (elect_one()) {
                    if (leader) mbar_expect_tx(&full[stage], NCTA * C::B_BYTES);
                    if (NCTA == 1) tma_load_2d(s, &tmX, kb * BK, (int)n0, &full[stage]);
                    else tma_load_2d_2sm(s, &tmX, kb * BK, (int)(n0 + cta_rank * C::B_ROWS), &full[stage]);
                }
                __syncwarp();
                if (++stage == (uint32_t)C::STAGES) { stage = 0; phase ^= 1; }
            }
        }
    } else if (warp == 1) {
        if (leader) {
            constexpr uint32_t idesc = make_idesc(NCTA * QM, RN);
            uint32_t stage = 0, phase = 0, as = 0, aphase = 0;
            mbar_wait(afull, 0);
            tc_fence_after();
            const uint32_t sa = smem_u32(smem);
            for (int64_t t = grp; t < tiles; t += groups) {
                mbar_wait(&tempty[as], aphase ^ 1);
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + as * RN;
                for (int kb = 0; kb < KB; kb++) {
                    mbar_wait(&full[stage], phase);
                    tc_fence_after();
                    const uint64_t da = make_smem_desc(sa + kb * C::A_KB_BYTES);
                    const uint64_t db = make_smem_desc(smem_u32(smem_b + stage * C::B_BYTES));
                    if (elect_one()) {
#pragma unroll
                        for (int k = 0; k < BK / 16; k++) umma_f16<NCTA>(d_tmem, da + 2 * k, db + 2 * k, idesc, (kb | k) != 0 ? 1u : 0u);
                        umma_commit<NCTA>(&empty[stage]);
                    }
                    __syncwarp();
                    if (++stage == (uint32_t)C::STAGES) { stage = 0; phase ^= 1; }
                }
                if (elect_one()) umma_commit<NCTA>(&tfull[as]);
                __syncwarp();
                as ^= 1;
                if (as == 0) aphase ^= 1;
            }
        }
    } else {
        const int q = warp & 3, half = (warp - 2) >> 2;
        uint32_t as = 0, aphase = 0;
        const int query = qrow + q * 32 + lane;
        const bool q_ok = query < nq;
        const float my_thr = q_ok ? __ldcg(thr_eff + query) : INFINITY;
        uint64_t *my_cand = cand + (size_t)query * kCap;
        for (int64_t t = grp; t < tiles; t += groups) {
            const int64_t n0 = r0 + t * RN;
            mbar_wait(&tfull[as], aphase);
            tc_fence_after();
#pragma unroll 1
            for (int c = 0; c < (RN / 2) / 32; c++) {
                uint32_t v[32];
                const int col0 = half * (RN / 2) + c * 32;
                tmem_ld_32x32(tmem_base + ((uint32_t)(q * 32) << 16) + as * RN + col0, v);
                tmem_ld_wait();
                const int64_t row_base = n0 + col0;
                uint32_t pass = 0;
#pragma unroll
                for (int j = 0; j < 32; j++) {
                    const bool p = (__uint_as_float(v[j]) > my_thr) && (row_base + j < r1);
                    pass |= (p ? 1u : 0u) << j;
                }
                if (pass) {
                    const uint32_t npass = __popc(pass);
                    uint32_t pos = atomicAdd(cnt + query, npass);      // keeps counting past kCap: overflow = cnt > kCap
#pragma unroll
                    for (int j = 0; j < 32; j++) {
                        if (pass & (1u << j)) {
                            if (pos < (uint32_t)kCap)
                                my_cand[pos] = make_comp(f2key(__uint_as_float(v[j])), (uint32_t)(row_base + j));
                            pos++;
                        }
                    }
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) {
                if (NCTA == 1) mbar_arrive_relaxed(&tempty[as]);
                else mbar_arrive_cluster(&tempty[as], 0);
            }
            as ^= 1;
            if (as == 0) aphase ^= 1;
        }
    }

    tc_fence_before();
    __syncthreads();
    if (NCTA > 1) cluster_sync_all();
    if (warp == 1) {
        __syncwarp();
        tmem_dealloc<NCTA>(tmem_base, 512);
    }
}

// per query (one warp each): fp16 copy of the query for the tensor cores, the filter margin, and a
// fresh selection state.  Rows >= nq of the padded block are zero and never selected.
__global__ void batch_prep_kernel(const float *__restrict__ q, int nq, int nq_pad, __half *__restrict__ qh,
                                  float *__restrict__ thr, float *__restrict__ thr_eff, float *__restrict__ margin,
                                  uint32_t *__restrict__ cnt, uint32_t *__restrict__ kept, uint32_t *__restrict__ bad,
                                  const float *__restrict__ max_norm2) {
    const int lane = threadIdx.x & 31;
    const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (row >= nq_pad) return;
    float n2 = 0.f, amax = 0.f;
#pragma unroll
    for (int c = 0; c < 4; c++) {
        const int col = (c * 32 + lane) * 4;
        float4 v = row < nq ? *reinterpret_cast<const float4 *>(q + (size_t)row * kD + col) : make_float4(0, 0, 0, 0);
        amax = fmaxf(amax, fmaxf(fmaxf(fabsf(v.x), fabsf(v.y)), fmaxf(fabsf(v.z), fabsf(v.w))));
        n2 = fmaf(v.x, v.x, n2); n2 = fmaf(v.y, v.y, n2); n2 = fmaf(v.z, v.z, n2); n2 = fmaf(v.w, v.w, n2);
        __half2 a = __floats2half2_rn(v.x, v.y), b = __floats2half2_rn(v.z, v.w);
        uint2 o;
        o.x = *reinterpret_cast<uint32_t *>(&a);
        o.y = *reinterpret_cast<uint32_t *>(&b);
        *reinterpret_cast<uint2 *>(qh + (size_t)row * kD + col) = o;
    }
    n2 = warp_sum(n2);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) amax = fmaxf(amax, __shfl_xor_sync(0xffffffffu, amax, o));
    if (lane == 0) {
        // a query outside fp16's range (or not finite) cannot be filtered on the tensor cores: its block
        // selects exactly from the rows themselves (the rescue path) in every range
        bad[row] = (amax < 60000.f && n2 < 3.0e38f) ? 0u : 1u;
        // |<q - fp16(q), x>| <= 2^-11 |q| |x| (normal range; subnormal elements add < 1e-6) and the tensor
        // core's fp32 accumulation of 512 products is within 2^-16 |q| |x|
        margin[row] = (4.8828125e-4f + 1.52587890625e-5f) * sqrtf(n2) * sqrtf(__ldg(max_norm2)) + 1e-6f;
        thr[row] = -INFINITY;
        thr_eff[row] = -INFINITY;
        cnt[row] = 0;
        kept[row] = 0;
    }
}

// descending bitonic sort of p2 composite keys in shared memory by the whole block
__device__ void sort_desc(uint64_t *a, uint32_t p2) {
    for (uint32_t kk = 2; kk <= p2; kk <<= 1) {
        for (uint32_t j = kk >> 1; j > 0; j >>= 1) {
            for (uint32_t t = threadIdx.x; t < p2 / 2; t += blockDim.x) {
                uint32_t i = ((t & ~(j - 1)) << 1) | (t & (j - 1));
                uint32_t l = i | j;
                bool desc = (i & kk) == 0;
                uint64_t x = a[i], y = a[l];
                if ((x < y) == desc) { a[i] = y; a[l] = x; }
            }
            __syncthreads();
        }
    }
}
__device__ __forceinline__ uint32_t next_pow2(uint32_t c) {
    uint32_t p2 = 1;
    while (p2 < c) p2 <<= 1;
    return p2;
}

// One block per query, after each row range [r0, r1): exact re-score of the new candidates (or an
// exact re-read of the range if the list overflowed), sort, keep k, raise the threshold; on the
// final range also write D / I (faiss padding) -- possibly into the root GPU's mailbox.
// The usual list (k kept + a few hundred new) is handled in 16 KB of shared memory, so many query
// blocks share an SM; longer lists and the rescue work in the query's global candidate buffer.
constexpr int kSmemCap = 2048;

template <int THREADS>
__global__ void __launch_bounds__(THREADS)
batch_compact_kernel(uint64_t *__restrict__ cand, uint32_t *__restrict__ cnt, uint32_t *__restrict__ kept,
                     float *__restrict__ thr, float *__restrict__ thr_eff, const float *__restrict__ margin,
                     const uint32_t *__restrict__ bad, const uint4 *__restrict__ rows, const float *__restrict__ xq,
                     int64_t k, int64_t r0, int64_t r1, int final_pass, const IdMap ids, const PeerOut po,
                     float *__restrict__ D, int64_t *__restrict__ I, unsigned long long *rescued) {
    __shared__ uint64_t s_c[kSmemCap];
    __shared__ uint32_t s_n;
    const int q = blockIdx.x;
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    constexpr int nwarps = THREADS / 32;
    uint64_t *mine = cand + (size_t)q * kCap;
    const uint32_t c_raw = cnt[q], kp = kept[q];
    const bool rescue = c_raw > (uint32_t)kCap || bad[q] != 0;
    if (!final_pass && !rescue && c_raw == kp) return;  // nothing new in this range
    float qr[16];
    load_q_lane_f16(xq + (size_t)q * kD, lane, qr);
    const bool in_smem = !rescue && c_raw <= (uint32_t)kSmemCap;
    uint64_t *buf = in_smem ? s_c : mine;              // entries [0, kp) of `mine` are the kept (exact, sorted) ones
    if (in_smem)
        for (uint32_t i = threadIdx.x; i < kp; i += THREADS) s_c[i] = mine[i];
    uint32_t c;
    if (rescue) {
        // RESCUE: survivors of this range were dropped (or the query cannot be filtered in fp16).
        // Select exactly over [r0, r1) here, in the global buffer.
        if (threadIdx.x == 0) { s_n = kp; atomicAdd(rescued, 1ull); }
        __syncthreads();
        float cur = kp == (uint32_t)k ? thr[q] : -INFINITY;
        constexpr int64_t kBatch = 2048;
        for (int64_t b0 = r0; b0 < r1; b0 += kBatch) {
            if (s_n + kBatch > (uint32_t)kCap) {
                // make room: sort, keep k, raise the running threshold (uniform branch: s_n is shared)
                const uint32_t n_now = s_n, p2 = next_pow2(n_now);
                for (uint32_t i = n_now + threadIdx.x; i < p2; i += THREADS) buf[i] = 0ull;
                __syncthreads();
                sort_desc(buf, p2);
                const uint32_t keep = min(n_now, (uint32_t)k);
                if (keep == (uint32_t)k) cur = key2f((uint32_t)(buf[k - 1] >> 32));
                __syncthreads();
                if (threadIdx.x == 0) s_n = keep;
                __syncthreads();
            }
            const int64_t b1 = min(b0 + kBatch, r1);
            for (int64_t r = b0 + wid; r < b1; r += nwarps) {
                const float sc = exact_score_f16(rows + r * 64, qr, lane);
                // rows ascend in id: an equal score later in the shard loses the tie
                if (lane == 0 && sc > cur) buf[atomicAdd(&s_n, 1u)] = make_comp(f2key(sc), (uint32_t)r);
            }
            __syncthreads();
        }
        c = s_n;
    } else {
        c = c_raw;
        // New candidates carry the tensor cores' score s~ of the fp16-rounded query: the exact score lies in
        // [s~ - m, s~ + m].  Before any row is fetched, order everything by its LOWER bound (exact score for
        // the kept entries, s~ - m for the new ones): the k-th largest lower bound T cannot exceed the k-th
        // largest exact score, so only entries whose UPPER bound reaches T can still make the top k.  That is
        // ~k entries instead of the ~(g - 1) k that passed the stale threshold: the gather from HBM shrinks by
        // the range growth factor g.
        const float m = margin[q];
        for (uint32_t i = kp + threadIdx.x; i < c; i += THREADS) {
            const uint64_t e = mine[i];
            buf[i] = ((uint64_t)f2key(key2f((uint32_t)(e >> 32)) - m) << 32) | (e & 0xffffffffull);
        }
        __syncthreads();
        uint32_t P = c;
        if (c > (uint32_t)k) {
            const uint32_t p2a = next_pow2(c);
            for (uint32_t i = c + threadIdx.x; i < p2a; i += THREADS) buf[i] = 0ull;
            __syncthreads();
            sort_desc(buf, p2a);
            const float cut = key2f((uint32_t)(buf[k - 1] >> 32)) - 2.0f * m;
            if (threadIdx.x == 0) s_n = c;
            __syncthreads();
            // entries are sorted by lower bound: the contenders are a prefix
            for (uint32_t i = (uint32_t)k + threadIdx.x; i < c; i += THREADS)
                if (key2f((uint32_t)(buf[i] >> 32)) < cut && !(key2f((uint32_t)(buf[i - 1] >> 32)) < cut)) s_n = i;
            __syncthreads();
            P = s_n;
        }
        // exact fp32 score of the contenders that are new (their rows lie in this range): one warp per
        // row, four rows in flight per warp
        for (uint32_t i0 = 4 * wid; i0 < P; i0 += 4 * nwarps) {
            uint32_t id[4];
            bool isnew[4];
            uint4 v[4][2];
#pragma unroll
            for (int j = 0; j < 4; j++) {
                const uint32_t i = min(i0 + j, P - 1);
                id[j] = 0xffffffffu - (uint32_t)buf[i];
                isnew[j] = (int64_t)id[j] >= r0 && i0 + j < P;
                if (isnew[j]) {
                    const uint4 *p = rows + (size_t)id[j] * 64;
                    v[j][0] = ld_stream_v4(p + lane);
                    v[j][1] = ld_stream_v4(p + lane + 32);
                }
            }
#pragma unroll
            for (int j = 0; j < 4; j++) {
                if (!isnew[j]) continue;                   // warp-uniform
                float t = 0.f;
                t += dot8_h(v[j][0], &qr[0]);
                t += dot8_h(v[j][1], &qr[8]);
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
                if (lane == 0) buf[i0 + j] = make_comp(f2key(t + 0.0f), id[j]);
            }
        }
        __syncthreads();
        c = P;
    }
    const uint32_t keep = min(c, (uint32_t)k);
    {
        // every entry of buf[0, c) now carries its exact score: final order, and the kept list goes back
        // to the query's global buffer (new entries may have displaced old ones even when c == kp)
        const uint32_t p2 = next_pow2(c);
        for (uint32_t i = c + threadIdx.x; i < p2; i += THREADS) buf[i] = 0ull;
        __syncthreads();
        sort_desc(buf, p2);
        if (in_smem)
            for (uint32_t i = threadIdx.x; i < keep; i += THREADS) mine[i] = s_c[i];
    }
    if (threadIdx.x == 0) {
        cnt[q] = keep;
        kept[q] = keep;
        if (keep == (uint32_t)k) {
            const float t = key2f((uint32_t)(buf[k - 1] >> 32));
            thr[q] = t;
            thr_eff[q] = t - margin[q];
        }
    }
    if (final_pass) {
        peer_wait_slot(po);
        for (int64_t j = threadIdx.x; j < k; j += THREADS) {
            if (j < keep) {
                const uint64_t e = buf[j];
                D[(size_t)q * k + j] = key2f((uint32_t)(e >> 32));
                I[(size_t)q * k + j] = map_id(ids, 0xffffffffu - (uint32_t)e);
            } else {
                D[(size_t)q * k + j] = -3.4028234663852886e38f;
                I[(size_t)q * k + j] = -1;
            }
        }
        peer_signal(po, 1u);
    }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                  const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int make_map(CUtensorMap *m, const void *base, uint64_t rows, uint64_t cols, uint32_t box_rows) {
    static EncodeTiledFn fn = nullptr;
    static std::once_flag once;
    std::call_once(once, [] {
        void *p = nullptr;
        cudaDriverEntryPointQueryResult qr;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qr) == cudaSuccess &&
            qr == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(p);
    });
    if (!fn) { set_error("cuTensorMapEncodeTiled entry point not available"); return CB_ERR_CUDA; }
    cuuint64_t gdim[2] = {cols, rows};
    cuuint64_t gstride[1] = {cols * 2};
    cuuint32_t box[2] = {(cuuint32_t)BK, box_rows};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = fn(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, const_cast<void *>(base), gdim, gstride, box, estr,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled failed (%d)", (int)r); return CB_ERR_CUDA; }
    return CB_OK;
}

}  // namespace

// workspace owned by the index (see flatip.cu)
struct BatchWs {
    __half *qh = nullptr;
    float *thr = nullptr, *thr_eff = nullptr, *margin = nullptr;
    uint32_t *cnt = nullptr, *kept = nullptr, *bad = nullptr;
    uint64_t *cand = nullptr;
    unsigned long long *rescued = nullptr;
    int nq_pad_cap = 0;
};

int batch_ws_ensure(BatchWs *w, int nq_pad) {
    if (!w->rescued) {
        CB_CUDA(cudaMalloc(&w->rescued, 8));
        CB_CUDA(cudaMemset(w->rescued, 0, 8));
    }
    if (nq_pad <= w->nq_pad_cap) return CB_OK;
    cudaFree(w->qh); cudaFree(w->thr); cudaFree(w->thr_eff); cudaFree(w->margin);
    cudaFree(w->cnt); cudaFree(w->kept); cudaFree(w->bad); cudaFree(w->cand);
    w->qh = nullptr; w->thr = w->thr_eff = w->margin = nullptr; w->cnt = w->kept = w->bad = nullptr; w->cand = nullptr;
    w->nq_pad_cap = 0;
    CB_CUDA(cudaMalloc(&w->qh, (size_t)nq_pad * kD * 2));
    CB_CUDA(cudaMalloc(&w->thr, (size_t)nq_pad * 4));
    CB_CUDA(cudaMalloc(&w->thr_eff, (size_t)nq_pad * 4));
    CB_CUDA(cudaMalloc(&w->margin, (size_t)nq_pad * 4));
    CB_CUDA(cudaMalloc(&w->cnt, (size_t)nq_pad * 4));
    CB_CUDA(cudaMalloc(&w->kept, (size_t)nq_pad * 4));
    CB_CUDA(cudaMalloc(&w->bad, (size_t)nq_pad * 4));
    CB_CUDA(cudaMalloc(&w->cand, (size_t)nq_pad * kCap * 8));
    w->nq_pad_cap = nq_pad;
    return CB_OK;
}

void batch_ws_free(BatchWs *w) {
    cudaFree(w->qh); cudaFree(w->thr); cudaFree(w->thr_eff); cudaFree(w->margin);
    cudaFree(w->cnt); cudaFree(w->kept); cudaFree(w->bad); cudaFree(w->cand); cudaFree(w->rescued);
    *w = BatchWs();
}

BatchWs *batch_ws_new() { return new BatchWs(); }
void batch_ws_delete(BatchWs *w) { if (w) { batch_ws_free(w); delete w; } }

// (query, range) pairs that overflowed their candidate list and were re-selected exactly; synchronises `s`
int batch_ws_stats(BatchWs *w, int64_t *rescued, cudaStream_t s) {
    *rescued = 0;
    if (!w->rescued) return CB_OK;
    unsigned long long v = 0;
    CB_CUDA(cudaMemcpyAsync(&v, w->rescued, 8, cudaMemcpyDeviceToHost, s));
    CB_CUDA(cudaStreamSynchronize(s));
    *rescued = (int64_t)v;
    return CB_OK;
}

template <int NCTA>
static int launch_range(const CUtensorMap &tmQ, const CUtensorMap &tmX, int64_t r0, int64_t r1, int nq, int q_blocks,
                        int sms, BatchWs *w, cudaStream_t s) {
    using C = BCfg<NCTA>;
    const int64_t tiles = (r1 - r0 + RN - 1) / RN;
    // every cluster keeps one query block: the cluster count is a multiple of q_blocks
    int clusters = (int)std::min<int64_t>(tiles * q_blocks, sms / NCTA);
    clusters = std::max(q_blocks, clusters / q_blocks * q_blocks);
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(clusters * NCTA);
    cfg.blockDim = dim3(kThreads);
    cfg.dynamicSmemBytes = C::SMEM_BYTES;
    cfg.stream = s;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = NCTA; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    CB_CUDA(cudaLaunchKernelEx(&cfg, flatip_batch_kernel<NCTA>, tmQ, tmX, r0, r1, nq, q_blocks,
                               (const float *)w->thr_eff, w->cnt, w->cand));
    CB_LAUNCH_CHECK();
    return CB_OK;
}

// one pass sequence over the shard for <= kMaxGroup queries; nothing synchronises
static int search_group(BatchWs *w, const void *rows_f16, int64_t n, int sms, int nq, const float *q_dev, int64_t k,
                        float *D_dev, int64_t *I_dev, const IdMap &ids, const PeerOut &po, const float *max_norm2,
                        cudaStream_t s) {
    const int ncta = nq > QM ? 2 : 1;
    const int qblk = ncta * QM;
    const int nq_pad = (nq + qblk - 1) / qblk * qblk;
    const int q_blocks = nq_pad / qblk;
    batch_prep_kernel<<<(nq_pad + 7) / 8, 256, 0, s>>>(q_dev, nq, nq_pad, w->qh, w->thr, w->thr_eff, w->margin, w->cnt,
                                                      w->kept, w->bad, max_norm2);
    CB_LAUNCH_CHECK();
    CUtensorMap tmQ, tmX;
    int rc;
    if ((rc = make_map(&tmQ, w->qh, (uint64_t)nq_pad, kD, QM))) return rc;
    if ((rc = make_map(&tmX, rows_f16, (uint64_t)n, kD, RN / ncta))) return rc;
    // row ranges grow geometrically (x g): the expected survivors of a range are k m / n_seen ~ (g - 1) k, and
    // together with the k kept entries they should fit the compaction's shared-memory list: g = 16 up to
    // k = 113, 6 at k = 256, 2 from k = 512.  Few queries can afford longer ranges (fewer launches on the
    // HBM-bound small-nq pass)
    const int64_t growth = std::max<int64_t>(2, std::min<int64_t>(16, kSmemCap / k - 2));
    const int64_t span_max = nq <= QM ? (16ll << 20) : (4ll << 20);
    int64_t r0 = 0, span = 1024;
    while (r0 < n) {
        const int64_t r1 = std::min(n, r0 + span);
        rc = ncta == 2 ? launch_range<2>(tmQ, tmX, r0, r1, nq, q_blocks, sms, w, s)
                       : launch_range<1>(tmQ, tmX, r0, r1, nq, q_blocks, sms, w, s);
        if (rc) return rc;
        const int final_pass = r1 >= n ? 1 : 0;
        // few queries: one wave of big blocks (the re-score is a chain of dependent row gathers per warp);
        // many queries: small blocks, eight per SM
        if (nq <= 2 * sms)
            batch_compact_kernel<512><<<(unsigned)nq, 512, 0, s>>>(
                w->cand, w->cnt, w->kept, w->thr, w->thr_eff, w->margin, w->bad, (const uint4 *)rows_f16, q_dev, k, r0,
                r1, final_pass, ids, po, D_dev, I_dev, w->rescued);
        else
            batch_compact_kernel<256><<<(unsigned)nq, 256, 0, s>>>(
                w->cand, w->cnt, w->kept, w->thr, w->thr_eff, w->margin, w->bad, (const uint4 *)rows_f16, q_dev, k, r0,
                r1, final_pass, ids, po, D_dev, I_dev, w->rescued);
        CB_LAUNCH_CHECK();
        r0 = r1;
        span = std::min<int64_t>(span * growth, span_max);
    }
    return CB_OK;
}

int flatip_search_batch(BatchWs *w, const void *rows_f16, int64_t n, int device, int64_t nq, const float *q_dev,
                        int64_t k, float *D_dev, int64_t *I_dev, const IdMap &ids, const PeerOut &po,
                        const float *max_norm2, cudaStream_t s) {
    int rc = batch_ws_ensure(w, (int)std::min<int64_t>((nq + 255) / 256 * 256, kMaxGroup));
    if (rc) return rc;
    static std::once_flag attr_once[64];       // function attributes are per device
    cudaError_t attr_err = cudaSuccess;
    std::call_once(attr_once[device & 63], [&] {
        attr_err = cudaFuncSetAttribute(flatip_batch_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, BCfg<1>::SMEM_BYTES);
        if (attr_err == cudaSuccess)
            attr_err = cudaFuncSetAttribute(flatip_batch_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, BCfg<2>::SMEM_BYTES);
    });
    CB_CUDA(attr_err);
    int sms = kNumSMs;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device);
    for (int64_t g0 = 0; g0 < nq; g0 += kMaxGroup) {
        const int m = (int)std::min<int64_t>(kMaxGroup, nq - g0);
        rc = search_group(w, rows_f16, n, sms, m, q_dev + g0 * kD, k, D_dev + g0 * k, I_dev + g0 * k, ids, po, max_norm2, s);
        if (rc) return rc;
    }
    return CB_OK;
}

}  // namespace cb
