// Exact inner-product top-k for QUERY BATCHES on the tensor cores (sm_100a).
//
// Replaces faiss.IndexFlatIP.search for large nq -- the path faiss serves with blocked
// sgemm (reference call site /root/reference/query-index.py:111,
// BASELINE configs[2] "batch-1024 throughput" and configs[4]).
//
//   S[q, i] = <q, x_i>  as a tcgen05 GEMM: A = queries, B = fp16 database rows (both K-major),
//   fp32 accumulators in TMEM.  Queries are fp32 in the reference; to stay exact they are
//   split q = hi + lo (two fp16 values, lo may be subnormal: |q - hi - lo| <= 2^-25 for unit
//   vectors) and both halves are multiplied against the SAME shared-memory tile of rows, so
//   the database is streamed once and score error stays ~1e-7 (well inside the 1e-5 rule).
//
//   The score matrix (nq x N, 40 GB for 1024 x 10M) is never materialised: the epilogue
//   compares each score with its query's running threshold (the k-th best so far) and
//   appends survivors to a per-query candidate list.  The shard is processed in row ranges
//   that grow geometrically (4K, +16K, +64K, ...); between ranges a compaction kernel sorts
//   each list, keeps the best k and raises the threshold, so the expected survivors per
//   range stay ~4k.  Row ranges are ascending in id, so `score > threshold` is also exact
//   for ties (a later row with an equal score has a higher id and loses).  If a list ever
//   overflows (adversarial ordering) the caller falls back to the streaming-scan path.
//
// Kernel shape: CTA pairs (cta_group::2), 256 queries x 256 rows per pair tile, K = 512 in
// 8 k-blocks, 2 x 4 MMAs per k-block (hi, lo), 4-stage TMA ring (A_hi, A_lo, B-half per CTA),
// double-buffered TMEM accumulators, 8 epilogue warps per CTA.
#include "common.cuh"
#include "tc_ptx.cuh"

#include <algorithm>

namespace cb {
namespace {

using namespace tc;

constexpr int QM = 128;                 // queries per CTA
constexpr int RN = 256;                 // database rows per pair tile
constexpr int BK = 64;
constexpr int KB = 512 / BK;
constexpr int kStages = 4;
constexpr int A_BYTES = QM * BK * 2;
constexpr int B_BYTES = (RN / 2) * BK * 2;
constexpr int STAGE_BYTES = 2 * A_BYTES + B_BYTES;
constexpr int kEpiWarps = 8;
constexpr int kThreads = 64 + kEpiWarps * 32;
constexpr int SMEM_BYTES = kStages * STAGE_BYTES + 256 + 1024;
constexpr int kCap = 8192;              // candidate slots per query between compactions

__global__ void __launch_bounds__(kThreads, 1)
flatip_batch_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmX,
                    const int64_t r0, const int64_t r1, const int nq, const int q_blocks,
                    const float *__restrict__ thr, uint32_t *__restrict__ cnt, uint64_t *__restrict__ cand,
                    int *__restrict__ overflow) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t *smem = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    uint64_t *full = reinterpret_cast<uint64_t *>(smem + kStages * STAGE_BYTES);
    uint64_t *empty = full + kStages;
    uint64_t *tfull = empty + kStages;
    uint64_t *tempty = tfull + 2;
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(tempty + 2);

    const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;   // warp-uniform
    const uint32_t cta_rank = cluster_ctarank();
    const bool leader = cta_rank == 0;
    const int pair = blockIdx.x >> 1, num_pairs = gridDim.x >> 1;

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tmQ);
        tma_prefetch_desc(&tmX);
    }
    if (warp == 1) {
        if (lane == 0) {
            for (int i = 0; i < kStages; i++) { mbar_init(&full[i], 1); mbar_init(&empty[i], 1); }
            for (int i = 0; i < 2; i++) { mbar_init(&tfull[i], 1); mbar_init(&tempty[i], kEpiWarps * 2); }
            fence_barrier_init();
        }
        __syncwarp();
        tmem_alloc<2>(tmem_slot, 512);
    }
    tc_fence_before();
    __syncthreads();
    cluster_sync_all();
    tc_fence_after();
    const uint32_t tmem_base = *reinterpret_cast<volatile uint32_t *>(tmem_slot);

    const int64_t tiles = (r1 - r0 + RN - 1) / RN;
    const int64_t items = tiles * q_blocks;          // item = (row tile, query block); query block fastest

    // the TMA-issue and MMA-issue roles are warp-uniform loops with the issuing instructions behind
    // elect.sync (as in gemm.cu): descriptors and coordinates stay in uniform registers
    if (warp == 0) {
        uint32_t stage = 0, phase = 0;
        for (int64_t it = pair; it < items; it += num_pairs) {
            const int qb = (int)(it % q_blocks);
            const int64_t n0 = r0 + (it / q_blocks) * RN;
            const int qrow = qb * 2 * QM + (int)cta_rank * QM;
            for (int kb = 0; kb < KB; kb++) {
                mbar_wait(&empty[stage], phase ^ 1);
                uint8_t *s = smem + stage * STAGE_BYTES;
                if (elect_one()) {
                    if (leader) mbar_expect_tx(&full[stage], 2 * STAGE_BYTES);
                    tma_load_2d_2sm(s, &tmQ, kb * BK, qrow, &full[stage]);                    // hi half
                    tma_load_2d_2sm(s + A_BYTES, &tmQ, 512 + kb * BK, qrow, &full[stage]);    // lo half
                    tma_load_2d_2sm(s + 2 * A_BYTES, &tmX, kb * BK, (int)(n0 + cta_rank * (RN / 2)), &full[stage]);
                }
                __syncwarp();
                if (++stage == kStages) { stage = 0; phase ^= 1; }
            }
        }
    } else if (warp == 1) {
        if (leader) {
            constexpr uint32_t idesc = make_idesc(2 * QM, RN);
            uint32_t stage = 0, phase = 0, as = 0, aphase = 0;
            for (int64_t it = pair; it < items; it += num_pairs) {
                mbar_wait(&tempty[as], aphase ^ 1);
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + as * RN;
                for (int kb = 0; kb < KB; kb++) {
                    mbar_wait(&full[stage], phase);
                    tc_fence_after();
                    const uint32_t s = smem_u32(smem + stage * STAGE_BYTES);
                    const uint64_t dh = make_smem_desc(s), dl = make_smem_desc(s + A_BYTES);
                    const uint64_t db = make_smem_desc(s + 2 * A_BYTES);
                    if (elect_one()) {
#pragma unroll
                        for (int k = 0; k < BK / 16; k++) umma_f16<2>(d_tmem, dh + 2 * k, db + 2 * k, idesc, (kb | k) != 0 ? 1u : 0u);
#pragma unroll
                        for (int k = 0; k < BK / 16; k++) umma_f16<2>(d_tmem, dl + 2 * k, db + 2 * k, idesc, 1u);
                        umma_commit<2>(&empty[stage]);
                    }
                    __syncwarp();
                    if (++stage == kStages) { stage = 0; phase ^= 1; }
                }
                if (elect_one()) umma_commit<2>(&tfull[as]);
                __syncwarp();
                as ^= 1;
                if (as == 0) aphase ^= 1;
            }
        }
    } else {
        const int q = warp & 3, half = (warp - 2) >> 2;
        uint32_t as = 0, aphase = 0;
        for (int64_t it = pair; it < items; it += num_pairs) {
            const int qb = (int)(it % q_blocks);
            const int64_t n0 = r0 + (it / q_blocks) * RN;
            const int query = qb * 2 * QM + (int)cta_rank * QM + q * 32 + lane;
            const bool q_ok = query < nq;
            const float my_thr = q_ok ? __ldcg(thr + query) : INFINITY;
            uint64_t *my_cand = cand + (size_t)query * kCap;
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
                    uint32_t pos = atomicAdd(cnt + query, npass);
#pragma unroll
                    for (int j = 0; j < 32; j++) {
                        if (pass & (1u << j)) {
                            if (pos < (uint32_t)kCap)
                                my_cand[pos] = ((uint64_t)f2key(__uint_as_float(v[j])) << 32) |
                                               (uint64_t)(0xffffffffu - (uint32_t)(row_base + j));
                            else
                                *overflow = 1;
                            pos++;
                        }
                    }
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive_cluster(&tempty[as], 0);
            as ^= 1;
            if (as == 0) aphase ^= 1;
        }
    }

    tc_fence_before();
    __syncthreads();
    cluster_sync_all();
    if (warp == 1) {
        __syncwarp();
        tmem_dealloc<2>(tmem_base, 512);
    }
}

// fp32 queries -> [nq_pad, 1024] fp16 (hi | lo); resets the per-query selection state
__global__ void batch_prep_kernel(const float *__restrict__ q, int nq, int nq_pad, __half *__restrict__ qh,
                                  float *__restrict__ thr, uint32_t *__restrict__ cnt, int *overflow) {
    const int64_t total = (int64_t)nq_pad * 512;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int row = (int)(i >> 9), col = (int)(i & 511);
        float v = row < nq ? q[i] : 0.f;
        __half hi = __float2half_rn(v);
        __half lo = __float2half_rn(v - __half2float(hi));
        qh[(size_t)row * 1024 + col] = hi;
        qh[(size_t)row * 1024 + 512 + col] = lo;
        if (col == 0) { thr[row] = -INFINITY; cnt[row] = 0; }
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) *overflow = 0;
}

// one block per query: sort the candidate list (descending by (score, ~id)), keep the best k,
// raise the threshold; on the final call also write D / I with faiss padding
__global__ void __launch_bounds__(256)
batch_compact_kernel(uint64_t *__restrict__ cand, uint32_t *__restrict__ cnt, float *__restrict__ thr, int64_t k,
                     int final_pass, int64_t id_base, float *__restrict__ D, int64_t *__restrict__ I) {
    extern __shared__ uint64_t s_c[];
    const int q = blockIdx.x;
    uint64_t *mine = cand + (size_t)q * kCap;
    const uint32_t c = min(cnt[q], (uint32_t)kCap);
    if (!final_pass && c <= (uint32_t)k) return;
    uint32_t p2 = 1;
    while (p2 < c) p2 <<= 1;
    for (uint32_t i = threadIdx.x; i < p2; i += blockDim.x) s_c[i] = i < c ? mine[i] : 0ull;
    __syncthreads();
    for (uint32_t kk = 2; kk <= p2; kk <<= 1) {
        for (uint32_t j = kk >> 1; j > 0; j >>= 1) {
            for (uint32_t t = threadIdx.x; t < p2 / 2; t += blockDim.x) {
                uint32_t i = ((t & ~(j - 1)) << 1) | (t & (j - 1));
                uint32_t l = i | j;
                bool desc = (i & kk) == 0;
                uint64_t x = s_c[i], y = s_c[l];
                if ((x < y) == desc) { s_c[i] = y; s_c[l] = x; }
            }
            __syncthreads();
        }
    }
    const uint32_t keep = min(c, (uint32_t)k);
    for (uint32_t i = threadIdx.x; i < keep; i += blockDim.x) mine[i] = s_c[i];
    if (threadIdx.x == 0) {
        cnt[q] = keep;
        if (keep == (uint32_t)k) thr[q] = key2f((uint32_t)(s_c[k - 1] >> 32));
    }
    if (final_pass) {
        for (int64_t j = threadIdx.x; j < k; j += blockDim.x) {
            if (j < keep) {
                uint64_t e = s_c[j];
                D[(size_t)q * k + j] = key2f((uint32_t)(e >> 32));
                I[(size_t)q * k + j] = id_base + (int64_t)(0xffffffffu - (uint32_t)e);
            } else {
                D[(size_t)q * k + j] = -3.4028234663852886e38f;
                I[(size_t)q * k + j] = -1;
            }
        }
    }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                  const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int make_map(CUtensorMap *m, const void *base, uint64_t rows, uint64_t cols, uint32_t box_rows) {
    static EncodeTiledFn fn = nullptr;
    if (!fn) {
        void *p = nullptr;
        cudaDriverEntryPointQueryResult qr;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qr) != cudaSuccess ||
            qr != cudaDriverEntryPointSuccess) {
            set_error("cuTensorMapEncodeTiled entry point not available");
            return CB_ERR_CUDA;
        }
        fn = reinterpret_cast<EncodeTiledFn>(p);
    }
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
    float *thr = nullptr;
    uint32_t *cnt = nullptr;
    uint64_t *cand = nullptr;
    int *overflow = nullptr;
    int *h_overflow = nullptr;
    int nq_pad_cap = 0;
};

int batch_ws_ensure(BatchWs *w, int nq_pad) {
    if (nq_pad <= w->nq_pad_cap) return CB_OK;
    cudaFree(w->qh); cudaFree(w->thr); cudaFree(w->cnt); cudaFree(w->cand);
    w->qh = nullptr; w->thr = nullptr; w->cnt = nullptr; w->cand = nullptr;
    CB_CUDA(cudaMalloc(&w->qh, (size_t)nq_pad * 1024 * 2));
    CB_CUDA(cudaMalloc(&w->thr, (size_t)nq_pad * 4));
    CB_CUDA(cudaMalloc(&w->cnt, (size_t)nq_pad * 4));
    CB_CUDA(cudaMalloc(&w->cand, (size_t)nq_pad * kCap * 8));
    if (!w->overflow) {
        CB_CUDA(cudaMalloc(&w->overflow, 4));
        CB_CUDA(cudaMallocHost(&w->h_overflow, 4));
    }
    w->nq_pad_cap = nq_pad;
    return CB_OK;
}

void batch_ws_free(BatchWs *w) {
    cudaFree(w->qh); cudaFree(w->thr); cudaFree(w->cnt); cudaFree(w->cand); cudaFree(w->overflow);
    if (w->h_overflow) cudaFreeHost(w->h_overflow);
    *w = BatchWs();
}

BatchWs *batch_ws_new() { return new BatchWs(); }
void batch_ws_delete(BatchWs *w) { if (w) { batch_ws_free(w); delete w; } }

// Returns CB_OK and sets *overflowed.  Synchronises `s` once (to read the overflow flag).
int flatip_search_batch(BatchWs *w, const void *rows_f16, int64_t n, int device, int64_t nq, const float *q_dev,
                        int64_t k, float *D_dev, int64_t *I_dev, int64_t id_base, cudaStream_t s,
                        bool *overflowed) {
    *overflowed = false;
    const int nq_pad = (int)((nq + 2 * QM - 1) / (2 * QM) * (2 * QM));
    int rc = batch_ws_ensure(w, nq_pad);
    if (rc) return rc;
    static bool attr_done[64] = {false};       // function attributes are per device
    if (!attr_done[device & 63]) {
        CB_CUDA(cudaFuncSetAttribute(flatip_batch_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
        CB_CUDA(cudaFuncSetAttribute(batch_compact_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kCap * 8));
        attr_done[device & 63] = true;
    }
    batch_prep_kernel<<<std::min(nq_pad * 2, kNumSMs * 8), 256, 0, s>>>(q_dev, (int)nq, nq_pad, w->qh, w->thr, w->cnt, w->overflow);
    CB_LAUNCH_CHECK();
    CUtensorMap tmQ, tmX;
    if ((rc = make_map(&tmQ, w->qh, (uint64_t)nq_pad, 1024, QM))) return rc;
    if ((rc = make_map(&tmX, rows_f16, (uint64_t)n, 512, RN / 2))) return rc;
    int sms = kNumSMs;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device);
    const int q_blocks = nq_pad / (2 * QM);
    // row ranges grow geometrically so the expected survivors per range stay ~4k
    int64_t r0 = 0, span = 4096;
    while (r0 < n) {
        const int64_t r1 = std::min(n, r0 + span);
        const int64_t items = ((r1 - r0 + RN - 1) / RN) * q_blocks;
        const int pairs = (int)std::min<int64_t>(items, sms / 2);
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3(pairs * 2);
        cfg.blockDim = dim3(kThreads);
        cfg.dynamicSmemBytes = SMEM_BYTES;
        cfg.stream = s;
        cudaLaunchAttribute attr[1];
        attr[0].id = cudaLaunchAttributeClusterDimension;
        attr[0].val.clusterDim.x = 2; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
        cfg.attrs = attr;
        cfg.numAttrs = 1;
        CB_CUDA(cudaLaunchKernelEx(&cfg, flatip_batch_kernel, tmQ, tmX, r0, r1, (int)nq, q_blocks,
                                   (const float *)w->thr, w->cnt, w->cand, w->overflow));
        CB_LAUNCH_CHECK();
        const int final_pass = r1 >= n ? 1 : 0;
        batch_compact_kernel<<<(unsigned)nq, 256, kCap * 8, s>>>(w->cand, w->cnt, w->thr, k, final_pass, id_base, D_dev, I_dev);
        CB_LAUNCH_CHECK();
        r0 = r1;
        span = std::min<int64_t>(span * 4, 4ll << 20);
    }
    CB_CUDA(cudaMemcpyAsync(w->h_overflow, w->overflow, 4, cudaMemcpyDeviceToHost, s));
    CB_CUDA(cudaStreamSynchronize(s));
    *overflowed = *w->h_overflow != 0;
    return CB_OK;
}

}  // namespace cb
