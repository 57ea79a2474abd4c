// Hand-written tcgen05 GEMM (sm_100a): TMA -> 128B-swizzled smem ring -> tcgen05.mma
// (kind::f16, K=16 per instruction) -> fp32 accumulators in TMEM (double buffered) ->
// tcgen05.ld epilogue fused with bias / QuickGELU / residual / pos-emb, staged through
// shared memory so global stores (and the residual read) are row-coalesced.
//
// Two flavours of the same kernel (template NCTA):
//   NCTA = 2  cta_group::2: a cluster of two CTAs (one SM pair) computes a 256 x BN tile.
//             Each CTA TMA-loads its own 128 rows of A and HALF of the W tile; the leader
//             CTA issues one M=256 MMA that reads both halves, so per-SM L2->smem operand
//             traffic drops from 16+BN/8 KB to 16+BN/16 KB per k-block.  Used for the big
//             tower GEMMs (the 1-CTA form is operand-bandwidth bound at ~60 % tensor pipe).
//   NCTA = 1  cta_group::1, 128 x BN tile per CTA: small problems (M <= 128 rows or fewer
//             tiles than SM pairs).
//
// Replaces the cuBLAS/cuDNN calls torch dispatches for openai/CLIP's conv1, in_proj,
// out_proj, c_fc, c_proj, proj  (reference
// call sites /root/reference/build-index.py:49, query-index.py:108).
//
// Persistent kernel, one CTA per SM, static tile schedule (n fastest so concurrently
// running CTAs share A tiles through L2).  Warp roles (320 threads):
//   warp 0      TMA producer (one elected lane)
//   warp 1      TMEM allocator + MMA issuer (one elected lane)
//   warps 2..9  epilogue; warp w owns TMEM lanes 32*(w%4) .. +31 (= tile rows) and the
//               column half (w-2)/4 of the tile
#include "common.cuh"
#include "gemm.cuh"
#include "tc_ptx.cuh"

#include <cuda.h>

#include <algorithm>
#include <mutex>

namespace cb {
namespace {

constexpr int BM = 128;
constexpr int BK = 64;          // 64 fp16 = 128 B = one swizzle atom row
constexpr int UMMA_K = 16;
constexpr int kEpiWarps = 8;     // two per TMEM lane quadrant, each owns half of the tile's columns
constexpr int kThreads = 64 + kEpiWarps * 32;

// staging boxes per epilogue warp: the bias / bias+GELU epilogues stream their 32-column chunks through a
// ring of kRingBoxes (the bulk store of a box overlaps the next chunk's math), which leaves room
// for one or two more pipeline stages; the residual and patch epilogues need all chunks resident
constexpr int kRingBoxes = 2;
template <int EPI>
__host__ __device__ constexpr bool kRingEpi() { return EPI == EPI_BIAS || EPI == EPI_BIAS_GELU; }

template <int BN, int NCTA, int EPI>
struct Cfg {
    static constexpr int A_BYTES = BM * BK * 2;
    static constexpr int B_ROWS = BN / NCTA;                  // rows of W this CTA loads per k-block
    static constexpr int B_BYTES = B_ROWS * BK * 2;
    static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
    static constexpr int TMEM_COLS = 2 * BN <= 256 ? 256 : 512;
    static constexpr int BAR_BYTES = 256;
    // epilogue staging: per epilogue warp 32 rows x BN/2 fp16 as dense boxes of 32 rows x 32
    // columns (64-byte rows, 64B-swizzled: the layout a TMA store expects, and bank-conflict free
    // for both row-per-thread and row-contiguous accesses), followed by the warp's BN/2 fp32 bias
    static constexpr int EPI_COLS = BN / 2;
    static constexpr int EPI_BOX_BYTES = 32 * 64;
    static constexpr int NBOX = kRingEpi<EPI>() && EPI_COLS / 32 > kRingBoxes ? kRingBoxes : EPI_COLS / 32;
    static constexpr int EPI_BIAS_OFF = NBOX * EPI_BOX_BYTES;
    static constexpr int EPI_WARP_BYTES = EPI_BIAS_OFF + 2 * EPI_COLS * 4;   // + bias and colsum slices
    static constexpr int kMaxSmem = 232448;                   // 227 KB opt-in limit
    static constexpr int FIXED = BAR_BYTES + kEpiWarps * EPI_WARP_BYTES + 1024 + 1024;   // + alignment slack
    static constexpr int STAGES_FIT = (kMaxSmem - FIXED) / STAGE_BYTES;
    static constexpr int STAGES = STAGES_FIT > 8 ? 8 : STAGES_FIT;
    static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + FIXED;
    static_assert(STAGES >= 3, "pipeline too shallow");
};

using namespace tc;

template <int EPI>
__host__ __device__ constexpr bool kHasBiasT() { return EPI == EPI_BIAS || EPI == EPI_BIAS_GELU || EPI == EPI_BIAS_RESID; }

// ---- the kernel -------------------------------------------------------------------------
template <int BN, int EPI, int NCTA>
__global__ void __launch_bounds__(kThreads, 1)
gemm_tcgen05_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                    const __grid_constant__ CUtensorMap tmC, const GemmArgs g, const int m_tiles, const int n_tiles, const int stages_and_dbg) {
    const int num_stages = stages_and_dbg & 0xff, raster = stages_and_dbg >> 16;
#ifdef CLIPB200_EXPERIMENTS
    const int dbg_mode = (stages_and_dbg >> 8) & 0xff;     // result-corrupting perf probes (profiles/gemm_decompose.py)
#else
    constexpr int dbg_mode = 0;                            // compiled out of the product library
#endif
    // tile -> (m tile, n tile).  raster 0: n fastest (concurrent CTAs share an A tile); 1: m fastest (share a
    // W tile); g >= 2: groups of g m-tiles, m fastest inside a group (share both)
    auto decode = [&](int tile, int &mt, int &nt) {
        if (raster == 0) { mt = tile / n_tiles; nt = tile % n_tiles; }
        else if (raster == 1) { mt = tile % m_tiles; nt = tile / m_tiles; }
        else {
            const int per = raster * n_tiles, grp = tile / per, in = tile - grp * per;
            const int rows = min(raster, m_tiles - grp * raster);
            mt = grp * raster + in % rows; nt = in / rows;
        }
    };
    // m_tiles counts (128*NCTA)-row tiles; a cluster of NCTA CTAs owns one tile at a time
    using C = Cfg<BN, NCTA, EPI>;
    const uint32_t cta_rank = NCTA == 1 ? 0u : cluster_ctarank();
    const bool leader = cta_rank == 0;
    const int cluster_id = blockIdx.x / NCTA, num_clusters = gridDim.x / NCTA;
    extern __shared__ uint8_t smem_raw[];
    uint8_t *smem = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    // the ring depth is a launch parameter: the shared-memory layout follows it, so a shallower ring really
    // leaves the rest of the SM's shared memory to co-resident CTAs of other kernels
    uint64_t *full = reinterpret_cast<uint64_t *>(smem + num_stages * C::STAGE_BYTES);
    uint64_t *empty = full + num_stages;
    uint64_t *tfull = empty + num_stages;
    uint64_t *tempty = tfull + 2;
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(tempty + 2);

    const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);    // warp-uniform by construction
    const int lane = threadIdx.x & 31;

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tmA);
        tma_prefetch_desc(&tmB);
    }
    if (warp == 1) {
        if (lane == 0) {
            for (int i = 0; i < num_stages; i++) {
                mbar_init(&full[i], 1);
                mbar_init(&empty[i], 1);
            }
            for (int i = 0; i < 2; i++) {
                mbar_init(&tfull[i], 1);
                mbar_init(&tempty[i], kEpiWarps * NCTA);   // one arrive per epilogue warp of every CTA
            }
            fence_barrier_init();
        }
        __syncwarp();
        tmem_alloc<NCTA>(tmem_slot, C::TMEM_COLS);
    }
    tc_fence_before();
    __syncthreads();
    if (NCTA > 1) cluster_sync_all();          // peer barriers are initialised before any remote arrive
    tc_fence_after();
    const uint32_t tmem_base = *reinterpret_cast<volatile uint32_t *>(tmem_slot);
    // everything above touched only this CTA's shared / tensor memory: under programmatic dependent launch it
    // ran while the previous kernel of the stream was still draining.  From here on global memory is read.
    pdl_launch_dependents();
    pdl_wait();
    if (g.stamp != nullptr && threadIdx.x == 0) {      // live timing: this launch's work starts here
        unsigned long long t;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
        atomicMin(g.stamp, t);
    }

    const int num_tiles = m_tiles * n_tiles;
    const int KB = g.K / BK;

    if (warp == 0) {
        // warp-uniform producer loop; one elected lane arms the barrier and issues the TMA loads
        uint32_t stage = 0, phase = 0;
        for (int tile = cluster_id; tile < num_tiles; tile += num_clusters) {
            int mt_, nt_;
            decode(tile, mt_, nt_);
            const int m_blk = mt_ * NCTA + (int)cta_rank, n_blk = nt_;
            for (int kb = 0; kb < KB; kb++) {
                mbar_wait(&empty[stage], phase ^ 1);
                uint8_t *sa = smem + stage * C::STAGE_BYTES;
                if (elect_one()) {
                    if (NCTA == 1) {
                        mbar_expect_tx(&full[stage], C::STAGE_BYTES);
                        tma_load_2d(sa, &tmA, kb * BK, m_blk * BM, &full[stage]);
                        tma_load_2d(sa + C::A_BYTES, &tmB, kb * BK, n_blk * BN, &full[stage]);
                    } else if ((dbg_mode == 3 || dbg_mode == 5) && tile != cluster_id) {
                        // experiment: traffic of a W-stationary schedule (W loaded for the first tile only;
                        // results are garbage)
                        if (leader) mbar_expect_tx(&full[stage], NCTA * C::A_BYTES);
                        tma_load_2d_2sm(sa, &tmA, kb * BK, m_blk * BM, &full[stage]);
                    } else {
                        // both CTAs' bytes are credited to the leader's barrier
                        if (leader) mbar_expect_tx(&full[stage], NCTA * C::STAGE_BYTES);
                        tma_load_2d_2sm(sa, &tmA, kb * BK, m_blk * BM, &full[stage]);
                        tma_load_2d_2sm(sa + C::A_BYTES, &tmB, kb * BK, n_blk * BN + (int)cta_rank * C::B_ROWS,
                                        &full[stage]);
                    }
                }
                __syncwarp();
                if (++stage == (uint32_t)num_stages) { stage = 0; phase ^= 1; }
            }
        }
    } else if (warp == 1) {
        if (leader) {
            // the whole warp walks the schedule (uniform control flow, uniform registers); one elected lane
            // issues the MMAs and commits
            constexpr uint32_t idesc = make_idesc(BM * NCTA, BN);
            uint32_t stage = 0, phase = 0, as = 0, aphase = 0;
            for (int tile = cluster_id; tile < num_tiles; tile += num_clusters) {
                mbar_wait(&tempty[as], aphase ^ 1);
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + as * BN;
                for (int kb = 0; kb < KB; kb++) {
                    mbar_wait(&full[stage], phase);
                    tc_fence_after();
                    const uint32_t sa = smem_u32(smem + stage * C::STAGE_BYTES);
                    const uint64_t da = make_smem_desc(sa);
                    const uint64_t db = make_smem_desc(sa + C::A_BYTES);
                    if (elect_one()) {
#pragma unroll
                        for (int k = 0; k < BK / UMMA_K; k++) {
                            // advance 16 elements = 32 B along K inside the swizzle atom: +2 in 16-B units
                            umma_f16<NCTA>(d_tmem, da + 2 * k, db + 2 * k, idesc, (kb | k) != 0 ? 1u : 0u);
                        }
                        umma_commit<NCTA>(&empty[stage]);    // frees the smem slot (in both CTAs) when the MMAs retire
                    }
                    __syncwarp();
                    if (++stage == (uint32_t)num_stages) { stage = 0; phase ^= 1; }
                }
                if (elect_one()) umma_commit<NCTA>(&tfull[as]);   // accumulator complete -> epilogue (both CTAs)
                __syncwarp();
                as ^= 1;
                if (as == 0) aphase ^= 1;
            }
        }
    } else {
        const int q = warp & 3;                          // TMEM lane quadrant this warp may touch
        const int half = (warp - 2) >> 2;                // which half of the tile's columns
        uint32_t as = 0, aphase = 0;
        constexpr int EC = C::EPI_COLS;                  // columns per epilogue warp
        constexpr int CPR = EC / 8;                      // 16-byte chunks per staged row
        constexpr bool kTmaStore = EPI != EPI_F32 && EPI != EPI_PATCH;
        // staging area starts 1024-byte aligned (TMA + swizzle pattern alignment)
        const uint32_t stage_area = (smem_u32(smem + num_stages * C::STAGE_BYTES + C::BAR_BYTES) + 1023u) & ~1023u;
        constexpr int NBOX = C::NBOX;
        const uint32_t stage_base = stage_area + (uint32_t)(warp - 2) * C::EPI_BOX_BYTES * NBOX;
        const uint32_t bias_smem = stage_area + kEpiWarps * C::EPI_BOX_BYTES * NBOX + (uint32_t)(warp - 2) * EC * 8;
        const uint32_t csum_smem = bias_smem + EC * 4;
        const bool ln_fold = kHasBiasT<EPI>() && g.ln_stats != nullptr;
        const bool emit_stats = EPI == EPI_BIAS_RESID && g.stats_out != nullptr;
        const int out_slices = g.N / EC;
        // 16-byte chunk j (of this warp's EC columns) of row r: box j/4, 64-byte rows, 64B swizzle
        auto stg = [&](int r, int j) -> uint32_t {
            return stage_base + (uint32_t)((j >> 2) % NBOX) * C::EPI_BOX_BYTES + (uint32_t)r * 64u +
                   (uint32_t)((((j & 3) ^ ((r >> 1) & 3))) << 4);
        };
        constexpr bool kHasBias = EPI == EPI_BIAS || EPI == EPI_BIAS_GELU || EPI == EPI_BIAS_RESID;
        if (kTmaStore && lane == 0) tma_prefetch_desc(&tmC);
        for (int tile = cluster_id; tile < num_tiles; tile += num_clusters) {
            int mt_, nt_;
            decode(tile, mt_, nt_);
            const int m_blk = mt_ * NCTA + (int)cta_rank, n_blk = nt_;
            const int row0 = m_blk * BM + q * 32;        // first row of this warp's slice
            const int row = row0 + lane;
            const bool row_ok = row < g.M;
            const int n_base = n_blk * BN + half * EC;   // first column of this warp's slice
            if (kTmaStore && EPI == EPI_BIAS_RESID) {
                // the residual prefetch below overwrites every staging box: the previous tile's bulk
                // stores must have finished reading them
                if (elect_one()) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
                __syncwarp();
            }
            // -- before the accumulator is ready: stage bias and (coalesced, async) the residual
            if (kHasBias) {
                for (int j = lane; j < EC / 4; j += 32) {
                    float4 b = g.bias ? __ldg(reinterpret_cast<const float4 *>(g.bias + n_base) + j) : make_float4(0, 0, 0, 0);
                    asm volatile("st.shared.v4.f32 [%0], {%1,%2,%3,%4};" ::"r"(bias_smem + j * 16), "f"(b.x), "f"(b.y), "f"(b.z), "f"(b.w) : "memory");
                }
            }
            float ln_mean = 0.f, ln_rstd = 1.f;
            if (ln_fold) {
                for (int j = lane; j < EC / 4; j += 32) {
                    float4 b = __ldg(reinterpret_cast<const float4 *>(g.colsum + n_base) + j);
                    asm volatile("st.shared.v4.f32 [%0], {%1,%2,%3,%4};" ::"r"(csum_smem + j * 16), "f"(b.x), "f"(b.y), "f"(b.z), "f"(b.w) : "memory");
                }
                if (row_ok) {
                    const float2 *sp = reinterpret_cast<const float2 *>(g.ln_stats) + (size_t)row * g.ln_slices;
                    float sx = 0.f, sxx = 0.f;
                    for (int t = 0; t < g.ln_slices; t++) { float2 p = __ldcg(sp + t); sx += p.x; sxx += p.y; }
                    const float inv = 1.0f / (float)g.K;
                    ln_mean = sx * inv;
                    ln_rstd = rsqrtf(fmaxf(sxx * inv - ln_mean * ln_mean, 0.f) + 1e-5f);
                }
            }
            float st_sum = 0.f, st_sq = 0.f;
            if (EPI == EPI_BIAS_RESID) {
#pragma unroll 4
                for (int c = lane; c < 32 * CPR; c += 32) {
                    const int r = c / CPR, j = c - r * CPR;
                    if (row0 + r < g.M)
                        cp_async16(stg(r, j), g.resid + (size_t)(row0 + r) * g.N + n_base + j * 8);
                }
            }
            mbar_wait(&tfull[as], aphase);
            tc_fence_after();
            if (EPI == EPI_BIAS_RESID) cp_async_wait_all();
            __syncwarp();
            if (dbg_mode == 1 || dbg_mode == 5) {      // experiment: MMA-only throughput (results are not written)
                tc_fence_before();
                __syncwarp();
                if (lane == 0) {
                    if (NCTA == 1) mbar_arrive_relaxed(&tempty[as]);
                    else mbar_arrive_cluster(&tempty[as], 0);
                }
                as ^= 1;
                if (as == 0) aphase ^= 1;
                continue;
            }
            const float *pos_row = nullptr;
            if (EPI == EPI_PATCH) {
                const int img = row / 49, p = row - img * 49;
                pos_row = g.pos + (size_t)(1 + p) * g.N;
            }
#pragma unroll 1
            for (int c = 0; c < EC / 32; c++) {
                if (kTmaStore && EPI != EPI_BIAS_RESID) {
                    // box c is rewritten below: its bulk store of the previous tile (one group per box, so
                    // NB - 1 younger groups may still be pending) must have finished reading it
                    if (elect_one()) asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(NBOX - 1) : "memory");
                    __syncwarp();
                }
                uint32_t v[32];
                tmem_ld_32x32(tmem_base + ((uint32_t)(q * 32) << 16) + as * BN + half * EC + c * 32, v);
                const int n0 = n_base + c * 32;
                float add[32];
                if (kHasBias) {
#pragma unroll
                    for (int j = 0; j < 8; j++) {
                        float4 b;
                        asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(b.x), "=f"(b.y), "=f"(b.z), "=f"(b.w) : "r"(bias_smem + c * 128 + j * 16));
                        add[4 * j] = b.x; add[4 * j + 1] = b.y; add[4 * j + 2] = b.z; add[4 * j + 3] = b.w;
                    }
                } else if (EPI == EPI_PATCH) {
                    const float4 *p4 = reinterpret_cast<const float4 *>(pos_row + n0);
#pragma unroll
                    for (int j = 0; j < 8; j++) {
                        float4 b = row_ok ? __ldg(p4 + j) : make_float4(0, 0, 0, 0);
                        add[4 * j] = b.x; add[4 * j + 1] = b.y; add[4 * j + 2] = b.z; add[4 * j + 3] = b.w;
                    }
                } else {
#pragma unroll
                    for (int j = 0; j < 32; j++) add[j] = 0.f;
                }
                uint4 rz[4];
                if (EPI == EPI_BIAS_RESID) {
#pragma unroll
                    for (int j = 0; j < 4; j++) rz[j] = lds128(stg(lane, c * 4 + j));
                }
                tmem_ld_wait();
                float o[32];
                if (ln_fold) {
                    // y = rstd * (acc - mean * colsum[n]) + b'[n]
                    const float nrm = -ln_rstd * ln_mean;
#pragma unroll
                    for (int j = 0; j < 8; j++) {
                        float4 cs;
                        asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(cs.x), "=f"(cs.y), "=f"(cs.z), "=f"(cs.w) : "r"(csum_smem + c * 128 + j * 16));
                        o[4 * j] = fmaf(ln_rstd, __uint_as_float(v[4 * j]), fmaf(nrm, cs.x, add[4 * j]));
                        o[4 * j + 1] = fmaf(ln_rstd, __uint_as_float(v[4 * j + 1]), fmaf(nrm, cs.y, add[4 * j + 1]));
                        o[4 * j + 2] = fmaf(ln_rstd, __uint_as_float(v[4 * j + 2]), fmaf(nrm, cs.z, add[4 * j + 2]));
                        o[4 * j + 3] = fmaf(ln_rstd, __uint_as_float(v[4 * j + 3]), fmaf(nrm, cs.w, add[4 * j + 3]));
                    }
                } else {
#pragma unroll
                    for (int j = 0; j < 32; j++) o[j] = __uint_as_float(v[j]) + add[j];
                }
                if (EPI == EPI_BIAS_GELU) {
                    // QuickGELU x*sigmoid(1.702x) = 0.5x + 0.5x*tanh(0.851x): one SFU op per element
#pragma unroll
                    for (int j = 0; j < 32; j++) {
                        float t;
                        asm("tanh.approx.f32 %0, %1;" : "=f"(t) : "f"(0.851f * o[j]));
                        const float hx = 0.5f * o[j];
                        o[j] = fmaf(hx, t, hx);
                    }
                }
                if (EPI == EPI_BIAS_RESID) {
#pragma unroll
                    for (int j = 0; j < 4; j++) {
                        const __half2 *h = reinterpret_cast<const __half2 *>(&rz[j]);
#pragma unroll
                        for (int e = 0; e < 4; e++) {
                            float2 f = __half22float2(h[e]);
                            o[8 * j + 2 * e] += f.x;
                            o[8 * j + 2 * e + 1] += f.y;
                        }
                    }
                }
                if (emit_stats) {
                    // row statistics for the next (folded) LayerNorm, from the fp32 values about to be
                    // rounded to fp16 (the rounding noise is zero-mean, 2^-11 relative per element)
#pragma unroll
                    for (int j = 0; j < 32; j++) {
                        st_sum += o[j];
                        st_sq = fmaf(o[j], o[j], st_sq);
                    }
                }
                if (EPI == EPI_F32) {
                    if (row_ok) {
                        float4 *dst = reinterpret_cast<float4 *>(reinterpret_cast<float *>(g.C) + (size_t)row * g.ldc + n0);
#pragma unroll
                        for (int j = 0; j < 8; j++) dst[j] = make_float4(o[4 * j], o[4 * j + 1], o[4 * j + 2], o[4 * j + 3]);
                    }
                } else {
                    // stage the fp16 row slice in the swizzled box of this 32-column chunk
#pragma unroll
                    for (int j = 0; j < 4; j++)
                        sts128(stg(lane, c * 4 + j),
                               make_uint4(pack_h2(o[8 * j], o[8 * j + 1]), pack_h2(o[8 * j + 2], o[8 * j + 3]),
                                          pack_h2(o[8 * j + 4], o[8 * j + 5]), pack_h2(o[8 * j + 6], o[8 * j + 7])));
                    if (kTmaStore && dbg_mode != 2) {
                        // bulk-store this box at once (rows >= M are clipped by the tensor map); one group
                        // per box so the next tile can reuse box c while later boxes are still in flight
                        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                        __syncwarp();
                        // one elected lane (always the same one: bulk groups are per thread) issues the store;
                        // behind elect.sync the tensor-map address and coordinates stay in uniform registers
                        if (elect_one()) {
                            asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];"
                                         ::"l"(reinterpret_cast<uint64_t>(&tmC)), "r"(stage_base + (c % NBOX) * C::EPI_BOX_BYTES),
                                           "r"(n0), "r"(row0)
                                         : "memory");
                            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
                        }
                    }
                }
            }
            if (emit_stats && row_ok)
                reinterpret_cast<float2 *>(g.stats_out)[(size_t)row * out_slices + (n_base / EC)] = make_float2(st_sum, st_sq);
            // all TMEM reads of this accumulator stage are done: hand it back to the MMA warp
            tc_fence_before();
            __syncwarp();
            if (elect_one()) {
                if (NCTA == 1) mbar_arrive_relaxed(&tempty[as]);
                else mbar_arrive_cluster(&tempty[as], 0);    // the leader's MMA warp waits for both CTAs
            }
            as ^= 1;
            if (as == 0) aphase ^= 1;
            if (EPI == EPI_PATCH) {
                // scattered rows (class-token gaps): coalesced manual copy-out from the staged boxes
                __half *Cb = reinterpret_cast<__half *>(g.C);
                for (int c = lane; c < 32 * CPR; c += 32) {
                    const int r = c / CPR, j = c - r * CPR;
                    const int grow = row0 + r;
                    if (grow < g.M) {
                        const uint4 val = lds128(stg(r, j));
                        *reinterpret_cast<uint4 *>(Cb + (size_t)(grow + grow / 49 + 1) * g.ldc + n_base + j * 8) = val;
                    }
                }
                __syncwarp();                            // staging boxes are reused by the next tile
            }
        }
        if (kTmaStore) {
            if (elect_one()) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
            __syncwarp();
        }
    }

    tc_fence_before();
    __syncthreads();
    if (NCTA > 1) cluster_sync_all();          // nobody exits while the peer may still touch its smem/TMEM
    if (warp == 1) {
        __syncwarp();
        tmem_dealloc<NCTA>(tmem_base, C::TMEM_COLS);
    }
    if (g.stamp != nullptr && threadIdx.x == 0) {
        unsigned long long t;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
        atomicMax(g.stamp + 1, t);
    }
}

// ---- single-row-block GEMMs (M <= 128): a lean kernel, optionally split-K over a thread-block cluster ----
// A single query (50 / 77 rows) makes every tower GEMM a read of the weight matrix with one 128-row tile of
// A: one 128 x 64 output tile per CTA, 192 threads, a ring only as deep as the k-loop, operands of the
// epilogue prefetched while the MMAs run, direct row stores.  With a cluster of S CTAs per tile, CTA r
// multiplies k-blocks [r KB/S, (r+1) KB/S), parks its fp32 partial tile in a scratch buffer (L2), the cluster
// synchronises (release / acquire at cluster scope) and the leader adds the S partials IN RANK ORDER
// (deterministic) and runs the epilogue -- built, tested, and measured SLOWER than S = 1 (skinny_split below).
constexpr int kSkinnyThreads = 192;     // warp 0 TMA, warp 1 TMEM + MMA, warps 2..5 epilogue (one TMEM lane quadrant each)
constexpr int kSkinnyBN = 64;
constexpr int kSkinnyStageBytes = BM * BK * 2 + kSkinnyBN * BK * 2;     // A 16 KB + W 8 KB
constexpr int kSkinnyMaxStages = 6;
__host__ __device__ constexpr int skinny_smem(int stages) { return stages * kSkinnyStageBytes + 256 + 1024 + 2 * kSkinnyBN * 4; }

__device__ __forceinline__ uint32_t cluster_nctarank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_nctarank;" : "=r"(r));
    return r;
}

template <int EPI>
__global__ void __launch_bounds__(kSkinnyThreads, 1)
gemm_skinny_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const GemmArgs g,
                   const int kb_per_cta, const int num_stages) {
    // the ring is as deep as this CTA's share of K (at most 6 stages): with a split of 4 or 8 that is 24 - 72 KB,
    // so two or more CTAs share an SM and every cluster of the launch is resident at once
    extern __shared__ uint8_t smem_raw[];
    uint8_t *smem = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    uint64_t *full = reinterpret_cast<uint64_t *>(smem + num_stages * kSkinnyStageBytes);
    uint64_t *empty = full + kSkinnyMaxStages;
    uint64_t *tfull = empty + kSkinnyMaxStages;
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(tfull + 1);
    float *s_bias = reinterpret_cast<float *>(reinterpret_cast<uint8_t *>(full) + 256);     // [64] bias, [64] colsum
    float *s_csum = s_bias + kSkinnyBN;

    const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);
    const int lane = threadIdx.x & 31;
    const uint32_t S = cluster_nctarank(), crank = cluster_ctarank();
    const int n_blk = blockIdx.x / S;
    const int kb0 = (int)crank * kb_per_cta;

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tmA);
        tma_prefetch_desc(&tmB);
    }
    if (warp == 1) {
        if (lane == 0) {
            for (int i = 0; i < num_stages; i++) { mbar_init(&full[i], 1); mbar_init(&empty[i], 1); }
            mbar_init(tfull, 1);
            fence_barrier_init();
        }
        __syncwarp();
        tmem_alloc<1>(tmem_slot, 64);
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *reinterpret_cast<volatile uint32_t *>(tmem_slot);
    pdl_launch_dependents();
    pdl_wait();
    if (g.stamp != nullptr && threadIdx.x == 0) {
        unsigned long long t;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
        atomicMin(g.stamp, t);
    }

    if (warp == 0) {
        uint32_t stage = 0, phase = 0;
        for (int i = 0; i < kb_per_cta; i++) {
            mbar_wait(&empty[stage], phase ^ 1);
            uint8_t *sa = smem + stage * kSkinnyStageBytes;
            if (elect_one()) {
                mbar_expect_tx(&full[stage], kSkinnyStageBytes);
                tma_load_2d(sa, &tmA, (kb0 + i) * BK, 0, &full[stage]);
                tma_load_2d(sa + BM * BK * 2, &tmB, (kb0 + i) * BK, n_blk * kSkinnyBN, &full[stage]);
            }
            __syncwarp();
            if (++stage == (uint32_t)num_stages) { stage = 0; phase ^= 1; }
        }
    } else if (warp == 1) {
        constexpr uint32_t idesc = make_idesc(BM, kSkinnyBN);
        uint32_t stage = 0, phase = 0;
        for (int i = 0; i < kb_per_cta; i++) {
            mbar_wait(&full[stage], phase);
            tc_fence_after();
            const uint32_t sa = smem_u32(smem + stage * kSkinnyStageBytes);
            const uint64_t da = make_smem_desc(sa), db = make_smem_desc(sa + BM * BK * 2);
            if (elect_one()) {
#pragma unroll
                for (int k = 0; k < BK / UMMA_K; k++) umma_f16<1>(tmem_base, da + 2 * k, db + 2 * k, idesc, (i | k) != 0 ? 1u : 0u);
                umma_commit<1>(&empty[stage]);
            }
            __syncwarp();
            if (++stage == (uint32_t)num_stages) { stage = 0; phase ^= 1; }
        }
        if (elect_one()) umma_commit<1>(tfull);
        __syncwarp();
    }

    // accumulator -> registers: thread = one row of the 128 x 64 tile
    const int q = warp & 3;
    const int row = q * 32 + lane;
    const bool row_ok = row < g.M;
    const int n0 = n_blk * kSkinnyBN;
    constexpr bool kHasBias = EPI == EPI_BIAS || EPI == EPI_BIAS_GELU || EPI == EPI_BIAS_RESID;
    const bool ln_fold = kHasBias && g.ln_stats != nullptr;
    float acc[kSkinnyBN];
    float ln_rstd = 1.f, ln_nrm = 0.f;
    uint4 resid_row[EPI == EPI_BIAS_RESID ? kSkinnyBN / 8 : 1];
    if (warp >= 2 && crank == 0) {
        // the leader's epilogue operands do not depend on the MMAs: fetch them while the tensor core works
        const int t = threadIdx.x - 64;
        if (t < kSkinnyBN) {
            s_bias[t] = (kHasBias && g.bias) ? __ldg(g.bias + n0 + t) : 0.f;
            s_csum[t] = ln_fold ? __ldg(g.colsum + n0 + t) : 0.f;
        }
        if (ln_fold && row_ok) {
            const float2 *sp = reinterpret_cast<const float2 *>(g.ln_stats) + (size_t)row * g.ln_slices;
            float sx = 0.f, sxx = 0.f;
            for (int i = 0; i < g.ln_slices; i++) { float2 p2 = __ldcg(sp + i); sx += p2.x; sxx += p2.y; }
            const float inv = 1.0f / (float)g.K;
            const float mean = sx * inv;
            ln_rstd = rsqrtf(fmaxf(sxx * inv - mean * mean, 0.f) + 1e-5f);
            ln_nrm = -ln_rstd * mean;
        }
        if (EPI == EPI_BIAS_RESID && row_ok) {
            const uint4 *rp = reinterpret_cast<const uint4 *>(g.resid + (size_t)row * g.N + n0);
#pragma unroll
            for (int j = 0; j < kSkinnyBN / 8; j++) resid_row[j] = rp[j];
        }
    }
    if (warp >= 2) {
        mbar_wait(tfull, 0);
        tc_fence_after();
        uint32_t v[32];
#pragma unroll
        for (int c = 0; c < 2; c++) {
            tmem_ld_32x32(tmem_base + ((uint32_t)(q * 32) << 16) + c * 32, v);
            tmem_ld_wait();
#pragma unroll
            for (int j = 0; j < 32; j++) acc[c * 32 + j] = __uint_as_float(v[j]);
        }
        if (S > 1 && row_ok) {
            float4 *dst = reinterpret_cast<float4 *>(g.skinny_scratch + ((size_t)blockIdx.x * BM + row) * kSkinnyBN);
#pragma unroll
            for (int j = 0; j < kSkinnyBN / 4; j++) dst[j] = make_float4(acc[4 * j], acc[4 * j + 1], acc[4 * j + 2], acc[4 * j + 3]);
        }
    }
    tc_fence_before();
    if (S > 1) cluster_sync_all();           // every CTA's partial tile is visible to the leader
    else __syncthreads();

    if (warp >= 2 && crank == 0 && row_ok) {
        if (S > 1) {
            // partials in rank order: the sum does not depend on which CTA finished first
#pragma unroll
            for (int j = 0; j < kSkinnyBN; j++) acc[j] = 0.f;
            for (uint32_t r = 0; r < S; r++) {
                const float4 *src = reinterpret_cast<const float4 *>(g.skinny_scratch + ((size_t)(blockIdx.x + r) * BM + row) * kSkinnyBN);
#pragma unroll
                for (int j = 0; j < kSkinnyBN / 4; j++) {
                    const float4 f = __ldcg(src + j);
                    acc[4 * j] += f.x; acc[4 * j + 1] += f.y; acc[4 * j + 2] += f.z; acc[4 * j + 3] += f.w;
                }
            }
        }
        if (ln_fold) {
            // y = rstd * (acc - mean * colsum[n]) + b'[n]
#pragma unroll
            for (int j = 0; j < kSkinnyBN; j++) acc[j] = fmaf(ln_rstd, acc[j], fmaf(ln_nrm, s_csum[j], s_bias[j]));
        } else if (kHasBias) {
#pragma unroll
            for (int j = 0; j < kSkinnyBN; j++) acc[j] += s_bias[j];
        } else if (EPI == EPI_PATCH) {
            const float *pos_row = g.pos + (size_t)(1 + row % 49) * g.N + n0;
#pragma unroll
            for (int j = 0; j < kSkinnyBN; j++) acc[j] += __ldg(pos_row + j);
        }
        if (EPI == EPI_BIAS_GELU) {
#pragma unroll
            for (int j = 0; j < kSkinnyBN; j++) {
                float t;
                asm("tanh.approx.f32 %0, %1;" : "=f"(t) : "f"(0.851f * acc[j]));
                const float hx = 0.5f * acc[j];
                acc[j] = fmaf(hx, t, hx);
            }
        }
        if (EPI == EPI_BIAS_RESID) {
#pragma unroll
            for (int j = 0; j < kSkinnyBN / 8; j++) {
                const uint4 u = resid_row[j];
                const __half2 *h = reinterpret_cast<const __half2 *>(&u);
#pragma unroll
                for (int e = 0; e < 4; e++) {
                    const float2 f = __half22float2(h[e]);
                    acc[8 * j + 2 * e] += f.x;
                    acc[8 * j + 2 * e + 1] += f.y;
                }
            }
            if (g.stats_out != nullptr) {
                // the folded LayerNorm of the next GEMM: (sum, sum of squares) per 32-column slice, as the
                // general kernel's 64-wide tile writes them (gemm_out_slices)
                const int out_slices = g.N / 32;
#pragma unroll
                for (int c = 0; c < 2; c++) {
                    float st_sum = 0.f, st_sq = 0.f;
#pragma unroll
                    for (int j = 0; j < 32; j++) { st_sum += acc[c * 32 + j]; st_sq = fmaf(acc[c * 32 + j], acc[c * 32 + j], st_sq); }
                    reinterpret_cast<float2 *>(g.stats_out)[(size_t)row * out_slices + n0 / 32 + c] = make_float2(st_sum, st_sq);
                }
            }
        }
        if (EPI == EPI_F32) {
            float4 *dst = reinterpret_cast<float4 *>(reinterpret_cast<float *>(g.C) + (size_t)row * g.ldc + n0);
#pragma unroll
            for (int j = 0; j < kSkinnyBN / 4; j++) dst[j] = make_float4(acc[4 * j], acc[4 * j + 1], acc[4 * j + 2], acc[4 * j + 3]);
        } else {
            const size_t orow = EPI == EPI_PATCH ? (size_t)(row + row / 49 + 1) : (size_t)row;
            uint4 *dst = reinterpret_cast<uint4 *>(reinterpret_cast<__half *>(g.C) + orow * g.ldc + n0);
#pragma unroll
            for (int j = 0; j < kSkinnyBN / 8; j++)
                dst[j] = make_uint4(pack_h2(acc[8 * j], acc[8 * j + 1]), pack_h2(acc[8 * j + 2], acc[8 * j + 3]),
                                    pack_h2(acc[8 * j + 4], acc[8 * j + 5]), pack_h2(acc[8 * j + 6], acc[8 * j + 7]));
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        __syncwarp();
        tmem_dealloc<1>(tmem_base, 64);
    }
    if (g.stamp != nullptr && threadIdx.x == 0) {
        unsigned long long t;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
        atomicMax(g.stamp + 1, t);
    }
}

// ---- host ---------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                  const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn get_encode_fn() {
    static EncodeTiledFn fn = nullptr;
    if (!fn) {
        void *p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess ||
            q != cudaDriverEntryPointSuccess)
            return nullptr;
        fn = reinterpret_cast<EncodeTiledFn>(p);
    }
    return fn;
}

// 2-D fp16 row-major [rows, cols] tensor, box = box_rows x 64 columns, 128B swizzle
int make_map(CUtensorMap *m, const void *base, uint64_t rows, uint64_t cols, uint32_t box_rows) {
    EncodeTiledFn fn = get_encode_fn();
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

// output map for the epilogue's bulk stores: fp16 [rows, cols], box = 32 rows x 32 columns
// (64-byte rows), 64B swizzle
int make_out_map(CUtensorMap *m, const void *base, uint64_t rows, uint64_t cols, uint64_t ld_elems) {
    EncodeTiledFn fn = get_encode_fn();
    if (!fn) { set_error("cuTensorMapEncodeTiled entry point not available"); return CB_ERR_CUDA; }
    cuuint64_t gdim[2] = {cols, rows};
    cuuint64_t gstride[1] = {ld_elems * 2};
    cuuint32_t box[2] = {32, 32};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = fn(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, const_cast<void *>(base), gdim, gstride, box, estr,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled (output) failed (%d)", (int)r); return CB_ERR_CUDA; }
    return CB_OK;
}

template <int BN, int EPI, int NCTA>
int launch(const GemmArgs &g, cudaStream_t s) {
    using C = Cfg<BN, NCTA, EPI>;
    auto kern = gemm_tcgen05_kernel<BN, EPI, NCTA>;
    int cur_dev = 0;
    CB_CUDA(cudaGetDevice(&cur_dev));
    // function attributes are per device; handles on different GPUs launch from different host threads
    static std::once_flag attr_once[64];
    cudaError_t attr_err = cudaSuccess;
    std::call_once(attr_once[cur_dev & 63], [&] {
        attr_err = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM_BYTES);
    });
    CB_CUDA(attr_err);
    CUtensorMap tmA, tmB;
    int rc = make_map(&tmA, g.A, (uint64_t)g.M, (uint64_t)g.K, BM);
    if (rc) return rc;
    rc = make_map(&tmB, g.W, (uint64_t)g.N, (uint64_t)g.K, C::B_ROWS);
    if (rc) return rc;
    CUtensorMap tmC = tmA;                     // placeholder when the epilogue does not bulk-store
    if (EPI != EPI_F32 && EPI != EPI_PATCH) {
        rc = make_out_map(&tmC, g.C, (uint64_t)g.M, (uint64_t)g.N, (uint64_t)g.ldc);
        if (rc) return rc;
    }
    const int m_tiles = (g.M + BM * NCTA - 1) / (BM * NCTA), n_tiles = g.N / BN;
    int dev = 0, sms = kNumSMs;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const int clusters = std::min(m_tiles * n_tiles, sms / NCTA);
    int stages = C::STAGES;
    if (tune(T_GEMM_STAGES) > 0) stages = std::max(2, std::min(C::STAGES, (int)tune(T_GEMM_STAGES)));
    if (EPI == EPI_BIAS_RESID && tune(T_GEMM_RESID_STAGES) > 0) stages = std::max(2, std::min(stages, (int)tune(T_GEMM_RESID_STAGES)));
#ifdef CLIPB200_EXPERIMENTS
    if (tune(T_GEMM_DEBUG) > 0) stages |= (int)tune(T_GEMM_DEBUG) << 8;
#endif
    if (tune(T_GEMM_RASTER) > 0) stages |= (int)tune(T_GEMM_RASTER) << 16;
    const size_t smem_bytes = (size_t)(stages & 0xff) * C::STAGE_BYTES + C::FIXED;
    CB_CUDA(launch_ex(kern, dim3(clusters * NCTA), dim3(kThreads), smem_bytes, s, NCTA, true, tmA, tmB, tmC, g, m_tiles,
                      n_tiles, stages));
    CB_LAUNCH_CHECK();
    return CB_OK;
}

template <int BN, int NCTA>
int dispatch_epi(const GemmArgs &g, cudaStream_t s) {
    switch (g.epilogue) {
        case EPI_BIAS: return launch<BN, EPI_BIAS, NCTA>(g, s);
        case EPI_BIAS_GELU: return launch<BN, EPI_BIAS_GELU, NCTA>(g, s);
        case EPI_BIAS_RESID: return launch<BN, EPI_BIAS_RESID, NCTA>(g, s);
        case EPI_PATCH: return launch<BN, EPI_PATCH, NCTA>(g, s);
        case EPI_F32: return launch<BN, EPI_F32, NCTA>(g, s);
    }
    set_error("gemm_f16: unknown epilogue %d", g.epilogue);
    return CB_ERR_INVALID;
}


template <int EPI>
int launch_skinny(const GemmArgs &g, int split, cudaStream_t s) {
    auto kern = gemm_skinny_kernel<EPI>;
    int cur_dev = 0;
    CB_CUDA(cudaGetDevice(&cur_dev));
    static std::once_flag attr_once[64];
    cudaError_t attr_err = cudaSuccess;
    std::call_once(attr_once[cur_dev & 63], [&] {
        attr_err = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, skinny_smem(kSkinnyMaxStages));
    });
    CB_CUDA(attr_err);
    CUtensorMap tmA, tmB;
    int rc = make_map(&tmA, g.A, (uint64_t)g.M, (uint64_t)g.K, BM);
    if (rc) return rc;
    if ((rc = make_map(&tmB, g.W, (uint64_t)g.N, (uint64_t)g.K, kSkinnyBN))) return rc;
    const int n_tiles = g.N / kSkinnyBN, kb_per_cta = g.K / BK / split;
    const int stages = std::min(kb_per_cta, kSkinnyMaxStages);
    CB_CUDA(launch_ex(kern, dim3(n_tiles * split), dim3(kSkinnyThreads), skinny_smem(stages), s, split, true, tmA, tmB, g,
                      kb_per_cta, stages));
    CB_LAUNCH_CHECK();
    return CB_OK;
}

// Split factor of a single-row-block GEMM.  MEASURED NEGATIVE RESULT (profiles/r02_latency_skinny_sweep.txt):
// single-query encode_image takes 0.476 ms with split 1, 0.529 / 0.585 / 0.601 ms with split 2 / 4 / 8 (general
// kernel: 0.487 ms).  The ~7 us per launch of a single-query forward pass is launch + prologue + one HBM round
// trip + epilogue, not the serial walk over K: the cluster barrier and the scratch round trip cost more than the
// shorter k-loop saves -- also when only the long-K shapes (c_proj, patch embed) are split: image unchanged,
// text 0.333 -> 0.373 ms.  The default is therefore NO split; the knob gemm_skinny = 2 / 4 / 8 forces one (tests).
int skinny_split(const GemmArgs &g, int sms) {
    (void)sms;
    if (g.skinny_scratch == nullptr) return 1;
    const int n_tiles = g.N / kSkinnyBN, kb = g.K / BK;
    if (tune(T_GEMM_SKINNY) > 0) {
        const int f = (int)tune(T_GEMM_SKINNY);
        if ((f == 1 || f == 2 || f == 4 || f == 8) && kb % f == 0 && (size_t)n_tiles * f * BM * kSkinnyBN <= kSkinnyScratchFloats) return f;
    }
    return 1;
}

int dispatch_skinny(const GemmArgs &g, int sms, cudaStream_t s) {
    const int split = skinny_split(g, sms);
    switch (g.epilogue) {
        case EPI_BIAS: return launch_skinny<EPI_BIAS>(g, split, s);
        case EPI_BIAS_GELU: return launch_skinny<EPI_BIAS_GELU>(g, split, s);
        case EPI_BIAS_RESID: return launch_skinny<EPI_BIAS_RESID>(g, split, s);
        case EPI_PATCH: return launch_skinny<EPI_PATCH>(g, split, s);
        case EPI_F32: return launch_skinny<EPI_F32>(g, split, s);
    }
    set_error("gemm_f16: unknown epilogue %d", g.epilogue);
    return CB_ERR_INVALID;
}

// Pick cluster size and N tile.  A CTA pair is used whenever the problem has more than one
// 128-row block (it halves the W traffic per SM and measured faster on every tower shape,
// profiles/r01_gemm_sweep.txt); the N tile is the one that minimises
// rounds-of-the-static-schedule x tile width / measured relative tile efficiency.
void pick_config(int M, int N, int sms, int *ncta_out, int *bn_out) {
    const int ncta = M > BM ? 2 : 1;
    if (M <= BM && N % 64 == 0) {
        // one row block (a single query, or a handful): the GEMM is a read of W, and how fast that goes
        // depends on how many SMs pull it - the narrowest tile gives N/64 CTAs (12..48 on the towers)
        *ncta_out = 1;
        *bn_out = 64;
        return;
    }
    const int m_tiles = (M + BM * ncta - 1) / (BM * ncta);
    const int slots = sms / ncta;
    double best_cost = 1e30;
    *ncta_out = ncta;
    *bn_out = 128;
    for (int bn : {256, 192, 128}) {
        if (N % bn) continue;
        const int tiles = m_tiles * (N / bn);
        const int rounds = (tiles + slots - 1) / slots;
        const double eff = bn == 256 ? 1.0 : (bn == 192 ? 0.90 : 0.66);
        const double cost = (double)rounds * bn / eff;
        if (cost < best_cost) { best_cost = cost; *bn_out = bn; }
    }
}

}  // namespace

int gemm_out_slices(int M, int N) {
    int dev = 0, sms = kNumSMs;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    int ncta = 1, bn = 128;
    pick_config(M, N, sms, &ncta, &bn);
    return N / (bn / 2);
}

int gemm_f16(const GemmArgs &g, cudaStream_t stream) {
    CB_REQUIRE(g.A && g.W && g.C, "gemm_f16: null operand");
    CB_REQUIRE(g.M > 0 && g.N > 0 && g.K > 0, "gemm_f16: empty shape");
    CB_REQUIRE(g.K % BK == 0, "gemm_f16: K=%d must be a multiple of %d", g.K, BK);
    CB_REQUIRE(g.N % 128 == 0, "gemm_f16: N=%d must be a multiple of 128", g.N);
    CB_REQUIRE(((uintptr_t)g.A & 15) == 0 && ((uintptr_t)g.W & 15) == 0 && ((uintptr_t)g.C & 15) == 0,
               "gemm_f16: operands must be 16-byte aligned");
    CB_REQUIRE(g.epilogue != EPI_BIAS_RESID || g.resid, "gemm_f16: residual epilogue without residual");
    CB_REQUIRE(g.epilogue != EPI_PATCH || g.pos, "gemm_f16: patch epilogue without pos-emb");
    CB_REQUIRE(!g.ln_stats || (g.colsum && g.ln_slices > 0 && (g.epilogue == EPI_BIAS || g.epilogue == EPI_BIAS_GELU)),
               "gemm_f16: LayerNorm fold needs colsum, ln_slices and a bias / bias+GELU epilogue");
    CB_REQUIRE(!g.stats_out || g.epilogue == EPI_BIAS_RESID, "gemm_f16: stats_out needs the residual epilogue");
    int dev = 0, sms = kNumSMs;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    int ncta = 1, bn = 128;
    pick_config(g.M, g.N, sms, &ncta, &bn);
    // one row block: the split-K cluster kernel (same 64-wide tile, same statistics layout as the general kernel)
    if (g.M <= BM && g.N % kSkinnyBN == 0 && tune(T_GEMM_SKINNY) != 0 && tune(T_GEMM_BN) <= 0 && tune(T_GEMM_NCTA) <= 0 &&
        ((uintptr_t)g.C & 15) == 0 && (g.ldc % 8) == 0)
        return dispatch_skinny(g, sms, stream);
    if (!g.stats_out) {      // test/experiment overrides (the stats layout depends on the default choice)
        if (tune(T_GEMM_BN) > 0) {
            int v = (int)tune(T_GEMM_BN);
            if ((v == 128 || v == 192 || v == 256 || (v == 64 && ncta == 1)) && g.N % v == 0) bn = v;
        }
        if (tune(T_GEMM_NCTA) > 0) {
            int v = (int)tune(T_GEMM_NCTA);
            if ((v == 1 || (v == 2 && g.M > BM)) && !(v == 2 && bn == 64)) ncta = v;
        }
    }
    if (ncta == 2) {
        switch (bn) {
            case 256: return dispatch_epi<256, 2>(g, stream);
            case 192: return dispatch_epi<192, 2>(g, stream);
            case 128: return dispatch_epi<128, 2>(g, stream);
        }
    } else {
        switch (bn) {
            case 256: return dispatch_epi<256, 1>(g, stream);
            case 192: return dispatch_epi<192, 1>(g, stream);
            case 128: return dispatch_epi<128, 1>(g, stream);
            case 64: return dispatch_epi<64, 1>(g, stream);
        }
    }
    set_error("gemm_f16: no tile shape for N=%d", g.N);
    return CB_ERR_INVALID;
}

}  // namespace cb

// the stand-alone entry points own no workspace: a single-row-block GEMM borrows its split-K scratch from the
// stream-ordered allocator for the duration of the launch
static int gemm_with_scratch(cb::GemmArgs &g, cudaStream_t s) {
    if (g.M > 128) return cb::gemm_f16(g, s);
    float *scratch = nullptr;
    CB_CUDA(cudaMallocAsync(&scratch, cb::kSkinnyScratchFloats * sizeof(float), s));
    g.skinny_scratch = scratch;
    const int rc = cb::gemm_f16(g, s);
    CB_CUDA(cudaFreeAsync(scratch, s));
    return rc;
}

extern "C" int cb_gemm_f16_ex_device(int M, int N, int K, const void *A, const void *W, const float *bias,
                                     const void *resid, void *C, int epilogue, const float *ln_stats, int ln_slices,
                                     const float *colsum, float *stats_out, void *stream) {
    cb::GemmArgs g;
    g.A = (const __half *)A; g.W = (const __half *)W; g.bias = bias; g.resid = (const __half *)resid;
    g.pos = nullptr; g.C = C; g.M = M; g.N = N; g.K = K; g.ldc = N; g.epilogue = epilogue;
    g.ln_stats = ln_stats; g.ln_slices = ln_slices; g.colsum = colsum; g.stats_out = stats_out;
    return gemm_with_scratch(g, (cudaStream_t)stream);
}

extern "C" int cb_gemm_out_slices(int M, int N) { return cb::gemm_out_slices(M, N); }

extern "C" int cb_gemm_f16_device(int M, int N, int K, const void *A, const void *W, const float *bias,
                                  const void *resid, const float *pos, void *C, int ldc, int epilogue,
                                  void *stream) {
    cb::GemmArgs g;
    g.A = (const __half *)A; g.W = (const __half *)W; g.bias = bias; g.resid = (const __half *)resid;
    g.pos = pos; g.C = C; g.M = M; g.N = N; g.K = K; g.ldc = ldc > 0 ? ldc : N; g.epilogue = epilogue;
    return gemm_with_scratch(g, (cudaStream_t)stream);
}
