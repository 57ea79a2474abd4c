// tcgen05 / TMEM / TMA GEMM for the CLIP towers (sm_100a).
//   C[M,N] = epilogue( A[M,K] (fp16, row-major) x W[N,K]^T (fp16, row-major = torch Linear weight) )
// fp32 accumulation in tensor memory; fused epilogues.
#pragma once
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace cb {

enum GemmEpilogue : int {
    EPI_BIAS = 0,        // C = acc + bias                      (in_proj)
    EPI_BIAS_GELU = 1,   // C = quickgelu(acc + bias)           (c_fc)      x*sigmoid(1.702x)
    EPI_BIAS_RESID = 2,  // C = resid + acc + bias              (out_proj, c_proj); C may alias resid
    EPI_PATCH = 3,       // C[row + row/49 + 1] = acc + pos[1 + row%49]   (conv1 as GEMM + pos-emb)
    EPI_F32 = 4,         // Cf32 = acc                          (visual.proj / text_projection)
};

struct GemmArgs {
    const __half *A;     // [M,K]
    const __half *W;     // [N,K]
    const float *bias;   // [N] fp32 (EPI_BIAS*, may be null = 0)
    const __half *resid; // [M,N] (EPI_BIAS_RESID)
    const float *pos;    // [50,N] fp32 (EPI_PATCH)
    void *C;             // fp16 [M,N] (row stride ldc) or fp32 for EPI_F32
    int M, N, K;
    int ldc;             // elements
    int epilogue;
    // LayerNorm folded into this GEMM (EPI_BIAS / EPI_BIAS_GELU): A holds the RAW residual
    // stream x, W holds gamma-scaled weights W'[n][k] = gamma[k] W[n][k], bias holds
    // b'[n] = b[n] + sum_k beta[k] W[n][k], and the epilogue computes
    //     y = rstd_r * (acc - mean_r * colsum[n]) + b'[n],   colsum[n] = sum_k W'[n][k]
    // from per-row partial sums ln_stats[r][s] = (sum x, sum x^2) over column slice s
    // (LayerNorm width = K, eps 1e-5).  Null = plain epilogue.
    const float *ln_stats = nullptr;
    int ln_slices = 0;
    const float *colsum = nullptr;
    // EPI_BIAS_RESID: also emit the per-row partial (sum, sum of squares) of the output (fp32
    // values before the fp16 store) over each epilogue warp's column slice: stats_out[r][gemm_out_slices(M,N)][2].
    // One writer per slot: deterministic, nothing to zero.
    float *stats_out = nullptr;
    // live timing (bench.py roofline): when set, every CTA folds %globaltimer into stamp[0] (min at
    // kernel entry) and stamp[1] (max at exit), so the launch's duration is measured on the device
    // without host-side event gaps.  Null in the product path.
    unsigned long long *stamp = nullptr;
    // fp32 scratch for the split-K form of single-row-block GEMMs (M <= 128): kSkinnyScratchFloats floats,
    // private to the stream the GEMM runs on.  Null: no split (each CTA walks all of K).
    float *skinny_scratch = nullptr;
};

// scratch size for GemmArgs::skinny_scratch: (N / 64) n-tiles x split factor <= 148 CTAs, 128 x 64 floats each
constexpr size_t kSkinnyScratchFloats = (size_t)148 * 128 * 64;

// returns a CB_* code; launches on `stream`
int gemm_f16(const GemmArgs &g, cudaStream_t stream);

// number of column slices a GEMM of this shape writes per row into stats_out
int gemm_out_slices(int M, int N);

}  // namespace cb
