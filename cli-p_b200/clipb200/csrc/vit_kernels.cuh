// HBM/L2-bound helper kernels of the CLIP towers (declarations).
#pragma once
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace cb {

// [B,224,224,3] u8 HWC -> patch-major im2col [B*49, 3072] fp16, column = py*96 + px*3 + c,
// value = (u8/255 - mean_c) / std_c          (clip._transform's ToTensor + Normalize)
int preprocess_u8(const uint8_t *img, __half *patches, int B, cudaStream_t s);
// [B,3,224,224] fp32 NCHW (already normalised by `transform`) -> same patch layout
int preprocess_f32(const float *img, __half *patches, int B, cudaStream_t s);

// LayerNorm over `width` (768 or 512) columns, eps 1e-5, fp32 statistics.
//   in row r  = in + (gather ? gather[r] : r * in_row_stride) * width
//   out row r = out + r * width
//   cls_fill != null: input rows with (row % cls_period == 0) are taken from cls_fill
//   (fp32 [width] = class_embedding + positional_embedding[0]) instead of memory.
//   stats_out != null: also writes (sum, sum of squares) of the fp16-rounded OUTPUT row to
//   stats_out[2*r .. 2*r+1] (consumed by a LayerNorm-folded GEMM, gemm.cuh)
int layernorm_f16(const __half *in, __half *out, const float *gamma, const float *beta, int rows, int width,
                  int in_row_stride, const int *gather, const float *cls_fill, int cls_period, cudaStream_t s,
                  float *stats_out = nullptr);

// rows of 512 fp32: out = in / ||in||  (no epsilon: build-index.py:50)
int l2norm_rows_f32(const float *in, float *out, int rows, int width, cudaStream_t s);

// text: x[b*77+t] = token_embedding[ids[b,t]] + positional_embedding[t] (fp16 out);
// eot_row[b] = b*77 + argmax_t ids[b,t] (first maximum)
// stats_out != null: (sum, sum of squares) of every fp16-rounded output row, as above
int text_embed(const int32_t *ids, const float *tok_emb, const float *pos_emb, __half *x, int *eot_row, int B,
               int ctx, int width, int vocab, cudaStream_t s, float *stats_out = nullptr);

// fused multi-head attention over packed qkv rows [B*L, 3*W] (q|k|v, heads of 64) ->
// out [B*L, W];  L = 50 (no mask) or 77 (causal)
int attention_f16(const __half *qkv, __half *out, int B, int L, int heads, bool causal, cudaStream_t s);

}  // namespace cb
