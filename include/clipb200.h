/* clipb200 -- C ABI of the B200-native hot paths behind CLI-P's surface.
 *
 * The reference (ps-auxw/CLI-P) has no FFI of its own: its boundary is Python
 * duck typing on two imported modules, `faiss` and `clip`.  Every entry point
 * below names the reference call it replaces (file:line under /root/reference).
 * Conventions (modelled on faiss's own c_api): int return code, 0 = OK,
 * non-zero = error with a thread-local message from cb_last_error(); caller
 * owns all in/out buffers; the library owns device-resident database shards and
 * model weights; ids are int64 positions in add() order.  No function aborts,
 * none has a CPU fallback: creating a handle without a CUDA device is an error.
 *
 * "_device" variants take device pointers and a cudaStream_t (passed as void*)
 * and never synchronise; the plain variants take HOST pointers, do the
 * host<->device copies themselves and return when the results are in host
 * memory.
 */
#ifndef CLIPB200_H
#define CLIPB200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define CB_OK 0
#define CB_ERR_INVALID 1   /* bad argument (shape, k <= 0, null pointer ...) */
#define CB_ERR_CUDA 2      /* a CUDA runtime/driver call failed */
#define CB_ERR_NOGPU 3     /* no CUDA device: there is no CPU fallback */
#define CB_ERR_OOM 4
#define CB_ERR_IO 5

/* storage dtype of database rows / dtype of a device source buffer */
#define CB_F32 0
#define CB_F16 1

typedef struct cb_index cb_index;

/* ---- misc ------------------------------------------------------------- */
const char *cb_last_error(void);
int cb_abi_version(void);
int cb_device_count(int *n);

/* Process-wide tuning knobs (tests, profiling).  Each knob is read ONCE from the environment variable
 * CLIPB200_<NAME IN CAPITALS> when the library is loaded; launch paths never call getenv.  -1 = default.
 *   gemm_bn, gemm_ncta, gemm_stages, gemm_raster   tcgen05 GEMM tile shape / ring depth / tile order
 *   batch_min_nq       smallest query batch served by the tensor-core search (default 2 on shards of
 *                      >= 1M rows, 16 below)
 *   no_graph           1: small encode_* calls are never replayed as CUDA graphs
 *   ln_blocks_per_sm   LayerNorm grid size
 *   ln_fold            0 / 1 / unset: see cb_clip_finalize
 *   pdl                0: the towers' kernels are launched without programmatic dependent launch
 *   gemm_skinny        0: single-row-block GEMMs (M <= 128) use the general kernel; 2 / 4 / 8: force a
 *                      cluster split-K of that factor (measured slower than no split; kept for tests)
 *   gemm_resid_stages  ring depth of the residual-epilogue GEMMs
 *   attn_tc            0: vision-tower attention runs the mma.sync kernel instead of the tcgen05 kernel that
 *                      puts two images of a head on one 128-row tile (tests compare the two)
 * Result-corrupting perf probes (gemm_debug, skip) exist only in -DCLIPB200_EXPERIMENTS builds. */
int cb_tuning_set(const char *name, int64_t value);
int cb_tuning_get(const char *name, int64_t *value);

/* ---- exact inner-product index (replaces faiss.IndexFlatIP) ----------- */

/* faiss.IndexFlatIP(512)                       build-index.py:80
 * d must be 512 (CLIP ViT-B/32 embedding width); storage is CB_F16 (1024
 * B/row, BASELINE configs 3-5) or CB_F32 (2048 B/row, what the reference
 * stores).  The shard lives on CUDA device `device`. */
int cb_flatip_create(int d, int storage_dtype, int device, cb_index **out);
void cb_flatip_free(cb_index *ix);

/* index.ntotal */
int64_t cb_flatip_ntotal(const cb_index *ix);
int cb_flatip_dim(const cb_index *ix);
int cb_flatip_storage_dtype(const cb_index *ix);
int cb_flatip_device(const cb_index *ix);

/* pre-size the shard (optional; add() grows geometrically otherwise) */
int cb_flatip_reserve(cb_index *ix, int64_t n_rows);

/* index.add(float32[n,512])                    build-index.py:99,107
 * rows are appended; ids are implicit and sequential from ntotal. */
int cb_flatip_add(cb_index *ix, int64_t n, const float *x_host);
/* rows already on the shard's device; queued on `stream`.  The index's own stream (the one the
 * host-pointer entry points run on) is ordered after the add.  Calls that touch one index from
 * several streams (add_device / search_device) must be ordered by the caller. */
int cb_flatip_add_device(cb_index *ix, int64_t n, const void *x_dev, int src_dtype,
                         void *stream);

/* drop all rows, keep the allocation (index.reset()) */
int cb_flatip_reset(cb_index *ix);

/* D, I = index.search(float32[nq,512], k)      query-index.py:111
 * D float32[nq,k] descending, I int64[nq,k]; unfilled slots I=-1,
 * D=-3.4028235e38; equal scores: lower id first.  id_base is added to every
 * returned id (the shard's first global row). */
int cb_flatip_search(cb_index *ix, int64_t nq, const float *q_host, int64_t k,
                     float *D_host, int64_t *I_host);
int cb_flatip_search_device(cb_index *ix, int64_t nq, const float *q_dev, int64_t k,
                            float *D_dev, int64_t *I_dev, int64_t id_base, void *stream);

/* search dispatch: one or two queries (or fp32 storage, or k > 1024, or a shard under 8192 rows)
 * stream the shard once per 4 queries (HBM-bound scan); larger batches on fp16 shards run a tcgen05
 * GEMM of the fp16-rounded queries with a fused per-query threshold filter, then re-score the
 * survivors exactly in fp32 (tensor-bound; flatip_batch.cu).  Both paths give the same answer bit
 * for bit and neither synchronises.  Counters for tests / benches: batch searches served by the
 * tensor-core path, and (query, row range) pairs whose candidate list overflowed and were
 * re-selected exactly on the device (adversarial row order; this call synchronises). */
int cb_flatip_batch_stats(cb_index *ix, int64_t *n_batch_searches, int64_t *n_rescued);

/* Merge R per-shard results into [nq][k] by (-score, id); ids < 0 are padding.
 * Shard r's block [nq][k] starts at D_in + r*shard_stride_D (elements) and
 * I_in + r*shard_stride_I; a stride <= 0 means densely packed [R][nq][k].  Runs
 * on the current device.  This is the step after the single cross-GPU gather
 * (SURVEY.md 8e), whose receive buffer holds each rank's D block and I block
 * back to back. */
int cb_topk_merge_device(int R, int64_t nq, int64_t k, const float *D_in,
                         const int64_t *I_in, int64_t shard_stride_D, int64_t shard_stride_I,
                         float *D_out, int64_t *I_out, void *stream);

/* ---- the database sharded over GPUs --------------------------------------------
 * Rows are split over the shards; a query is replicated, every shard selects its local top-k with
 * GLOBAL ids, and the block that finishes a query's list stores it straight into a mailbox in the
 * ROOT GPU's memory over NVLink (peer stores + a system-scope release on a per-rank counter).  The
 * root's merge kernel acquires the counters and merges by (-score, id): one kernel chain per GPU,
 * no collective, no host round trip, bit-identical to the unsharded answer.
 *
 * (1) one process per GPU (torchrun): every rank owns a cb_index; rank 0 allocates the mailbox and
 *     publishes its cudaIpc handle, the others open it. */
int cb_flatip_p2p_init(cb_index *ix, int rank, int world, int64_t max_elems, void *ipc_handle_out64);
int cb_flatip_p2p_connect(cb_index *ix, const void *root_ipc_handle64);
/* every rank calls this with the same (nq, k) sequence and the same queries; (D, I) are written on
 * rank 0 only (may be NULL elsewhere).  id_base = the shard's first global row.  nq*k is chunked to
 * max_elems per mailbox slot; k <= max_elems. */
int cb_flatip_search_p2p_device(cb_index *ix, int64_t nq, const float *q_dev, int64_t k, float *D_dev,
                                int64_t *I_dev, int64_t id_base, void *stream);
/* host-buffer form: q_host in, (D_host, I_host) out on rank 0 (may be NULL elsewhere); returns when done */
int cb_flatip_search_p2p(cb_index *ix, int64_t nq, const float *q_host, int64_t k, float *D_host,
                         int64_t *I_host, int64_t id_base);
/* Pipelined form for query streams: queue one search and return; results are ordered on a stream by
 * cb_flatip_join.  Searches alternate between two lanes (own stream, workspace and mailbox slot) and each
 * lane's kernels take half of the SMs' residency, so the selection / exchange / merge tail of one query
 * overlaps the pass over the shard of the next.  With an attached mailbox every rank submits the same
 * sequence and (D, I) are written on rank 0; without one it is the plain local search.  The query and
 * output buffers of a submitted search must stay untouched until the join; nq * k <= max_elems. */
int cb_flatip_submit_search_device(cb_index *ix, int64_t nq, const float *q_dev, int64_t k, float *D_dev,
                                   int64_t *I_dev, int64_t id_base, void *after_stream);
int cb_flatip_join(cb_index *ix, void *stream);
/* sticky error flag of the mailbox (a peer did not deliver within 20 s); synchronises */
int cb_flatip_p2p_status(cb_index *ix, int *error);

/* (2) ONE process driving several GPUs -- the REPL of query-index.py:29-30,111 with
 *     faiss.IndexFlatIP(devices=[...]): the same protocol between the devices of this process. */
typedef struct cb_sharded cb_sharded;
int cb_sharded_create(int d, int storage_dtype, int ndev, const int *devices, cb_sharded **out);
void cb_sharded_free(cb_sharded *s);
int64_t cb_sharded_ntotal(const cb_sharded *s);
int cb_sharded_num_shards(const cb_sharded *s);
cb_index *cb_sharded_shard(cb_sharded *s, int r);          /* borrowed: stats, timing, device rows */
int cb_sharded_reserve(cb_sharded *s, int64_t n_rows_total);
int cb_sharded_reset(cb_sharded *s);
/* index.add: the n rows get the next n ids and are split contiguously over the shards */
int cb_sharded_add(cb_sharded *s, int64_t n, const float *x_host);
/* rows already on shard `shard`'s device get the next n ids */
int cb_sharded_add_device(cb_sharded *s, int shard, int64_t n, const void *x_dev, int src_dtype, void *stream);
/* index.search: one call fans the query out, every device runs its chain, the root merges */
int cb_sharded_search(cb_sharded *s, int64_t nq, const float *q_host, int64_t k, float *D_host, int64_t *I_host);
/* q, D, I on devices[0]; `stream` is a stream of devices[0]; never synchronises */
int cb_sharded_search_device(cb_sharded *s, int64_t nq, const float *q_dev, int64_t k, float *D_dev,
                             int64_t *I_dev, void *stream);
int cb_sharded_get_rows(cb_sharded *s, int64_t start, int64_t n, float *out_host);

/* copy rows [start, start+n) back to the host as float32 (index.reconstruct_n;
 * used by write_index, build-index.py:109) */
int cb_flatip_get_rows(cb_index *ix, int64_t start, int64_t n, float *out_host);

/* device pointer of the shard (rows are contiguous, ntotal x d of the storage
 * dtype) -- for zero-copy ingest checks and benches */
const void *cb_flatip_device_rows(const cb_index *ix);

/* Live timing of the search kernel (scan + select in one cooperative launch; the dominant
 * kernel of a search): when enabled, every launch is bracketed by CUDA events on its own stream
 * (up to 256 launches); timing_read waits for them, returns the summed duration and count
 * and clears the list.  Used by bench.py for the roofline figure. */
int cb_flatip_timing(cb_index *ix, int enable);
int cb_flatip_timing_read(cb_index *ix, double *scan_ms_total, int *n_scans);
/* phases of the most recent timed launch of the search kernel, from %globaltimer stamps inside it:
 * ms3 = {pass over the shard up to the grid barrier, k-th-bin decision (+ radix levels if any),
 * gather + sort + write}.  Synchronises. */
int cb_flatip_phase_times(cb_index *ix, double *ms3);

/* ---- CLIP ViT-B/32 towers (replaces `import clip`'s model) ----------------- */
typedef struct cb_clip cb_clip;

/* model, transform = clip.load("ViT-B/32", device, jit=False)
 *                                   build-index.py:18, query-index.py:21
 * Creates an empty model on CUDA device `device` with activation workspace for
 * up to max_image_batch images / max_text_batch token rows per forward (larger
 * calls are processed in chunks).  Parameters are then supplied by their
 * openai/CLIP state-dict names as host float32 arrays (cb_clip_set_param; GEMM
 * weights are stored as fp16, everything else fp32, like clip.load does on a
 * CUDA device) and checked by cb_clip_finalize (all 302 tensors present with
 * ViT-B/32 shapes). */
int cb_clip_create(int device, int max_image_batch, int max_text_batch, cb_clip **out);
int cb_clip_set_param(cb_clip *m, const char *name, const float *data_host, int64_t numel);
int cb_clip_finalize(cb_clip *m);
void cb_clip_free(cb_clip *m);
/* ln_1 / ln_2 are folded into the QKV / c_fc GEMMs (gamma in the weights, row statistics from the
 * previous epilogue).  With the ln_fold knob unset, cb_clip_finalize runs a small calibration batch
 * through both forms and keeps the fold only if the embeddings agree to cosine >= 0.9998; otherwise
 * LayerNorm stays a separate fp32 launch, as in the reference.  Reports the decision. */
int cb_clip_ln_fold_status(cb_clip *m, int *folded, double *min_cosine);

/* image_features = model.encode_image(image)            build-index.py:49
 * (+ `/ image_features.norm(dim=-1, keepdim=True)`      build-index.py:50  when normalize != 0)
 * _u8 : B x 224 x 224 x 3 uint8 HWC pixels; clip._transform's ToTensor +
 *       Normalize(mean, std) run on the GPU, fused with the im2col that feeds the
 *       patch-embedding GEMM (build-index.py:48 for inputs that are already 224x224).
 * _f32: B x 3 x 224 x 224 float32 NCHW, exactly what `transform(image)` returns
 *       (the unchanged reference call).
 * out : B x 512 float32. */
int cb_clip_encode_image_u8_device(cb_clip *m, int64_t B, const uint8_t *hwc_dev, float *out_dev,
                                   int normalize, void *stream);
int cb_clip_encode_image_f32_device(cb_clip *m, int64_t B, const float *nchw_dev, float *out_dev,
                                    int normalize, void *stream);
int cb_clip_encode_image_u8(cb_clip *m, int64_t B, const uint8_t *hwc_host, float *out_host,
                            int normalize);

/* Pipelined forms for the index-time loop (build-index.py:30-58 re-shaped to
 * batches): submit() queues the work and returns.  Batches alternate between two
 * lanes (own activation workspace + streams), so the H2D copy of one batch
 * overlaps the forward pass of the other and the HBM-bound kernels of one pass
 * (LayerNorm, attention) overlap the tensor-bound GEMMs of the other.  B <=
 * max_image_batch.  Host variant: out_host is valid after cb_clip_sync(); host
 * buffers should be pinned and must outlive the sync.  Device variant: work is
 * ordered after `after_stream`; cb_clip_join(stream) makes `stream` wait for
 * all submitted work without blocking the host. */
int cb_clip_submit_image_u8(cb_clip *m, int64_t B, const uint8_t *hwc_host, float *out_host,
                            int normalize);
int cb_clip_submit_image_u8_device(cb_clip *m, int64_t B, const uint8_t *hwc_dev, float *out_dev,
                                   int normalize, void *after_stream);
int cb_clip_join(cb_clip *m, void *stream);
int cb_clip_sync(cb_clip *m);

/* text_features = model.encode_text(texts)              query-index.py:108
 * ids: B x 77 int32 tokens as produced by clip.tokenize (query-index.py:107);
 * the EOT position is argmax(ids) per row. out: B x 512 float32. */
int cb_clip_encode_text_device(cb_clip *m, int64_t B, const int32_t *ids_dev, float *out_dev,
                               int normalize, void *stream);
int cb_clip_encode_text(cb_clip *m, int64_t B, const int32_t *ids_host, float *out_host,
                        int normalize);

/* live timing of the GEMM launches inside encode_* (bench.py roofline): summed
 * CUDA-event duration, algorithmic FLOPs (2*M*N*K) and count since enable/read */
int cb_clip_timing(cb_clip *m, int enable);   /* 0 off, 1 GEMM launches, 2 every kernel class */
int cb_clip_timing_read(cb_clip *m, double *gemm_ms_total, double *gemm_flops, int *n_gemms);
/* call before cb_clip_timing_read: per timed GEMM launch, in launch order, its device-side duration
 * and its start relative to the first one (ms) */
int cb_clip_timing_launches(cb_clip *m, double *ms_out, double *t0_ms_out, int cap, int *n);
/* call before cb_clip_timing_read: summed live duration per kernel class since enable,
 * ms_by_class4 = {gemm, attention, layernorm, other} (level 2 only for the last three) */
int cb_clip_timing_breakdown(cb_clip *m, double *ms_by_class4);

/* building blocks of the towers, exported for unit tests (device pointers) */
int cb_layernorm_f16_device(const void *in, void *out, const float *gamma, const float *beta,
                            int rows, int width, int in_row_stride, const int *gather,
                            const float *cls_fill, int cls_period, void *stream);
int cb_attention_f16_device(const void *qkv, void *out, int B, int L, int heads, int causal,
                            void *stream);
int cb_preprocess_u8_device(const uint8_t *hwc, void *patches_f16, int B, void *stream);
int cb_preprocess_f32_device(const float *nchw, void *patches_f16, int B, void *stream);
int cb_l2norm_f32_device(const float *in, float *out, int rows, int width, void *stream);

/* transform(image), first half                           build-index.py:48
 * clip._transform's Resize(224, BICUBIC) on the shorter side + CenterCrop(224) for
 * one uint8 RGB image [h][w][3] already on the device -> [224][224][3] uint8.
 * Bit-identical to Pillow's ImagingResample (22-bit fixed-point weights,
 * horizontal then vertical pass); feed the result to cb_clip_encode_image_u8*. */
int cb_resize224_u8_device(const uint8_t *src_hwc, int h, int w, uint8_t *dst_224x224x3,
                           void *stream);

/* ---- JPEG files -> pixel batches in HBM (index-time caller of path A) -------
 * Image.open(tfn) + transform(image)                      build-index.py:47-48
 * for a whole batch of files in one call: `threads` host threads (0 = one per
 * core, at most 16), each with its own nvjpeg decoder state and CUDA stream,
 * read + entropy-decode the files and the GPU finishes them as interleaved RGB
 * straight into out_dev[i] (uint8 [224][224][3]); images of another size go
 * through cb_resize224_u8_device first.  nvjpeg is library code, loaded with
 * dlopen; pixels may differ from libjpeg-turbo's by +-1.
 * status[i]: 0 ok, 1 file unreadable, 2 not a decodable JPEG, 3 CUDA error,
 * 4 unsupported here (the caller decodes that file on the CPU instead).
 * The call returns when every out_dev[i] with status 0 is complete. */
typedef struct cb_jpeg cb_jpeg;
int cb_jpeg_create(int device, int threads, cb_jpeg **out);
void cb_jpeg_free(cb_jpeg *j);
int cb_jpeg_threads(const cb_jpeg *j);
int cb_jpeg_decode_files(cb_jpeg *j, int64_t n, const char *const *paths, uint8_t *out_dev,
                         int32_t *status);
int cb_jpeg_decode_memory(cb_jpeg *j, int64_t n, const uint8_t *const *data, const int64_t *sizes,
                          uint8_t *out_dev, int32_t *status);

/* ---- tcgen05 GEMM building block (exported for unit tests and benches) ----
 * C[M,N] = epilogue(A[M,K] fp16 row-major x W[N,K]^T fp16 row-major), fp32
 * accumulation in tensor memory.  Replaces the cuBLAS calls torch dispatches for
 * openai/CLIP's nn.Linear / conv1 / `@ proj` inside model.encode_image /
 * encode_text (build-index.py:49, query-index.py:108).  epilogue: 0 bias,
 * 1 bias+QuickGELU, 2 bias+residual (C may alias resid), 3 patch-embed scatter
 * + pos-emb, 4 plain fp32 output.  K % 64 == 0, N % 128 == 0; all device
 * pointers, 16-byte aligned; ldc in elements (0 = N). */
int cb_gemm_f16_device(int M, int N, int K, const void *A, const void *W, const float *bias,
                       const void *resid, const float *pos, void *C, int ldc, int epilogue,
                       void *stream);

/* Extended form used by the towers: LayerNorm folded into the GEMM and per-row statistics.
 *  - ln_stats != NULL (epilogue 0 or 1): A is the RAW residual stream x, W holds
 *    gamma-scaled weights, bias the beta-corrected bias, colsum[n] = sum_k W[n][k]; the
 *    epilogue computes rstd_r*(acc - mean_r*colsum[n]) + bias[n] from ln_stats[r][s] =
 *    (sum x, sum x^2) partial sums over ln_slices column slices (LayerNorm width K, eps 1e-5).
 *    Replaces ln_1 / ln_2 + in_proj / c_fc of openai/CLIP's ResidualAttentionBlock.
 *  - stats_out != NULL (epilogue 2): also writes the partial (sum, sum^2) of the fp16-rounded
 *    output per row and column slice, [M][cb_gemm_out_slices(M,N)][2] floats. */
int cb_gemm_f16_ex_device(int M, int N, int K, const void *A, const void *W, const float *bias,
                          const void *resid, void *C, int epilogue, const float *ln_stats,
                          int ln_slices, const float *colsum, float *stats_out, void *stream);
int cb_gemm_out_slices(int M, int N);

/* kernel launches issued by this library on the calling thread since the last
 * call with reset != 0 (bench.py's gpu_launches) */
int64_t cb_launch_count(int reset);

#ifdef __cplusplus
}
#endif
#endif /* CLIPB200_H */
