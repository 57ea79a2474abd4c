// CLIP ViT-B/32 towers on the tcgen05 GEMM + helper kernels: model handle, parameter
// loading by openai/CLIP state-dict name, encode_image / encode_text.
//
// Replaces `model.encode_image(image)` (/root/reference/build-index.py:49, with the
// L2 normalisation of :50 fused at the end) and `model.encode_text(texts)`
// (/root/reference/query-index.py:108).  Activations are fp16 with fp32 accumulation and
// fp32 LayerNorm statistics -- the precision the reference's own GPU path runs at
// (clip.load converts weights to fp16 on CUDA devices).
//
// HBM layout (B images, rows = B*50, W = 768):
//   patches [B*49, 3072] fp16   im2col, column = py*96 + px*3 + c (conv1 weight permuted to match)
//   x       [rows, W]    fp16   residual stream (token-major per image: class token first)
//   h       [rows, W]    fp16   LayerNorm output = GEMM A operand
//   qkv     [rows, 3W]   fp16   q | k | v, heads of 64 columns
//   att     [rows, W]    fp16
//   mlp     [rows, 4W]   fp16
//   cls     [B, W] fp16 -> emb [B, 512] fp32 -> out (L2-normalised)
// The text tower reuses the same buffers with rows = B*77, W = 512.
#include "common.cuh"
#include "gemm.cuh"
#include "vit_kernels.cuh"

#include <cmath>
#include <map>
#include <string>
#include <vector>

using namespace cb;

namespace {

struct Param {
    void *dev = nullptr;
    int64_t numel = 0;
    bool half = false;
};

struct LayerW {
    const __half *w_qkv, *w_o, *w_fc, *w_proj;
    const float *b_qkv, *b_o, *b_fc, *b_proj, *ln1_g, *ln1_b, *ln2_g, *ln2_b;
    // LayerNorm-folded form (gemm.cuh): w_qkv / w_fc hold gamma-scaled weights, b_qkv / b_fc the
    // beta-corrected biases, cs_* the per-output-column sums of the scaled weights
    const float *cs_qkv = nullptr, *cs_fc = nullptr;
};

constexpr int VW = 768, VL = 50, VH = 12, TW = 512, TL = 77, TH = 8, LAYERS = 12, ED = 512, VOCAB = 49408;

// activation workspace of one forward pass in flight
struct Ws {
    __half *patches = nullptr, *x = nullptr, *h = nullptr, *qkv = nullptr, *att = nullptr, *mlp = nullptr, *cls = nullptr;
    float *emb = nullptr;
    int *eot = nullptr;
    float *st1 = nullptr, *st2 = nullptr;   // per-row (sum, sum^2) slices for the folded ln_1 / ln_2
    float *skinny = nullptr;                // split-K partial tiles of single-row-block GEMMs (gemm.cuh)
};

constexpr int kMaxStatSlices = 24;      // 768 / 32: the BN = 64 tile of single-row-block GEMMs

// a lane = workspace + streams + staging for the pipelined submit API; two lanes run
// concurrently so the memory-bound kernels of one batch overlap the GEMMs of the other
struct Lane {
    Ws ws;
    cudaStream_t stream = nullptr, copy_stream = nullptr;
    uint8_t *img[2] = {nullptr, nullptr};   // two input slots: the copy of this lane's next batch does
    float *out = nullptr;                   // not have to wait for its current forward pass
    cudaEvent_t copied[2] = {nullptr, nullptr}, consumed[2] = {nullptr, nullptr}, done = nullptr;
    bool slot_used[2] = {false, false};
    int n_host = 0;
    bool used = false;
};

}  // namespace

struct cb_clip {
    int device = 0;
    int max_img = 0, max_txt = 0;
    std::map<std::string, Param> params;
    std::map<std::string, std::vector<float>> host;   // fp32 copies needed to fold LayerNorm at finalize
    bool ln_fold = true;          // ln_1 / ln_2 folded into the QKV / c_fc GEMMs (decided by cb_clip_finalize)
    double ln_fold_cos = -2.0;    // calibration result: min cosine folded vs unfolded (-2: not calibrated)
    bool finalized = false;
    LayerW vis[LAYERS], txt[LAYERS];
    const __half *conv1_w = nullptr, *vproj_w = nullptr, *tproj_w = nullptr;
    const float *vpos = nullptr, *ln_pre_g = nullptr, *ln_pre_b = nullptr, *ln_post_g = nullptr, *ln_post_b = nullptr;
    const float *tok_emb = nullptr, *tpos = nullptr, *lnf_g = nullptr, *lnf_b = nullptr;
    float *cls_pos = nullptr;     // class_embedding + positional_embedding[0]
    Ws ws;                        // workspace of the caller-stream (_device) entry points
    // staging for the host-pointer entry points
    uint8_t *d_img = nullptr;
    int32_t *d_ids = nullptr;
    float *d_out = nullptr;
    cudaStream_t stream = nullptr;
    // pipelined submit API: two lanes in flight
    Lane lanes[2];
    bool lanes_ready = false;
    int next_lane = 0;
    cudaEvent_t fence_ev = nullptr;
    // CUDA graphs of small forward passes (<= kGraphMaxBatch rows): a single query is ~70 tiny launches,
    // i.e. launch-bound; the replay costs one launch.  Keyed by (entry point, batch, normalize).
    struct GraphEntry { cudaGraphExec_t exec = nullptr; int launches = 0; bool seen = false, failed = false; };
    std::map<int, GraphEntry> graphs;
    float *d_img_f32 = nullptr;     // fixed input slot of the fp32-image graphs (allocated on first use)
    // live GEMM timing (bench.py roofline)
    int timing = 0;                 // 0 off, 1 GEMM launches only, 2 every kernel class
    unsigned long long *stamps = nullptr;   // device: (entry min, exit max) %globaltimer per timed GEMM launch
    int stamp_n = 0;
    static constexpr int kStamps = 8192;
    std::vector<cudaEvent_t> ev;
    std::vector<int> ev_cat;        // category of each event pair: 0 gemm, 1 attention, 2 layernorm, 3 other
    int ev_n = 0;
    double gemm_flops = 0;
};

namespace {

int upload(cb_clip *m, const std::string &name, const float *host, int64_t numel, bool as_half) {
    Param &p = m->params[name];
    if (p.dev) { cudaFree(p.dev); p.dev = nullptr; }
    p.numel = numel;
    p.half = as_half;
    if (as_half) {
        std::vector<__half> tmp((size_t)numel);
        for (int64_t i = 0; i < numel; i++) tmp[i] = __float2half_rn(host[i]);
        CB_CUDA(cudaMalloc(&p.dev, (size_t)numel * 2));
        CB_CUDA(cudaMemcpy(p.dev, tmp.data(), (size_t)numel * 2, cudaMemcpyHostToDevice));
    } else {
        CB_CUDA(cudaMalloc(&p.dev, (size_t)std::max<int64_t>(numel, 1) * 4));
        CB_CUDA(cudaMemcpy(p.dev, host, (size_t)numel * 4, cudaMemcpyHostToDevice));
    }
    return CB_OK;
}

bool ends_with(const std::string &s, const char *suf) {
    const size_t n = strlen(suf);
    return s.size() >= n && s.compare(s.size() - n, n, suf) == 0;
}

template <typename T>
int need(cb_clip *m, const std::string &name, int64_t numel, bool half, const T **out) {
    auto it = m->params.find(name);
    if (it == m->params.end()) { set_error("cb_clip_finalize: parameter %s was never set", name.c_str()); return CB_ERR_INVALID; }
    if (it->second.numel != numel || it->second.half != half) {
        set_error("cb_clip_finalize: parameter %s has %lld elements, expected %lld", name.c_str(),
                  (long long)it->second.numel, (long long)numel);
        return CB_ERR_INVALID;
    }
    *out = reinterpret_cast<const T *>(it->second.dev);
    return CB_OK;
}

// W'[n][k] = gamma[k] W[n][k] (rounded to fp16), colsum[n] = sum_k W'[n][k] (of the rounded values),
// b'[n] = b[n] + sum_k beta[k] W[n][k]; replaces the raw weight / bias on the device
int fold_one(cb_clip *m, const std::string &wname, const std::string &bname, const std::string &gname,
             const std::string &btname, const std::string &csname, int N, int K) {
    auto need_host = [&](const std::string &n, size_t cnt) -> const std::vector<float> * {
        auto it = m->host.find(n);
        if (it == m->host.end() || it->second.size() != cnt) return nullptr;
        return &it->second;
    };
    const std::vector<float> *Wv = need_host(wname, (size_t)N * K), *bv = need_host(bname, N);
    const std::vector<float> *gv = need_host(gname, K), *btv = need_host(btname, K);
    if (!Wv || !bv || !gv || !btv) {
        set_error("cb_clip_finalize: %s / %s / %s / %s missing or mis-sized", wname.c_str(), bname.c_str(),
                  gname.c_str(), btname.c_str());
        return CB_ERR_INVALID;
    }
    std::vector<float> Wf((size_t)N * K), cs(N), bf(N);
    for (int n = 0; n < N; n++) {
        double s_cs = 0.0, s_b = 0.0;
        const float *wr = Wv->data() + (size_t)n * K;
        for (int k = 0; k < K; k++) {
            const float wp = __half2float(__float2half_rn((*gv)[k] * wr[k]));
            Wf[(size_t)n * K + k] = wp;
            s_cs += wp;
            s_b += (double)(*btv)[k] * wr[k];
        }
        cs[n] = (float)s_cs;
        bf[n] = (*bv)[n] + (float)s_b;
    }
    int rc;
    if ((rc = upload(m, wname, Wf.data(), (int64_t)N * K, true))) return rc;
    if ((rc = upload(m, bname, bf.data(), N, false))) return rc;
    return upload(m, csname, cs.data(), N, false);
}

int fold_layernorm(cb_clip *m, const char *prefix, int W) {
    for (int i = 0; i < LAYERS; i++) {
        const std::string b = std::string(prefix) + ".resblocks." + std::to_string(i) + ".";
        int rc;
        if ((rc = fold_one(m, b + "attn.in_proj_weight", b + "attn.in_proj_bias", b + "ln_1.weight", b + "ln_1.bias",
                           b + "attn.in_proj_colsum", 3 * W, W))) return rc;
        if ((rc = fold_one(m, b + "mlp.c_fc.weight", b + "mlp.c_fc.bias", b + "ln_2.weight", b + "ln_2.bias",
                           b + "mlp.c_fc_colsum", 4 * W, W))) return rc;
    }
    return CB_OK;
}

int bind_blocks(cb_clip *m, const char *prefix, int W, LayerW *lw) {
    for (int i = 0; i < LAYERS; i++) {
        const std::string b = std::string(prefix) + ".resblocks." + std::to_string(i) + ".";
        int rc = 0;
        rc |= need(m, b + "attn.in_proj_weight", 3ll * W * W, true, &lw[i].w_qkv);
        rc |= need(m, b + "attn.out_proj.weight", 1ll * W * W, true, &lw[i].w_o);
        rc |= need(m, b + "mlp.c_fc.weight", 4ll * W * W, true, &lw[i].w_fc);
        rc |= need(m, b + "mlp.c_proj.weight", 4ll * W * W, true, &lw[i].w_proj);
        rc |= need(m, b + "attn.in_proj_bias", 3ll * W, false, &lw[i].b_qkv);
        rc |= need(m, b + "attn.out_proj.bias", 1ll * W, false, &lw[i].b_o);
        rc |= need(m, b + "mlp.c_fc.bias", 4ll * W, false, &lw[i].b_fc);
        rc |= need(m, b + "mlp.c_proj.bias", 1ll * W, false, &lw[i].b_proj);
        rc |= need(m, b + "ln_1.weight", W, false, &lw[i].ln1_g);
        rc |= need(m, b + "ln_1.bias", W, false, &lw[i].ln1_b);
        rc |= need(m, b + "ln_2.weight", W, false, &lw[i].ln2_g);
        rc |= need(m, b + "ln_2.bias", W, false, &lw[i].ln2_b);
        if (m->ln_fold) {
            rc |= need(m, b + "attn.in_proj_colsum", 3ll * W, false, &lw[i].cs_qkv);
            rc |= need(m, b + "mlp.c_fc_colsum", 4ll * W, false, &lw[i].cs_fc);
        }
        if (rc) return CB_ERR_INVALID;
    }
    return CB_OK;
}

// GEMM launches are timed ON THE DEVICE (every CTA folds %globaltimer into a per-launch (min entry, max
// exit) pair): the duration of a launch then excludes the host-side gaps that CUDA events bracket in
int timed_gemm(cb_clip *m, const Ws &w, const GemmArgs &g0, cudaStream_t s) {
    GemmArgs g = g0;
    g.skinny_scratch = w.skinny;
    if (!m->timing || m->stamp_n >= cb_clip::kStamps) return gemm_f16(g, s);
    g.stamp = m->stamps + 2 * (size_t)m->stamp_n;
    int rc = gemm_f16(g, s);
    if (rc) return rc;
    m->stamp_n++;
    m->gemm_flops += 2.0 * g.M * g.N * g.K;
    return CB_OK;
}

// bracket a non-GEMM launch with events when the full breakdown is requested
template <typename F>
int timed_other(cb_clip *m, int cat, cudaStream_t s, F &&launch) {
    const bool t = m->timing >= 2 && m->ev_n + 2 <= (int)m->ev.size();
    if (t) CB_CUDA(cudaEventRecord(m->ev[m->ev_n], s));
    int rc = launch();
    if (rc) return rc;
    if (t) {
        CB_CUDA(cudaEventRecord(m->ev[m->ev_n + 1], s));
        m->ev_cat[m->ev_n / 2] = cat;
        m->ev_n += 2;
    }
    return CB_OK;
}

GemmArgs mk(const __half *A, const __half *W, const float *bias, const __half *resid, void *C, int M, int N, int K, int epi) {
    GemmArgs g;
    g.A = A; g.W = W; g.bias = bias; g.resid = resid; g.pos = nullptr; g.C = C;
    g.M = M; g.N = N; g.K = K; g.ldc = N; g.epilogue = epi;
    return g;
}

// 12 residual attention blocks over x [B*L, W] (in place)
int run_blocks(cb_clip *m, Ws &w, const LayerW *lw, int W, int heads, int B, int L, bool causal, cudaStream_t s) {
    const int rows = B * L;
#ifdef CLIPB200_EXPERIMENTS
    // perf experiments only (results become wrong): skip=1 drops ln_1/ln_2, =2 drops attention
    const int skip = tune(T_SKIP) > 0 ? (int)tune(T_SKIP) : 0;
#else
    constexpr int skip = 0;
#endif
    if (m->ln_fold) {
        // ln_1 / ln_2 live inside the QKV / c_fc GEMMs: A operand = raw residual stream, per-row
        // statistics come from whoever wrote x last (ln_pre / text_embed for block 0, then the
        // residual GEMM epilogues), w.st1 feeds ln_1 and w.st2 feeds ln_2
        const int res_slices = gemm_out_slices(rows, W);
        for (int i = 0; i < LAYERS; i++) {
            int rc;
            GemmArgs g = mk(w.x, lw[i].w_qkv, lw[i].b_qkv, nullptr, w.qkv, rows, 3 * W, W, EPI_BIAS);
            g.ln_stats = w.st1; g.ln_slices = i == 0 ? 1 : res_slices; g.colsum = lw[i].cs_qkv;
            if ((rc = timed_gemm(m, w, g, s))) return rc;
            if (!(skip & 2))
            if ((rc = timed_other(m, 1, s, [&] { return attention_f16(w.qkv, w.att, B, L, heads, causal, s); }))) return rc;
            g = mk(w.att, lw[i].w_o, lw[i].b_o, w.x, w.x, rows, W, W, EPI_BIAS_RESID);
            g.stats_out = w.st2;
            if ((rc = timed_gemm(m, w, g, s))) return rc;
            g = mk(w.x, lw[i].w_fc, lw[i].b_fc, nullptr, w.mlp, rows, 4 * W, W, EPI_BIAS_GELU);
            g.ln_stats = w.st2; g.ln_slices = res_slices; g.colsum = lw[i].cs_fc;
            if ((rc = timed_gemm(m, w, g, s))) return rc;
            g = mk(w.mlp, lw[i].w_proj, lw[i].b_proj, w.x, w.x, rows, W, 4 * W, EPI_BIAS_RESID);
            g.stats_out = w.st1;
            if ((rc = timed_gemm(m, w, g, s))) return rc;
        }
        return CB_OK;
    }
    for (int i = 0; i < LAYERS; i++) {
        int rc;
        if (!(skip & 1))
        if ((rc = timed_other(m, 2, s, [&] { return layernorm_f16(w.x, w.h, lw[i].ln1_g, lw[i].ln1_b, rows, W, 1, nullptr, nullptr, 0, s); }))) return rc;
        if ((rc = timed_gemm(m, w, mk(w.h, lw[i].w_qkv, lw[i].b_qkv, nullptr, w.qkv, rows, 3 * W, W, EPI_BIAS), s))) return rc;
        if (!(skip & 2))
        if ((rc = timed_other(m, 1, s, [&] { return attention_f16(w.qkv, w.att, B, L, heads, causal, s); }))) return rc;
        if ((rc = timed_gemm(m, w, mk(w.att, lw[i].w_o, lw[i].b_o, w.x, w.x, rows, W, W, EPI_BIAS_RESID), s))) return rc;
        if (!(skip & 1))
        if ((rc = timed_other(m, 2, s, [&] { return layernorm_f16(w.x, w.h, lw[i].ln2_g, lw[i].ln2_b, rows, W, 1, nullptr, nullptr, 0, s); }))) return rc;
        if ((rc = timed_gemm(m, w, mk(w.h, lw[i].w_fc, lw[i].b_fc, nullptr, w.mlp, rows, 4 * W, W, EPI_BIAS_GELU), s))) return rc;
        if ((rc = timed_gemm(m, w, mk(w.mlp, lw[i].w_proj, lw[i].b_proj, w.x, w.x, rows, W, 4 * W, EPI_BIAS_RESID), s))) return rc;
    }
    return CB_OK;
}

cudaError_t alloc_ws(const cb_clip *m, Ws &w) {
    // the larger of the two towers, element counts
    const size_t rows_v = (size_t)m->max_img * VL, rows_t = (size_t)m->max_txt * TL;
    const size_t n_x = std::max(rows_v * VW, rows_t * TW);
    const size_t n_cls = std::max((size_t)m->max_img * VW, (size_t)m->max_txt * TW);
    const size_t nb = std::max(m->max_img, m->max_txt);
    cudaError_t e = cudaSuccess;
    auto A = [&](void **p, size_t bytes) { if (e == cudaSuccess) e = cudaMalloc(p, std::max<size_t>(bytes, 256)); };
    A((void **)&w.patches, (size_t)m->max_img * 49 * 3072 * 2);
    A((void **)&w.x, n_x * 2); A((void **)&w.h, n_x * 2); A((void **)&w.att, n_x * 2);
    A((void **)&w.qkv, 3 * n_x * 2); A((void **)&w.mlp, 4 * n_x * 2);
    A((void **)&w.cls, n_cls * 2); A((void **)&w.emb, nb * ED * 4); A((void **)&w.eot, nb * 4);
    const size_t rows = std::max(rows_v, rows_t);
    A((void **)&w.st1, rows * kMaxStatSlices * 8); A((void **)&w.st2, rows * kMaxStatSlices * 8);
    A((void **)&w.skinny, kSkinnyScratchFloats * 4);
    return e;
}

void free_ws(Ws &w) {
    void *bufs[] = {w.patches, w.x, w.h, w.qkv, w.att, w.mlp, w.cls, w.emb, w.eot, w.st1, w.st2, w.skinny};
    for (void *p : bufs) cudaFree(p);
    w = Ws();
}

// patches (already in w.patches) -> out [B,512] fp32
int vision_from_patches(cb_clip *m, Ws &w, int B, float *out_dev, int normalize, cudaStream_t s) {
    int rc;
    GemmArgs g = mk(w.patches, m->conv1_w, nullptr, nullptr, w.x, B * 49, VW, 3072, EPI_PATCH);
    g.pos = m->vpos;
    if ((rc = timed_gemm(m, w, g, s))) return rc;
    // ln_pre in place; class-token rows (row % 50 == 0) come from class_embedding + pos[0]
    if ((rc = timed_other(m, 2, s, [&] { return layernorm_f16(w.x, w.x, m->ln_pre_g, m->ln_pre_b, B * VL, VW, 1, nullptr, m->cls_pos, VL, s,
                                                                  m->ln_fold ? w.st1 : nullptr); }))) return rc;
    if ((rc = run_blocks(m, w, m->vis, VW, VH, B, VL, false, s))) return rc;
    if ((rc = layernorm_f16(w.x, w.cls, m->ln_post_g, m->ln_post_b, B, VW, VL, nullptr, nullptr, 0, s))) return rc;
    float *emb = normalize ? w.emb : out_dev;
    if ((rc = timed_gemm(m, w, mk(w.cls, m->vproj_w, nullptr, nullptr, emb, B, ED, VW, EPI_F32), s))) return rc;
    if (normalize && (rc = l2norm_rows_f32(emb, out_dev, B, ED, s))) return rc;
    return CB_OK;
}

int encode_u8_chunk(cb_clip *m, Ws &w, int b, const uint8_t *hwc_dev, float *out_dev, int normalize, cudaStream_t s) {
    int rc = timed_other(m, 3, s, [&] { return preprocess_u8(hwc_dev, w.patches, b, s); });
    if (rc) return rc;
    return vision_from_patches(m, w, b, out_dev, normalize, s);
}

int text_forward(cb_clip *m, Ws &w, int B, const int32_t *ids_dev, float *out_dev, int normalize, cudaStream_t s) {
    int rc;
    if ((rc = text_embed(ids_dev, m->tok_emb, m->tpos, w.x, w.eot, B, TL, TW, VOCAB, s, m->ln_fold ? w.st1 : nullptr))) return rc;
    if ((rc = run_blocks(m, w, m->txt, TW, TH, B, TL, true, s))) return rc;
    // ln_final only on the EOT rows (LayerNorm is per-row, so gathering first is exact)
    if ((rc = layernorm_f16(w.x, w.cls, m->lnf_g, m->lnf_b, B, TW, 1, w.eot, nullptr, 0, s))) return rc;
    float *emb = normalize ? w.emb : out_dev;
    if ((rc = timed_gemm(m, w, mk(w.cls, m->tproj_w, nullptr, nullptr, emb, B, ED, TW, EPI_F32), s))) return rc;
    if (normalize && (rc = l2norm_rows_f32(emb, out_dev, B, ED, s))) return rc;
    return CB_OK;
}

constexpr double kFoldMinCos = 0.9998;   // folded vs unfolded embeddings on the calibration batch

// restore the raw in_proj / c_fc weights and biases from the host copies (undo fold_layernorm)
int unfold_layernorm(cb_clip *m, const char *prefix, int W) {
    for (int i = 0; i < LAYERS; i++) {
        const std::string b = std::string(prefix) + ".resblocks." + std::to_string(i) + ".";
        for (const char *suffix : {"attn.in_proj_weight", "attn.in_proj_bias", "mlp.c_fc.weight", "mlp.c_fc.bias"}) {
            auto it = m->host.find(b + suffix);
            if (it == m->host.end()) { set_error("cb_clip_finalize: host copy of %s%s missing", b.c_str(), suffix); return CB_ERR_INVALID; }
            const bool w = ends_with(it->first, "weight");
            int rc = upload(m, it->first, it->second.data(), (int64_t)it->second.size(), w);
            if (rc) return rc;
        }
    }
    return CB_OK;
}

// A small deterministic forward pass through both towers with the CURRENT binding (folded or not):
// four images (noise, ramp, checkerboard, flat grey + noise) and four token rows.  out = the
// un-normalised embeddings, [rows][512].
int calibration_pass(cb_clip *m, std::vector<float> &out) {
    out.clear();
    const bool was = m->finalized;
    m->finalized = true;
    int rc = CB_OK;
    uint32_t lcg = 12345u;
    auto rnd = [&]() { lcg = lcg * 1664525u + 1013904223u; return lcg >> 24; };
    const int nb = std::min(4, m->max_img);
    if (nb > 0) {
        std::vector<uint8_t> img((size_t)nb * 224 * 224 * 3);
        for (int b = 0; b < nb; b++)
            for (int y = 0; y < 224; y++)
                for (int x = 0; x < 224; x++)
                    for (int c = 0; c < 3; c++) {
                        uint32_t v;
                        switch (b) {
                            case 0: v = rnd(); break;
                            case 1: v = (uint32_t)((x + y * 2 + c * 40) * 255 / (224 * 3 + 80)); break;
                            case 2: v = (((x >> 4) ^ (y >> 4)) & 1) ? 230u - 20u * c : 25u + 30u * c; break;
                            default: v = 118u + (rnd() & 15u); break;
                        }
                        img[(((size_t)b * 224 + y) * 224 + x) * 3 + c] = (uint8_t)std::min(v, 255u);
                    }
        CB_CUDA(cudaMemcpyAsync(m->d_img, img.data(), img.size(), cudaMemcpyHostToDevice, m->stream));
        rc = encode_u8_chunk(m, m->ws, nb, m->d_img, m->d_out, 0, m->stream);
        if (rc == CB_OK) {
            out.resize((size_t)nb * ED);
            CB_CUDA(cudaMemcpyAsync(out.data(), m->d_out, out.size() * 4, cudaMemcpyDeviceToHost, m->stream));
            CB_CUDA(cudaStreamSynchronize(m->stream));
        }
    }
    const int nt = std::min(4, m->max_txt);
    if (rc == CB_OK && nt > 0) {
        std::vector<int32_t> ids((size_t)nt * TL, 0);
        for (int b = 0; b < nt; b++) {
            const int len = 3 + b * 5;
            ids[(size_t)b * TL] = 49406;
            for (int t = 1; t <= len; t++) ids[(size_t)b * TL + t] = 1000 + (int)((rnd() * 151u + (uint32_t)t * 977u) % 39000u);
            ids[(size_t)b * TL + len + 1] = 49407;
        }
        CB_CUDA(cudaMemcpyAsync(m->d_ids, ids.data(), ids.size() * 4, cudaMemcpyHostToDevice, m->stream));
        rc = text_forward(m, m->ws, nt, m->d_ids, m->d_out, 0, m->stream);
        if (rc == CB_OK) {
            const size_t at = out.size();
            out.resize(at + (size_t)nt * ED);
            CB_CUDA(cudaMemcpyAsync(out.data() + at, m->d_out, (size_t)nt * ED * 4, cudaMemcpyDeviceToHost, m->stream));
            CB_CUDA(cudaStreamSynchronize(m->stream));
        }
    }
    m->finalized = was;
    return rc;
}

constexpr int kGraphMaxBatch = 32;

void drop_graphs(cb_clip *m) {
    for (auto &kv : m->graphs)
        if (kv.second.exec) cudaGraphExecDestroy(kv.second.exec);
    m->graphs.clear();
}

// Run `body(in, out, stream)` - one small forward pass - as a CUDA graph replay.
// The graph is captured once per (kind, b, normalize) on the handle's own stream with FIXED
// buffers (`slot_in` -> workspace -> m->d_out); a replay on the caller's stream is
//   copy in_dev -> slot_in (skipped when the caller already passed the slot), graph launch,
//   copy m->d_out -> out_dev.
// The first call of a key runs the plain launches (it also sets the kernels' attributes); timing
// mode and CLIPB200_NO_GRAPH=1 always do.  A failed capture falls back to plain launches for good.
template <typename Body>
int graphed_forward(cb_clip *m, int kind, int b, int normalize, const void *in_dev, void *slot_in, size_t in_bytes,
                    float *out_dev, cudaStream_t s, Body &&body) {
    const bool off = tune(T_NO_GRAPH) > 0;
    if (off || m->timing || b > kGraphMaxBatch) return body(in_dev, out_dev, s);
    cb_clip::GraphEntry &e = m->graphs[(kind << 16) | (b << 1) | (normalize ? 1 : 0)];
    if (!e.seen || e.failed) {
        e.seen = true;
        return body(in_dev, out_dev, s);
    }
    if (!e.exec) {
        const int64_t before = g_launches;
        cudaGraph_t graph = nullptr;
        cudaError_t ce = cudaStreamBeginCapture(m->stream, cudaStreamCaptureModeThreadLocal);
        int rc = ce == cudaSuccess ? body(slot_in, m->d_out, m->stream) : CB_ERR_CUDA;
        if (ce == cudaSuccess) ce = cudaStreamEndCapture(m->stream, &graph);
        if (rc == CB_OK && ce == cudaSuccess && graph) ce = cudaGraphInstantiate(&e.exec, graph, 0);
        if (graph) cudaGraphDestroy(graph);
        e.launches = (int)(g_launches - before);
        g_launches = before;                       // a capture launches nothing
        if (rc != CB_OK || ce != cudaSuccess || !e.exec) {
            cudaGetLastError();
            e.exec = nullptr;
            e.failed = true;
            return body(in_dev, out_dev, s);
        }
    }
    if (in_dev != slot_in) CB_CUDA(cudaMemcpyAsync(slot_in, in_dev, in_bytes, cudaMemcpyDeviceToDevice, s));
    CB_CUDA(cudaGraphLaunch(e.exec, s));
    g_launches += e.launches;
    if (out_dev != m->d_out)
        CB_CUDA(cudaMemcpyAsync(out_dev, m->d_out, (size_t)b * ED * 4, cudaMemcpyDeviceToDevice, s));
    return CB_OK;
}

}  // namespace

extern "C" {

int cb_clip_create(int device, int max_image_batch, int max_text_batch, cb_clip **out) {
    CB_REQUIRE(out != nullptr, "cb_clip_create: out is null");
    *out = nullptr;
    CB_REQUIRE(max_image_batch >= 0 && max_text_batch >= 0 && max_image_batch + max_text_batch > 0,
               "cb_clip_create: batch capacities must be >= 0 and not both 0");
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
        cudaGetLastError();
        set_error("cb_clip_create: no CUDA device (this library has no CPU fallback)");
        return CB_ERR_NOGPU;
    }
    CB_REQUIRE(device >= 0 && device < ndev, "cb_clip_create: device %d out of range", device);
    DeviceGuard g(device);
    cb_clip *m = new (std::nothrow) cb_clip();
    if (!m) { set_error("out of host memory"); return CB_ERR_OOM; }
    m->device = device;
    m->max_img = max_image_batch;
    m->max_txt = max_text_batch;
    const size_t nb = std::max(max_image_batch, max_text_batch);
    cudaError_t e = alloc_ws(m, m->ws);
    auto A = [&](void **p, size_t bytes) { if (e == cudaSuccess) e = cudaMalloc(p, std::max<size_t>(bytes, 256)); };
    A((void **)&m->d_img, (size_t)max_image_batch * 224 * 224 * 3);
    A((void **)&m->d_ids, (size_t)max_text_batch * TL * 4);
    A((void **)&m->d_out, nb * ED * 4);
    A((void **)&m->cls_pos, VW * 4);
    if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&m->stream, cudaStreamNonBlocking);
    if (e != cudaSuccess) {
        set_error("cb_clip_create: %s", cudaGetErrorString(e));
        cb_clip_free(m);
        return e == cudaErrorMemoryAllocation ? CB_ERR_OOM : CB_ERR_CUDA;
    }
    *out = m;
    return CB_OK;
}

void cb_clip_free(cb_clip *m) {
    if (!m) return;
    DeviceGuard g(m->device);
    if (m->stream) cudaStreamSynchronize(m->stream);
    drop_graphs(m);
    for (auto &kv : m->params) cudaFree(kv.second.dev);
    free_ws(m->ws);
    void *bufs[] = {m->d_img, m->d_ids, m->d_out, m->cls_pos, m->d_img_f32, m->stamps};
    for (void *p : bufs) cudaFree(p);
    for (cudaEvent_t ev : m->ev) cudaEventDestroy(ev);
    for (Lane &l : m->lanes) {
        if (l.stream) cudaStreamSynchronize(l.stream);
        free_ws(l.ws);
        cudaFree(l.out);
        for (int i = 0; i < 2; i++) {
            cudaFree(l.img[i]);
            if (l.copied[i]) cudaEventDestroy(l.copied[i]);
            if (l.consumed[i]) cudaEventDestroy(l.consumed[i]);
        }
        if (l.done) cudaEventDestroy(l.done);
        if (l.copy_stream) cudaStreamDestroy(l.copy_stream);
        if (l.stream) cudaStreamDestroy(l.stream);
    }
    if (m->fence_ev) cudaEventDestroy(m->fence_ev);
    if (m->stream) cudaStreamDestroy(m->stream);
    delete m;
}

int cb_clip_set_param(cb_clip *m, const char *name_c, const float *host, int64_t numel) {
    CB_REQUIRE(m && name_c && host, "cb_clip_set_param: null argument");
    CB_REQUIRE(numel >= 0, "cb_clip_set_param: numel < 0");
    DeviceGuard g(m->device);
    const std::string name(name_c);
    m->finalized = false;
    if (name == "visual.conv1.weight") {
        // [768][c][py][px] -> [768][py][px][c] so a patch row is 96 contiguous values
        CB_REQUIRE(numel == 768ll * 3072, "visual.conv1.weight: expected 768*3*32*32 elements");
        std::vector<float> t((size_t)numel);
        for (int o = 0; o < 768; o++)
            for (int c = 0; c < 3; c++)
                for (int p = 0; p < 1024; p++) t[(size_t)o * 3072 + p * 3 + c] = host[(size_t)o * 3072 + c * 1024 + p];
        return upload(m, name, t.data(), numel, true);
    }
    if (name == "visual.proj" || name == "text_projection") {
        // stored [in, 512] and applied as x @ P: the GEMM wants [512, in] (K-major)
        CB_REQUIRE(numel % ED == 0, "%s: element count not a multiple of 512", name_c);
        const int64_t in = numel / ED;
        std::vector<float> t((size_t)numel);
        for (int64_t i = 0; i < in; i++)
            for (int64_t o = 0; o < ED; o++) t[(size_t)o * in + i] = host[(size_t)i * ED + o];
        return upload(m, name, t.data(), numel, true);
    }
    const bool gemm_w = ends_with(name, "in_proj_weight") || ends_with(name, "out_proj.weight") ||
                        ends_with(name, "c_fc.weight") || ends_with(name, "c_proj.weight");
    // everything LayerNorm folding needs at finalize: the two consuming weights + biases and ln_1 / ln_2
    if (name.find(".resblocks.") != std::string::npos &&
        (ends_with(name, "in_proj_weight") || ends_with(name, "in_proj_bias") || ends_with(name, "c_fc.weight") ||
         ends_with(name, "c_fc.bias") || name.find(".ln_1.") != std::string::npos || name.find(".ln_2.") != std::string::npos))
        m->host[name].assign(host, host + numel);
    return upload(m, name, host, numel, gemm_w);
}

int cb_clip_finalize(cb_clip *m) {
    CB_REQUIRE(m != nullptr, "cb_clip_finalize: null handle");
    DeviceGuard g(m->device);
    drop_graphs(m);                 // captured graphs hold the old parameter pointers
    int rc = 0;
    const float *cls_emb = nullptr;
    rc |= need(m, "visual.conv1.weight", 768ll * 3072, true, &m->conv1_w);
    rc |= need(m, "visual.class_embedding", VW, false, &cls_emb);
    rc |= need(m, "visual.positional_embedding", 1ll * VL * VW, false, &m->vpos);
    rc |= need(m, "visual.ln_pre.weight", VW, false, &m->ln_pre_g);
    rc |= need(m, "visual.ln_pre.bias", VW, false, &m->ln_pre_b);
    rc |= need(m, "visual.ln_post.weight", VW, false, &m->ln_post_g);
    rc |= need(m, "visual.ln_post.bias", VW, false, &m->ln_post_b);
    rc |= need(m, "visual.proj", 1ll * VW * ED, true, &m->vproj_w);
    rc |= need(m, "token_embedding.weight", 1ll * VOCAB * TW, false, &m->tok_emb);
    rc |= need(m, "positional_embedding", 1ll * TL * TW, false, &m->tpos);
    rc |= need(m, "ln_final.weight", TW, false, &m->lnf_g);
    rc |= need(m, "ln_final.bias", TW, false, &m->lnf_b);
    rc |= need(m, "text_projection", 1ll * TW * ED, true, &m->tproj_w);
    if (rc) return CB_ERR_INVALID;
    // class token row before ln_pre = class_embedding + positional_embedding[0]
    {
        std::vector<float> a(VW), b(VW);
        CB_CUDA(cudaMemcpy(a.data(), cls_emb, VW * 4, cudaMemcpyDeviceToHost));
        CB_CUDA(cudaMemcpy(b.data(), m->vpos, VW * 4, cudaMemcpyDeviceToHost));
        for (int i = 0; i < VW; i++) a[i] += b[i];
        CB_CUDA(cudaMemcpy(m->cls_pos, a.data(), VW * 4, cudaMemcpyHostToDevice));
    }
    // LayerNorm fold (ln_fold knob: 0 never, 1 always, unset = fold if a calibration batch agrees with the
    // unfolded model).  The folded GEMM consumes the raw fp16 residual stream and removes mean * colsum
    // afterwards, with variance = E[x^2] - mean^2 in fp32: exact enough for sane activations, but a
    // checkpoint whose residual stream carries a large common-mode offset would lose digits there.  The
    // reference normalises in fp32 BEFORE the fp16 cast, so when the two disagree the unfolded form wins.
    const int64_t knob = tune(T_LN_FOLD);
    m->ln_fold_cos = -2.0;
    std::vector<float> e_plain;
    if (knob != 1) {
        m->ln_fold = false;
        if ((rc = bind_blocks(m, "visual.transformer", VW, m->vis))) return rc;
        if ((rc = bind_blocks(m, "transformer", TW, m->txt))) return rc;
        if (knob < 0 && (rc = calibration_pass(m, e_plain))) return rc;
    }
    if (knob != 0) {
        if ((rc = fold_layernorm(m, "visual.transformer", VW))) return rc;
        if ((rc = fold_layernorm(m, "transformer", TW))) return rc;
        m->ln_fold = true;
        if ((rc = bind_blocks(m, "visual.transformer", VW, m->vis))) return rc;
        if ((rc = bind_blocks(m, "transformer", TW, m->txt))) return rc;
        if (knob < 0) {
            std::vector<float> e_fold;
            if ((rc = calibration_pass(m, e_fold))) return rc;
            double worst = 1.0;
            for (size_t r = 0; r + ED <= e_fold.size() && r + ED <= e_plain.size(); r += ED) {
                double ab = 0, aa = 0, bb = 0;
                for (int i = 0; i < ED; i++) {
                    ab += (double)e_fold[r + i] * e_plain[r + i];
                    aa += (double)e_fold[r + i] * e_fold[r + i];
                    bb += (double)e_plain[r + i] * e_plain[r + i];
                }
                const double c = (aa > 0 && bb > 0) ? ab / std::sqrt(aa * bb) : -1.0;
                worst = std::min(worst, std::isfinite(c) ? c : -1.0);
            }
            m->ln_fold_cos = worst;
            if (worst < kFoldMinCos) {
                // put the raw weights / biases back and run LayerNorm as its own launch
                if ((rc = unfold_layernorm(m, "visual.transformer", VW))) return rc;
                if ((rc = unfold_layernorm(m, "transformer", TW))) return rc;
                m->ln_fold = false;
                if ((rc = bind_blocks(m, "visual.transformer", VW, m->vis))) return rc;
                if ((rc = bind_blocks(m, "transformer", TW, m->txt))) return rc;
            }
        }
    }
    m->host.clear();
    m->finalized = true;
    return CB_OK;
}

#define CLIP_READY(m, fn)                                                              \
    CB_REQUIRE((m) != nullptr, fn ": null handle");                                     \
    CB_REQUIRE((m)->finalized, fn ": parameters not finalized (call cb_clip_finalize)")

int cb_clip_encode_image_u8_device(cb_clip *m, int64_t B, const uint8_t *hwc_dev, float *out_dev, int normalize,
                                   void *stream) {
    CLIP_READY(m, "cb_clip_encode_image_u8_device");
    CB_REQUIRE(B >= 0, "cb_clip_encode_image_u8_device: B < 0");
    if (B == 0) return CB_OK;
    CB_REQUIRE(hwc_dev && out_dev, "cb_clip_encode_image_u8_device: null buffer");
    CB_REQUIRE(m->max_img > 0, "cb_clip_encode_image_u8_device: handle created with max_image_batch = 0");
    DeviceGuard g(m->device);
    cudaStream_t s = (cudaStream_t)stream;
    for (int64_t lo = 0; lo < B; lo += m->max_img) {
        const int b = (int)std::min<int64_t>(m->max_img, B - lo);
        const size_t in_bytes = (size_t)b * 224 * 224 * 3;
        int rc = graphed_forward(m, 0, b, normalize, hwc_dev + (size_t)lo * 224 * 224 * 3, m->d_img, in_bytes,
                                 out_dev + lo * ED, s, [&](const void *in, float *out, cudaStream_t cs) {
                                     return encode_u8_chunk(m, m->ws, b, (const uint8_t *)in, out, normalize, cs);
                                 });
        if (rc) return rc;
    }
    return CB_OK;
}

int cb_clip_encode_image_f32_device(cb_clip *m, int64_t B, const float *nchw_dev, float *out_dev, int normalize,
                                    void *stream) {
    CLIP_READY(m, "cb_clip_encode_image_f32_device");
    CB_REQUIRE(B >= 0, "cb_clip_encode_image_f32_device: B < 0");
    if (B == 0) return CB_OK;
    CB_REQUIRE(nchw_dev && out_dev, "cb_clip_encode_image_f32_device: null buffer");
    CB_REQUIRE(m->max_img > 0, "cb_clip_encode_image_f32_device: handle created with max_image_batch = 0");
    DeviceGuard g(m->device);
    cudaStream_t s = (cudaStream_t)stream;
    for (int64_t lo = 0; lo < B; lo += m->max_img) {
        const int b = (int)std::min<int64_t>(m->max_img, B - lo);
        const size_t in_bytes = (size_t)b * 3 * 224 * 224 * 4;
        if (b <= kGraphMaxBatch && !m->d_img_f32)
            CB_CUDA(cudaMalloc(&m->d_img_f32, (size_t)std::min(kGraphMaxBatch, m->max_img) * 3 * 224 * 224 * 4));
        int rc = graphed_forward(m, 1, b, normalize, nchw_dev + (size_t)lo * 3 * 224 * 224, m->d_img_f32, in_bytes,
                                 out_dev + lo * ED, s, [&](const void *in, float *out, cudaStream_t cs) {
                                     int r = preprocess_f32((const float *)in, m->ws.patches, b, cs);
                                     return r ? r : vision_from_patches(m, m->ws, b, out, normalize, cs);
                                 });
        if (rc) return rc;
    }
    return CB_OK;
}

int cb_clip_encode_text_device(cb_clip *m, int64_t B, const int32_t *ids_dev, float *out_dev, int normalize,
                               void *stream) {
    CLIP_READY(m, "cb_clip_encode_text_device");
    CB_REQUIRE(B >= 0, "cb_clip_encode_text_device: B < 0");
    if (B == 0) return CB_OK;
    CB_REQUIRE(ids_dev && out_dev, "cb_clip_encode_text_device: null buffer");
    CB_REQUIRE(m->max_txt > 0, "cb_clip_encode_text_device: handle created with max_text_batch = 0");
    DeviceGuard g(m->device);
    cudaStream_t s = (cudaStream_t)stream;
    for (int64_t lo = 0; lo < B; lo += m->max_txt) {
        const int b = (int)std::min<int64_t>(m->max_txt, B - lo);
        int rc = graphed_forward(m, 2, b, normalize, ids_dev + lo * TL, m->d_ids, (size_t)b * TL * 4, out_dev + lo * ED, s,
                                 [&](const void *in, float *out, cudaStream_t cs) {
                                     return text_forward(m, m->ws, b, (const int32_t *)in, out, normalize, cs);
                                 });
        if (rc) return rc;
    }
    return CB_OK;
}

int cb_clip_encode_image_u8(cb_clip *m, int64_t B, const uint8_t *hwc_host, float *out_host, int normalize) {
    CLIP_READY(m, "cb_clip_encode_image_u8");
    CB_REQUIRE(B >= 0, "cb_clip_encode_image_u8: B < 0");
    if (B == 0) return CB_OK;
    CB_REQUIRE(hwc_host && out_host, "cb_clip_encode_image_u8: null buffer");
    CB_REQUIRE(m->max_img > 0, "cb_clip_encode_image_u8: handle created with max_image_batch = 0");
    DeviceGuard g(m->device);
    const size_t img_bytes = 224 * 224 * 3;
    for (int64_t lo = 0; lo < B; lo += m->max_img) {
        const int b = (int)std::min<int64_t>(m->max_img, B - lo);
        CB_CUDA(cudaMemcpyAsync(m->d_img, hwc_host + (size_t)lo * img_bytes, (size_t)b * img_bytes,
                                cudaMemcpyHostToDevice, m->stream));
        int rc = cb_clip_encode_image_u8_device(m, b, m->d_img, m->d_out, normalize, m->stream);
        if (rc) return rc;
        CB_CUDA(cudaMemcpyAsync(out_host + lo * ED, m->d_out, (size_t)b * ED * 4, cudaMemcpyDeviceToHost, m->stream));
        CB_CUDA(cudaStreamSynchronize(m->stream));
    }
    return CB_OK;
}

// Pipelined variants: return as soon as the work is queued.  Batches alternate between two
// lanes (own workspace + streams): the H2D copy of one batch overlaps the forward pass of the
// other, and the HBM-bound kernels of one forward pass (LayerNorm, attention) overlap the
// tensor-bound GEMMs of the other.  Results are complete after cb_clip_sync() (host
// variant) / cb_clip_join() (device variant).  Host buffers should be pinned and must stay
// valid until then.
static int lanes_init(cb_clip *m) {
    if (m->lanes_ready) return CB_OK;
    const size_t img_bytes = 224 * 224 * 3;
    for (Lane &l : m->lanes) {
        cudaError_t e = alloc_ws(m, l.ws);
        if (e != cudaSuccess) { set_error("lane workspace: %s", cudaGetErrorString(e)); return CB_ERR_OOM; }
        CB_CUDA(cudaStreamCreateWithFlags(&l.stream, cudaStreamNonBlocking));
        CB_CUDA(cudaStreamCreateWithFlags(&l.copy_stream, cudaStreamNonBlocking));
        for (int i = 0; i < 2; i++) {
            CB_CUDA(cudaMalloc(&l.img[i], (size_t)std::max(m->max_img, 1) * img_bytes));
            CB_CUDA(cudaEventCreateWithFlags(&l.copied[i], cudaEventDisableTiming));
            CB_CUDA(cudaEventCreateWithFlags(&l.consumed[i], cudaEventDisableTiming));
        }
        CB_CUDA(cudaMalloc(&l.out, (size_t)std::max(m->max_img, 1) * ED * 4));
        CB_CUDA(cudaEventCreateWithFlags(&l.done, cudaEventDisableTiming));
    }
    CB_CUDA(cudaEventCreateWithFlags(&m->fence_ev, cudaEventDisableTiming));
    m->lanes_ready = true;
    return CB_OK;
}

int cb_clip_submit_image_u8(cb_clip *m, int64_t B, const uint8_t *hwc_host, float *out_host, int normalize) {
    CLIP_READY(m, "cb_clip_submit_image_u8");
    CB_REQUIRE(B > 0 && B <= m->max_img, "cb_clip_submit_image_u8: B must be in [1, max_image_batch]");
    CB_REQUIRE(hwc_host && out_host, "cb_clip_submit_image_u8: null buffer");
    DeviceGuard g(m->device);
    int rc = lanes_init(m);
    if (rc) return rc;
    Lane &l = m->lanes[m->next_lane];
    m->next_lane ^= 1;
    const int sl = l.n_host & 1;
    l.n_host++;
    // the forward pass that last read this input slot (two submissions ago on this lane) must be
    // past its preprocess kernel before the slot is overwritten
    if (l.slot_used[sl]) CB_CUDA(cudaStreamWaitEvent(l.copy_stream, l.consumed[sl], 0));
    CB_CUDA(cudaMemcpyAsync(l.img[sl], hwc_host, (size_t)B * 224 * 224 * 3, cudaMemcpyHostToDevice, l.copy_stream));
    CB_CUDA(cudaEventRecord(l.copied[sl], l.copy_stream));
    CB_CUDA(cudaStreamWaitEvent(l.stream, l.copied[sl], 0));
    if ((rc = timed_other(m, 3, l.stream, [&] { return preprocess_u8(l.img[sl], l.ws.patches, (int)B, l.stream); }))) return rc;
    CB_CUDA(cudaEventRecord(l.consumed[sl], l.stream));      // pixels are in the im2col buffer now
    if ((rc = vision_from_patches(m, l.ws, (int)B, l.out, normalize, l.stream))) return rc;
    CB_CUDA(cudaMemcpyAsync(out_host, l.out, (size_t)B * ED * 4, cudaMemcpyDeviceToHost, l.stream));
    CB_CUDA(cudaEventRecord(l.done, l.stream));
    l.slot_used[sl] = true;
    l.used = true;
    return CB_OK;
}

// device-resident input and output; ordered after everything queued on `after_stream` so far
int cb_clip_submit_image_u8_device(cb_clip *m, int64_t B, const uint8_t *hwc_dev, float *out_dev, int normalize,
                                   void *after_stream) {
    CLIP_READY(m, "cb_clip_submit_image_u8_device");
    CB_REQUIRE(B > 0 && B <= m->max_img, "cb_clip_submit_image_u8_device: B must be in [1, max_image_batch]");
    CB_REQUIRE(hwc_dev && out_dev, "cb_clip_submit_image_u8_device: null buffer");
    DeviceGuard g(m->device);
    int rc = lanes_init(m);
    if (rc) return rc;
    Lane &l = m->lanes[m->next_lane];
    m->next_lane ^= 1;
    CB_CUDA(cudaEventRecord(m->fence_ev, (cudaStream_t)after_stream));
    CB_CUDA(cudaStreamWaitEvent(l.stream, m->fence_ev, 0));
    if ((rc = encode_u8_chunk(m, l.ws, (int)B, hwc_dev, out_dev, normalize, l.stream))) return rc;
    CB_CUDA(cudaEventRecord(l.done, l.stream));
    l.used = true;
    return CB_OK;
}

// make `stream` wait for everything submitted so far (no host synchronisation)
int cb_clip_join(cb_clip *m, void *stream) {
    CB_REQUIRE(m != nullptr, "cb_clip_join: null handle");
    DeviceGuard g(m->device);
    for (Lane &l : m->lanes)
        if (l.used) CB_CUDA(cudaStreamWaitEvent((cudaStream_t)stream, l.done, 0));
    return CB_OK;
}

int cb_clip_sync(cb_clip *m) {
    CB_REQUIRE(m != nullptr, "cb_clip_sync: null handle");
    DeviceGuard g(m->device);
    for (Lane &l : m->lanes) {
        if (l.copy_stream) CB_CUDA(cudaStreamSynchronize(l.copy_stream));
        if (l.stream) CB_CUDA(cudaStreamSynchronize(l.stream));
    }
    CB_CUDA(cudaStreamSynchronize(m->stream));
    return CB_OK;
}

int cb_clip_encode_text(cb_clip *m, int64_t B, const int32_t *ids_host, float *out_host, int normalize) {
    CLIP_READY(m, "cb_clip_encode_text");
    CB_REQUIRE(B >= 0, "cb_clip_encode_text: B < 0");
    if (B == 0) return CB_OK;
    CB_REQUIRE(ids_host && out_host, "cb_clip_encode_text: null buffer");
    CB_REQUIRE(m->max_txt > 0, "cb_clip_encode_text: handle created with max_text_batch = 0");
    DeviceGuard g(m->device);
    for (int64_t lo = 0; lo < B; lo += m->max_txt) {
        const int b = (int)std::min<int64_t>(m->max_txt, B - lo);
        CB_CUDA(cudaMemcpyAsync(m->d_ids, ids_host + lo * TL, (size_t)b * TL * 4, cudaMemcpyHostToDevice, m->stream));
        int rc = cb_clip_encode_text_device(m, b, m->d_ids, m->d_out, normalize, m->stream);
        if (rc) return rc;
        CB_CUDA(cudaMemcpyAsync(out_host + lo * ED, m->d_out, (size_t)b * ED * 4, cudaMemcpyDeviceToHost, m->stream));
        CB_CUDA(cudaStreamSynchronize(m->stream));
    }
    return CB_OK;
}

int cb_clip_ln_fold_status(cb_clip *m, int *folded, double *min_cosine) {
    CB_REQUIRE(m && folded && min_cosine, "cb_clip_ln_fold_status: null argument");
    *folded = m->ln_fold ? 1 : 0;
    *min_cosine = m->ln_fold_cos;
    return CB_OK;
}

int cb_clip_timing(cb_clip *m, int enable) {
    CB_REQUIRE(m != nullptr, "cb_clip_timing: null handle");
    DeviceGuard g(m->device);
    m->timing = enable;
    m->ev_n = 0;
    m->gemm_flops = 0;
    m->stamp_n = 0;
    if (m->timing) {
        if (!m->stamps) CB_CUDA(cudaMalloc(&m->stamps, (size_t)cb_clip::kStamps * 16));
        // (min, max) pairs start at (all ones, 0); wait for launches that may still write the old ones
        CB_CUDA(cudaDeviceSynchronize());
        std::vector<unsigned long long> init((size_t)cb_clip::kStamps * 2);
        for (size_t i = 0; i < init.size(); i += 2) { init[i] = ~0ull; init[i + 1] = 0ull; }
        CB_CUDA(cudaMemcpy(m->stamps, init.data(), init.size() * 8, cudaMemcpyHostToDevice));
    }
    if (m->timing && m->ev.empty()) {
        m->ev.resize(8192);
        m->ev_cat.assign(4096, 0);
        for (auto &e : m->ev) CB_CUDA(cudaEventCreate(&e));
    }
    return CB_OK;
}

int cb_clip_timing_breakdown(cb_clip *m, double *ms_by_class4) {
    CB_REQUIRE(m && ms_by_class4, "cb_clip_timing_breakdown: null argument");
    DeviceGuard g(m->device);
    for (int i = 0; i < 4; i++) ms_by_class4[i] = 0;
    CB_CUDA(cudaDeviceSynchronize());
    if (m->stamp_n > 0) {
        std::vector<unsigned long long> st((size_t)m->stamp_n * 2);
        CB_CUDA(cudaMemcpy(st.data(), m->stamps, st.size() * 8, cudaMemcpyDeviceToHost));
        for (int i = 0; i < m->stamp_n; i++)
            if (st[2 * i + 1] >= st[2 * i]) ms_by_class4[0] += (double)(st[2 * i + 1] - st[2 * i]) * 1e-6;
    }
    for (int i = 0; i + 1 < m->ev_n; i += 2) {
        CB_CUDA(cudaEventSynchronize(m->ev[i + 1]));
        float ms = 0;
        CB_CUDA(cudaEventElapsedTime(&ms, m->ev[i], m->ev[i + 1]));
        ms_by_class4[m->ev_cat[i / 2] & 3] += ms;
    }
    return CB_OK;
}

int cb_clip_timing_launches(cb_clip *m, double *ms_out, double *t0_ms_out, int cap, int *n) {
    CB_REQUIRE(m && ms_out && t0_ms_out && n, "cb_clip_timing_launches: null argument");
    DeviceGuard g(m->device);
    CB_CUDA(cudaDeviceSynchronize());
    *n = std::min(cap, m->stamp_n);
    if (*n <= 0) { *n = 0; return CB_OK; }
    std::vector<unsigned long long> st((size_t)*n * 2);
    CB_CUDA(cudaMemcpy(st.data(), m->stamps, st.size() * 8, cudaMemcpyDeviceToHost));
    for (int i = 0; i < *n; i++) {
        ms_out[i] = st[2 * i + 1] >= st[2 * i] ? (double)(st[2 * i + 1] - st[2 * i]) * 1e-6 : -1.0;
        t0_ms_out[i] = (double)(st[2 * i] - st[0]) * 1e-6;
    }
    return CB_OK;
}

int cb_clip_timing_read(cb_clip *m, double *gemm_ms_total, double *gemm_flops, int *n_gemms) {
    CB_REQUIRE(m && gemm_ms_total && gemm_flops && n_gemms, "cb_clip_timing_read: null argument");
    DeviceGuard g(m->device);
    CB_CUDA(cudaDeviceSynchronize());
    double tot = 0;
    int ng = 0;
    if (m->stamp_n > 0) {
        std::vector<unsigned long long> st((size_t)m->stamp_n * 2);
        CB_CUDA(cudaMemcpy(st.data(), m->stamps, st.size() * 8, cudaMemcpyDeviceToHost));
        for (int i = 0; i < m->stamp_n; i++)
            if (st[2 * i + 1] >= st[2 * i]) { tot += (double)(st[2 * i + 1] - st[2 * i]) * 1e-6; ng++; }
    }
    *gemm_ms_total = tot;
    *gemm_flops = m->gemm_flops;
    *n_gemms = ng;
    m->ev_n = 0;
    m->gemm_flops = 0;
    if (m->timing) return cb_clip_timing(m, m->timing);      // re-arm the stamp slots
    m->stamp_n = 0;
    return CB_OK;
}

// ---- building blocks exported for unit tests ------------------------------------------
int cb_layernorm_f16_device(const void *in, void *out, const float *gamma, const float *beta, int rows, int width,
                            int in_row_stride, const int *gather, const float *cls_fill, int cls_period,
                            void *stream) {
    CB_REQUIRE(in && out && gamma && beta, "cb_layernorm_f16_device: null buffer");
    return layernorm_f16((const __half *)in, (__half *)out, gamma, beta, rows, width, in_row_stride, gather, cls_fill,
                         cls_period, (cudaStream_t)stream);
}
int cb_attention_f16_device(const void *qkv, void *out, int B, int L, int heads, int causal, void *stream) {
    CB_REQUIRE(qkv && out, "cb_attention_f16_device: null buffer");
    return attention_f16((const __half *)qkv, (__half *)out, B, L, heads, causal != 0, (cudaStream_t)stream);
}
int cb_preprocess_u8_device(const uint8_t *hwc, void *patches_f16, int B, void *stream) {
    CB_REQUIRE(hwc && patches_f16, "cb_preprocess_u8_device: null buffer");
    return preprocess_u8(hwc, (__half *)patches_f16, B, (cudaStream_t)stream);
}
int cb_preprocess_f32_device(const float *nchw, void *patches_f16, int B, void *stream) {
    CB_REQUIRE(nchw && patches_f16, "cb_preprocess_f32_device: null buffer");
    return preprocess_f32(nchw, (__half *)patches_f16, B, (cudaStream_t)stream);
}
int cb_l2norm_f32_device(const float *in, float *out, int rows, int width, void *stream) {
    CB_REQUIRE(in && out, "cb_l2norm_f32_device: null buffer");
    return l2norm_rows_f32(in, out, rows, width, (cudaStream_t)stream);
}

}  // extern "C"
