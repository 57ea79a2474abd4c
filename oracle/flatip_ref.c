/* ORACLE (test infrastructure, not product code).
 *
 * Plain-C restatement of the algorithm faiss uses for
 *   IndexFlatIP::search -> knn_inner_product   (small nq: one dot product per
 *   (query, row), results kept in a k-element min-heap, then heap-sorted
 *   descending; unfilled slots id=-1 / score=-FLT_MAX)
 * which is what the reference calls at /root/reference/query-index.py:111 on
 * the index built at /root/reference/build-index.py:80-81.  faiss itself is an
 * un-vendored, un-pinned dependency (setup.sh:12), so this follows its
 * published behaviour, not its source.
 *
 * It exists to pin oracle/flatip_ref.py with an independently written second
 * implementation, and to serve as the multi-threaded "port" CPU baseline in
 * bench.py (OpenMP over database slices, like faiss's OpenMP loops).
 *
 * Tie rule: a candidate displaces the heap root only if strictly greater, or
 * equal with a smaller id -- i.e. the total order (-score, id).
 *
 * Build: see oracle/Makefile  (gcc -O3 -march=native -fopenmp -shared -fPIC)
 */
#include <float.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

typedef struct { float s; int64_t id; } ent_t;

/* a "worse" than b  <=>  a sorts after b in (-score, id) order */
static inline int worse(ent_t a, ent_t b) {
    return a.s < b.s || (a.s == b.s && a.id > b.id);
}

static void sift_down(ent_t *h, int n, int i) {
    for (;;) {
        int l = 2 * i + 1, r = l + 1, m = i;
        if (l < n && worse(h[l], h[m])) m = l;
        if (r < n && worse(h[r], h[m])) m = r;
        if (m == i) return;
        ent_t t = h[i]; h[i] = h[m]; h[m] = t; i = m;
    }
}

static void heap_offer(ent_t *h, int *n, int k, ent_t e) {
    if (*n < k) {                       /* fill phase: sift up */
        int i = (*n)++;
        h[i] = e;
        while (i > 0) {
            int p = (i - 1) / 2;
            if (!worse(h[i], h[p])) break;
            ent_t t = h[i]; h[i] = h[p]; h[p] = t; i = p;
        }
    } else if (worse(h[0], e)) {        /* root is the current worst */
        h[0] = e;
        sift_down(h, k, 0);
    }
}

static int cmp_desc(const void *pa, const void *pb) {
    ent_t a = *(const ent_t *)pa, b = *(const ent_t *)pb;
    if (worse(a, b)) return 1;
    if (worse(b, a)) return -1;
    return 0;
}

/* fp16 rows are widened with the compiler's _Float16 (IEEE binary16; F16C when
 * -mf16c is given), exactly numpy's astype(float32). */
static inline float half_to_float(uint16_t h) {
    _Float16 v;
    memcpy(&v, &h, 2);
    return (float)v;
}

static inline float dot_f32(const float *q, const float *x, int d) {
    float acc[8] = {0};
    for (int j = 0; j + 8 <= d; j += 8)
        for (int l = 0; l < 8; l++) acc[l] += q[j + l] * x[j + l];
    float s = 0.f;
    for (int l = 0; l < 8; l++) s += acc[l];
    for (int j = d & ~7; j < d; j++) s += q[j] * x[j];
    return s;
}
static inline float dot_f16(const float *q, const uint16_t *x, int d) {
    /* 8 independent partial sums so the loop vectorises (faiss's SIMD kernels sum in
     * lanes too); the order differs from numpy's BLAS by O(1e-7), inside tolerance */
    float acc[8] = {0};
    const _Float16 *h = (const _Float16 *)x;
    for (int j = 0; j + 8 <= d; j += 8)
        for (int l = 0; l < 8; l++) acc[l] += q[j + l] * (float)h[j + l];
    float s = 0.f;
    for (int l = 0; l < 8; l++) s += acc[l];
    for (int j = d & ~7; j < d; j++) s += q[j] * (float)h[j];
    return s;
}

/* dtype: 0 = fp32 rows, 1 = fp16 rows.  Returns 0 on success. */
int oracle_flatip_search(const void *xb, int dtype, int64_t n, int d,
                         const float *xq, int64_t nq, int64_t k,
                         float *D, int64_t *I, int nthreads) {
    if (k <= 0 || d <= 0) return 1;
#ifdef _OPENMP
    if (nthreads > 0) omp_set_num_threads(nthreads);
    int T = omp_get_max_threads();
#else
    int T = 1;
#endif
    (void)nthreads;
    for (int64_t qi = 0; qi < nq; qi++) {
        const float *q = xq + qi * d;
        ent_t *heaps = (ent_t *)malloc(sizeof(ent_t) * (size_t)k * T);
        int *cnt = (int *)calloc(T, sizeof(int));
        if (!heaps || !cnt) return 2;
#pragma omp parallel num_threads(T)
        {
#ifdef _OPENMP
            int t = omp_get_thread_num();
#else
            int t = 0;
#endif
            ent_t *h = heaps + (size_t)t * k;
            int hn = 0;
#pragma omp for schedule(static)
            for (int64_t i = 0; i < n; i++) {
                ent_t e;
                e.s = dtype ? dot_f16(q, (const uint16_t *)xb + i * d, d)
                            : dot_f32(q, (const float *)xb + i * d, d);
                e.s += 0.0f;
                e.id = i;
                heap_offer(h, &hn, (int)k, e);
            }
            cnt[t] = hn;
        }
        /* merge the per-thread heaps */
        int64_t tot = 0;
        for (int t = 0; t < T; t++) tot += cnt[t];
        ent_t *all = (ent_t *)malloc(sizeof(ent_t) * (size_t)(tot ? tot : 1));
        int64_t p = 0;
        for (int t = 0; t < T; t++)
            for (int j = 0; j < cnt[t]; j++) all[p++] = heaps[(size_t)t * k + j];
        qsort(all, (size_t)tot, sizeof(ent_t), cmp_desc);
        for (int64_t j = 0; j < k; j++) {
            if (j < tot) { D[qi * k + j] = all[j].s; I[qi * k + j] = all[j].id; }
            else         { D[qi * k + j] = -FLT_MAX; I[qi * k + j] = -1; }
        }
        free(all); free(heaps); free(cnt);
    }
    return 0;
}
