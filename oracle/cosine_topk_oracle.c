/*
 * cosine_topk_oracle.c -- CPU ORACLE (TEST INFRASTRUCTURE, NOT PRODUCT CODE)
 *
 * A plain-C restatement of the one hot path of Iamdarika/Spotify_recommender:
 * cosine similarity of a query song against every song's 12-float feature
 * row, followed by top-N selection.  Only tests/, __graft_entry__.smoke() and
 * bench.py's cpu_baseline / --impl reference legs may load this library; the
 * product path (spotify_recommender_b200/csrc) never links or calls it.
 *
 * Parity pin: every function below is checked in tests/test_oracle.py against
 *   (a) the known-answer vectors of SURVEY.md Appendix A (tests/golden/), which
 *       were captured from the reference's own CPU build, and
 *   (b) oracle/_ref/libref_cpu.so -- the UNMODIFIED reference Recommender.cu
 *       compiled with -DDISABLE_CUDA from /root/reference (see oracle/Makefile)
 *       whenever that file is present.
 *
 * Arithmetic rules (reference Recommender.cu:256-273, "calculateSimilaritiesCPU"):
 *   - accumulators start at 0.0f, features visited j = 0..11 ascending;
 *   - every product is rounded to FP32 and then added (NO fused multiply-add:
 *     this file must be compiled with -ffp-contract=off, the Makefile does);
 *   - queryNorm = sqrtf(sum q_j*q_j) by the same recipe      (:259-261)
 *   - den = sqrtf(norm) * queryNorm                          (:270)
 *   - score = den > 1e-8f ? max(-1, min(1, dot / den)) : 0   (:271)
 *     with std::min / std::max semantics (second argument wins only when
 *     strictly smaller / larger, so NaN collapses the way libstdc++ does).
 * Selection rules (reference Recommender.cu:275-318, "recommendByIndex"):
 *   - the query song itself is skipped by index only          (:296)
 *   - at most min(K, N-1) results, best first                 (:300-315)
 *   - CANONICAL order used by the build: score descending, then LOWER song
 *     index first (north_star).  The reference's own order on exact ties is a
 *     libstdc++ heap artefact; sr_oracle_topk_refheap() restates that artefact
 *     (std::push_heap / std::pop_heap of bits/stl_heap.h) so the oracle can be
 *     pinned on the reference's tie cases as well.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define SR_FEATURES 12 /* reference Song.h:12 FEATURE_COUNT */

/* Returns 1 when this object really was built without FMA contraction: with
 * a = 1 + 2^-12 the rounded square drops the 2^-24 term, a fused a*a - r keeps
 * it.  tests/test_oracle.py asserts this. */
int sr_oracle_selfcheck_unfused(void)
{
    volatile float a = 1.0f + 1.0f / 4096.0f;
    float r = a * a;
    volatile float nr = -r;
    float p = a * a;
    float d = p + nr; /* 0 when unfused; 2^-24 if contracted into an FMA */
    return d == 0.0f;
}

/* reference Recommender.cu:259-261 -- norm of the query row, unfused */
float sr_oracle_query_norm(const float *q)
{
    float acc = 0.0f;
    for (int j = 0; j < SR_FEATURES; ++j) {
        float p = q[j] * q[j];
        acc = acc + p;
    }
    return sqrtf(acc);
}

/* reference Recommender.cu:263-271 -- one (query, song) score */
float sr_oracle_pair_score(const float *q, float query_norm, const float *f)
{
    float dot = 0.0f;
    float norm = 0.0f;
    for (int j = 0; j < SR_FEATURES; ++j) {
        float pd = q[j] * f[j];
        dot = dot + pd;
        float pn = f[j] * f[j];
        norm = norm + pn;
    }
    norm = sqrtf(norm) * query_norm;
    if (norm > 1e-8f) {
        float s = dot / norm;
        float lo = (s < 1.0f) ? s : 1.0f;      /* std::min(1.0f, s)  */
        float hi = (-1.0f < lo) ? lo : -1.0f;  /* std::max(-1.0f, lo) */
        return hi;
    }
    return 0.0f;
}

/* reference Recommender.cu:256-273 -- all N scores of one query row.
 * features: dense row-major N x 12.  threads <= 1 => strictly serial. */
void sr_oracle_scores(const float *features, int64_t n, const float *q,
                      float *out, int threads)
{
    const float qn = sr_oracle_query_norm(q);
    (void)threads;
#ifdef _OPENMP
#pragma omp parallel for schedule(static) num_threads(threads > 1 ? threads : 1)
#endif
    for (int64_t i = 0; i < n; ++i)
        out[i] = sr_oracle_pair_score(q, qn, features + i * SR_FEATURES);
}

/* ---- canonical selection: score desc, index asc ------------------------- */

static inline int sr_better(float sa, int32_t ia, float sb, int32_t ib)
{
    /* is (sa, ia) strictly ahead of (sb, ib) in canonical order? */
    return (sa > sb) || (sa == sb && ia < ib);
}

/* bounded insertion into a sorted (best first) list of capacity k */
static inline int sr_insert(float *sc, int32_t *ix, int len, int k, float s,
                            int32_t i)
{
    if (len == k && !sr_better(s, i, sc[k - 1], ix[k - 1]))
        return len;
    int pos = (len < k) ? len : k - 1;
    while (pos > 0 && sr_better(s, i, sc[pos - 1], ix[pos - 1])) {
        sc[pos] = sc[pos - 1];
        ix[pos] = ix[pos - 1];
        --pos;
    }
    sc[pos] = s;
    ix[pos] = i;
    return (len < k) ? len + 1 : k;
}

/* Canonical top-K over precomputed scores.  `exclude` (< 0: none) is skipped by
 * index (reference :296).  id_base is added to every emitted index (row-shard
 * base).  Returns the number of results written (<= k); the tail of out_idx is
 * filled with -1 and of out_score with 0. */
int sr_oracle_topk_canonical(const float *scores, int64_t n, int64_t exclude,
                             int k, int32_t id_base, int32_t *out_idx,
                             float *out_score)
{
    int len = 0;
    if (k <= 0)
        return 0;
    for (int64_t i = 0; i < n; ++i) {
        if (i == exclude)
            continue;
        len = sr_insert(out_score, out_idx, len, k, scores[i],
                        (int32_t)(i + id_base));
    }
    for (int r = len; r < k; ++r) {
        out_idx[r] = -1;
        out_score[r] = 0.0f;
    }
    return len;
}

/* ---- the reference's own selection, heap artefact included -------------- *
 * std::priority_queue<Recommendation> with operator< == "similarity >"
 * (reference Recommender.h:19-21) is a MIN-heap on similarity.  push = vector
 * push_back + std::push_heap (sift the new leaf up while parent "<" value);
 * pop = std::pop_heap (move root out, walk the hole to a leaf always taking the
 * child that is not "<" its sibling, drop the former last element in the hole,
 * sift it up) + pop_back.  This mirrors libstdc++ bits/stl_heap.h
 * (__push_heap / __adjust_heap), GCC 13. */
typedef struct {
    int32_t idx;
    float sim;
} sr_rec;

static inline int sr_rec_less(const sr_rec *a, const sr_rec *b)
{
    return a->sim > b->sim; /* reference Recommender.h:19-21 */
}

static void sr_push_heap(sr_rec *h, int64_t hole, int64_t top, sr_rec v)
{
    int64_t parent = (hole - 1) / 2;
    while (hole > top && sr_rec_less(&h[parent], &v)) {
        h[hole] = h[parent];
        hole = parent;
        parent = (hole - 1) / 2;
    }
    h[hole] = v;
}

static void sr_adjust_heap(sr_rec *h, int64_t hole, int64_t len, sr_rec v)
{
    const int64_t top = hole;
    int64_t child = hole;
    while (child < (len - 1) / 2) {
        child = 2 * (child + 1);
        if (sr_rec_less(&h[child], &h[child - 1]))
            --child;
        h[hole] = h[child];
        hole = child;
    }
    if ((len & 1) == 0 && child == (len - 2) / 2) {
        child = 2 * (child + 1);
        h[hole] = h[child - 1];
        hole = child - 1;
    }
    sr_push_heap(h, hole, top, v);
}

static void sr_heap_pop(sr_rec *h, int64_t *len)
{
    if (*len > 1) {
        sr_rec last = h[*len - 1];
        h[*len - 1] = h[0];
        sr_adjust_heap(h, 0, *len - 1, last);
    }
    --*len;
}

/* reference Recommender.cu:293-315.  Returns count written; k <= 0 returns 0
 * (the reference would dereference an empty heap: SURVEY Appendix B.2). */
int sr_oracle_topk_refheap(const float *scores, int64_t n, int64_t exclude,
                           int k, int32_t *out_idx)
{
    if (k <= 0 || n <= 0)
        return 0;
    int64_t cap = (k < n) ? k : n;
    sr_rec *h = (sr_rec *)malloc((size_t)(cap + 1) * sizeof(sr_rec));
    int64_t len = 0;
    for (int64_t i = 0; i < n; ++i) {
        if (i == exclude)
            continue;
        sr_rec r = {(int32_t)i, scores[i]};
        if (len < k) {
            h[len] = r;
            ++len;
            sr_push_heap(h, len - 1, 0, r);
        } else if (r.sim > h[0].sim) {
            sr_heap_pop(h, &len);
            h[len] = r;
            ++len;
            sr_push_heap(h, len - 1, 0, r);
        }
    }
    int cnt = (int)len;
    /* drain (worst first) then reverse => best first */
    for (int w = cnt - 1; w >= 0; --w) {
        out_idx[w] = h[0].idx;
        sr_heap_pop(h, &len);
    }
    free(h);
    return cnt;
}

/* ---- batch drivers ------------------------------------------------------ */

/* Canonical top-K for a batch of query ROWS (q: nq x 12).  exclude[qi] is the
 * LOCAL row to skip or < 0.  out_idx / out_score are nq x k (-1 / 0 padded).
 * threads > 1 parallelises over songs inside a query when nq is small and over
 * queries otherwise; results do not depend on the thread count. */
void sr_oracle_query_rows(const float *features, int64_t n, const float *q,
                          const int64_t *exclude, int nq, int k,
                          int32_t id_base, int32_t *out_idx, float *out_score,
                          int threads)
{
    if (threads < 1)
        threads = 1;
    if (nq >= threads && threads > 1) {
#ifdef _OPENMP
#pragma omp parallel num_threads(threads)
#endif
        {
            float *sc = (float *)malloc((size_t)(n > 0 ? n : 1) * sizeof(float));
#ifdef _OPENMP
#pragma omp for schedule(dynamic, 1)
#endif
            for (int qi = 0; qi < nq; ++qi) {
                sr_oracle_scores(features, n, q + (size_t)qi * SR_FEATURES, sc, 1);
                sr_oracle_topk_canonical(sc, n, exclude ? exclude[qi] : -1, k,
                                         id_base, out_idx + (size_t)qi * k,
                                         out_score + (size_t)qi * k);
            }
            free(sc);
        }
        return;
    }
    float *sc = (float *)malloc((size_t)(n > 0 ? n : 1) * sizeof(float));
    for (int qi = 0; qi < nq; ++qi) {
        sr_oracle_scores(features, n, q + (size_t)qi * SR_FEATURES, sc, threads);
        sr_oracle_topk_canonical(sc, n, exclude ? exclude[qi] : -1, k, id_base,
                                 out_idx + (size_t)qi * k,
                                 out_score + (size_t)qi * k);
    }
    free(sc);
}

/* Batch of in-database queries by song index: the row is read from the store
 * and the song itself is excluded (reference recommendByIndex, :275-318). */
void sr_oracle_query_index(const float *features, int64_t n,
                           const int32_t *qidx, int nq, int k,
                           int32_t *out_idx, float *out_score, int threads)
{
    float *rows = (float *)malloc((size_t)(nq > 0 ? nq : 1) * SR_FEATURES * sizeof(float));
    int64_t *ex = (int64_t *)malloc((size_t)(nq > 0 ? nq : 1) * sizeof(int64_t));
    for (int qi = 0; qi < nq; ++qi) {
        memcpy(rows + (size_t)qi * SR_FEATURES,
               features + (size_t)qidx[qi] * SR_FEATURES,
               SR_FEATURES * sizeof(float));
        ex[qi] = qidx[qi];
    }
    sr_oracle_query_rows(features, n, rows, ex, nq, k, 0, out_idx, out_score,
                         threads);
    free(rows);
    free(ex);
}

/* Merge `parts` sorted candidate lists per query (each k long, -1 padded, as a
 * row-sharded run produces) into one canonical list: the CPU statement of the
 * multi-GPU merge step.  in_idx/in_score: parts x nq x k. */
void sr_oracle_merge_parts(const int32_t *in_idx, const float *in_score,
                           int parts, int nq, int k, int32_t *out_idx,
                           float *out_score)
{
    for (int qi = 0; qi < nq; ++qi) {
        int32_t *oi = out_idx + (size_t)qi * k;
        float *os = out_score + (size_t)qi * k;
        int len = 0;
        for (int p = 0; p < parts; ++p) {
            const int32_t *pi = in_idx + ((size_t)p * nq + qi) * k;
            const float *ps = in_score + ((size_t)p * nq + qi) * k;
            for (int r = 0; r < k; ++r) {
                if (pi[r] < 0)
                    continue;
                len = sr_insert(os, oi, len, k, ps[r], pi[r]);
            }
        }
        for (int r = len; r < k; ++r) {
            oi[r] = -1;
            os[r] = 0.0f;
        }
    }
}

/* ---- the min-max normalisation step of preprocessing ----------------------
 * Restates DataManager.cpp:270-301.  raw: n x 11 (danceability .. tempo, the order of
 * DataManager.cpp:156-159), genre_id: n, out: n x 12.
 *   - min/max per column with std::min / std::max semantics: the running value is kept
 *     unless the new one is strictly smaller / larger, so NaN entries never enter (:273-281);
 *     starting values numeric_limits<float>::max() / lowest() (:270-271);
 *   - range = max - min; range > 0.0001f ? (x - min) / range : 0.5f            (:291-296)
 *   - out[11] = (float)genre_id / max(1, n_genres - 1)                          (:299)
 * One deliberate normalisation of an order artefact: a column minimum / maximum of zero is
 * taken as +0.0f whatever the signs and order of the zeros in the column (the reference keeps
 * whichever zero it met first).  minmax_out (22 floats: 11 minima, 11 maxima) may be NULL. */
void sr_oracle_minmax_normalize(const float *raw, const int32_t *genre_id, int64_t n, int32_t n_genres,
                                float *out, float *minmax_out)
{
    float mn[SR_FEATURES - 1], mx[SR_FEATURES - 1];
    for (int j = 0; j < SR_FEATURES - 1; ++j) {
        mn[j] = 3.402823466e+38f;
        mx[j] = -3.402823466e+38f;
    }
    for (int64_t i = 0; i < n; ++i)
        for (int j = 0; j < SR_FEATURES - 1; ++j) {
            const float v = raw[i * (SR_FEATURES - 1) + j];
            if (v < mn[j]) mn[j] = v;
            if (v > mx[j]) mx[j] = v;
        }
    for (int j = 0; j < SR_FEATURES - 1; ++j) {
        if (mn[j] == 0.0f) mn[j] = 0.0f;
        if (mx[j] == 0.0f) mx[j] = 0.0f;
    }
    const int gden_i = n_genres - 1 > 1 ? n_genres - 1 : 1;
    const float gden = (float)gden_i;
    for (int64_t i = 0; i < n; ++i) {
        for (int j = 0; j < SR_FEATURES - 1; ++j) {
            const float range = mx[j] - mn[j];
            out[i * SR_FEATURES + j] = range > 0.0001f ? (raw[i * (SR_FEATURES - 1) + j] - mn[j]) / range : 0.5f;
        }
        out[i * SR_FEATURES + SR_FEATURES - 1] = (float)genre_id[i] / gden;
    }
    if (minmax_out)
        for (int j = 0; j < SR_FEATURES - 1; ++j) {
            minmax_out[j] = mn[j];
            minmax_out[SR_FEATURES - 1 + j] = mx[j];
        }
}

int sr_oracle_max_threads(void)
{
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}
