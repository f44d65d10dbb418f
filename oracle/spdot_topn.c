/* oracle/spdot_topn.c -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * CPU restatement of the sparse cosine top-n step that the reference delegates to the
 * third-party package `sparse_dot_topn` (pin: sparse_dot_topn>=0.2.9, reference setup.py:28;
 * API awesome_cossim_topn exists only in releases < 1.0).  Call site being restated:
 *     polyfuzz/models/_utils.py:82   awesome_cossim_topn(from_vector, to_vector.T, top_n+1, min_similarity)
 * followed by the reference's own post-processing
 *     polyfuzz/models/_utils.py:84-87   (self-match: zero the diagonal)
 *     polyfuzz/models/_utils.py:128-136 (_top_n_idx_sparse)
 *     polyfuzz/models/_utils.py:139-146 (_top_n_similarities_sparse)
 *
 * The package source is not available in this environment (no network), so this follows its
 * PUBLISHED algorithm (Gustavson row-wise SpGEMM with a dense accumulator + touched list, keep
 * values strictly greater than lower_bound, per-row top-n selection):  "parity unpinned" at the
 * package boundary -- the reference's tests hold no numeric vectors for it.  What IS pinned
 * (tests/test_oracle_golden.py): agreement with the unmodified reference `sklearn` branch
 * (polyfuzz/models/_utils.py:94-102) run in the build container, via tests/golden/ fixtures.
 *
 * Canonical parity contract because the reference leaves tie order undefined:
 *   - accumulate in fp64, terms of the from-row visited in ascending column order, product
 *     rounded before the add (no FMA contraction: build with -ffp-contract=off),
 *   - a candidate must have score > lower_bound (strict) and, in self-match mode, j != i,
 *   - ranking key: (score descending, to-index ascending),
 *   - unused slots: idx = -1, score = 0.
 *
 * B is passed as the inverted index of the to-matrix (CSR of to_vector.T, i.e. term-major
 * posting lists with ascending doc ids) exactly as awesome_cossim_topn receives it.
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

typedef struct { double v; int32_t j; } cand_t;

static int cand_before(const cand_t *a, const cand_t *b) {
    if (a->v != b->v) return a->v > b->v;
    return a->j < b->j;
}

/* insert into a sorted (best first) list of at most k entries */
static void topk_insert(cand_t *top, int *cnt, int k, cand_t c) {
    if (*cnt == k && !cand_before(&c, &top[k - 1])) return;
    int pos = (*cnt < k) ? *cnt : k - 1;
    while (pos > 0 && cand_before(&c, &top[pos - 1])) { top[pos] = top[pos - 1]; --pos; }
    top[pos] = c;
    if (*cnt < k) ++*cnt;
}

/* returns 0 on success.  n_threads <= 1 : single thread (how the reference calls the package,
 * _utils.py:82 passes no use_threads); > 1 : OpenMP over from-rows (the package's n_jobs variant). */
int oracle_spdot_topn(
    int32_t n_from, int32_t n_to,
    const int32_t *a_indptr, const int32_t *a_indices, const double *a_data,     /* from CSR  */
    const int32_t *bt_indptr, const int32_t *bt_indices, const double *bt_data,  /* to^T CSR  */
    int32_t k, double lower_bound,
    int32_t self_match,        /* 1: exclude j == from_index_base + i - to_index_base           */
    int64_t from_index_base,   /* global index of from-row 0                                    */
    int64_t to_index_base,     /* global index of to-row 0 (shards); output idx are GLOBAL       */
    int32_t *top_idx, double *top_val, int32_t n_threads)
{
    if (k <= 0) return 0;
    int nt = n_threads > 1 ? n_threads : 1;
    int err = 0;
#ifdef _OPENMP
#pragma omp parallel num_threads(nt)
#endif
    {
        double *sums = (double *)calloc((size_t)(n_to > 0 ? n_to : 1), sizeof(double));
        int32_t *touched = (int32_t *)malloc(sizeof(int32_t) * (size_t)(n_to > 0 ? n_to : 1));
        uint8_t *mark = (uint8_t *)calloc((size_t)(n_to > 0 ? n_to : 1), 1);
        cand_t *top = (cand_t *)malloc(sizeof(cand_t) * (size_t)k);
        if (!sums || !touched || !mark || !top) { err = 1; }
        else {
#ifdef _OPENMP
#pragma omp for schedule(dynamic, 64)
#endif
        for (int32_t i = 0; i < n_from; ++i) {
            int32_t nt_touched = 0;
            for (int32_t p = a_indptr[i]; p < a_indptr[i + 1]; ++p) {       /* ascending term id */
                int32_t t = a_indices[p];
                double v = a_data[p];
                for (int32_t q = bt_indptr[t]; q < bt_indptr[t + 1]; ++q) {
                    int32_t j = bt_indices[q];
                    double prod = v * bt_data[q];
                    if (!mark[j]) { mark[j] = 1; touched[nt_touched++] = j; }
                    sums[j] = sums[j] + prod;
                }
            }
            int cnt = 0;
            int64_t self_j = from_index_base + i - to_index_base;
            for (int32_t u = 0; u < nt_touched; ++u) {
                int32_t j = touched[u];
                double s = sums[j];
                sums[j] = 0.0; mark[j] = 0;
                if (!(s > lower_bound)) continue;
                if (self_match && (int64_t)j == self_j) continue;
                cand_t c; c.v = s; c.j = j;
                topk_insert(top, &cnt, k, c);
            }
            for (int r = 0; r < k; ++r) {
                if (r < cnt) { top_idx[(size_t)i * k + r] = (int32_t)(top[r].j + to_index_base); top_val[(size_t)i * k + r] = top[r].v; }
                else         { top_idx[(size_t)i * k + r] = -1; top_val[(size_t)i * k + r] = 0.0; }
            }
        }
        }
        free(sums); free(touched); free(mark); free(top);
    }
    return err;
}

/* Merge of per-shard top-k lists (the multi-GPU exchange step has no reference equivalent;
 * this is the CPU statement of the same (score desc, idx asc) key used to check pfz_topk_merge). */
int oracle_topk_merge(int32_t n_shards, int32_t n_from, int32_t k_in, int32_t k_out,
                      const int32_t *idx, const double *val, /* [n_shards][n_from][k_in] */
                      int32_t *out_idx, double *out_val)
{
    cand_t *top = (cand_t *)malloc(sizeof(cand_t) * (size_t)(k_out > 0 ? k_out : 1));
    if (!top) return 1;
    for (int32_t i = 0; i < n_from; ++i) {
        int cnt = 0;
        for (int32_t s = 0; s < n_shards; ++s)
            for (int32_t r = 0; r < k_in; ++r) {
                size_t o = ((size_t)s * n_from + i) * k_in + r;
                if (idx[o] < 0) continue;
                cand_t c; c.v = val[o]; c.j = idx[o];
                topk_insert(top, &cnt, k_out, c);
            }
        for (int r = 0; r < k_out; ++r) {
            if (r < cnt) { out_idx[(size_t)i * k_out + r] = top[r].j; out_val[(size_t)i * k_out + r] = top[r].v; }
            else         { out_idx[(size_t)i * k_out + r] = -1; out_val[(size_t)i * k_out + r] = 0.0; }
        }
    }
    free(top);
    return 0;
}
