/* CPU oracle, C restatement.  TEST INFRASTRUCTURE ONLY — never linked into libcldrd.so.
 *
 * PARITY UNPINNED (see oracle/flat_ip.py): restates the published semantics of faiss'
 * IndexIDMap(IndexFlatIP).search as the reference calls it
 * (retriever/retrieval_utils.py:135,143): fp32 inner products, per-query top-k by descending
 * score, ties -> lower row, rows -> ids through id_map, (-FLT_MAX, -1) padding.
 * Written for clarity: one dot product at a time, a binary heap per query, OpenMP over queries.
 * Used by tests to cross-check the numpy oracle and by bench.py as a scalar-code CPU baseline.
 */
#include <float.h>
#include <stdint.h>
#include <stdlib.h>

typedef struct { float s; int64_t r; } hit_t;

/* "a ranks after b": lower score, or equal score and higher row */
static int worse(hit_t a, hit_t b) { return a.s < b.s || (a.s == b.s && a.r > b.r); }

static void sift_down(hit_t* h, int n, int i) {
    for (;;) {
        int l = 2 * i + 1, r = l + 1, m = i;
        if (l < n && worse(h[l], h[m])) m = l;
        if (r < n && worse(h[r], h[m])) m = r;
        if (m == i) return;
        hit_t t = h[i]; h[i] = h[m]; h[m] = t; i = m;
    }
}

static int cmp_best_first(const void* pa, const void* pb) {
    hit_t a = *(const hit_t*)pa, b = *(const hit_t*)pb;
    if (worse(b, a)) return -1;
    if (worse(a, b)) return 1;
    return 0;
}

/* xb [n][d], ids [n] or NULL, xq [nq][d] -> D [nq][k], I [nq][k] */
int oracle_flat_ip_search(const float* xb, const int64_t* ids, int64_t n, int d, const float* xq,
                          int64_t nq, int k, float* D, int64_t* I) {
    int failed = 0;
#pragma omp parallel for schedule(dynamic, 1)
    for (int64_t qi = 0; qi < nq; ++qi) {
        hit_t* heap = (hit_t*)malloc(sizeof(hit_t) * (size_t)k);  /* min-heap of the k best */
        if (!heap) { failed = 1; continue; }
        int hn = 0;
        const float* q = xq + qi * d;
        for (int64_t r = 0; r < n; ++r) {
            const float* b = xb + r * d;
            float acc = 0.f;
            for (int c = 0; c < d; ++c) acc += q[c] * b[c];
            hit_t h = {acc, r};
            if (hn < k) {
                heap[hn++] = h;
                if (hn == k) for (int i = k / 2 - 1; i >= 0; --i) sift_down(heap, hn, i);
            } else if (worse(heap[0], h)) {
                heap[0] = h;
                sift_down(heap, hn, 0);
            }
        }
        qsort(heap, (size_t)hn, sizeof(hit_t), cmp_best_first);
        for (int j = 0; j < k; ++j) {
            if (j < hn) {
                D[qi * k + j] = heap[j].s;
                I[qi * k + j] = ids ? ids[heap[j].r] : heap[j].r;
            } else {
                D[qi * k + j] = -FLT_MAX;
                I[qi * k + j] = -1;
            }
        }
        free(heap);
    }
    return failed;
}
