/*
 * pgv_oracle.c -- TEST INFRASTRUCTURE ONLY.  CPU restatement of the arithmetic the
 * reference delegates to pgvector for the path
 *     ORDER BY embedding_vec <=> :q LIMIT :k        (app/services/vector_store.py:274-287,
 *                                                    app/services/corpus_search.py:1525-1536)
 *
 * PARITY UNPINNED: pgvector's C source is NOT under /root/reference (it is an external
 * Postgres extension, pinned only as `--branch v0.5.1` in
 * scripts/install_pgvector_for_postgresql14.sh:21) and the reference's own tests hold no
 * golden vectors for retrieval.  What follows restates pgvector v0.5.1's
 * published algorithm (src/vector.c, cosine_distance):
 *     float dot = 0, na = 0, nb = 0;                       // float4 accumulation
 *     for i < dim: dot += a[i]*b[i]; na += a[i]*a[i]; nb += b[i]*b[i];
 *     double sim = (double)dot / sqrt((double)na * (double)nb);
 *     clamp sim to [-1, 1];  return 1.0 - sim;             // float8
 * pgvector builds with -ftree-vectorize -fassociative-math, so the summation ORDER is
 * unspecified; this file is compiled with the same flags (oracle/Makefile).
 * The SQL around it: `1 - (v <=> q)` is evaluated in float8; ORDER BY distance ASC puts
 * NaN last (Postgres float8 ordering); ties have no defined order (we break them by row).
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference leg
 * may load this library.  The product never does.
 */
#include <math.h>
#include <pthread.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

/* pgvector src/vector.c cosine_distance(), returns the float8 distance */
double pgv_cosine_distance(const float* a, const float* b, int dim) {
    float distance = 0.0f, norma = 0.0f, normb = 0.0f;
    for (int i = 0; i < dim; i++) {
        distance += a[i] * b[i];
        norma += a[i] * a[i];
        normb += b[i] * b[i];
    }
    double similarity = (double)distance / sqrt((double)norma * (double)normb);
    if (similarity > 1) similarity = 1.0;
    else if (similarity < -1) similarity = -1.0;
    return 1.0 - similarity;
}

/*
 * Sequential scan: distance of q to every row with mask[r] != 0 (mask == NULL: all rows).
 * Rows that do not pass get dist = +inf marker via pass[r] = 0.  X is row-major, pitch ld.
 * One backend process = one thread (pgv_scan); pgv_set_threads(t) makes the scan partition the
 * rows over t pthreads the way a Postgres parallel seq scan partitions the heap.
 */
static int g_threads = 1;
void pgv_set_threads(int t) { g_threads = t < 1 ? 1 : (t > 256 ? 256 : t); }
int pgv_get_threads(void) { return g_threads; }

typedef struct {
    const float* X; int64_t lo, hi; int dim; int64_t ld; const float* q;
    const uint8_t* mask; double* dist;
} scan_job;

static void* scan_range(void* p) {
    scan_job* j = (scan_job*)p;
    for (int64_t r = j->lo; r < j->hi; r++) {
        if (j->mask && !j->mask[r]) { j->dist[r] = INFINITY; continue; }
        j->dist[r] = pgv_cosine_distance(j->X + r * j->ld, j->q, j->dim);
    }
    return NULL;
}

void pgv_scan(const float* X, int64_t n, int dim, int64_t ld, const float* q,
              const uint8_t* mask, double* dist) {
    int t = g_threads;
    if (t <= 1 || n < 4096) {
        scan_job j = { X, 0, n, dim, ld, q, mask, dist };
        scan_range(&j);
        return;
    }
    pthread_t th[256]; scan_job jobs[256];
    int64_t per = (n + t - 1) / t;
    for (int i = 0; i < t; i++) {
        int64_t lo = i * per, hi = lo + per > n ? n : lo + per;
        if (lo > n) lo = n;
        jobs[i] = (scan_job){ X, lo, hi, dim, ld, q, mask, dist };
        pthread_create(&th[i], NULL, scan_range, &jobs[i]);
    }
    for (int i = 0; i < t; i++) pthread_join(th[i], NULL);
}

typedef struct { double d; int64_t r; } ent_t;

/* Postgres float8 ORDER BY ... ASC: NaN sorts after every non-NaN value. Ties -> row asc. */
static int ent_less(const ent_t* a, const ent_t* b) {
    int an = isnan(a->d), bn = isnan(b->d);
    if (an != bn) return bn;            /* non-NaN first */
    if (!an && a->d != b->d) return a->d < b->d;
    return a->r < b->r;
}

static void heap_sift_down(ent_t* h, int64_t n, int64_t i) {
    /* max-heap on ent_less (root = worst kept entry) */
    for (;;) {
        int64_t l = 2 * i + 1, r = l + 1, m = i;
        if (l < n && ent_less(&h[m], &h[l])) m = l;
        if (r < n && ent_less(&h[m], &h[r])) m = r;
        if (m == i) return;
        ent_t t = h[i]; h[i] = h[m]; h[m] = t;
        i = m;
    }
}

static int ent_cmp_qsort(const void* a, const void* b) {
    const ent_t* x = (const ent_t*)a; const ent_t* y = (const ent_t*)b;
    if (ent_less(x, y)) return -1;
    if (ent_less(y, x)) return 1;
    return 0;
}

/*
 * ORDER BY dist ASC LIMIT k over rows with mask[r] != 0 (top-N heapsort, as the
 * executor does for ORDER BY .. LIMIT).  Writes rows and similarity = 1 - dist (float8,
 * the SELECT list expression).  Returns the number of rows produced (<= k).
 */
int64_t pgv_topk(const double* dist, const uint8_t* mask, int64_t n, int64_t k,
                 int64_t* rows_out, double* sim_out) {
    if (k <= 0) return 0;
    ent_t* h = (ent_t*)malloc(sizeof(ent_t) * (size_t)k);
    int64_t m = 0;
    for (int64_t r = 0; r < n; r++) {
        if (mask && !mask[r]) continue;
        ent_t e = { dist[r], r };
        if (m < k) {
            h[m++] = e;
            if (m == k) for (int64_t i = k / 2 - 1; i >= 0; i--) heap_sift_down(h, k, i);
        } else if (ent_less(&e, &h[0])) {
            h[0] = e;
            heap_sift_down(h, k, 0);
        }
    }
    qsort(h, (size_t)m, sizeof(ent_t), ent_cmp_qsort);
    for (int64_t i = 0; i < m; i++) {
        rows_out[i] = h[i].r;
        sim_out[i] = 1.0 - h[i].d;
    }
    free(h);
    return m;
}

/* The whole statement for one query: scan + top-N.  scratch = n doubles. */
int64_t pgv_search(const float* X, int64_t n, int dim, int64_t ld, const float* q,
                   const uint8_t* mask, int64_t k, int64_t* rows_out, double* sim_out,
                   double* scratch) {
    pgv_scan(X, n, dim, ld, q, mask, scratch);
    return pgv_topk(scratch, mask, n, k, rows_out, sim_out);
}

/* nq queries back to back (what nq SQL statements do). */
void pgv_search_batch(const float* X, int64_t n, int dim, int64_t ld, const float* Q, int nq,
                      const uint8_t* mask, int64_t k, int64_t* rows_out, double* sim_out,
                      int64_t* counts_out, double* scratch) {
    for (int i = 0; i < nq; i++) {
        int64_t m = pgv_search(X, n, dim, ld, Q + (int64_t)i * dim, mask, k,
                               rows_out + (int64_t)i * k, sim_out + (int64_t)i * k, scratch);
        counts_out[i] = m;
        for (int64_t j = m; j < k; j++) { rows_out[(int64_t)i * k + j] = -1; sim_out[(int64_t)i * k + j] = NAN; }
    }
}

/* round-to-nearest-even fp32 -> bf16 -> fp32, the storage rounding of MRAG_BF16 mode */
void pgv_round_bf16(const float* in, float* out, int64_t n) {
    for (int64_t i = 0; i < n; i++) {
        uint32_t u; memcpy(&u, &in[i], 4);
        if ((u & 0x7fffffffu) > 0x7f800000u) { u |= 0x00400000u; u &= 0xffff0000u; }
        else { u += 0x7fffu + ((u >> 16) & 1u); u &= 0xffff0000u; }
        memcpy(&out[i], &u, 4);
    }
}

int pgv_oracle_abi(void) { return 1; }
