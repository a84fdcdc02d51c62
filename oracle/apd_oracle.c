/*
 * apd_oracle.c -- CPU restatement of the reference's all-pairs DTW + UPGMA handoff.
 * TEST INFRASTRUCTURE ONLY (see apd_oracle.h).  Build: oracle/Makefile
 * (gcc -O2 -ffp-contract=off -fno-fast-math: Rust never contracts a*b+c into an
 * FMA nor reassociates, and neither may this file).
 *
 * Citations are file:line in /root/reference/.
 */
#include "apd_oracle.h"

#include <math.h>
#include <pthread.h>
#include <stdlib.h>
#include <string.h>

/* ------------------------------------------------------------------------- */
/* numerics.rs                                                               */
/* ------------------------------------------------------------------------- */

/* src/numerics.rs:114-120 -- sequential f32 accumulation from 0.0 of
 * powf(x-y, 2.0) (LLVM folds powf(v, 2.0) to v*v in the --release build that
 * generate_report.sh:4 makes), then IEEE sqrt. */
float apd_oracle_euclidean(const float *x, const float *y, size_t dim)
{
    float distance = 0.0f;
    for (size_t i = 0; i < dim; i++) {
        float t = x[i] - y[i];
        distance += t * t;
    }
    return sqrtf(distance);
}

/* src/numerics.rs:138-144 */
static uint64_t abs_u(uint64_t n, uint64_t m) { return n > m ? n - m : m - n; }
/* src/numerics.rs:149-155 */
static uint64_t diff_u(uint64_t n, uint64_t m) { return n > m ? n - m : 0; }

/* Rust `f32 as usize`: truncation toward zero, saturating, NaN -> 0. */
static uint64_t f32_as_usize(float v)
{
    if (!(v == v)) return 0;
    if (v <= 0.0f) return 0;
    if (v >= 18446744073709551616.0f) return UINT64_MAX;
    return (uint64_t)v;
}

/* src/discovery.rs:40 */
uint64_t apd_oracle_warping_band(float pct, uint64_t len)
{
    float prod = pct * (float)len;
    return f32_as_usize(prod);
}

/* src/alignments.rs:173 (wrapping add as in a --release build) */
uint64_t apd_oracle_window(uint64_t band, uint64_t n, uint64_t m)
{
    uint64_t a = abs_u(n, m);
    return (band > a ? band : a) + 2;
}

/* src/alignments.rs:174-175 */
uint64_t apd_oracle_cells_visited(uint64_t n, uint64_t m, uint64_t w)
{
    uint64_t cells = 0;
    for (uint64_t i = 1; i <= n; i++) {
        uint64_t lo = diff_u(i, w);
        if (lo < 1) lo = 1;
        uint64_t hi = i + w; /* exclusive */
        if (hi > m + 1) hi = m + 1;
        if (hi > lo) cells += hi - lo;
    }
    return cells;
}

/* ------------------------------------------------------------------------- */
/* sparse (i,j) -> f32 map standing in for HashMap<(usize,usize),f32>        */
/* (src/alignments.rs:100-104).  Open addressing, doubled at 7/8 load like   */
/* hashbrown; the hash is cheaper than SipHash, which only flatters the CPU  */
/* baseline.                                                                 */
/* ------------------------------------------------------------------------- */
typedef struct {
    uint64_t i, j;
    float v;
    uint32_t used;
} slot_t;

typedef struct {
    slot_t *slots;
    uint64_t cap; /* power of two */
    uint64_t len;
} sparse_t;

static uint64_t mix(uint64_t i, uint64_t j)
{
    uint64_t h = i * 0x9E3779B97F4A7C15ull ^ (j + 0x7F4A7C15ull) * 0xC2B2AE3D27D4EB4Full;
    h ^= h >> 29;
    h *= 0xBF58476D1CE4E5B9ull;
    h ^= h >> 32;
    return h;
}

static int sparse_init(sparse_t *s)
{
    s->cap = 16;
    s->len = 0;
    s->slots = (slot_t *)calloc(s->cap, sizeof(slot_t));
    return s->slots ? 0 : -1;
}

static void sparse_free(sparse_t *s) { free(s->slots); s->slots = NULL; }

static const float *sparse_get(const sparse_t *s, uint64_t i, uint64_t j)
{
    uint64_t mask = s->cap - 1;
    uint64_t k = mix(i, j) & mask;
    for (;;) {
        const slot_t *e = &s->slots[k];
        if (!e->used) return NULL;
        if (e->i == i && e->j == j) return &e->v;
        k = (k + 1) & mask;
    }
}

static int sparse_insert(sparse_t *s, uint64_t i, uint64_t j, float v);

static int sparse_grow(sparse_t *s)
{
    sparse_t bigger;
    bigger.cap = s->cap * 2;
    bigger.len = 0;
    bigger.slots = (slot_t *)calloc(bigger.cap, sizeof(slot_t));
    if (!bigger.slots) return -1;
    for (uint64_t k = 0; k < s->cap; k++)
        if (s->slots[k].used) sparse_insert(&bigger, s->slots[k].i, s->slots[k].j, s->slots[k].v);
    free(s->slots);
    *s = bigger;
    return 0;
}

static int sparse_insert(sparse_t *s, uint64_t i, uint64_t j, float v)
{
    if ((s->len + 1) * 8 > s->cap * 7)
        if (sparse_grow(s)) return -1;
    uint64_t mask = s->cap - 1;
    uint64_t k = mix(i, j) & mask;
    for (;;) {
        slot_t *e = &s->slots[k];
        if (!e->used) {
            e->used = 1; e->i = i; e->j = j; e->v = v;
            s->len++;
            return 0;
        }
        if (e->i == i && e->j == j) { e->v = v; return 0; }
        k = (k + 1) & mask;
    }
}

/* ------------------------------------------------------------------------- */
/* alignments.rs: Alignment                                                  */
/* ------------------------------------------------------------------------- */

/* The three-way select of src/alignments.rs:153-159.  Returns the branch:
 * 0 = match (i-1,j-1), 1 = insertion (i-1,j), 2 = deletion (i,j-1). */
static int select_branch(float match_score, float insert_score, float delete_score)
{
    if (delete_score < match_score && delete_score < insert_score) return 2;
    if (insert_score < match_score && insert_score < delete_score) return 1;
    return 0;
}

static float node_value(int branch, float match_score, float insert_score, float delete_score,
                        float distance, const apd_oracle_params *p)
{
    /* penalty * distance is rounded to f32 before the add (no FMA). */
    if (branch == 2) { float t = p->deletion_penalty * distance; return delete_score + t; }
    if (branch == 1) { float t = p->insertion_penalty * distance; return insert_score + t; }
    { float t = p->match_penalty * distance; return match_score + t; }
}

static float lookup(const sparse_t *s, uint64_t i, uint64_t j)
{
    const float *v = sparse_get(s, i, j);
    return v ? *v : INFINITY; /* src/alignments.rs:139-152 */
}

float apd_oracle_dtw_literal(const float *x, uint64_t n, const float *y, uint64_t m,
                             uint64_t dim, const apd_oracle_params *p,
                             uint32_t *path_ij, uint64_t path_cap, uint64_t *path_len)
{
    sparse_t sp;
    if (path_len) *path_len = 0;
    if (sparse_init(&sp)) return NAN;
    sparse_insert(&sp, 0, 0, 0.0f); /* src/alignments.rs:107-111 */

    /* src/alignments.rs:165-180 */
    uint64_t w = apd_oracle_window(p->warping_band, n, m);
    for (uint64_t i = 1; i <= n; i++) {
        uint64_t lo = diff_u(i, w);
        if (lo < 1) lo = 1;
        uint64_t hi = i + w;
        if (hi < i) hi = UINT64_MAX; /* band wider than the address space: clamp */
        if (hi > m + 1) hi = m + 1;
        for (uint64_t j = lo; j < hi; j++) {
            /* src/alignments.rs:129-160 */
            float distance = apd_oracle_euclidean(x + (i - 1) * dim, y + (j - 1) * dim, dim);
            float match_score = lookup(&sp, i - 1, j - 1);
            float insert_score = lookup(&sp, i - 1, j);
            float delete_score = lookup(&sp, i, j - 1);
            int b = select_branch(match_score, insert_score, delete_score);
            float node = node_value(b, match_score, insert_score, delete_score, distance, p);
            sparse_insert(&sp, i, j, node);
        }
    }

    /* src/alignments.rs:116-125; n-1 / m-1 wrap in --release and then miss. */
    float score;
    if (m == 0 && n == 0) {
        score = INFINITY;
    } else if (n == 0 || m == 0) {
        score = INFINITY;
    } else {
        const float *v = sparse_get(&sp, n - 1, m - 1);
        score = v ? *v / (float)(n + m) : INFINITY;
        if (v && path_len && n >= 2 && m >= 2) {
            /* Trace-back (our definition). */
            uint64_t i = n - 1, j = m - 1, len = 0;
            uint64_t guard = n + m + 2;
            while (i >= 1 && j >= 1 && guard--) {
                if (path_ij && len < path_cap) {
                    path_ij[2 * len] = (uint32_t)i;
                    path_ij[2 * len + 1] = (uint32_t)j;
                }
                len++;
                int b = select_branch(lookup(&sp, i - 1, j - 1), lookup(&sp, i - 1, j),
                                      lookup(&sp, i, j - 1));
                if (b == 2) j -= 1;
                else if (b == 1) i -= 1;
                else { i -= 1; j -= 1; }
            }
            *path_len = len;
        }
    }
    sparse_free(&sp);
    return score;
}

float apd_oracle_dtw_dense(const float *x, uint64_t n, const float *y, uint64_t m,
                           uint64_t dim, const apd_oracle_params *p)
{
    if (n == 0 || m == 0) return INFINITY;
    uint64_t w = apd_oracle_window(p->warping_band, n, m);
    /* Two full rows of m+1 columns; cells outside the band are reset to +INF
     * as the band slides, which is what a missing map entry reads as. */
    float *prev = (float *)malloc((m + 2) * sizeof(float));
    float *cur = (float *)malloc((m + 2) * sizeof(float));
    float target = INFINITY;
    int have_target = 0;
    if (!prev || !cur) { free(prev); free(cur); return NAN; }
    for (uint64_t j = 0; j <= m + 1; j++) { prev[j] = INFINITY; cur[j] = INFINITY; }
    prev[0] = 0.0f; /* (0,0) */
    if (n - 1 == 0 && m - 1 == 0) { target = 0.0f; have_target = 1; }
    uint64_t prev_lo = 0, prev_hi = 1; /* row 0 holds column 0 only */
    for (uint64_t i = 1; i <= n; i++) {
        uint64_t lo = diff_u(i, w);
        if (lo < 1) lo = 1;
        uint64_t hi = i + w;
        if (hi < i) hi = UINT64_MAX;
        if (hi > m + 1) hi = m + 1;
        /* cur still holds row i-2, but every read below is guarded by the span of the row it
         * belongs to (prev_lo/prev_hi for row i-1, lo for this row), so nothing stale is read */
        for (uint64_t j = lo; j < hi; j++) {
            float distance = apd_oracle_euclidean(x + (i - 1) * dim, y + (j - 1) * dim, dim);
            float match_score = (j - 1 >= prev_lo && j - 1 < prev_hi) ? prev[j - 1] : INFINITY;
            float insert_score = (j >= prev_lo && j < prev_hi) ? prev[j] : INFINITY;
            float delete_score = (j - 1 >= lo) ? cur[j - 1] : INFINITY;
            int b = select_branch(match_score, insert_score, delete_score);
            cur[j] = node_value(b, match_score, insert_score, delete_score, distance, p);
        }
        if (i == n - 1 && m - 1 >= lo && m - 1 < hi) { target = cur[m - 1]; have_target = 1; }
        float *t = prev; prev = cur; cur = t;
        prev_lo = lo; prev_hi = hi > lo ? hi : lo;
    }
    free(prev); free(cur);
    return have_target ? target / (float)(n + m) : INFINITY;
}

/* ------------------------------------------------------------------------- */
/* alignments.rs: AlignmentWorkers::align_all                                */
/* ------------------------------------------------------------------------- */
typedef struct {
    const float *const *frames;
    const uint32_t *lens;
    uint32_t n, dim;
    float pct, ins, del, mat;
    int variant;
    /* row-block mode */
    uint32_t start, stop;
    float *out_nxn;
    /* pair-list mode */
    const uint32_t *pairs;
    uint64_t p_start, p_stop;
    float *out_pairs;
} job_t;

static float one_pair(const job_t *jb, uint32_t i, uint32_t j)
{
    /* src/alignments.rs:52-57 */
    uint64_t li = jb->lens[i], lj = jb->lens[j];
    uint64_t len = li > lj ? li : lj;
    apd_oracle_params p;
    p.warping_band = apd_oracle_warping_band(jb->pct, len); /* src/discovery.rs:38-45 */
    p.insertion_penalty = jb->ins;
    p.deletion_penalty = jb->del;
    p.match_penalty = jb->mat;
    if (jb->variant == 0)
        return apd_oracle_dtw_literal(jb->frames[i], li, jb->frames[j], lj, jb->dim, &p, NULL, 0, NULL);
    return apd_oracle_dtw_dense(jb->frames[i], li, jb->frames[j], lj, jb->dim, &p);
}

static void *row_worker(void *arg)
{
    job_t *jb = (job_t *)arg;
    for (uint32_t i = jb->start; i < jb->stop; i++)
        for (uint32_t j = 0; j < jb->n; j++)
            if (i != j) jb->out_nxn[(size_t)i * jb->n + j] = one_pair(jb, i, j);
    return NULL;
}

static void *pair_worker(void *arg)
{
    job_t *jb = (job_t *)arg;
    for (uint64_t k = jb->p_start; k < jb->p_stop; k++)
        jb->out_pairs[k] = one_pair(jb, jb->pairs[2 * k], jb->pairs[2 * k + 1]);
    return NULL;
}

int apd_oracle_align_all(const float *const *frames, const uint32_t *lens, uint32_t n,
                         uint32_t dim, float pct, float ins, float del, float mat,
                         uint32_t workers, int variant, float *out_nxn)
{
    if (workers == 0) return -1; /* n / 0 panics at src/alignments.rs:33 */
    memset(out_nxn, 0, (size_t)n * n * sizeof(float)); /* src/alignments.rs:20-23 */
    uint32_t batch_size = n / workers + 1;             /* src/alignments.rs:33 */
    job_t *jobs = (job_t *)calloc(workers, sizeof(job_t));
    pthread_t *th = (pthread_t *)calloc(workers, sizeof(pthread_t));
    int *started = (int *)calloc(workers, sizeof(int));
    if (!jobs || !th || !started) { free(jobs); free(th); free(started); return -2; }
    for (uint32_t b = 0; b < workers; b++) {
        uint64_t start = (uint64_t)b * batch_size;
        uint64_t stop = (uint64_t)(b + 1) * batch_size;
        if (stop > n) stop = n;
        if (start >= stop) continue; /* empty range: the Rust thread does nothing */
        job_t *jb = &jobs[b];
        jb->frames = frames; jb->lens = lens; jb->n = n; jb->dim = dim;
        jb->pct = pct; jb->ins = ins; jb->del = del; jb->mat = mat; jb->variant = variant;
        jb->start = (uint32_t)start; jb->stop = (uint32_t)stop; jb->out_nxn = out_nxn;
        if (pthread_create(&th[b], NULL, row_worker, jb) == 0) started[b] = 1;
        else row_worker(jb);
    }
    for (uint32_t b = 0; b < workers; b++)
        if (started[b]) pthread_join(th[b], NULL);
    free(jobs); free(th); free(started);
    return 0;
}

int apd_oracle_align_pairs(const float *const *frames, const uint32_t *lens, uint32_t n,
                           uint32_t dim, float pct, float ins, float del, float mat,
                           const uint32_t *pairs_ij, uint64_t n_pairs, uint32_t workers,
                           int variant, float *out)
{
    if (workers == 0) return -1;
    for (uint64_t k = 0; k < 2 * n_pairs; k++)
        if (pairs_ij[k] >= n) return -3;
    job_t *jobs = (job_t *)calloc(workers, sizeof(job_t));
    pthread_t *th = (pthread_t *)calloc(workers, sizeof(pthread_t));
    int *started = (int *)calloc(workers, sizeof(int));
    if (!jobs || !th || !started) { free(jobs); free(th); free(started); return -2; }
    uint64_t per = (n_pairs + workers - 1) / workers;
    for (uint32_t b = 0; b < workers; b++) {
        uint64_t start = (uint64_t)b * per, stop = start + per;
        if (stop > n_pairs) stop = n_pairs;
        if (start >= stop) continue;
        job_t *jb = &jobs[b];
        jb->frames = frames; jb->lens = lens; jb->n = n; jb->dim = dim;
        jb->pct = pct; jb->ins = ins; jb->del = del; jb->mat = mat; jb->variant = variant;
        jb->pairs = pairs_ij; jb->p_start = start; jb->p_stop = stop; jb->out_pairs = out;
        if (pthread_create(&th[b], NULL, pair_worker, jb) == 0) started[b] = 1;
        else pair_worker(jb);
    }
    for (uint32_t b = 0; b < workers; b++)
        if (started[b]) pthread_join(th[b], NULL);
    free(jobs); free(th); free(started);
    return 0;
}

/* ------------------------------------------------------------------------- */
/* numerics.rs: percentile; clustering.rs: UPGMA                             */
/* ------------------------------------------------------------------------- */
static int cmp_f32(const void *a, const void *b)
{
    float x = *(const float *)a, y = *(const float *)b;
    return (x > y) - (x < y);
}

/* src/numerics.rs:125-133 */
int apd_oracle_percentile(const float *x, uint64_t len, float perc, float *out)
{
    float nf = (float)len * perc;
    float *numbers = (float *)malloc((len ? len : 1) * sizeof(float));
    if (!numbers) return -2;
    uint64_t cnt = 0;
    for (uint64_t k = 0; k < len; k++)
        if (x[k] == x[k]) numbers[cnt++] = x[k];
    qsort(numbers, cnt, sizeof(float), cmp_f32);
    uint64_t idx = f32_as_usize(nf);
    if (idx >= cnt) { free(numbers); return -1; } /* index out of bounds panic */
    *out = numbers[idx];
    free(numbers);
    return 0;
}

typedef struct {
    uint32_t *parents;
    uint32_t n_parents;
    const float *distances;
    uint32_t n_instances;
    uint32_t n_clusters;
} dendro_t;

/* src/clustering.rs:115-121 */
static uint32_t cluster_of(const dendro_t *d, uint32_t i)
{
    uint32_t p = i;
    while (p != d->parents[p]) p = d->parents[p];
    return p;
}

/* src/clustering.rs:153-170 -- one f32 accumulator over x-major (x asc, y asc). */
static float linkage(const dendro_t *d, const uint32_t *assignment, uint32_t i, uint32_t j)
{
    float size_x = 0.0f, size_y = 0.0f, distance = 0.0f;
    uint32_t n = d->n_instances;
    for (uint32_t x = 0; x < n; x++) {
        if (assignment[x] == i) {
            size_y = 0.0f;
            for (uint32_t y = 0; y < n; y++) {
                if (assignment[y] == j) {
                    distance += d->distances[(size_t)x * n + y];
                    size_y += 1.0f;
                }
            }
            size_x += 1.0f;
        }
    }
    return distance / (size_x * size_y);
}

static int cmp_u32(const void *a, const void *b)
{
    uint32_t x = *(const uint32_t *)a, y = *(const uint32_t *)b;
    return (x > y) - (x < y);
}

int apd_oracle_upgma(const float *dist_nxn, uint32_t n, float perc, apd_oracle_merge *ops,
                     uint32_t *n_ops, float *threshold_out, uint32_t *assignment_out)
{
    dendro_t d;
    *n_ops = 0;
    d.parents = (uint32_t *)malloc((2 * (size_t)n + 1) * sizeof(uint32_t));
    uint32_t *assignment = (uint32_t *)malloc((n ? n : 1) * sizeof(uint32_t));
    uint32_t *roots = (uint32_t *)malloc((n ? n : 1) * sizeof(uint32_t));
    if (!d.parents || !assignment || !roots) { free(d.parents); free(assignment); free(roots); return -2; }
    for (uint32_t i = 0; i < n; i++) d.parents[i] = i; /* src/clustering.rs:88-91 */
    d.n_parents = n;
    d.distances = dist_nxn;
    d.n_instances = n;
    d.n_clusters = n;

    float threshold;
    int rc = apd_oracle_percentile(dist_nxn, (uint64_t)n * n, perc, &threshold); /* :101 */
    if (rc) { free(d.parents); free(assignment); free(roots); return rc; }
    if (threshold_out) *threshold_out = threshold;

    float distance = 0.0f;
    while (d.n_clusters > 1 && distance < threshold) { /* src/clustering.rs:104 */
        /* merge(): src/clustering.rs:175-209 */
        for (uint32_t i = 0; i < n; i++) assignment[i] = cluster_of(&d, i);
        uint32_t n_roots = 0;
        memcpy(roots, assignment, n * sizeof(uint32_t));
        qsort(roots, n, sizeof(uint32_t), cmp_u32);
        for (uint32_t i = 0; i < n; i++)
            if (i == 0 || roots[i] != roots[i - 1]) roots[n_roots++] = roots[i];

        float min_linkage = INFINITY;
        uint32_t min_p = 0, min_q = 0, tie = 0;
        for (uint32_t a = 0; a < n_roots; a++) {
            for (uint32_t b = 0; b < n_roots; b++) {
                if (roots[a] == roots[b]) continue;
                float l = linkage(&d, assignment, roots[a], roots[b]);
                if (l < min_linkage) {
                    min_linkage = l; min_p = roots[a]; min_q = roots[b]; tie = 0;
                } else if (l == min_linkage &&
                           !((roots[a] == min_q && roots[b] == min_p))) {
                    /* a different unordered pair reaches the same minimum: the
                     * reference's answer would depend on HashSet order */
                    tie = 1;
                }
            }
        }
        /* merge_clusters(): src/clustering.rs:133-141 */
        uint32_t k = d.n_parents;
        d.parents[min_p] = k;
        d.parents[min_q] = k;
        d.parents[k] = k;
        d.n_parents++;
        d.n_clusters--;
        apd_oracle_merge *op = &ops[(*n_ops)++];
        op->merge_i = min_p; op->merge_j = min_q; op->into = k;
        op->distance = min_linkage; op->tie = tie;
        distance = min_linkage;
    }
    if (assignment_out)
        for (uint32_t i = 0; i < n; i++) assignment_out[i] = cluster_of(&d, i);
    free(d.parents); free(assignment); free(roots);
    return 0;
}

/* ------------------------------------------------------------------------- */
/* neural.rs / spectrogram.rs: the embedding step in front of the DTW path    */
/* ------------------------------------------------------------------------- */

/* AutoEncoder::predict on ONE frame (src/neural.rs:55-71), which is how
 * NDSequence::encoded calls it (src/spectrogram.rs:103-121: a 1 x n_bins Mat per frame):
 *   Mat::mul      src/numerics.rs:305-319  flat[j] += x[k] * w[k*cols + j], k ascending from 0.0
 *                                          (separate multiply and add: Rust never contracts)
 *   add_col       src/numerics.rs:246-257  += b[j]
 *   sigmoid       src/numerics.rs:222-232  1.0 / (1.0 + f32::exp(-x))   (libm expf)
 *   scale(255.0)  src/numerics.rs:296-301
 *   mean / std    src/numerics.rs:12-29    sequential sums over the n_latent values, /len;
 *                                          powf(v - mu, 2.0) folds to a product
 *   sigma = max(std, 1.0); z_score (src/numerics.rs:71-73): (x - mu) / sigma
 * w_encode is n_bins x n_latent row-major (Mat{flat, cols: n_latent}), b_encode 1 x n_latent. */
void apd_oracle_ae_predict(const float *x, uint32_t n_bins, const float *w_encode,
                           const float *b_encode, uint32_t n_latent, float *out)
{
    for (uint32_t j = 0; j < n_latent; j++) {
        float acc = 0.0f;
        for (uint32_t k = 0; k < n_bins; k++) {
            float prod = x[k] * w_encode[(size_t)k * n_latent + j];
            acc += prod;
        }
        acc += b_encode[j];
        float s = 1.0f / (1.0f + expf(-acc));
        out[j] = s * 255.0f;
    }
    float mu = 0.0f;
    for (uint32_t j = 0; j < n_latent; j++) mu += out[j];
    mu = mu / (float)n_latent;
    float sd = 0.0f;
    for (uint32_t j = 0; j < n_latent; j++) {
        float t = out[j] - mu;
        sd += t * t;
    }
    sd = sqrtf(sd / (float)n_latent);
    float sigma = fmaxf(sd, 1.0f); /* f32::max: a NaN std yields 1.0, like fmaxf */
    for (uint32_t j = 0; j < n_latent; j++) out[j] = (out[j] - mu) / sigma;
}

/* NDSequence::encoded (src/spectrogram.rs:103-121): predict() frame by frame. */
void apd_oracle_ae_encode(const float *frames, uint64_t len, uint32_t n_bins, const float *w_encode,
                          const float *b_encode, uint32_t n_latent, float *out)
{
    for (uint64_t t = 0; t < len; t++)
        apd_oracle_ae_predict(frames + t * n_bins, n_bins, w_encode, b_encode, n_latent, out + t * n_latent);
}
