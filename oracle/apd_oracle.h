/*
 * apd_oracle.h -- CPU oracle for the all-pairs banded weighted DTW path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under audio_pattern_discovery_b200/ may
 * include, link or call this; only tests/, __graft_entry__.smoke() and the
 * cpu_baseline / --impl reference legs of bench.py do, and only as the checker
 * or as the reported CPU baseline.
 *
 * PARITY STATUS: "parity unpinned" by the reference's own tests -- the
 * reference (dkohlsdorf/audio_pattern_discovery, Rust) ships no tests, golden
 * vectors or fixtures and cannot be compiled here (no
 * cargo/rustc in the image, Cargo.lock absent).  The oracle is therefore pinned
 * by (1) the hand-derived known-answer tests of SURVEY.md Appendix B and
 * (2) bit-for-bit agreement of three independent restatements on a committed
 * seeded corpus: the hash-map literal and the dense rolling-band version in
 * this file, and the pure-Python transliteration in oracle/literal.py.
 *
 * All file:line citations are relative to /root/reference/.
 */
#ifndef APD_ORACLE_H
#define APD_ORACLE_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* src/alignments.rs:77-83 */
typedef struct {
    uint64_t warping_band;
    float insertion_penalty;
    float deletion_penalty;
    float match_penalty;
} apd_oracle_params;

/* src/numerics.rs:114-120 */
float apd_oracle_euclidean(const float *x, const float *y, size_t dim);

/* src/discovery.rs:38-45: (pct * len as f32) as usize, saturating, NaN -> 0 */
uint64_t apd_oracle_warping_band(float pct, uint64_t len);

/* src/alignments.rs:173: w = max(band, abs(n, m)) + 2 */
uint64_t apd_oracle_window(uint64_t band, uint64_t n, uint64_t m);

/* cells the reference visits for one ordered pair
 * (src/alignments.rs:174-175). */
uint64_t apd_oracle_cells_visited(uint64_t n, uint64_t m, uint64_t w);

/*
 * Literal restatement of Alignment::{new,construct_alignment,score}
 * (src/alignments.rs:106-180): sparse (i,j)->f32 map seeded with (0,0)=0,
 * missing cells read as +INF, strict-< three-way select, score read from
 * (n-1, m-1) and divided by (n+m).
 *
 * If path_ij != NULL the warping path is traced back from (n-1, m-1) to the
 * cell whose predecessor is (0,0), re-applying the forward rule at every cell
 * (SURVEY.md Appendix A.8 -- the reference keeps `sparse` but never traces it;
 * this definition is ours).  Pairs are written end-to-start as (i, j), 1-based,
 * at most path_cap pairs; *path_len receives the full length (0 if the score
 * cell is absent).
 */
float apd_oracle_dtw_literal(const float *x, uint64_t n, const float *y, uint64_t m,
                             uint64_t dim, const apd_oracle_params *p,
                             uint32_t *path_ij, uint64_t path_cap, uint64_t *path_len);

/* Same arithmetic on two rolling dense band rows; must equal the literal bit for bit. */
float apd_oracle_dtw_dense(const float *x, uint64_t n, const float *y, uint64_t m,
                           uint64_t dim, const apd_oracle_params *p);

/*
 * AlignmentWorkers::align_all (src/alignments.rs:31-67): every ordered pair
 * i != j, len = max(len_i, len_j), params from pct (src/discovery.rs:38-45),
 * static row blocks of n / workers + 1 rows per thread, diagonal left at 0.
 * variant 0 = literal (hash map), 1 = dense.  Returns 0 on success,
 * -1 on workers == 0 (the reference divides by it and panics).
 */
int apd_oracle_align_all(const float *const *frames, const uint32_t *lens, uint32_t n,
                         uint32_t dim, float pct, float ins, float del, float mat,
                         uint32_t workers, int variant, float *out_nxn);

/*
 * The body of the pair loop (src/alignments.rs:52-57) for an explicit list of
 * ordered pairs (i, j) = pairs_ij[2k], pairs_ij[2k+1]; out[k] = score.  Used to
 * spot-check matrices too large to recompute on the CPU.  Pairs are dealt to
 * `workers` threads in contiguous blocks.
 */
int apd_oracle_align_pairs(const float *const *frames, const uint32_t *lens, uint32_t n,
                           uint32_t dim, float pct, float ins, float del, float mat,
                           const uint32_t *pairs_ij, uint64_t n_pairs, uint32_t workers,
                           int variant, float *out);

/* src/numerics.rs:125-133: index from the unfiltered length, NaNs dropped. */
int apd_oracle_percentile(const float *x, uint64_t len, float perc, float *out);

/* src/clustering.rs:18-25 */
typedef struct {
    uint32_t merge_i;
    uint32_t merge_j;
    uint32_t into;
    float distance;
    uint32_t tie; /* 1 if another ordered root pair had exactly the same linkage */
} apd_oracle_merge;

/*
 * AgglomerativeClustering::clustering (src/clustering.rs:81-110) with merge /
 * linkage (153-209) restated literally; the root set is iterated in ascending
 * id order where the reference iterates a HashSet (random order), so (p,q)
 * orientation and exact-tie resolution are ours; `tie` flags the latter.
 * ops must hold n-1 entries.  assignment_out (n entries, may be NULL) receives
 * the final root id of each instance.
 */
int apd_oracle_upgma(const float *dist_nxn, uint32_t n, float perc, apd_oracle_merge *ops,
                     uint32_t *n_ops, float *threshold_out, uint32_t *assignment_out);

/*
 * AutoEncoder::predict on one frame (src/neural.rs:55-71) and NDSequence::encoded
 * (src/spectrogram.rs:103-121): the step that produces the reference's real DTW input
 * (SURVEY.md section 8 row f3).  w_encode: n_bins x n_latent row-major, b_encode: n_latent.
 * f32::exp is the C library's expf, as in a Rust build on this platform.
 */
void apd_oracle_ae_predict(const float *x, uint32_t n_bins, const float *w_encode,
                           const float *b_encode, uint32_t n_latent, float *out);
void apd_oracle_ae_encode(const float *frames, uint64_t len, uint32_t n_bins, const float *w_encode,
                          const float *b_encode, uint32_t n_latent, float *out);

#ifdef __cplusplus
}
#endif
#endif
