/*
 * apd.h -- C ABI of libapd_b200: the B200-native replacement for the all-pairs
 * weighted, Sakoe-Chiba-banded DTW distance matrix of
 * dkohlsdorf/audio_pattern_discovery (the path SURVEY.md section 8 scopes).
 *
 * This is the boundary a Rust `apd-sys` crate binds with `extern "C"` (see
 * INTEGRATION.md and rust/); every entry point names the reference interface it
 * replaces (file:line relative to the reference repository root).
 *
 * Conventions: plain pointers and sizes only; every function returns an
 * apd_status (0 = OK) and never unwinds; a context is thread-compatible (one
 * caller at a time); the library copies inputs before returning and writes only
 * into caller-allocated outputs.  There is NO CPU fallback: without a CUDA
 * device apd_create() fails with APD_ERR_NO_DEVICE.
 */
#ifndef APD_H
#define APD_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define APD_ABI_VERSION 2

typedef enum {
    APD_OK = 0,
    APD_ERR_INVALID = 1,     /* bad argument (message in apd_last_error) */
    APD_ERR_NO_DEVICE = 2,   /* no usable CUDA device: this library has no CPU path */
    APD_ERR_CUDA = 3,        /* a CUDA runtime call failed */
    APD_ERR_UNSUPPORTED = 4, /* e.g. dim > APD_MAX_DIM */
    APD_ERR_STATE = 5,       /* call order (no sequences set, ...) */
    APD_ERR_INTERNAL = 6
} apd_status;

#define APD_MAX_DIM 32
#define APD_MAX_DEVICES 8   /* devices of one single-process group (apd_create_multi) */
#define APD_AE_MAX_BINS 64  /* widest input frame of apd_set_sequences_encoded */

/* Arithmetic variants of the distance kernel. */
#define APD_MODE_STRICT 0u /* bit-exact with the reference's f32 operation sequence (default) */
#define APD_MODE_FAST 1u   /* packed-FMA accumulation + approximate sqrt; <= 1e-5 relative */

typedef struct apd_ctx apd_ctx;

/* Replaces src/discovery.rs:38-45 (Discovery::alignment_params) + src/alignments.rs:77-83
 * (AlignmentParams).  The per-pair band is derived on the device exactly as the
 * reference does: band = (warping_band_percentage * max(len_i, len_j) as f32) as usize. */
typedef struct {
    float warping_band_percentage; /* project/config/Discovery.toml:17 */
    float insertion_penalty;       /* :18  weight on predecessor (i-1, j)   */
    float deletion_penalty;        /* :19  weight on predecessor (i, j-1)   */
    float match_penalty;           /* :20  weight on predecessor (i-1, j-1) */
    uint32_t mode;                 /* APD_MODE_* */
} apd_params;

typedef struct {
    uint64_t n_sequences;
    uint64_t ordered_pairs;      /* n(n-1): what src/alignments.rs:50-51 enumerates */
    uint64_t units_total;        /* 32-pair work units over the whole matrix */
    uint64_t units_local;        /* ... handled by this context's shard */
    uint64_t cells_reference;    /* reference cell updates of this shard's ordered pairs
                                    (src/alignments.rs:174-175 visit rule) */
    uint64_t cells_computed;     /* cell updates the kernels actually execute (both orientations) */
    uint32_t kernel_launches;    /* DTW + scatter launches of the last align call */
    float kernel_ms;             /* CUDA-event time of the DTW kernels of the last call */
    float scatter_ms;            /* packed -> n x n scatter kernel */
    float h2d_ms, d2h_ms;        /* host<->device copies inside the last host-buffer call */
    uint64_t h2d_bytes, d2h_bytes;
    float sm_clock_mhz;          /* clock rate the device reports (max), for rooflines */
    uint32_t sm_count;
    float select_ms;             /* radix-select passes of the last apd_percentile_* call */
    float path_ms;               /* forward + trace-back kernels of the last apd_align_pair(s) call */
} apd_stats;

/* ---- lifecycle ------------------------------------------------------------ */

/* One context drives one CUDA device.  Replaces AlignmentWorkers::new's role of
 * owning the data (src/alignments.rs:17-26). */
apd_status apd_create(int device_id, apd_ctx **out);

/* One context that drives n_dev devices of this box from ONE host process -- what the
 * reference's single blocking call `workers.align_all(&discover)` (src/main.rs:189-195,
 * src/alignments.rs:31-67: alignment_workers threads in one process) becomes on a multi-GPU
 * box.  device_ids == NULL: devices 0..n_dev-1; n_dev <= 0: every visible device (at most
 * APD_MAX_DEVICES).  Every entry point that takes host buffers (apd_set_sequences*,
 * apd_align_all, apd_align_pair(s), apd_percentile_matrix, apd_get_stats) works on the group:
 * every device holds the whole sequence arena (one upload, then NVLink copies), computes the
 * work units u with u % n_dev == its index, and stores its packed results straight into every
 * other device's gathered buffer through NVLink peer mappings from inside the DTW kernel (the
 * all-gather is fused into the kernel; without peer access the shards are copied afterwards);
 * each device expands the matrix and copies its own slab of rows to out_nxn.  The
 * device-buffer stages (apd_set_shard, apd_packed_len, apd_align_packed, apd_scatter_packed)
 * belong to the one-process-per-GPU form and return APD_ERR_UNSUPPORTED on a group. */
apd_status apd_create_multi(const int *device_ids, int n_dev, apd_ctx **out);
apd_status apd_device_count(int *count); /* visible CUDA devices; APD_ERR_NO_DEVICE if none */
/* Devices of the context's group (1 for apd_create); *peer_stores (may be NULL) = 1 if the
 * kernels store into peer memory directly. */
apd_status apd_group_size(apd_ctx *ctx, uint32_t *n_dev, uint32_t *peer_stores);
void apd_destroy(apd_ctx *ctx);
const char *apd_last_error(const apd_ctx *ctx); /* ctx may be NULL: last create error */
uint32_t apd_abi_version(void);

/* ---- sequence packing (the spectrogram.rs / discovery.rs glue) ------------- */

/* Replaces the Vec<NDSequence> hand-over of src/main.rs:189 and the NDSequence
 * layout contract (src/spectrogram.rs:13-24, vec() 99-101, len() 152-154):
 * frames[s] points at lens[s] * dim row-major f32 values.  Sequences are copied,
 * sorted by length and packed into one 16-byte-aligned device arena. */
apd_status apd_set_sequences(apd_ctx *ctx, const float *const *frames, const uint32_t *lens,
                             uint32_t n, uint32_t dim);

/* Same, from one flat buffer: sequence s starts at flat + offsets[s] (in floats). */
apd_status apd_set_sequences_flat(apd_ctx *ctx, const float *flat, const uint64_t *offsets,
                                  const uint32_t *lens, uint32_t n, uint32_t dim);

/* The embedding step in front of the path, on the device:
 * replaces `NDSequence::new(..).encoded(&nn)` of src/main.rs:150-161, i.e.
 * NDSequence::encoded (src/spectrogram.rs:103-121) = AutoEncoder::predict (src/neural.rs:55-71)
 * frame by frame: sigmoid(x W + b) * 255, then the per-frame z-score with sigma = max(std, 1.0),
 * every f32 operation in the reference's order.  cepstra[s] points at lens[s] * n_bins
 * row-major f32 values (n_bins <= APD_AE_MAX_BINS; 26 in the reference, src/spectrogram.rs:76);
 * w_encode is n_bins x n_latent row-major (Mat{flat, cols}, src/numerics.rs:169-173), b_encode
 * n_latent values (n_latent <= APD_MAX_DIM; 10 in project/config/Discovery.toml:5).  The
 * embeddings are written straight into the arena (frame width n_latent); they never exist on
 * the host.  apd_get_sequence reads one sequence of the arena back (lens * dim floats). */
apd_status apd_set_sequences_encoded(apd_ctx *ctx, const float *const *cepstra, const uint32_t *lens,
                                     uint32_t n, uint32_t n_bins, const float *w_encode,
                                     const float *b_encode, uint32_t n_latent);
apd_status apd_get_sequence(apd_ctx *ctx, uint32_t index, float *out, uint64_t cap_floats);

/* One-process-per-GPU jobs: every rank declares the same batch, ONE rank uploads it with apd_set_sequences and the
 * others receive the packed arena over NVLink (ncclBroadcast) instead of packing and uploading it again:
 *   apd_set_sequences_layout  the batch's lengths and frame width only -- builds the same arena layout and tables every
 *                             rank derives from them, allocates the arena, leaves its contents undefined;
 *   apd_arena_device          the arena's device pointer and size in floats (also valid after apd_set_sequences on the
 *                             uploading rank, once apd_synchronize(ctx, 0) has returned);
 *   apd_arena_commit          the caller has filled the arena, in stream order before any later call that names the
 *                             same stream (or synchronised): the batch is ready.
 * The layout is a pure function of (lens, dim), so the arenas of all ranks are byte-identical. */
apd_status apd_set_sequences_layout(apd_ctx *ctx, const uint32_t *lens, uint32_t n, uint32_t dim);
apd_status apd_arena_device(apd_ctx *ctx, void **d_arena, uint64_t *n_floats);
apd_status apd_arena_commit(apd_ctx *ctx);

/* Multi-process sharding (one process per GPU, e.g. under torchrun): this
 * context computes work units u with u % world == rank.  Default 0 / 1. */
apd_status apd_set_shard(apd_ctx *ctx, uint32_t rank, uint32_t world);

/* ---- the hot path ---------------------------------------------------------- */

/* Replaces AlignmentWorkers::align_all + the result hand-over
 * (src/alignments.rs:31-67, src/main.rs:191-195): fills out_nxn (host memory,
 * n*n floats, row-major, result[i*n+j] = score(data[i], data[j]), diagonal 0.0).
 * With world > 1 only this shard's pairs are non-zero; use the device-side calls
 * below plus an all-gather to assemble the full matrix. */
apd_status apd_align_all(apd_ctx *ctx, const apd_params *p, float *out_nxn);

/* Device-resident stages of the same call, for callers that own device buffers
 * (PyTorch tensors, NCCL):
 *   1. apd_packed_len: floats in one shard's packed result buffer (identical on
 *      every rank so that the buffers can be all-gathered; depends on p only
 *      through warping_band_percentage);
 *   2. apd_align_packed: run the DTW kernels of this shard, results ->
 *      d_packed (device pointer, apd_packed_len floats);
 *   3. apd_scatter_packed: expand `world` gathered shards (device pointer,
 *      world * apd_packed_len floats, rank-major; world == 1 on a sharded context
 *      means "this shard's buffer only") into the device matrix d_out_nxn (n*n
 *      floats, diagonal 0).
 * All work is enqueued on `stream` (a cudaStream_t).  stream == 0 / NULL does NOT mean the
 * legacy default stream: it selects the context's own private non-blocking stream, which has
 * no implicit ordering with any other stream -- call apd_synchronize(ctx, 0) (or pass your own
 * stream everywhere) before another stream touches d_packed / d_out_nxn.  Calls on one
 * context are ordered among themselves whatever streams they name: each one waits (on the
 * device) for the work the previous one enqueued. */
apd_status apd_packed_len(apd_ctx *ctx, const apd_params *p, uint64_t *n_floats);
apd_status apd_align_packed(apd_ctx *ctx, const apd_params *p, float *d_packed, void *stream);
apd_status apd_scatter_packed(apd_ctx *ctx, const float *d_gathered, uint32_t world,
                              float *d_out_nxn, void *stream);
/* Waits for `stream` (0 = the context's own), collects kernel timings and checks the
 * device-side error flag of the preceding apd_align_packed. */
apd_status apd_synchronize(apd_ctx *ctx, void *stream);

/* Replaces Alignment::new + construct_alignment + score for one ordered pair
 * (src/alignments.rs:106-125,165-180), with the warping path traced on the
 * device (the reference keeps `sparse` for this but never walks it; the path
 * definition is SURVEY.md Appendix A.8).  path_ij receives up to path_cap
 * (i, j) pairs, 1-based, end-to-start; *path_len the full length.  path_ij may
 * be NULL. */
apd_status apd_align_pair(apd_ctx *ctx, const apd_params *p, uint32_t i, uint32_t j,
                          float *score, uint32_t *path_ij, uint64_t path_cap,
                          uint64_t *path_len);

/* Batch form: pairs_ij = n_pairs ordered (i, j) pairs; scores[k]; pair k's path at
 * paths_ij + k * path_cap * 2 (may be NULL); path_lens[k] (may be NULL). */
apd_status apd_align_pairs(apd_ctx *ctx, const apd_params *p, const uint32_t *pairs_ij,
                           uint64_t n_pairs, float *scores, uint32_t *paths_ij,
                           uint64_t path_cap, uint64_t *path_lens);

/* Same with the caller's own AlignmentParams.warping_band (src/alignments.rs:79) in
 * place of the percentage-derived one: what Alignment::construct_alignment receives
 * when it is called outside align_all (p->warping_band_percentage is ignored). */
apd_status apd_align_pairs_band(apd_ctx *ctx, const apd_params *p, uint64_t warping_band,
                                const uint32_t *pairs_ij, uint64_t n_pairs, float *scores,
                                uint32_t *paths_ij, uint64_t path_cap, uint64_t *path_lens);

/* ---- threshold for the handoff to clustering ---------------------------------- */

/* Replaces numerics::percentile (src/numerics.rs:125-133) as AgglomerativeClustering::clustering
 * calls it on the whole matrix (src/clustering.rs:101): the element at index
 * (len as f32 * perc) as usize of the ascending non-NaN entries -- NaNs are dropped but the
 * index comes from the unfiltered length, +INF and the diagonal zeros take part.  An index
 * past the filtered length is the reference's out-of-bounds panic: APD_ERR_INVALID.
 * apd_percentile_matrix works on the device copy of the matrix the last apd_align_all left
 * behind; apd_percentile_device on any 16-byte-aligned device buffer (e.g. the matrix
 * apd_scatter_packed wrote), enqueued on `stream` and synchronised before returning. */
apd_status apd_percentile_matrix(apd_ctx *ctx, float perc, float *out);
apd_status apd_percentile_device(apd_ctx *ctx, const float *d_x, uint64_t len, float perc,
                                 void *stream, float *out);

/* ---- clustering on the host (the consumer of the matrix) ------------------------ */

/* src/clustering.rs:18-25 (ClusteringOperation) */
typedef struct {
    uint32_t merge_i, merge_j, into;
    float distance;
    uint32_t operation; /* src/clustering.rs:7-13: 0 Sequence2Sequence, 1 Sequence2Cluster,
                           2 Cluster2Sequence, 3 Cluster2Cluster */
    uint32_t tie;       /* 1 if another unordered root pair had exactly the same linkage (the
                           reference's choice then depends on HashSet iteration order) */
} apd_merge;

/* A result-identical fast form of AgglomerativeClustering::clustering (src/clustering.rs:81-110
 * with merge/linkage 153-209): same f32 linkage values (x-major sums recomputed only for the
 * merged cluster), strict-`<` argmin over roots in ascending id, the reference's stop rule
 * (`while n_clusters > 1 && distance < threshold`, so the merge that reaches the threshold is
 * still applied).  HOST code and host buffers: dist_nxn is the n*n row-major matrix, ops holds
 * n-1 entries, assignment (n entries, may be NULL) receives each instance's final root.
 * threshold_in == NULL: the threshold is numerics::percentile(dist, perc) computed here;
 * otherwise *threshold_in is used (e.g. from apd_percentile_matrix).  Returns APD_ERR_INVALID
 * where the reference panics (percentile index out of bounds). */
apd_status apd_upgma(const float *dist_nxn, uint32_t n, float perc, const float *threshold_in,
                     apd_merge *ops, uint32_t *n_ops, float *threshold_out,
                     uint32_t *assignment_out);

/* ---- persistence (host code) ---------------------------------------------------- */

/* The matrix the reference keeps only in an Arc<Mutex<Vec<f32>>> for the length of learn()
 * (src/main.rs:194-195), on disk: <stem>.apdm = n*n little-endian f32, row-major, byte for
 * byte the Vec<f32> handed to clustering() (src/clustering.rs:81-85); <stem>.apdm.json = a
 * small header {"format": "apd-matrix-1", "n", "dtype": "<f4", "params", "sha256"}.
 * params_json: a JSON object to store with it (NULL = {}).  apd_load_matrix with
 * out_nxn == NULL only reports n; verify != 0 checks the payload digest.
 * Paths (README.md:67 "alignment path information"): <stem>.apdp.json, cells 1-based,
 * end to start, in the layout apd_align_pairs fills (pair k at paths_ij + k*path_cap*2).
 * apd_load_paths stores at most cap_pairs entries / path_cap cells each and reports the
 * full counts in *n_pairs / path_lens.  Errors: apd_last_error(NULL).  Same files as
 * audio_pattern_discovery_b200/matrix_io.py reads and writes. */
apd_status apd_save_matrix(const char *stem, const float *dist_nxn, uint32_t n, const char *params_json);
apd_status apd_load_matrix(const char *stem, float *out_nxn, uint64_t cap_floats, uint32_t *n_out, int verify);
apd_status apd_save_paths(const char *stem, const uint32_t *pairs_ij, uint64_t n_pairs, const float *scores,
                          const uint32_t *paths_ij, uint64_t path_cap, const uint64_t *path_lens);
apd_status apd_load_paths(const char *stem, uint32_t *pairs_ij, float *scores, uint64_t *path_lens,
                          uint64_t cap_pairs, uint32_t *paths_ij, uint64_t path_cap, uint64_t *n_pairs);

/* ---- introspection --------------------------------------------------------- */
apd_status apd_get_stats(apd_ctx *ctx, apd_stats *out);
/* JSON array describing the launch classes of the last DTW enqueue on (the first device of)
 * this context: where each class keeps its boundary ring (tmem / smem / global), ring height
 * in tiles, units, grid and residency.  Valid until the next call on this thread. */
const char *apd_last_launch_plan(apd_ctx *ctx);

#ifdef __cplusplus
}
#endif
#endif /* APD_H */
