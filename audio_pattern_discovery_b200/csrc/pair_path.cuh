// pair_path.cuh -- K2: one ordered pair's banded weighted DTW with the branch taken
// at every cell recorded on the device and the warping path traced back on the
// device (BASELINE.json north_star (e): "backtracking and alignment paths run
// on-device only for pairs the reporting layer requests").
//
// Replaces Alignment::{new, construct_alignment, score} (src/alignments.rs:106-180)
// for single pairs; the reference keeps the `sparse` map that would allow a
// trace-back but never walks it (README.md:67 only promises it), so the path is
// defined by SURVEY.md Appendix A.8: start at (n-1, m-1), step to the predecessor
// the forward rule selected, stop after the cell whose predecessor is (0,0).
//
// Forward kernel (pair_wave_kernel): ONE WARP PER PAIR, an anti-diagonal wavefront of 4x4
// register tiles.  The DP rows are cut into slabs of 128 rows; lane l of the warp owns rows
// 4l .. 4l+3 of the slab (its 4 x frames stay in registers for the whole slab) and at step s
// computes the tile of column block J = s - l, so the 32 tiles of a step lie on one
// anti-diagonal: the row above a lane's tile is the bottom row of its upper neighbour's tile
// of the previous step (4 warp shuffles), the column to the left is its own previous tile
// (registers).  y frames are staged through a shared-memory ring (cp.async, 8 steps ahead,
// conflict-free padded tile stride), the row that links two slabs through a small L2-resident
// buffer.  The branch taken at each of a tile's 16 cells is packed into one 32-bit word
// (2 bits per cell) and stored step-major, lane-minor, so a warp step writes one coalesced
// 128-byte line: 4.2 MB of scratch for a 4096 x 4096 pair.
// Trace-back kernel (pair_trace_kernel): one thread per pair walks the direction words from
// (n-1, m-1); all requested pairs are walked concurrently.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <string>
#include <vector>

#include "dtw_core.h"
#include "host_plan.h"

namespace apd {

enum { PW_SLAB_ROWS = 128, PW_YRING = 64, PW_YLOOK = 8, PW_WARPS = 4 };

struct PairJob {
    uint32_t xs, ys;       // sorted positions of x = data[i], y = data[j]
    uint32_t out;          // index of the request this job answers (jobs are run most expensive first)
    int32_t w;             // window (src/alignments.rs:173)
    uint32_t steps_max;    // direction words are stored at [(slab * steps_max + step) * 32 + lane]
    uint32_t pad;
    uint64_t dir_off;      // first direction word of this pair in the scratch buffer (uint32 units)
    uint64_t row_off;      // first float of this pair's slab-link row in the scratch buffer (float units)
};

// Column blocks [Jlo, Jhi] (4 columns each, column j = 4J + c + 1) that slab k (rows
// 128k + 1 .. 128k + 128) touches inside the band j - i in [-w, w-1] and the matrix
// (rows 1..np, columns 1..mp).  Empty (Jhi < Jlo) when the slab lies below the band.
APD_HD void pw_slab_range(int k, int np, int mp, int w, int& Jlo, int& Jhi)
{
    const long long ia = 128ll * k + 1;
    long long ib = ia + 127;
    if (ib > np) ib = np;
    long long jlo = ia - w, jhi = ib + w - 1;
    if (jlo < 1) jlo = 1;
    if (jhi > mp) jhi = mp;
    if (jhi < jlo || ib < ia) { Jlo = 0; Jhi = -1; return; }
    Jlo = (int)((jlo - 1) >> 2);
    Jhi = (int)((jhi - 1) >> 2);
}

// Context-owned device scratch of the trace-back path (grow-only, reused across calls).
struct PathScratch {
    void* d_buf = nullptr; size_t cap = 0;        // direction words + slab-link rows
    void* d_jobs = nullptr; size_t jobs_cap = 0;
    float* d_scores = nullptr; unsigned long long* d_lens = nullptr; size_t res_cap = 0;
    uint32_t* d_paths = nullptr; size_t paths_cap = 0;
    unsigned int* d_counter = nullptr;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    void release();
};

// Host runner (apd_api.cu): orders the requests by cost, chunks them so the direction scratch
// fits, runs the two kernels, copies scores / paths back.  Returns cudaSuccess and an empty
// `err` on success; a non-empty `err` with cudaSuccess is an argument problem.
// band_override >= 0 replaces the pct-derived band (AlignmentParams.warping_band,
// src/alignments.rs:79) for callers of construct_alignment that pass their own.
// *ms (may be NULL) receives the device time of the kernels (CUDA events).
cudaError_t pair_paths_run(const Arena& arena, const float* d_arena, const uint32_t* d_off,
                           const uint32_t* d_len, const uint32_t* pairs_ij, uint64_t n_pairs, float pct,
                           long long band_override, float ins, float del, float mat, bool strict, float* scores,
                           uint32_t* paths_ij, uint64_t path_cap, uint64_t* path_lens, int sm_count, cudaStream_t stream,
                           PathScratch& scratch, float* ms, std::string& err);

}  // namespace apd
