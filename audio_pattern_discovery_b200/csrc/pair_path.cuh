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
// One CTA per requested pair; anti-diagonal wavefront over t = i + j with three
// rolling diagonals in shared memory (indexed by i) and one direction byte per
// cell stored diagonal-major in global scratch so a diagonal's writes coalesce.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <string>
#include <vector>

#include "dtw_core.h"
#include "host_plan.h"

namespace apd {

struct PairJob {
    uint32_t xs, ys;      // sorted positions of x = data[i], y = data[j]
    uint64_t dir_off;     // byte offset of this pair's direction scratch
    uint32_t stride;      // bytes per anti-diagonal in the scratch
};


// Host runner (apd_api.cu): chunks the request so the direction scratch fits, runs
// the kernel, copies scores / paths back.  Returns cudaSuccess and an empty `err`
// on success; a non-empty `err` with cudaSuccess is an argument problem.
// band_override >= 0 replaces the pct-derived band (AlignmentParams.warping_band,
// src/alignments.rs:79) for callers of construct_alignment that pass their own.
cudaError_t pair_paths_run(const Arena& arena, const float* d_arena, const uint32_t* d_off,
                           const uint32_t* d_len, const uint32_t* pairs_ij, uint64_t n_pairs, float pct,
                           long long band_override, float ins, float del, float mat, bool strict, float* scores,
                           uint32_t* paths_ij, uint64_t path_cap, uint64_t* path_lens, int sm_count, cudaStream_t stream,
                           float* ms, std::string& err);

}  // namespace apd
