// percentile.cuh -- see percentile.cu.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <string.h>

#include <string>

namespace apd {

// (len as f32 * perc) as usize, src/numerics.rs:126,132
uint64_t percentile_index(uint64_t len, float perc);

// Exact k-th smallest non-NaN entry of d_x[0..len) with k = percentile_index(len, perc).
// d_hist: 256 device counters.  A non-empty `err` with cudaSuccess is the reference's panic.
cudaError_t percentile_select(const float* d_x, uint64_t len, float perc, unsigned long long* d_hist, int sm_count,
                              cudaStream_t stream, float* out, uint64_t* n_valid, float* ms, std::string& err);

}  // namespace apd
