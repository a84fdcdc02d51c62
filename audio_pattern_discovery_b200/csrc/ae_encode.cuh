// ae_encode.cuh -- see ae_encode.cu.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace apd {

// raw: dense cepstra in sorted order (sequence s at raw + src_off[s], len[s] x n_bins floats);
// arena frame (off[s] + t) receives predict(frame t) in its first n_latent floats (the arena
// was zeroed, so the pads stay 0).  w: n_bins x n_latent row-major, b: n_latent.
cudaError_t ae_encode_launch(const float* d_raw, const uint64_t* d_src_off, const uint32_t* d_off, const uint32_t* d_len,
                             uint32_t n, uint32_t n_bins, uint32_t n_latent, uint32_t dpad, const float* d_w,
                             const float* d_b, float* d_arena, int sm_count, cudaStream_t stream);

}  // namespace apd
