// percentile.cu -- the threshold step of the handoff to clustering
// (src/clustering.rs:101 -> src/numerics.rs:125-133) as an exact order statistic on the
// device: an 8-bit-per-pass radix select over the n*n distances, with the reference's
// quirks kept: NaN entries are dropped but the index is computed from the UNFILTERED
// length ((len as f32 * perc) as usize), entries compare like partial_cmp (so +INF sorts
// last and the n diagonal zeros take part), and an index past the filtered length is the
// reference's out-of-bounds panic (reported as an error, never clamped).
//
// HBM-bound: 4 passes, each reads the matrix once (vectorised, grid = SMs x 8) and builds
// a 256-bin histogram in shared memory.  Algorithmic bytes = 4 * 4 * len.
#include "percentile.cuh"

#include <algorithm>

namespace apd {

namespace {

__device__ __forceinline__ uint32_t order_key(float v)
{
    // -0.0 and +0.0 compare equal in the reference (partial_cmp, src/numerics.rs:129): both get the
    // key of +0.0, so a selected zero is always reported as +0.0 (a DTW matrix never holds -0.0:
    // its entries are sums of square roots divided by a positive count)
    const uint32_t b = (__float_as_uint(v) == 0x80000000u) ? 0u : __float_as_uint(v);
    return (b & 0x80000000u) ? ~b : (b | 0x80000000u);  // ascending uint order == ascending float order
}

// Distances crowd into a few exponent values, so most consecutive votes of a thread hit the
// same bin: votes are run-length merged in registers and only runs reach the shared atomics.
struct BinRun {
    uint32_t bin, count;
};

__device__ __forceinline__ void vote_bin(uint32_t* hist, BinRun& run, float v, uint32_t prefix, uint32_t prefix_mask,
                                         int shift)
{
    if (v != v) return;  // NaN: filtered out like src/numerics.rs:128
    const uint32_t k = order_key(v);
    if ((k & prefix_mask) != prefix) return;
    const uint32_t b = (k >> shift) & 0xffu;
    if (b == run.bin) { run.count++; return; }
    if (run.count) atomicAdd(&hist[run.bin], run.count);
    run.bin = b;
    run.count = 1;
}

__global__ void __launch_bounds__(512) select_hist_kernel(const float* __restrict__ x, uint64_t len, uint32_t prefix,
                                                          uint32_t prefix_mask, int shift,
                                                          unsigned long long* __restrict__ hist_out)
{
    __shared__ uint32_t hist[256];
    for (int k = threadIdx.x; k < 256; k += blockDim.x) hist[k] = 0;
    __syncthreads();
    const uint64_t tid = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x, nth = (uint64_t)gridDim.x * blockDim.x;
    // the buffer comes from cudaMalloc (256-byte aligned): whole float4s, then the tail
    const uint64_t n4 = len / 4;
    const float4* x4 = reinterpret_cast<const float4*>(x);
    BinRun run;
    run.bin = 0; run.count = 0;
    for (uint64_t i = tid; i < n4; i += nth) {
        const float4 v = __ldg(x4 + i);
        vote_bin(hist, run, v.x, prefix, prefix_mask, shift);
        vote_bin(hist, run, v.y, prefix, prefix_mask, shift);
        vote_bin(hist, run, v.z, prefix, prefix_mask, shift);
        vote_bin(hist, run, v.w, prefix, prefix_mask, shift);
    }
    for (uint64_t i = 4 * n4 + tid; i < len; i += nth) vote_bin(hist, run, x[i], prefix, prefix_mask, shift);
    if (run.count) atomicAdd(&hist[run.bin], run.count);
    __syncthreads();
    for (int k = threadIdx.x; k < 256; k += blockDim.x)
        if (hist[k]) atomicAdd(&hist_out[k], (unsigned long long)hist[k]);
}

}  // namespace

uint64_t percentile_index(uint64_t len, float perc)
{
    // Rust: `x.len() as f32 * perc` then `as usize` (truncation, saturation, NaN -> 0)
    const float n = (float)len * perc;
    if (!(n == n) || n <= 0.0f) return 0;
    if (n >= 18446744073709551616.0f) return UINT64_MAX;
    return (uint64_t)n;
}

cudaError_t percentile_select(const float* d_x, uint64_t len, float perc, unsigned long long* d_hist, int sm_count,
                              cudaStream_t stream, float* out, uint64_t* n_valid, float* ms, std::string& err)
{
    err.clear();
    if (((uintptr_t)d_x & 15u) != 0) { err = "percentile input must be 16-byte aligned"; return cudaSuccess; }
    const uint64_t index = percentile_index(len, perc);
    cudaEvent_t e0, e1;
    cudaError_t e;
    if ((e = cudaEventCreate(&e0)) != cudaSuccess) return e;
    if ((e = cudaEventCreate(&e1)) != cudaSuccess) { cudaEventDestroy(e0); return e; }
    uint32_t prefix = 0, mask = 0;
    uint64_t k = index;
    unsigned long long h[256];
    const int grid = (int)std::max<uint64_t>(1, std::min<uint64_t>((uint64_t)sm_count * 8, (len / 4 + 511) / 512));
    cudaEventRecord(e0, stream);
    for (int pass = 0; pass < 4 && e == cudaSuccess; pass++) {
        const int shift = 24 - 8 * pass;
        if ((e = cudaMemsetAsync(d_hist, 0, 256 * sizeof(unsigned long long), stream)) != cudaSuccess) break;
        select_hist_kernel<<<grid, 512, 0, stream>>>(d_x, len, prefix, mask, shift, d_hist);
        if ((e = cudaGetLastError()) != cudaSuccess) break;
        if ((e = cudaMemcpyAsync(h, d_hist, sizeof(h), cudaMemcpyDeviceToHost, stream)) != cudaSuccess) break;
        if ((e = cudaStreamSynchronize(stream)) != cudaSuccess) break;
        if (pass == 0) {
            uint64_t valid = 0;
            for (int b = 0; b < 256; b++) valid += h[b];
            if (n_valid) *n_valid = valid;
            if (index >= valid) {
                err = "percentile index out of bounds: the reference panics here (src/numerics.rs:132)";
                break;
            }
        }
        int b = 0;
        while (b < 255 && k >= h[b]) { k -= h[b]; b++; }
        prefix |= (uint32_t)b << shift;
        mask |= 0xffu << shift;
    }
    cudaEventRecord(e1, stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(stream);
    if (e == cudaSuccess && ms) cudaEventElapsedTime(ms, e0, e1);
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    if (e != cudaSuccess || !err.empty()) return e;
    const uint32_t bits = (prefix & 0x80000000u) ? (prefix & 0x7fffffffu) : ~prefix;
    float v;
    memcpy(&v, &bits, 4);
    *out = v;
    return cudaSuccess;
}

}  // namespace apd
