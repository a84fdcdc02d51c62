// pair_path.cu -- host runner of the on-device DTW trace-back kernel (pair_path.cuh).
#include "pair_path.cuh"

#include <algorithm>

namespace apd {

namespace {

// Cells of anti-diagonal t (i + j = t) inside rows 1..n', columns 1..m' and the band
// j - i in [-w, w-1] (src/alignments.rs:174-175): i in [lo, hi].
__device__ __forceinline__ void diag_range(int t, int np, int mp, int w, int& lo, int& hi)
{
    // j = t - i;  j - i = t - 2i in [-w, w-1]  <=>  i in [ceil((t-w+1)/2), floor((t+w)/2)]
    int a = t - w + 1;
    lo = (a >= 0) ? (a + 1) >> 1 : -((-a) >> 1);
    hi = (t + w) >> 1;
    if (lo < 1) lo = 1;
    if (lo < t - mp) lo = t - mp;
    if (hi > np) hi = np;
    if (hi > t - 1) hi = t - 1;
}

__device__ __forceinline__ float pair_distance(const float* __restrict__ x, const float* __restrict__ y,
                                               int dpad, bool strict)
{
    if (strict) {
        float acc = 0.0f;
        for (int k = 0; k < dpad; k++) {
            float d = __fadd_rn(x[k], -y[k]);
            acc = __fadd_rn(acc, __fmul_rn(d, d));
        }
        return __fsqrt_rn(acc);
    }
    // Same association as the FAST A stage of row_step(): even / odd partial sums with FMAs.
    float a0 = 0.0f, a1 = 0.0f;
    for (int k = 0; k < dpad; k += 2) {
        float d0 = x[k] - y[k], d1 = x[k + 1] - y[k + 1];
        a0 = (k == 0) ? __fmul_rn(d0, d0) : __fmaf_rn(d0, d0, a0);
        a1 = (k == 0) ? __fmul_rn(d1, d1) : __fmaf_rn(d1, d1, a1);
    }
    return sqrt_fast(a0 + a1);
}

__global__ void __launch_bounds__(256) pair_path_kernel(
    const float* __restrict__ arena, const uint32_t* __restrict__ off, const uint32_t* __restrict__ len,
    const PairJob* __restrict__ jobs, int dpad, float pct, long long band_override, float pins, float pdel,
    float pmat, int strict,
    uint8_t* __restrict__ dirs, float* __restrict__ scores, uint32_t* __restrict__ paths, uint64_t path_cap,
    unsigned long long* __restrict__ path_lens)
{
    extern __shared__ float diag_smem[];
    const PairJob job = jobs[blockIdx.x];
    const int n = (int)len[job.xs], m = (int)len[job.ys];
    const int np = n - 1, mp = m - 1;
    if (threadIdx.x == 0) path_lens[blockIdx.x] = 0;
    if (n < 1 || m < 1 || np == 0 || mp == 0) {
        // src/alignments.rs:116-125: the score cell (n-1, m-1) is the seed (0,0) when
        // n == m == 1 (0 / 2 = 0) and absent (+INF) when only one side has length 1 or
        // a side is empty.  No path in either case.
        if (threadIdx.x == 0) scores[blockIdx.x] = (n == 1 && m == 1) ? 0.0f : APD_INF;
        return;
    }
    const int w = (band_override >= 0) ? window_of_band(band_override, n, m) : window_of(pct, n, m);
    const float* x0 = arena + (size_t)off[job.xs] * dpad;
    const float* y0 = arena + (size_t)off[job.ys] * dpad;
    uint8_t* dir = dirs + job.dir_off;
    const int L = np + 2;  // diagonals are indexed by i in 0..np
    float* d0 = diag_smem;          // t-2
    float* d1 = diag_smem + L;      // t-1
    float* d2 = diag_smem + 2 * L;  // t
    for (int i = threadIdx.x; i < L; i += blockDim.x) { d0[i] = APD_INF; d1[i] = APD_INF; d2[i] = APD_INF; }
    __syncthreads();
    if (threadIdx.x == 0) d0[0] = 0.0f;  // t = 0: the seed (0,0) = 0 (src/alignments.rs:107-111)
    __syncthreads();
    // t = 1 holds only boundary cells (0,1) and (1,0): absent -> d1 stays +INF.
    for (int t = 2; t <= np + mp; t++) {
        int lo, hi;
        diag_range(t, np, mp, w, lo, hi);
        // Entries of d2 outside [lo, hi] must read as +INF two diagonals later; the band
        // moves by at most one row per diagonal, so clearing a margin of 2 suffices.
        for (int i = lo - 2 + (int)threadIdx.x; i <= hi + 2; i += blockDim.x)
            if (i >= 0 && i < L && (i < lo || i > hi)) d2[i] = APD_INF;
        for (int i = lo + (int)threadIdx.x; i <= hi; i += blockDim.x) {
            const int j = t - i;
            const float dist = pair_distance(x0 + (size_t)(i - 1) * dpad, y0 + (size_t)(j - 1) * dpad, dpad, strict != 0);
            const float M = d0[i - 1];  // (i-1, j-1)
            const float I = d1[i - 1];  // (i-1, j)   insertion
            const float E = d1[i];      // (i, j-1)   deletion
            // src/alignments.rs:153-159
            int b = 0;
            if (E < M && E < I) b = 2;
            else if (I < M && I < E) b = 1;
            const float base = (b == 2) ? E : (b == 1 ? I : M);
            const float pen = (b == 2) ? pdel : (b == 1 ? pins : pmat);
            d2[i] = __fadd_rn(base, __fmul_rn(pen, dist));
            dir[(size_t)t * job.stride + (i - lo)] = (uint8_t)b;
        }
        __syncthreads();
        float* tmp = d0; d0 = d1; d1 = d2; d2 = tmp;
    }
    // After the rotation d1 holds diagonal np+mp, whose cell i = np is (n', m').
    if (threadIdx.x == 0) {
        const float acc = d1[np];
        scores[blockIdx.x] = finish_score(acc, n, m);
        unsigned long long plen = 0;
        int i = np, j = mp;
        uint32_t* out = paths ? paths + (size_t)blockIdx.x * path_cap * 2 : nullptr;
        while (i >= 1 && j >= 1) {
            if (out && plen < path_cap) { out[2 * plen] = (uint32_t)i; out[2 * plen + 1] = (uint32_t)j; }
            plen++;
            int lo, hi;
            const int t = i + j;
            diag_range(t, np, mp, w, lo, hi);
            const int b = dir[(size_t)t * job.stride + (i - lo)];
            if (b == 2) j -= 1;
            else if (b == 1) i -= 1;
            else { i -= 1; j -= 1; }
        }
        path_lens[blockIdx.x] = plen;
    }
}


struct DevBuf {
    void* p = nullptr;
    ~DevBuf() { if (p) cudaFree(p); }
    cudaError_t alloc(size_t bytes) { return cudaMalloc(&p, std::max<size_t>(bytes, 1)); }
};
}  // namespace

cudaError_t pair_paths_run(const Arena& ar, const float* d_arena, const uint32_t* d_off,
                           const uint32_t* d_len, const uint32_t* pairs_ij, uint64_t n_pairs, float pct,
                           long long band_override, float ins, float del, float mat, bool strict, float* scores,
                           uint32_t* paths_ij, uint64_t path_cap, uint64_t* path_lens, int sm_count, cudaStream_t stream,
                           float* ms, std::string& err)
{
    err.clear();
    if (paths_ij && path_cap == 0) paths_ij = nullptr;
    std::vector<uint32_t> inv(ar.n);
    for (uint32_t s = 0; s < ar.n; s++) inv[ar.perm[s]] = s;

    int dev = 0, smem_optin = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return e;
    e = cudaDeviceGetAttribute(&smem_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev);
    if (e != cudaSuccess) return e;
    size_t free_b = 0, total_b = 0;
    e = cudaMemGetInfo(&free_b, &total_b);
    if (e != cudaSuccess) return e;
    const uint64_t budget = std::max<uint64_t>(std::min<uint64_t>(free_b / 2, 8ull << 30), 64ull << 20);

    uint64_t done = 0;
    while (done < n_pairs) {
        // Greedy chunk: as many pairs as fit the scratch budget (at least one).
        std::vector<PairJob> jobs;
        uint64_t dir_bytes = 0;
        size_t smem_max = 0;
        uint64_t k = done;
        for (; k < n_pairs; k++) {
            PairJob j;
            j.xs = inv[pairs_ij[2 * k]];
            j.ys = inv[pairs_ij[2 * k + 1]];
            const int n = (int)ar.len[j.xs], m = (int)ar.len[j.ys];
            const int np = n - 1, mp = m - 1;
            uint64_t bytes = 0;
            j.stride = 1;
            if (np >= 1 && mp >= 1) {
                const int w = (band_override >= 0) ? window_of_band(band_override, n, m) : window_of(pct, n, m);
                j.stride = (uint32_t)std::min(std::min(np, mp), w + 1) + 1;
                bytes = (uint64_t)(np + mp + 1) * j.stride;
                size_t smem = (size_t)3 * (np + 2) * sizeof(float);
                if (smem > (size_t)smem_optin) {
                    err = "sequence too long for the on-device trace-back kernel";
                    return cudaSuccess;
                }
                smem_max = std::max(smem_max, smem);
            }
            bytes = (bytes + 15) & ~15ull;
            if (!jobs.empty() && (dir_bytes + bytes > budget || jobs.size() >= 65535)) break;
            j.dir_off = dir_bytes;
            dir_bytes += bytes;
            jobs.push_back(j);
        }
        const size_t cnt = jobs.size();
        DevBuf d_jobs, d_dirs, d_scores, d_paths, d_lens;
        if ((e = d_jobs.alloc(cnt * sizeof(PairJob))) != cudaSuccess) return e;
        if ((e = d_dirs.alloc(dir_bytes)) != cudaSuccess) return e;
        if ((e = d_scores.alloc(cnt * sizeof(float))) != cudaSuccess) return e;
        if ((e = d_lens.alloc(cnt * sizeof(unsigned long long))) != cudaSuccess) return e;
        if (paths_ij && (e = d_paths.alloc(cnt * path_cap * 2 * sizeof(uint32_t))) != cudaSuccess) return e;
        if ((e = cudaMemcpyAsync(d_jobs.p, jobs.data(), cnt * sizeof(PairJob), cudaMemcpyHostToDevice, stream)) != cudaSuccess) return e;
        if ((e = cudaFuncSetAttribute(pair_path_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)std::max<size_t>(smem_max, 16))) != cudaSuccess) return e;
        pair_path_kernel<<<(unsigned)cnt, 256, std::max<size_t>(smem_max, 16), stream>>>(
            d_arena, d_off, d_len, (const PairJob*)d_jobs.p, (int)ar.dpad, pct, band_override, ins, del, mat, strict ? 1 : 0,
            (uint8_t*)d_dirs.p, (float*)d_scores.p, paths_ij ? (uint32_t*)d_paths.p : nullptr, path_cap,
            (unsigned long long*)d_lens.p);
        if ((e = cudaGetLastError()) != cudaSuccess) return e;
        if ((e = cudaMemcpyAsync(scores + done, d_scores.p, cnt * sizeof(float), cudaMemcpyDeviceToHost, stream)) != cudaSuccess) return e;
        std::vector<unsigned long long> lens_h(cnt);
        if ((e = cudaMemcpyAsync(lens_h.data(), d_lens.p, cnt * sizeof(unsigned long long), cudaMemcpyDeviceToHost, stream)) != cudaSuccess) return e;
        if (paths_ij)
            if ((e = cudaMemcpyAsync(paths_ij + done * path_cap * 2, d_paths.p, cnt * path_cap * 2 * sizeof(uint32_t), cudaMemcpyDeviceToHost, stream)) != cudaSuccess) return e;
        if ((e = cudaStreamSynchronize(stream)) != cudaSuccess) return e;
        if (path_lens)
            for (size_t q = 0; q < cnt; q++) path_lens[done + q] = lens_h[q];
        done += cnt;
    }
    return cudaSuccess;
}

}  // namespace apd
