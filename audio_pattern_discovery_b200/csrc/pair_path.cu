// pair_path.cu -- K2, the on-device DTW trace-back for requested pairs: kernels and host runner
// (design notes in pair_path.cuh).
#include "pair_path.cuh"

#include "apd_internal.h"

#include <cuda_bf16.h>

#include <algorithm>
#include <cstdlib>
#include <numeric>

namespace apd {

namespace {

struct WaveArgs {
    const float* arena;
    const uint32_t* off;
    const uint32_t* len;
    const PairJob* jobs;
    const uint32_t* order;   // job indices, most expensive first (the warps' fetch order)
    uint32_t n_jobs;
    unsigned int* counter;
    float pins, pdel, pmat;
    uint32_t* scratch;       // direction words (uint32) and slab-link rows (float) live in one buffer
    float* scores;           // indexed by job
};

// Arithmetic of the local frame distance (the product ships DIST_STRICT and DIST_FAST; the
// others exist only to MEASURE what a tensor-core formulation of the distance would do to the
// final score, north_star (d): they emulate, on the CUDA cores, the numerics of
// |x|^2 + |y|^2 - 2 x.y with the dot product taken from split-precision tensor-core MMAs --
// in their most favourable form: exact products, f32 FMA accumulation).  Selected with
// APD_EXPERIMENT_DIST=<n> for FAST-mode apd_align_pairs calls on 20-wide frames
// (tools/tensor_core_score_error.py); never used by apd_align_all.
enum {
    DIST_STRICT = 0,      // the reference's operation sequence (bit-exact)
    DIST_FAST = 1,        // difference form, FMA accumulation, approximate sqrt
    DIST_DOT_F32 = 2,     // dot form, every product and sum in f32 FMA (the best any dot form can do)
    DIST_DOT_3XTF32 = 3,  // dot form, x.y = xh.yh + xh.yl + xl.yh with tf32 pieces (3 MMAs, f32 accumulate)
    DIST_DOT_1XTF32 = 4,  // dot form, single-pass tf32
    DIST_DOT_3XTF32_CENTERED = 5,  // as 3, after subtracting the row sequence's mean frame from x and y
    DIST_DOT_BF16X3 = 6   // dot form, 3 bf16 pieces per operand, the 6 largest cross products
};

__device__ __forceinline__ float to_tf32(float a)
{
    uint32_t r;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(a));
    return __uint_as_float(r);
}
__device__ __forceinline__ float to_bf16(float a) { return __bfloat162float(__float2bfloat16_rn(a)); }

// src/numerics.rs:114-120 for one cell.  STRICT: packed subtract and square, sequential f32
// accumulation in dimension order, IEEE sqrt -- the same bits as K1 and the reference.  FAST:
// the association of K1's FAST A stage (even / odd partial sums with FMAs, approximate sqrt).
template <int DPAD, int MODE>
__device__ __forceinline__ float cell_distance(const F2 (&x)[DPAD / 2], const F2 (&y)[DPAD / 2])
{
    if (MODE == DIST_STRICT) {
        float acc = 0.0f;
#pragma unroll
        for (int k = 0; k < DPAD / 2; k++) {
            const F2 t = sub2_rn(x[k], y[k]);
            const F2 p = mul2_rn(t, t);
            acc = (k == 0) ? p.x : add_rn(acc, p.x);
            acc = add_rn(acc, p.y);
        }
        return sqrt_rn(acc);
    }
    if (MODE == DIST_FAST) {
        F2 acc2 = mk2(0.0f, 0.0f);
#pragma unroll
        for (int k = 0; k < DPAD / 2; k++) {
            const F2 t = sub2_rn(x[k], y[k]);
            acc2 = (k == 0) ? mul2_rn(t, t) : fma2_rn(t, t, acc2);
        }
        return sqrt_fast(acc2.x + acc2.y);
    }
    // ---- experiments: d^2 = |x|^2 + |y|^2 - 2 x.y ----
    float nx = 0.0f, ny = 0.0f, dot = 0.0f;
#pragma unroll
    for (int k = 0; k < DPAD; k++) {
        const float a = (k & 1) ? x[k / 2].y : x[k / 2].x, b = (k & 1) ? y[k / 2].y : y[k / 2].x;
        nx = fmaf(a, a, nx);   // the norms are per-frame constants computed once in f32 in any real kernel
        ny = fmaf(b, b, ny);
        if (MODE == DIST_DOT_F32) {
            dot = fmaf(a, b, dot);
        } else if (MODE == DIST_DOT_1XTF32) {
            dot = fmaf(to_tf32(a), to_tf32(b), dot);
        } else if (MODE == DIST_DOT_BF16X3) {
            const float a1 = to_bf16(a), a2 = to_bf16(a - a1), a3 = to_bf16(a - a1 - a2);
            const float b1 = to_bf16(b), b2 = to_bf16(b - b1), b3 = to_bf16(b - b1 - b2);
            dot = fmaf(a1, b1, dot); dot = fmaf(a1, b2, dot); dot = fmaf(a2, b1, dot);
            dot = fmaf(a2, b2, dot); dot = fmaf(a1, b3, dot); dot = fmaf(a3, b1, dot);
        } else {
            const float ah = to_tf32(a), al = to_tf32(a - ah), bh = to_tf32(b), bl = to_tf32(b - bh);
            dot = fmaf(ah, bh, dot); dot = fmaf(ah, bl, dot); dot = fmaf(al, bh, dot);
        }
    }
    const float d2 = fmaxf(fmaf(-2.0f, dot, nx + ny), 0.0f);
    return sqrt_fast(d2);
}

// The 4 distances of one tile column (4 x rows against one y frame).  STRICT / FAST: dimension-major with
// one accumulator per row, so four independent add chains are in flight (the per-cell operation order --
// and therefore every bit -- is that of cell_distance()).
template <int DPAD, int MODE>
__device__ __forceinline__ void column_distances(const F2 (&xr)[TILE][DPAD / 2], const F2 (&y)[DPAD / 2], float (&d)[TILE])
{
    if (MODE == DIST_STRICT) {
        float acc[TILE];
#pragma unroll
        for (int k = 0; k < DPAD / 2; k++) {
            F2 p[TILE];
#pragma unroll
            for (int r = 0; r < TILE; r++) { const F2 t = sub2_rn(xr[r][k], y[k]); p[r] = mul2_rn(t, t); }
#pragma unroll
            for (int r = 0; r < TILE; r++) acc[r] = (k == 0) ? p[r].x : add_rn(acc[r], p[r].x);
#pragma unroll
            for (int r = 0; r < TILE; r++) acc[r] = add_rn(acc[r], p[r].y);
        }
#pragma unroll
        for (int r = 0; r < TILE; r++) d[r] = sqrt_rn(acc[r]);
    } else if (MODE == DIST_FAST) {
        F2 acc2[TILE];
#pragma unroll
        for (int k = 0; k < DPAD / 2; k++) {
#pragma unroll
            for (int r = 0; r < TILE; r++) {
                const F2 t = sub2_rn(xr[r][k], y[k]);
                acc2[r] = (k == 0) ? mul2_rn(t, t) : fma2_rn(t, t, acc2[r]);
            }
        }
#pragma unroll
        for (int r = 0; r < TILE; r++) d[r] = sqrt_fast(acc2[r].x + acc2[r].y);
    } else {
#pragma unroll
        for (int r = 0; r < TILE; r++) d[r] = cell_distance<DPAD, MODE>(xr[r], y);
    }
}

template <int DPAD, int MODE>
__global__ void __launch_bounds__(32 * PW_WARPS) pair_wave_kernel(const WaveArgs a)
{
    extern __shared__ float4 smem4[];
    constexpr int TS = DPAD + 1;  // float4 per staged y tile: 4 frames (DPAD float4) + 1 pad -> conflict-free LDS.128
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    float4* ring = smem4 + warp * (PW_YRING * TS);
    const uint32_t ring_smem = (uint32_t)__cvta_generic_to_shared(ring);
    const float INF = APD_INF;

    for (;;) {
        unsigned int q = 0;
        if (lane == 0) q = atomicAdd(a.counter, 1u);
        q = __shfl_sync(0xffffffffu, q, 0);
        if (q >= a.n_jobs) break;
        const uint32_t ji = a.order[q];
        const PairJob jb = a.jobs[ji];
        const int n = (int)a.len[jb.xs], m = (int)a.len[jb.ys];
        const int np = n - 1, mp = m - 1, w = jb.w;      // the host only queues pairs with np >= 1 and mp >= 1
        const float4* xbase4 = reinterpret_cast<const float4*>(a.arena) + (size_t)a.off[jb.xs] * (DPAD / 4);
        const float4* ybase4 = reinterpret_cast<const float4*>(a.arena) + (size_t)a.off[jb.ys] * (DPAD / 4);
        uint32_t* dir = a.scratch + jb.dir_off;
        float* rb = reinterpret_cast<float*>(a.scratch) + jb.row_off;   // rb[j - 1]: last row of the previous slab at column j
        const int Jt = (mp + 3) >> 2;
        for (int j = lane; j < 4 * Jt; j += 32) rb[j] = INF;            // row 0 of the DP: absent cells
        __syncwarp();

        // experiment DIST_DOT_3XTF32_CENTERED: the row sequence's mean frame, subtracted from x and y alike
        // (x - y is unchanged, |x|^2 and |y|^2 shrink towards the size of the differences)
        F2 ctr[DPAD / 2];
#pragma unroll
        for (int v = 0; v < DPAD / 2; v++) ctr[v] = mk2(0.0f, 0.0f);
        if (MODE == DIST_DOT_3XTF32_CENTERED) {
            for (int t = lane; t < n; t += 32) {
                const float4* p = xbase4 + (size_t)t * (DPAD / 4);
#pragma unroll
                for (int v = 0; v < DPAD / 4; v++) {
                    const float4 f = __ldg(p + v);
                    ctr[2 * v].x += f.x; ctr[2 * v].y += f.y; ctr[2 * v + 1].x += f.z; ctr[2 * v + 1].y += f.w;
                }
            }
#pragma unroll
            for (int v = 0; v < DPAD / 2; v++) {
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) {
                    ctr[v].x += __shfl_xor_sync(0xffffffffu, ctr[v].x, o);
                    ctr[v].y += __shfl_xor_sync(0xffffffffu, ctr[v].y, o);
                }
                ctr[v].x /= (float)n; ctr[v].y /= (float)n;
            }
        }

        // the score cell (n', m') (src/alignments.rs:116-125)
        const int kstar = (np - 1) >> 7, lstar = ((np - 1) & 127) >> 2, rstar = (np - 1) & 3;
        const int Jstar = (mp - 1) >> 2, cstar = (mp - 1) & 3;
        float ans = INF;

        const int n_slabs = (np + PW_SLAB_ROWS - 1) / PW_SLAB_ROWS;
        for (int k = 0; k < n_slabs; k++) {
            int Jlo, Jhi;
            pw_slab_range(k, np, mp, w, Jlo, Jhi);
            if (Jhi < Jlo) continue;
            const int i0 = PW_SLAB_ROWS * k + 1 + 4 * lane;   // first DP row of this lane
            const bool has_rows = i0 <= np;
            // x frames of rows i0 .. i0+3 (frame i - 1), clamped to the last row used
            F2 xr[TILE][DPAD / 2];
#pragma unroll
            for (int r = 0; r < TILE; r++) {
                int i = i0 + r;
                if (i > np) i = np;
                const float4* p = xbase4 + (size_t)(i - 1) * (DPAD / 4);
#pragma unroll
                for (int v = 0; v < DPAD / 4; v++) {
                    const float4 f = __ldg(p + v);
                    xr[r][2 * v] = make_float2(f.x, f.y);
                    xr[r][2 * v + 1] = make_float2(f.z, f.w);
                    if (MODE == DIST_DOT_3XTF32_CENTERED) {
                        xr[r][2 * v] = sub2_rn(xr[r][2 * v], ctr[2 * v]);
                        xr[r][2 * v + 1] = sub2_rn(xr[r][2 * v + 1], ctr[2 * v + 1]);
                    }
                }
            }
            float left[TILE], bottom[TILE];
#pragma unroll
            for (int r = 0; r < TILE; r++) { left[r] = INF; bottom[r] = INF; }
            // cell above-left of the lane's first tile: out of band for every lane but the first, whose
            // upper neighbour row belongs to the previous slab (or is row 0 with the seed (0,0) = 0)
            float diag = INF;
            if (lane == 0) diag = (Jlo == 0) ? (k == 0 ? 0.0f : INF) : rb[4 * Jlo - 1];

            auto y_fetch = [&](int J) {
                if (J <= Jhi && lane < DPAD) {
                    const uint32_t dst = ring_smem + (uint32_t)(((J & (PW_YRING - 1)) * TS + lane) * 16);
                    const float4* src = ybase4 + (size_t)J * DPAD + lane;   // frames 4J .. 4J+3 are contiguous
                    asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src));
                }
                asm volatile("cp.async.commit_group;");
            };
            __syncwarp();   // every lane is done with the ring of the previous slab
#pragma unroll 1
            for (int t = 0; t < PW_YLOOK; t++) y_fetch(Jlo + t);

#pragma unroll 1
            for (int s = Jlo; s <= Jhi + 31; s++) {
                y_fetch(s + PW_YLOOK);
                const int J = s - lane;
                const bool act = has_rows && J >= Jlo && J <= Jhi;
                // row above the tile: the upper neighbour's bottom row of the previous step; lane 0: the slab link
                float top[TILE];
#pragma unroll
                for (int c = 0; c < TILE; c++) top[c] = __shfl_up_sync(0xffffffffu, bottom[c], 1);
                if (lane == 0) {
                    if (act) {
                        const float4 v = *reinterpret_cast<const float4*>(rb + 4 * J);
                        top[0] = v.x; top[1] = v.y; top[2] = v.z; top[3] = v.w;
                    } else {
#pragma unroll
                        for (int c = 0; c < TILE; c++) top[c] = INF;
                    }
                }
                asm volatile("cp.async.wait_group %0;" ::"n"((int)PW_YLOOK + 0) : "memory");
                __syncwarp();
                uint32_t word = 0;
                if (act) {
                    const float4* yt = ring + (J & (PW_YRING - 1)) * TS;
                    const int j0 = 4 * J + 1;
                    // cell (r, c) is a real in-band cell iff r <= rmax, c <= cmax and dlo <= c - r <= dhi
                    const int rmax = np - i0, cmax = mp - j0;
                    const int dlo = -w - (j0 - i0), dhi = (w - 1) - (j0 - i0);
                    const bool cap = (k == kstar) && (lane == lstar) && (J == Jstar);
                    float colprev[TILE];
#pragma unroll
                    for (int r = 0; r < TILE; r++) colprev[r] = left[r];
#pragma unroll
                    for (int c = 0; c < TILE; c++) {
                        F2 yv[DPAD / 2];
#pragma unroll
                        for (int v = 0; v < DPAD / 4; v++) {
                            const float4 f = yt[c * (DPAD / 4) + v];
                            yv[2 * v] = make_float2(f.x, f.y);
                            yv[2 * v + 1] = make_float2(f.z, f.w);
                            if (MODE == DIST_DOT_3XTF32_CENTERED) {
                                yv[2 * v] = sub2_rn(yv[2 * v], ctr[2 * v]);
                                yv[2 * v + 1] = sub2_rn(yv[2 * v + 1], ctr[2 * v + 1]);
                            }
                        }
                        float dcol[TILE];
                        column_distances<DPAD, MODE>(xr, yv, dcol);
                        float up = top[c];
                        float dg = (c == 0) ? diag : top[c - 1];
#pragma unroll
                        for (int r = 0; r < TILE; r++) {
                            const float d = dcol[r];
                            const float E = colprev[r];   // (i, j-1)   deletion
                            const float I = up;           // (i-1, j)   insertion
                            const float M = dg;           // (i-1, j-1) match
                            // src/alignments.rs:153-159
                            uint32_t b = 0;
                            if (E < M && E < I) b = 2;
                            else if (I < M && I < E) b = 1;
                            const float base = (b == 2) ? E : (b == 1 ? I : M);
                            const float pen = (b == 2) ? a.pdel : (b == 1 ? a.pins : a.pmat);
                            float v = add_rn(base, mul_rn(pen, d));
                            const bool ok = (r <= rmax) && (c <= cmax) && (c - r >= dlo) && (c - r <= dhi);
                            v = ok ? v : INF;             // what a missing map entry reads as
                            b = ok ? b : 0u;
                            word |= b << (2 * (4 * r + c));
                            if (cap && r == rstar && c == cstar) ans = v;
                            dg = colprev[r];
                            colprev[r] = v;
                            up = v;
                        }
                        bottom[c] = up;
                    }
#pragma unroll
                    for (int r = 0; r < TILE; r++) left[r] = colprev[r];
                    dir[((size_t)k * jb.steps_max + (size_t)(s - Jlo)) * 32 + lane] = word;
                    if (lane == 31) *reinterpret_cast<float4*>(rb + 4 * J) = make_float4(bottom[0], bottom[1], bottom[2], bottom[3]);
                } else {
#pragma unroll
                    for (int c = 0; c < TILE; c++) bottom[c] = INF;
                }
                diag = top[TILE - 1];
            }
            asm volatile("cp.async.wait_group 0;" ::: "memory");
            __syncwarp();   // lane 31's slab-link row is visible to lane 0 of the next slab
        }
        // one lane holds the score cell
        const float got = __shfl_sync(0xffffffffu, ans, lstar);
        if (lane == 0) a.scores[ji] = finish_score(got, n, m);
        __syncwarp();
    }
}

__global__ void __launch_bounds__(64) pair_trace_kernel(const PairJob* __restrict__ jobs, uint32_t n_jobs,
                                                        const uint32_t* __restrict__ len, const uint32_t* __restrict__ scratch,
                                                        uint32_t* __restrict__ paths, uint64_t path_cap,
                                                        unsigned long long* __restrict__ path_lens)
{
    const uint32_t ji = blockIdx.x * blockDim.x + threadIdx.x;
    if (ji >= n_jobs) return;
    const PairJob jb = jobs[ji];
    const int np = (int)len[jb.xs] - 1, mp = (int)len[jb.ys] - 1, w = jb.w;
    const uint32_t* dir = scratch + jb.dir_off;
    uint32_t* out = paths ? paths + (size_t)ji * path_cap * 2 : nullptr;
    unsigned long long plen = 0;
    int i = np, j = mp, kc = -1, Jlo = 0, Jhi = -1;
    while (i >= 1 && j >= 1) {
        if (out && plen < path_cap) { out[2 * plen] = (uint32_t)i; out[2 * plen + 1] = (uint32_t)j; }
        plen++;
        const int k = (i - 1) >> 7, l = ((i - 1) & 127) >> 2, r = (i - 1) & 3, J = (j - 1) >> 2, c = (j - 1) & 3;
        if (k != kc) { pw_slab_range(k, np, mp, w, Jlo, Jhi); kc = k; }
        uint32_t b = 0;   // a cell outside the band reads as a missing map entry everywhere: MATCH
        if (J >= Jlo && J <= Jhi) {
            const uint32_t word = dir[((size_t)k * jb.steps_max + (size_t)(J + l - Jlo)) * 32 + l];
            b = (word >> (2 * (4 * r + c))) & 3u;
        }
        if (b == 2) j -= 1;
        else if (b == 1) i -= 1;
        else { i -= 1; j -= 1; }
    }
    path_lens[ji] = plen;
}

template <int DPAD, int MODE>
cudaError_t launch_wave_mode(const WaveArgs& a, int grid, cudaStream_t stream)
{
    const size_t smem = (size_t)PW_WARPS * PW_YRING * (DPAD + 1) * sizeof(float4);
    cudaError_t e = cudaFuncSetAttribute(pair_wave_kernel<DPAD, MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    pair_wave_kernel<DPAD, MODE><<<grid, 32 * PW_WARPS, smem, stream>>>(a);
    return cudaGetLastError();
}

template <int DPAD>
cudaError_t launch_wave(const WaveArgs& a, bool strict, int grid, cudaStream_t stream)
{
    if (strict) return launch_wave_mode<DPAD, DIST_STRICT>(a, grid, stream);
    if (DPAD == 20) {   // measurement-only distance arithmetic (see the DIST_* enum)
        const char* ex = getenv("APD_EXPERIMENT_DIST");
        const int mode = ex ? atoi(ex) : 0;
        switch (mode) {
            case DIST_DOT_F32: return launch_wave_mode<20, DIST_DOT_F32>(a, grid, stream);
            case DIST_DOT_3XTF32: return launch_wave_mode<20, DIST_DOT_3XTF32>(a, grid, stream);
            case DIST_DOT_1XTF32: return launch_wave_mode<20, DIST_DOT_1XTF32>(a, grid, stream);
            case DIST_DOT_3XTF32_CENTERED: return launch_wave_mode<20, DIST_DOT_3XTF32_CENTERED>(a, grid, stream);
            case DIST_DOT_BF16X3: return launch_wave_mode<20, DIST_DOT_BF16X3>(a, grid, stream);
            default: break;
        }
    }
    return launch_wave_mode<DPAD, DIST_FAST>(a, grid, stream);
}

cudaError_t launch_wave_dpad(uint32_t dpad, const WaveArgs& a, bool strict, int grid, cudaStream_t stream)
{
    switch (dpad) {
        case 4: return launch_wave<4>(a, strict, grid, stream);
        case 8: return launch_wave<8>(a, strict, grid, stream);
        case 12: return launch_wave<12>(a, strict, grid, stream);
        case 16: return launch_wave<16>(a, strict, grid, stream);
        case 20: return launch_wave<20>(a, strict, grid, stream);
        case 24: return launch_wave<24>(a, strict, grid, stream);
        case 28: return launch_wave<28>(a, strict, grid, stream);
        case 32: return launch_wave<32>(a, strict, grid, stream);
        default: return cudaErrorInvalidValue;
    }
}

template <class T>
cudaError_t grow(T*& p, size_t& cap, size_t need_bytes)
{
    if (p && cap >= need_bytes) return cudaSuccess;
    if (p) { cudaFree(p); p = nullptr; cap = 0; }
    need_bytes = std::max<size_t>(need_bytes, 256);
    cudaError_t e = cudaMalloc((void**)&p, need_bytes);
    if (e == cudaSuccess) cap = need_bytes;
    return e;
}

}  // namespace

void PathScratch::release()
{
    if (d_buf) cudaFree(d_buf);
    if (d_jobs) cudaFree(d_jobs);
    if (d_scores) cudaFree(d_scores);
    if (d_lens) cudaFree(d_lens);
    if (d_paths) cudaFree(d_paths);
    if (d_counter) cudaFree(d_counter);
    if (ev0) cudaEventDestroy(ev0);
    if (ev1) cudaEventDestroy(ev1);
    *this = PathScratch();
}

cudaError_t pair_paths_run(const Arena& ar, const float* d_arena, const uint32_t* d_off,
                           const uint32_t* d_len, const uint32_t* pairs_ij, uint64_t n_pairs, float pct,
                           long long band_override, float ins, float del, float mat, bool strict, float* scores,
                           uint32_t* paths_ij, uint64_t path_cap, uint64_t* path_lens, int sm_count, cudaStream_t stream,
                           PathScratch& sc, float* ms, std::string& err)
{
    err.clear();
    if (ms) *ms = 0.f;
    if (paths_ij && path_cap == 0) paths_ij = nullptr;
    std::vector<uint32_t> inv(ar.n);
    for (uint32_t s = 0; s < ar.n; s++) inv[ar.perm[s]] = s;
    cudaError_t e;
    if (!sc.ev0) {
        if ((e = cudaEventCreate(&sc.ev0)) != cudaSuccess) return e;
        if ((e = cudaEventCreate(&sc.ev1)) != cudaSuccess) return e;
        if ((e = cudaMalloc((void**)&sc.d_counter, sizeof(unsigned int))) != cudaSuccess) return e;
    }
    size_t free_b = 0, total_b = 0;
    if ((e = cudaMemGetInfo(&free_b, &total_b)) != cudaSuccess) return e;
    // direction scratch budget: what is free now plus what this context already holds for it
    // (8 GB hold ~1 900 pairs of 4096 x 4096: more than one wave of warps; the buffer is allocated once and reused)
    const uint64_t budget = std::max<uint64_t>(std::min<uint64_t>((free_b + sc.cap) / 2, 8ull << 30), 64ull << 20);

    PhaseTimer pt;
    uint64_t done = 0;
    while (done < n_pairs) {
        // Greedy chunk in request order: as many pairs as fit the scratch budget (at least one).
        std::vector<PairJob> jobs;
        std::vector<uint32_t> job_req;      // request index (relative to `done`) of each job
        std::vector<uint64_t> cost;
        uint64_t words = 0;                 // scratch cursor, in 4-byte units
        uint64_t k = done;
        for (; k < n_pairs; k++) {
            PairJob j{};
            j.xs = inv[pairs_ij[2 * k]];
            j.ys = inv[pairs_ij[2 * k + 1]];
            j.out = (uint32_t)(k - done);
            const int n = (int)ar.len[j.xs], m = (int)ar.len[j.ys];
            const int np = n - 1, mp = m - 1;
            if (n < 1 || m < 1 || np == 0 || mp == 0) continue;   // no DP: answered on the host below
            const int w = (band_override >= 0) ? window_of_band(band_override, n, m) : window_of(pct, n, m);
            j.w = w;
            const int n_slabs = (np + PW_SLAB_ROWS - 1) / PW_SLAB_ROWS;
            int steps_max = 1;
            uint64_t cells = 0;
            for (int s = 0; s < n_slabs; s++) {
                int Jlo, Jhi;
                pw_slab_range(s, np, mp, w, Jlo, Jhi);
                if (Jhi >= Jlo) {
                    steps_max = std::max(steps_max, Jhi - Jlo + 1 + 31);
                    cells += (uint64_t)(Jhi - Jlo + 1 + 31) * 512;
                }
            }
            j.steps_max = (uint32_t)steps_max;
            const uint64_t dir_words = (uint64_t)n_slabs * steps_max * 32;
            const uint64_t row_words = ((uint64_t)4 * ((mp + 3) >> 2) + 3) & ~3ull;
            if (!jobs.empty() && ((words + dir_words + row_words) * 4 > budget || jobs.size() >= (1u << 24))) break;
            j.dir_off = words;
            words += dir_words;
            j.row_off = words;              // multiple of 4 words: float4 accesses are aligned
            words += row_words;
            jobs.push_back(j);
            job_req.push_back(j.out);
            cost.push_back(cells);
        }
        const size_t cnt_req = (size_t)(k - done);   // requests covered by this chunk (jobs + trivial ones)
        const size_t cnt = jobs.size();
        // host-answered requests (src/alignments.rs:116-125: n == m == 1 scores 0/2 = 0, a missing score cell +INF)
        for (uint64_t q = done; q < k; q++) {
            const int n = (int)ar.len[inv[pairs_ij[2 * q]]], m = (int)ar.len[inv[pairs_ij[2 * q + 1]]];
            if (n < 1 || m < 1 || n - 1 == 0 || m - 1 == 0) {
                scores[q] = (n == 1 && m == 1) ? 0.0f : INFINITY;
                if (path_lens) path_lens[q] = 0;
            }
        }
        pt.lap("paths: job list");
        if (cnt) {
            std::vector<uint32_t> order(cnt);
            std::iota(order.begin(), order.end(), 0u);
            std::stable_sort(order.begin(), order.end(), [&](uint32_t x, uint32_t y) { return cost[x] > cost[y]; });
            size_t jobs_bytes = cnt * sizeof(PairJob) + cnt * sizeof(uint32_t);
            if ((e = grow(sc.d_jobs, sc.jobs_cap, jobs_bytes)) != cudaSuccess) return e;
            if ((e = grow(sc.d_buf, sc.cap, words * 4)) != cudaSuccess) return e;
            if (sc.res_cap < cnt) {
                if (sc.d_scores) cudaFree(sc.d_scores);
                if (sc.d_lens) cudaFree(sc.d_lens);
                sc.d_scores = nullptr; sc.d_lens = nullptr; sc.res_cap = 0;
                if ((e = cudaMalloc((void**)&sc.d_scores, cnt * sizeof(float))) != cudaSuccess) return e;
                if ((e = cudaMalloc((void**)&sc.d_lens, cnt * sizeof(unsigned long long))) != cudaSuccess) return e;
                sc.res_cap = cnt;
            }
            if (paths_ij && (e = grow(sc.d_paths, sc.paths_cap, cnt * path_cap * 2 * sizeof(uint32_t))) != cudaSuccess) return e;
            pt.lap("paths: scratch allocation");
            PairJob* d_jobs = static_cast<PairJob*>(sc.d_jobs);
            uint32_t* d_order = reinterpret_cast<uint32_t*>(d_jobs + cnt);
            if ((e = cudaMemcpyAsync(d_jobs, jobs.data(), cnt * sizeof(PairJob), cudaMemcpyHostToDevice, stream)) != cudaSuccess) return e;
            if ((e = cudaMemcpyAsync(d_order, order.data(), cnt * sizeof(uint32_t), cudaMemcpyHostToDevice, stream)) != cudaSuccess) return e;
            if ((e = cudaMemsetAsync(sc.d_counter, 0, sizeof(unsigned int), stream)) != cudaSuccess) return e;
            WaveArgs a{};
            a.arena = d_arena; a.off = d_off; a.len = d_len;
            a.jobs = d_jobs; a.order = d_order; a.n_jobs = (uint32_t)cnt; a.counter = sc.d_counter;
            a.pins = ins; a.pdel = del; a.pmat = mat;
            a.scratch = static_cast<uint32_t*>(sc.d_buf);
            a.scores = sc.d_scores;
            const int grid = (int)std::min<size_t>((cnt + PW_WARPS - 1) / PW_WARPS, (size_t)sm_count * 2);
            if ((e = cudaEventRecord(sc.ev0, stream)) != cudaSuccess) return e;
            if ((e = launch_wave_dpad(ar.dpad, a, strict, grid, stream)) != cudaSuccess) return e;
            pair_trace_kernel<<<(unsigned)((cnt + 63) / 64), 64, 0, stream>>>(d_jobs, (uint32_t)cnt, d_len, static_cast<const uint32_t*>(sc.d_buf),
                                                                             paths_ij ? sc.d_paths : nullptr, path_cap, sc.d_lens);
            if ((e = cudaGetLastError()) != cudaSuccess) return e;
            if ((e = cudaEventRecord(sc.ev1, stream)) != cudaSuccess) return e;
            std::vector<float> sc_h(cnt);
            std::vector<unsigned long long> lens_h(cnt);
            if ((e = cudaMemcpyAsync(sc_h.data(), sc.d_scores, cnt * sizeof(float), cudaMemcpyDeviceToHost, stream)) != cudaSuccess) return e;
            if ((e = cudaMemcpyAsync(lens_h.data(), sc.d_lens, cnt * sizeof(unsigned long long), cudaMemcpyDeviceToHost, stream)) != cudaSuccess) return e;
            if ((e = cudaStreamSynchronize(stream)) != cudaSuccess) return e;
            pt.lap("paths: kernels + scores back");
            // paths: jobs keep the request order, so runs of consecutive requests copy as one block
            if (paths_ij) {
                size_t q = 0;
                while (q < cnt) {
                    size_t r = q + 1;
                    while (r < cnt && job_req[r] == job_req[r - 1] + 1) r++;
                    if ((e = cudaMemcpyAsync(paths_ij + (done + job_req[q]) * path_cap * 2, sc.d_paths + q * path_cap * 2,
                                             (r - q) * path_cap * 2 * sizeof(uint32_t), cudaMemcpyDeviceToHost, stream)) != cudaSuccess) return e;
                    q = r;
                }
                if ((e = cudaStreamSynchronize(stream)) != cudaSuccess) return e;
            }
            pt.lap("paths: path cells back");
            for (size_t q = 0; q < cnt; q++) {
                scores[done + job_req[q]] = sc_h[q];
                if (path_lens) path_lens[done + job_req[q]] = lens_h[q];
            }
            float t = 0.f;
            if (cudaEventElapsedTime(&t, sc.ev0, sc.ev1) == cudaSuccess && ms) *ms += t;
        }
        done += cnt_req;
    }
    return cudaSuccess;
}

}  // namespace apd
