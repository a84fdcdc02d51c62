// dtw_kernels.cuh -- sm_100a kernels of the all-pairs banded DTW (K1 of SURVEY.md
// section 2.1) built around the lane program of dtw_core.h.
//
// Launch shape: persistent grid of single-warp CTAs (grid = SMs x resident CTAs per
// SM, the latter set by the ring's shared-memory footprint), each warp pulling
// 32-pair work units from an atomic counter in LPT order.
//
// Data layout (see host_plan.h): one arena of zero-padded DPAD-float frames, every
// frame 16-byte aligned, sequences sorted by length.  Per warp in shared memory:
//   xs   : 2 x (4 frames x DPAD floats)  double-buffered stage of the shared row
//          sequence x (coalesced LDG.128 by lanes < DPAD -> STS.128, read back as
//          warp-broadcast LDS.128)
//   ring : St tiles x 4 rows x 32 lanes x float2 -- boundary column of the previous
//          column block, lane-contiguous so LDS.64/STS.64 are conflict-free.
// GSTATE kernels keep the ring in a per-CTA slice of a global scratch buffer
// instead (same layout, coalesced 256-byte rows, L2 resident) for bands too wide
// for shared memory.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "dtw_core.h"
#include "host_plan.h"

namespace apd {

struct KernelArgs {
    const float* arena;
    const uint32_t* off;   // frame offset of frame 0, sorted position
    const uint32_t* len;   // length, sorted position
    const Unit* units;     // ordered unit list (all ranks)
    uint32_t N;
    uint32_t rank, world;  // this launch handles units u = rank + world * k
    uint64_t k_begin;      // first local index k of the launch class
    uint32_t k_count;      // number of local units in the class
    unsigned int* counter; // work-fetch counter (zeroed before the launch)
    float pct;
    Penalties pen;
    int St;                // ring size in tiles
    float2* out;           // packed results: out[k * 32 + lane] = (score(a,b), score(b,a))
    float2* gstate;        // GSTATE: gridDim.x rings of St*4*32 float2
    int* error_flag;       // set to 1 if a unit needs a bigger ring than St (planner bug)
    unsigned long long* tiles_done;  // optional: lane-tiles executed (statistics)
};

#if defined(__CUDACC__)

template <int DPAD>
struct DevCtx {
    LaneGeom lg;
    RowGeom rg;
    int lane;
    const float4* xbase4;  // frame 0 of x
    const float4* ybase4;  // frame 0 of this lane's y
    float4* xs4;           // 2 x DPAD float4
    F2* st;                // this lane's ring column; row r lives at st[r * 32]
    int cur;
    float4 xreg;
    unsigned int tiles;

    APD_D void row_range(int J, int& Ilo, int& Ihi) const
    {
        int lo, hi;
        lane_row_range(lg, rg, J, lo, hi);
        Ilo = __reduce_min_sync(0xffffffffu, lo);
        Ihi = __reduce_max_sync(0xffffffffu, hi);
    }
    APD_D bool interior(int I, int J) const
    {
        return __all_sync(0xffffffffu, lane_tile_interior(lg, rg, I, J));
    }
    APD_D const float4* xaddr(int I) const
    {
        return xbase4 + (ptrdiff_t)(4 * I - rg.rho - 1) * (DPAD / 4) + lane;
    }
    APD_D void x_preload(int I)
    {
        __syncwarp();
        if (lane < DPAD) xs4[lane] = __ldg(xaddr(I));
        cur = 0;
        __syncwarp();
    }
    APD_D void x_prefetch(int I)
    {
        if (lane < DPAD) xreg = __ldg(xaddr(I));
        tiles++;
    }
    APD_D const float* x_tile() const { return reinterpret_cast<const float*>(xs4 + cur * DPAD); }
    APD_D void x_commit()
    {
        if (lane < DPAD) xs4[(cur ^ 1) * DPAD + lane] = xreg;
        cur ^= 1;
        __syncwarp();
    }
    APD_D F2 st_load(int row) const { return st[row * 32]; }
    APD_D void st_store(int row, F2 v) { st[row * 32] = v; }
    APD_D void load_y(int J, F2 (&yv)[TILE][DPAD / 2]) const
    {
        const float4* p = ybase4 + (ptrdiff_t)(4 * J - lg.gamma - 1) * (DPAD / 4);
#pragma unroll
        for (int c = 0; c < TILE; c++)
#pragma unroll
            for (int q = 0; q < DPAD / 4; q++) {
                float4 v = __ldg(p + c * (DPAD / 4) + q);
                yv[c][2 * q] = make_float2(v.x, v.y);
                yv[c][2 * q + 1] = make_float2(v.z, v.w);
            }
    }
};

template <int DPAD, bool STRICT, bool UNITW, bool GSTATE>
__global__ void __launch_bounds__(32) dtw_units_kernel(const KernelArgs a)
{
    extern __shared__ float4 smem4[];
    const int lane = threadIdx.x;
    DevCtx<DPAD> ctx;
    ctx.lane = lane;
    ctx.xs4 = smem4;
    F2* ring = GSTATE ? a.gstate + (size_t)blockIdx.x * ((size_t)a.St * TILE * 32)
                      : reinterpret_cast<F2*>(smem4 + 2 * DPAD);
    ctx.st = ring + lane;
    ctx.tiles = 0;
    const float4* arena4 = reinterpret_cast<const float4*>(a.arena);

    for (;;) {
        unsigned int k = 0;
        if (lane == 0) k = atomicAdd(a.counter, 1u);
        k = __shfl_sync(0xffffffffu, k, 0);
        if (k >= a.k_count) break;
        const uint64_t kk = a.k_begin + k;
        const Unit un = a.units[(uint64_t)a.rank + (uint64_t)a.world * kk];
        const uint32_t b = 32u * un.B + (uint32_t)lane;
        const bool exists = (b > un.a) && (b < a.N);
        const int n = (int)a.len[un.a];
        const int m = exists ? (int)a.len[b] : 0;
        ctx.rg = row_geometry(n);
        ctx.lg = lane_geometry(exists, n, m, a.pct);
        const int Jt_max = __reduce_max_sync(0xffffffffu, ctx.lg.Jt);
        const int wmax = __reduce_max_sync(0xffffffffu, ctx.lg.active ? ctx.lg.w : 0);
        float s1 = APD_INF, s2 = APD_INF;  // an empty side scores +INF (src/alignments.rs:116-125)
        if (ring_tiles_needed(wmax, ctx.rg.It > 0 ? ctx.rg.It : 1) > a.St) {
            if (lane == 0) atomicExch(a.error_flag, 1);
            s1 = s2 = __int_as_float(0x7fc00000);
        } else if (Jt_max > 0) {
            ctx.xbase4 = arena4 + (size_t)a.off[un.a] * (DPAD / 4);
            ctx.ybase4 = arena4 + (size_t)(exists ? a.off[b] : a.off[un.a]) * (DPAD / 4);
            F2 acc = run_unit<DPAD, STRICT, UNITW>(ctx, ctx.lg, ctx.rg, Jt_max, a.St, a.pen);
            if (ctx.lg.active) {
                s1 = finish_score(acc.x, n, m);
                s2 = finish_score(acc.y, n, m);
            }
        }
        a.out[kk * 32 + lane] = make_float2(s1, s2);
        __syncwarp();
    }
    if (a.tiles_done) {
        unsigned int t = __reduce_add_sync(0xffffffffu, ctx.tiles);
        if (lane == 0) atomicAdd(a.tiles_done, (unsigned long long)t);
    }
}

#endif  // __CUDACC__

// Per-DPAD launchers (dtw_inst.cu, one object per padded frame width).
typedef cudaError_t (*dtw_launch_fn)(const KernelArgs& a, bool strict, bool unitw, bool gstate,
                                     int grid, size_t smem, cudaStream_t stream);
typedef cudaError_t (*dtw_occupancy_fn)(bool strict, bool unitw, bool gstate, size_t smem,
                                        int* blocks_per_sm);

#define APD_DECLARE_DPAD(D)                                                                    \
    cudaError_t dtw_launch_##D(const KernelArgs& a, bool strict, bool unitw, bool gstate,      \
                               int grid, size_t smem, cudaStream_t stream);                    \
    cudaError_t dtw_occupancy_##D(bool strict, bool unitw, bool gstate, size_t smem,           \
                                  int* blocks_per_sm);
APD_DECLARE_DPAD(4)
APD_DECLARE_DPAD(8)
APD_DECLARE_DPAD(12)
APD_DECLARE_DPAD(16)
APD_DECLARE_DPAD(20)
APD_DECLARE_DPAD(24)
APD_DECLARE_DPAD(28)
APD_DECLARE_DPAD(32)

inline size_t dtw_smem_bytes(int dpad, int St, bool gstate)
{
    size_t x = (size_t)2 * 4 * dpad * sizeof(float);
    return gstate ? x : x + (size_t)St * TILE * 32 * sizeof(float2);
}

}  // namespace apd
