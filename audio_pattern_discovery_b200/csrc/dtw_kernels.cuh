// dtw_kernels.cuh -- sm_100a kernels of the all-pairs banded DTW (K1 of SURVEY.md
// section 2.1) built around the lane program of dtw_core.h.
//
// Launch shape: persistent grid (grid = SMs x resident CTAs per SM; CTAs of 4 independent
// warps when the boundary ring lives in tensor memory, single-warp CTAs otherwise), each warp
// pulling 32-pair work units from an atomic counter in LPT order.  The results of a unit are
// stored to every buffer of KernelArgs::out -- in a single-process device group these are the
// gathered buffers of all member GPUs, written through NVLink peer mappings.
//
// Data layout (see host_plan.h): one arena of zero-padded DPAD-float frames, every
// frame 16-byte aligned, sequences sorted by length.  Per warp in shared memory:
//   xs   : 4 x (4 frames x DPAD floats)  stage of the shared row sequence x, two tiles
//          ahead of the recurrence: asynchronous 16-byte copies (cp.async -> LDGSTS), one
//          group per tile, read back as warp-broadcast LDS.128 (a TMA bulk-copy variant is
//          kept behind APD_X_STAGE_TMA; it measured slower)
//   ring : St tiles x 2 halves x 32 lanes x float4 -- boundary column of the previous
//          column block ((D1, D2) of 2 rows per float4), lane-contiguous so LDS.128 /
//          STS.128 are conflict-free.
// Where the ring lives is a kernel variant (RING_*):
//   RING_TMEM    in Blackwell tensor memory: each lane owns one TMEM lane of its warp's
//                32-lane quadrant, a ring tile is 8 consecutive 32-bit columns moved with
//                tcgen05.st / tcgen05.ld (.32x32b.x8).  TMEM is otherwise idle here (no
//                MMA), it costs no shared memory and no L2 traffic, and leaves occupancy to
//                the register file: CTAs of 4 warps, 256 columns each, 2 CTAs per SM.
//                Rings a little taller than the 32 tiles that fit (up to TMEM_RING_TILES +
//                TMEM_SPILL_TILES) keep their first 32 slots in tensor memory and the rest in
//                shared memory (same layout as RING_SMEM), which still leaves room for 2 CTAs
//                of 4 warps per SM -- 8 resident warps instead of the 3-4 of a pure
//                shared-memory ring.
//   RING_SMEM    in shared memory (layout above), single-warp CTAs; only when tensor memory
//                is ruled out (APD_RING=smem).
//   RING_GLOBAL  in a per-warp slice of a global scratch buffer (same layout, coalesced
//                512-byte rows, L2 resident); bands too tall for shared memory.
#pragma once
#include <cuda_runtime.h>

#ifndef APD_X_STAGE_TMA
#define APD_X_STAGE_TMA 0  // 1: stage x rows with cp.async.bulk + mbarrier (measured slower, see DevCtx::x_fetch)
#endif
#include <stdint.h>

#include "dtw_core.h"
#include "host_plan.h"

namespace apd {

enum { APD_MAX_GROUP = 8 };  // devices of one single-process group (one 8 x B200 box)

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
    // Packed results: out[q][k * 32 + lane] = (score(a,b), score(b,a)) for q < n_out.  A single
    // device writes one buffer.  In a single-process device group (apd_create_multi) every
    // member's kernel stores its results straight into the gathered buffer of EVERY member
    // through NVLink peer mappings (out[q] = member q's buffer + this member's rank slot): the
    // all-gather of the packed shards is fused into the kernel epilogue, 8 bytes per pair and
    // peer, and no collective follows the kernel.
    float2* out[APD_MAX_GROUP];
    uint32_t n_out;
    float2* gstate;        // GSTATE: gridDim.x rings of St*4*32 float2
    int* error_flag;       // set to 1 if a unit needs a bigger ring than St (planner bug)
    unsigned long long* tiles_done;  // optional: lane-tile columns (TILE cells each) executed (statistics)
    int carveout;          // host side only: preferred shared-memory carve-out of the launch (percent, -1 = driver default)
};

#if defined(__CUDACC__)

enum { X_BAR_F4 = 2 };  // float4 slots holding the X_STAGES mbarriers behind the stage buffers
enum { RING_SMEM = 0, RING_GLOBAL = 1, RING_TMEM = 2, RING_WIDE = 3 };
enum { TMEM_WARPS = 4, TMEM_COLS = 256, TMEM_RING_TILES = TMEM_COLS / 8 };
enum { TMEM_SPILL_TILES = 25 };
// Tile width of the 8-warps-per-SM kernels: 4 columns of y in registers (DPAD/2 register pairs
// each) -- 2 columns from 28-wide frames on, where 4 no longer fit the register file without
// spilling (26 is the reference's raw cepstrum width, src/spectrogram.rs:76).
template <int DPAD>
struct TileCols {
    static constexpr int value = DPAD >= 28 ? 2 : TILE;
};  // slots of a RING_TMEM ring that may live in shared memory: 2 CTAs x 4 warps x 25 KB < 227 KB

template <int DPAD, int RING, int TC = TILE>
struct DevCtx {
    LaneGeom lg;
    RowGeom rg;
    int lane;
    const float4* xbase4;  // frame 0 of x
    const float4* ybase4;  // frame 0 of this lane's y
    float4* xs4;           // X_STAGES x DPAD float4
    uint32_t xs_smem;      // shared-space address of xs4
    uint32_t bar_smem;     // shared-space address of the X_STAGES mbarriers (8 bytes each)
    uint32_t phase;        // bit b: parity the next wait on buffer b expects
    uint32_t pending;      // bit b: a bulk copy into buffer b has not been awaited yet
    float4* ring4;         // this lane's ring column: tile s, half h at ring4[(2 * s + h) * 32]
    int St;                // ring size in tiles
    uint32_t taddr;        // RING_TMEM: (first lane of the warp's quadrant << 16) | first column
    int tcap;              // RING_TMEM: ring slots held in tensor memory (the rest, if any, in shared memory)
    unsigned int tiles;

    APD_D void sweep_info(int J, int& Ilo, int& Ihi, int& Nlo, int& Nhi) const
    {
        int lo, hi, nlo, nhi;
        lane_row_range(lg, rg, J, lo, hi);
        lane_interior_range(lg, rg, J, nlo, nhi);
        Ilo = __reduce_min_sync(0xffffffffu, lo);
        Ihi = __reduce_max_sync(0xffffffffu, hi);
        Nlo = __reduce_max_sync(0xffffffffu, nlo);
        Nhi = __reduce_min_sync(0xffffffffu, nhi);
    }
    APD_D const float4* xaddr(int I) const
    {
        return xbase4 + (ptrdiff_t)(4 * I - rg.rho - 1) * (DPAD / 4) + lane;
    }
#if APD_X_STAGE_TMA
    // EXPERIMENT (-DAPD_X_STAGE_TMA=1, build.py --variant tma): x rows staged with the TMA unit's 1-D bulk
    // copy (cp.async.bulk, SASS UBLKCP): one elected lane arms the buffer's mbarrier with the byte
    // count and issues the copy of the tile's 4 contiguous frames; readers wait on the phase
    // parity.  Correct (GPU parity suite green) but measured 12 % SLOWER on the headline shape
    // (profiles/r2i_*TMA_staging_variant*: 669 vs 758 GCUPS STRICT, 926 vs 1131 FAST; FMA pipe 62.6 %
    // vs 69.0 %, stall_wait 0.79 vs 0.56 per issue): the ~90-cycle mbarrier.try_wait round trip is
    // exposed once per tile by in-order issue, and the bulk copy bypasses L1 where re-read x rows
    // otherwise hit.  Kept for reference; the default is cp.async below.
    APD_D void x_fetch(int I, int buf, bool valid)
    {
        __syncwarp();
        if (valid && lane == 0) {
            const uint32_t bar = bar_smem + 8u * (uint32_t)buf;
            const uint32_t dst = xs_smem + (uint32_t)(buf * DPAD * 16);
            const float4* src = xbase4 + (ptrdiff_t)(4 * I - rg.rho - 1) * (DPAD / 4);
            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"((uint32_t)(DPAD * 16)));
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
                         "l"(src), "r"((uint32_t)(DPAD * 16)), "r"(bar));
        }
        if (valid) pending |= 1u << buf;
    }
    APD_D void x_wait(int buf)
    {
        if ((pending >> buf) & 1u) {
            const uint32_t bar = bar_smem + 8u * (uint32_t)buf;
            const uint32_t parity = (phase >> buf) & 1u;
            asm volatile(
                "{\n.reg .pred p;\n"
                "WAIT_%=:\n"
                "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
                "@!p bra WAIT_%=;\n}"
                :
                : "r"(bar), "r"(parity)
                : "memory");
            phase ^= 1u << buf;
            pending &= ~(1u << buf);
        }
    }
    APD_D void x_preload(int buf, int I)
    {
        x_fetch(I, buf, true);
        x_wait(buf);
    }
#else
    // x rows are staged with asynchronous 16-byte copies (cp.async -> SASS LDGSTS): lanes
    // < DPAD each move one float4 of the tile's 4 contiguous frames straight from L1/L2 into
    // the stage buffer -- no staging register, no STS -- one committed group per pipeline step
    // (an empty one when there is nothing left to fetch, so the group count stays in step).
    // A tile's group was committed X_LOOK steps before its first read: wait_group X_LOOK-1.
    APD_D void x_fetch(int I, int buf, bool valid)
    {
        if (valid && lane < DPAD) {
            const uint32_t dst = xs_smem + (uint32_t)((buf * DPAD + lane) * 16);
            const float4* src = xbase4 + (ptrdiff_t)(4 * I - rg.rho - 1) * (DPAD / 4) + lane;
            asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src));
        }
        asm volatile("cp.async.commit_group;");
    }
    APD_D void x_wait(int)
    {
        asm volatile("cp.async.wait_group %0;" ::"n"(X_LOOK - 1) : "memory");
        __syncwarp();  // the other lanes' copies are visible, and nobody is still reading the
                       // buffer the next x_fetch overwrites
    }
    APD_D void x_preload(int buf, int I)
    {
        x_fetch(I, buf, true);
        asm volatile("cp.async.wait_group 0;" ::: "memory");
        __syncwarp();
    }
#endif
    APD_D void x_init()
    {
#if APD_X_STAGE_TMA
        if (lane == 0) {
#pragma unroll
            for (int b = 0; b < X_STAGES; b++)
                asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar_smem + 8u * (uint32_t)b) : "memory");
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        }
#endif
        phase = 0;
        pending = 0;
        for (int k = lane; k < X_STAGES * DPAD; k += 32) xs4[k] = make_float4(0.f, 0.f, 0.f, 0.f);
#if APD_X_STAGE_TMA
        // the zero fill (generic proxy) is ordered before later bulk copies (async proxy) into the same bytes
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
#endif
        __syncwarp();
    }
    APD_D void note_step(int) const {}
    APD_D const float* x_tile(int buf) const { return reinterpret_cast<const float*>(xs4 + buf * DPAD); }
    APD_D void ring_load(int slot, F2 (&v)[TILE]) const
    {
        if (RING == RING_TMEM && slot < tcap) {
            // Outputs go straight into v's registers: nothing may read them before ring_wait().
            asm volatile(
                "tcgen05.wait::st.sync.aligned;\n"
                "tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];\n"
                : "=f"(v[0].x), "=f"(v[0].y), "=f"(v[1].x), "=f"(v[1].y), "=f"(v[2].x), "=f"(v[2].y), "=f"(v[3].x),
                  "=f"(v[3].y)
                : "r"(taddr + 8u * (uint32_t)slot));
        } else {
            if (RING == RING_TMEM) slot -= tcap;   // the ring's tail in shared memory
            const float4 a = ring4[(2 * slot) * 32], b = ring4[(2 * slot + 1) * 32];
            v[0] = make_float2(a.x, a.y); v[1] = make_float2(a.z, a.w);
            v[2] = make_float2(b.x, b.y); v[3] = make_float2(b.z, b.w);
            if (RING == RING_GLOBAL) {
                // Rings this tall (unbanded long pairs: up to 1 MB per warp) stream from L2 / HBM;
                // they are walked sequentially, so pull the tile 4 steps ahead into L1 now.
                int sp = slot + 4;
                if (sp >= St) sp -= St;
                if (sp < St) {
                    asm volatile("prefetch.global.L1 [%0];" ::"l"(ring4 + (2 * sp) * 32));
                    asm volatile("prefetch.global.L1 [%0];" ::"l"(ring4 + (2 * sp + 1) * 32));
                }
            }
        }
    }
    // The tile requested by ring_load() may not be read before this (a no-op for the
    // scoreboarded shared / global loads; the registers are tied to the wait for TMEM).
    APD_D void ring_wait(F2 (&v)[TILE]) const
    {
        if (RING == RING_TMEM) {
            asm volatile("tcgen05.wait::ld.sync.aligned;\n"
                         : "+f"(v[0].x), "+f"(v[0].y), "+f"(v[1].x), "+f"(v[1].y), "+f"(v[2].x), "+f"(v[2].y),
                           "+f"(v[3].x), "+f"(v[3].y));
        }
    }
    // Completion of the store is awaited by the next ring_load (tcgen05.wait::st there).
    APD_D void ring_store(int slot, const F2 (&v)[TILE])
    {
        if (RING == RING_TMEM && slot < tcap) {
            asm volatile(
                "tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};\n"
                :
                : "r"(taddr + 8u * (uint32_t)slot), "r"(__float_as_uint(v[0].x)), "r"(__float_as_uint(v[0].y)),
                  "r"(__float_as_uint(v[1].x)), "r"(__float_as_uint(v[1].y)), "r"(__float_as_uint(v[2].x)),
                  "r"(__float_as_uint(v[2].y)), "r"(__float_as_uint(v[3].x)), "r"(__float_as_uint(v[3].y)));
        } else {
            if (RING == RING_TMEM) slot -= tcap;
            ring4[(2 * slot) * 32] = make_float4(v[0].x, v[0].y, v[1].x, v[1].y);
            ring4[(2 * slot + 1) * 32] = make_float4(v[2].x, v[2].y, v[3].x, v[3].y);
        }
        tiles++;
    }
    APD_D F2 ring_load_last(int slot) const
    {
        if (RING == RING_TMEM && slot < tcap) {
            uint32_t r0, r1;
            asm volatile(
                "tcgen05.wait::st.sync.aligned;\n"
                "tcgen05.ld.sync.aligned.32x32b.x2.b32 {%0, %1}, [%2];\n"
                "tcgen05.wait::ld.sync.aligned;\n"
                : "=r"(r0), "=r"(r1)
                : "r"(taddr + 8u * (uint32_t)slot + 6u));
            return make_float2(__uint_as_float(r0), __uint_as_float(r1));
        }
        if (RING == RING_TMEM) slot -= tcap;
        const float2* p = reinterpret_cast<const float2*>(ring4 + (2 * slot + 1) * 32);
        return p[1];
    }
    APD_D const float4* yaddr(int J) const
    {
        return ybase4 + (ptrdiff_t)(TC * J - lg.gamma - 1) * (DPAD / 4);
    }
    APD_D void switch_y(int J, F2 (&yv)[TC][DPAD / 2]) const
    {
        if (J < lg.Jt) {
            const float4* p = yaddr(J);
#pragma unroll
            for (int c = 0; c < TC; c++)
#pragma unroll
                for (int q = 0; q < DPAD / 4; q++) {
                    float4 v = __ldg(p + c * (DPAD / 4) + q);
                    yv[c][2 * q] = make_float2(v.x, v.y);
                    yv[c][2 * q + 1] = make_float2(v.z, v.w);
                }
        }
        if (J + 1 < lg.Jt) {  // the next block's frames: pull them towards the SM
            const char* p = reinterpret_cast<const char*>(yaddr(J + 1));
#pragma unroll
            for (int o = 0; o < TC * DPAD * 4; o += 128) asm volatile("prefetch.global.L2 [%0];" ::"l"(p + o));
            asm volatile("prefetch.global.L2 [%0];" ::"l"(p + TC * DPAD * 4 - 4));
        }
    }
};

// The persistent work loop of one warp: pulls 32-pair units from the class counter.
template <int DPAD, bool STRICT, bool UNITW, int RING, int TC>
APD_D void warp_unit_loop(const KernelArgs& a, DevCtx<DPAD, RING, TC>& ctx)
{
    const int lane = ctx.lane;
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
        ctx.lg = lane_geometry(exists, n, m, a.pct, TC);
        {   // the row grid that costs the fewest tiles for most of the warp's lanes (dtw_core.h: choose_rho)
            const unsigned int v = lane_rho_votes(ctx.lg, n);
            unsigned int votes = 0;
#pragma unroll
            for (int r = 0; r < 4; r++) votes |= (unsigned int)__popc(__ballot_sync(0xffffffffu, (v >> r) & 1u)) << (8 * r);
            ctx.rg = row_geometry(n, choose_rho(votes, n));
        }
        const int Jt_max = __reduce_max_sync(0xffffffffu, ctx.lg.Jt);
        const int wmax = __reduce_max_sync(0xffffffffu, ctx.lg.active ? ctx.lg.w : 0);
        float s1 = APD_INF, s2 = APD_INF;  // an empty side scores +INF (src/alignments.rs:116-125)
        if (ring_tiles_needed(wmax, ctx.rg.It > 0 ? ctx.rg.It : 1, TC) > a.St) {
            if (lane == 0) atomicExch(a.error_flag, 1);
            s1 = s2 = __int_as_float(0x7fc00000);
        } else if (Jt_max > 0) {
            ctx.xbase4 = arena4 + (size_t)a.off[un.a] * (DPAD / 4);
            ctx.ybase4 = arena4 + (size_t)(exists ? a.off[b] : a.off[un.a]) * (DPAD / 4);
            SqrtFlags fl;
            flags_reset(fl);
            F2 acc = run_unit<DPAD, TC, STRICT, UNITW>(ctx, ctx.lg, ctx.rg, Jt_max, a.St, a.pen, fl);
            if (STRICT && __any_sync(0xffffffffu, ctx.lg.active && flags_bad(fl))) {
                __syncwarp();
                acc = run_unit_exact<DPAD, TC, UNITW>(ctx, ctx.lg, ctx.rg, Jt_max, a.St, a.pen);
            }
            if (ctx.lg.active) {
                s1 = finish_score(acc.x, n, m);
                s2 = finish_score(acc.y, n, m);
            }
        }
        const float2 res = make_float2(s1, s2);
#pragma unroll 1
        for (uint32_t q = 0; q < a.n_out; q++) a.out[q][kk * 32 + lane] = res;
        __syncwarp();
    }
    if (a.tiles_done) {
        unsigned int t = __reduce_add_sync(0xffffffffu, ctx.tiles);
        if (lane == 0) atomicAdd(a.tiles_done, (unsigned long long)t * TC);   // in tile columns (TILE cells each)
    }
}

// RING_SMEM / RING_GLOBAL: single-warp CTAs.
template <int DPAD, bool STRICT, bool UNITW, int RING>
__global__ void __launch_bounds__(32) dtw_units_kernel(const KernelArgs a)
{
    extern __shared__ float4 smem4[];
    const int lane = threadIdx.x;
    DevCtx<DPAD, RING, TileCols<DPAD>::value> ctx;
    ctx.lane = lane;
    ctx.xs4 = smem4;
    ctx.xs_smem = (uint32_t)__cvta_generic_to_shared(smem4);
    ctx.bar_smem = ctx.xs_smem + X_STAGES * DPAD * 16;
    float4* ring = (RING == RING_GLOBAL) ? reinterpret_cast<float4*>(a.gstate) + (size_t)blockIdx.x * ((size_t)a.St * 2 * 32)
                                         : smem4 + X_STAGES * DPAD + X_BAR_F4;
    ctx.ring4 = ring + lane;
    ctx.taddr = 0;
    ctx.tcap = 0;
    ctx.St = a.St;
    ctx.tiles = 0;
    ctx.x_init();
    warp_unit_loop<DPAD, STRICT, UNITW, RING, TileCols<DPAD>::value>(a, ctx);
}

// RING_TMEM: CTAs of TMEM_WARPS independent warps sharing one tensor-memory allocation of
// TMEM_COLS columns; warp w owns the TMEM lanes 32w .. 32w+31 (the only ones tcgen05.ld/st
// issued by that warp can reach), lane l of the warp owns TMEM lane 32w + l.
template <int DPAD, bool STRICT, bool UNITW>
__global__ void __launch_bounds__(32 * TMEM_WARPS, 2) dtw_units_tmem_kernel(const KernelArgs a)
{
    extern __shared__ float4 smem4[];
    __shared__ uint32_t tmem_base_smem;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (warp == 0) {
        const uint32_t dst = (uint32_t)__cvta_generic_to_shared(&tmem_base_smem);
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;\n" ::"r"(dst), "r"((uint32_t)TMEM_COLS));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n");
    }
    asm volatile("tcgen05.fence::before_thread_sync;\n");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;\n");
    const uint32_t tmem_base = tmem_base_smem;

    DevCtx<DPAD, RING_TMEM, TileCols<DPAD>::value> ctx;
    ctx.lane = lane;
    const int spill = a.St > TMEM_RING_TILES ? a.St - TMEM_RING_TILES : 0;   // ring slots kept in shared memory
    float4* wbase = smem4 + warp * (X_STAGES * DPAD + X_BAR_F4 + spill * 2 * 32);
    ctx.xs4 = wbase;
    ctx.xs_smem = (uint32_t)__cvta_generic_to_shared(ctx.xs4);
    ctx.bar_smem = ctx.xs_smem + X_STAGES * DPAD * 16;
    ctx.ring4 = wbase + X_STAGES * DPAD + X_BAR_F4 + lane;
    ctx.taddr = tmem_base + ((uint32_t)(32 * warp) << 16);
    ctx.tcap = TMEM_RING_TILES;
    ctx.St = a.St;
    ctx.tiles = 0;
    ctx.x_init();
    warp_unit_loop<DPAD, STRICT, UNITW, RING_TMEM, TileCols<DPAD>::value>(a, ctx);

    asm volatile("tcgen05.fence::before_thread_sync;\n");
    __syncthreads();
    if (warp == 0) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;\n" ::"r"(tmem_base), "r"((uint32_t)TMEM_COLS));
    }
}

// RING_WIDE: ONE CTA of WIDE_WARPS = 12 independent warps per SM (3 per scheduler instead of 2)
// running the 4 x 2-column tile program.  Two columns of y cost 40 registers instead of 80, so the
// lane program fits the 168 registers a scheduler's register file allows three warps, and the
// third warp fills the issue slots the other two leave while they wait on each other's FMA-pipe
// work.  The CTA owns all 512 tensor-memory columns: warps 0..7 keep their boundary ring there
// (quadrant = warp & 3, column half = warp >> 2: 256 columns = 32 tiles each), warps 8..11 keep
// theirs in shared memory (same layout as RING_SMEM).  Rings taller than 32 tiles stay with the
// 8-warp kernels above.
enum { WIDE_WARPS = 12, WIDE_TMEM_WARPS = 8, WIDE_TC = 2 };

template <int DPAD, bool STRICT, bool UNITW>
__global__ void __launch_bounds__(32 * WIDE_WARPS, 1) dtw_units_wide_kernel(const KernelArgs a)
{
    extern __shared__ float4 smem4[];
    __shared__ uint32_t tmem_base_smem;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (warp == 0) {
        const uint32_t dst = (uint32_t)__cvta_generic_to_shared(&tmem_base_smem);
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;\n" ::"r"(dst), "r"(512u));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n");
    }
    asm volatile("tcgen05.fence::before_thread_sync;\n");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;\n");
    const uint32_t tmem_base = tmem_base_smem;

    DevCtx<DPAD, RING_TMEM, WIDE_TC> ctx;
    ctx.lane = lane;
    constexpr int XW = X_STAGES * DPAD + X_BAR_F4;   // float4 per warp: x stage + barriers
    ctx.xs4 = smem4 + warp * XW;
    ctx.xs_smem = (uint32_t)__cvta_generic_to_shared(ctx.xs4);
    ctx.bar_smem = ctx.xs_smem + X_STAGES * DPAD * 16;
    if (warp < WIDE_TMEM_WARPS) {
        ctx.taddr = tmem_base + ((uint32_t)(32 * (warp & 3)) << 16) + (uint32_t)((warp >> 2) * TMEM_COLS);
        ctx.tcap = TMEM_RING_TILES;
        ctx.ring4 = nullptr;
    } else {
        ctx.taddr = 0;
        ctx.tcap = 0;   // every slot in shared memory
        ctx.ring4 = smem4 + WIDE_WARPS * XW + (size_t)(warp - WIDE_TMEM_WARPS) * ((size_t)a.St * 2 * 32) + lane;
    }
    ctx.St = a.St;
    ctx.tiles = 0;
    ctx.x_init();
    warp_unit_loop<DPAD, STRICT, UNITW, RING_TMEM, WIDE_TC>(a, ctx);

    asm volatile("tcgen05.fence::before_thread_sync;\n");
    __syncthreads();
    if (warp == 0) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;\n" ::"r"(tmem_base), "r"(512u));
    }
}

#endif  // __CUDACC__

// Per-DPAD launchers (dtw_inst.cu, one object per padded frame width).
typedef cudaError_t (*dtw_launch_fn)(const KernelArgs& a, bool strict, bool unitw, int ring,
                                     int grid, size_t smem, cudaStream_t stream);
typedef cudaError_t (*dtw_occupancy_fn)(bool strict, bool unitw, int ring, size_t smem,
                                        int* blocks_per_sm);

#define APD_DECLARE_DPAD(D)                                                                    \
    cudaError_t dtw_launch_##D(const KernelArgs& a, bool strict, bool unitw, int ring,         \
                               int grid, size_t smem, cudaStream_t stream);                    \
    cudaError_t dtw_occupancy_##D(bool strict, bool unitw, int ring, size_t smem,              \
                                  int* blocks_per_sm);
APD_DECLARE_DPAD(4)
APD_DECLARE_DPAD(8)
APD_DECLARE_DPAD(12)
APD_DECLARE_DPAD(16)
APD_DECLARE_DPAD(20)
APD_DECLARE_DPAD(24)
APD_DECLARE_DPAD(28)
APD_DECLARE_DPAD(32)

// Dynamic shared memory per CTA (one warp, or TMEM_WARPS warps for RING_TMEM).
inline size_t dtw_smem_bytes(int dpad, int St, int ring)
{
    size_t x = (size_t)X_STAGES * 4 * dpad * sizeof(float) + X_BAR_F4 * 16;
    if (ring == RING_WIDE) return x * 12 + (size_t)4 * St * TILE * 32 * sizeof(float2);   // WIDE_WARPS, 4 shared-memory rings
    if (ring == RING_TMEM) {
        const int spill = St > TMEM_RING_TILES ? St - TMEM_RING_TILES : 0;
        return (x + (size_t)spill * TILE * 32 * sizeof(float2)) * TMEM_WARPS;
    }
    return ring == RING_GLOBAL ? x : x + (size_t)St * TILE * 32 * sizeof(float2);
}
inline int dtw_cta_warps(int ring) { return ring == RING_WIDE ? 12 : (ring == RING_TMEM ? 4 : 1); }

}  // namespace apd
