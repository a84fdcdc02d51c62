// dtw_emul.cpp -- host-side emulator of the CUDA kernel's schedule (TEST ONLY).
//
// Compiles audio_pattern_discovery_b200/csrc/dtw_core.h (the per-lane program the
// kernels run) and host_plan.cpp (arena packing + unit planning) with g++ and runs
// every lane of every work unit sequentially, with the warp reductions replaced by
// loops over the 32 lanes' geometry.  It lets the CPU test-suite check tiling, band
// masks, the boundary ring and the planner against the oracle without a GPU.  It is
// NOT a fallback: nothing in the package loads it.
#include <cstdint>
#include <algorithm>
#include <cstring>
#include <string>
#include <vector>

#include "../../audio_pattern_discovery_b200/csrc/dtw_core.h"
#include "../../audio_pattern_discovery_b200/csrc/host_plan.h"

using namespace apd;

namespace {

template <int DPAD, int TC>
struct HostCtx {
    const LaneGeom* lanes;  // 32
    RowGeom rg;
    int lane;
    const float* xbase;  // frame 0 of x
    const float* ybase;  // frame 0 of this lane's y
    std::vector<F2> state;
    uint64_t tiles = 0;

    void sweep_info(int J, int& Ilo, int& Ihi, int& Nlo, int& Nhi) const
    {
        Ilo = 0x3fffffff; Ihi = -1; Nlo = -0x3fffffff; Nhi = 0x3fffffff;
        for (int l = 0; l < 32; l++) {
            int lo, hi, nlo, nhi;
            lane_row_range(lanes[l], rg, J, lo, hi);
            lane_interior_range(lanes[l], rg, J, nlo, nhi);
            if (lo < Ilo) Ilo = lo;
            if (hi > Ihi) Ihi = hi;
            if (nlo > Nlo) Nlo = nlo;
            if (nhi < Nhi) Nhi = nhi;
        }
        // cross-check the closed-form interior range against the per-tile predicate
        for (int I = Ilo; I <= Ihi; I++) {
            bool all = true;
            for (int l = 0; l < 32; l++) all = all && lane_tile_interior(lanes[l], rg, I, J);
            if (all != (I >= Nlo && I <= Nhi)) mismatch = true;
        }
    }
    const float* xaddr(int I) const { return xbase + (int64_t)(4 * I - rg.rho - 1) * DPAD; }
    void note_step(int kind) { if (lane == first_lane) steps[kind]++; }
    void x_preload(int buf, int I) { stage[buf] = xaddr(I); }
    void x_fetch(int I, int buf, bool valid) { if (valid) stage[buf] = xaddr(I); }
    void x_wait(int) const {}
    // the final pipeline step computes distances of stale rows (discarded): any readable memory will do
    const float* x_tile(int buf) const { return stage[buf] ? stage[buf] : xaddr(0); }
    void ring_load(int slot, F2 (&v)[TILE]) const { for (int r = 0; r < TILE; r++) v[r] = state[slot * TILE + r]; }
    void ring_wait(F2 (&)[TILE]) const {}
    void ring_store(int slot, const F2 (&v)[TILE]) { for (int r = 0; r < TILE; r++) state[slot * TILE + r] = v[r]; tiles++; }
    F2 ring_load_last(int slot) const { return state[slot * TILE + TILE - 1]; }
    void switch_y(int J, F2 (&yv)[TC][DPAD / 2]) const
    {
        if (J >= lanes[lane].Jt) return;
        const float* p = ybase + (int64_t)(TC * J - lanes[lane].gamma - 1) * DPAD;
        for (int c = 0; c < TC; c++)
            for (int k = 0; k < DPAD / 2; k++) yv[c][k] = mk2(p[c * DPAD + 2 * k], p[c * DPAD + 2 * k + 1]);
    }
    const float* stage[X_STAGES] = {nullptr, nullptr, nullptr, nullptr};
    mutable bool mismatch = false;
    int first_lane = -1;
    uint64_t* steps = nullptr;
};

static uint64_t exact_units = 0;
static int g_tile_cols = 4;           // tests: 4 (the 8-warp kernels' tiles) or 2 (the 12-warp kernel's)
static int g_force_rho = -1;          // tests: run every unit on one fixed row grid
static uint64_t rho_used[4] = {0, 0, 0, 0};
static uint64_t step_counts[3] = {0, 0, 0};

template <int DPAD, int TC>
int run_all(const Arena& ar, const UnitPlan& plan, float pct, Penalties pen, int strict,
            bool unitw, uint32_t rank, uint32_t world, float* out, uint64_t* tiles_out)
{
    const uint32_t N = ar.n;
    uint64_t tiles = 0;
    for (const UnitClass& uc : plan.classes) {
        for (uint64_t u = uc.begin; u < uc.end; u++) {
            if (u % world != rank) continue;
            const Unit un = plan.units[u];
            const int n = (int)ar.len[un.a];
            LaneGeom lanes[32];
            int Jt_max = 0, wmax = 0;
            unsigned int votes = 0;
            for (int l = 0; l < 32; l++) {
                uint32_t b = 32 * un.B + l;
                bool exists = b > un.a && b < N;
                lanes[l] = lane_geometry(exists, n, exists ? (int)ar.len[b] : 0, pct, TC);
                if (lanes[l].Jt > Jt_max) Jt_max = lanes[l].Jt;
                if (lanes[l].active && lanes[l].w > wmax) wmax = lanes[l].w;
                const unsigned int v = lane_rho_votes(lanes[l], n);
                for (int r = 0; r < 4; r++) votes += ((v >> r) & 1u) << (8 * r);
            }
            RowGeom rg = row_geometry(n, g_force_rho >= 0 ? g_force_rho : choose_rho(votes, n));
            rho_used[rg.rho]++;
            if (ring_tiles_needed(wmax, rg.It > 0 ? rg.It : 1, TC) > uc.St) return -10;  // planner bug
            for (int l = 0; l < 32; l++) {
                uint32_t b = 32 * un.B + l;
                if (!(b > un.a && b < N)) continue;
                const int m = (int)ar.len[b];
                float s1, s2;
                if (!lanes[l].active) {
                    s1 = s2 = INFINITY;  // src/alignments.rs:116-125 with an empty side
                } else {
                    HostCtx<DPAD, TC> ctx;
                    ctx.lanes = lanes; ctx.rg = rg; ctx.lane = l;
                    ctx.steps = step_counts;
                    for (int q = 0; q < 32 && ctx.first_lane < 0; q++) if (lanes[q].active) ctx.first_lane = q;
                    ctx.xbase = ar.data.data() + (size_t)ar.off[un.a] * DPAD;
                    ctx.ybase = ar.data.data() + (size_t)ar.off[b] * DPAD;
                    F2 poison = mk2(-12345.0f, -54321.0f);  // stale ring entries must never be read
                    ctx.state.assign((size_t)uc.St * TILE, poison);
                    F2 acc;
                    SqrtFlags fl;
                    flags_reset(fl);
                    if (strict && unitw) acc = run_unit<DPAD, TC, true, true>(ctx, lanes[l], rg, Jt_max, uc.St, pen, fl);
                    else if (strict) acc = run_unit<DPAD, TC, true, false>(ctx, lanes[l], rg, Jt_max, uc.St, pen, fl);
                    else if (unitw) acc = run_unit<DPAD, TC, false, true>(ctx, lanes[l], rg, Jt_max, uc.St, pen, fl);
                    else acc = run_unit<DPAD, TC, false, false>(ctx, lanes[l], rg, Jt_max, uc.St, pen, fl);
                    // The host sqrt is exact everywhere, so the flags never change a result here;
                    // strict == 2 forces the cold path so that its schedule is exercised too.
                    if (strict == 2 || (strict && flags_bad(fl))) {
                        exact_units++;
                        ctx.state.assign((size_t)uc.St * TILE, poison);
                        acc = unitw ? run_unit_exact<DPAD, TC, true>(ctx, lanes[l], rg, Jt_max, uc.St, pen)
                                    : run_unit_exact<DPAD, TC, false>(ctx, lanes[l], rg, Jt_max, uc.St, pen);
                    }
                    if (ctx.mismatch) return -11;  // interior range formula disagrees with the tile predicate
                    s1 = finish_score(acc.x, n, m);
                    s2 = finish_score(acc.y, n, m);
                    if (l == 0 || tiles == 0) tiles += 0;
                    tiles += ctx.tiles;
                }
                const uint32_t ia = ar.perm[un.a], ib = ar.perm[b];
                out[(size_t)ia * N + ib] = s1;
                out[(size_t)ib * N + ia] = s2;
            }
        }
    }
    if (tiles_out) *tiles_out = tiles;
    return 0;
}

}  // namespace

extern "C" {

// Emulates apd_align_all (include/apd.h) for shard (rank, world) on the host.
// info (may be NULL, 16 entries; [12..14] pipeline steps by mask kind, [15] exact re-runs) receives: [0] units, [1] classes, [2] reference cells of the
// shard, [3] lane-tiles executed, [4..7] St of up to four classes, [8..11] 1 if gstate.
int apd_emul_align_all(const float* const* frames, const uint32_t* lens, uint32_t n, uint32_t dim,
                       float pct, float ins, float del, float mat, int strict, uint32_t rank,
                       uint32_t world, float* out_nxn, uint64_t* info)
{
    Arena ar;
    std::string err = build_arena(frames, lens, n, dim, ar);
    if (!err.empty()) return -1;
    if (ar.dpad > 32) return -2;
    UnitPlan plan;
    build_unit_plan(ar, pct, plan);
    std::memset(out_nxn, 0, (size_t)n * n * sizeof(float));
    Penalties pen; pen.ins = ins; pen.del = del; pen.mat = mat;
    const bool unitw = (ins == 1.0f && del == 1.0f && mat == 1.0f);
    uint64_t tiles = 0;
    int rc;
    switch (ar.dpad) {
        case 4: rc = g_tile_cols == 2 ? run_all<4, 2>(ar, plan, pct, pen, strict, unitw, rank, world, out_nxn, &tiles)
                                         : run_all<4, 4>(ar, plan, pct, pen, strict, unitw, rank, world, out_nxn, &tiles); break;
        case 8: rc = g_tile_cols == 2 ? run_all<8, 2>(ar, plan, pct, pen, strict, unitw, rank, world, out_nxn, &tiles)
                                         : run_all<8, 4>(ar, plan, pct, pen, strict, unitw, rank, world, out_nxn, &tiles); break;
        case 12: rc = g_tile_cols == 2 ? run_all<12, 2>(ar, plan, pct, pen, strict, unitw, rank, world, out_nxn, &tiles)
                                         : run_all<12, 4>(ar, plan, pct, pen, strict, unitw, rank, world, out_nxn, &tiles); break;
        case 16: rc = g_tile_cols == 2 ? run_all<16, 2>(ar, plan, pct, pen, strict, unitw, rank, world, out_nxn, &tiles)
                                         : run_all<16, 4>(ar, plan, pct, pen, strict, unitw, rank, world, out_nxn, &tiles); break;
        case 20: rc = g_tile_cols == 2 ? run_all<20, 2>(ar, plan, pct, pen, strict, unitw, rank, world, out_nxn, &tiles)
                                         : run_all<20, 4>(ar, plan, pct, pen, strict, unitw, rank, world, out_nxn, &tiles); break;
        case 24: rc = g_tile_cols == 2 ? run_all<24, 2>(ar, plan, pct, pen, strict, unitw, rank, world, out_nxn, &tiles)
                                         : run_all<24, 4>(ar, plan, pct, pen, strict, unitw, rank, world, out_nxn, &tiles); break;
        case 28: rc = g_tile_cols == 2 ? run_all<28, 2>(ar, plan, pct, pen, strict, unitw, rank, world, out_nxn, &tiles)
                                         : run_all<28, 4>(ar, plan, pct, pen, strict, unitw, rank, world, out_nxn, &tiles); break;
        case 32: rc = g_tile_cols == 2 ? run_all<32, 2>(ar, plan, pct, pen, strict, unitw, rank, world, out_nxn, &tiles)
                                         : run_all<32, 4>(ar, plan, pct, pen, strict, unitw, rank, world, out_nxn, &tiles); break;
        default: return -3;
    }
    if (info) {
        std::memset(info, 0, 16 * sizeof(uint64_t));
        info[0] = plan.units.size();
        info[1] = plan.classes.size();
        info[2] = reference_cells(ar, plan, rank, world);
        info[3] = tiles;
        info[12] = step_counts[0]; info[13] = step_counts[1]; info[14] = step_counts[2]; info[15] = exact_units;
        step_counts[0] = step_counts[1] = step_counts[2] = 0; exact_units = 0;
        for (size_t c = 0; c < plan.classes.size() && c < 4; c++) {
            info[4 + c] = (uint64_t)plan.classes[c].St;
            info[8 + c] = plan.classes[c].gstate ? 1 : 0;
        }
    }
    return rc;
}

// Planner only (no DP): info[0] units, [1] classes, [2] reference cells of shard (rank, world),
// [3] sum of the planner's tile estimates, [4..7] St per class, [8..11] gstate per class,
// [12..15] units per class.
int apd_emul_plan_info(const uint32_t* lens, uint32_t n, uint32_t dim, float pct, uint32_t rank, uint32_t world,
                       uint64_t* info)
{
    Arena ar;
    std::string err = build_arena_layout(lens, n, dim, ar);
    if (!err.empty()) return -1;
    UnitPlan plan;
    build_unit_plan(ar, pct, plan);
    std::memset(info, 0, 16 * sizeof(uint64_t));
    info[0] = plan.units.size();
    info[1] = plan.classes.size();
    info[2] = reference_cells(ar, plan, rank, world);
    info[3] = plan.tiles_estimate;
    for (size_t c = 0; c < plan.classes.size() && c < 4; c++) {
        info[4 + c] = (uint64_t)plan.classes[c].St;
        info[8 + c] = plan.classes[c].gstate ? 1 : 0;
        info[12 + c] = plan.classes[c].end - plan.classes[c].begin;
    }
    // every unordered pair appears in exactly one unit
    uint64_t pairs = 0;
    for (const Unit& u : plan.units) {
        uint32_t b0 = std::max(32 * u.B, u.a + 1), b1 = std::min(32 * u.B + 32, n);
        if (b1 > b0) pairs += b1 - b0;
    }
    if (pairs != (uint64_t)n * (n - 1) / 2) return -2;
    return 0;
}

// Forces the row grid (0..3) of every unit, -1 = the kernel's own choice; returns how many units
// used each grid since the last call (4 counters).
void apd_emul_force_rho(int rho, uint64_t* used4)
{
    g_force_rho = rho;
    if (used4) for (int r = 0; r < 4; r++) used4[r] = rho_used[r];
    for (int r = 0; r < 4; r++) rho_used[r] = 0;
}

// Tile width of the emulated lane program: 4 or 2 columns.
void apd_emul_tile_cols(int tc) { g_tile_cols = (tc == 2) ? 2 : 4; }

// The ordered unit list for a given enumeration block (row_block = 32 x sharers): units as (a << 32 | B),
// at most cap entries; returns the number of units, classes via info[0..3] = units per class.
uint64_t apd_emul_plan_units(const uint32_t* lens, uint32_t n, uint32_t dim, float pct, uint32_t row_block,
                             uint64_t* out, uint64_t cap, uint64_t* info)
{
    Arena ar;
    if (!build_arena_layout(lens, n, dim, ar).empty()) return 0;
    UnitPlan plan;
    build_unit_plan(ar, pct, plan, row_block);
    for (uint64_t k = 0; k < plan.units.size() && k < cap; k++) out[k] = ((uint64_t)plan.units[k].a << 32) | plan.units[k].B;
    if (info) for (size_t c = 0; c < 4; c++) info[c] = c < plan.classes.size() ? plan.classes[c].end - plan.classes[c].begin : 0;
    return plan.units.size();
}

uint64_t apd_emul_cells_visited(uint64_t n, uint64_t m, uint64_t w) { return cells_visited(n, m, w); }

int apd_emul_window(float pct, int n, int m) { return window_of(pct, n, m); }

}  // extern "C"
