// host_plan.cpp -- see host_plan.h.
#include "host_plan.h"

#include <algorithm>
#include <cstring>
#include <numeric>
#include <thread>

#include "dtw_core.h"

namespace apd {

const int kSmemRingCaps[SMEM_RING_CAPS] = {16, 32, 56};

std::string build_arena_layout(const uint32_t* lens, uint32_t n, uint32_t dim, Arena& out)
{
    if (dim == 0) return "dim must be >= 1";
    if (n > 0 && !lens) return "null length table";
    out = Arena();
    out.n = n;
    out.dim = dim;
    out.dpad = (dim + 3u) & ~3u;
    out.perm.resize(n);
    std::iota(out.perm.begin(), out.perm.end(), 0u);
    // Length-sorted ("length-bucketed") order: the 32 column sequences of a unit
    // are neighbours in this order and therefore have similar band geometry.
    std::stable_sort(out.perm.begin(), out.perm.end(),
                     [&](uint32_t x, uint32_t y) { return lens[x] < lens[y]; });
    out.len.resize(n);
    out.off.resize(n);
    uint64_t frames_total = 0;
    for (uint32_t s = 0; s < n; s++) {
        uint32_t src = out.perm[s];
        if (lens[src] > (1u << 30)) return "sequence too long";
        out.len[s] = lens[src];
        frames_total += PRE_PAD_FRAMES;
        if (frames_total + lens[src] >= (1ull << 32)) return "arena exceeds 2^32 frames";
        out.off[s] = (uint32_t)frames_total;
        frames_total += lens[src];
    }
    frames_total += PRE_PAD_FRAMES;  // slack behind the last sequence
    out.total_frames = frames_total;
    return "";
}

static void fill_arena_range(const Arena& ar, const float* const* frames, float* dst, uint32_t s0, uint32_t s1)
{
    for (uint32_t s = s0; s < s1; s++) {
        // zero the pre-pad in front of the sequence (and the slack behind the last one)
        float* pad = dst + (size_t)(ar.off[s] - PRE_PAD_FRAMES) * ar.dpad;
        std::memset(pad, 0, (size_t)PRE_PAD_FRAMES * ar.dpad * sizeof(float));
        if (s + 1 == ar.n)
            std::memset(dst + (size_t)(ar.off[s] + ar.len[s]) * ar.dpad, 0, (size_t)PRE_PAD_FRAMES * ar.dpad * sizeof(float));
        const float* src = frames[ar.perm[s]];
        float* d = dst + (size_t)ar.off[s] * ar.dpad;
        if (!ar.len[s]) continue;
        if (ar.dpad == ar.dim) {
            std::memcpy(d, src, (size_t)ar.len[s] * ar.dim * sizeof(float));
        } else {
            for (uint32_t t = 0; t < ar.len[s]; t++) {
                float* row = d + (size_t)t * ar.dpad;
                std::memcpy(row, src + (size_t)t * ar.dim, ar.dim * sizeof(float));
                for (uint32_t k = ar.dim; k < ar.dpad; k++) row[k] = 0.0f;
            }
        }
    }
}

void fill_arena(const Arena& ar, const float* const* frames, float* dst)
{
    if (ar.n == 0) {
        std::memset(dst, 0, (size_t)ar.total_frames * ar.dpad * sizeof(float));
        return;
    }
    // The copy is memory bound; a few host threads get it close to the DRAM rate.
    const uint64_t bytes = (uint64_t)ar.total_frames * ar.dpad * sizeof(float);
    unsigned nt = std::thread::hardware_concurrency();
    nt = std::max(1u, std::min(nt, 8u));
    if (bytes < (8u << 20) || ar.n < 2 * nt) nt = 1;
    if (nt == 1) { fill_arena_range(ar, frames, dst, 0, ar.n); return; }
    std::vector<std::thread> th;
    for (unsigned t = 0; t < nt; t++) {
        const uint32_t s0 = (uint32_t)((uint64_t)ar.n * t / nt), s1 = (uint32_t)((uint64_t)ar.n * (t + 1) / nt);
        th.emplace_back([&ar, frames, dst, s0, s1] { fill_arena_range(ar, frames, dst, s0, s1); });
    }
    for (auto& x : th) x.join();
}

// Frames [F0, F1) of the arena (frame indices: every sequence is PRE_PAD_FRAMES zero frames,
// then its own frames; PRE_PAD_FRAMES zero frames of slack close the arena) -> dst, where
// dst[0] is the first float of frame F0.  Used by the chunked upload: the arena is packed
// piecewise into a small pinned ring while earlier pieces are already on their way to HBM.
void fill_arena_frames(const Arena& ar, const float* const* frames, uint64_t F0, uint64_t F1, float* dst)
{
    if (F1 <= F0) return;
    const size_t dpad = ar.dpad, dim = ar.dim;
    // first sequence whose region [off - PRE_PAD, off + len) ends behind F0
    uint32_t lo = 0, hi = ar.n;
    while (lo < hi) {
        const uint32_t mid = lo + (hi - lo) / 2;
        if ((uint64_t)ar.off[mid] + ar.len[mid] <= F0) lo = mid + 1; else hi = mid;
    }
    uint64_t f = F0;
    for (uint32_t s = lo; s < ar.n && f < F1; s++) {
        const uint64_t d0 = ar.off[s], d1 = d0 + ar.len[s];
        if (f < d0) {  // pre-pad zeros
            const uint64_t e = std::min<uint64_t>(d0, F1);
            std::memset(dst + (f - F0) * dpad, 0, (size_t)(e - f) * dpad * sizeof(float));
            f = e;
        }
        if (f >= F1) break;
        const uint64_t e = std::min<uint64_t>(d1, F1);
        if (e > f) {
            const float* src = frames[ar.perm[s]] + (size_t)(f - d0) * dim;
            float* d = dst + (f - F0) * dpad;
            if (dpad == dim) {
                std::memcpy(d, src, (size_t)(e - f) * dim * sizeof(float));
            } else {
                for (uint64_t t = 0; t < e - f; t++) {
                    float* row = d + (size_t)t * dpad;
                    std::memcpy(row, src + (size_t)t * dim, dim * sizeof(float));
                    for (size_t k = dim; k < dpad; k++) row[k] = 0.0f;
                }
            }
            f = e;
        }
    }
    if (f < F1) std::memset(dst + (f - F0) * dpad, 0, (size_t)(F1 - f) * dpad * sizeof(float));  // closing slack
}

std::string build_arena(const float* const* frames, const uint32_t* lens, uint32_t n,
                        uint32_t dim, Arena& out)
{
    std::string err = build_arena_layout(lens, n, dim, out);
    if (!err.empty()) return err;
    for (uint32_t s = 0; s < n; s++)
        if (lens[s] > 0 && (!frames || !frames[s])) return "null frames pointer for a non-empty sequence";
    out.data.resize((size_t)out.total_frames * out.dpad);
    fill_arena(out, frames, out.data.data());
    return "";
}

// Per-unit planning data of row sequence a against column block B (both sorted positions).
static inline void unit_cost(const Arena& ar, float pct, uint32_t a, uint32_t B, uint32_t& cost, int& cls, int& need)
{
    const uint32_t N = ar.n;
    const int n = (int)ar.len[a];
    const uint32_t last = std::min(32 * B + 31, N - 1);
    const int mmax = (int)ar.len[last];  // sorted ascending: the block's longest
    // b > a in sorted order => m >= n, and window_of is monotone in m there.
    const int wmax = window_of(pct, n, mmax);
    const int It = (n + 3) >> 2, Jt = (mmax + 3) >> 2;
    need = ring_tiles_needed(wmax, It + 1);  // + 1: the kernel may shift the row grid by up to 3 rows
    const int span = need - 1;
    const uint64_t cost64 = (uint64_t)std::max(Jt, 1) * (uint64_t)std::max(span, 1);
    cost = (uint32_t)std::min<uint64_t>(cost64, 0xffffffffu);
    cls = SMEM_RING_CAPS;  // gstate
    for (int k = 0; k < SMEM_RING_CAPS; k++)
        if (need <= kSmemRingCaps[k]) { cls = k; break; }
}

// The list is a pure function of (sorted lengths, pct): class by class, expensive units first
// inside a class (LPT order for the kernels' dynamic unit fetch).  It is built as a parallel
// counting sort over (class, cost bucket): the row sequences are dealt to host threads in
// contiguous ranges, every thread histograms its units, a prefix sum over (bucket, thread)
// gives each thread its private output cursor per bucket -- so the result is identical to the
// serial enumeration order inside a bucket whatever the thread count -- and a
// second pass writes the units.  10 000 sequences (1.57 M units): ~100 ms serial, ~15 ms on 8 threads;
// this is on the critical path of the first align call.
void build_unit_plan(const Arena& ar, float pct, UnitPlan& out, uint32_t row_block)
{
    out = UnitPlan();
    out.pct = pct;
    if (row_block < 32) row_block = 32;
    row_block = (row_block + 31) & ~31u;
    out.row_block = row_block;
    const uint32_t N = ar.n;
    if (N < 2) return;
    const uint32_t nblocks = (N + 31) / 32;
    const int n_cls = SMEM_RING_CAPS + 1;
    const int NB = 1024;
    const size_t n_bkt = (size_t)n_cls * NB;

    unsigned nt = std::thread::hardware_concurrency();
    nt = std::max(1u, std::min(nt, 16u));
    if ((uint64_t)N * nblocks < (1u << 16)) nt = 1;
    // Enumeration order inside a cost bucket: blocks of `row_block` row sequences, then the column
    // block B, then the row sequence inside its block -- row_block consecutive units share the 32
    // column sequences of B, so the ~1200 warps resident on a GPU work on a few dozen column blocks at
    // a time (tens of MB, L2 resident) instead of streaming the whole arena past every row sequence
    // (C3: 413 MB > L2).  row_block = 32 x (devices or ranks sharing the list): units are dealt
    // u mod world, so every device still sees 32 consecutive units of its own per column block.
    // Row blocks are dealt to the host threads in contiguous ranges with about equal unit counts.
    const uint32_t RB = row_block;
    const uint32_t nrb = (N - 1 + RB - 1) / RB;   // row sequences are 0 .. N-2
    auto units_of_rb = [&](uint32_t rb) -> uint64_t {
        uint64_t c = 0;
        const uint32_t a1 = std::min(rb * RB + RB, N - 1);
        for (uint32_t a = rb * RB; a < a1; a++) c += nblocks - (a + 1) / 32;
        return c;
    };
    std::vector<uint32_t> rb_begin(nt + 1, 0);
    {
        uint64_t total = 0;
        for (uint32_t rb = 0; rb < nrb; rb++) total += units_of_rb(rb);
        uint64_t acc = 0;
        unsigned t = 1;
        for (uint32_t rb = 0; rb < nrb && t < nt; rb++) {
            acc += units_of_rb(rb);
            if (acc * nt >= total * t) rb_begin[t++] = rb + 1;
        }
        for (; t < nt; t++) rb_begin[t] = nrb;
        rb_begin[nt] = nrb;
    }
    auto for_threads = [&](auto&& fn) {
        if (nt == 1) { fn(0u); return; }
        std::vector<std::thread> th;
        for (unsigned t = 0; t < nt; t++) th.emplace_back([&fn, t] { fn(t); });
        for (auto& x : th) x.join();
    };
    // visit(t, f): f(a, B) for every unit of thread t's row blocks, in enumeration order
    auto visit = [&](unsigned t, auto&& f) {
        for (uint32_t rb = rb_begin[t]; rb < rb_begin[t + 1]; rb++) {
            const uint32_t a0 = rb * RB, a1 = std::min(a0 + RB, N - 1);
            for (uint32_t B = (a0 + 1) / 32; B < nblocks; B++)
                for (uint32_t a = a0; a < a1; a++)
                    if (B >= (a + 1) / 32) f(a, B);
        }
    };

    // pass 1: maximum cost (bucket scale), per-class ring need and counts
    struct Acc { uint32_t max_cost = 1; int cls_need[SMEM_RING_CAPS + 1] = {0, 0, 0, 0}; uint64_t tiles = 0; };
    std::vector<Acc> acc(nt);
    for_threads([&](unsigned t) {
        Acc A;
        visit(t, [&](uint32_t a, uint32_t B) {
            uint32_t cost; int cls, need;
            unit_cost(ar, pct, a, B, cost, cls, need);
            A.max_cost = std::max(A.max_cost, cost);
            A.cls_need[cls] = std::max(A.cls_need[cls], need);
            A.tiles += cost;
        });
        acc[t] = A;
    });
    uint32_t max_cost = 1;
    int cls_need[SMEM_RING_CAPS + 1] = {0, 0, 0, 0};
    for (const Acc& A : acc) {
        max_cost = std::max(max_cost, A.max_cost);
        for (int c = 0; c < n_cls; c++) cls_need[c] = std::max(cls_need[c], A.cls_need[c]);
        out.tiles_estimate += A.tiles;
    }
    auto bucket = [&](uint32_t cost, int cls) {
        const uint64_t q = (uint64_t)cost * (NB - 1) / max_cost;  // 0..NB-1, monotone in cost
        return (size_t)cls * NB + (size_t)(NB - 1 - q);
    };
    // pass 2: histogram per thread
    std::vector<std::vector<uint64_t>> hist(nt, std::vector<uint64_t>(n_bkt, 0));
    for_threads([&](unsigned t) {
        std::vector<uint64_t>& h = hist[t];
        visit(t, [&](uint32_t a, uint32_t B) {
            uint32_t cost; int cls, need;
            unit_cost(ar, pct, a, B, cost, cls, need);
            h[bucket(cost, cls)]++;
        });
    });
    // exclusive prefix over (bucket major, thread minor)
    uint64_t pos = 0;
    uint64_t cls_count[SMEM_RING_CAPS + 1] = {0, 0, 0, 0};
    for (size_t b = 0; b < n_bkt; b++)
        for (unsigned t = 0; t < nt; t++) {
            const uint64_t cnt = hist[t][b];
            hist[t][b] = pos;
            pos += cnt;
            cls_count[b / NB] += cnt;
        }
    out.units.resize(pos);
    // pass 3: write
    for_threads([&](unsigned t) {
        std::vector<uint64_t>& cur = hist[t];
        visit(t, [&](uint32_t a, uint32_t B) {
            uint32_t cost; int cls, need;
            unit_cost(ar, pct, a, B, cost, cls, need);
            Unit u; u.a = a; u.B = B;
            out.units[cur[bucket(cost, cls)]++] = u;
        });
    });
    pos = 0;
    for (int c = 0; c < n_cls; c++) {
        if (!cls_count[c]) continue;
        UnitClass uc;
        uc.begin = pos;
        uc.end = pos + cls_count[c];
        uc.St = cls_need[c];
        uc.gstate = (c == SMEM_RING_CAPS);
        out.classes.push_back(uc);
        pos = uc.end;
    }
}

// sum_{k=0}^{K-1} max(0, min(A, B-k))
static uint64_t band_sum(int64_t A, int64_t B, int64_t K)
{
    if (A <= 0 || B <= 0 || K <= 0) return 0;
    int64_t k1 = std::min<int64_t>(std::max<int64_t>(B - A + 1, 0), K);  // terms equal to A
    int64_t k2 = std::min<int64_t>(B, K);                                // positive terms
    uint64_t s = (uint64_t)(A * k1);
    if (k2 > k1) s += (uint64_t)((k2 - k1) * B - (k1 + k2 - 1) * (k2 - k1) / 2);
    return s;
}

uint64_t cells_visited(uint64_t n, uint64_t m, uint64_t w)
{
    // diagonals j-i = 0..w-1 hold min(n, m-k) cells, diagonals j-i = -1..-w hold min(m, n-k)
    return band_sum((int64_t)n, (int64_t)m, (int64_t)w) +
           band_sum((int64_t)m, (int64_t)n - 1, (int64_t)w);
}

uint64_t reference_cells(const Arena& ar, const UnitPlan& plan, uint32_t rank, uint32_t world)
{
    uint64_t total = 0;
    const uint32_t N = ar.n;
    for (uint64_t u = rank; u < plan.units.size(); u += world) {
        const Unit& un = plan.units[u];
        const uint64_t n = ar.len[un.a];
        uint32_t b = std::max(32 * un.B, un.a + 1);
        const uint32_t bend = std::min(32 * un.B + 32, N);
        while (b < bend) {  // runs of equal length share one closed-form evaluation
            uint32_t e = b + 1;
            while (e < bend && ar.len[e] == ar.len[b]) e++;
            const uint64_t m = ar.len[b];
            if (n >= 1 && m >= 1) {
                uint64_t w = (uint64_t)window_of(plan.pct, (int)n, (int)m);
                total += (uint64_t)(e - b) * (cells_visited(n, m, w) + cells_visited(m, n, w));
            }
            b = e;
        }
    }
    return total;
}

}  // namespace apd
