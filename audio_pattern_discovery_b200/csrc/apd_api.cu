// apd_api.cu -- the C ABI of include/apd.h on top of the sm_100a kernels.
//
// There is no CPU path in this file: every entry point that computes anything needs
// a CUDA device, and apd_create() fails with APD_ERR_NO_DEVICE without one.
//
// A context drives one device (apd_create) or, from ONE host process, a group of up to
// APD_MAX_GROUP devices of one box (apd_create_multi).  In a group the member with index 0 is
// the leader: it owns the host-side state (arena layout, unit plan, statistics); every member
// holds the whole sequence arena, computes the work units u with u % G == its index, and its
// DTW kernel stores the packed results directly into the gathered buffer of every member
// through NVLink peer mappings (KernelArgs::out) -- the all-gather is fused into the kernel.
// Each member then expands the gathered buffer into the n x n matrix and copies its own slab
// of rows to the caller's host buffer, so the device->host copy runs on G PCIe links at once.
#include <cuda_runtime.h>

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <new>
#include <string>
#include <thread>
#include <vector>

#include "../../include/apd.h"
#include "ae_encode.cuh"
#include "apd_internal.h"
#include "dtw_kernels.cuh"
#include "host_plan.h"
#include "pair_path.cuh"
#include "percentile.cuh"

using namespace apd;

// The layouts the ctypes / Rust / C++ bindings mirror (tests/test_abi.py checks the Python side,
// tests/test_rust_sources.py the Rust side).
static_assert(sizeof(apd_params) == 20, "apd_params layout is part of the ABI");
static_assert(sizeof(apd_stats) == 104, "apd_stats layout is part of the ABI");
static_assert(sizeof(apd_merge) == 24, "apd_merge layout is part of the ABI");
static_assert(APD_MAX_DEVICES == APD_MAX_GROUP, "include/apd.h and dtw_kernels.cuh disagree on the group size");

namespace {

thread_local std::string g_create_error;

struct OccKey { int dpad, strict, unitw, ring; size_t smem; int occ; };

struct DevStatus {            // device-side flags of one align call, read back in one copy
    int error;                // a unit needed a bigger ring than planned
    int pad;
    unsigned long long tiles; // lane-tile columns executed (statistics)
};

enum { STAGE_BUFS = 3, MAX_CLASSES = SMEM_RING_CAPS + 1 };
const size_t kStageBytesMax = 16u << 20;  // one buffer of the pinned upload ring (small uploads get smaller buffers:
                                          // pinning costs ~0.7 ms per MB, which a cold first call of a small job would notice)

}  // namespace

namespace apd {
void set_thread_error(const std::string& msg) { g_create_error = msg; }
}  // namespace apd

struct apd_ctx {
    // ---- device-local state ------------------------------------------------------------
    int device = 0;
    cudaStream_t stream = nullptr;
    int sm_count = 0;
    float sm_clock_mhz = 0.f;
    size_t smem_optin = 0;

    float* d_arena = nullptr; size_t arena_cap = 0;
    uint32_t* d_off = nullptr; uint32_t* d_len = nullptr; uint32_t* d_perm = nullptr;
    uint64_t* d_srcoff = nullptr; size_t table_cap = 0;
    float* d_raw = nullptr; size_t raw_cap = 0;
    float* d_aux = nullptr; size_t aux_cap = 0;          // auto-encoder weights (apd_set_sequences_encoded)

    Unit* d_units = nullptr; size_t units_cap = 0;
    uint64_t units_serial = 0;                           // serial of the plan d_units holds

    float* d_packed = nullptr; size_t packed_cap = 0;    // packed results: own shard, or (group) every member's, rank-major
    float* d_matrix = nullptr; size_t matrix_cap = 0;
    float2* d_gstate = nullptr; size_t gstate_cap = 0;
    unsigned int* d_counters = nullptr;                  // one work counter per launch class
    DevStatus* d_status = nullptr;
    DevStatus h_status_buf{};                            // read back with one small copy per call (pageable: pinning
    DevStatus* h_status = &h_status_buf;                 // 16 bytes would cost more at creation than it ever saves)
    unsigned long long* d_hist = nullptr;                // 256 radix-select counters
    PathScratch path_scratch;                            // trace-back scratch (apd_align_pair(s)), grow-only

    cudaStream_t class_stream[MAX_CLASSES] = {nullptr, nullptr, nullptr, nullptr};
    cudaEvent_t ev_fork = nullptr, ev_join[MAX_CLASSES] = {nullptr, nullptr, nullptr, nullptr};
    cudaEvent_t ev_k0 = nullptr, ev_k1 = nullptr, ev_s0 = nullptr, ev_s1 = nullptr;
    cudaEvent_t ev_h0 = nullptr, ev_h1 = nullptr, ev_d0 = nullptr, ev_d1 = nullptr;
    cudaEvent_t ev_done = nullptr;   // the last work enqueued by this library that touches the context's device state
    cudaEvent_t ev_dtw = nullptr;    // group: this member's DTW kernels (and their peer stores) are complete
    bool inflight = false;
    bool timed_kernel = false, timed_scatter = false, timed_h2d = false, timed_d2h = false;
    std::vector<OccKey> occ_cache;
    float kernel_ms = 0.f, scatter_ms = 0.f;             // per-member timings of the last call
    uint32_t launches = 0;
    uint64_t local_units = 0;
    unsigned long long tiles = 0;

    // ---- group ---------------------------------------------------------------------------
    apd_ctx* lead = nullptr;                 // == this for a single-device context and for the leader
    std::vector<apd_ctx*> members;           // leader only: every member, members[0] == leader
    bool p2p = false;                        // leader: every member can store into every member's memory
    uint32_t rank = 0, world = 1;            // shard: multi-process (apd_set_shard) or index / size of the group

    // ---- host-side state (leader / single) -------------------------------------------------
    Arena arena;                             // layout only (data lives on the device)
    bool have_sequences = false;
    bool arena_pending = false;              // layout declared, contents to be filled by the caller (apd_set_sequences_layout)
    UnitPlan plan; bool plan_valid = false; uint64_t plan_serial = 0;
    uint64_t cells_ref = 0; bool cells_ref_valid = false;
    bool matrix_valid = false;               // d_matrix (of the leader) holds the last apd_align_all result
    float* h_stage[STAGE_BUFS] = {nullptr, nullptr, nullptr};   // pinned upload ring
    void* h_stage_raw[STAGE_BUFS] = {nullptr, nullptr, nullptr};
    size_t stage_bytes = 0;                  // size of each ring buffer
    cudaEvent_t ev_stage[STAGE_BUFS] = {nullptr, nullptr, nullptr};
    int ring_floor = RING_TMEM;              // APD_RING / APD_FORCE_GSTATE, read once at creation
    bool debug = false;                      // APD_DEBUG
    bool concurrent_classes = true;          // APD_SERIAL_CLASSES=1 turns it off
    bool uniform_carveout = true;            // APD_CARVEOUT=0 turns it off (see run_dtw)
    int wide = -1;                           // APD_WIDE: 1 / 0 force the 12-warps-per-SM kernel on / off, unset = by penalties

    apd_stats stats{};
    std::string err;
    std::string launch_desc;                 // launch classes of the last DTW enqueue (apd_last_launch_plan)
};

namespace {

apd_status fail(apd_ctx* c, apd_status s, const std::string& msg)
{
    if (c) { c->err = msg; if (c->lead && c->lead != c) c->lead->err = msg; }
    else g_create_error = msg;
    return s;
}

#define APD_CUDA(c, call)                                                                  \
    do {                                                                                   \
        cudaError_t e_ = (call);                                                           \
        if (e_ != cudaSuccess)                                                             \
            return fail(c, APD_ERR_CUDA, std::string(#call) + ": " + cudaGetErrorString(e_)); \
    } while (0)

template <class T>
apd_status ensure_device(apd_ctx* c, T*& ptr, size_t& cap, size_t need_elems)
{
    if (need_elems <= cap && ptr) return APD_OK;
    if (ptr) { cudaFree(ptr); ptr = nullptr; cap = 0; }
    size_t n = std::max<size_t>(need_elems, 1);
    APD_CUDA(c, cudaMalloc((void**)&ptr, n * sizeof(T)));
    cap = n;
    return APD_OK;
}

inline bool grouped(const apd_ctx* c) { return c->lead->members.size() > 1; }

// Work this library enqueued earlier (possibly on a caller's stream) may still be using the
// context's shared device state (unit list, counters, status, ring scratch, packed buffer):
// anything that is about to touch that state on `stream` is ordered behind it ...
apd_status guard_begin(apd_ctx* c, cudaStream_t stream)
{
    if (c->inflight) APD_CUDA(c, cudaStreamWaitEvent(stream, c->ev_done, 0));
    return APD_OK;
}
// ... and becomes the new "last work".
apd_status guard_end(apd_ctx* c, cudaStream_t stream)
{
    APD_CUDA(c, cudaEventRecord(c->ev_done, stream));
    c->inflight = true;
    return APD_OK;
}

struct LaunchFns { dtw_launch_fn launch; dtw_occupancy_fn occ; };

bool pick_launcher(uint32_t dpad, LaunchFns& f)
{
    switch (dpad) {
        case 4: f = {dtw_launch_4, dtw_occupancy_4}; return true;
        case 8: f = {dtw_launch_8, dtw_occupancy_8}; return true;
        case 12: f = {dtw_launch_12, dtw_occupancy_12}; return true;
        case 16: f = {dtw_launch_16, dtw_occupancy_16}; return true;
        case 20: f = {dtw_launch_20, dtw_occupancy_20}; return true;
        case 24: f = {dtw_launch_24, dtw_occupancy_24}; return true;
        case 28: f = {dtw_launch_28, dtw_occupancy_28}; return true;
        case 32: f = {dtw_launch_32, dtw_occupancy_32}; return true;
        default: return false;
    }
}

// Expands gathered packed shards into the row-major n x n matrix the reference's
// AlignmentWorkers.result holds (src/alignments.rs:56-57): result[i*n+j] for the
// caller's original indices i, j; the diagonal is zeroed by the caller (memset).
__global__ void scatter_packed_kernel(const float2* __restrict__ gathered, uint64_t k_per_rank,
                                      uint32_t world, uint32_t rank0, const Unit* __restrict__ units,
                                      uint64_t n_units, const uint32_t* __restrict__ perm,
                                      uint32_t N, float* __restrict__ out)
{
    const uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const uint64_t slot = t >> 5;  // rank-major: slot = (r - rank0) * k_per_rank + k
    const uint32_t lane = (uint32_t)(t & 31);
    const uint64_t r = slot / k_per_rank + rank0, k = slot % k_per_rank;
    if (r >= world) return;
    const uint64_t u = r + (uint64_t)world * k;
    if (u >= n_units) return;
    const Unit un = units[u];
    const uint32_t b = 32u * un.B + lane;
    if (!(b > un.a && b < N)) return;
    const float2 v = gathered[slot * 32 + lane];
    const uint32_t ia = perm[un.a], ib = perm[b];
    out[(size_t)ia * N + ib] = v.x;
    out[(size_t)ib * N + ia] = v.y;
}

// Gathers the caller's flat buffer into the sorted, padded arena (the device half of
// the sequence-packing glue).  One block per sequence (grid-stride), threads over
// the sequence's padded elements.
__global__ void pack_arena_kernel(const float* __restrict__ raw, const uint64_t* __restrict__ src_off,
                                  const uint32_t* __restrict__ off, const uint32_t* __restrict__ len,
                                  uint32_t n, uint32_t dim, uint32_t dpad, float* __restrict__ arena)
{
    for (uint32_t s = blockIdx.x; s < n; s += gridDim.x) {
        const float* src = raw + src_off[s];
        float* dst = arena + (size_t)off[s] * dpad;
        const uint32_t total = len[s] * dpad;
        for (uint32_t e = threadIdx.x; e < total; e += blockDim.x) {
            uint32_t t = e / dpad, k = e - t * dpad;
            dst[e] = (k < dim) ? src[(size_t)t * dim + k] : 0.0f;
        }
    }
}

// Sorted-order tables of the leader's arena layout -> device m.
apd_status upload_tables(apd_ctx* m, const uint64_t* src_off_sorted)
{
    const Arena& ar = m->lead->arena;
    const size_t n = ar.n;
    if (n > m->table_cap || !m->d_off) {
        if (m->d_off) cudaFree(m->d_off);
        if (m->d_len) cudaFree(m->d_len);
        if (m->d_perm) cudaFree(m->d_perm);
        if (m->d_srcoff) cudaFree(m->d_srcoff);
        m->d_off = m->d_len = m->d_perm = nullptr; m->d_srcoff = nullptr; m->table_cap = 0;
        size_t cap = std::max<size_t>(n, 1);
        APD_CUDA(m, cudaMalloc((void**)&m->d_off, cap * sizeof(uint32_t)));
        APD_CUDA(m, cudaMalloc((void**)&m->d_len, cap * sizeof(uint32_t)));
        APD_CUDA(m, cudaMalloc((void**)&m->d_perm, cap * sizeof(uint32_t)));
        APD_CUDA(m, cudaMalloc((void**)&m->d_srcoff, cap * sizeof(uint64_t)));
        m->table_cap = cap;
    }
    if (n) {
        APD_CUDA(m, cudaMemcpyAsync(m->d_off, ar.off.data(), n * sizeof(uint32_t), cudaMemcpyHostToDevice, m->stream));
        APD_CUDA(m, cudaMemcpyAsync(m->d_len, ar.len.data(), n * sizeof(uint32_t), cudaMemcpyHostToDevice, m->stream));
        APD_CUDA(m, cudaMemcpyAsync(m->d_perm, ar.perm.data(), n * sizeof(uint32_t), cudaMemcpyHostToDevice, m->stream));
        if (src_off_sorted)
            APD_CUDA(m, cudaMemcpyAsync(m->d_srcoff, src_off_sorted, n * sizeof(uint64_t), cudaMemcpyHostToDevice, m->stream));
    }
    return APD_OK;
}

// Start of every apd_set_sequences*: validates, builds the layout on the leader, sizes the
// arena on every member, and orders the coming uploads behind any kernels still in flight.
apd_status begin_sequences(apd_ctx* c, const uint32_t* lens, uint32_t n, uint32_t dim)
{
    if (!c) return APD_ERR_INVALID;
    if (c->lead != c) return fail(c, APD_ERR_INVALID, "not a context handle returned by apd_create / apd_create_multi");
    if (dim == 0) return fail(c, APD_ERR_INVALID, "dim must be >= 1");
    if (dim > APD_MAX_DIM) return fail(c, APD_ERR_UNSUPPORTED, "dim > APD_MAX_DIM (32) is not supported");
    if (n > 0 && !lens) return fail(c, APD_ERR_INVALID, "lens is NULL");
    const bool had = c->have_sequences;
    c->have_sequences = false;
    c->arena_pending = false;
    c->matrix_valid = false;
    Arena next;
    std::string e = build_arena_layout(lens, n, dim, next);
    if (!e.empty()) return fail(c, APD_ERR_INVALID, e);
    // The unit plan and the reference cell count depend on the lengths only: a new batch
    // with the same lengths (the usual case when a caller re-aligns re-encoded slices)
    // keeps them and the uploaded unit list.
    const bool same_layout = had && next.n == c->arena.n && next.dim == c->arena.dim && next.len == c->arena.len &&
                             next.perm == c->arena.perm;
    if (!same_layout) {
        c->plan_valid = false;
        c->cells_ref_valid = false;
    }
    c->arena = std::move(next);
    for (apd_ctx* m : c->members) {
        APD_CUDA(m, cudaSetDevice(m->device));
        apd_status s = guard_begin(m, m->stream);
        if (s != APD_OK) return s;
        s = ensure_device(m, m->d_arena, m->arena_cap, (size_t)c->arena.total_frames * c->arena.dpad);
        if (s != APD_OK) return s;
    }
    APD_CUDA(c, cudaSetDevice(c->device));
    return APD_OK;
}

// End of every apd_set_sequences*: the leader's arena is complete on its stream (event ev_h1
// recorded by the caller); the other members of a group pull it over NVLink.
apd_status broadcast_arena(apd_ctx* c)
{
    const size_t bytes = (size_t)c->arena.total_frames * c->arena.dpad * sizeof(float);
    for (apd_ctx* m : c->members) {
        if (m == c) continue;
        APD_CUDA(m, cudaSetDevice(m->device));
        apd_status s = upload_tables(m, nullptr);
        if (s != APD_OK) return s;
        APD_CUDA(m, cudaStreamWaitEvent(m->stream, c->ev_h1, 0));
        APD_CUDA(m, cudaMemcpyPeerAsync(m->d_arena, m->device, c->d_arena, c->device, bytes, m->stream));
        s = guard_end(m, m->stream);
        if (s != APD_OK) return s;
    }
    APD_CUDA(c, cudaSetDevice(c->device));
    return guard_end(c, c->stream);
}

void k_range(uint64_t begin, uint64_t end, uint32_t rank, uint32_t world, uint64_t& k0, uint64_t& k1)
{
    // smallest k with rank + world*k >= x
    auto first_k = [&](uint64_t x) -> uint64_t { return x <= rank ? 0 : (x - rank + world - 1) / world; };
    k0 = first_k(begin);
    k1 = first_k(end);
}

// Packed (score(a,b), score(b,a)) entries per shard -- identical on every rank / member.
uint64_t packed_entries(const apd_ctx* c)
{
    const apd_ctx* L = c->lead;
    uint64_t U = L->plan.units.size();
    return (U + c->world - 1) / c->world;
}

// Host plan on the leader (a pure function of the sorted lengths and pct), then the unit list
// on every member: the leader's copy comes from the host, the other members pull it from the
// leader over NVLink.
apd_status ensure_plan(apd_ctx* c, float pct)
{
    // The plan depends on pct only through the band (bit pattern compare keeps NaN stable).
    const uint32_t sharers = c->members.size() > 1 ? (uint32_t)c->members.size() : c->world;
    const uint32_t row_block = 32u * std::min<uint32_t>(std::max<uint32_t>(sharers, 1u), 16u);
    if (!(c->plan_valid && std::memcmp(&c->plan.pct, &pct, sizeof(float)) == 0 && c->plan.row_block == row_block)) {
        build_unit_plan(c->arena, pct, c->plan, row_block);
        c->cells_ref_valid = false;
        c->plan_valid = true;
        c->plan_serial++;
    }
    const size_t nu = c->plan.units.size();
    if (c->units_serial != c->plan_serial) {
        APD_CUDA(c, cudaSetDevice(c->device));
        apd_status s = guard_begin(c, c->stream);   // kernels of an earlier call may still read d_units
        if (s != APD_OK) return s;
        s = ensure_device(c, c->d_units, c->units_cap, nu);
        if (s != APD_OK) return s;
        // (cudaMemcpyAsync from pageable memory returns after staging the source; the vector is owned by the context)
        if (nu) APD_CUDA(c, cudaMemcpyAsync(c->d_units, c->plan.units.data(), nu * sizeof(Unit), cudaMemcpyHostToDevice, c->stream));
        s = guard_end(c, c->stream);
        if (s != APD_OK) return s;
        c->units_serial = c->plan_serial;
    }
    for (apd_ctx* m : c->members) {
        if (m == c || m->units_serial == c->plan_serial) continue;
        APD_CUDA(m, cudaSetDevice(m->device));
        apd_status s = guard_begin(m, m->stream);
        if (s != APD_OK) return s;
        s = ensure_device(m, m->d_units, m->units_cap, nu);
        if (s != APD_OK) return s;
        APD_CUDA(m, cudaStreamWaitEvent(m->stream, c->ev_done, 0));
        if (nu) APD_CUDA(m, cudaMemcpyPeerAsync(m->d_units, m->device, c->d_units, c->device, nu * sizeof(Unit), m->stream));
        s = guard_end(m, m->stream);
        if (s != APD_OK) return s;
        m->units_serial = c->plan_serial;
    }
    APD_CUDA(c, cudaSetDevice(c->device));
    return APD_OK;
}

apd_status occupancy(apd_ctx* c, const LaunchFns& f, int dpad, bool strict, bool unitw, int ring,
                     size_t smem, int& occ)
{
    for (const OccKey& k : c->occ_cache)
        if (k.dpad == dpad && k.strict == strict && k.unitw == unitw && k.ring == ring && k.smem == smem) {
            occ = k.occ;
            return APD_OK;
        }
    int o = 0;
    APD_CUDA(c, f.occ(strict, unitw, ring, smem, &o));
    if (o < 1) return fail(c, APD_ERR_INTERNAL, "kernel does not fit on an SM");
    c->occ_cache.push_back({dpad, strict, unitw, ring, smem, o});
    occ = o;
    return APD_OK;
}

apd_status check_params(apd_ctx* c, const apd_params* p)
{
    if (!p) return fail(c, APD_ERR_INVALID, "params is NULL");
    if (p->mode != APD_MODE_STRICT && p->mode != APD_MODE_FAST) return fail(c, APD_ERR_INVALID, "unknown mode");
    return APD_OK;
}

// Enqueues the DTW kernels of member / context m's shard on `stream`: one launch per class of
// the plan, most expensive class first, each class on its own stream so that the tail of one
// launch is filled by the next (the classes share nothing but the output buffers).
// outs[0..n_out): where the packed results go (see KernelArgs::out).
apd_status run_dtw(apd_ctx* m, const apd_params* p, float* const* outs, uint32_t n_out, cudaStream_t stream)
{
    apd_ctx* L = m->lead;
    const Arena& ar = L->arena;
    const UnitPlan& plan = L->plan;
    LaunchFns f;
    if (!pick_launcher(ar.dpad, f)) return fail(m, APD_ERR_UNSUPPORTED, "unsupported frame width");
    const bool strict = (p->mode == APD_MODE_STRICT);
    const bool unitw = (p->insertion_penalty == 1.0f && p->deletion_penalty == 1.0f && p->match_penalty == 1.0f);
    // Where the boundary ring lives (dtw_kernels.cuh): tensor memory if it fits 256 columns,
    // else shared memory, else global scratch.  APD_RING=tmem|smem|global (read at creation)
    // narrows the choice for experiments and tests (a ring that does not fit the forced home
    // moves down the list).
    const int ring_floor = L->ring_floor;

    apd_status s = guard_begin(m, stream);
    if (s != APD_OK) return s;
    APD_CUDA(m, cudaMemsetAsync(m->d_counters, 0, 8 * sizeof(unsigned int), stream));
    APD_CUDA(m, cudaMemsetAsync(m->d_status, 0, sizeof(DevStatus), stream));
    APD_CUDA(m, cudaEventRecord(m->ev_k0, stream));
    uint32_t launches = 0;
    uint64_t local_units = 0;
    const size_t ncls = plan.classes.size();
    if (ncls > MAX_CLASSES) return fail(m, APD_ERR_INTERNAL, "too many launch classes");

    // ring scratch of the global-ring class is sized before anything is launched (no realloc mid-flight)
    struct Cls { uint64_t k0, k1; int ring, occ, grid; size_t smem; };
    Cls cls[MAX_CLASSES];
    size_t gstate_need = 0;
    for (size_t ci = 0; ci < ncls; ci++) {
        const UnitClass& uc = plan.classes[ci];
        Cls& q = cls[ci];
        k_range(uc.begin, uc.end, m->rank, m->world, q.k0, q.k1);
        if (q.k1 <= q.k0) continue;
        q.ring = RING_GLOBAL;
        // (a ring up to TMEM_SPILL_TILES taller than tensor memory holds spills its tail to shared memory
        // and still runs 2 CTAs x 4 warps per SM)
        if (ring_floor == RING_TMEM && uc.St <= TMEM_RING_TILES + TMEM_SPILL_TILES) q.ring = RING_TMEM;
        else if (ring_floor != RING_GLOBAL && !uc.gstate) q.ring = RING_SMEM;
        // 12 warps per SM on 4 x 2-column tiles (dtw_units_wide_kernel) where the ring fits tensor memory outright
        // and the frames are at most 24 wide (beyond that even two columns of y exceed the 168 registers three
        // warps per scheduler leave).  Measured (profiles/README.md, r2t): the weighted recurrence is bound by
        // dependent-issue latency and gains 11 % from the third warp (C2 524 -> 581 GCUPS); with unit penalties
        // the 8-warp kernel is 1 % ahead (C3 758 vs 751), so that is the default policy.  APD_WIDE=0|1 overrides.
        const bool use_wide = L->wide == 1 || (L->wide < 0 && !unitw);
        if (q.ring == RING_TMEM && use_wide && uc.St <= TMEM_RING_TILES && ar.dpad <= 24) q.ring = RING_WIDE;
        q.smem = dtw_smem_bytes((int)ar.dpad, uc.St, q.ring);
        if (q.smem > m->smem_optin) return fail(m, APD_ERR_INTERNAL, "ring does not fit in shared memory");
        s = occupancy(m, f, (int)ar.dpad, strict, unitw, q.ring, q.smem, q.occ);
        if (s != APD_OK) return s;
        const uint64_t wpc = (uint64_t)dtw_cta_warps(q.ring);
        q.grid = (int)std::min<uint64_t>((q.k1 - q.k0 + wpc - 1) / wpc, (uint64_t)m->sm_count * q.occ);
        if (q.ring == RING_GLOBAL) gstate_need += (size_t)q.grid * uc.St * TILE * 32;
    }
    s = ensure_device(m, m->d_gstate, m->gstate_cap, gstate_need);
    if (s != APD_OK) return s;

    m->launch_desc.clear();
    // Kernels whose shared-memory needs differ run with different L1 / shared-memory splits, and an SM does not
    // change its split while CTAs of the other kind are resident -- the classes of one call would run one after
    // the other.  When a call launches more than one class they all ask for the same (maximum shared) split.
    int active_classes = 0;
    for (size_t ci = 0; ci < ncls; ci++) active_classes += (cls[ci].k1 > cls[ci].k0) ? 1 : 0;
    const int carveout = (L->uniform_carveout && L->concurrent_classes && active_classes > 1) ? (int)cudaSharedmemCarveoutMaxShared
                                                                                             : (int)cudaSharedmemCarveoutDefault;
    const bool fork = L->concurrent_classes && ncls > 1;
    if (fork) APD_CUDA(m, cudaEventRecord(m->ev_fork, stream));
    size_t gstate_used = 0;
    for (size_t cr = 0; cr < ncls; cr++) {
        const size_t ci = ncls - 1 - cr;   // classes are ordered by ring height: the tallest (most expensive units) first
        const UnitClass& uc = plan.classes[ci];
        const Cls& q = cls[ci];
        if (q.k1 <= q.k0) continue;
        local_units += q.k1 - q.k0;
        KernelArgs a{};
        a.arena = m->d_arena; a.off = m->d_off; a.len = m->d_len; a.units = m->d_units;
        a.N = ar.n; a.rank = m->rank; a.world = m->world;
        a.k_begin = q.k0; a.k_count = (uint32_t)(q.k1 - q.k0);
        a.counter = m->d_counters + ci;
        a.pct = p->warping_band_percentage;
        a.pen.ins = p->insertion_penalty; a.pen.del = p->deletion_penalty; a.pen.mat = p->match_penalty;
        a.St = uc.St;
        for (uint32_t o = 0; o < n_out; o++) a.out[o] = reinterpret_cast<float2*>(outs[o]);
        a.n_out = n_out;
        a.carveout = carveout;
        a.error_flag = &m->d_status->error;
        a.tiles_done = &m->d_status->tiles;
        if (q.ring == RING_GLOBAL) {
            a.gstate = m->d_gstate + gstate_used;
            gstate_used += (size_t)q.grid * uc.St * TILE * 32;
        }
        {
            char line[256];
            snprintf(line, sizeof(line), "%s{\"ring\": \"%s\", \"ring_tiles\": %d, \"units\": %llu, \"ctas\": %d, \"warps_per_cta\": %d, "
                     "\"ctas_per_sm\": %d, \"smem_bytes\": %zu}", m->launch_desc.empty() ? "" : ", ",
                     q.ring == RING_WIDE ? "tmem+smem, 12 warps" : (q.ring == RING_TMEM ? (uc.St > TMEM_RING_TILES ? "tmem+smem" : "tmem") : (q.ring == RING_SMEM ? "smem" : "global")), uc.St,
                     (unsigned long long)(q.k1 - q.k0), q.grid, dtw_cta_warps(q.ring), q.occ, q.smem);
            m->launch_desc += line;
            if (L->debug) fprintf(stderr, "[apd] dev %d class %zu: %s\n", m->device, ci, line);
        }
        cudaStream_t ls = stream;
        if (fork && launches > 0) {
            ls = m->class_stream[ci];
            APD_CUDA(m, cudaStreamWaitEvent(ls, m->ev_fork, 0));
        }
        APD_CUDA(m, f.launch(a, strict, unitw, q.ring, q.grid, q.smem, ls));
        if (ls != stream) {
            APD_CUDA(m, cudaEventRecord(m->ev_join[ci], ls));
            APD_CUDA(m, cudaStreamWaitEvent(stream, m->ev_join[ci], 0));
        }
        launches++;
    }
    APD_CUDA(m, cudaEventRecord(m->ev_k1, stream));
    m->timed_kernel = true;
    m->launches = launches;
    m->local_units = local_units;
    return guard_end(m, stream);
}

apd_status run_scatter(apd_ctx* m, const float* d_gathered, uint32_t nranks, float* d_out, cudaStream_t stream)
{
    apd_ctx* L = m->lead;
    const uint64_t N = L->arena.n;
    APD_CUDA(m, cudaEventRecord(m->ev_s0, stream));
    if (N) APD_CUDA(m, cudaMemsetAsync(d_out, 0, N * N * sizeof(float), stream));
    const uint64_t kpr = packed_entries(m);
    const uint64_t threads = (uint64_t)nranks * kpr * 32;
    if (threads) {
        const int bs = 256;
        const uint64_t blocks = (threads + bs - 1) / bs;
        if (blocks > 0x7fffffffull) return fail(m, APD_ERR_UNSUPPORTED, "matrix too large for one scatter launch");
        // nranks == world: buffer holds all ranks (rank-major).  nranks == 1 on a sharded
        // context: buffer holds only this rank's units.
        const uint32_t world = m->world;
        scatter_packed_kernel<<<(unsigned)blocks, bs, 0, stream>>>(
            reinterpret_cast<const float2*>(d_gathered), kpr, world, nranks == world ? 0u : m->rank, m->d_units,
            L->plan.units.size(), m->d_perm, (uint32_t)N, d_out);
        APD_CUDA(m, cudaGetLastError());
        m->launches++;
    }
    APD_CUDA(m, cudaEventRecord(m->ev_s1, stream));
    m->timed_scatter = true;
    return guard_end(m, stream);
}

// Waits for `stream` of member m and collects its status / timings.
apd_status finish_member(apd_ctx* m, cudaStream_t stream)
{
    APD_CUDA(m, cudaMemcpyAsync(m->h_status, m->d_status, sizeof(DevStatus), cudaMemcpyDeviceToHost, stream));
    APD_CUDA(m, cudaStreamSynchronize(stream));
    m->inflight = false;
    m->tiles = m->h_status->tiles;
    if (m->timed_kernel) { cudaEventElapsedTime(&m->kernel_ms, m->ev_k0, m->ev_k1); m->timed_kernel = false; }
    if (m->timed_scatter) { cudaEventElapsedTime(&m->scatter_ms, m->ev_s0, m->ev_s1); m->timed_scatter = false; }
    if (m->h_status->error) return fail(m, APD_ERR_INTERNAL, "a work unit needed a larger boundary ring than planned");
    return APD_OK;
}

// Statistics of the last align call, aggregated over the members (max of the times, sums of the counts).
void collect_stats(apd_ctx* c)
{
    float kms = 0.f, sms = 0.f;
    uint64_t tiles = 0, units = 0;
    uint32_t launches = 0;
    for (apd_ctx* m : c->members) {
        kms = std::max(kms, m->kernel_ms); sms = std::max(sms, m->scatter_ms);
        tiles += m->tiles; units += m->local_units; launches += m->launches;
    }
    c->stats.kernel_ms = kms; c->stats.scatter_ms = sms;
    c->stats.cells_computed = tiles * TILE * 2;   // tile columns x TILE rows x 2 orientations
    c->stats.units_local = units;
    c->stats.units_total = c->plan.units.size();
    c->stats.kernel_launches = launches;
    if (c->timed_h2d) { cudaEventElapsedTime(&c->stats.h2d_ms, c->ev_h0, c->ev_h1); c->timed_h2d = false; }
    if (c->timed_d2h) { cudaEventElapsedTime(&c->stats.d2h_ms, c->ev_d0, c->ev_d1); c->timed_d2h = false; }
}

void set_sequence_stats(apd_ctx* c, uint32_t n, uint64_t h2d_bytes)
{
    c->stats.h2d_bytes = h2d_bytes;
    c->stats.n_sequences = n;
    c->stats.ordered_pairs = (uint64_t)n * (n ? n - 1 : 0);
}

void release_stage_ring(apd_ctx* c)
{
    for (int b = 0; b < STAGE_BUFS; b++) {
        if (c->h_stage_raw[b]) { cudaHostUnregister(c->h_stage_raw[b]); free(c->h_stage_raw[b]); }
        c->h_stage_raw[b] = nullptr;
        c->h_stage[b] = nullptr;
    }
    c->stage_bytes = 0;
}

// The pinned ring the uploads are staged through: STAGE_BUFS buffers of about a third of the
// upload each (1 MB .. 16 MB), grown when a later upload is bigger.
apd_status ensure_stage_ring(apd_ctx* c, size_t upload_bytes)
{
    size_t want = (upload_bytes / STAGE_BUFS + (1u << 20)) & ~((size_t)(1u << 20) - 1);
    want = std::min(std::max<size_t>(want, 1u << 20), kStageBytesMax);
    if (c->h_stage[0] && c->stage_bytes >= want) return APD_OK;
    if (c->h_stage[0]) {
        for (int b = 0; b < STAGE_BUFS; b++) APD_CUDA(c, cudaEventSynchronize(c->ev_stage[b]));
        release_stage_ring(c);
    }
    for (int b = 0; b < STAGE_BUFS; b++) {
        // page-aligned host memory, registered (pinned) in place: several times cheaper than
        // cudaMallocHost, which matters for the first call of a fresh context
        void* raw = nullptr;
        if (posix_memalign(&raw, 4096, want) != 0) return fail(c, APD_ERR_INTERNAL, "out of host memory");
        std::memset(raw, 0, want);
        cudaError_t e = cudaHostRegister(raw, want, cudaHostRegisterPortable);
        if (e != cudaSuccess) { free(raw); return fail(c, APD_ERR_CUDA, std::string("cudaHostRegister: ") + cudaGetErrorString(e)); }
        c->h_stage_raw[b] = raw;
        c->h_stage[b] = static_cast<float*>(raw);
        if (!c->ev_stage[b]) APD_CUDA(c, cudaEventCreateWithFlags(&c->ev_stage[b], cudaEventDisableTiming));
    }
    c->stage_bytes = want;
    return APD_OK;
}

apd_status create_one(int device_id, apd_ctx** out)
{
    apd_ctx* c = new (std::nothrow) apd_ctx();
    if (!c) return fail(nullptr, APD_ERR_INTERNAL, "out of host memory");
    c->device = device_id;
    c->lead = c;
    cudaDeviceProp prop;
#define APD_CREATE_CUDA(call)                                                            \
    do {                                                                                 \
        cudaError_t e_ = (call);                                                         \
        if (e_ != cudaSuccess) {                                                         \
            g_create_error = std::string(#call) + ": " + cudaGetErrorString(e_);         \
            apd_destroy(c);                                                              \
            return APD_ERR_CUDA;                                                         \
        }                                                                                \
    } while (0)
    PhaseTimer pt;
    APD_CREATE_CUDA(cudaSetDevice(device_id));
    APD_CREATE_CUDA(cudaFree(0));
    pt.lap("create: device context");
    APD_CREATE_CUDA(cudaGetDeviceProperties(&prop, device_id));
    if (prop.major < 10) {
        g_create_error = "device is not sm_100-class (compute capability " + std::to_string(prop.major) + "." +
                         std::to_string(prop.minor) + "); libapd_b200 ships sm_100a code only";
        apd_destroy(c);
        return APD_ERR_NO_DEVICE;
    }
    c->sm_count = prop.multiProcessorCount;
    int khz = 0;
    cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, device_id);
    c->sm_clock_mhz = khz / 1000.0f;
    c->smem_optin = prop.sharedMemPerBlockOptin;
    APD_CREATE_CUDA(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
    for (int k = 0; k < MAX_CLASSES; k++) {
        APD_CREATE_CUDA(cudaStreamCreateWithFlags(&c->class_stream[k], cudaStreamNonBlocking));
        APD_CREATE_CUDA(cudaEventCreateWithFlags(&c->ev_join[k], cudaEventDisableTiming));
    }
    APD_CREATE_CUDA(cudaMalloc((void**)&c->d_counters, 8 * sizeof(unsigned int)));
    APD_CREATE_CUDA(cudaMalloc((void**)&c->d_status, sizeof(DevStatus)));
    APD_CREATE_CUDA(cudaMemset(c->d_status, 0, sizeof(DevStatus)));
    APD_CREATE_CUDA(cudaMalloc((void**)&c->d_hist, 256 * sizeof(unsigned long long)));
    cudaEvent_t* evs[] = {&c->ev_k0, &c->ev_k1, &c->ev_s0, &c->ev_s1, &c->ev_h0, &c->ev_h1, &c->ev_d0, &c->ev_d1};
    for (cudaEvent_t* ev : evs) APD_CREATE_CUDA(cudaEventCreate(ev));
    APD_CREATE_CUDA(cudaEventCreateWithFlags(&c->ev_done, cudaEventDisableTiming));
    APD_CREATE_CUDA(cudaEventCreateWithFlags(&c->ev_dtw, cudaEventDisableTiming));
    APD_CREATE_CUDA(cudaEventCreateWithFlags(&c->ev_fork, cudaEventDisableTiming));
    pt.lap("create: streams, events, small buffers");
#undef APD_CREATE_CUDA
    const char* ring_env = getenv("APD_RING");
    const char* force_g = getenv("APD_FORCE_GSTATE");
    if (ring_env && !strcmp(ring_env, "smem")) c->ring_floor = RING_SMEM;
    if ((ring_env && !strcmp(ring_env, "global")) || (force_g && force_g[0] == '1')) c->ring_floor = RING_GLOBAL;
    c->debug = getenv("APD_DEBUG") != nullptr;
    const char* ser = getenv("APD_SERIAL_CLASSES");
    c->concurrent_classes = !(ser && ser[0] == '1');
    const char* cv = getenv("APD_CARVEOUT");
    c->uniform_carveout = !(cv && cv[0] == '0');
    const char* wide = getenv("APD_WIDE");
    c->wide = wide ? (wide[0] == '1' ? 1 : 0) : -1;
    c->stats.sm_clock_mhz = c->sm_clock_mhz;
    c->stats.sm_count = (uint32_t)c->sm_count;
    c->members.push_back(c);
    *out = c;
    return APD_OK;
}

void destroy_one(apd_ctx* c)
{
    if (!c) return;
    cudaSetDevice(c->device);
    if (c->stream) cudaStreamSynchronize(c->stream);
    for (int k = 0; k < MAX_CLASSES; k++) {
        if (c->class_stream[k]) { cudaStreamSynchronize(c->class_stream[k]); cudaStreamDestroy(c->class_stream[k]); }
        if (c->ev_join[k]) cudaEventDestroy(c->ev_join[k]);
    }
    void* dptrs[] = {c->d_arena, c->d_off, c->d_len, c->d_perm, c->d_srcoff, c->d_raw, c->d_aux, c->d_units, c->d_packed,
                     c->d_matrix, c->d_gstate, c->d_counters, c->d_status, c->d_hist};
    for (void* p : dptrs) if (p) cudaFree(p);
    c->path_scratch.release();
    release_stage_ring(c);
    for (int b = 0; b < STAGE_BUFS; b++)
        if (c->ev_stage[b]) cudaEventDestroy(c->ev_stage[b]);
    cudaEvent_t evs[] = {c->ev_k0, c->ev_k1, c->ev_s0, c->ev_s1, c->ev_h0, c->ev_h1, c->ev_d0, c->ev_d1,
                         c->ev_done, c->ev_dtw, c->ev_fork};
    for (cudaEvent_t ev : evs) if (ev) cudaEventDestroy(ev);
    if (c->stream) cudaStreamDestroy(c->stream);
    delete c;
}

// Chunked upload of the arena: host threads pack the next piece of the sorted, padded arena
// into a pinned ring buffer while the previous pieces are on their way to HBM.  Returns when
// the caller's buffers are no longer needed (the last pieces may still be in flight).
apd_status upload_arena(apd_ctx* c, const float* const* frames)
{
    PhaseTimer pt;
    const Arena& ar = c->arena;
    apd_status s = ensure_stage_ring(c, (size_t)ar.total_frames * ar.dpad * sizeof(float));
    if (s != APD_OK) return s;
    pt.lap("upload: pinned ring ready");
    const uint64_t frames_per_buf = std::max<uint64_t>(c->stage_bytes / (ar.dpad * sizeof(float)), 1);
    unsigned nt = std::thread::hardware_concurrency();
    nt = std::max(1u, std::min(nt, 8u));
    int b = 0;
    uint64_t chunk = 0;
    for (uint64_t F0 = 0; F0 < ar.total_frames; F0 += frames_per_buf, chunk++, b = (b + 1) % STAGE_BUFS) {
        const uint64_t F1 = std::min<uint64_t>(F0 + frames_per_buf, ar.total_frames);
        APD_CUDA(c, cudaEventSynchronize(c->ev_stage[b]));  // the last copy out of this buffer (if any) is done
        float* dst = c->h_stage[b];
        const uint64_t nf = F1 - F0;
        const unsigned use = (nf * ar.dpad * sizeof(float) < (1u << 20)) ? 1u : nt;
        if (use == 1) {
            fill_arena_frames(ar, frames, F0, F1, dst);
        } else {
            std::vector<std::thread> th;
            for (unsigned t = 0; t < use; t++) {
                const uint64_t a = F0 + nf * t / use, e = F0 + nf * (t + 1) / use;
                th.emplace_back([&ar, frames, a, e, dst, F0] { fill_arena_frames(ar, frames, a, e, dst + (a - F0) * ar.dpad); });
            }
            for (auto& x : th) x.join();
        }
        APD_CUDA(c, cudaMemcpyAsync(c->d_arena + F0 * ar.dpad, dst, nf * ar.dpad * sizeof(float), cudaMemcpyHostToDevice, c->stream));
        APD_CUDA(c, cudaEventRecord(c->ev_stage[b], c->stream));
    }
    return APD_OK;
}

}  // namespace

extern "C" {

uint32_t apd_abi_version(void) { return APD_ABI_VERSION; }

const char* apd_last_error(const apd_ctx* ctx) { return ctx ? ctx->err.c_str() : g_create_error.c_str(); }

apd_status apd_device_count(int* count)
{
    if (!count) return APD_ERR_INVALID;
    *count = 0;
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n == 0)
        return fail(nullptr, APD_ERR_NO_DEVICE,
                    std::string("no CUDA device (this library has no CPU path): ") +
                        (e != cudaSuccess ? cudaGetErrorString(e) : "device count is 0"));
    *count = n;
    return APD_OK;
}

apd_status apd_create(int device_id, apd_ctx** out)
{
    if (!out) return fail(nullptr, APD_ERR_INVALID, "out is NULL");
    *out = nullptr;
    int count = 0;
    apd_status s = apd_device_count(&count);
    if (s != APD_OK) return s;
    if (device_id < 0 || device_id >= count) return fail(nullptr, APD_ERR_INVALID, "device_id out of range");
    return create_one(device_id, out);
}

apd_status apd_create_multi(const int* device_ids, int n_dev, apd_ctx** out)
{
    if (!out) return fail(nullptr, APD_ERR_INVALID, "out is NULL");
    *out = nullptr;
    int count = 0;
    apd_status s = apd_device_count(&count);
    if (s != APD_OK) return s;
    if (n_dev <= 0) { n_dev = std::min(count, (int)APD_MAX_DEVICES); device_ids = nullptr; }
    if (n_dev > APD_MAX_DEVICES) return fail(nullptr, APD_ERR_UNSUPPORTED, "more than APD_MAX_DEVICES devices in one group");
    std::vector<int> ids(n_dev);
    for (int k = 0; k < n_dev; k++) {
        ids[k] = device_ids ? device_ids[k] : k;
        if (ids[k] < 0 || ids[k] >= count) return fail(nullptr, APD_ERR_INVALID, "device id out of range");
        for (int q = 0; q < k; q++)
            if (ids[q] == ids[k]) return fail(nullptr, APD_ERR_INVALID, "a device appears twice in device_ids");
    }
    // one host thread per device: the primary contexts are created side by side
    std::vector<apd_ctx*> ms(n_dev, nullptr);
    std::vector<apd_status> st(n_dev, APD_OK);
    std::vector<std::string> errs(n_dev);
    {
        std::vector<std::thread> th;
        for (int k = 0; k < n_dev; k++)
            th.emplace_back([&, k] {
                st[k] = create_one(ids[k], &ms[k]);
                if (st[k] != APD_OK) errs[k] = g_create_error;  // thread_local: carry it over
            });
        for (auto& x : th) x.join();
    }
    for (int k = 0; k < n_dev; k++)
        if (st[k] != APD_OK) {
            g_create_error = errs[k];
            for (apd_ctx* m : ms) destroy_one(m);
            return st[k];
        }
    apd_ctx* L = ms[0];
    L->members.assign(ms.begin(), ms.end());
    for (int k = 0; k < n_dev; k++) {
        ms[k]->lead = L;
        ms[k]->rank = (uint32_t)k;
        ms[k]->world = (uint32_t)n_dev;
        if (k) ms[k]->members.clear();
    }
    // Peer mappings: with all of them in place the DTW kernels store their results into every
    // member's memory (NVLink / NVSwitch); without, the shards are copied after the kernels.
    bool p2p = true;
    for (int a = 0; a < n_dev && p2p; a++)
        for (int b = 0; b < n_dev; b++) {
            if (a == b) continue;
            int can = 0;
            if (cudaDeviceCanAccessPeer(&can, ids[a], ids[b]) != cudaSuccess || !can) { p2p = false; break; }
        }
    const char* nop2p = getenv("APD_NO_P2P");
    if (nop2p && nop2p[0] == '1') p2p = false;
    if (p2p)
        for (int a = 0; a < n_dev; a++) {
            cudaSetDevice(ids[a]);
            for (int b = 0; b < n_dev; b++) {
                if (a == b) continue;
                cudaError_t e = cudaDeviceEnablePeerAccess(ids[b], 0);
                if (e == cudaErrorPeerAccessAlreadyEnabled) { cudaGetLastError(); e = cudaSuccess; }
                if (e != cudaSuccess) { cudaGetLastError(); p2p = false; }
            }
        }
    L->p2p = p2p && n_dev > 1;
    cudaSetDevice(L->device);
    *out = L;
    return APD_OK;
}

void apd_destroy(apd_ctx* c)
{
    if (!c) return;
    std::vector<apd_ctx*> ms = c->members;   // copy: destroy_one deletes the leader too
    if (ms.empty()) { destroy_one(c); return; }
    for (size_t k = ms.size(); k-- > 0;) destroy_one(ms[k]);
}

apd_status apd_group_size(apd_ctx* c, uint32_t* n_dev, uint32_t* peer_stores)
{
    if (!c || !n_dev) return APD_ERR_INVALID;
    *n_dev = (uint32_t)c->lead->members.size();
    if (peer_stores) *peer_stores = c->lead->p2p ? 1u : 0u;
    return APD_OK;
}

apd_status apd_set_sequences(apd_ctx* c, const float* const* frames, const uint32_t* lens, uint32_t n, uint32_t dim)
{
    PhaseTimer pt;
    apd_status s = begin_sequences(c, lens, n, dim);
    if (s != APD_OK) return s;
    pt.lap("set_sequences: layout + device arena");
    if (n > 0 && !frames) return fail(c, APD_ERR_INVALID, "frames is NULL");
    for (uint32_t k = 0; k < n; k++)
        if (lens[k] > 0 && !frames[k]) return fail(c, APD_ERR_INVALID, "frames[k] is NULL for a non-empty sequence");
    const size_t floats = (size_t)c->arena.total_frames * c->arena.dpad;
    APD_CUDA(c, cudaEventRecord(c->ev_h0, c->stream));
    // CPU half of the packing glue: sort + pad, piecewise through the pinned ring, overlapped with the H2D copies.
    s = upload_arena(c, frames);
    if (s != APD_OK) return s;
    pt.lap("set_sequences: pack + H2D (pinned ring)");
    s = upload_tables(c, nullptr);
    if (s != APD_OK) return s;
    APD_CUDA(c, cudaEventRecord(c->ev_h1, c->stream));
    c->timed_h2d = true;
    set_sequence_stats(c, n, floats * sizeof(float));
    s = broadcast_arena(c);
    if (s != APD_OK) return s;
    // The caller's buffers have been copied; the tail of the upload is still in flight on the
    // context's stream, and everything that follows is ordered behind it.
    c->have_sequences = true;
    return APD_OK;
}

apd_status apd_set_sequences_flat(apd_ctx* c, const float* flat, const uint64_t* offsets, const uint32_t* lens,
                                  uint32_t n, uint32_t dim)
{
    apd_status s = begin_sequences(c, lens, n, dim);
    if (s != APD_OK) return s;
    if (n > 0 && (!flat || !offsets)) return fail(c, APD_ERR_INVALID, "flat/offsets is NULL");
    uint64_t extent = 0;
    std::vector<uint64_t> src_sorted(n);
    for (uint32_t k = 0; k < n; k++) extent = std::max<uint64_t>(extent, offsets[k] + (uint64_t)lens[k] * dim);
    for (uint32_t sidx = 0; sidx < n; sidx++) src_sorted[sidx] = offsets[c->arena.perm[sidx]];
    s = ensure_device(c, c->d_raw, c->raw_cap, (size_t)extent);
    if (s != APD_OK) return s;
    APD_CUDA(c, cudaEventRecord(c->ev_h0, c->stream));
    if (extent)
        APD_CUDA(c, cudaMemcpyAsync(c->d_raw, flat, extent * sizeof(float), cudaMemcpyHostToDevice, c->stream));
    s = upload_tables(c, src_sorted.data());
    if (s != APD_OK) return s;
    // Device half of the packing glue: zero the arena (pads), then gather.
    APD_CUDA(c, cudaMemsetAsync(c->d_arena, 0, (size_t)c->arena.total_frames * c->arena.dpad * sizeof(float), c->stream));
    if (n) {
        int grid = (int)std::min<uint32_t>(n, (uint32_t)c->sm_count * 16);
        pack_arena_kernel<<<grid, 256, 0, c->stream>>>(c->d_raw, c->d_srcoff, c->d_off, c->d_len, n, dim,
                                                       c->arena.dpad, c->d_arena);
        APD_CUDA(c, cudaGetLastError());
    }
    APD_CUDA(c, cudaEventRecord(c->ev_h1, c->stream));
    c->timed_h2d = true;
    set_sequence_stats(c, n, extent * sizeof(float));
    s = broadcast_arena(c);
    if (s != APD_OK) return s;
    APD_CUDA(c, cudaStreamSynchronize(c->stream));  // src_sorted and the caller's (possibly pageable) buffer may go away
    c->have_sequences = true;
    return APD_OK;
}

apd_status apd_set_sequences_encoded(apd_ctx* c, const float* const* cepstra, const uint32_t* lens, uint32_t n,
                                     uint32_t n_bins, const float* w_encode, const float* b_encode, uint32_t n_latent)
{
    if (!c) return APD_ERR_INVALID;
    if (n_bins == 0 || n_bins > APD_AE_MAX_BINS) return fail(c, APD_ERR_UNSUPPORTED, "n_bins must be 1..APD_AE_MAX_BINS");
    if (!w_encode || !b_encode) return fail(c, APD_ERR_INVALID, "w_encode/b_encode is NULL");
    apd_status s = begin_sequences(c, lens, n, n_latent);
    if (s != APD_OK) return s;
    if (n > 0 && !cepstra) return fail(c, APD_ERR_INVALID, "cepstra is NULL");
    for (uint32_t k = 0; k < n; k++)
        if (lens[k] > 0 && !cepstra[k]) return fail(c, APD_ERR_INVALID, "cepstra[k] is NULL for a non-empty sequence");
    const Arena& ar = c->arena;
    // Raw cepstra go up in sorted order, densely (no padding): sequence s starts at src_sorted[s] floats.
    std::vector<uint64_t> src_sorted(n);
    uint64_t total_frames = 0;
    for (uint32_t sidx = 0; sidx < n; sidx++) { src_sorted[sidx] = total_frames * n_bins; total_frames += ar.len[sidx]; }
    const uint64_t extent = total_frames * n_bins;
    s = ensure_device(c, c->d_raw, c->raw_cap, (size_t)extent);
    if (s != APD_OK) return s;
    const size_t wfloats = (size_t)n_bins * n_latent + n_latent;
    s = ensure_device(c, c->d_aux, c->aux_cap, wfloats);
    if (s != APD_OK) return s;
    s = ensure_stage_ring(c, (size_t)extent * sizeof(float));
    if (s != APD_OK) return s;
    APD_CUDA(c, cudaEventRecord(c->ev_h0, c->stream));
    // weights: one small pageable copy each (staged before the call returns)
    APD_CUDA(c, cudaMemcpyAsync(c->d_aux, w_encode, (size_t)n_bins * n_latent * sizeof(float), cudaMemcpyHostToDevice, c->stream));
    APD_CUDA(c, cudaMemcpyAsync(c->d_aux + (size_t)n_bins * n_latent, b_encode, n_latent * sizeof(float), cudaMemcpyHostToDevice, c->stream));
    // cepstra through the pinned ring, sequence by sequence in sorted order
    {
        const uint64_t buf_floats = c->stage_bytes / sizeof(float);
        int b = 0;
        uint64_t chunk = 0, fill = 0, dst0 = 0;
        auto flush = [&]() -> apd_status {
            if (!fill) return APD_OK;
            APD_CUDA(c, cudaMemcpyAsync(c->d_raw + dst0, c->h_stage[b], fill * sizeof(float), cudaMemcpyHostToDevice, c->stream));
            APD_CUDA(c, cudaEventRecord(c->ev_stage[b], c->stream));
            dst0 += fill; fill = 0; chunk++; b = (b + 1) % STAGE_BUFS;
            APD_CUDA(c, cudaEventSynchronize(c->ev_stage[b]));  // the last copy out of the next buffer (if any) is done
            return APD_OK;
        };
        APD_CUDA(c, cudaEventSynchronize(c->ev_stage[0]));
        for (uint32_t sidx = 0; sidx < n; sidx++) {
            const float* src = cepstra[ar.perm[sidx]];
            uint64_t left = (uint64_t)ar.len[sidx] * n_bins;
            while (left) {
                const uint64_t take = std::min<uint64_t>(left, buf_floats - fill);
                std::memcpy(c->h_stage[b] + fill, src, take * sizeof(float));
                fill += take; src += take; left -= take;
                if (fill == buf_floats) { s = flush(); if (s != APD_OK) return s; }
            }
        }
        s = flush();
        if (s != APD_OK) return s;
    }
    s = upload_tables(c, src_sorted.data());
    if (s != APD_OK) return s;
    APD_CUDA(c, cudaMemsetAsync(c->d_arena, 0, (size_t)ar.total_frames * ar.dpad * sizeof(float), c->stream));
    if (n) {
        cudaError_t e = ae_encode_launch(c->d_raw, c->d_srcoff, c->d_off, c->d_len, n, n_bins, n_latent, ar.dpad, c->d_aux,
                                         c->d_aux + (size_t)n_bins * n_latent, c->d_arena, c->sm_count, c->stream);
        if (e != cudaSuccess) return fail(c, APD_ERR_CUDA, std::string("ae_encode_kernel: ") + cudaGetErrorString(e));
    }
    APD_CUDA(c, cudaEventRecord(c->ev_h1, c->stream));
    c->timed_h2d = true;
    set_sequence_stats(c, n, extent * sizeof(float));
    s = broadcast_arena(c);
    if (s != APD_OK) return s;
    APD_CUDA(c, cudaStreamSynchronize(c->stream));  // src_sorted goes out of scope
    c->have_sequences = true;
    return APD_OK;
}

apd_status apd_set_sequences_layout(apd_ctx* c, const uint32_t* lens, uint32_t n, uint32_t dim)
{
    apd_status s = begin_sequences(c, lens, n, dim);
    if (s != APD_OK) return s;
    if (grouped(c)) return fail(c, APD_ERR_UNSUPPORTED, "a device group uploads once and copies over NVLink itself");
    s = upload_tables(c, nullptr);
    if (s != APD_OK) return s;
    set_sequence_stats(c, n, 0);
    c->arena_pending = true;   // the caller fills the arena (apd_arena_device) and then calls apd_arena_commit
    return guard_end(c, c->stream);
}

apd_status apd_arena_device(apd_ctx* c, void** d_arena, uint64_t* n_floats)
{
    if (!c || !d_arena || !n_floats) return APD_ERR_INVALID;
    if (c->lead != c || grouped(c)) return fail(c, APD_ERR_UNSUPPORTED, "one-device contexts only");
    if (!c->have_sequences && !c->arena_pending) return fail(c, APD_ERR_STATE, "no sequence batch declared");
    *d_arena = c->d_arena;
    *n_floats = (uint64_t)c->arena.total_frames * c->arena.dpad;
    return APD_OK;
}

apd_status apd_arena_commit(apd_ctx* c)
{
    if (!c) return APD_ERR_INVALID;
    if (!c->arena_pending) return fail(c, APD_ERR_STATE, "apd_set_sequences_layout has not been called");
    c->arena_pending = false;
    c->have_sequences = true;
    return APD_OK;
}

apd_status apd_get_sequence(apd_ctx* c, uint32_t index, float* out, uint64_t cap_floats)
{
    if (!c) return APD_ERR_INVALID;
    if (c->lead != c) return APD_ERR_INVALID;
    if (!c->have_sequences) return fail(c, APD_ERR_STATE, "apd_set_sequences has not been called");
    const Arena& ar = c->arena;
    if (index >= ar.n) return fail(c, APD_ERR_INVALID, "sequence index out of range");
    uint32_t spos = 0;
    for (; spos < ar.n; spos++) if (ar.perm[spos] == index) break;
    const uint64_t need = (uint64_t)ar.len[spos] * ar.dim;
    if (need > cap_floats) return fail(c, APD_ERR_INVALID, "output buffer too small");
    if (!need) return APD_OK;
    if (!out) return fail(c, APD_ERR_INVALID, "out is NULL");
    APD_CUDA(c, cudaSetDevice(c->device));
    APD_CUDA(c, cudaMemcpy2DAsync(out, ar.dim * sizeof(float), c->d_arena + (size_t)ar.off[spos] * ar.dpad, ar.dpad * sizeof(float),
                                  ar.dim * sizeof(float), ar.len[spos], cudaMemcpyDeviceToHost, c->stream));
    APD_CUDA(c, cudaStreamSynchronize(c->stream));
    return APD_OK;
}

apd_status apd_set_shard(apd_ctx* c, uint32_t rank, uint32_t world)
{
    if (!c) return APD_ERR_INVALID;
    if (grouped(c)) return fail(c, APD_ERR_UNSUPPORTED, "a device group shards internally: apd_set_shard is for one-device contexts");
    if (world == 0 || rank >= world) return fail(c, APD_ERR_INVALID, "need 0 <= rank < world");
    c->rank = rank;
    c->world = world;
    c->cells_ref_valid = false;
    return APD_OK;
}

apd_status apd_packed_len(apd_ctx* c, const apd_params* p, uint64_t* n_floats)
{
    if (!c || !n_floats) return APD_ERR_INVALID;
    if (grouped(c)) return fail(c, APD_ERR_UNSUPPORTED, "device-buffer stages are for one-device contexts");
    if (!c->have_sequences) return fail(c, APD_ERR_STATE, "apd_set_sequences has not been called");
    apd_status s = check_params(c, p);
    if (s != APD_OK) return s;
    APD_CUDA(c, cudaSetDevice(c->device));
    s = ensure_plan(c, p->warping_band_percentage);
    if (s != APD_OK) return s;
    *n_floats = packed_entries(c) * 32 * 2;
    return APD_OK;
}

apd_status apd_align_packed(apd_ctx* c, const apd_params* p, float* d_packed, void* stream)
{
    if (!c) return APD_ERR_INVALID;
    if (grouped(c)) return fail(c, APD_ERR_UNSUPPORTED, "device-buffer stages are for one-device contexts");
    if (!c->have_sequences) return fail(c, APD_ERR_STATE, "apd_set_sequences has not been called");
    apd_status s = check_params(c, p);
    if (s != APD_OK) return s;
    if (!d_packed && c->arena.n >= 2) return fail(c, APD_ERR_INVALID, "d_packed is NULL");
    APD_CUDA(c, cudaSetDevice(c->device));
    s = ensure_plan(c, p->warping_band_percentage);
    if (s != APD_OK) return s;
    cudaStream_t st = stream ? (cudaStream_t)stream : c->stream;
    float* outs[1] = {d_packed};
    return run_dtw(c, p, outs, 1, st);   // ordered behind the uploads on the context's stream by the in-flight guard
}

apd_status apd_scatter_packed(apd_ctx* c, const float* d_gathered, uint32_t world, float* d_out_nxn, void* stream)
{
    if (!c) return APD_ERR_INVALID;
    if (grouped(c)) return fail(c, APD_ERR_UNSUPPORTED, "device-buffer stages are for one-device contexts");
    if (!c->have_sequences || !c->plan_valid) return fail(c, APD_ERR_STATE, "no aligned plan: call apd_align_packed first");
    if (world != c->world && world != 1) return fail(c, APD_ERR_INVALID, "world must equal the shard world (or 1 for this shard only)");
    if (c->arena.n && !d_out_nxn) return fail(c, APD_ERR_INVALID, "d_out_nxn is NULL");
    APD_CUDA(c, cudaSetDevice(c->device));
    cudaStream_t st = stream ? (cudaStream_t)stream : c->stream;
    apd_status s = guard_begin(c, st);
    if (s != APD_OK) return s;
    return run_scatter(c, d_gathered, world, d_out_nxn, st);
}

apd_status apd_synchronize(apd_ctx* c, void* stream)
{
    if (!c) return APD_ERR_INVALID;
    APD_CUDA(c, cudaSetDevice(c->device));
    apd_status s = finish_member(c, stream ? (cudaStream_t)stream : c->stream);
    collect_stats(c->lead);
    return s;
}

apd_status apd_align_all(apd_ctx* c, const apd_params* p, float* out_nxn)
{
    if (!c) return APD_ERR_INVALID;
    if (c->lead != c) return fail(c, APD_ERR_INVALID, "not a context handle returned by apd_create / apd_create_multi");
    if (!c->have_sequences) return fail(c, APD_ERR_STATE, "apd_set_sequences has not been called");
    apd_status s = check_params(c, p);
    if (s != APD_OK) return s;
    const uint64_t N = c->arena.n;
    if (N && !out_nxn) return fail(c, APD_ERR_INVALID, "out_nxn is NULL");
    c->matrix_valid = false;
    PhaseTimer pt;
    s = ensure_plan(c, p->warping_band_percentage);
    if (s != APD_OK) return s;
    pt.lap("align_all: unit plan");
    const uint32_t G = (uint32_t)c->members.size();
    const bool group = G > 1;
    // A shard of a multi-process job (apd_set_shard) expands only its own units; a group gathers all of them.
    const uint32_t gather_ranks = group ? G : 1u;
    const size_t pk = (size_t)packed_entries(c) * 64;   // floats per shard
    for (apd_ctx* m : c->members) {
        APD_CUDA(m, cudaSetDevice(m->device));
        s = guard_begin(m, m->stream);
        if (s != APD_OK) return s;
        s = ensure_device(m, m->d_packed, m->packed_cap, pk * gather_ranks);
        if (s != APD_OK) return s;
        s = ensure_device(m, m->d_matrix, m->matrix_cap, (size_t)(N * N));
        if (s != APD_OK) return s;
    }
    pt.lap("align_all: result buffers");
    // DTW kernels on every member; with peer mappings they store into every member's gathered buffer.
    for (apd_ctx* m : c->members) {
        APD_CUDA(m, cudaSetDevice(m->device));
        float* outs[APD_MAX_GROUP];
        uint32_t n_out = 0;
        if (group && c->p2p) {
            for (apd_ctx* q : c->members) outs[n_out++] = q->d_packed + (size_t)m->rank * pk;
        } else {
            outs[n_out++] = m->d_packed + (group ? (size_t)m->rank * pk : 0);
        }
        s = run_dtw(m, p, outs, n_out, m->stream);
        if (s != APD_OK) return s;
        if (group) APD_CUDA(m, cudaEventRecord(m->ev_dtw, m->stream));
    }
    // Every member's matrix needs every member's shard.
    if (group)
        for (apd_ctx* m : c->members) {
            APD_CUDA(m, cudaSetDevice(m->device));
            for (apd_ctx* q : c->members) {
                if (q == m) continue;
                APD_CUDA(m, cudaStreamWaitEvent(m->stream, q->ev_dtw, 0));
                if (!c->p2p)  // no peer mappings: pull q's shard after its kernels
                    APD_CUDA(m, cudaMemcpyPeerAsync(m->d_packed + (size_t)q->rank * pk, m->device,
                                                    q->d_packed + (size_t)q->rank * pk, q->device, pk * sizeof(float), m->stream));
            }
        }
    for (apd_ctx* m : c->members) {
        APD_CUDA(m, cudaSetDevice(m->device));
        s = run_scatter(m, m->d_packed, gather_ranks, m->d_matrix, m->stream);
        if (s != APD_OK) return s;
    }
    pt.lap("align_all: enqueue");
    // Device -> host: member g copies rows [g*N/G, (g+1)*N/G) of the matrix it assembled, so a group
    // uses G PCIe links at once (one host thread per member: the copy into pageable memory blocks).
    c->stats.d2h_bytes = N * N * sizeof(float);
    std::vector<apd_status> st(G, APD_OK);
    auto d2h = [&](uint32_t g) {
        apd_ctx* m = c->members[g];
        auto body = [&]() -> apd_status {
            APD_CUDA(m, cudaSetDevice(m->device));
            const uint64_t r0 = N * g / G, r1 = N * (g + 1) / G;
            if (m == c) APD_CUDA(m, cudaEventRecord(c->ev_d0, m->stream));
            if (r1 > r0)
                APD_CUDA(m, cudaMemcpyAsync(out_nxn + r0 * N, m->d_matrix + r0 * N, (r1 - r0) * N * sizeof(float),
                                            cudaMemcpyDeviceToHost, m->stream));
            if (m == c) { APD_CUDA(m, cudaEventRecord(c->ev_d1, m->stream)); c->timed_d2h = true; }
            return finish_member(m, m->stream);
        };
        st[g] = body();
    };
    if (!group) {
        d2h(0);
    } else {
        std::vector<std::thread> th;
        for (uint32_t g = 0; g < G; g++) th.emplace_back(d2h, g);
        for (auto& x : th) x.join();
    }
    pt.lap("align_all: kernels + D2H");
    APD_CUDA(c, cudaSetDevice(c->device));
    collect_stats(c);
    for (uint32_t g = 0; g < G; g++)
        if (st[g] != APD_OK) { c->err = c->members[g]->err; return st[g]; }
    c->matrix_valid = true;
    return APD_OK;
}

static apd_status align_pairs_impl(apd_ctx* c, const apd_params* p, long long band_override, const uint32_t* pairs_ij,
                                   uint64_t n_pairs, float* scores, uint32_t* paths_ij, uint64_t path_cap,
                                   uint64_t* path_lens)
{
    if (!c) return APD_ERR_INVALID;
    if (c->lead != c) return APD_ERR_INVALID;
    if (!c->have_sequences) return fail(c, APD_ERR_STATE, "apd_set_sequences has not been called");
    apd_status s = check_params(c, p);
    if (s != APD_OK) return s;
    if (n_pairs == 0) return APD_OK;
    if (!pairs_ij || !scores) return fail(c, APD_ERR_INVALID, "pairs_ij/scores is NULL");
    for (uint64_t k = 0; k < 2 * n_pairs; k++)
        if (pairs_ij[k] >= c->arena.n) return fail(c, APD_ERR_INVALID, "pair index out of range");
    // Requested pairs are dealt round-robin to the members of a group (every member holds the arena).
    const uint32_t G = (uint32_t)c->members.size();
    std::vector<apd_status> st(G, APD_OK);
    std::vector<float> path_ms(G, 0.f);
    auto run = [&](uint32_t g) {
        apd_ctx* m = c->members[g];
        auto body = [&]() -> apd_status {
            const uint64_t k0 = n_pairs * g / G, k1 = n_pairs * (g + 1) / G;
            if (k1 <= k0) return APD_OK;
            APD_CUDA(m, cudaSetDevice(m->device));
            apd_status gs = guard_begin(m, m->stream);
            if (gs != APD_OK) return gs;
            std::string err;
            cudaError_t e = pair_paths_run(c->arena, m->d_arena, m->d_off, m->d_len, pairs_ij + 2 * k0, k1 - k0,
                                           p->warping_band_percentage, band_override, p->insertion_penalty, p->deletion_penalty,
                                           p->match_penalty, p->mode == APD_MODE_STRICT, scores + k0,
                                           paths_ij ? paths_ij + k0 * path_cap * 2 : nullptr, path_cap,
                                           path_lens ? path_lens + k0 : nullptr, m->sm_count, m->stream, m->path_scratch, &path_ms[g], err);
            if (e != cudaSuccess) return fail(m, APD_ERR_CUDA, err.empty() ? cudaGetErrorString(e) : err);
            if (!err.empty()) return fail(m, APD_ERR_INVALID, err);
            return APD_OK;
        };
        st[g] = body();
    };
    if (G == 1 || n_pairs < 2 * G) {
        // a handful of pairs: the leader alone
        apd_ctx* m = c;
        APD_CUDA(m, cudaSetDevice(m->device));
        s = guard_begin(m, m->stream);
        if (s != APD_OK) return s;
        std::string err;
        cudaError_t e = pair_paths_run(c->arena, m->d_arena, m->d_off, m->d_len, pairs_ij, n_pairs, p->warping_band_percentage,
                                       band_override, p->insertion_penalty, p->deletion_penalty, p->match_penalty,
                                       p->mode == APD_MODE_STRICT, scores, paths_ij, path_cap, path_lens, m->sm_count, m->stream,
                                       m->path_scratch, &path_ms[0], err);
        if (e != cudaSuccess) return fail(c, APD_ERR_CUDA, err.empty() ? cudaGetErrorString(e) : err);
        if (!err.empty()) return fail(c, APD_ERR_INVALID, err);
    } else {
        std::vector<std::thread> th;
        for (uint32_t g = 0; g < G; g++) th.emplace_back(run, g);
        for (auto& x : th) x.join();
        APD_CUDA(c, cudaSetDevice(c->device));
        for (uint32_t g = 0; g < G; g++)
            if (st[g] != APD_OK) { c->err = c->members[g]->err; return st[g]; }
    }
    float ms = 0.f;
    for (float v : path_ms) ms = std::max(ms, v);
    c->stats.path_ms = ms;
    return APD_OK;
}

apd_status apd_align_pairs(apd_ctx* c, const apd_params* p, const uint32_t* pairs_ij, uint64_t n_pairs,
                           float* scores, uint32_t* paths_ij, uint64_t path_cap, uint64_t* path_lens)
{
    return align_pairs_impl(c, p, -1, pairs_ij, n_pairs, scores, paths_ij, path_cap, path_lens);
}

apd_status apd_align_pairs_band(apd_ctx* c, const apd_params* p, uint64_t warping_band, const uint32_t* pairs_ij,
                                uint64_t n_pairs, float* scores, uint32_t* paths_ij, uint64_t path_cap,
                                uint64_t* path_lens)
{
    const long long b = warping_band > 1000000000ull ? 1000000000ll : (long long)warping_band;
    return align_pairs_impl(c, p, b, pairs_ij, n_pairs, scores, paths_ij, path_cap, path_lens);
}

apd_status apd_align_pair(apd_ctx* c, const apd_params* p, uint32_t i, uint32_t j, float* score, uint32_t* path_ij,
                          uint64_t path_cap, uint64_t* path_len)
{
    uint32_t pr[2] = {i, j};
    return apd_align_pairs(c, p, pr, 1, score, path_ij, path_cap, path_len);
}

static apd_status percentile_impl(apd_ctx* c, const float* d_x, uint64_t len, float perc, cudaStream_t st, float* out)
{
    if (!out) return fail(c, APD_ERR_INVALID, "out is NULL");
    if (len == 0) return fail(c, APD_ERR_INVALID, "percentile of an empty slice: the reference panics (index out of bounds)");
    apd_status s = guard_begin(c, st);
    if (s != APD_OK) return s;
    std::string err;
    uint64_t valid = 0;
    cudaError_t e = percentile_select(d_x, len, perc, c->d_hist, c->sm_count, st, out, &valid, &c->lead->stats.select_ms, err);
    if (e != cudaSuccess) return fail(c, APD_ERR_CUDA, cudaGetErrorString(e));
    if (!err.empty()) return fail(c, APD_ERR_INVALID, err);
    return APD_OK;
}

apd_status apd_percentile_matrix(apd_ctx* c, float perc, float* out)
{
    if (!c) return APD_ERR_INVALID;
    if (!c->matrix_valid) return fail(c, APD_ERR_STATE, "no matrix on the device: call apd_align_all first");
    APD_CUDA(c, cudaSetDevice(c->device));
    const uint64_t N = c->arena.n;
    return percentile_impl(c, c->d_matrix, N * N, perc, c->stream, out);
}

apd_status apd_percentile_device(apd_ctx* c, const float* d_x, uint64_t len, float perc, void* stream, float* out)
{
    if (!c) return APD_ERR_INVALID;
    if (!d_x && len) return fail(c, APD_ERR_INVALID, "d_x is NULL");
    APD_CUDA(c, cudaSetDevice(c->device));
    return percentile_impl(c, d_x, len, perc, stream ? (cudaStream_t)stream : c->stream, out);
}

const char* apd_last_launch_plan(apd_ctx* c)
{
    if (!c) return "";
    // device 0 of the group describes the launch (every member launches the same classes)
    static thread_local std::string buf;
    buf = "[" + c->lead->launch_desc + "]";
    return buf.c_str();
}

apd_status apd_get_stats(apd_ctx* c, apd_stats* out)
{
    if (!c || !out) return APD_ERR_INVALID;
    if (c->have_sequences && c->plan_valid && !c->cells_ref_valid) {
        // a group covers every unit; a multi-process shard only its own
        const bool group = c->members.size() > 1;
        c->cells_ref = reference_cells(c->arena, c->plan, group ? 0 : c->rank, group ? 1 : c->world);
        c->cells_ref_valid = true;
    }
    c->stats.cells_reference = c->cells_ref_valid ? c->cells_ref : 0;
    *out = c->stats;
    return APD_OK;
}

}  // extern "C"
