// apd_api.cu -- the C ABI of include/apd.h on top of the sm_100a kernels.
//
// There is no CPU path in this file: every entry point that computes anything needs
// a CUDA device, and apd_create() fails with APD_ERR_NO_DEVICE without one.
#include <cuda_runtime.h>

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <new>
#include <string>
#include <thread>
#include <vector>

#include "../../include/apd.h"
#include "dtw_kernels.cuh"
#include "host_plan.h"
#include "pair_path.cuh"
#include "percentile.cuh"

using namespace apd;

// The layouts the ctypes / Rust / C++ bindings mirror (tests/test_abi.py checks the Python side).
static_assert(sizeof(apd_params) == 20, "apd_params layout is part of the ABI");
static_assert(sizeof(apd_stats) == 104, "apd_stats layout is part of the ABI");
static_assert(sizeof(apd_merge) == 24, "apd_merge layout is part of the ABI");

namespace {

thread_local std::string g_create_error;

struct OccKey { int dpad, strict, unitw, ring; size_t smem; int occ; };

}  // namespace

struct apd_ctx {
    int device = 0;
    cudaStream_t stream = nullptr;
    int sm_count = 0;
    float sm_clock_mhz = 0.f;
    size_t smem_optin = 0;

    Arena arena;  // layout only (data lives on the device)
    bool have_sequences = false;
    float* d_arena = nullptr; size_t arena_cap = 0;
    uint32_t* d_off = nullptr; uint32_t* d_len = nullptr; uint32_t* d_perm = nullptr;
    uint64_t* d_srcoff = nullptr; size_t table_cap = 0;
    float* h_stage = nullptr; size_t stage_cap = 0;  // pinned
    float* d_raw = nullptr; size_t raw_cap = 0;

    uint32_t rank = 0, world = 1;

    UnitPlan plan; bool plan_valid = false;
    Unit* d_units = nullptr; size_t units_cap = 0;
    uint64_t cells_ref = 0; bool cells_ref_valid = false;

    float* d_packed = nullptr; size_t packed_cap = 0;   // own packed buffer (apd_align_all)
    float* d_matrix = nullptr; size_t matrix_cap = 0;
    float2* d_gstate = nullptr; size_t gstate_cap = 0;
    unsigned int* d_counters = nullptr;                 // one work counter per class
    int* d_error = nullptr;
    unsigned long long* d_tiles = nullptr;
    unsigned long long* d_hist = nullptr;               // 256 radix-select counters
    bool matrix_valid = false;                          // d_matrix holds the last apd_align_all result

    cudaEvent_t ev_k0 = nullptr, ev_k1 = nullptr, ev_s0 = nullptr, ev_s1 = nullptr;
    cudaEvent_t ev_h0 = nullptr, ev_h1 = nullptr, ev_d0 = nullptr, ev_d1 = nullptr;
    bool timed_kernel = false, timed_scatter = false, timed_h2d = false, timed_d2h = false;

    std::vector<OccKey> occ_cache;
    apd_stats stats{};
    std::string err;
};

namespace {

apd_status fail(apd_ctx* c, apd_status s, const std::string& msg)
{
    if (c) c->err = msg; else g_create_error = msg;
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

apd_status upload_tables(apd_ctx* c, const uint64_t* src_off_sorted)
{
    const Arena& ar = c->arena;
    const size_t n = ar.n;
    if (n > c->table_cap || !c->d_off) {
        if (c->d_off) cudaFree(c->d_off);
        if (c->d_len) cudaFree(c->d_len);
        if (c->d_perm) cudaFree(c->d_perm);
        if (c->d_srcoff) cudaFree(c->d_srcoff);
        c->d_off = c->d_len = c->d_perm = nullptr; c->d_srcoff = nullptr; c->table_cap = 0;
        size_t cap = std::max<size_t>(n, 1);
        APD_CUDA(c, cudaMalloc((void**)&c->d_off, cap * sizeof(uint32_t)));
        APD_CUDA(c, cudaMalloc((void**)&c->d_len, cap * sizeof(uint32_t)));
        APD_CUDA(c, cudaMalloc((void**)&c->d_perm, cap * sizeof(uint32_t)));
        APD_CUDA(c, cudaMalloc((void**)&c->d_srcoff, cap * sizeof(uint64_t)));
        c->table_cap = cap;
    }
    if (n) {
        APD_CUDA(c, cudaMemcpyAsync(c->d_off, ar.off.data(), n * sizeof(uint32_t), cudaMemcpyHostToDevice, c->stream));
        APD_CUDA(c, cudaMemcpyAsync(c->d_len, ar.len.data(), n * sizeof(uint32_t), cudaMemcpyHostToDevice, c->stream));
        APD_CUDA(c, cudaMemcpyAsync(c->d_perm, ar.perm.data(), n * sizeof(uint32_t), cudaMemcpyHostToDevice, c->stream));
        if (src_off_sorted)
            APD_CUDA(c, cudaMemcpyAsync(c->d_srcoff, src_off_sorted, n * sizeof(uint64_t), cudaMemcpyHostToDevice, c->stream));
    }
    return APD_OK;
}

apd_status begin_sequences(apd_ctx* c, const uint32_t* lens, uint32_t n, uint32_t dim)
{
    if (!c) return APD_ERR_INVALID;
    if (dim == 0) return fail(c, APD_ERR_INVALID, "dim must be >= 1");
    if (dim > APD_MAX_DIM) return fail(c, APD_ERR_UNSUPPORTED, "dim > APD_MAX_DIM (32) is not supported");
    if (n > 0 && !lens) return fail(c, APD_ERR_INVALID, "lens is NULL");
    APD_CUDA(c, cudaSetDevice(c->device));
    const bool had = c->have_sequences;
    c->have_sequences = false;
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
    return ensure_device(c, c->d_arena, c->arena_cap, (size_t)c->arena.total_frames * c->arena.dpad);
}

void k_range(uint64_t begin, uint64_t end, uint32_t rank, uint32_t world, uint64_t& k0, uint64_t& k1)
{
    // smallest k with rank + world*k >= x
    auto first_k = [&](uint64_t x) -> uint64_t { return x <= rank ? 0 : (x - rank + world - 1) / world; };
    k0 = first_k(begin);
    k1 = first_k(end);
}

uint64_t packed_entries(const apd_ctx* c)
{
    uint64_t U = c->plan.units.size();
    return (U + c->world - 1) / c->world;  // identical on every rank
}

apd_status ensure_plan(apd_ctx* c, float pct)
{
    // The plan depends on pct only through the band (bit pattern compare keeps NaN stable).
    if (c->plan_valid && std::memcmp(&c->plan.pct, &pct, sizeof(float)) == 0) return APD_OK;
    build_unit_plan(c->arena, pct, c->plan);
    c->cells_ref_valid = false;
    apd_status s = ensure_device(c, c->d_units, c->units_cap, c->plan.units.size());
    if (s != APD_OK) return s;
    if (!c->plan.units.empty())
        APD_CUDA(c, cudaMemcpyAsync(c->d_units, c->plan.units.data(), c->plan.units.size() * sizeof(Unit),
                                    cudaMemcpyHostToDevice, c->stream));
    // the host vector must outlive the async copy from pageable memory: it does (owned by ctx),
    // and cudaMemcpyAsync from pageable memory returns only after staging the source.
    c->plan_valid = true;
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

apd_status run_dtw(apd_ctx* c, const apd_params* p, float* d_packed, cudaStream_t stream)
{
    LaunchFns f;
    if (!pick_launcher(c->arena.dpad, f)) return fail(c, APD_ERR_UNSUPPORTED, "unsupported frame width");
    const bool strict = (p->mode == APD_MODE_STRICT);
    const bool unitw = (p->insertion_penalty == 1.0f && p->deletion_penalty == 1.0f && p->match_penalty == 1.0f);
    // Where the boundary ring lives (dtw_kernels.cuh): tensor memory if it fits 256 columns,
    // else shared memory, else global scratch.  APD_RING=tmem|smem|global narrows the choice
    // for experiments and tests (a ring that does not fit the forced home moves down the list).
    const char* ring_env = getenv("APD_RING");
    const char* force_g = getenv("APD_FORCE_GSTATE");
    int ring_floor = RING_TMEM;  // best allowed: TMEM > SMEM > GLOBAL
    if (ring_env && !strcmp(ring_env, "smem")) ring_floor = RING_SMEM;
    if ((ring_env && !strcmp(ring_env, "global")) || (force_g && force_g[0] == '1')) ring_floor = RING_GLOBAL;

    APD_CUDA(c, cudaMemsetAsync(c->d_counters, 0, 8 * sizeof(unsigned int), stream));
    APD_CUDA(c, cudaMemsetAsync(c->d_error, 0, sizeof(int), stream));
    APD_CUDA(c, cudaMemsetAsync(c->d_tiles, 0, sizeof(unsigned long long), stream));
    APD_CUDA(c, cudaEventRecord(c->ev_k0, stream));
    uint32_t launches = 0;
    uint64_t local_units = 0;
    for (size_t ci = 0; ci < c->plan.classes.size(); ci++) {
        const UnitClass& uc = c->plan.classes[ci];
        uint64_t k0, k1;
        k_range(uc.begin, uc.end, c->rank, c->world, k0, k1);
        if (k1 <= k0) continue;
        local_units += k1 - k0;
        int ring = RING_GLOBAL;
        if (ring_floor == RING_TMEM && uc.St <= TMEM_RING_TILES) ring = RING_TMEM;
        else if (ring_floor != RING_GLOBAL && !uc.gstate) ring = RING_SMEM;
        const size_t smem = dtw_smem_bytes((int)c->arena.dpad, uc.St, ring);
        if (smem > c->smem_optin) return fail(c, APD_ERR_INTERNAL, "ring does not fit in shared memory");
        int occ = 0;
        apd_status s = occupancy(c, f, (int)c->arena.dpad, strict, unitw, ring, smem, occ);
        if (s != APD_OK) return s;
        const uint64_t warps_per_cta = (uint64_t)dtw_cta_warps(ring);
        uint64_t grid64 = std::min<uint64_t>((k1 - k0 + warps_per_cta - 1) / warps_per_cta, (uint64_t)c->sm_count * occ);
        int grid = (int)grid64;
        KernelArgs a{};
        a.arena = c->d_arena; a.off = c->d_off; a.len = c->d_len; a.units = c->d_units;
        a.N = c->arena.n; a.rank = c->rank; a.world = c->world;
        a.k_begin = k0; a.k_count = (uint32_t)(k1 - k0);
        a.counter = c->d_counters + ci;
        a.pct = p->warping_band_percentage;
        a.pen.ins = p->insertion_penalty; a.pen.del = p->deletion_penalty; a.pen.mat = p->match_penalty;
        a.St = uc.St;
        a.out = reinterpret_cast<float2*>(d_packed);
        a.error_flag = c->d_error;
        a.tiles_done = c->d_tiles;
        if (ring == RING_GLOBAL) {
            size_t need = (size_t)grid * uc.St * TILE * 32;
            s = ensure_device(c, c->d_gstate, c->gstate_cap, need);
            if (s != APD_OK) return s;
            a.gstate = c->d_gstate;
        }
        if (getenv("APD_DEBUG"))
            fprintf(stderr, "[apd] class %zu: ring=%s St=%d units=%llu grid=%d x %d warps occ=%d smem=%zu\n", ci,
                    ring == RING_TMEM ? "tmem" : (ring == RING_SMEM ? "smem" : "global"), uc.St,
                    (unsigned long long)(k1 - k0), grid, (int)warps_per_cta, occ, smem);
        APD_CUDA(c, f.launch(a, strict, unitw, ring, grid, smem, stream));
        launches++;
    }
    APD_CUDA(c, cudaEventRecord(c->ev_k1, stream));
    c->timed_kernel = true;
    c->stats.kernel_launches = launches;
    c->stats.units_local = local_units;
    c->stats.units_total = c->plan.units.size();
    return APD_OK;
}

apd_status run_scatter(apd_ctx* c, const float* d_gathered, uint32_t nranks, float* d_out, cudaStream_t stream)
{
    const uint64_t N = c->arena.n;
    APD_CUDA(c, cudaEventRecord(c->ev_s0, stream));
    if (N) APD_CUDA(c, cudaMemsetAsync(d_out, 0, N * N * sizeof(float), stream));
    const uint64_t kpr = packed_entries(c);
    const uint64_t threads = (uint64_t)nranks * kpr * 32;
    if (threads) {
        const int bs = 256;
        const uint64_t blocks = (threads + bs - 1) / bs;
        if (blocks > 0x7fffffffull) return fail(c, APD_ERR_UNSUPPORTED, "matrix too large for one scatter launch");
        // nranks == world: buffer holds all ranks (rank-major).  nranks == 1 on a sharded
        // context: buffer holds only this rank's units.
        const uint32_t world = c->world;
        if (nranks == world) {
            scatter_packed_kernel<<<(unsigned)blocks, bs, 0, stream>>>(
                reinterpret_cast<const float2*>(d_gathered), kpr, world, 0, c->d_units,
                c->plan.units.size(), c->d_perm, (uint32_t)N, d_out);
        } else {
            scatter_packed_kernel<<<(unsigned)blocks, bs, 0, stream>>>(
                reinterpret_cast<const float2*>(d_gathered), kpr, world, c->rank, c->d_units,
                c->plan.units.size(), c->d_perm, (uint32_t)N, d_out);
        }
        APD_CUDA(c, cudaGetLastError());
        c->stats.kernel_launches++;
    }
    APD_CUDA(c, cudaEventRecord(c->ev_s1, stream));
    c->timed_scatter = true;
    return APD_OK;
}

apd_status finish(apd_ctx* c, cudaStream_t stream)
{
    APD_CUDA(c, cudaStreamSynchronize(stream));
    int herr = 0;
    unsigned long long tiles = 0;
    APD_CUDA(c, cudaMemcpy(&herr, c->d_error, sizeof(int), cudaMemcpyDeviceToHost));
    APD_CUDA(c, cudaMemcpy(&tiles, c->d_tiles, sizeof(tiles), cudaMemcpyDeviceToHost));
    c->stats.cells_computed = (uint64_t)tiles * TILE * TILE * 2;
    if (c->timed_kernel) { cudaEventElapsedTime(&c->stats.kernel_ms, c->ev_k0, c->ev_k1); c->timed_kernel = false; }
    if (c->timed_scatter) { cudaEventElapsedTime(&c->stats.scatter_ms, c->ev_s0, c->ev_s1); c->timed_scatter = false; }
    if (c->timed_h2d) { cudaEventElapsedTime(&c->stats.h2d_ms, c->ev_h0, c->ev_h1); c->timed_h2d = false; }
    if (c->timed_d2h) { cudaEventElapsedTime(&c->stats.d2h_ms, c->ev_d0, c->ev_d1); c->timed_d2h = false; }
    if (herr) return fail(c, APD_ERR_INTERNAL, "a work unit needed a larger boundary ring than planned");
    return APD_OK;
}

}  // namespace

extern "C" {

uint32_t apd_abi_version(void) { return APD_ABI_VERSION; }

const char* apd_last_error(const apd_ctx* ctx) { return ctx ? ctx->err.c_str() : g_create_error.c_str(); }

apd_status apd_create(int device_id, apd_ctx** out)
{
    if (!out) return fail(nullptr, APD_ERR_INVALID, "out is NULL");
    *out = nullptr;
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess || count == 0)
        return fail(nullptr, APD_ERR_NO_DEVICE,
                    std::string("no CUDA device (this library has no CPU path): ") +
                        (e != cudaSuccess ? cudaGetErrorString(e) : "device count is 0"));
    if (device_id < 0 || device_id >= count) return fail(nullptr, APD_ERR_INVALID, "device_id out of range");
    apd_ctx* c = new (std::nothrow) apd_ctx();
    if (!c) return fail(nullptr, APD_ERR_INTERNAL, "out of host memory");
    c->device = device_id;
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
    APD_CREATE_CUDA(cudaSetDevice(device_id));
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
    APD_CREATE_CUDA(cudaMalloc((void**)&c->d_counters, 8 * sizeof(unsigned int)));
    APD_CREATE_CUDA(cudaMalloc((void**)&c->d_error, sizeof(int)));
    APD_CREATE_CUDA(cudaMalloc((void**)&c->d_tiles, sizeof(unsigned long long)));
    APD_CREATE_CUDA(cudaMalloc((void**)&c->d_hist, 256 * sizeof(unsigned long long)));
    cudaEvent_t* evs[] = {&c->ev_k0, &c->ev_k1, &c->ev_s0, &c->ev_s1, &c->ev_h0, &c->ev_h1, &c->ev_d0, &c->ev_d1};
    for (cudaEvent_t* ev : evs) APD_CREATE_CUDA(cudaEventCreate(ev));
#undef APD_CREATE_CUDA
    c->stats.sm_clock_mhz = c->sm_clock_mhz;
    c->stats.sm_count = (uint32_t)c->sm_count;
    *out = c;
    return APD_OK;
}

void apd_destroy(apd_ctx* c)
{
    if (!c) return;
    cudaSetDevice(c->device);
    if (c->stream) cudaStreamSynchronize(c->stream);
    void* dptrs[] = {c->d_arena, c->d_off, c->d_len, c->d_perm, c->d_srcoff, c->d_raw, c->d_units, c->d_packed,
                     c->d_matrix, c->d_gstate, c->d_counters, c->d_error, c->d_tiles, c->d_hist};
    for (void* p : dptrs) if (p) cudaFree(p);
    if (c->h_stage) cudaFreeHost(c->h_stage);
    cudaEvent_t evs[] = {c->ev_k0, c->ev_k1, c->ev_s0, c->ev_s1, c->ev_h0, c->ev_h1, c->ev_d0, c->ev_d1};
    for (cudaEvent_t ev : evs) if (ev) cudaEventDestroy(ev);
    if (c->stream) cudaStreamDestroy(c->stream);
    delete c;
}

apd_status apd_set_sequences(apd_ctx* c, const float* const* frames, const uint32_t* lens, uint32_t n, uint32_t dim)
{
    apd_status s = begin_sequences(c, lens, n, dim);
    if (s != APD_OK) return s;
    if (n > 0 && !frames) return fail(c, APD_ERR_INVALID, "frames is NULL");
    for (uint32_t k = 0; k < n; k++)
        if (lens[k] > 0 && !frames[k]) return fail(c, APD_ERR_INVALID, "frames[k] is NULL for a non-empty sequence");
    const size_t floats = (size_t)c->arena.total_frames * c->arena.dpad;
    if (floats > c->stage_cap) {
        if (c->h_stage) cudaFreeHost(c->h_stage);
        c->h_stage = nullptr; c->stage_cap = 0;
        APD_CUDA(c, cudaMallocHost((void**)&c->h_stage, floats * sizeof(float)));
        c->stage_cap = floats;
    }
    // CPU half of the packing glue: sort + pad into the pinned staging buffer.
    fill_arena(c->arena, frames, c->h_stage);
    APD_CUDA(c, cudaEventRecord(c->ev_h0, c->stream));
    APD_CUDA(c, cudaMemcpyAsync(c->d_arena, c->h_stage, floats * sizeof(float), cudaMemcpyHostToDevice, c->stream));
    s = upload_tables(c, nullptr);
    if (s != APD_OK) return s;
    APD_CUDA(c, cudaEventRecord(c->ev_h1, c->stream));
    c->timed_h2d = true;
    c->stats.h2d_bytes = floats * sizeof(float);
    c->stats.n_sequences = n;
    c->stats.ordered_pairs = (uint64_t)n * (n ? n - 1 : 0);
    APD_CUDA(c, cudaStreamSynchronize(c->stream));
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
    APD_CUDA(c, cudaEventRecord(c->ev_h1, c->stream));
    c->timed_h2d = true;
    // Device half of the packing glue: zero the arena (pads), then gather.
    APD_CUDA(c, cudaMemsetAsync(c->d_arena, 0, (size_t)c->arena.total_frames * c->arena.dpad * sizeof(float), c->stream));
    if (n) {
        int grid = (int)std::min<uint32_t>(n, (uint32_t)c->sm_count * 16);
        pack_arena_kernel<<<grid, 256, 0, c->stream>>>(c->d_raw, c->d_srcoff, c->d_off, c->d_len, n, dim,
                                                       c->arena.dpad, c->d_arena);
        APD_CUDA(c, cudaGetLastError());
    }
    c->stats.h2d_bytes = extent * sizeof(float);
    c->stats.n_sequences = n;
    c->stats.ordered_pairs = (uint64_t)n * (n ? n - 1 : 0);
    APD_CUDA(c, cudaStreamSynchronize(c->stream));  // src_sorted and the caller's buffer may go away
    c->have_sequences = true;
    return APD_OK;
}

apd_status apd_set_shard(apd_ctx* c, uint32_t rank, uint32_t world)
{
    if (!c) return APD_ERR_INVALID;
    if (world == 0 || rank >= world) return fail(c, APD_ERR_INVALID, "need 0 <= rank < world");
    c->rank = rank;
    c->world = world;
    c->cells_ref_valid = false;
    return APD_OK;
}

apd_status apd_packed_len(apd_ctx* c, const apd_params* p, uint64_t* n_floats)
{
    if (!c || !n_floats) return APD_ERR_INVALID;
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
    if (!c->have_sequences) return fail(c, APD_ERR_STATE, "apd_set_sequences has not been called");
    apd_status s = check_params(c, p);
    if (s != APD_OK) return s;
    if (!d_packed && c->arena.n >= 2) return fail(c, APD_ERR_INVALID, "d_packed is NULL");
    APD_CUDA(c, cudaSetDevice(c->device));
    s = ensure_plan(c, p->warping_band_percentage);
    if (s != APD_OK) return s;
    cudaStream_t st = stream ? (cudaStream_t)stream : c->stream;
    if (st != c->stream) APD_CUDA(c, cudaStreamSynchronize(c->stream));  // plan upload happens on ctx->stream
    return run_dtw(c, p, d_packed, st);
}

apd_status apd_scatter_packed(apd_ctx* c, const float* d_gathered, uint32_t world, float* d_out_nxn, void* stream)
{
    if (!c) return APD_ERR_INVALID;
    if (!c->have_sequences || !c->plan_valid) return fail(c, APD_ERR_STATE, "no aligned plan: call apd_align_packed first");
    if (world != c->world && world != 1) return fail(c, APD_ERR_INVALID, "world must equal the shard world (or 1 for this shard only)");
    if (c->arena.n && !d_out_nxn) return fail(c, APD_ERR_INVALID, "d_out_nxn is NULL");
    APD_CUDA(c, cudaSetDevice(c->device));
    cudaStream_t st = stream ? (cudaStream_t)stream : c->stream;
    return run_scatter(c, d_gathered, world, d_out_nxn, st);
}

apd_status apd_synchronize(apd_ctx* c, void* stream)
{
    if (!c) return APD_ERR_INVALID;
    APD_CUDA(c, cudaSetDevice(c->device));
    return finish(c, stream ? (cudaStream_t)stream : c->stream);
}

apd_status apd_align_all(apd_ctx* c, const apd_params* p, float* out_nxn)
{
    if (!c) return APD_ERR_INVALID;
    if (!c->have_sequences) return fail(c, APD_ERR_STATE, "apd_set_sequences has not been called");
    apd_status s = check_params(c, p);
    if (s != APD_OK) return s;
    const uint64_t N = c->arena.n;
    if (N && !out_nxn) return fail(c, APD_ERR_INVALID, "out_nxn is NULL");
    APD_CUDA(c, cudaSetDevice(c->device));
    s = ensure_plan(c, p->warping_band_percentage);
    if (s != APD_OK) return s;
    size_t pk = (size_t)packed_entries(c) * 64;
    s = ensure_device(c, c->d_packed, c->packed_cap, pk);
    if (s != APD_OK) return s;
    s = ensure_device(c, c->d_matrix, c->matrix_cap, (size_t)(N * N));
    if (s != APD_OK) return s;
    s = run_dtw(c, p, c->d_packed, c->stream);
    if (s != APD_OK) return s;
    s = run_scatter(c, c->d_packed, 1, c->d_matrix, c->stream);
    if (s != APD_OK) return s;
    APD_CUDA(c, cudaEventRecord(c->ev_d0, c->stream));
    if (N) APD_CUDA(c, cudaMemcpyAsync(out_nxn, c->d_matrix, N * N * sizeof(float), cudaMemcpyDeviceToHost, c->stream));
    APD_CUDA(c, cudaEventRecord(c->ev_d1, c->stream));
    c->timed_d2h = true;
    c->stats.d2h_bytes = N * N * sizeof(float);
    s = finish(c, c->stream);
    c->matrix_valid = (s == APD_OK);
    return s;
}

static apd_status align_pairs_impl(apd_ctx* c, const apd_params* p, long long band_override, const uint32_t* pairs_ij,
                                   uint64_t n_pairs, float* scores, uint32_t* paths_ij, uint64_t path_cap,
                                   uint64_t* path_lens)
{
    if (!c) return APD_ERR_INVALID;
    if (!c->have_sequences) return fail(c, APD_ERR_STATE, "apd_set_sequences has not been called");
    apd_status s = check_params(c, p);
    if (s != APD_OK) return s;
    if (n_pairs == 0) return APD_OK;
    if (!pairs_ij || !scores) return fail(c, APD_ERR_INVALID, "pairs_ij/scores is NULL");
    for (uint64_t k = 0; k < 2 * n_pairs; k++)
        if (pairs_ij[k] >= c->arena.n) return fail(c, APD_ERR_INVALID, "pair index out of range");
    APD_CUDA(c, cudaSetDevice(c->device));
    std::string err;
    cudaError_t e = pair_paths_run(c->arena, c->d_arena, c->d_off, c->d_len, pairs_ij, n_pairs, p->warping_band_percentage,
                                   band_override, p->insertion_penalty, p->deletion_penalty, p->match_penalty,
                                   p->mode == APD_MODE_STRICT, scores, paths_ij, path_cap, path_lens, c->stream, err);
    if (e != cudaSuccess) return fail(c, APD_ERR_CUDA, err.empty() ? cudaGetErrorString(e) : err);
    if (!err.empty()) return fail(c, APD_ERR_INVALID, err);
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
    std::string err;
    uint64_t valid = 0;
    cudaError_t e = percentile_select(d_x, len, perc, c->d_hist, c->sm_count, st, out, &valid, &c->stats.select_ms, err);
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

apd_status apd_get_stats(apd_ctx* c, apd_stats* out)
{
    if (!c || !out) return APD_ERR_INVALID;
    if (c->have_sequences && c->plan_valid && !c->cells_ref_valid) {
        c->cells_ref = reference_cells(c->arena, c->plan, c->rank, c->world);
        c->cells_ref_valid = true;
    }
    c->stats.cells_reference = c->cells_ref_valid ? c->cells_ref : 0;
    *out = c->stats;
    return APD_OK;
}

}  // extern "C"
