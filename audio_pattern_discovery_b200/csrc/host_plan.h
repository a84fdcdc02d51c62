// host_plan.h -- host-side planning for the all-pairs DTW launch: the sequence
// arena (the spectrogram.rs/discovery.rs "packing glue" of SURVEY.md section 8 row a1)
// and the 32-pair work-unit list with its launch classes.  Pure C++ (no CUDA) so the
// CPU test-suite can exercise it together with dtw_core.h.
#pragma once
#include <stdint.h>

#include <string>
#include <vector>

namespace apd {

// Sequences sorted by length (stable), each stored as
//   [PRE_PAD_FRAMES zero frames][len frames of dpad floats, zero padded from dim]
// so that every frame starts on a 16-byte boundary and the end-anchored tiles of
// dtw_core.h may read up to 4 frames in front of a sequence.
struct Arena {
    uint32_t n = 0, dim = 0, dpad = 0;
    std::vector<uint32_t> perm;   // sorted position -> caller's index
    std::vector<uint32_t> len;    // length at sorted position
    std::vector<uint32_t> off;    // frame index (in the arena) of frame 0, sorted position
    uint64_t total_frames = 0;
    std::vector<float> data;      // total_frames * dpad floats (host staging copy)
};

// Computes the layout (perm / len / off / total_frames); `data` stays empty.
// Returns "" or an error message.
std::string build_arena_layout(const uint32_t* lens, uint32_t n, uint32_t dim, Arena& out);

// Fills dst (total_frames * dpad floats, e.g. a pinned staging buffer) from per-sequence
// pointers: frames[s] -> lens[s] * dim floats in the caller's order.
void fill_arena(const Arena& layout, const float* const* frames, float* dst);

// Frames [F0, F1) of the arena only: dst[0] is the first float of frame F0 (chunked upload).
void fill_arena_frames(const Arena& layout, const float* const* frames, uint64_t F0, uint64_t F1, float* dst);

// Layout + fill into out.data (used by the host-side emulator).
std::string build_arena(const float* const* frames, const uint32_t* lens, uint32_t n,
                        uint32_t dim, Arena& out);

struct Unit {
    uint32_t a;  // sorted position of the shared row sequence
    uint32_t B;  // lanes own sorted positions 32*B .. 32*B+31 (only those > a are pairs)
};

// A launch class: a contiguous slice of the ordered unit list whose units all fit a
// boundary ring of `St` tiles; `gstate` classes keep the ring in global memory.
struct UnitClass {
    uint64_t begin = 0, end = 0;
    int St = 0;
    bool gstate = false;
};

struct UnitPlan {
    float pct = 0.f;
    std::vector<Unit> units;        // ordered: class by class, expensive units first
    std::vector<UnitClass> classes;
    uint64_t tiles_estimate = 0;    // sum of the per-unit cost estimates (tiles)
    uint32_t row_block = 32;        // enumeration granularity the list was built with (see build_unit_plan)
};

// Largest ring (in tiles) kept in shared memory; bigger units go to the gstate class.
enum { SMEM_RING_CAPS = 3 };
extern const int kSmemRingCaps[SMEM_RING_CAPS];

// row_block: 32 x the number of devices / ranks that deal the list among themselves (u mod world).
void build_unit_plan(const Arena& arena, float pct, UnitPlan& out, uint32_t row_block = 32);

// Reference cell updates (src/alignments.rs:174-175 visit rule) of one ordered pair.
uint64_t cells_visited(uint64_t n, uint64_t m, uint64_t w);

// Sum of cells_visited over both orientations of every pair of the units
// u in [0, units.size()) with u % world == rank.
uint64_t reference_cells(const Arena& arena, const UnitPlan& plan, uint32_t rank, uint32_t world);

}  // namespace apd
