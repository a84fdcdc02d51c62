// apd_host.hpp -- C++ host-side mirror of the reference's Rust interface for the DTW hot path,
// on top of the C ABI of apd.h.  The reference is compiled code (Rust) and its toolchain is
// absent from the build image, so this header is the compiled-language drop-in surface: same
// names, argument meaning and failure behaviour (a Rust panic is a thrown apd_host::Panic).
// File:line citations are relative to the reference repository root.
//
//   NDSequence                         src/spectrogram.rs:13-24, vec() 99-101, len() 152-154
//   Discovery, alignment_params()      src/discovery.rs:7-46
//   AlignmentParams, ::Default(len)    src/alignments.rs:77-94
//   AlignmentWorkers, ::align_all      src/alignments.rs:11-68
//   Alignment, ::construct_alignment, ::score    src/alignments.rs:99-181
//   AgglomerativeClustering::{clustering, cluster_sets}, ClusteringOperation, Merge
//                                      src/clustering.rs:7-110
//
// Header only; link with -lapd_b200.  Nothing here computes: every number comes from the
// library (CUDA for the alignment and the threshold, host C++ for the UPGMA).
#pragma once
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <fstream>
#include <limits>
#include <map>
#include <memory>
#include <mutex>
#include <set>
#include <sstream>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>

#include "apd.h"

namespace apd_host {

// A Rust `panic!` / failed `unwrap()` at the same place in the reference.
struct Panic : std::runtime_error {
    explicit Panic(const std::string& what) : std::runtime_error(what) {}
};

// src/spectrogram.rs:13-24 -- only the members the hot path touches.
struct NDSequence {
    size_t n_bins = 0;
    std::vector<float> frames;  // flat `[x00 ... x0D ... xT0 ... xTD]`
    size_t audio_id = 0;

    NDSequence() = default;
    NDSequence(size_t bins, std::vector<float> flat, size_t id = 0) : n_bins(bins), frames(std::move(flat)), audio_id(id) {}
    const float* vec(size_t t) const { return frames.data() + t * n_bins; }  // src/spectrogram.rs:99-101
    size_t len() const { return n_bins ? frames.size() / n_bins : 0; }       // src/spectrogram.rs:152-154
};

// src/alignments.rs:77-94
struct AlignmentParams {
    size_t warping_band = 0;
    float insertion_penalty = 1.0f, deletion_penalty = 1.0f, match_penalty = 1.0f;
    static AlignmentParams Default(size_t len) { return AlignmentParams{len, 1.0f, 1.0f, 1.0f}; }
};

// Rust `f32 as usize`: truncation, saturation, NaN -> 0.
inline size_t f32_as_usize(float v)
{
    if (!(v == v) || v <= 0.0f) return 0;
    if (v >= 18446744073709551616.0f) return std::numeric_limits<size_t>::max();
    return (size_t)v;
}

// src/discovery.rs:7-26 (every field of project/config/Discovery.toml is required, like serde).
struct Discovery {
    size_t dft_win = 256, dft_step = 128, ceps_filter = 32, vat_moving = 15;
    float vat_percentile = 0.95f;
    size_t vat_min_len = 150, alignment_workers = 4;
    float clustering_percentile = 0.05f, warping_band_percentage = 1.0f;
    float insertion_penalty = 1.0f, deletion_penalty = 1.0f, match_penalty = 1.0f;
    size_t auto_encoder = 10;
    float learning_rate = 0.1f;
    size_t epochs = 25;
    float epoch_drop = 5.0f, drop = 0.5f;

    // src/discovery.rs:29-36.  The shipped file is flat `key = value  # comment` lines.
    static Discovery from_toml(const std::string& file)
    {
        std::ifstream in(file);
        if (!in) throw Panic("Template file not found");
        std::map<std::string, std::string> kv;
        std::string line;
        while (std::getline(in, line)) {
            const size_t hash = line.find('#');
            if (hash != std::string::npos) line.erase(hash);
            const size_t eq = line.find('=');
            if (eq == std::string::npos) continue;
            auto trim = [](std::string s) {
                const char* ws = " \t\r\n";
                const size_t a = s.find_first_not_of(ws), b = s.find_last_not_of(ws);
                return a == std::string::npos ? std::string() : s.substr(a, b - a + 1);
            };
            kv[trim(line.substr(0, eq))] = trim(line.substr(eq + 1));
        }
        auto need = [&](const char* k) -> const std::string& {
            auto it = kv.find(k);
            if (it == kv.end()) throw Panic(std::string("missing field `") + k + "`");
            return it->second;
        };
        Discovery d;
        d.dft_win = std::stoul(need("dft_win")); d.dft_step = std::stoul(need("dft_step"));
        d.ceps_filter = std::stoul(need("ceps_filter")); d.vat_moving = std::stoul(need("vat_moving"));
        d.vat_percentile = std::stof(need("vat_percentile")); d.vat_min_len = std::stoul(need("vat_min_len"));
        d.alignment_workers = std::stoul(need("alignment_workers"));
        d.clustering_percentile = std::stof(need("clustering_percentile"));
        d.warping_band_percentage = std::stof(need("warping_band_percentage"));
        d.insertion_penalty = std::stof(need("insertion_penalty")); d.deletion_penalty = std::stof(need("deletion_penalty"));
        d.match_penalty = std::stof(need("match_penalty")); d.auto_encoder = std::stoul(need("auto_encoder"));
        d.learning_rate = std::stof(need("learning_rate")); d.epochs = std::stoul(need("epochs"));
        d.epoch_drop = std::stof(need("epoch_drop")); d.drop = std::stof(need("drop"));
        return d;
    }

    // src/discovery.rs:38-45: the band is an f32 product truncated to usize.
    AlignmentParams alignment_params(size_t n_size) const
    {
        return AlignmentParams{f32_as_usize(warping_band_percentage * (float)n_size), insertion_penalty, deletion_penalty,
                               match_penalty};
    }
};

// The encoder half of src/neural.rs:13-19 (AutoEncoder): what NDSequence::encoded needs.
// w_encode is n_bins x n_latent row-major (Mat{flat, cols}, src/numerics.rs:169-173).
struct AutoEncoder {
    std::vector<float> w_encode, b_encode;
    size_t n_bins = 0;
    size_t n_latent() const { return b_encode.size(); }  // src/neural.rs:22-24
};

namespace detail {

inline void check(apd_ctx* ctx, apd_status st, const char* what)
{
    if (st != APD_OK) {
        const char* msg = apd_last_error(ctx);
        throw Panic(std::string(what) + " failed (status " + std::to_string((int)st) + "): " + (msg ? msg : ""));
    }
}

struct AllDevices {};

struct Context {
    apd_ctx* raw = nullptr;
    explicit Context(int device = 0) { check(nullptr, apd_create(device, &raw), "apd_create"); }
    // every visible GPU of the box in one context: the reference's single blocking align_all call
    // (src/main.rs:189-195) fans out inside the library
    explicit Context(AllDevices) { check(nullptr, apd_create_multi(nullptr, 0, &raw), "apd_create_multi"); }
    ~Context() { apd_destroy(raw); }
    Context(const Context&) = delete;
    Context& operator=(const Context&) = delete;
    // The spectrogram.rs glue: one pointer + length per NDSequence; the library copies and packs.
    // euclidean() in the reference silently assumes equal widths (src/numerics.rs:114-120 iterates
    // x.len()); a narrower sequence would be read past its end here, so a mismatch is a Panic.
    void set_sequences(const std::vector<const NDSequence*>& data)
    {
        std::vector<const float*> ptrs(data.size());
        std::vector<uint32_t> lens(data.size());
        const size_t dim = data.empty() ? 1u : data[0]->n_bins;
        for (size_t k = 0; k < data.size(); k++) {
            if (data[k]->n_bins != dim)
                throw Panic("sequence " + std::to_string(k) + " has n_bins " + std::to_string(data[k]->n_bins) +
                            " but sequence 0 has " + std::to_string(dim));
            ptrs[k] = data[k]->frames.data();
            lens[k] = (uint32_t)data[k]->len();
        }
        check(raw, apd_set_sequences(raw, ptrs.data(), lens.data(), (uint32_t)data.size(), (uint32_t)dim), "apd_set_sequences");
    }
    void set_sequences(const std::vector<NDSequence>& data)
    {
        std::vector<const NDSequence*> refs(data.size());
        for (size_t k = 0; k < data.size(); k++) refs[k] = &data[k];
        set_sequences(refs);
    }
    // `NDSequence::new(..).encoded(&nn)` of src/main.rs:150-161 on the device: the cepstra go up,
    // AutoEncoder::predict (src/neural.rs:55-71) runs per frame on the GPU, the embeddings land in the arena.
    void set_sequences_encoded(const std::vector<NDSequence>& cepstra, const AutoEncoder& nn)
    {
        std::vector<const float*> ptrs(cepstra.size());
        std::vector<uint32_t> lens(cepstra.size());
        for (size_t k = 0; k < cepstra.size(); k++) {
            if (cepstra[k].n_bins != nn.n_bins) throw Panic("assertion failed: self.cols == other.rows()");  // Mat::mul, src/numerics.rs:306
            ptrs[k] = cepstra[k].frames.data();
            lens[k] = (uint32_t)cepstra[k].len();
        }
        check(raw, apd_set_sequences_encoded(raw, ptrs.data(), lens.data(), (uint32_t)cepstra.size(), (uint32_t)nn.n_bins,
                                             nn.w_encode.data(), nn.b_encode.data(), (uint32_t)nn.n_latent()),
              "apd_set_sequences_encoded");
    }
};

// One single-GPU context shared by every Alignment::construct_alignment call of the process
// (creating a CUDA context per pair would cost more than the pair).
inline Context& pair_context(std::unique_lock<std::mutex>& held)
{
    static std::mutex mu;
    static Context ctx(0);
    held = std::unique_lock<std::mutex>(mu);
    return ctx;
}

}  // namespace detail

// Arc<Mutex<Vec<f32>>> (src/alignments.rs:13): `result->lock()` yields the flat matrix.
class SharedResult {
public:
    explicit SharedResult(size_t n) : values_(n, 0.0f) {}
    struct Guard {
        std::unique_lock<std::mutex> held;
        std::vector<float>& values;
        std::vector<float>& unwrap() { return values; }
        std::vector<float>* operator->() { return &values; }
    };
    Guard lock() { return Guard{std::unique_lock<std::mutex>(mutex_), values_}; }

private:
    std::mutex mutex_;
    std::vector<float> values_;
};

// src/alignments.rs:11-68
class AlignmentWorkers {
public:
    std::shared_ptr<std::vector<NDSequence>> data;
    std::shared_ptr<SharedResult> result;
    uint32_t mode = APD_MODE_STRICT;
    apd_stats stats{};

    explicit AlignmentWorkers(std::vector<NDSequence> sequences)
        : data(std::make_shared<std::vector<NDSequence>>(std::move(sequences))),
          result(std::make_shared<SharedResult>(data->size() * data->size()))  // n*n zeros, diagonal stays 0.0 (:20-23, :51)
    {
    }

    // Cepstra in, embeddings computed on the device (row f3): `data` holds the raw cepstra and the
    // auto-encoder is applied inside align_all instead of by NDSequence::encoded on the host.
    AlignmentWorkers(std::vector<NDSequence> cepstra, AutoEncoder nn) : AlignmentWorkers(std::move(cepstra))
    {
        encoder_ = std::make_shared<AutoEncoder>(std::move(nn));
    }

    // Blocking; `alignment_workers` is accepted and ignored beyond the division the reference
    // performs with it (src/alignments.rs:33 panics on 0).
    void align_all(const Discovery& params)
    {
        if (params.alignment_workers == 0) throw Panic("attempt to divide by zero");
        detail::Context ctx{detail::AllDevices{}};
        if (encoder_) ctx.set_sequences_encoded(*data, *encoder_);
        else ctx.set_sequences(*data);
        const apd_params p{params.warping_band_percentage, params.insertion_penalty, params.deletion_penalty,
                           params.match_penalty, mode};
        auto guard = result->lock();  // locked once, not once per pair (src/alignments.rs:56)
        detail::check(ctx.raw, apd_align_all(ctx.raw, &p, guard.values.data()), "apd_align_all");
        apd_get_stats(ctx.raw, &stats);
        std::printf("Aligned %llu ordered pairs, %llu cells in %.3f s on the GPU (%.1f GCUPS)\n",
                    (unsigned long long)stats.ordered_pairs, (unsigned long long)stats.cells_reference,
                    stats.kernel_ms / 1e3, stats.kernel_ms > 0 ? stats.cells_reference / (stats.kernel_ms * 1e6) : 0.0);
    }

private:
    std::shared_ptr<AutoEncoder> encoder_;
};

// The matrix the reference keeps only in memory (src/main.rs:194-195), on disk and back
// (<stem>.apdm + <stem>.apdm.json, see include/apd.h).
inline void save_matrix(const std::string& stem, const std::vector<float>& distances, size_t n_instances,
                        const std::string& params_json = "{}")
{
    if (distances.size() != n_instances * n_instances) throw Panic("distances must hold n_instances^2 entries");
    detail::check(nullptr, apd_save_matrix(stem.c_str(), distances.data(), (uint32_t)n_instances, params_json.c_str()),
                  "apd_save_matrix");
}
inline std::vector<float> load_matrix(const std::string& stem, size_t* n_instances)
{
    uint32_t n = 0;
    detail::check(nullptr, apd_load_matrix(stem.c_str(), nullptr, 0, &n, 1), "apd_load_matrix");
    std::vector<float> d((size_t)n * n);
    detail::check(nullptr, apd_load_matrix(stem.c_str(), d.data(), d.size(), &n, 1), "apd_load_matrix");
    if (n_instances) *n_instances = n;
    return d;
}

// src/alignments.rs:99-181.  `sparse` is not materialised (the reference never reads more than
// the score cell); the traced warping path is in `path`.
class Alignment {
public:
    size_t n = 0, m = 0;
    std::vector<std::pair<size_t, size_t>> path;  // (i, j), 1-based, end to start
    uint32_t mode = APD_MODE_STRICT;

    float score() const  // src/alignments.rs:116-125
    {
        if (!constructed_) return std::numeric_limits<float>::infinity();  // Alignment::new(): n == m == 0
        return score_;
    }

    void construct_alignment(const NDSequence& x, const NDSequence& y, const AlignmentParams& params)
    {
        n = x.len();
        m = y.len();
        std::unique_lock<std::mutex> held;
        detail::Context& ctx = detail::pair_context(held);
        ctx.set_sequences(std::vector<const NDSequence*>{&x, &y});
        const apd_params p{0.0f, params.insertion_penalty, params.deletion_penalty, params.match_penalty, mode};
        const uint64_t cap = n + m + 2;
        std::vector<uint32_t> cells(2 * cap);
        const uint32_t pair[2] = {0, 1};
        uint64_t plen = 0;
        detail::check(ctx.raw,
                      apd_align_pairs_band(ctx.raw, &p, params.warping_band, pair, 1, &score_, cells.data(), cap, &plen),
                      "apd_align_pairs_band");
        path.clear();
        for (uint64_t k = 0; k < plen && k < cap; k++) path.emplace_back(cells[2 * k], cells[2 * k + 1]);
        constructed_ = true;
    }

private:
    float score_ = 0.0f;
    bool constructed_ = false;
};

// src/clustering.rs:7-13
enum class Merge { Sequence2Sequence = 0, Sequence2Cluster = 1, Cluster2Sequence = 2, Cluster2Cluster = 3 };

// src/clustering.rs:18-25
struct ClusteringOperation {
    size_t merge_i, merge_j, into;
    float distance;
    Merge operation;
    bool tie;  // another root pair had exactly this linkage (HashSet-order dependent upstream)
};

struct AgglomerativeClustering {
    // src/clustering.rs:81-110 through the result-identical fast form apd_upgma.
    static std::pair<std::vector<ClusteringOperation>, std::set<size_t>> clustering(const std::vector<float>& distances,
                                                                                     size_t n_instances, float perc)
    {
        if (distances.size() != n_instances * n_instances) throw Panic("distances must hold n_instances^2 entries");
        std::vector<apd_merge> ops(n_instances ? n_instances : 1);
        std::vector<uint32_t> assignment(n_instances ? n_instances : 1);
        uint32_t n_ops = 0;
        float threshold = 0.0f;
        std::printf("\tset parents to self\n\tbuild initial dendrogram\n\testimate threshold\n");
        if (apd_upgma(distances.data(), (uint32_t)n_instances, perc, nullptr, ops.data(), &n_ops, &threshold,
                      assignment.data()) != APD_OK)
            throw Panic("index out of bounds: percentile(distances, perc)");  // src/numerics.rs:132
        std::printf("Clustering with %g\n", threshold);
        std::vector<ClusteringOperation> out;
        for (uint32_t k = 0; k < n_ops; k++)
            out.push_back({ops[k].merge_i, ops[k].merge_j, ops[k].into, ops[k].distance, (Merge)ops[k].operation, ops[k].tie != 0});
        std::set<size_t> clusters;
        for (size_t i = 0; i < n_instances; i++) clusters.insert(assignment[i]);
        return {out, clusters};
    }

    // src/clustering.rs:40-76
    static std::vector<std::vector<size_t>> cluster_sets(const std::vector<ClusteringOperation>& operations,
                                                         const std::set<size_t>& cluster_ids, size_t n_instances)
    {
        std::map<size_t, std::vector<size_t>> results;
        for (const ClusteringOperation& op : operations) {
            std::vector<size_t> cluster;
            for (size_t side : {op.merge_i, op.merge_j}) {
                auto it = results.find(side);
                if (it != results.end()) cluster.insert(cluster.end(), it->second.begin(), it->second.end());
                else cluster.push_back(side);
            }
            results[op.into] = cluster;
        }
        std::vector<std::vector<size_t>> grouped;
        for (size_t c : cluster_ids) {
            auto it = results.find(c);
            if (it == results.end()) { std::printf("Cluster not found: %zu | Singular cluster\n", c); continue; }
            std::vector<size_t> g;
            for (size_t i : it->second) if (i < n_instances) g.push_back(i);
            grouped.push_back(g);
        }
        return grouped;
    }
};

}  // namespace apd_host
