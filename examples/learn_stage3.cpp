// learn_stage3.cpp -- the reference's call site of the hot path (stage 3 of `learn()`,
// src/main.rs:187-203) written against the C++ mirror of its interface (include/apd_host.hpp):
//
//     let mut workers = alignments::AlignmentWorkers::new(signals);        // :189
//     workers.align_all(&discover);                                        // :191
//     let distances = workers.result.lock().unwrap().clone();              // :194-195
//     let (operations, clusters) = AgglomerativeClustering::clustering(distances, n, perc);   // :196-200
//     let grouped = AgglomerativeClustering::cluster_sets(&operations, &clusters, n);         // :203
//
//   learn_stage3 <sequences.bin> <Discovery.toml> <out_stem> [<encoder.bin>]
//   learn_stage3 --cluster-only <matrix.apdm> <n> <perc> <out_stem>      (host only: no GPU needed)
//
// sequences.bin: "APDS", u32 n, u32 dim, n x u32 lengths, then the frames (f32, row-major).
// encoder.bin (optional): "APDE", u32 n_bins, u32 n_latent, w_encode (n_bins x n_latent f32), b_encode;
//   the sequences are then raw cepstra and `NDSequence::new(..).encoded(&nn)` (src/main.rs:150-161)
//   runs on the device inside align_all.
// Writes <out_stem>.apdm + .apdm.json (the Vec<f32> handed to clustering, apd_save_matrix) and
// <out_stem>.merges.txt.
// A failure that is a panic in the reference prints Rust's panic line and exits with 101.
#include <cstdio>
#include <cstring>
#include <fstream>
#include <iostream>
#include <string>
#include <vector>

#include "apd_host.hpp"

using namespace apd_host;

static std::vector<NDSequence> read_sequences(const std::string& path)
{
    std::ifstream in(path, std::ios::binary);
    if (!in) throw Panic("cannot open " + path);
    char magic[4];
    uint32_t n = 0, dim = 0;
    in.read(magic, 4);
    in.read(reinterpret_cast<char*>(&n), 4);
    in.read(reinterpret_cast<char*>(&dim), 4);
    if (!in || std::memcmp(magic, "APDS", 4) != 0) throw Panic("not an APDS file: " + path);
    std::vector<uint32_t> lens(n);
    in.read(reinterpret_cast<char*>(lens.data()), 4 * (std::streamsize)n);
    std::vector<NDSequence> out;
    for (uint32_t k = 0; k < n; k++) {
        std::vector<float> flat((size_t)lens[k] * dim);
        in.read(reinterpret_cast<char*>(flat.data()), (std::streamsize)(flat.size() * 4));
        if (!in) throw Panic("truncated APDS file");
        out.emplace_back(dim, std::move(flat), k);
    }
    return out;
}

static AutoEncoder read_encoder(const std::string& path)
{
    std::ifstream in(path, std::ios::binary);
    if (!in) throw Panic("cannot open " + path);
    char magic[4];
    uint32_t n_bins = 0, n_latent = 0;
    in.read(magic, 4);
    in.read(reinterpret_cast<char*>(&n_bins), 4);
    in.read(reinterpret_cast<char*>(&n_latent), 4);
    if (!in || std::memcmp(magic, "APDE", 4) != 0) throw Panic("not an APDE file: " + path);
    AutoEncoder nn;
    nn.n_bins = n_bins;
    nn.w_encode.resize((size_t)n_bins * n_latent);
    nn.b_encode.resize(n_latent);
    in.read(reinterpret_cast<char*>(nn.w_encode.data()), (std::streamsize)(nn.w_encode.size() * 4));
    in.read(reinterpret_cast<char*>(nn.b_encode.data()), (std::streamsize)(nn.b_encode.size() * 4));
    if (!in) throw Panic("truncated APDE file");
    return nn;
}

static void write_outputs(const std::string& stem, const std::vector<float>& distances,
                          const std::vector<ClusteringOperation>& ops, const std::vector<std::vector<size_t>>& grouped)
{
    if (!distances.empty()) {
        size_t n = 0;
        while (n * n < distances.size()) n++;
        save_matrix(stem, distances, n);
    }
    std::ofstream t(stem + ".merges.txt");
    for (const ClusteringOperation& op : ops) {
        uint32_t bits;
        std::memcpy(&bits, &op.distance, 4);
        t << op.merge_i << " " << op.merge_j << " " << op.into << " " << bits << " " << (int)op.operation << " " << (op.tie ? 1 : 0)
          << "\n";
    }
    t << "# clusters " << grouped.size() << "\n";
}

int main(int argc, char** argv)
{
    try {
        if (argc == 6 && std::string(argv[1]) == "--cluster-only") {
            const size_t n = std::stoul(argv[3]);
            std::vector<float> distances(n * n);
            std::ifstream in(argv[2], std::ios::binary);
            in.read(reinterpret_cast<char*>(distances.data()), (std::streamsize)(distances.size() * 4));
            if (!in) throw Panic("cannot read matrix");
            auto res = AgglomerativeClustering::clustering(distances, n, std::stof(argv[4]));
            auto grouped = AgglomerativeClustering::cluster_sets(res.first, res.second, n);
            write_outputs(argv[5], {}, res.first, grouped);
            return 0;
        }
        if (argc != 4 && argc != 5) {
            std::fprintf(stderr, "usage: %s <sequences.bin> <Discovery.toml> <out_stem> [<encoder.bin>]\n", argv[0]);
            return 2;
        }
        const Discovery discover = Discovery::from_toml(argv[2]);
        std::vector<NDSequence> signals = read_sequences(argv[1]);
        std::printf("==== Starting Alignment And Clustering ==== \n");
        const size_t n = signals.size();
        AlignmentWorkers workers = (argc == 5) ? AlignmentWorkers(std::move(signals), read_encoder(argv[4]))
                                               : AlignmentWorkers(std::move(signals));
        workers.align_all(discover);
        const std::vector<float> distances = workers.result->lock().unwrap();  // .clone()
        auto res = AgglomerativeClustering::clustering(distances, n, discover.clustering_percentile);
        auto grouped = AgglomerativeClustering::cluster_sets(res.first, res.second, n);
        write_outputs(argv[3], distances, res.first, grouped);
        return 0;
    } catch (const Panic& p) {
        std::fprintf(stderr, "thread 'main' panicked at '%s'\n", p.what());
        return 101;
    }
}
