// matrix_io.cpp -- on-disk form of the distance matrix and of alignment paths behind the C ABI
// (SURVEY.md section 8 row f4), so that a Rust / C++ host can persist what the reference keeps
// only in memory: the Vec<f32> handed to clustering() (src/main.rs:194-200) and the alignment
// paths README.md:67 promises.  Same format as audio_pattern_discovery_b200/matrix_io.py
// (either side reads what the other wrote):
//   <stem>.apdm        n*n little-endian f32, row-major, result[x*n+y]
//   <stem>.apdm.json   {"format": "apd-matrix-1", "n": .., "dtype": "<f4", "params": {..}, "sha256": ".."}
//   <stem>.apdp.json   {"format": "apd-paths-1", "paths": [{"i": .., "j": .., "score": .., "path": [[i, j], ..]}]}
// Host code only; no CUDA.
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/apd.h"
#include "apd_internal.h"

namespace {

// ---- SHA-256 (FIPS 180-4) ---------------------------------------------------------------
struct Sha256 {
    uint32_t h[8];
    uint8_t buf[64];
    uint64_t len = 0;
    size_t fill = 0;
    Sha256()
    {
        static const uint32_t init[8] = {0x6a09e667, 0xbb67ae85, 0x3c6ef372, 0xa54ff53a, 0x510e527f, 0x9b05688c, 0x1f83d9ab, 0x5be0cd19};
        std::memcpy(h, init, sizeof(h));
    }
    static uint32_t rotr(uint32_t x, int n) { return (x >> n) | (x << (32 - n)); }
    void block(const uint8_t* p)
    {
        static const uint32_t K[64] = {
            0x428a2f98, 0x71374491, 0xb5c0fbcf, 0xe9b5dba5, 0x3956c25b, 0x59f111f1, 0x923f82a4, 0xab1c5ed5, 0xd807aa98, 0x12835b01, 0x243185be,
            0x550c7dc3, 0x72be5d74, 0x80deb1fe, 0x9bdc06a7, 0xc19bf174, 0xe49b69c1, 0xefbe4786, 0x0fc19dc6, 0x240ca1cc, 0x2de92c6f, 0x4a7484aa,
            0x5cb0a9dc, 0x76f988da, 0x983e5152, 0xa831c66d, 0xb00327c8, 0xbf597fc7, 0xc6e00bf3, 0xd5a79147, 0x06ca6351, 0x14292967, 0x27b70a85,
            0x2e1b2138, 0x4d2c6dfc, 0x53380d13, 0x650a7354, 0x766a0abb, 0x81c2c92e, 0x92722c85, 0xa2bfe8a1, 0xa81a664b, 0xc24b8b70, 0xc76c51a3,
            0xd192e819, 0xd6990624, 0xf40e3585, 0x106aa070, 0x19a4c116, 0x1e376c08, 0x2748774c, 0x34b0bcb5, 0x391c0cb3, 0x4ed8aa4a, 0x5b9cca4f,
            0x682e6ff3, 0x748f82ee, 0x78a5636f, 0x84c87814, 0x8cc70208, 0x90befffa, 0xa4506ceb, 0xbef9a3f7, 0xc67178f2};
        uint32_t w[64];
        for (int i = 0; i < 16; i++) w[i] = (uint32_t)p[4 * i] << 24 | (uint32_t)p[4 * i + 1] << 16 | (uint32_t)p[4 * i + 2] << 8 | p[4 * i + 3];
        for (int i = 16; i < 64; i++) {
            const uint32_t s0 = rotr(w[i - 15], 7) ^ rotr(w[i - 15], 18) ^ (w[i - 15] >> 3);
            const uint32_t s1 = rotr(w[i - 2], 17) ^ rotr(w[i - 2], 19) ^ (w[i - 2] >> 10);
            w[i] = w[i - 16] + s0 + w[i - 7] + s1;
        }
        uint32_t a = h[0], b = h[1], c = h[2], d = h[3], e = h[4], f = h[5], g = h[6], hh = h[7];
        for (int i = 0; i < 64; i++) {
            const uint32_t S1 = rotr(e, 6) ^ rotr(e, 11) ^ rotr(e, 25), ch = (e & f) ^ (~e & g);
            const uint32_t t1 = hh + S1 + ch + K[i] + w[i];
            const uint32_t S0 = rotr(a, 2) ^ rotr(a, 13) ^ rotr(a, 22), mj = (a & b) ^ (a & c) ^ (b & c);
            const uint32_t t2 = S0 + mj;
            hh = g; g = f; f = e; e = d + t1; d = c; c = b; b = a; a = t1 + t2;
        }
        h[0] += a; h[1] += b; h[2] += c; h[3] += d; h[4] += e; h[5] += f; h[6] += g; h[7] += hh;
    }
    void update(const void* data, size_t n)
    {
        const uint8_t* p = static_cast<const uint8_t*>(data);
        len += n;
        if (fill) {
            const size_t take = std::min(n, 64 - fill);
            std::memcpy(buf + fill, p, take);
            fill += take; p += take; n -= take;
            if (fill == 64) { block(buf); fill = 0; }
        }
        for (; n >= 64; p += 64, n -= 64) block(p);
        if (n) { std::memcpy(buf, p, n); fill = n; }
    }
    std::string hex()
    {
        const uint64_t bits = len * 8;
        const uint8_t one = 0x80, zero = 0;
        update(&one, 1);
        while (fill != 56) update(&zero, 1);
        uint8_t be[8];
        for (int i = 0; i < 8; i++) be[i] = (uint8_t)(bits >> (56 - 8 * i));
        update(be, 8);
        char out[65];
        for (int i = 0; i < 8; i++) std::snprintf(out + 8 * i, 9, "%08x", h[i]);
        return std::string(out, 64);
    }
};

// ---- a JSON reader just big enough for the two headers -------------------------------------
struct Json {
    const char* p;
    const char* end;
    bool ok = true;
    void ws() { while (p < end && (*p == ' ' || *p == '\n' || *p == '\t' || *p == '\r')) p++; }
    bool eat(char c) { ws(); if (p < end && *p == c) { p++; return true; } return false; }
    bool peek(char c) { ws(); return p < end && *p == c; }
    std::string str()
    {
        std::string s;
        if (!eat('"')) { ok = false; return s; }
        while (p < end && *p != '"') {
            if (*p == '\\' && p + 1 < end) {
                p++;
                switch (*p) {
                    case 'n': s += '\n'; break; case 't': s += '\t'; break; case 'r': s += '\r'; break;
                    case 'b': s += '\b'; break; case 'f': s += '\f'; break;
                    case 'u': s += '?'; p += (end - p > 4 ? 4 : 0); break;
                    default: s += *p;
                }
                p++;
            } else s += *p++;
        }
        if (!eat('"')) ok = false;
        return s;
    }
    double num()
    {
        ws();
        char* e = nullptr;
        const double v = std::strtod(p, &e);
        if (e == p) { ok = false; return 0; }
        p = e;
        return v;
    }
    void skip()  // any value
    {
        ws();
        if (p >= end) { ok = false; return; }
        if (*p == '"') { str(); return; }
        if (*p == '{') {
            p++;
            if (eat('}')) return;
            do { str(); if (!eat(':')) { ok = false; return; } skip(); } while (ok && eat(','));
            if (!eat('}')) ok = false;
            return;
        }
        if (*p == '[') {
            p++;
            if (eat(']')) return;
            do { skip(); } while (ok && eat(','));
            if (!eat(']')) ok = false;
            return;
        }
        if (!std::strncmp(p, "true", 4) || !std::strncmp(p, "null", 4)) { p += 4; return; }
        if (!std::strncmp(p, "false", 5)) { p += 5; return; }
        num();
    }
};

bool read_file(const std::string& path, std::string& out)
{
    FILE* f = std::fopen(path.c_str(), "rb");
    if (!f) return false;
    char buf[1 << 16];
    size_t n;
    out.clear();
    while ((n = std::fread(buf, 1, sizeof(buf), f)) > 0) out.append(buf, n);
    std::fclose(f);
    return true;
}

apd_status io_fail(apd_status s, const std::string& msg)
{
    apd::set_thread_error(msg);
    return s;
}

std::string score_json(float v)
{
    if (v != v) return "\"nan\"";
    if (std::isinf(v)) return v > 0 ? "\"inf\"" : "\"-inf\"";
    char b[32];
    std::snprintf(b, sizeof(b), "%.9g", (double)v);
    return b;
}

float score_from(Json& j)
{
    if (j.peek('"')) {
        const std::string s = j.str();
        if (s == "inf") return INFINITY;
        if (s == "-inf") return -INFINITY;
        return NAN;
    }
    return (float)j.num();
}

}  // namespace

extern "C" {

apd_status apd_save_matrix(const char* stem, const float* dist_nxn, uint32_t n, const char* params_json)
{
    if (!stem || (n && !dist_nxn)) return io_fail(APD_ERR_INVALID, "stem/dist_nxn is NULL");
    const std::string payload = std::string(stem) + ".apdm";
    const size_t bytes = (size_t)n * n * sizeof(float);   // the library targets little-endian hosts (x86-64, aarch64)
    FILE* f = std::fopen(payload.c_str(), "wb");
    if (!f) return io_fail(APD_ERR_INVALID, "cannot create " + payload);
    const bool wrote = bytes == 0 || std::fwrite(dist_nxn, 1, bytes, f) == bytes;
    if (std::fclose(f) != 0 || !wrote) return io_fail(APD_ERR_INTERNAL, "short write to " + payload);
    Sha256 sha;
    sha.update(dist_nxn, bytes);
    const std::string meta = payload + ".json";
    f = std::fopen(meta.c_str(), "wb");
    if (!f) return io_fail(APD_ERR_INVALID, "cannot create " + meta);
    std::fprintf(f, "{\"dtype\": \"<f4\", \"format\": \"apd-matrix-1\", \"n\": %u, \"params\": %s, \"sha256\": \"%s\"}\n", n,
                 (params_json && params_json[0]) ? params_json : "{}", sha.hex().c_str());
    if (std::fclose(f) != 0) return io_fail(APD_ERR_INTERNAL, "short write to " + meta);
    return APD_OK;
}

apd_status apd_load_matrix(const char* stem, float* out_nxn, uint64_t cap_floats, uint32_t* n_out, int verify)
{
    if (!stem || !n_out) return io_fail(APD_ERR_INVALID, "stem/n_out is NULL");
    const std::string payload = std::string(stem) + ".apdm";
    std::string meta;
    if (!read_file(payload + ".json", meta)) return io_fail(APD_ERR_INVALID, "cannot read " + payload + ".json");
    Json j{meta.data(), meta.data() + meta.size()};
    std::string format, dtype, sha;
    double n = -1;
    if (!j.eat('{')) return io_fail(APD_ERR_INVALID, "not a JSON object: " + payload + ".json");
    if (!j.eat('}')) {
        do {
            const std::string key = j.str();
            if (!j.eat(':')) { j.ok = false; break; }
            if (key == "format") format = j.str();
            else if (key == "dtype") dtype = j.str();
            else if (key == "sha256") sha = j.str();
            else if (key == "n") n = j.num();
            else j.skip();
        } while (j.ok && j.eat(','));
    }
    if (!j.ok || format != "apd-matrix-1" || dtype != "<f4" || n < 0 || n > 4294967295.0)
        return io_fail(APD_ERR_INVALID, "not an apd-matrix-1 header: " + payload + ".json");
    *n_out = (uint32_t)n;
    const uint64_t count = (uint64_t)*n_out * *n_out;
    if (!out_nxn) return APD_OK;   // size query
    if (cap_floats < count) return io_fail(APD_ERR_INVALID, "output buffer too small for the stored matrix");
    FILE* f = std::fopen(payload.c_str(), "rb");
    if (!f) return io_fail(APD_ERR_INVALID, "cannot read " + payload);
    const size_t got = count ? std::fread(out_nxn, sizeof(float), count, f) : 0;
    char extra;
    const bool longer = std::fread(&extra, 1, 1, f) == 1;
    std::fclose(f);
    if (got != count || longer) return io_fail(APD_ERR_INVALID, "payload size does not match n");
    if (verify) {
        Sha256 s2;
        s2.update(out_nxn, count * sizeof(float));
        if (s2.hex() != sha) return io_fail(APD_ERR_INVALID, "payload checksum mismatch");
    }
    return APD_OK;
}

apd_status apd_save_paths(const char* stem, const uint32_t* pairs_ij, uint64_t n_pairs, const float* scores,
                          const uint32_t* paths_ij, uint64_t path_cap, const uint64_t* path_lens)
{
    if (!stem || (n_pairs && (!pairs_ij || !scores || !path_lens))) return io_fail(APD_ERR_INVALID, "NULL argument");
    const std::string out = std::string(stem) + ".apdp.json";
    FILE* f = std::fopen(out.c_str(), "wb");
    if (!f) return io_fail(APD_ERR_INVALID, "cannot create " + out);
    std::fprintf(f, "{\"format\": \"apd-paths-1\", \"paths\": [");
    for (uint64_t k = 0; k < n_pairs; k++) {
        std::fprintf(f, "%s{\"i\": %u, \"j\": %u, \"score\": %s, \"path\": [", k ? ", " : "", pairs_ij[2 * k], pairs_ij[2 * k + 1],
                     score_json(scores[k]).c_str());
        const uint64_t L = paths_ij ? (path_lens[k] < path_cap ? path_lens[k] : path_cap) : 0;
        const uint32_t* p = paths_ij ? paths_ij + k * path_cap * 2 : nullptr;
        for (uint64_t q = 0; q < L; q++) std::fprintf(f, "%s[%u, %u]", q ? ", " : "", p[2 * q], p[2 * q + 1]);
        std::fprintf(f, "]}");
    }
    std::fprintf(f, "]}\n");
    if (std::fclose(f) != 0) return io_fail(APD_ERR_INTERNAL, "short write to " + out);
    return APD_OK;
}

apd_status apd_load_paths(const char* stem, uint32_t* pairs_ij, float* scores, uint64_t* path_lens, uint64_t cap_pairs,
                          uint32_t* paths_ij, uint64_t path_cap, uint64_t* n_pairs)
{
    if (!stem || !n_pairs) return io_fail(APD_ERR_INVALID, "stem/n_pairs is NULL");
    const std::string in = std::string(stem) + ".apdp.json";
    std::string doc;
    if (!read_file(in, doc)) return io_fail(APD_ERR_INVALID, "cannot read " + in);
    Json j{doc.data(), doc.data() + doc.size()};
    std::string format;
    uint64_t count = 0;
    bool seen_paths = false;
    if (!j.eat('{')) return io_fail(APD_ERR_INVALID, "not a JSON object: " + in);
    if (!j.eat('}')) {
        do {
            const std::string key = j.str();
            if (!j.eat(':')) { j.ok = false; break; }
            if (key == "format") format = j.str();
            else if (key == "paths") {
                seen_paths = true;
                if (!j.eat('[')) { j.ok = false; break; }
                if (j.eat(']')) continue;
                do {  // one entry
                    uint32_t pi = 0, pj = 0;
                    float sc = NAN;
                    uint64_t L = 0;
                    const bool keep = count < cap_pairs;
                    if (!j.eat('{')) { j.ok = false; break; }
                    do {
                        const std::string k2 = j.str();
                        if (!j.eat(':')) { j.ok = false; break; }
                        if (k2 == "i") pi = (uint32_t)j.num();
                        else if (k2 == "j") pj = (uint32_t)j.num();
                        else if (k2 == "score") sc = score_from(j);
                        else if (k2 == "path") {
                            if (!j.eat('[')) { j.ok = false; break; }
                            if (!j.eat(']')) {
                                do {
                                    if (!j.eat('[')) { j.ok = false; break; }
                                    const uint32_t a = (uint32_t)j.num();
                                    if (!j.eat(',')) { j.ok = false; break; }
                                    const uint32_t b = (uint32_t)j.num();
                                    if (!j.eat(']')) { j.ok = false; break; }
                                    if (keep && paths_ij && L < path_cap) { paths_ij[(count * path_cap + L) * 2] = a; paths_ij[(count * path_cap + L) * 2 + 1] = b; }
                                    L++;
                                } while (j.ok && j.eat(','));
                                if (!j.eat(']')) j.ok = false;
                            }
                        } else j.skip();
                    } while (j.ok && j.eat(','));
                    if (!j.eat('}')) j.ok = false;
                    if (keep) {
                        if (pairs_ij) { pairs_ij[2 * count] = pi; pairs_ij[2 * count + 1] = pj; }
                        if (scores) scores[count] = sc;
                        if (path_lens) path_lens[count] = L;
                    }
                    count++;
                } while (j.ok && j.eat(','));
                if (!j.eat(']')) j.ok = false;
            } else j.skip();
        } while (j.ok && j.eat(','));
    }
    if (!j.ok || format != "apd-paths-1" || !seen_paths) return io_fail(APD_ERR_INVALID, "not an apd-paths-1 file: " + in);
    *n_pairs = count;
    return APD_OK;
}

}  // extern "C"
