// upgma.cpp -- host-side, result-identical fast form of the reference's naive UPGMA
// (AgglomerativeClustering::clustering / merge / linkage, src/clustering.rs:81-209), the
// consumer of the distance matrix ("next" row f2).
//
// The reference recomputes every linkage from scratch for every merge (O(C^2 n + C n^2) per
// merge, ~O(n^4) overall: unusable at the 10 000-sequence matrices the GPU path produces).
// This version keeps the matrix of linkages between the current roots and, after a merge,
// recomputes only the rows / columns of the new cluster -- with the reference's own loop
// (one f32 accumulator over x ascending, y ascending, then one division by size_x * size_y),
// so every linkage is bit-identical to what `linkage()` would return and the strict-`<`
// argmin takes the same decision.  Where the reference iterates a HashSet (random order),
// roots are visited in ascending id (the convention of the literal restatement the tests
// compare against): exact ties are flagged.
//
// Pure host code (the north_star keeps UPGMA on the host); no CUDA in this file.
#include <algorithm>
#include <cmath>
#include <cstring>
#include <limits>
#include <thread>
#include <vector>

#include "../../include/apd.h"

namespace {

struct Root {
    uint32_t id;                    // cluster id (instances 0..n-1, merged clusters n, n+1, ...)
    std::vector<uint32_t> members;  // ascending instance ids
};

// src/clustering.rs:153-170
float linkage(const float* dist, uint32_t n, const std::vector<uint32_t>& xi, const std::vector<uint32_t>& yj)
{
    float distance = 0.0f;
    for (uint32_t x : xi) {
        const float* row = dist + (size_t)x * n;
        for (uint32_t y : yj) distance += row[y];
    }
    return distance / ((float)xi.size() * (float)yj.size());
}

// src/numerics.rs:125-133
bool percentile_host(const float* x, uint64_t len, float perc, float* out)
{
    std::vector<float> v;
    v.reserve(len);
    for (uint64_t k = 0; k < len; k++)
        if (x[k] == x[k]) v.push_back(x[k]);
    const float nf = (float)len * perc;
    uint64_t idx;
    if (!(nf == nf) || nf <= 0.0f) idx = 0;
    else if (nf >= 18446744073709551616.0f) idx = std::numeric_limits<uint64_t>::max();
    else idx = (uint64_t)nf;
    if (idx >= v.size()) return false;  // index out of bounds: the reference panics
    std::nth_element(v.begin(), v.begin() + idx, v.end());
    *out = v[idx];
    return true;
}

}  // namespace

extern "C" apd_status apd_upgma(const float* dist_nxn, uint32_t n, float perc, const float* threshold_in,
                                apd_merge* ops, uint32_t* n_ops, float* threshold_out, uint32_t* assignment_out)
{
    if (!n_ops || (n > 0 && !dist_nxn) || (n > 1 && !ops)) return APD_ERR_INVALID;
    *n_ops = 0;
    float threshold;
    if (threshold_in) threshold = *threshold_in;
    else if (!percentile_host(dist_nxn, (uint64_t)n * n, perc, &threshold)) return APD_ERR_INVALID;
    if (threshold_out) *threshold_out = threshold;

    const float INF = std::numeric_limits<float>::infinity();
    std::vector<Root> roots(n);           // indexed by slot
    std::vector<uint8_t> alive(n, 1);
    std::vector<uint32_t> slot_of_id(2 * (size_t)n + 1, UINT32_MAX);
    std::vector<uint32_t> parent(2 * (size_t)n + 1);
    for (uint32_t i = 0; i < n; i++) {
        roots[i].id = i;
        roots[i].members.assign(1, i);
        slot_of_id[i] = i;
        parent[i] = i;
    }
    uint32_t n_parents = n, n_clusters = n;
    // L[a*n + b]: linkage(root in slot a, root in slot b); singletons: dist / (1*1) == dist
    std::vector<float> L((size_t)n * n);
    if (n) std::memcpy(L.data(), dist_nxn, (size_t)n * n * sizeof(float));
    std::vector<float> rmin(n, INF);
    std::vector<uint32_t> rarg(n, UINT32_MAX);  // slot of the row minimum with the smallest id
    auto rescan = [&](uint32_t a) {
        float best = INF;
        uint32_t arg = UINT32_MAX;
        const float* row = L.data() + (size_t)a * n;
        for (uint32_t b = 0; b < n; b++) {
            if (!alive[b] || b == a) continue;
            const float l = row[b];
            if (l < best || (l == best && arg != UINT32_MAX && roots[b].id < roots[arg].id)) { best = l; arg = b; }
        }
        rmin[a] = best;
        rarg[a] = arg;
    };
    for (uint32_t a = 0; a < n; a++) rescan(a);

    const unsigned hw = std::max(1u, std::min(16u, std::thread::hardware_concurrency()));
    float distance = 0.0f;
    while (n_clusters > 1 && distance < threshold) {  // src/clustering.rs:104
        // -- argmin over ordered root pairs, rows and columns in ascending id, strict `<`
        float best = INF;
        uint32_t pa = UINT32_MAX;
        for (uint32_t a = 0; a < n; a++) {
            if (!alive[a] || rarg[a] == UINT32_MAX) continue;
            if (rmin[a] < best || (rmin[a] == best && pa != UINT32_MAX && roots[a].id < roots[pa].id)) {
                if (rmin[a] < INF) { best = rmin[a]; pa = a; }
            }
        }
        uint32_t p_id = 0, q_id = 0, tie = 0;
        uint32_t sp = UINT32_MAX, sq = UINT32_MAX;
        if (pa != UINT32_MAX) {
            sp = pa; sq = rarg[pa];
            p_id = roots[sp].id; q_id = roots[sq].id;
            // another unordered pair with exactly the same linkage: the reference's choice
            // would depend on HashSet iteration order
            for (uint32_t a = 0; a < n && !tie; a++) {
                if (!alive[a] || rmin[a] != best) continue;
                const float* row = L.data() + (size_t)a * n;
                for (uint32_t b = 0; b < n; b++) {
                    if (!alive[b] || b == a || row[b] != best) continue;
                    if ((a == sp && b == sq) || (a == sq && b == sp)) continue;
                    tie = 1;
                    break;
                }
            }
        }
        if (pa == UINT32_MAX) {
            // no finite linkage: the reference keeps its initial (0, 0) / +INF; an INF linkage
            // "ties" with that initial value in the literal restatement
            for (uint32_t a = 0; a < n && !tie; a++) {
                if (!alive[a]) continue;
                for (uint32_t b = 0; b < n; b++)
                    if (alive[b] && b != a && L[(size_t)a * n + b] == INF) { tie = 1; break; }
            }
        }
        // -- merge_clusters(): src/clustering.rs:133-141 (p = q = 0 if no finite linkage exists,
        // exactly what the reference's (0, 0) initial value does)
        const uint32_t k = n_parents;
        parent[p_id] = k;
        parent[q_id] = k;
        parent[k] = k;
        n_parents++;
        n_clusters--;
        apd_merge& op = ops[(*n_ops)++];
        op.merge_i = p_id; op.merge_j = q_id; op.into = k; op.distance = best; op.tie = tie;
        // src/clustering.rs:193-201
        op.operation = (p_id < n && q_id < n) ? 0u : ((p_id >= n && q_id >= n) ? 3u : ((p_id >= n && q_id < n) ? 2u : 1u));
        distance = best;
        if (sp == UINT32_MAX) {
            // degenerate (0, 0) merge: instance 0's chain now ends in k; keep the bookkeeping sane
            const uint32_t s0 = slot_of_id[0];
            if (s0 != UINT32_MAX && alive[s0]) { roots[s0].id = k; slot_of_id[k] = s0; slot_of_id[0] = UINT32_MAX; }
            continue;
        }
        // -- the new root takes p's slot
        std::vector<uint32_t> merged(roots[sp].members.size() + roots[sq].members.size());
        std::merge(roots[sp].members.begin(), roots[sp].members.end(), roots[sq].members.begin(),
                   roots[sq].members.end(), merged.begin());
        roots[sp].members.swap(merged);
        roots[sp].id = k;
        slot_of_id[k] = sp;
        slot_of_id[p_id] = slot_of_id[q_id] = UINT32_MAX;
        alive[sq] = 0;
        roots[sq].members.clear();
        roots[sq].members.shrink_to_fit();
        // -- linkages of the new root against every other root, both directions
        std::vector<uint32_t> others;
        for (uint32_t c = 0; c < n; c++)
            if (alive[c] && c != sp) others.push_back(c);
        const size_t work = roots[sp].members.size() * (size_t)n;
        const unsigned nt = (work > (1u << 16) && others.size() > 64) ? hw : 1;
        auto body = [&](size_t lo, size_t hi) {
            for (size_t t = lo; t < hi; t++) {
                const uint32_t c = others[t];
                L[(size_t)sp * n + c] = linkage(dist_nxn, n, roots[sp].members, roots[c].members);
                L[(size_t)c * n + sp] = linkage(dist_nxn, n, roots[c].members, roots[sp].members);
            }
        };
        if (nt == 1) {
            body(0, others.size());
        } else {
            std::vector<std::thread> th;
            for (unsigned t = 0; t < nt; t++)
                th.emplace_back(body, others.size() * t / nt, others.size() * (t + 1) / nt);
            for (auto& x : th) x.join();
        }
        // -- row minima
        rescan(sp);
        for (uint32_t c : others) {
            if (rarg[c] == sp || rarg[c] == sq) {
                rescan(c);
            } else {
                const float l = L[(size_t)c * n + sp];
                // k is the largest id: it loses every tie, so only a strictly smaller value moves the minimum
                if (l < rmin[c]) { rmin[c] = l; rarg[c] = sp; }
            }
        }
    }
    if (assignment_out) {
        for (uint32_t i = 0; i < n; i++) {
            uint32_t r = i;
            while (r != parent[r]) r = parent[r];
            assignment_out[i] = r;
        }
    }
    return APD_OK;
}
