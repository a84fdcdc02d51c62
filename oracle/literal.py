"""Pure-Python transliteration of the reference DTW -- TEST INFRASTRUCTURE ONLY.

An independent restatement (dict-backed, numpy.float32 scalar arithmetic so every
operation rounds to f32 exactly like Rust's f32) used to cross-check the C oracle
bit for bit on small cases.  Citations are file:line in /root/reference/.
"PARITY UNPINNED" by the reference's own tests: it has none.
"""
import numpy as np

F = np.float32
INF = F(np.inf)
_USIZE_MAX = (1 << 64) - 1


def euclidean(x, y):
    """src/numerics.rs:114-120"""
    distance = F(0.0)
    for i in range(len(x)):
        t = F(x[i]) - F(y[i])
        distance = F(distance + F(t * t))
    return F(np.sqrt(distance))


def f32_as_usize(v):
    """Rust `f32 as usize`: truncate, saturate, NaN -> 0."""
    v = float(v)
    if v != v or v <= 0.0:
        return 0
    if v >= 18446744073709551616.0:
        return _USIZE_MAX
    return int(v)


def warping_band(pct, length):
    """src/discovery.rs:40 -- f32 product, then truncation."""
    return f32_as_usize(F(F(pct) * F(length)))


def window(band, n, m):
    """src/alignments.rs:173"""
    return max(band, abs(n - m)) + 2


def cells_visited(n, m, w):
    """src/alignments.rs:174-175"""
    total = 0
    for i in range(1, n + 1):
        lo = max(i - w if i > w else 0, 1)
        hi = min(i + w, m + 1)
        total += max(0, hi - lo)
    return total


def select_branch(match_score, insert_score, delete_score):
    """src/alignments.rs:153-159: 0 match, 1 insertion, 2 deletion."""
    if delete_score < match_score and delete_score < insert_score:
        return 2
    if insert_score < match_score and insert_score < delete_score:
        return 1
    return 0


def construct_alignment(x, y, band, ins, dele, mat):
    """src/alignments.rs:165-180; x, y are (T, D) float32 arrays.  Returns the sparse map."""
    n, m = len(x), len(y)
    ins, dele, mat = F(ins), F(dele), F(mat)
    sparse = {(0, 0): F(0.0)}
    w = window(band, n, m)
    for i in range(1, n + 1):
        lo = max(i - w if i > w else 0, 1)
        hi = min(i + w, m + 1)
        for j in range(lo, hi):
            distance = euclidean(x[i - 1], y[j - 1])
            match_score = sparse.get((i - 1, j - 1), INF)
            insert_score = sparse.get((i - 1, j), INF)
            delete_score = sparse.get((i, j - 1), INF)
            b = select_branch(match_score, insert_score, delete_score)
            if b == 2:
                node = F(delete_score + F(dele * distance))
            elif b == 1:
                node = F(insert_score + F(ins * distance))
            else:
                node = F(match_score + F(mat * distance))
            sparse[(i, j)] = node
    return sparse


def score(sparse, n, m):
    """src/alignments.rs:116-125"""
    if n == 0 or m == 0:
        return INF
    v = sparse.get((n - 1, m - 1))
    if v is None:
        return INF
    return F(v / F(n + m))


def dtw(x, y, pct, ins=1.0, dele=1.0, mat=1.0):
    """The pair-loop body, src/alignments.rs:52-57."""
    n, m = len(x), len(y)
    band = warping_band(pct, max(n, m))
    return score(construct_alignment(x, y, band, ins, dele, mat), n, m)


def path(sparse, n, m):
    """Trace-back from (n-1, m-1) by re-applying the forward rule (our definition,
    SURVEY.md Appendix A.8).  Returns [(i, j), ...] end-to-start, 1-based."""
    if n < 2 or m < 2 or (n - 1, m - 1) not in sparse:
        return []
    out = []
    i, j = n - 1, m - 1
    while i >= 1 and j >= 1:
        out.append((i, j))
        b = select_branch(sparse.get((i - 1, j - 1), INF), sparse.get((i - 1, j), INF),
                          sparse.get((i, j - 1), INF))
        if b == 2:
            j -= 1
        elif b == 1:
            i -= 1
        else:
            i -= 1
            j -= 1
    return out


def align_all(seqs, pct, ins=1.0, dele=1.0, mat=1.0):
    """src/alignments.rs:31-67 without the threads: n x n, diagonal 0."""
    n = len(seqs)
    out = np.zeros((n, n), dtype=np.float32)
    for i in range(n):
        for j in range(n):
            if i != j:
                out[i, j] = dtw(seqs[i], seqs[j], pct, ins, dele, mat)
    return out
