/*
 * TEST INFRASTRUCTURE ONLY (oracle/): CPU restatement of the stage that CONSUMES the overlap matrix B, the next row of the
 * hot-path scope table.  Nothing under elba_b200/ may include, link or execute this.
 *
 * Restates, on the 2-bit arena of DnaBuffer (src/DnaSeq.cpp:7-29: 4 bases per byte, base i at bits 6 - 2 (i % 4)):
 *   xdrop_aligner                 src/XDropAligner.cpp:224-282   seed check, strand, left + right extension, score
 *   _extend_seed_one_direction    src/XDropAligner.cpp:46-196    antidiagonal X-drop extension
 *   classify_alignment            src/XDropAligner.cpp:7-44
 *   Overlap::extend_overlap       src/Overlap.cpp:20-73          direction / suffix fields of the overlap edge
 * Pinned by tests/test_oracle_xdrop.py against oracle/_ref (the reference's own XDropAligner.cpp + Overlap.cpp compiled
 * unmodified) on every aligned pair of the fixtures.
 *
 * The extension keeps three antidiagonals of the DP matrix (cell (c, r): c bases of the query side, r of the target
 * side consumed; antidiagonal d = c + r).  Here they are arrays indexed by the ABSOLUTE column c; the reference indexes
 * the same cells relative to a moving offset, which changes nothing that is observable.
 */
#include <cstdint>
#include <vector>
#include <algorithm>
#include <limits>

namespace {

struct Read
{
    const uint8_t *mem; int len;
    int at(int i) const { return (mem[i >> 2] >> (6 - 2 * (i & 3))) & 3; }                /* DnaSeq::operator[] */
    int rc_at(int i) const { return 3 - at(len - 1 - i); }                               /* DnaSeq::revcomp_at */
};

struct Seed { int begQ, endQ, begT, endT; bool rc; };

/* One direction.  Returns the best extension score and moves the seed's begin (left) or end (right). */
int extend(const Read &Q, const Read &T, bool left, Seed &sd, int mat, int mis, int gap, int drop)
{
    const int extQ = left ? sd.begQ : Q.len - sd.endQ;
    const int extT = left ? sd.begT : T.len - sd.endT;
    const int cols = extQ + 1, rows = extT + 1;
    if (rows == 1 || cols == 1) return 0;

    const int imin = std::numeric_limits<int>::min();
    const int floor_score = imin / (2 * std::max(cols, rows));
    gap = std::max(gap, floor_score);
    mis = std::max(mis, floor_score);
    const int NONE = imin - gap - mis;                     /* a pruned cell; NONE + gap and NONE + mis do not wrap */

    /* prev2 / prev1 / cur: antidiagonals d - 2, d - 1, d.  Only the columns the reference keeps are ever read. */
    std::vector<int> prev2(cols + 2, NONE), prev1(cols + 2, NONE), cur(cols + 2, NONE);
    prev1[0] = 0;                                          /* antidiagonal 0: the seed's corner */
    cur[0] = cur[1] = (-gap > drop) ? NONE : gap;          /* antidiagonal 1: one gap either way */
    int top1 = 0, top = 1;                                 /* highest column stored in prev1 / cur */

    int lo = 1, hi = 2;                                    /* columns [lo, hi) of the next antidiagonal */
    int d = 1, best = 0;
    int best_c = 0, best_r = 0, best_score = 0;
    while (lo < hi)
    {
        ++d;
        prev2.swap(prev1); prev1.swap(cur);                /* cur now holds the stale d - 2 values: overwritten below */
        top1 = top;
        const int base = lo - 1;                           /* lowest column stored for this antidiagonal */
        top = hi;
        for (int c = base; c <= top; ++c) cur[c] = NONE;
        if ((long long)d * gap > (long long)best - drop)
        {
            if (base == 0) cur[0] = d * gap;               /* column 0: gaps only */
            if (d - hi == 0) cur[hi] = d * gap;            /* row 0: gaps only */
        }
        int diag_best = d * gap;
        for (int c = lo; c < hi; ++c)
        {
            const int r = d - c;
            const int pq = left ? cols - 1 - c : c - 1 + sd.endQ;
            const int pt = left ? rows - 1 - r : r - 1 + sd.endT;
            const int tb = sd.rc ? T.rc_at(pt) : T.at(pt);
            int v = std::max(prev1[c - 1], prev1[c]) + gap;
            v = std::max(v, prev2[c - 1] + (Q.at(pq) == tb ? mat : mis));
            if (v < best - drop) cur[c] = NONE;
            else { cur[c] = v; diag_best = std::max(diag_best, v); }
            /* compared with the best of the EARLIER antidiagonals: the last improving cell of an antidiagonal wins */
            if (v > best) { best_c = c; best_r = r; best_score = v; }
        }
        best = std::max(best, diag_best);
        /* shrink the band from both ends past cells pruned on two consecutive antidiagonals */
        while (lo <= top && cur[lo] == NONE && lo - 1 <= top1 && prev1[lo - 1] == NONE) ++lo;
        while (hi > base && cur[hi - 1] == NONE && prev1[hi - 1] == NONE) --hi;
        ++hi;
        lo = std::max(lo, d + 2 - rows);
        hi = std::min(hi, cols);
    }
    if (left) { sd.begT -= best_r; sd.begQ -= best_c; }
    else      { sd.endT += best_r; sd.endQ += best_c; }
    return best_score;
}

struct Aln { int begQ, endQ, begT, endT, score; bool rc; };

/* xdrop_aligner: -1 (and score -1, everything else 0) when the seed is not an exact k-mer match */
int align(const Read &Q, const Read &T, int k, int sq, int st, int mat, int mis, int gap, int drop, Aln &out)
{
    out = Aln{0, 0, 0, 0, -1, false};
    if (sq < 0 || sq + k > Q.len) return -1;
    if (st < 0 || st + k > T.len) return -1;
    if (sq == 0 && st == 0) return -1;
    const bool rc = Q.at(sq + (k >> 1)) != T.at(st + (k >> 1));
    for (int i = 0; i < k; ++i)
        if (Q.at(sq + i) != (rc ? T.rc_at(T.len - st - k + i) : T.at(st + i))) return -1;
    Seed s; s.begQ = sq; s.endQ = sq + k; s.begT = rc ? T.len - st - k : st; s.endT = s.begT + k; s.rc = rc;
    Seed l = s, r = s;
    const int ls = extend(Q, T, true, l, mat, mis, gap, drop);
    const int rs = extend(Q, T, false, r, mat, mis, gap, drop);
    out.begQ = l.begQ; out.endQ = r.endQ;
    out.begT = rc ? T.len - r.endT : l.begT;
    out.endT = rc ? T.len - l.begT : r.endT;
    out.rc = rc; out.score = ls + rs + mat * k;
    return out.score;
}

enum Kind { BAD = 0, FIRST_IN = 1, SECOND_IN = 2, FIRST_TO_SECOND = 3, SECOND_TO_FIRST = 4 };

Kind classify(const Aln &a, int lenQ, int lenT)
{
    if (a.score <= 0) return BAD;
    const int bT = a.rc ? lenT - a.endT : a.begT, eT = a.rc ? lenT - a.begT : a.endT;
    const int maplen = ((a.endT - a.begT) + (a.endQ - a.begQ)) / 2;
    const int overhang = std::min(a.begQ, bT) + std::min(lenQ - a.endQ, lenT - eT);
    const int overlap = maplen + overhang;
    const float thr = (1.0 - 0.1) * (0.99 * overlap);                          /* DELTACHERNOFF = 0.1, XDropAligner.hpp:9 */
    if (a.begQ <= bT && lenQ - a.endQ <= lenT - eT) return FIRST_IN;
    if (a.begQ >= bT && lenQ - a.endQ >= lenT - eT) return SECOND_IN;
    if (a.score < thr || overlap < 500) return BAD;
    return a.begQ > bT ? FIRST_TO_SECOND : SECOND_TO_FIRST;
}

} // namespace

extern "C" {

/* fields per pair, in this order */
enum { XO_BEGQ, XO_ENDQ, XO_BEGT, XO_ENDT, XO_SCORE, XO_RC, XO_PASSED, XO_CONTQ, XO_CONTT, XO_DIR, XO_DIRT, XO_SUFFIX, XO_SUFFIXT, XO_FIELDS };

int elba_oracle_xdrop_fields(void) { return XO_FIELDS; }

/* Overlap(len, seed).extend_overlap(seqQ, seqT, ...) for every pair (src/PairwiseAlignment.cpp:82-91, src/Overlap.cpp:20-73) */
void elba_oracle_xdrop_batch(const uint8_t *buf, const uint64_t *off, const uint64_t *len, int k,
                             const int64_t *rows, const int64_t *cols, const uint32_t *seedq, const uint32_t *seedt, uint64_t npairs,
                             int mat, int mis, int gap, int drop, int32_t *out)
{
    for (uint64_t p = 0; p < npairs; ++p)
    {
        const Read Q{buf + off[rows[p]], (int)len[rows[p]]}, T{buf + off[cols[p]], (int)len[cols[p]]};
        Aln a;
        align(Q, T, k, (int)seedq[p], (int)seedt[p], mat, mis, gap, drop, a);
        const Kind kind = classify(a, Q.len, T.len);
        int32_t *o = out + p * XO_FIELDS;
        o[XO_BEGQ] = a.begQ; o[XO_ENDQ] = a.endQ; o[XO_BEGT] = a.begT; o[XO_ENDT] = a.endT; o[XO_SCORE] = a.score; o[XO_RC] = a.rc;
        o[XO_PASSED] = kind != BAD; o[XO_CONTQ] = kind == FIRST_IN; o[XO_CONTT] = kind == SECOND_IN;
        int dir = -1, dirT = -1, suffix = 0, suffixT = 0;
        const int bT = a.rc ? T.len - a.endT : a.begT, eT = a.rc ? T.len - a.begT : a.endT;
        if (kind == FIRST_TO_SECOND)
        {
            dir = a.rc ? 0 : 1; dirT = a.rc ? 0 : 2;
            suffix = (T.len - eT) - (Q.len - a.endQ); suffixT = a.begQ - bT;
        }
        else if (kind == SECOND_TO_FIRST)
        {
            dir = a.rc ? 3 : 2; dirT = a.rc ? 3 : 1;
            suffix = bT - a.begQ; suffixT = (Q.len - a.endQ) - (T.len - eT);
        }
        o[XO_DIR] = dir; o[XO_DIRT] = dirT; o[XO_SUFFIX] = suffix; o[XO_SUFFIXT] = suffixT;
    }
}

} // extern "C"
