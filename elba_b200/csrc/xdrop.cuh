// X-drop seed-and-extend of the nonzeros of B: the stage that consumes the overlap matrix.
//
// Reference: PairwiseAlignment (src/PairwiseAlignment.cpp:5-106) walks the local block of B, keeps the nonzeros of the
// strict upper triangle (:52), and for each runs Overlap(len, seeds[0]).extend_overlap (:90-91, src/Overlap.cpp:20-73) =
// xdrop_aligner + classify_alignment (src/XDropAligner.cpp:7-282): check that the seed is an exact k-mer match and on
// which strand, extend left and right along antidiagonals under an X-drop, classify the overlap.
//
//   k_xdrop_select   one thread per nonzero of B: is it aligned (row < column), and where does it go in the pair list
//   k_xdrop          one WARP per pair.  Cell (c, r) of the extension matrix = c bases of the query side and r of the target
//                    side consumed; the warp sweeps antidiagonal d = c + r, lanes take the columns [lo, hi) of the live band
//                    32 at a time.  Three antidiagonals are kept, indexed by the absolute column, in a per-warp scratch
//                    in global memory (a band is tens to hundreds of cells wide and stays in L1/L2; the reads are up to
//                    tens of thousands of bases, so no fixed shared-memory band would be safe).  The band ends are
//                    trimmed with ballots instead of the reference's two serial scans.  All arithmetic is the reference's
//                    32-bit integer arithmetic, pruned cells included, so scores and end points are bit-exact; the one
//                    floating-point comparison of classify_alignment is evaluated with the same expression.
//
// This is the first correct version of this row (parity against the CPU oracle and the reference's own digests,
// tests/test_gpu_xdrop.py); it is
// not tuned: one warp per pair regardless of its length, both directions one after the other.
#pragma once
#include "common.cuh"

namespace elba {

static constexpr int XD_FIELDS = 13;     // begQ endQ begT endT score rc passed containedQ containedT direction directionT suffix suffixT

struct XdropArgs
{
    const uint8_t *buf; const u64 *off; const u32 *len;        // the 2-bit arena (src/DnaSeq.cpp:7-29)
    int k, mat, mis, gap, drop;
    const u32 *prow, *pcol, *sq, *st; u64 npairs;              // aligned pairs: row read, column read, seed position in each
    u32 row_base, col_base;                                    // read index of the block's first row / column read in buf / off / len
    int *scratch; u64 stride;                                  // per warp: 3 * stride ints
    int32_t *out;                                              // [npairs][XD_FIELDS]
};

struct XRead
{
    const uint8_t *mem; int len;
    __device__ __forceinline__ int at(int i) const { return (__ldg(mem + (i >> 2)) >> (6 - 2 * (i & 3))) & 3; }     // DnaSeq::operator[]
    __device__ __forceinline__ int rc_at(int i) const { return 3 - at(len - 1 - i); }                                // DnaSeq::revcomp_at
};

// nonzero e of B belongs to row r: rowptr[r] <= e < rowptr[r + 1]
__global__ void k_xdrop_select(const int64_t *__restrict__ rowptr, const u32 *__restrict__ col, u32 nrows, u64 nnz, int64_t row0, int64_t col0,
                               int global_rule, u64 *__restrict__ flag, u32 *__restrict__ rowof)
{
    const u64 e = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    if (e > nnz) return;
    if (e == nnz) { flag[e] = 0; return; }
    u32 lo = 0, hi = nrows;                       // rowptr[lo] <= e < rowptr[hi]
    while (hi - lo > 1) { const u32 mid = (lo + hi) >> 1; if ((u64)rowptr[mid] <= e) lo = mid; else hi = mid; }
    rowof[e] = lo;
    const int64_t lr = lo, lc = col[e];
    // src/PairwiseAlignment.cpp:52: upper triangle of the local block, its diagonal only where it lies above the global one.
    // That rule pairs block (i, j) with block (j, i) and so needs a SQUARE grid (the reference only runs on those); on a
    // pr != pc grid every unordered pair is taken where it lies in the global upper triangle instead.
    flag[e] = (global_rule ? (lr + row0 < lc + col0) : ((lr < lc) || (lr <= lc && lr + row0 < lc + col0))) ? 1 : 0;
}

__global__ void k_xdrop_pairs(const u64 *__restrict__ slot, const u32 *__restrict__ rowof, const u32 *__restrict__ col, const u32 *__restrict__ seeds, u64 nnz,
                              u32 *__restrict__ prow, u32 *__restrict__ pcol, u32 *__restrict__ sq, u32 *__restrict__ st, u64 *__restrict__ nzidx)
{
    const u64 e = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= nnz) return;
    const u64 s = slot[e];
    if (slot[e + 1] == s) return;                 // not aligned
    prow[s] = rowof[e]; pcol[s] = col[e]; sq[s] = seeds[4 * e]; st[s] = seeds[4 * e + 1]; nzidx[s] = e;      // seeds[0], :90
}

struct XSeedD { int begQ, endQ, begT, endT; bool rc; };

// _extend_seed_one_direction (src/XDropAligner.cpp:46-196) by one warp; every lane returns the same values
__device__ int xdrop_extend(const XRead &Q, const XRead &T, bool left, XSeedD &sd, int mat, int mis, int gap, int drop,
                            int *b0, int *b1, int *b2, u32 lane)
{
    const int extQ = left ? sd.begQ : Q.len - sd.endQ;
    const int extT = left ? sd.begT : T.len - sd.endT;
    const int cols = extQ + 1, rows = extT + 1;
    if (rows == 1 || cols == 1) return 0;
    const int imin = (int)0x80000000;
    const int floor_score = imin / (2 * max(cols, rows));
    gap = max(gap, floor_score);
    mis = max(mis, floor_score);
    const int NONE = imin - gap - mis;
    int *prev2 = b0, *prev1 = b1, *cur = b2;
    if (lane == 0) { prev1[0] = 0; cur[0] = cur[1] = (-gap > drop) ? NONE : gap; }
    __syncwarp();
    int top1 = 0, top = 1, lo = 1, hi = 2, d = 1, best = 0;
    int best_c = 0, best_r = 0, best_score = 0;
    while (lo < hi)
    {
        ++d;
        { int *t = prev2; prev2 = prev1; prev1 = cur; cur = t; }
        top1 = top;
        const int base = lo - 1;
        top = hi;
        if (lane == 0)
        {
            cur[base] = NONE; cur[top] = NONE;
            if ((long long)d * gap > (long long)best - drop)
            {
                if (base == 0) cur[0] = d * gap;
                if (d - hi == 0) cur[hi] = d * gap;
            }
        }
        int diag_best = d * gap;
        int win_c = -1, win_v = 0;                                  // the highest column of this lane that beats `best`
        for (int c = lo + (int)lane; c < hi; c += 32)
        {
            const int r = d - c;
            const int pq = left ? cols - 1 - c : c - 1 + sd.endQ;
            const int pt = left ? rows - 1 - r : r - 1 + sd.endT;
            const int tb = sd.rc ? T.rc_at(pt) : T.at(pt);
            int v = max(prev1[c - 1], prev1[c]) + gap;
            v = max(v, prev2[c - 1] + (Q.at(pq) == tb ? mat : mis));
            if (v < best - drop) cur[c] = NONE;
            else { cur[c] = v; diag_best = max(diag_best, v); }
            if (v > best) { win_c = c; win_v = v; }
        }
        // the last improving cell of the antidiagonal (highest column) is the one the reference ends up with
        int wc = win_c;
        for (int o = 16; o; o >>= 1) { wc = max(wc, __shfl_xor_sync(0xffffffffu, wc, o)); diag_best = max(diag_best, __shfl_xor_sync(0xffffffffu, diag_best, o)); }
        if (wc >= 0)
        {
            const unsigned who = __ballot_sync(0xffffffffu, win_c == wc);
            const int v = __shfl_sync(0xffffffffu, win_v, __ffs(who) - 1);
            best_c = wc; best_r = d - wc; best_score = v;
        }
        best = max(best, diag_best);
        __syncwarp();                                               // the antidiagonal is complete and visible to the warp
        // trim pruned cells from both ends of the band (two consecutive antidiagonals pruned there)
        while (true)
        {
            const int l = lo + (int)lane;
            const bool dead = l <= top && cur[l] == NONE && l - 1 <= top1 && prev1[l - 1] == NONE;
            const unsigned m = __ballot_sync(0xffffffffu, dead);
            const int run = (m == 0xffffffffu) ? 32 : __ffs(~m) - 1;                       // leading dead cells
            lo += run;
            if (run < 32) break;
        }
        while (true)
        {
            const int h = hi - (int)lane;
            const bool dead = h > base && cur[h - 1] == NONE && prev1[h - 1] == NONE;
            const unsigned m = __ballot_sync(0xffffffffu, dead);
            const int run = (m == 0xffffffffu) ? 32 : __ffs(~m) - 1;
            hi -= run;
            if (run < 32) break;
        }
        ++hi;
        lo = max(lo, d + 2 - rows);
        hi = min(hi, cols);
    }
    if (left) { sd.begT -= best_r; sd.begQ -= best_c; }
    else      { sd.endT += best_r; sd.endQ += best_c; }
    return best_score;
}

__global__ void __launch_bounds__(128) k_xdrop(XdropArgs A)
{
    const u32 lane = threadIdx.x & 31;
    const u64 warp = ((u64)blockIdx.x * blockDim.x + threadIdx.x) >> 5, nwarps = ((u64)gridDim.x * blockDim.x) >> 5;
    int *b0 = A.scratch + warp * 3 * A.stride, *b1 = b0 + A.stride, *b2 = b1 + A.stride;
    for (u64 p = warp; p < A.npairs; p += nwarps)
    {
        const u32 rq = A.prow[p] + A.row_base, rt = A.pcol[p] + A.col_base;
        XRead Q{A.buf + A.off[rq], (int)A.len[rq]}, T{A.buf + A.off[rt], (int)A.len[rt]};
        const int k = A.k, sq = (int)A.sq[p], st = (int)A.st[p];
        int begQ = 0, endQ = 0, begT = 0, endT = 0, score = -1; bool rc = false;
        // xdrop_aligner, src/XDropAligner.cpp:224-282
        bool ok = !(sq < 0 || sq + k > Q.len) && !(st < 0 || st + k > T.len) && !(sq == 0 && st == 0);
        if (ok)
        {
            rc = Q.at(sq + (k >> 1)) != T.at(st + (k >> 1));
            bool same = true;
            for (int i = (int)lane; i < k; i += 32) same = same && (Q.at(sq + i) == (rc ? T.rc_at(T.len - st - k + i) : T.at(st + i)));
            ok = __all_sync(0xffffffffu, same);
            if (!ok) rc = false;
        }
        if (ok)
        {
            XSeedD s; s.begQ = sq; s.endQ = sq + k; s.begT = rc ? T.len - st - k : st; s.endT = s.begT + k; s.rc = rc;
            XSeedD l = s, r = s;
            const int ls = xdrop_extend(Q, T, true, l, A.mat, A.mis, A.gap, A.drop, b0, b1, b2, lane);
            __syncwarp();
            const int rs = xdrop_extend(Q, T, false, r, A.mat, A.mis, A.gap, A.drop, b0, b1, b2, lane);
            __syncwarp();
            begQ = l.begQ; endQ = r.endQ;
            begT = rc ? T.len - r.endT : l.begT;
            endT = rc ? T.len - l.begT : r.endT;
            score = ls + rs + A.mat * k;
        }
        if (lane == 0)
        {
            // classify_alignment (src/XDropAligner.cpp:7-44) + the edge fields of Overlap::extend_overlap (src/Overlap.cpp:35-72)
            int kind = 0;     // 0 bad, 1 first contained, 2 second contained, 3 first -> second, 4 second -> first
            const int lenQ = Q.len, lenT = T.len;
            const int bT = rc ? lenT - endT : begT, eT = rc ? lenT - begT : endT;
            if (score > 0)
            {
                const int maplen = ((endT - begT) + (endQ - begQ)) / 2;
                const int overhang = min(begQ, bT) + min(lenQ - endQ, lenT - eT);
                const int overlap = maplen + overhang;
                const float thr = (float)((1.0 - 0.1) * (0.99 * (double)overlap));
                if (begQ <= bT && lenQ - endQ <= lenT - eT) kind = 1;
                else if (begQ >= bT && lenQ - endQ >= lenT - eT) kind = 2;
                else if ((float)score < thr || overlap < 500) kind = 0;
                else kind = begQ > bT ? 3 : 4;
            }
            int dir = -1, dirT = -1, suffix = 0, suffixT = 0;
            if (kind == 3) { dir = rc ? 0 : 1; dirT = rc ? 0 : 2; suffix = (lenT - eT) - (lenQ - endQ); suffixT = begQ - bT; }
            else if (kind == 4) { dir = rc ? 3 : 2; dirT = rc ? 3 : 1; suffix = bT - begQ; suffixT = (lenQ - endQ) - (lenT - eT); }
            int32_t *o = A.out + p * XD_FIELDS;
            o[0] = begQ; o[1] = endQ; o[2] = begT; o[3] = endT; o[4] = score; o[5] = rc; o[6] = kind != 0; o[7] = kind == 1; o[8] = kind == 2;
            o[9] = dir; o[10] = dirT; o[11] = suffix; o[12] = suffixT;
        }
    }
}

__global__ void k_max_u32(const u32 *__restrict__ v, u64 n, u32 *__restrict__ out)
{
    u32 m = 0;
    for (u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (u64)gridDim.x * blockDim.x) m = max(m, v[i]);
    for (int o = 16; o; o >>= 1) m = max(m, __shfl_xor_sync(0xffffffffu, m, o));
    if ((threadIdx.x & 31) == 0 && m) atomicMax(out, m);
}

} // namespace elba
