// Transitive reduction of the overlap graph: TransitiveReduction(R), src/TransitiveReduction.cpp:3-92, with the
// functors of include/TransitiveReduction.hpp:19-110 (MinPlusSR, GreaterThanSR, PlusFuzzSRing, TransitiveRemoval ...) and
// Overlap::Transpose / Overlap::arrows (include/Overlap.hpp:36-74).
//
// What the reference's loop amounts to (DESIGN.md has the derivation; the tests pin it against the reference's own file run
// on the CPU): R is made symmetric with the query / target fields swapped in the mirror image; N = R (x) R under the
// min-plus semiring keeps, per pair (i, j) and per combination of arrow ends, the shortest two-hop suffix; an edge (i, j) of R
// is transitive if suffix + FUZZ >= N(i, j).suffix_paths[direction]; transitive edges are removed in both orientations; the
// second round multiplies a matrix of default Overlaps (no arrows) and changes nothing.
//
// On the device the product is never formed: the test of edge (i, j) only needs N at (i, j), i.e. the two-hop paths
// i -> k -> j over the neighbours k of i, each a binary search for j in row k.  Rows of an overlap graph hold a few dozen
// entries (the coverage), so one thread per edge walks row i and probes the rows of its neighbours; everything is 32-bit
// gathers out of L2.  Small matrices, launch-bound: correctness first, no tuning in this version.
#pragma once
#include "common.cuh"

namespace elba {

struct TrGraph                     // the symmetric R in CSR, columns ascending
{
    const int64_t *rowptr;         // [n + 1]
    const u32 *row, *col;          // [m]
    const int32_t *dir, *dirT, *suf, *sufT;
    u64 m; u32 n;
};

// 2 nnz sort items: the triple itself and its mirror image.  key = row << 32 | col, payload = item index (mirror: nnz + e);
// the originals come first, so a stable sort leaves R's own entry in front of a mirror image at the same coordinate
// (R += RT keeps the left operand where both hold an entry: Overlap's operator+ returns lhs, include/Overlap.hpp:76-77).
__global__ void k_tr_items(const int64_t *__restrict__ row, const int64_t *__restrict__ col, u64 nnz, u64 *__restrict__ key, u32 *__restrict__ val)
{
    const u64 e = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= nnz) return;
    const u64 r = (u64)row[e], c = (u64)col[e];
    key[e] = (r << 32) | c;        val[e] = (u32)e;
    key[nnz + e] = (c << 32) | r;  val[nnz + e] = (u32)(nnz + e);
}

__global__ void k_tr_heads(const u64 *__restrict__ key, u64 m2, u64 *__restrict__ head)
{
    const u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    if (i > m2) return;
    head[i] = (i < m2 && (i == 0 || key[i] != key[i - 1])) ? 1ull : 0ull;
}

// the first item of every coordinate becomes an entry of the symmetric R; a mirror image carries Overlap::Transpose's swaps
__global__ void k_tr_entries(const u64 *__restrict__ key, const u32 *__restrict__ val, const u64 *__restrict__ rank, u64 m2, u64 nnz,
                             const int32_t *__restrict__ fields, u64 *__restrict__ okey, u32 *__restrict__ orow, u32 *__restrict__ ocol,
                             int32_t *__restrict__ dir, int32_t *__restrict__ dirT, int32_t *__restrict__ suf, int32_t *__restrict__ sufT,
                             u32 *__restrict__ src, uint8_t *__restrict__ tr)
{
    const u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= m2) return;
    if (i != 0 && key[i] == key[i - 1]) return;
    const u64 o = rank[i];
    const u32 v = val[i]; const bool mirror = v >= nnz; const u64 e = mirror ? v - nnz : v;
    const int32_t f0 = fields[4 * e], f1 = fields[4 * e + 1], f2 = fields[4 * e + 2], f3 = fields[4 * e + 3];
    okey[o] = key[i]; orow[o] = (u32)(key[i] >> 32); ocol[o] = (u32)key[i];
    dir[o] = mirror ? f1 : f0; dirT[o] = mirror ? f0 : f1; suf[o] = mirror ? f3 : f2; sufT[o] = mirror ? f2 : f3;
    src[o] = (u32)e; tr[o] = mirror ? 1 : 0;
}

__device__ __forceinline__ int64_t tr_find(const TrGraph &g, u32 r, u32 c)
{
    int64_t lo = __ldg(g.rowptr + r), hi = __ldg(g.rowptr + r + 1);
    const int64_t end = hi;
    while (lo < hi) { const int64_t mid = (lo + hi) >> 1; if (__ldg(g.col + mid) < c) lo = mid + 1; else hi = mid; }
    return (lo < end && __ldg(g.col + lo) == c) ? lo : -1;
}

// I(i, j): GreaterThanSR on F = R + FUZZ and N = R (x) R (src/TransitiveReduction.cpp:49,60; include/TransitiveReduction.hpp:45-56,83-103)
__global__ void k_tr_mark(TrGraph g, int32_t fuzz, uint8_t *__restrict__ I)
{
    const u64 p = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= g.m) return;
    const int32_t d = g.dir[p];
    uint8_t flag = 0;
    if (d != -1)
    {
        const u32 i = g.row[p], j = g.col[p];
        int best = 0x7FFFFFFF;                                        // N(i, j).suffix_paths[d]
        const int64_t b = g.rowptr[i], e = g.rowptr[i + 1];
        for (int64_t a = b; a < e; ++a)
        {
            const int32_t d1 = __ldg(g.dir + a);
            if (d1 == -1) continue;                                   // no arrows: MinPlusSR::multiply gives no path
            const int t1 = (d1 >> 1) & 1, h1 = d1 & 1;
            if (((2 * t1) | (d & 1)) != d) continue;                  // only paths that end up in suffix_paths[d]: 2 t1 + h2 == d
            const int64_t q = tr_find(g, __ldg(g.col + a), j);
            if (q < 0) continue;
            const int32_t d2 = __ldg(g.dir + q);
            if (d2 == -1) continue;
            const int t2 = (d2 >> 1) & 1, h2 = d2 & 1;
            if (t2 == h1 || 2 * t1 + h2 != d) continue;
            best = min(best, __ldg(g.suf + a) + __ldg(g.suf + q));
        }
        flag = (best != 0x7FFFFFFF && g.suf[p] + fuzz >= best) ? 1 : 0;
    }
    I[p] = flag;
}

// T = I + transpose(I) (src/TransitiveReduction.cpp:70-76), on the coordinates R holds
__global__ void k_tr_symmetric(TrGraph g, const uint8_t *__restrict__ I, uint8_t *__restrict__ T)
{
    const u64 p = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= g.m || !I[p]) return;
    T[p] = 1;
    const int64_t q = tr_find(g, g.col[p], g.row[p]);
    if (q >= 0) T[q] = 1;
}

// S = R without T (:86), without the explicit entry T starts with at (0, 0) (:27-28), without direction -1 (:88)
__global__ void k_tr_keep(TrGraph g, const uint8_t *__restrict__ T, u64 *__restrict__ keep)
{
    const u64 p = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    if (p > g.m) return;
    keep[p] = (p < g.m && !T[p] && !(g.row[p] == 0 && g.col[p] == 0) && g.dir[p] != -1) ? 1ull : 0ull;
}

__global__ void k_tr_output(TrGraph g, const uint8_t *__restrict__ T, const u64 *__restrict__ rank, const u32 *__restrict__ src, const uint8_t *__restrict__ tr,
                            int64_t *__restrict__ orow, int64_t *__restrict__ ocol, int32_t *__restrict__ ofields, u64 *__restrict__ osrc, uint8_t *__restrict__ otr)
{
    const u64 p = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= g.m) return;
    if (T[p] || (g.row[p] == 0 && g.col[p] == 0) || g.dir[p] == -1) return;
    const u64 o = rank[p];
    orow[o] = g.row[p]; ocol[o] = g.col[p];
    ofields[4 * o] = g.dir[p]; ofields[4 * o + 1] = g.dirT[p]; ofields[4 * o + 2] = g.suf[p]; ofields[4 * o + 3] = g.sufT[p];
    osrc[o] = src[p]; otr[o] = tr[p];
}

} // namespace elba
