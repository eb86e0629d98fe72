// B = A (x) A^T under the SharedSeeds semiring, version 2: expand by COLUMN, reduce by ROW, everything streamed.
//
// Reference: create_seed_matrix (src/SharedSeeds.cpp:4-10) = CombBLAS Mult_AnXBn_DoubleBuff with
//   multiply(posQ,posT) = {seeds[0]=(posQ,posT), n=1}             include/SharedSeeds.hpp:48-52
//   add(l,r)            = {l.seeds[0], r.seeds[0], l.n + r.n}     include/SharedSeeds.hpp:41-46
// followed by Prune(numshared <= 1).  Folding a nonzero's products in ascending column id (the canonical rule)
// gives seeds[0] = pair of the smallest shared column, seeds[1] = pair of the largest, numshared = number of shared columns.
//
// Why (profiles/r1_v8_spgemm.md): the row-wise hash SpGEMM of round 1 fetched, for every nonzero (i, c) of a row, the
// entries of column c from an 840 MB structure: two dependent sector-random reads per nonzero, 16.6 GB of DRAM traffic for
// 1.9 GB of algorithmic bytes (8.6 x), 9 % issue utilisation.  Nothing about the table was the problem; the gather was.
//
// Here the products are generated where the data lies in order:
//   k_sp2_expand   walks the COLUMNS (both operands column-major, read sequentially): every pair (i, j), i != j, of reads
//                  that share column c becomes one 16-byte tuple {j, c, pos_i, pos_j} appended to row i's region of the
//                  tuple array.  The regions are exact (the row's product count is known), the slot inside a region comes
//                  from one atomic on an L2-resident cursor per row.
//   k_sp2_*        one warp / CTA per row reads its tuples as ONE contiguous stream, hashes j into a shared-memory table
//                  with three commutative atomics (count, min c, max c), prunes, sorts the survivors by column and writes
//                  them; a second sweep over the same tuples (L2) drops the positions of the min- / max-column tuple
//                  into the output.  The diagonal entry of a row (reads share all their k-mers with themselves) comes
//                  from the row's own CSR entries, not from tuples.
// DRAM traffic: the two column-major operands once, the tuples written once and read once (twice from L2), B once.
#pragma once
#include "common.cuh"
#include "spgemm.cuh"

namespace elba {

struct __align__(16) Sp2Tuple { u32 j, c, pi, pj; };

struct Sp2Operands
{
    const u32 *l_cptr; const uint2 *l_cent;          // left operand by column: {row (local to the row block), pos}
    const u32 *r_cptr; const uint2 *r_cent;          // right operand by column: {row (local to the column block), pos}
    u64 ncol;                                        // reliable k-mers
    int64_t row0, col0;                              // global read id of left row 0 / right row 0
    u32 l_rows, r_rows;
};

// tuples of row i = products of the row minus the pairs with itself (diagonal blocks only)
__global__ void k_sp2_tuple_counts(const u64 *__restrict__ prod, const int64_t *__restrict__ rowptr, u32 nrows, int64_t row0, int64_t col0, u32 r_rows,
                                   u64 *__restrict__ tcnt)
{
    const u32 i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i > nrows) return;
    if (i == nrows) { tcnt[i] = 0; return; }
    const int64_t jd = row0 + (int64_t)i - col0;
    const u64 self = (jd >= 0 && jd < (int64_t)r_rows) ? (u64)(rowptr[i + 1] - rowptr[i]) : 0ull;
    tcnt[i] = prod[i] - self;
}

// rows [ra, rb) of this round; tup_off[i] - tup_off[ra] is where row i's region starts in `out`
__global__ void __launch_bounds__(256) k_sp2_expand(Sp2Operands op, const u64 *__restrict__ tup_off, u32 *__restrict__ cursor, u32 ra, u32 rb,
                                                    Sp2Tuple *__restrict__ out)
{
    const u32 lane = threadIdx.x & 31;
    const u64 base = tup_off[ra];
    const bool diag_block = op.row0 < op.col0 + (int64_t)op.r_rows && op.col0 < op.row0 + (int64_t)op.l_rows;
    const int64_t shift = op.row0 - op.col0;                     // right-local id of left row i = i + shift
    for (u64 c0 = ((u64)blockIdx.x * blockDim.x + threadIdx.x) & ~31ull; c0 < op.ncol; c0 += (u64)gridDim.x * blockDim.x)
    {
        const u64 c = c0 + lane;
        u32 lb = 0, le = 0, qb = 0, qe = 0;
        if (c < op.ncol) { lb = __ldg(op.l_cptr + c); le = __ldg(op.l_cptr + c + 1); qb = __ldg(op.r_cptr + c); qe = __ldg(op.r_cptr + c + 1); }
        const u32 work = (le - lb) * (qe - qb);
        // small columns (the usual case: <= UPPER entries each side): the lane does its own column
        if (work && work <= 64u)
        {
            for (u32 a = lb; a < le; ++a)
            {
                const uint2 x = __ldg(op.l_cent + a);
                if (x.x < ra || x.x >= rb) continue;
                const int64_t self = (int64_t)x.x + shift;
                u32 n = qe - qb;
                if (diag_block && self >= 0 && self < (int64_t)op.r_rows)
                    for (u32 q = qb; q < qe; ++q) if ((int64_t)__ldg(op.r_cent + q).x == self) { --n; break; }
                if (n == 0) continue;
                u64 o = tup_off[x.x] - base + atomicAdd(cursor + x.x, n);
                for (u32 q = qb; q < qe; ++q)
                {
                    const uint2 y = __ldg(op.r_cent + q);
                    if (diag_block && (int64_t)y.x == self) continue;
                    Sp2Tuple t; t.j = y.x; t.c = (u32)c; t.pi = x.y; t.pj = y.y;
                    out[o++] = t;
                }
            }
        }
        // large columns (large UPPER, repeats): the warp takes them one after the other, a left entry per lane
        unsigned big = __ballot_sync(0xffffffffu, work > 64u);
        while (big)
        {
            const int src = __ffs(big) - 1; big &= big - 1;
            const u32 b_lb = __shfl_sync(0xffffffffu, lb, src), b_le = __shfl_sync(0xffffffffu, le, src);
            const u32 b_qb = __shfl_sync(0xffffffffu, qb, src), b_qe = __shfl_sync(0xffffffffu, qe, src);
            const u32 cc = (u32)(c0 + src);
            for (u32 a = b_lb + lane; a < b_le; a += 32)
            {
                const uint2 x = __ldg(op.l_cent + a);
                if (x.x < ra || x.x >= rb) continue;
                const int64_t self = (int64_t)x.x + shift;
                u32 n = b_qe - b_qb;
                if (diag_block && self >= 0 && self < (int64_t)op.r_rows)
                {
                    // the right column is sorted by row: binary search for the row itself
                    u32 lo = b_qb, hi = b_qe;
                    while (lo < hi) { const u32 mid = (lo + hi) >> 1; if ((int64_t)__ldg(op.r_cent + mid).x < self) lo = mid + 1; else hi = mid; }
                    if (lo < b_qe && (int64_t)__ldg(op.r_cent + lo).x == self) --n;
                }
                if (n == 0) continue;
                u64 o = tup_off[x.x] - base + atomicAdd(cursor + x.x, n);
                for (u32 q = b_qb; q < b_qe; ++q)
                {
                    const uint2 y = __ldg(op.r_cent + q);
                    if (diag_block && (int64_t)y.x == self) continue;
                    Sp2Tuple t; t.j = y.x; t.c = cc; t.pi = x.y; t.pj = y.y;
                    out[o++] = t;
                }
            }
        }
    }
}

struct Sp2Args
{
    const int64_t *a_rowptr; const u32 *a_col; const u32 *a_pos;       // rows of the left operand (CSR): the diagonal entries
    const u64 *tup_off; const Sp2Tuple *tuples; u32 ra, rb;            // this round's rows and their tuple regions
    int64_t row0, col0; u32 r_rows;
    u32 nrows; int seed_count;
    u32 *t_col; int32_t *t_num; u32 *t_seeds; u64 cap;                 // unordered row storage
    u64 *counters;          // [0] output cursor, [1] nnz before prune, [2] overflow-list cursor
    u64 *row_off; u32 *row_nnz;
};

static constexpr u32 SP2_OUT = 0x80000000u;      // cnt word after the output phase: this slot survived, low bits = its place in the row

__device__ __forceinline__ u32 sp2_hash(u32 j, int shift) { return (j * 0x9E3779B1u) >> shift; }

// One output row.  keys/cnt/cmin/cmax/sortbuf: TS entries each; ctl: 8 words.  Returns false on table overflow.
template <bool BLOCK>
__device__ bool sp2_row(const Sp2Args &A, u32 row, u32 tid, u32 nth, u32 *keys, u32 *cnt, u32 *cmin, u32 *cmax, u32 *sortbuf, u32 TS, volatile u32 *ctl)
{
    const u32 mask = TS - 1;
    const u32 limit = TS - (TS >> 2);          // 75 % load
    const int shift = 32 - (31 - __clz(TS));
    for (u32 s = tid; s < TS; s += nth) { keys[s] = EMPTY32; cnt[s] = 0; cmin[s] = EMPTY32; cmax[s] = 0; }
    if (tid < 8) ctl[tid] = 0;
    group_sync<BLOCK>();

    const int64_t rs = A.a_rowptr[row];
    const u32 nnz = (u32)(A.a_rowptr[row + 1] - rs);
    const int64_t jd64 = A.row0 + (int64_t)row - A.col0;
    const bool has_diag = nnz > 0 && jd64 >= 0 && jd64 < (int64_t)A.r_rows;
    const u32 jd = (u32)jd64;
    if (has_diag && tid == 0)
    {
        // the read with itself: every column of the row is shared; first and last column (columns ascend inside a row)
        const u32 h = sp2_hash(jd, shift);
        keys[h] = jd; cnt[h] = nnz; cmin[h] = A.a_col[rs]; cmax[h] = A.a_col[rs + nnz - 1];
        ctl[0] = 1;
    }
    group_sync<BLOCK>();
    const Sp2Tuple *__restrict__ tp = A.tuples + (A.tup_off[row] - A.tup_off[A.ra]);
    const u32 nt = (u32)(A.tup_off[row + 1] - A.tup_off[row]);
    for (u32 e = tid; e < nt; e += nth)
    {
        const uint4 t = __ldcs(reinterpret_cast<const uint4*>(tp + e));
        const u32 j = t.x;
        u32 h = sp2_hash(j, shift);
        bool placed = false;
        while (ctl[2] == 0)
        {
            const u32 kk = ((volatile u32*)keys)[h];
            if (kk == j) { placed = true; break; }
            if (kk == EMPTY32)
            {
                const u32 prev = atomicCAS(&keys[h], EMPTY32, j);
                if (prev == EMPTY32)
                {
                    const u32 d = atomicAdd((u32*)&ctl[0], 1u);
                    if (d + 1 > limit) ctl[2] = 1;      // too many distinct columns for this table
                    placed = true; break;
                }
                if (prev == j) { placed = true; break; }
            }
            h = (h + 1) & mask;
        }
        if (placed)
        {
            atomicAdd(&cnt[h], 1u);
            if (t.y < ((volatile u32*)cmin)[h]) atomicMin(&cmin[h], t.y);
            if (t.y > ((volatile u32*)cmax)[h]) atomicMax(&cmax[h], t.y);
        }
    }
    group_sync<BLOCK>();
    if (ctl[2]) { group_sync<BLOCK>(); return false; }

    // survivors (numshared >= 2): Prune(numshared <= 1), src/SharedSeeds.cpp:8
    for (u32 s = tid; s < TS; s += nth)
        if (keys[s] != EMPTY32 && cnt[s] >= 2) { const u32 i = atomicAdd((u32*)&ctl[1], 1u); sortbuf[i] = s; }
    group_sync<BLOCK>();
    const u32 n = ctl[1];
    if (n <= 32)
    {
        if (tid < 32)
        {
            u32 v = tid < n ? sortbuf[tid] : EMPTY32;
            u32 kx = tid < n ? keys[v] : EMPTY32;
#pragma unroll
            for (u32 size = 2; size <= 32; size <<= 1)
#pragma unroll
                for (u32 stride = size >> 1; stride > 0; stride >>= 1)
                {
                    const u32 ok = __shfl_xor_sync(0xffffffffu, kx, stride), ov = __shfl_xor_sync(0xffffffffu, v, stride);
                    const bool up = (tid & size) == 0, low = (tid & stride) == 0;
                    const bool take = (low == up) ? (ok < kx) : (ok > kx);      // column ids of one row are distinct
                    if (take) { kx = ok; v = ov; }
                }
            if (tid < n) sortbuf[tid] = v;
        }
        group_sync<BLOCK>();
    }
    else
    {
        u32 m = 1; while (m < n) m <<= 1;
        for (u32 i = n + tid; i < m; i += nth) sortbuf[i] = EMPTY32;
        group_sync<BLOCK>();
        for (u32 size = 2; size <= m; size <<= 1)
            for (u32 stride = size >> 1; stride > 0; stride >>= 1)
            {
                for (u32 i = tid; i < (m >> 1); i += nth)
                {
                    const u32 lo = 2 * i - (i & (stride - 1)), hi = lo + stride;
                    const u32 a = sortbuf[lo], b = sortbuf[hi];
                    const u32 ka = a == EMPTY32 ? EMPTY32 : keys[a], kb = b == EMPTY32 ? EMPTY32 : keys[b];
                    const bool up = (lo & size) == 0;
                    if ((ka > kb) == up) { sortbuf[lo] = b; sortbuf[hi] = a; }
                }
                group_sync<BLOCK>();
            }
    }
    if (tid == 0)
    {
        const u64 off = n ? atomicAdd(&A.counters[0], (u64)n) : 0;
        atomicAdd(&A.counters[1], (u64)ctl[0]);
        A.row_off[row] = off; A.row_nnz[row] = n;
        ctl[4] = (u32)off; ctl[5] = (u32)(off >> 32);
    }
    group_sync<BLOCK>();
    const u64 off = ((u64)ctl[5] << 32) | ctl[4];
    if (off + n > A.cap) { group_sync<BLOCK>(); return true; }           // the host resizes and redoes the multiplication
    for (u32 e = tid; e < n; e += nth)
    {
        const u32 s = sortbuf[e];
        const u64 o = off + e;
        A.t_col[o] = keys[s]; A.t_num[o] = (int32_t)cnt[s];
        uint4 sd = make_uint4(0, 0, 0, 0);
        if (has_diag && keys[s] == jd)
        {
            sd.x = sd.y = A.a_pos[rs];
            if (A.seed_count > 1) sd.z = sd.w = A.a_pos[rs + nnz - 1];
        }
        reinterpret_cast<uint4*>(A.t_seeds)[o] = sd;
        cnt[s] = SP2_OUT | e;
    }
    group_sync<BLOCK>();
    // the positions of the smallest / largest shared column: second sweep over the row's tuples (still in L2)
    for (u32 e = tid; e < nt; e += nth)
    {
        const uint4 t = __ldg(reinterpret_cast<const uint4*>(tp + e));
        u32 h = sp2_hash(t.x, shift);
        while (keys[h] != t.x) h = (h + 1) & mask;
        const u32 w = cnt[h];
        if (w & SP2_OUT)
        {
            uint2 *sd = reinterpret_cast<uint2*>(A.t_seeds + 4 * (off + (w & ~SP2_OUT)));
            if (t.y == cmin[h]) sd[0] = make_uint2(t.z, t.w);
            if (t.y == cmax[h] && A.seed_count > 1) sd[1] = make_uint2(t.z, t.w);
        }
    }
    group_sync<BLOCK>();
    return true;
}

// A warp per row with a SMALL table: the table holds the row's distinct columns (a few dozen reads overlap a read), not its
// products (hundreds).  Rows with more distinct columns than 3/4 of the table go on to the CTA kernel.
static constexpr u32 SP2_WARP_TS = 128;
static constexpr int SP2_WARPS = 8;
static constexpr u32 SP2_WARP_MAXPROD = 8192;     // rows with more products go straight to the CTA kernel (one warp would take too long)

__global__ void __launch_bounds__(32 * SP2_WARPS) k_sp2_warp(Sp2Args A, const u32 *__restrict__ rows, const u32 *__restrict__ nrows_p,
                                                             u32 *__restrict__ big_rows, u32 *__restrict__ nbig)
{
    __shared__ u32 s_tab[SP2_WARPS][5 * SP2_WARP_TS];
    __shared__ u32 s_ctl[SP2_WARPS][8];
    const u32 w = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const u32 n = *nrows_p;
    u32 *t = s_tab[w];
    for (u32 i = blockIdx.x * SP2_WARPS + w; i < n; i += gridDim.x * SP2_WARPS)
    {
        const u32 row = rows[i];
        if (row < A.ra || row >= A.rb) continue;
        const bool ok = sp2_row<false>(A, row, lane, 32, t, t + SP2_WARP_TS, t + 2 * SP2_WARP_TS, t + 3 * SP2_WARP_TS, t + 4 * SP2_WARP_TS, SP2_WARP_TS, s_ctl[w]);
        if (!ok && lane == 0) big_rows[atomicAdd(nbig, 1u)] = row;
    }
}

__global__ void __launch_bounds__(SPG_BLOCK_THREADS) k_sp2_block(Sp2Args A, const u32 *__restrict__ rows, const u32 *__restrict__ nrows_p, u32 *__restrict__ overflow_rows)
{
    __shared__ u32 s_tab[5 * SPG_BLOCK_TS];
    __shared__ u32 s_ctl[8];
    const u32 n = *nrows_p;
    for (u32 i = blockIdx.x; i < n; i += gridDim.x)
    {
        const u32 row = rows[i];
        if (row < A.ra || row >= A.rb) continue;          // uniform
        const bool ok = sp2_row<true>(A, row, threadIdx.x, SPG_BLOCK_THREADS, s_tab, s_tab + SPG_BLOCK_TS, s_tab + 2 * SPG_BLOCK_TS,
                                      s_tab + 3 * SPG_BLOCK_TS, s_tab + 4 * SPG_BLOCK_TS, SPG_BLOCK_TS, s_ctl);
        if (!ok && threadIdx.x == 0) overflow_rows[atomicAdd(&A.counters[2], 1ull)] = row;
    }
}

// rows whose distinct-column count exceeded the shared-memory table: one global-memory table per CTA
__global__ void __launch_bounds__(SPG_BLOCK_THREADS) k_sp2_global(Sp2Args A, const u32 *__restrict__ rows, u32 n, u32 *__restrict__ scratch, u32 TS)
{
    __shared__ u32 s_ctl[8];
    u32 *t = scratch + (size_t)blockIdx.x * 5 * TS;
    for (u32 i = blockIdx.x; i < n; i += gridDim.x)
        sp2_row<true>(A, rows[i], threadIdx.x, SPG_BLOCK_THREADS, t, t + TS, t + 2 * (size_t)TS, t + 3 * (size_t)TS, t + 4 * (size_t)TS, TS, s_ctl);
}

// column-major operand from (column, {row, pos}) pairs sorted by column (stable: rows ascend inside a column): the column pointers
__global__ void k_sp2_colptr(const u32 *__restrict__ col_sorted, u64 n, u64 ncol, u32 *__restrict__ cptr)
{
    const u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    if (i > n) return;
    const u64 c_prev = i == 0 ? 0 : (u64)col_sorted[i - 1] + 1;           // first column whose pointer is not yet set
    const u64 c_here = i == n ? ncol + 1 : (u64)col_sorted[i] + 1;        // pointers of columns < c_here ... set to i
    for (u64 c = c_prev; c < c_here; ++c) cptr[c] = (u32)i;               // columns (c_prev .. col[i]] start at i; empty ones in between too
}

// A (key = read << col_bits | column, sorted) -> sort key (column) and value (read << 32 | pos) of the stable sort by column
__global__ void k_csr_to_colsort(const u64 *__restrict__ key, const u32 *__restrict__ pos, u64 n, int col_bits, u32 *__restrict__ col, u64 *__restrict__ val)
{
    const u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const u64 kk = key[i];
    col[i] = (u32)(kk & ((1ull << col_bits) - 1));
    val[i] = ((kk >> col_bits) << 32) | pos[i];
}
// the column-major operand in the C ABI's form: 64-bit column pointers, rows and positions in separate arrays
__global__ void k_transpose_api(const u32 *__restrict__ cptr, const uint2 *__restrict__ ent, u64 ncol, u64 nnz, int64_t *__restrict__ colptr, u32 *__restrict__ row, u32 *__restrict__ pos)
{
    const u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    if (i <= ncol) colptr[i] = (int64_t)cptr[i];
    if (i < nnz) { const uint2 e = ent[i]; row[i] = e.x; pos[i] = e.y; }
}

// slice [b, b + n) of the gathered A (key = global read << 32 | column, sorted by read): sort keys / values for the column-major operand
__global__ void k_sp2_slice(const u64 *__restrict__ key, const u32 *__restrict__ pos, u64 b, u64 n, u64 r0, u32 *__restrict__ col, u64 *__restrict__ val)
{
    const u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const u64 kk = key[b + i];
    col[i] = (u32)kk;
    val[i] = (((kk >> 32) - r0) << 32) | pos[b + i];
}
__global__ void k_sp2_unpack_val(const u64 *__restrict__ val, u64 n, uint2 *__restrict__ ent)
{
    const u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) { const u64 v = val[i]; ent[i] = make_uint2((u32)(v >> 32), (u32)v); }
}

} // namespace elba
