// B = A (x) A^T under the SharedSeeds semiring, as a row-wise hash SpGEMM.
//
// Reference: create_seed_matrix (src/SharedSeeds.cpp:4-10) = CombBLAS Mult_AnXBn_DoubleBuff with
//   multiply(posQ,posT) = {seeds[0]=(posQ,posT), n=1}             include/SharedSeeds.hpp:48-52
//   add(l,r)            = {l.seeds[0], r.seeds[0], l.n + r.n}     include/SharedSeeds.hpp:41-46
// followed by Prune(numshared <= 1).  Folding a nonzero's products in ascending column id (the
// canonical rule) gives seeds[0] = pair of the smallest shared column, seeds[1] = pair of
// the largest, numshared = number of shared columns.  So per output (i,j) the accumulator is
//   count, min t, max t        (t = index of the shared column inside row i of A, ascending by column)
// three commutative 32-bit shared-memory atomics; the position pairs are looked up in the epilogue.
//
// Row i: for every nonzero (c, pos_i) of A(i,:), for every (j, pos_j) of column c of A (<= UPPER entries),
// hash j into the row's table.  Rows are binned by their product count: a warp per light row
// (256-slot table), two warps per medium row (1024 slots, 11 rows in flight per SM), a CTA per heavy row (2048
// slots), a global-memory table for rows that overflow that.
#pragma once
#include "common.cuh"

namespace elba {

static constexpr u32 EMPTY32 = 0xFFFFFFFFu;
static constexpr u32 SPG_WARP_TS = 256;      // slots per warp-row table
static constexpr u32 SPG_WARP_MAXPROD = 128; // rows with <= this many products go to the warp kernel
static constexpr u32 SPG_MID_TS = 1024;      // slots per row table of the two-warp kernel
static constexpr u32 SPG_MID_MAXPROD = 704;  // <= 75 % of SPG_MID_TS distinct columns even if every product is one: never overflows
static constexpr int SPG_MID_THREADS = 64;
static constexpr u32 SPG_BLOCK_TS = 2048;    // slots per CTA-row table
static constexpr int SPG_BLOCK_THREADS = 256;
static constexpr int SPG_WARPS_PER_CTA = 8;

struct SpgemmArgs
{
    const int64_t *a_rowptr; const u32 *a_col; const u32 *a_pos;       // rows of the left operand (CSR)
    const u32 *at_ptr; const uint2 *at_ent;                            // right operand by column (CSC), rows ascending: 32-bit column
                                                                       // pointers and {row, pos} entries side by side (one sector per column)
    u32 nrows;
    int seed_count;
    u32 *t_col; int32_t *t_num; u32 *t_seeds; u64 cap;                 // unordered row storage
    u64 *counters;          // [0] output cursor, [1] nnz before prune, [2] overflow-list cursor
    u64 *row_off; u32 *row_nnz;
};

template <bool BLOCK> __device__ __forceinline__ void group_sync() { if (BLOCK) __syncthreads(); else __syncwarp(); }

// position of read j inside column c (rows ascending, <= UPPER entries)
__device__ __forceinline__ u32 pos_in_column(const SpgemmArgs &A, u32 c, u32 j)
{
    const u32 b = __ldg(A.at_ptr + c), e = __ldg(A.at_ptr + c + 1);
    for (u32 q = b; q < e; ++q) { const uint2 v = __ldg(A.at_ent + q); if (v.x == j) return v.y; }
    return 0;   // unreachable: j was found through this column
}

// One output row.  keys/cnt/tmin/tmax/sortbuf: TS entries each; ctl: 8 words.  Returns false on table overflow.
template <bool BLOCK>
__device__ bool spgemm_row(const SpgemmArgs &A, u32 row, u32 tid, u32 nth,
                           u32 *keys, u32 *cnt, u32 *tmin, u32 *tmax, u32 *sortbuf, u32 TS, volatile u32 *ctl)
{
    const u32 mask = TS - 1;
    const u32 limit = TS - (TS >> 2);          // 75 % load
    const int shift = 32 - (31 - __clz(TS));
    for (u32 s = tid; s < TS; s += nth) { keys[s] = EMPTY32; cnt[s] = 0; tmin[s] = EMPTY32; tmax[s] = 0; }
    if (tid < 8) ctl[tid] = 0;
    group_sync<BLOCK>();

    const int64_t rs = A.a_rowptr[row];
    const u32 nnz = (u32)(A.a_rowptr[row + 1] - rs);
    // ncu (profiles/r1_v7_spgemm.md): 9 % issue active, everything waits on three dependent random loads per nonzero
    // (column id -> column pointers -> entries).  Four nonzeros per thread are in flight at every step of that chain.
    constexpr int MLP = 4;
    for (u32 tb = tid; tb < nnz; tb += MLP * nth)
    {
        u32 c[MLP], qb[MLP], qe[MLP]; uint2 e0[MLP];
#pragma unroll
        for (int i = 0; i < MLP; ++i) { const u32 t = tb + i * nth; c[i] = t < nnz ? __ldg(A.a_col + rs + t) : EMPTY32; }
#pragma unroll
        for (int i = 0; i < MLP; ++i)
        {
            qb[i] = qe[i] = 0;
            if (c[i] != EMPTY32) { qb[i] = __ldg(A.at_ptr + c[i]); qe[i] = __ldg(A.at_ptr + c[i] + 1); }
        }
#pragma unroll
        for (int i = 0; i < MLP; ++i) { e0[i] = make_uint2(0, 0); if (qb[i] < qe[i]) e0[i] = __ldg(A.at_ent + qb[i]); }
#pragma unroll
        for (int i = 0; i < MLP; ++i)
        {
            const u32 t = tb + i * nth;
            for (u32 q = qb[i]; q < qe[i]; ++q)
            {
                const u32 j = q == qb[i] ? e0[i].x : __ldg(A.at_ent + q).x;
                u32 h = (j * 0x9E3779B1u) >> shift;
                bool placed = false;
                while (ctl[2] == 0)
                {
                    u32 kk = ((volatile u32*)keys)[h];
                    if (kk == j) { placed = true; break; }
                    if (kk == EMPTY32)
                    {
                        u32 prev = atomicCAS(&keys[h], EMPTY32, j);
                        if (prev == EMPTY32)
                        {
                            u32 d = atomicAdd((u32*)&ctl[0], 1u);
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
                    if (t < ((volatile u32*)tmin)[h]) atomicMin(&tmin[h], t);
                    if (t > ((volatile u32*)tmax)[h]) atomicMax(&tmax[h], t);
                }
            }
        }
    }
    group_sync<BLOCK>();
    if (ctl[2]) { group_sync<BLOCK>(); return false; }

    // survivors (numshared >= 2): Prune(numshared <= 1), src/SharedSeeds.cpp:8
    for (u32 s = tid; s < TS; s += nth)
        if (keys[s] != EMPTY32 && cnt[s] >= 2) { u32 i = atomicAdd((u32*)&ctl[1], 1u); sortbuf[i] = s; }
    group_sync<BLOCK>();
    const u32 n = ctl[1];
    if (n <= 32)
    {
        // the common case (a row keeps a handful of columns): one warp sorts in registers, no barrier per stage
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
        // bitonic sort of slot indices by column id
        for (u32 size = 2; size <= m; size <<= 1)
            for (u32 stride = size >> 1; stride > 0; stride >>= 1)
            {
                for (u32 i = tid; i < (m >> 1); i += nth)
                {
                    u32 lo = 2 * i - (i & (stride - 1)), hi = lo + stride;
                    u32 a = sortbuf[lo], b = sortbuf[hi];
                    u32 ka = a == EMPTY32 ? EMPTY32 : keys[a], kb = b == EMPTY32 ? EMPTY32 : keys[b];
                    bool up = (lo & size) == 0;
                    if ((ka > kb) == up) { sortbuf[lo] = b; sortbuf[hi] = a; }
                }
                group_sync<BLOCK>();
            }
    }
    if (tid == 0)
    {
        u64 off = n ? atomicAdd(&A.counters[0], (u64)n) : 0;
        atomicAdd(&A.counters[1], (u64)ctl[0]);
        A.row_off[row] = off; A.row_nnz[row] = n;
        ctl[4] = (u32)off; ctl[5] = (u32)(off >> 32);
    }
    group_sync<BLOCK>();
    const u64 off = ((u64)ctl[5] << 32) | ctl[4];
    for (u32 e = tid; e < n; e += nth)
    {
        u32 s = sortbuf[e];
        u32 j = keys[s], t0 = tmin[s], t1 = tmax[s];
        u64 o = off + e;
        if (o < A.cap)
        {
            u32 c0 = A.a_col[rs + t0], c1 = A.a_col[rs + t1];
            uint4 sd;
            sd.x = A.a_pos[rs + t0]; sd.y = pos_in_column(A, c0, j);
            if (A.seed_count > 1) { sd.z = A.a_pos[rs + t1]; sd.w = pos_in_column(A, c1, j); } else { sd.z = 0; sd.w = 0; }
            A.t_col[o] = j; A.t_num[o] = (int32_t)cnt[s];
            reinterpret_cast<uint4*>(A.t_seeds)[o] = sd;
        }
    }
    group_sync<BLOCK>();
    return true;
}

// bin rows by product count
__global__ void k_spgemm_bin(const u64 *__restrict__ prod, u32 nrows, u32 *__restrict__ small_rows, u32 *__restrict__ mid_rows, u32 *__restrict__ big_rows,
                             u32 *__restrict__ nbins /*[0] small, [1] big, [2] mid; [4..5] max products (u64)*/, u64 *__restrict__ row_off, u32 *__restrict__ row_nnz,
                             u64 *__restrict__ maxprod, u32 mid_max, u32 small_max = SPG_WARP_MAXPROD)
{
    u32 r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= nrows) return;
    u64 p = prod[r];
    if (p == 0) { row_off[r] = 0; row_nnz[r] = 0; return; }
    if (p <= small_max) small_rows[atomicAdd(&nbins[0], 1u)] = r;
    else if (p <= mid_max) mid_rows[atomicAdd(&nbins[2], 1u)] = r;
    else { big_rows[atomicAdd(&nbins[1], 1u)] = r; atomicMax(maxprod, p); }
}

__global__ void __launch_bounds__(32 * SPG_WARPS_PER_CTA) k_spgemm_warp(SpgemmArgs A, const u32 *__restrict__ rows, const u32 *__restrict__ nrows_p)
{
    __shared__ u32 s_tab[SPG_WARPS_PER_CTA][5 * SPG_WARP_TS];
    __shared__ u32 s_ctl[SPG_WARPS_PER_CTA][8];
    u32 w = threadIdx.x >> 5, lane = threadIdx.x & 31;
    u32 n = *nrows_p;
    u32 *t = s_tab[w];
    for (u32 i = blockIdx.x * SPG_WARPS_PER_CTA + w; i < n; i += gridDim.x * SPG_WARPS_PER_CTA)
        spgemm_row<false>(A, rows[i], lane, 32, t, t + SPG_WARP_TS, t + 2 * SPG_WARP_TS, t + 3 * SPG_WARP_TS, t + 4 * SPG_WARP_TS, SPG_WARP_TS, s_ctl[w]);
}

__global__ void __launch_bounds__(SPG_BLOCK_THREADS) k_spgemm_block(SpgemmArgs A, const u32 *__restrict__ rows, const u32 *__restrict__ nrows_p,
                                                                    u32 *__restrict__ overflow_rows)
{
    __shared__ u32 s_tab[5 * SPG_BLOCK_TS];
    __shared__ u32 s_ctl[8];
    u32 n = *nrows_p;
    for (u32 i = blockIdx.x; i < n; i += gridDim.x)
    {
        u32 row = rows[i];
        bool ok = spgemm_row<true>(A, row, threadIdx.x, SPG_BLOCK_THREADS, s_tab, s_tab + SPG_BLOCK_TS, s_tab + 2 * SPG_BLOCK_TS,
                                   s_tab + 3 * SPG_BLOCK_TS, s_tab + 4 * SPG_BLOCK_TS, SPG_BLOCK_TS, s_ctl);
        if (!ok && threadIdx.x == 0) overflow_rows[atomicAdd(&A.counters[2], 1ull)] = row;
    }
}

// two warps per medium row: the table can never overflow (products <= SPG_MID_MAXPROD)
__global__ void __launch_bounds__(SPG_MID_THREADS) k_spgemm_mid(SpgemmArgs A, const u32 *__restrict__ rows, const u32 *__restrict__ nrows_p)
{
    __shared__ u32 s_tab[5 * SPG_MID_TS];
    __shared__ u32 s_ctl[8];
    u32 n = *nrows_p;
    for (u32 i = blockIdx.x; i < n; i += gridDim.x)
        spgemm_row<true>(A, rows[i], threadIdx.x, SPG_MID_THREADS, s_tab, s_tab + SPG_MID_TS, s_tab + 2 * SPG_MID_TS,
                         s_tab + 3 * SPG_MID_TS, s_tab + 4 * SPG_MID_TS, SPG_MID_TS, s_ctl);
}

// rows whose distinct-column count exceeded the shared-memory table: one global-memory table per CTA
__global__ void __launch_bounds__(SPG_BLOCK_THREADS) k_spgemm_global(SpgemmArgs A, const u32 *__restrict__ rows, u32 n, u32 *__restrict__ scratch, u32 TS)
{
    __shared__ u32 s_ctl[8];
    u32 *t = scratch + (size_t)blockIdx.x * 5 * TS;
    for (u32 i = blockIdx.x; i < n; i += gridDim.x)
        spgemm_row<true>(A, rows[i], threadIdx.x, SPG_BLOCK_THREADS, t, t + TS, t + 2 * (size_t)TS, t + 3 * (size_t)TS, t + 4 * (size_t)TS, TS, s_ctl);
}

// unordered row storage -> CSR order
__global__ void k_gather_B(const u64 *__restrict__ row_off, const u32 *__restrict__ row_nnz, const int64_t *__restrict__ b_rowptr, u32 nrows,
                           const u32 *__restrict__ t_col, const int32_t *__restrict__ t_num, const u32 *__restrict__ t_seeds,
                           u32 *__restrict__ b_col, int32_t *__restrict__ b_num, u32 *__restrict__ b_seeds)
{
    u32 warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (warp >= nrows) return;
    u64 src = row_off[warp]; u32 n = row_nnz[warp]; int64_t dst = b_rowptr[warp];
    for (u32 e = lane; e < n; e += 32)
    {
        b_col[dst + e] = t_col[src + e]; b_num[dst + e] = t_num[src + e];
        reinterpret_cast<uint4*>(b_seeds)[dst + e] = reinterpret_cast<const uint4*>(t_seeds)[src + e];
    }
}

__global__ void k_u32_to_u64(const u32 *__restrict__ in, u64 n, u64 *__restrict__ out)
{
    u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    if (i <= n) out[i] = i < n ? in[i] : 0;
}

} // namespace elba
