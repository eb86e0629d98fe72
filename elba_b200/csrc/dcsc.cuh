// B column-major and doubly compressed: the arrays of CombBLAS' Dcsc<int64_t, SharedSeeds> that the consumer walks
// (src/PairwiseAlignment.cpp:16-56: nzc, cp[], jc[], ir[], numx[]; row ids ascending within a column, LOCAL indices of the rank's
// block), made on the device from the row-major block so that the hand-off needs no sort of tuples on the host.
//   key = column (32 bit), payload = row << 32 | entry   --stable radix sort by column-->   rows ascending inside a column
#pragma once
#include "common.cuh"

namespace elba {

// row of every entry of a CSR matrix (binary search over the row pointers), packed with the entry's index
__global__ void k_dcsc_keys(const int64_t *__restrict__ rowptr, const u32 *__restrict__ col, u32 nrows, u64 nnz, u32 *__restrict__ key, u64 *__restrict__ val)
{
    const u64 e = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= nnz) return;
    u32 lo = 0, hi = nrows;                                   // last row with rowptr[row] <= e
    while (hi - lo > 1) { const u32 mid = (lo + hi) >> 1; if ((u64)__ldg(rowptr + mid) <= e) lo = mid; else hi = mid; }
    key[e] = col[e];
    val[e] = ((u64)lo << 32) | (u32)e;
}

// head[e] = 1 where a column starts in the sorted order (head[nnz] = 0: the scan's total lands there)
__global__ void k_dcsc_heads(const u32 *__restrict__ key, u64 nnz, u64 *__restrict__ head)
{
    const u64 e = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    if (e > nnz) return;
    head[e] = (e < nnz && (e == 0 || key[e] != key[e - 1])) ? 1ull : 0ull;
}

// rank[e] = exclusive scan of head: the e-th sorted entry belongs to nonempty column number rank[e] (if it is its head)
__global__ void k_dcsc_write(const u32 *__restrict__ key, const u64 *__restrict__ val, const u64 *__restrict__ rank, u64 nnz,
                             const int32_t *__restrict__ num, const uint4 *__restrict__ seeds,
                             int64_t *__restrict__ jc, int64_t *__restrict__ cp, int64_t *__restrict__ ir, int32_t *__restrict__ onum, uint4 *__restrict__ oseeds)
{
    const u64 e = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    if (e > nnz) return;
    if (e == nnz) { cp[rank[nnz]] = (int64_t)nnz; return; }
    const u64 v = val[e]; const u32 src = (u32)v;
    ir[e] = (int64_t)(v >> 32);
    onum[e] = num[src];
    oseeds[e] = seeds[src];
    if (e == 0 || key[e] != key[e - 1]) { jc[rank[e]] = (int64_t)key[e]; cp[rank[e]] = (int64_t)e; }
}

} // namespace elba
