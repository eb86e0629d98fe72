// A (reads x reliable k-mers) straight into device CSR + CSC.
// Reference: create_kmer_matrix (src/KmerOps.cpp:361-401) builds (row, col, pos) triples and lets the
// CombBLAS constructor (SumDuplicates=false -> maximum) merge duplicates: the LARGEST position of a
// k-mer inside a read survives.  Transpose: src/main.cpp:272-273.
// Here: seeds sorted by (read, column) [radix sort], last-of-run keeps the run maximum, compaction,
// row pointers by binary search; the CSC is the same entries sorted by (column, read).
#pragma once
#include "common.cuh"

namespace elba {

// Keys are packed (major << minor_bits | minor).
// flag[i] = 1 iff i is the last entry of its (read, column) run; flag[n] = 0 so that an exclusive scan over n+1
// items leaves the survivor count in idx[n].
__global__ void k_mark_run_ends(const u64 *__restrict__ key, u64 n, u64 *__restrict__ flag)
{
    u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    if (i > n) return;
    flag[i] = (i < n) && ((i + 1 == n) || (key[i + 1] != key[i]));
}

// survivors: key, max position of the run
__global__ void k_dedupe_write(const u64 *__restrict__ key, const u32 *__restrict__ pos, const u64 *__restrict__ idx, u64 n,
                               u64 *__restrict__ okey, u32 *__restrict__ opos)
{
    u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    u64 kk = key[i];
    if (i + 1 < n && key[i + 1] == kk) return;
    u32 best = pos[i];
    for (u64 j = i; j > 0 && key[j - 1] == kk; --j) best = max(best, pos[j - 1]);
    u64 o = idx[i];
    okey[o] = kk; opos[o] = best;
}

// ptr[r] = first index with (key >> shift) >= r, r in [0, nseg]
__global__ void k_segment_ptr(const u64 *__restrict__ key, u64 n, u64 nseg, int shift, int64_t *__restrict__ ptr)
{
    u64 r = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    if (r > nseg) return;
    u64 lo = 0, hi = n;
    while (lo < hi) { u64 mid = (lo + hi) >> 1; if ((key[mid] >> shift) < r) lo = mid + 1; else hi = mid; }
    ptr[r] = (int64_t)lo;
}

// key = (major << minor_bits | minor): minor out; optionally the key with the two fields swapped
// (minor << major_bits | major) for the transposed sort.
__global__ void k_split_swap(const u64 *__restrict__ key, u64 n, int minor_bits, int major_bits, u32 *__restrict__ minor, u64 *__restrict__ swapped)
{
    u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    u64 kk = key[i];
    u64 mn = kk & ((1ull << minor_bits) - 1), mj = kk >> minor_bits;
    if (minor) minor[i] = (u32)mn;
    if (swapped) swapped[i] = (mn << major_bits) | mj;
}

// per row: products_i = sum over its columns of the column length; total F accumulates
__global__ void k_row_products(const int64_t *__restrict__ rowptr, const u32 *__restrict__ col, const int64_t *__restrict__ colptr,
                               u32 nrows, u64 *__restrict__ prod, u64 *__restrict__ total)
{
    u32 warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (warp >= nrows) return;
    int64_t b = rowptr[warp], e = rowptr[warp + 1];
    u64 s = 0;
    for (int64_t p = b + lane; p < e; p += 32) { u32 c = col[p]; s += (u64)(colptr[c + 1] - colptr[c]); }
    for (int o = 16; o; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if (lane == 0) { prod[warp] = s; if (s) atomicAdd(total, s); }
}

// the same with the 32-bit column pointers of the column-major operand
__global__ void k_row_products32(const int64_t *__restrict__ rowptr, const u32 *__restrict__ col, const u32 *__restrict__ cptr,
                                 u32 nrows, u64 *__restrict__ prod, u64 *__restrict__ total)
{
    u32 warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (warp >= nrows) return;
    int64_t b = rowptr[warp], e = rowptr[warp + 1];
    u64 s = 0;
    for (int64_t p = b + lane; p < e; p += 32) { u32 c = col[p]; s += (u64)(__ldg(cptr + c + 1) - __ldg(cptr + c)); }
    for (int o = 16; o; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if (lane == 0) { prod[warp] = s; if (s) atomicAdd(total, s); }
}

// the right operand of the SpGEMM as the kernel reads it: 32-bit column pointers (half the footprint: most of it stays in
// L2) and {row, pos} side by side, so one column is one sector instead of three
__global__ void k_spgemm_operand(const int64_t *__restrict__ colptr, u64 ncol, const u32 *__restrict__ row, const u32 *__restrict__ pos, u64 nnz,
                                 u32 *__restrict__ ptr32, uint2 *__restrict__ ent)
{
    const u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    if (i <= ncol) ptr32[i] = (u32)colptr[i];
    if (i < nnz) ent[i] = make_uint2(row[i], pos[i]);
}

} // namespace elba
