// Pieces shared by the SpGEMM kernels (spgemm2.cuh): row binning by product count, table geometry of the CTA-per-row
// kernel, the copy of the unordered row storage into CSR order.
// (The row-wise hash SpGEMM of round 1 lived here; its gather of A^T columns read 8.6 x the algorithmic bytes from DRAM,
// profiles/r1_v8_spgemm.md, and was replaced by the expand-by-column / reduce-by-row scheme of spgemm2.cuh.)
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

template <bool BLOCK> __device__ __forceinline__ void group_sync() { if (BLOCK) __syncthreads(); else __syncwarp(); }

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
