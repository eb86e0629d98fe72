// Result digests: order-independent 64-bit multiset hashes of what the path produces, computed where the results lie.
// They let a benchmark prove WHAT it timed without moving the results to the host: the digest of a run on 8 GPUs must equal
// the digest of the same input on 1 GPU and the value the CPU oracle gives (tests/common.py restates the same arithmetic in
// numpy).  Every entry is hashed with GLOBAL ids and the hashes are summed mod 2^64, so the value does not depend on how
// the entries are distributed or ordered.
//   kmers  : reliable k-mer, count                      (get_kmer_count_map_keys/values, src/KmerOps.cpp:18-350)
//   A      : global read, column id, position           (create_kmer_matrix, src/KmerOps.cpp:361-401)
//   B      : global row, global column, numshared       (create_seed_matrix + Prune, src/SharedSeeds.cpp:4-10)
//   seeds  : global row, global column, the four seed positions (include/SharedSeeds.hpp:94-95)
#pragma once
#include "common.cuh"
#include "count_smem.cuh"

namespace elba {

static constexpr u64 DG_C1 = 0x9E3779B97F4A7C15ull, DG_C2 = 0xC2B2AE3D27D4EB4Full, DG_C3 = 0x165667B19E3779F9ull;

__host__ __device__ __forceinline__ u64 dg_kmer(u64 kmer, u32 count) { return mix64(kmer ^ mix64((u64)count + DG_C1)); }
__host__ __device__ __forceinline__ u64 dg_a(u64 grow, u32 col, u32 pos) { return mix64(mix64((grow << 32) | col) ^ ((u64)pos + DG_C2)); }
__host__ __device__ __forceinline__ u64 dg_b(u64 grow, u64 gcol, u32 num) { return mix64(mix64((grow << 32) | gcol) + (u64)num * DG_C3); }
__host__ __device__ __forceinline__ u64 dg_seeds(u64 grow, u64 gcol, u32 s0, u32 s1, u32 s2, u32 s3)
{
    return mix64(mix64(mix64((grow << 32) | gcol) ^ (((u64)s0 << 32) | s1)) ^ ((((u64)s2 << 32) | s3) + DG_C1));
}

__device__ __forceinline__ void dg_flush(u64 *__restrict__ out, u64 v)
{
    for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if ((threadIdx.x & 31) == 0 && v) atomicAdd(out, v);
}

__global__ void __launch_bounds__(256) k_digest_kmers(const u64 *__restrict__ kmer, const u32 *__restrict__ cnt, u64 n, u64 *__restrict__ out)
{
    u64 acc = 0;
    for (u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (u64)gridDim.x * blockDim.x) acc += dg_kmer(kmer[i], cnt[i]);
    dg_flush(out, acc);
}

// one warp per row of a CSR matrix with a 32-bit payload
__global__ void __launch_bounds__(256) k_digest_A(const int64_t *__restrict__ rowptr, const u32 *__restrict__ col, const u32 *__restrict__ pos, u32 nrows, u64 row0,
                                                  u64 *__restrict__ out)
{
    const u32 lane = threadIdx.x & 31;
    u64 acc = 0;
    for (u32 r = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; r < nrows; r += (gridDim.x * blockDim.x) >> 5)
    {
        const int64_t b = rowptr[r], e = rowptr[r + 1];
        for (int64_t p = b + lane; p < e; p += 32) acc += dg_a(row0 + r, col[p], pos[p]);
    }
    dg_flush(out, acc);
}

__global__ void __launch_bounds__(256) k_digest_B(const int64_t *__restrict__ rowptr, const u32 *__restrict__ col, const int32_t *__restrict__ num, const u32 *__restrict__ seeds,
                                                  u32 nrows, u64 row0, u64 col0, u64 *__restrict__ out_b, u64 *__restrict__ out_s)
{
    const u32 lane = threadIdx.x & 31;
    u64 acc_b = 0, acc_s = 0;
    for (u32 r = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; r < nrows; r += (gridDim.x * blockDim.x) >> 5)
    {
        const int64_t b = rowptr[r], e = rowptr[r + 1];
        for (int64_t p = b + lane; p < e; p += 32)
        {
            const u64 gr = row0 + r, gc = col0 + col[p];
            const uint4 s = reinterpret_cast<const uint4*>(seeds)[p];
            acc_b += dg_b(gr, gc, (u32)num[p]);
            acc_s += dg_seeds(gr, gc, s.x, s.y, s.z, s.w);
        }
    }
    dg_flush(out_b, acc_b); dg_flush(out_s, acc_s);
}

} // namespace elba
