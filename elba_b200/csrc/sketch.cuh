// The reference's two sizing sketches, bit-exact on the device.
//   HyperLogLog  src/HyperLogLog.cpp:12-50 driven by KmerEstimateHandler (include/KmerOps.hpp:58-69):
//                hashes the k-byte ASCII string of the canonical k-mer with MurmurHash3_x64_128 seed 313.
//   Bloom        src/Bloom.cpp:44-73: a,b from four chained murmur3 over the 8 raw k-mer bytes, probes (a+i*b) % bits.
// Neither influences the result of the path when LOWER >= 2; they are exposed for
// parity and for callers that size by estimate as the reference does.
#pragma once
#include "common.cuh"

namespace elba {

// bytes 8w .. 8w+7 of the ASCII string of k-mer x (bytes past k are zero), little-endian word
__device__ __forceinline__ u64 ascii_word(u64 x, int w, int k)
{
    u64 out = 0;
#pragma unroll
    for (int b = 0; b < 8; ++b)
    {
        int i = 8 * w + b;
        if (i < k)
        {
            u32 c = (u32)(x >> (62 - 2 * i)) & 3u;
            u64 ch = (0x54474341u >> (8 * c)) & 0xffu;          // "ACGT"[c]
            out |= ch << (8 * b);
        }
    }
    return out;
}

// MurmurHash3_x64_128(ascii(x), k, 313), first word  (src/HashFuncs.cpp:40-117)
__device__ __forceinline__ u64 murmur3_ascii(u64 x, int k)
{
    const u64 C1 = 0x87c37b91114253d5ull, C2 = 0x4cf5ad432745937full;
    u64 h1 = 313, h2 = 313;
    int nb = k >> 4;
    for (int i = 0; i < nb; ++i)
    {
        u64 k1 = ascii_word(x, 2 * i, k), k2 = ascii_word(x, 2 * i + 1, k);
        k1 *= C1; k1 = rotl64(k1, 31); k1 *= C2; h1 ^= k1; h1 = rotl64(h1, 27); h1 += h2; h1 = h1 * 5 + 0x52dce729;
        k2 *= C2; k2 = rotl64(k2, 33); k2 *= C1; h2 ^= k2; h2 = rotl64(h2, 31); h2 += h1; h2 = h2 * 5 + 0x38495ab5;
    }
    int rem = k & 15;
    if (rem > 8) { u64 k2 = ascii_word(x, 2 * nb + 1, k); k2 *= C2; k2 = rotl64(k2, 33); k2 *= C1; h2 ^= k2; }
    if (rem > 0) { u64 k1 = ascii_word(x, 2 * nb, k);     k1 *= C1; k1 = rotl64(k1, 31); k1 *= C2; h1 ^= k1; }
    h1 ^= (u64)k; h2 ^= (u64)k; h1 += h2; h2 += h1; h1 = fmix64(h1); h2 = fmix64(h2); h1 += h2;
    return h1;
}

__global__ void __launch_bounds__(256) k_hll(ReadsView rv, int k, int stride, u32 *__restrict__ gregs /*4096*/)
{
    __shared__ u32 s_reg[4096];
    for (int i = threadIdx.x; i < 4096; i += blockDim.x) s_reg[i] = 0;
    __syncthreads();
    u64 step = (u64)gridDim.x * blockDim.x;
    for (u64 g = (u64)blockIdx.x * blockDim.x + threadIdx.x; g < rv.nchunks; g += step)
    {
        ChunkInfo ci;
        if (!locate_chunk(rv, g, k, ci)) continue;
        foreach_kmer_in_chunk(rv, ci, k, stride, [&](u64 x, u32, int) {
            u64 h = murmur3_ascii(x, k);
            u32 idx = (u32)(h >> 52);                       // top 12 bits
            u64 w = h << 12;
            u32 rank = min((u32)__clzll((long long)w), 52u) + 1u;   // rho(): src/HyperLogLog.cpp:12-23
            if (rank > s_reg[idx]) atomicMax(&s_reg[idx], rank);
        });
    }
    __syncthreads();
    for (int i = threadIdx.x; i < 4096; i += blockDim.x) if (s_reg[i]) atomicMax(&gregs[i], s_reg[i]);
}

__global__ void __launch_bounds__(256) k_bloom_add(ReadsView rv, int k, int stride, u32 *__restrict__ bf, u64 bits, int hashes)
{
    u64 step = (u64)gridDim.x * blockDim.x;
    for (u64 g = (u64)blockIdx.x * blockDim.x + threadIdx.x; g < rv.nchunks; g += step)
    {
        ChunkInfo ci;
        if (!locate_chunk(rv, g, k, ci)) continue;
        foreach_kmer_in_chunk(rv, ci, k, stride, [&](u64 x, u32, int) {
            u32 a1 = (u32)murmur3_8(x, 0x9747b28cu), a2 = (u32)murmur3_8(x, a1), b1 = (u32)murmur3_8(x, a2), b2 = (u32)murmur3_8(x, b1);
            u64 a = ((u64)a1 << 32) | a2, b = ((u64)b1 << 32) | b2;
            for (int i = 0; i < hashes; ++i)
            {
                u64 p = (a + (u64)i * b) % bits;
                u32 m = 1u << (u32)(p & 31);                 // byte p>>3, bit p%8 of a little-endian word array
                u32 *wd = bf + (p >> 5);
                if (!(__ldg(wd) & m)) atomicOr(wd, m);
            }
        });
    }
}

} // namespace elba
