// Shared device helpers: 2-bit read access, rolling canonical k-mers, hashes.
// sm_100a only (no multi-arch dispatch).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace elba {

typedef unsigned long long u64;
typedef uint32_t u32;

// A canonical k-mer can never be all ones: the twin of T^k is A^k = 0, which is smaller
// (src/Kmer.cpp:200-205), and for k < 32 the low 2*(32-k) bits are zero (include/Kmer.hpp).
static constexpr u64 EMPTY_KEY = 0xFFFFFFFFFFFFFFFFull;

// k-mer instances are processed in chunks of CHUNK consecutive window starts of one read.
// CHUNK is a multiple of 4 so every chunk starts on a byte boundary of the arena
// (reads start on byte boundaries, src/DnaBuffer.cpp:22-29).
static constexpr int CHUNK = 32;

struct ReadsView
{
    const uint8_t *buf;        // packed arena (+16 bytes of slack after the end)
    const u64 *off;            // [n] byte offset of read i
    const u32 *len;            // [n] bases in read i
    const u64 *chunk_start;    // [n+1] exclusive scan of chunks per read
    const u64 *kmer_start;     // [n+1] exclusive scan of k-mer instances per read (stride 1 numbering)
    u32 n;
    u64 nchunks;
};

// ---- hashes -------------------------------------------------------------------------------
__device__ __forceinline__ u64 rotl64(u64 v, int r) { return (v << r) | (v >> (64 - r)); }
__device__ __forceinline__ u64 fmix64(u64 v)
{
    v ^= v >> 33; v *= 0xff51afd7ed558ccdull; v ^= v >> 33; v *= 0xc4ceb9fe1a85ec53ull; v ^= v >> 33; return v;
}

// MurmurHash3_x64_128 of an 8-byte key, first 64-bit word: Kmer::GetHash (src/Kmer.cpp:207-213,
// src/HashFuncs.cpp:40-117, 231-236 with seed 313).  Used where the reference's own value matters
// (owner function, Bloom); the tables use the cheaper slot_hash below.
__device__ __forceinline__ u64 murmur3_8(u64 x, u32 seed)
{
    const u64 C1 = 0x87c37b91114253d5ull, C2 = 0x4cf5ad432745937full;
    u64 h1 = seed, h2 = seed;
    u64 k1 = x; k1 *= C1; k1 = rotl64(k1, 31); k1 *= C2; h1 ^= k1;
    h1 ^= 8; h2 ^= 8; h1 += h2; h2 += h1; h1 = fmix64(h1); h2 = fmix64(h2); h1 += h2;
    return h1;
}

// Table / partition hash: two 64-bit multiplies, full avalanche into the high bits.
__device__ __forceinline__ u64 slot_hash(u64 x)
{
    x ^= x >> 32; x *= 0x9E3779B97F4A7C15ull; x ^= x >> 29; x *= 0xBF58476D1CE4E5B9ull; x ^= x >> 32;
    return x;
}

// ---- reads ----------------------------------------------------------------------------------
// read of global chunk g: largest r with chunk_start[r] <= g
__device__ __forceinline__ u32 find_read(const u64 *__restrict__ chunk_start, u32 n, u64 g)
{
    u32 lo = 0, hi = n;            // invariant: chunk_start[lo] <= g < chunk_start[hi]
    while (hi - lo > 1) { u32 mid = (lo + hi) >> 1; if (__ldg(chunk_start + mid) <= g) lo = mid; else hi = mid; }
    return lo;
}

// 64 bases starting at byte address a of the arena, as two big-endian words (base 0 at bit 62 of w0).
__device__ __forceinline__ void load_bases64(const uint8_t *__restrict__ buf, u64 a, u64 &w0, u64 &w1)
{
    const u32 *p = reinterpret_cast<const u32*>(buf + (a & ~3ull));
    u32 sh = (u32)(a & 3) * 8;
    u32 x0 = __ldg(p), x1 = __ldg(p + 1), x2 = __ldg(p + 2), x3 = __ldg(p + 3), x4 = __ldg(p + 4);
    u32 y0 = __funnelshift_r(x0, x1, sh), y1 = __funnelshift_r(x1, x2, sh), y2 = __funnelshift_r(x2, x3, sh), y3 = __funnelshift_r(x3, x4, sh);
    w0 = ((u64)__byte_perm(y0, 0, 0x0123) << 32) | __byte_perm(y1, 0, 0x0123);
    w1 = ((u64)__byte_perm(y2, 0, 0x0123) << 32) | __byte_perm(y3, 0, 0x0123);
}

struct ChunkInfo { u32 read; u32 p0; u32 nk; u64 kmer_base; };

// Locate chunk g (read, first window start, number of window starts in it).
__device__ __forceinline__ bool locate_chunk(const ReadsView &rv, u64 g, int k, ChunkInfo &ci)
{
    if (g >= rv.nchunks) return false;
    u32 r = find_read(rv.chunk_start, rv.n, g);
    u32 len = __ldg(rv.len + r);
    u32 nkr = len - (u32)k + 1;                  // reads with len < k have no chunks, never found here
    u32 p0 = (u32)(g - __ldg(rv.chunk_start + r)) * CHUNK;
    ci.read = r; ci.p0 = p0; ci.nk = min((u32)CHUNK, nkr - p0);
    ci.kmer_base = __ldg(rv.kmer_start + r) + p0;
    return true;
}

// Visit the canonical k-mers of one chunk: f(canonical, pos_in_read, index_in_chunk).
// Rolling forward / reverse-complement windows (GetExtension src/Kmer.cpp:149-165, GetTwin :167-198,
// GetRep :200-205) with shifts only — no table.  stride: only window starts p with p % stride == 0.
template <class F>
__device__ __forceinline__ void foreach_kmer_in_chunk(const ReadsView &rv, const ChunkInfo &ci, int k, int stride, F f)
{
    u64 a = __ldg(rv.off + ci.read) + (ci.p0 >> 2);
    u64 w0, w1;
    load_bases64(rv.buf, a, w0, w1);
    const int lsh = 2 * (32 - k);
    const u64 kmask = (k == 32) ? ~0ull : (~0ull << lsh);
    u64 fwd = w0 & kmask;
    // twin of the first window: complement, reverse the 2-bit codes, left-align
    u64 y = ~fwd;
    y = ((y >> 2) & 0x3333333333333333ull) | ((y & 0x3333333333333333ull) << 2);
    y = ((y >> 4) & 0x0F0F0F0F0F0F0F0Full) | ((y & 0x0F0F0F0F0F0F0F0Full) << 4);
    y = ((u64)__byte_perm((u32)y, 0, 0x0123) << 32) | __byte_perm((u32)(y >> 32), 0, 0x0123);
    u64 rc = y << lsh;
    // bases k .. k+CHUNK-2 of the chunk, left-aligned: the codes that roll in
    u64 nxt = (k == 32) ? w1 : ((w0 << (2 * k)) | (w1 >> (64 - 2 * k)));
    const bool strided = stride > 1;
#pragma unroll
    for (int s = 0; s < CHUNK; ++s)
    {
        if (s < (int)ci.nk)
        {
            u32 p = ci.p0 + s;
            if (!strided || (p % (u32)stride) == 0) f(fwd < rc ? fwd : rc, p, s);
        }
        u64 c = nxt >> 62; nxt <<= 2;
        fwd = (fwd << 2) | (c << lsh);
        rc = ((rc >> 2) | ((3ull - c) << 62)) & kmask;
    }
}

} // namespace elba
