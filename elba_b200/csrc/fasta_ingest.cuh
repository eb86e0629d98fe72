// FASTA ingest on the device: raw FASTA bytes + the .fai records of this rank's reads -> the 2-bit arena.
// (reference: FastaIndex::getmydna, src/FastaIndex.cpp:191-290, which copies every record line by line into a char buffer
//  on ONE core per rank and packs it with DnaSeq::compress, src/DnaSeq.cpp:7-29; code table include/DnaSeq.hpp:136-154)
//
// A record {len, pos, bases} says where a read lies in the file: base i is the byte at pos + i + i / bases (every line of
// `bases` characters is followed by ONE separator byte, FastaIndex.cpp:265-266: `locpos += cnt + 1`).  So every output byte
// can be computed on its own, no scan over the text is needed.  HBM-bound byte work: 1 B/base read, 0.25 B/base written.
//
//   work item  = 512 consecutive output bytes (2048 bases) of ONE read: item_start[r] = exclusive scan of ceil(bytes_r / 512)
//   warp       = one item (found by a binary search over item_start, the same for all lanes: broadcast loads);
//                in every step the 32 lanes produce 32 consecutive output bytes from 128 consecutive bases
//                (129-130 consecutive input bytes with a line break inside): both sides coalesced
//   lane       = 4 bases -> one byte; the lane carries (line, column) of its position from step to step
#pragma once
#include "common.cuh"

namespace elba {

// include/DnaSeq.hpp:136-154: A/a, N/n -> 0, C/c -> 1, G/g -> 2, T/t -> 3, anything else -> 4
__device__ __forceinline__ u32 fasta_code(u32 c)
{
    const u32 u = c & 0xDFu;                       // fold the lower case letters onto the upper case ones
    u32 code = 4u;
    code = (u == 'A' || u == 'N') ? 0u : code;
    code = (u == 'C') ? 1u : code;
    code = (u == 'G') ? 2u : code;
    code = (u == 'T') ? 3u : code;
    return code;
}

static constexpr u32 FI_ITEM_BYTES = 512;          // output bytes per work item

struct FastaView
{
    const uint8_t *raw;        // the chunk [chunk_pos, chunk_pos + chunk_bytes) of the FASTA file, on the device
    const u64 *rec;            // [n][3] len, pos, bases (FastaIndex::Record, include/FastaIndex.hpp:10)
    const u64 *off;            // [n + 1] first arena byte of every read
    const u64 *item_start;     // [n + 1] first work item of every read
    u64 chunk_pos; u32 n;
};

// items [item_begin, item_end) -> arena bytes.  A character outside the table carries code 4; the reference then ORs
// uint8_t(4 << (6 - 2 i)) into the byte (src/DnaSeq.cpp:20-22), which for i > 0 sets the LOW bit of the base before it
// and for i = 0 vanishes: reproduced, so that the arena equals the reference's byte for byte on any input.
//
// Measured first (profiles/r2_v21_ncu_k_fasta_pack_summary.txt): with one 32-bit division per output byte and a compare chain
// per character the kernel issued 41 thread instructions per base (426 GB/s of text, 0.07 of HBM).  Now: the code table sits
// in shared memory (one LDS per character), and a lane keeps (line, column) of its position and advances it by the 128 bases
// a warp step covers (one division per lane and WORK ITEM, not per byte); lines of fewer than 4 bases and reads of 2^32 bases
// or more take the general per-byte path.
__global__ void __launch_bounds__(256) k_fasta_pack(FastaView fv, u64 item_begin, u64 item_end, uint8_t *__restrict__ arena)
{
    __shared__ uint8_t s_code[256];
    s_code[threadIdx.x & 255] = (uint8_t)fasta_code(threadIdx.x & 255);
    __syncthreads();
    const u32 lane = threadIdx.x & 31;
    const u64 nwarps = ((u64)gridDim.x * blockDim.x) >> 5;
    for (u64 it = item_begin + (((u64)blockIdx.x * blockDim.x + threadIdx.x) >> 5); it < item_end; it += nwarps)
    {
        // the read of this item: the last r with item_start[r] <= it
        u32 lo = 0, hi = fv.n;
        while (hi - lo > 1) { const u32 mid = (lo + hi) >> 1; if (__ldg(fv.item_start + mid) <= it) lo = mid; else hi = mid; }
        const u32 r = lo;
        const u64 len = __ldg(fv.rec + 3ull * r), pos = __ldg(fv.rec + 3ull * r + 1) - fv.chunk_pos, bases = __ldg(fv.rec + 3ull * r + 2);
        const u64 nbytes = (len + 3) >> 2;
        const u64 b0 = (it - __ldg(fv.item_start + r)) * FI_ITEM_BYTES;           // first output byte of the item inside the read
        const u64 bend = min(nbytes, b0 + FI_ITEM_BYTES);
        uint8_t *__restrict__ out = arena + __ldg(fv.off + r);
        const uint8_t *__restrict__ in = fv.raw + pos;
        if (bases >= 4 && ((len | bases) >> 32) == 0)
        {
            // at most one line break inside the 4 bases of a lane
            const u32 B = (u32)bases, L = (u32)len;
            const u32 sq = 128u / B, sr = 128u - sq * B;                           // what 128 bases add to (line, column)
            u32 i0 = (u32)(b0 + lane) << 2;                                        // first base of my byte
            u32 q = i0 / B, rem = i0 - q * B;                                      // its line and its column in the line
#pragma unroll 1
            for (u64 b = b0 + lane; b < bend; b += 32)
            {
                const uint8_t *__restrict__ p = in + ((u64)i0 + q);               // the byte of base i0
                const u32 nb = min(4u, L - i0);
                const u32 t = B - rem;                                             // bases of mine before the line ends (>= 1)
                u32 byte = (u32)s_code[__ldg(p)] << 6;
                if (nb > 1) byte |= ((u32)s_code[__ldg(p + 1 + (1u >= t ? 1 : 0))] << 4) & 0xFFu;
                if (nb > 2) byte |= ((u32)s_code[__ldg(p + 2 + (2u >= t ? 1 : 0))] << 2) & 0xFFu;
                if (nb > 3) byte |= (u32)s_code[__ldg(p + 3 + (3u >= t ? 1 : 0))];
                out[b] = (uint8_t)byte;
                i0 += 128u; q += sq; rem += sr;
                if (rem >= B) { rem -= B; ++q; }
            }
            continue;
        }
#pragma unroll 1
        for (u64 b = b0 + lane; b < bend; b += 32)
        {
            const u64 i0 = b << 2;                                                 // first base of this byte
            const u64 q = i0 / bases; u64 rem = i0 - q * bases;                    // i0 = q * bases + rem
            u64 p = i0 + q;                                                        // input byte of base i0
            const u32 nb = (u32)min((u64)4, len - i0);
            u32 byte = 0;
#pragma unroll
            for (u32 j = 0; j < 4; ++j)
            {
                if (j < nb)
                {
                    byte |= ((u32)s_code[__ldg(in + p)] << (6 - 2 * j)) & 0xFFu;
                    ++p; if (++rem == bases) { rem = 0; ++p; }                     // the line ends: skip its separator
                }
            }
            out[b] = (uint8_t)byte;
        }
    }
}

} // namespace elba
