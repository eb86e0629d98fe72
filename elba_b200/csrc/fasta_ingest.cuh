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
//   lane       = 4 bases -> one byte; the line-break offset i / bases is one 32-bit division per lane and step
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
__global__ void __launch_bounds__(256) k_fasta_pack(FastaView fv, u64 item_begin, u64 item_end, uint8_t *__restrict__ arena)
{
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
#pragma unroll 1
        for (u64 b = b0 + lane; b < bend; b += 32)
        {
            const u64 i0 = b << 2;                                                 // first base of this byte
            u64 q, rem;                                                            // i0 = q * bases + rem
            if (((i0 | bases) >> 32) == 0) { q = (u32)i0 / (u32)bases; rem = (u32)i0 - (u32)q * (u32)bases; }
            else { q = i0 / bases; rem = i0 - q * bases; }
            u64 p = i0 + q;                                                        // input byte of base i0
            const u32 nb = (u32)min((u64)4, len - i0);
            u32 byte = 0;
#pragma unroll
            for (u32 j = 0; j < 4; ++j)
            {
                if (j < nb)
                {
                    const u32 code = fasta_code(__ldg(in + p));
                    byte |= (code << (6 - 2 * j)) & 0xFFu;
                    ++p; if (++rem == bases) { rem = 0; ++p; }                     // the line ends: skip its separator
                }
            }
            out[b] = (uint8_t)byte;
        }
    }
}

} // namespace elba
