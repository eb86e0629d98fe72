// Runs elba_b200/csrc/common.cuh's chunked rolling parse on the CPU (one "thread" per chunk, serially)
// and writes the canonical k-mer stream; tests/test_host_parse.py compares it with the oracle.
#include "cuda_runtime.h"
#include "../elba_b200/csrc/common.cuh"
#include <vector>
#include <cstdio>
using namespace elba;
extern "C" uint64_t host_parse(const uint8_t *buf, const uint64_t *off, const uint64_t *len64, uint32_t n, int k, int stride,
                               uint64_t *out_kmer, uint32_t *out_pos, uint32_t *out_read)
{
    std::vector<u32> len(n); std::vector<u64> cs(n + 1, 0), ks(n + 1, 0);
    for (u32 i = 0; i < n; ++i) { len[i] = (u32)len64[i]; u64 c = len64[i] >= (u64)k ? len64[i] - k + 1 : 0; ks[i + 1] = ks[i] + c; cs[i + 1] = cs[i] + (c + CHUNK - 1) / CHUNK; }
    ReadsView rv{buf, reinterpret_cast<const u64*>(off), len.data(), cs.data(), ks.data(), n, cs[n]};
    uint64_t m = 0;
    for (u64 g = 0; g < rv.nchunks; ++g)
    {
        ChunkInfo ci;
        if (!locate_chunk(rv, g, k, ci)) continue;
        foreach_kmer_in_chunk(rv, ci, k, stride, [&](u64 x, u32 p, int) { out_kmer[m] = x; out_pos[m] = p; out_read[m] = ci.read; ++m; });
    }
    return m;
}
