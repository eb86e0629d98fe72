// K-mer counting kernels (reference: src/KmerOps.cpp:18-350, include/KmerOps.hpp:58-136).
//
//   k_prep_reads        per-read chunk / k-mer counts (then exclusive scans)
//   (count_smem.cuh)    the fast path: two-level partition + shared-memory counting
//   k_count_array       open-addressing count table in HBM/L2 over one partition buffer: the exact fallback for a
//                       partition whose sub-buckets overflow (heavy hitters), profiles/r1_count_v0.md
//   k_count_direct      the same, fed straight from the reads (single-partition mode, small inputs)
//   k_table_collect     one table scan: reliable k-mers out, table reset
//   k_lookup_build      reliable k-mer -> column id table
//   k_emit_seeds        sweep 2: (read, column, pos) of every instance of a reliable k-mer
//                       (get_kmer_count_map_values + create_kmer_matrix triples, KmerOps.cpp:283-394)
#pragma once
#include "common.cuh"
#include "count_smem.cuh"

namespace elba {

struct __align__(16) Slot { u64 key; u32 cnt; u32 aux; };

// ------------------------------------------------------------------------------------------
__global__ void k_prep_reads(const u64 *__restrict__ len64, u32 n, int k, int stride,
                             u32 *__restrict__ len32, u64 *__restrict__ chunks, u64 *__restrict__ nk, u64 *__restrict__ nks)
{
    u32 i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i > n) return;
    if (i == n) { chunks[n] = 0; nk[n] = 0; nks[n] = 0; return; }
    u64 l = len64[i];
    u64 c = l >= (u64)k ? l - k + 1 : 0;          // ForeachKmer skips reads shorter than k (KmerOps.hpp:118-119)
    len32[i] = (u32)l;
    nk[i] = c;
    nks[i] = (c + stride - 1) / stride;
    chunks[i] = (c + CHUNK - 1) / CHUNK;
}

// ------------------------------------------------------------------------------------------
// count table: open addressing, linear probing, 16-byte slots {key, count, aux}, `slots` arbitrary (< 2^32).
//
// ncu (profiles/r1_count_v0.md) showed the first version latency-bound on a 4-deep dependent chain
// (stream load -> slot load -> CAS -> atomicAdd-with-return), each link a full L2/DRAM round trip, with
// warps half diverged in the probe loop.  This version has ONE waited round trip per k-mer:
//   prev = CAS(slot.key, EMPTY, kmer)   -> claims an empty slot or returns the resident key
//   RED.ADD slot.count                  -> fire-and-forget, nothing waits on it
// ILP k-mers per thread are in flight at once, the next k-mers are loaded before the atomics are waited on,
// and equal k-mers inside a warp (poly-A runs land next to each other) are merged with match.any first.
// Reliable k-mers are found afterwards by one scan of the table that also resets it (k_table_collect).
struct TableRef { Slot *tab; u32 slots; };

// The count tables hold h = mix64(canonical k-mer) (count_smem.cuh): the slot comes from the LOW word of h, the
// partition digits from the high word.
__device__ __forceinline__ u32 slot_of(u64 h, u32 slots) { return __umulhi((u32)h, slots); }

static constexpr u32 MAX_PROBES = 1u << 14;      // a table this full means the distinct-ratio estimate was wrong: flag, host retries

// returns 1 if this call claimed a new slot (a new distinct k-mer)
__device__ __forceinline__ u32 table_insert_resume(const TableRef &T, u64 h, u32 s, u64 prev, u32 mult, u32 *__restrict__ err)
{
    u32 probes = 0;
    while (prev != EMPTY_H && prev != h)
    {
        if (++probes > MAX_PROBES) { atomicOr(err, 1u); return 0; }
        s = (s + 1 == T.slots) ? 0 : s + 1;
        prev = atomicCAS(&T.tab[s].key, EMPTY_H, h);
    }
    atomicAdd(&T.tab[s].cnt, mult);               // result unused -> RED
    return prev == EMPTY_H;
}

__device__ __forceinline__ u32 table_insert(const TableRef &T, u64 h, u32 mult, u32 *__restrict__ err)
{
    u32 s = slot_of(h, T.slots);
    u64 prev = atomicCAS(&T.tab[s].key, EMPTY_H, h);
    return table_insert_resume(T, h, s, prev, mult, err);
}

// warp-reduce a per-thread tally into one global counter
__device__ __forceinline__ void tally(u64 *__restrict__ g, u32 v)
{
    for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if ((threadIdx.x & 31) == 0 && v) atomicAdd(g, (u64)v);
}

__global__ void k_table_clear(Slot *__restrict__ tab, u64 slots, u64 empty)
{
    u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    u64 step = (u64)gridDim.x * blockDim.x;
    ulonglong2 e; e.x = empty; e.y = 0;
    for (; i < slots; i += step) reinterpret_cast<ulonglong2*>(tab)[i] = e;
}

// single-partition mode: straight from the reads (small inputs)
__global__ void __launch_bounds__(256) k_count_direct(ReadsView rv, int k, int stride, TableRef T, u32 *__restrict__ err, u64 *__restrict__ distinct)
{
    u64 step = (u64)gridDim.x * blockDim.x;
    u64 rounds = (rv.nchunks + step - 1) / step;
    u32 nd = 0;
    for (u64 it = 0; it < rounds; ++it)
    {
        u64 g = it * step + (u64)blockIdx.x * blockDim.x + threadIdx.x;
        ChunkInfo ci;
        if (locate_chunk(rv, g, k, ci))
            foreach_kmer_in_chunk(rv, ci, k, stride, [&](u64 x, u32, int) { nd += table_insert(T, mix64(x), 1u, err); });
    }
    tally(distinct, nd);
}

static constexpr int COUNT_ILP = 4;
__global__ void __launch_bounds__(256) k_count_array(const u64 *__restrict__ kmers, u64 n, TableRef T, u32 *__restrict__ err, u64 *__restrict__ distinct)
{
    const u64 tile = (u64)blockDim.x * COUNT_ILP;
    const u64 ntiles = (n + tile - 1) / tile;
    const int lane = threadIdx.x & 31;
    u32 nd = 0;
    u64 x[COUNT_ILP]; bool valid[COUNT_ILP];
    u64 t = blockIdx.x;
    if (t < ntiles)
    {
#pragma unroll
        for (int j = 0; j < COUNT_ILP; ++j) { u64 i = t * tile + (u64)j * blockDim.x + threadIdx.x; valid[j] = i < n; x[j] = valid[j] ? __ldcs(kmers + i) : 0; }
    }
    for (; t < ntiles; t += gridDim.x)
    {
        u32 mult[COUNT_ILP], s[COUNT_ILP]; u64 prev[COUNT_ILP], cur[COUNT_ILP];
#pragma unroll
        for (int j = 0; j < COUNT_ILP; ++j)
        {
            cur[j] = x[j];
            unsigned act = __ballot_sync(0xffffffffu, valid[j]);
            mult[j] = 0;
            if (valid[j])
            {
                unsigned m = __match_any_sync(act, cur[j]);
                if (lane == __ffs(m) - 1) mult[j] = __popc(m);
            }
        }
#pragma unroll
        for (int j = 0; j < COUNT_ILP; ++j)
            if (mult[j]) { s[j] = slot_of(cur[j], T.slots); prev[j] = atomicCAS(&T.tab[s[j]].key, EMPTY_H, cur[j]); }
        // next tile's k-mers are requested before the atomics above are waited on
        u64 tn = t + gridDim.x;
        if (tn < ntiles)
        {
#pragma unroll
            for (int j = 0; j < COUNT_ILP; ++j) { u64 i = tn * tile + (u64)j * blockDim.x + threadIdx.x; valid[j] = i < n; x[j] = valid[j] ? __ldcs(kmers + i) : 0; }
        }
#pragma unroll
        for (int j = 0; j < COUNT_ILP; ++j)
            if (mult[j]) nd += table_insert_resume(T, cur[j], s[j], prev[j], mult[j], err);
    }
    tally(distinct, nd);
}

// One scan of a table: reliable k-mers (lower <= count <= upper) are appended to the output, every occupied slot is
// reset for the next partition.  counters: [0] R cursor, [1] sum of the reliable counts.
__global__ void __launch_bounds__(256) k_table_collect(Slot *__restrict__ tab, u32 slots, u32 lower, u32 upper,
                                                       u64 *__restrict__ out_key, u32 *__restrict__ out_cnt, u64 *__restrict__ counters, u64 cap)
{
    const u32 step = gridDim.x * blockDim.x;
    const u32 rounds = (slots + step - 1) / step;
    const int lane = threadIdx.x & 31;
    ulonglong2 e; e.x = EMPTY_H; e.y = 0;
    for (u32 it = 0; it < rounds; ++it)
    {
        u32 i = it * step + blockIdx.x * blockDim.x + threadIdx.x;
        bool ok = false; u64 key = EMPTY_H; u32 cnt = 0;
        if (i < slots)
        {
            ulonglong2 v = __ldcg(reinterpret_cast<const ulonglong2*>(tab + i));
            key = v.x; cnt = (u32)v.y;
            if (key != EMPTY_H) { reinterpret_cast<ulonglong2*>(tab)[i] = e; ok = cnt >= lower && cnt <= upper; }
        }
        unsigned m = __ballot_sync(0xffffffffu, ok);
        if (!m) continue;
        u32 sum = ok ? cnt : 0;
        for (int o = 16; o; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
        u64 base = 0;
        if (lane == 0) { base = atomicAdd(&counters[0], (u64)__popc(m)); atomicAdd(&counters[1], (u64)sum); }
        base = __shfl_sync(0xffffffffu, base, 0);
        if (ok) { u64 o = base + __popc(m & ((1u << lane) - 1)); if (o < cap) { out_key[o] = key; out_cnt[o] = cnt; } }
    }
}

// ------------------------------------------------------------------------------------------
// Sweep 2: which instances belong to a reliable k-mer, and to which column.
//
// v1 probed a filter and then a DRAM-resident table per instance inside one divergent loop: 2-3 lanes of a warp
// waiting on DRAM at a time (profiles/r1_count_v1.md, 14.7 ps per instance).  Now two dense stages:
//   k_probe_filter   every instance: h = mix64(k-mer), one load from an L2-resident blocked Bloom filter
//                    (3 bits in one 64-bit word).  The 32 k-mers of a chunk stay in registers, loads go out in
//                    batches of 8 with no branch between them; hits become 16-byte candidate records.
//   k_resolve        every candidate (a few % of the instances): one thread each, all lanes busy, probes the
//                    k-mer -> column table in HBM; false positives of the filter are dropped here.
__device__ __forceinline__ u32 filter_word(u64 h, u32 fwords) { return __umulhi((u32)(h >> 32), fwords); }
__device__ __forceinline__ u64 filter_bits(u64 h) { return (1ull << (h & 63)) | (1ull << ((h >> 6) & 63)) | (1ull << ((h >> 12) & 63)); }
__device__ __forceinline__ u32 lut_slot(u64 h, u32 lslots) { return __umulhi((u32)h, lslots); }

__global__ void k_lookup_build(const u64 *__restrict__ keys, const u32 *__restrict__ cnts, u32 R, Slot *__restrict__ tab, u32 lslots,
                               u64 *__restrict__ filter, u32 fwords)
{
    u32 i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= R) return;
    u64 x = keys[i];
    u64 h = mix64(x);
    if (filter) atomicOr(&filter[filter_word(h, fwords)], filter_bits(h));
    u32 s = lut_slot(h, lslots);
    while (true)
    {
        u64 prev = atomicCAS(&tab[s].key, EMPTY_KEY, x);
        if (prev == EMPTY_KEY) { tab[s].cnt = cnts[i]; tab[s].aux = i; return; }
        s = (s + 1 == lslots) ? 0 : s + 1;
    }
}

struct __align__(16) Candidate { u64 kmer; u32 pos; u32 read; };

static constexpr int PF_THREADS = 256;
static constexpr u32 PF_CHUNK = 2048;       // candidate records a warp reserves at a time (one global atomic per 2048 hits)

// Candidates are appended into warp-private chunks of PF_CHUNK records; the unused tail of a chunk is marked with
// kmer = EMPTY_KEY (k_resolve skips those).  *cursor ends as the number of record slots handed out.
__global__ void __launch_bounds__(PF_THREADS, 2) k_probe_filter(ReadsView rv, int k, int stride, const u64 *__restrict__ filter, u32 fwords,
                                                                Candidate *__restrict__ out, u64 *__restrict__ cursor, u64 cap)
{
    const u64 step = (u64)gridDim.x * blockDim.x;
    const u64 rounds = (rv.nchunks + step - 1) / step;
    const int lane = threadIdx.x & 31;
    u64 cbase = 0; u32 cused = PF_CHUNK;     // warp-uniform
    Candidate hole; hole.kmer = EMPTY_KEY; hole.pos = 0; hole.read = 0;
    for (u64 it = 0; it < rounds; ++it)
    {
        const u64 g = it * step + (u64)blockIdx.x * blockDim.x + threadIdx.x;
        ChunkInfo ci; const bool have = locate_chunk(rv, g, k, ci);
        u64 x[CHUNK]; u32 vmask = 0;
        if (have) foreach_kmer_in_chunk(rv, ci, k, stride, [&](u64 c, u32, int s) { x[s] = c; vmask |= 1u << s; });
        u32 hits = 0;
#pragma unroll
        for (int b = 0; b < CHUNK; b += 8)
        {
            u64 w[8], bits[8];
#pragma unroll
            for (int j = 0; j < 8; ++j)
            {
                u64 h = mix64(x[b + j]);
                bits[j] = filter_bits(h);
                w[j] = (vmask >> (b + j)) & 1u ? __ldg(filter + filter_word(h, fwords)) : 0ull;
            }
#pragma unroll
            for (int j = 0; j < 8; ++j) hits |= ((w[j] & bits[j]) == bits[j] ? 1u : 0u) << (b + j);
        }
        hits &= vmask;
        const u32 n = __popc(hits);
        u32 incl = n;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { u32 t = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += t; }
        const u32 total = __shfl_sync(0xffffffffu, incl, 31);
        if (total == 0) continue;
        if (cused + total > PF_CHUNK)
        {
            for (u32 i = cused + lane; i < PF_CHUNK; i += 32) if (cbase + i < cap) out[cbase + i] = hole;
            if (lane == 0) cbase = atomicAdd(cursor, (u64)PF_CHUNK);
            cbase = __shfl_sync(0xffffffffu, cbase, 0);
            cused = 0;
        }
        u64 base = cbase + cused + (incl - n);
        cused += total;
        if (n)
        {
#pragma unroll
            for (int s = 0; s < CHUNK; ++s)
                if (hits & (1u << s))
                {
                    if (base < cap) { Candidate c; c.kmer = x[s]; c.pos = ci.p0 + s; c.read = ci.read; out[base] = c; }
                    ++base;
                }
        }
    }
    for (u32 i = cused + lane; i < PF_CHUNK; i += 32) if (cbase + i < cap) out[cbase + i] = hole;
}

// candidates -> triples: key = (local read << col_bits | column), val = pos.  One output reservation per CTA round.
__global__ void __launch_bounds__(256) k_resolve(const Candidate *__restrict__ cand, u64 n, const Slot *__restrict__ tab, u32 lslots,
                                                 u64 *__restrict__ out_key, u32 *__restrict__ out_pos, u64 *__restrict__ cursor, u64 cap, int col_bits)
{
    __shared__ u32 s_wcnt[8];
    __shared__ u64 s_base;
    const u64 step = (u64)gridDim.x * blockDim.x;
    const u64 rounds = (n + step - 1) / step;
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    for (u64 it = 0; it < rounds; ++it)
    {
        const u64 i = it * step + (u64)blockIdx.x * blockDim.x + threadIdx.x;
        bool ok = false; u32 col = 0; Candidate c; c.kmer = EMPTY_KEY; c.pos = 0; c.read = 0;
        if (i < n) c = cand[i];
        if (c.kmer != EMPTY_KEY)
        {
            u32 s = lut_slot(mix64(c.kmer), lslots);
            while (true)
            {
                ulonglong2 v = __ldg(reinterpret_cast<const ulonglong2*>(tab + s));
                if (v.x == c.kmer) { col = (u32)(v.y >> 32); ok = true; break; }
                if (v.x == EMPTY_KEY) break;
                s = (s + 1 == lslots) ? 0 : s + 1;
            }
        }
        const unsigned m = __ballot_sync(0xffffffffu, ok);
        if (lane == 0) s_wcnt[w] = __popc(m);
        __syncthreads();
        if (threadIdx.x == 0)
        {
            u32 tot = 0;
#pragma unroll
            for (int j = 0; j < 8; ++j) { u32 t = s_wcnt[j]; s_wcnt[j] = tot; tot += t; }
            s_base = tot ? atomicAdd(cursor, (u64)tot) : 0ull;
        }
        __syncthreads();
        const u64 base = s_base + s_wcnt[w] + __popc(m & ((1u << lane) - 1));
        if (ok && base < cap) { out_key[base] = ((u64)c.read << col_bits) | col; out_pos[base] = c.pos; }
        __syncthreads();
    }
}

// the raw canonical stream (debug / parity of the parse stage)
__global__ void k_kmer_stream(ReadsView rv, int k, u64 *__restrict__ out)
{
    u64 step = (u64)gridDim.x * blockDim.x;
    for (u64 g = (u64)blockIdx.x * blockDim.x + threadIdx.x; g < rv.nchunks; g += step)
    {
        ChunkInfo ci;
        if (!locate_chunk(rv, g, k, ci)) continue;
        foreach_kmer_in_chunk(rv, ci, k, 1, [&](u64 x, u32, int s) { out[ci.kmer_base + s] = x; });
    }
}

} // namespace elba
