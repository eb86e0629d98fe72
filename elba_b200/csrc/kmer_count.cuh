// K-mer counting kernels (reference: src/KmerOps.cpp:18-350, include/KmerOps.hpp:58-136).
//
//   k_prep_reads        per-read chunk / k-mer counts (then exclusive scans)
//   k_part_hist         sweep 0: canonical k-mers -> partition histogram          (partitioned mode)
//   k_part_scatter      sweep 1: canonical k-mers -> partition buffers, tile-sorted in shared memory,
//                       written in coalesced runs (the reference's per-owner buckets + Alltoallv pack,
//                       KmerOps.cpp:99-151, as one kernel)
//   k_count_array       open-addressing count table in HBM/L2 over one partition buffer
//   k_count_direct      the same, fed straight from the reads (single-partition mode)
//   k_collect_reliable  candidates (count reached LOWER) -> reliable list filtered by UPPER
//   k_lookup_build      reliable k-mer -> column id table
//   k_emit_seeds        sweep 2: (read, column, pos) of every instance of a reliable k-mer
//                       (get_kmer_count_map_values + create_kmer_matrix triples, KmerOps.cpp:283-394)
#pragma once
#include "common.cuh"

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
// count table
__device__ __forceinline__ Slot ld_slot(const Slot *p)
{
    // one 16-byte L2 load (L1 is useless for random table probes)
    ulonglong2 v = __ldcg(reinterpret_cast<const ulonglong2*>(p));
    Slot s; s.key = v.x; s.cnt = (u32)v.y; s.aux = (u32)(v.y >> 32); return s;
}

// Insert/increment.  cand/ncand: slots whose count just reached `lower` are appended (exactly once per key).
__device__ __forceinline__ u32 table_add(Slot *__restrict__ tab, u64 mask, u64 kmer, u32 lower, u32 upper,
                                         u32 *__restrict__ cand, u32 *__restrict__ ncand, u32 cand_cap)
{
    u64 s = slot_hash(kmer) & mask;
    u32 claimed = 0;
    while (true)
    {
        Slot v = ld_slot(tab + s);
        if (v.key == EMPTY_KEY)
        {
            u64 prev = atomicCAS(&tab[s].key, EMPTY_KEY, kmer);
            claimed = (prev == EMPTY_KEY);
            v.key = claimed ? kmer : prev;
            v.cnt = 0;
        }
        if (v.key == kmer)
        {
            // counts only matter up to upper+1: stop hammering hot keys (poly-A) once saturated
            if (v.cnt <= upper)
            {
                u32 old = atomicAdd(&tab[s].cnt, 1u);
                if (old + 1 == lower)
                {
                    u32 i = atomicAdd(ncand, 1u);
                    if (i < cand_cap) cand[i] = (u32)s;
                }
            }
            return claimed;
        }
        s = (s + 1) & mask;
    }
}

// warp-reduce a per-thread tally into one global counter
__device__ __forceinline__ void tally(u64 *__restrict__ g, u32 v)
{
    for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if ((threadIdx.x & 31) == 0 && v) atomicAdd(g, (u64)v);
}

__global__ void k_table_clear(Slot *__restrict__ tab, u64 slots)
{
    u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    u64 step = (u64)gridDim.x * blockDim.x;
    ulonglong2 e; e.x = EMPTY_KEY; e.y = 0;
    for (; i < slots; i += step) reinterpret_cast<ulonglong2*>(tab)[i] = e;
}

__global__ void __launch_bounds__(256) k_count_direct(ReadsView rv, int k, int stride, Slot *__restrict__ tab, u64 mask,
                                                      u32 lower, u32 upper, u32 *__restrict__ cand, u32 *__restrict__ ncand, u32 cand_cap,
                                                      u64 *__restrict__ distinct)
{
    u64 step = (u64)gridDim.x * blockDim.x;
    u64 rounds = (rv.nchunks + step - 1) / step;
    u32 nd = 0;
    for (u64 it = 0; it < rounds; ++it)
    {
        u64 g = it * step + (u64)blockIdx.x * blockDim.x + threadIdx.x;
        ChunkInfo ci;
        if (locate_chunk(rv, g, k, ci))
            foreach_kmer_in_chunk(rv, ci, k, stride, [&](u64 x, u32, int) { nd += table_add(tab, mask, x, lower, upper, cand, ncand, cand_cap); });
    }
    tally(distinct, nd);
}

__global__ void __launch_bounds__(256) k_count_array(const u64 *__restrict__ kmers, u64 n, Slot *__restrict__ tab, u64 mask,
                                                     u32 lower, u32 upper, u32 *__restrict__ cand, u32 *__restrict__ ncand, u32 cand_cap,
                                                     u64 *__restrict__ distinct)
{
    u64 step = (u64)gridDim.x * blockDim.x;
    u64 rounds = (n + step - 1) / step;
    u32 nd = 0;
    for (u64 it = 0; it < rounds; ++it)
    {
        u64 i = it * step + (u64)blockIdx.x * blockDim.x + threadIdx.x;
        if (i < n)
        {
            u64 x = __ldcs(kmers + i);        // streaming: do not displace the table from L2
            nd += table_add(tab, mask, x, lower, upper, cand, ncand, cand_cap);
        }
    }
    tally(distinct, nd);
}

// candidates -> reliable (count <= upper); also accumulates the instance total of the reliable k-mers
__global__ void k_collect_reliable(const Slot *__restrict__ tab, const u32 *__restrict__ cand, const u32 *__restrict__ ncand_p, u32 upper,
                                   u64 *__restrict__ out_key, u32 *__restrict__ out_cnt, u64 *__restrict__ counters /*[0]=R cursor, [1]=sum cnt*/, u64 cap)
{
    u32 i = blockIdx.x * blockDim.x + threadIdx.x;
    u32 ncand = *ncand_p;
    bool ok = false; Slot v; v.key = 0; v.cnt = 0;
    if (i < ncand) { v = ld_slot(tab + cand[i]); ok = v.cnt <= upper; }
    unsigned m = __ballot_sync(0xffffffffu, ok);
    if (!m) return;
    int lane = threadIdx.x & 31;
    u32 sum = ok ? v.cnt : 0;
    for (int o = 16; o; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
    u64 base = 0;
    if (lane == 0) { base = atomicAdd(&counters[0], (u64)__popc(m)); atomicAdd(&counters[1], (u64)sum); }
    base = __shfl_sync(0xffffffffu, base, 0);
    if (ok) { u64 o = base + __popc(m & ((1u << lane) - 1)); if (o < cap) { out_key[o] = v.key; out_cnt[o] = v.cnt; } }
}

// ------------------------------------------------------------------------------------------
// partitioning
// partition = high bits of the hash (the table slot uses the low bits)
__device__ __forceinline__ u32 part_of(u64 x, u32 P) { return (u32)__umul64hi(slot_hash(x), (u64)P); }

__global__ void __launch_bounds__(256) k_part_hist(ReadsView rv, int k, int stride, u32 P, u64 *__restrict__ ghist)
{
    extern __shared__ u32 s_hist[];
    for (u32 i = threadIdx.x; i < P; i += blockDim.x) s_hist[i] = 0;
    __syncthreads();
    u64 step = (u64)gridDim.x * blockDim.x;
    for (u64 g = (u64)blockIdx.x * blockDim.x + threadIdx.x; g < rv.nchunks; g += step)
    {
        ChunkInfo ci;
        if (!locate_chunk(rv, g, k, ci)) continue;
        foreach_kmer_in_chunk(rv, ci, k, stride, [&](u64 x, u32, int) { atomicAdd(&s_hist[part_of(x, P)], 1u); });
    }
    __syncthreads();
    for (u32 i = threadIdx.x; i < P; i += blockDim.x) if (s_hist[i]) atomicAdd(&ghist[i], (u64)s_hist[i]);
}

// One tile = SCATTER_BLOCK chunks (SCATTER_BLOCK*32 k-mers).  Shared memory: sorted[tile k-mers] | gbase[P] | cnt[P] | off[P+1]
static constexpr int SCATTER_BLOCK = 256;
__global__ void __launch_bounds__(SCATTER_BLOCK) k_part_scatter(ReadsView rv, int k, int stride, u32 P,
                                                                u64 *__restrict__ gcursor /*[P] running write cursors, pre-set to partition starts*/,
                                                                u64 *__restrict__ out)
{
    extern __shared__ __align__(16) unsigned char s_raw[];
    u64 *s_sorted = reinterpret_cast<u64*>(s_raw);            // [SCATTER_BLOCK*CHUNK]
    u64 *s_gbase = s_sorted + SCATTER_BLOCK * CHUNK;          // [P]
    u32 *s_cnt = reinterpret_cast<u32*>(s_gbase + P);         // [P]
    u32 *s_off = s_cnt + P;                                   // [P+1]
    __shared__ u32 s_total;

    u64 ntiles = (rv.nchunks + SCATTER_BLOCK - 1) / SCATTER_BLOCK;
    for (u64 tile = blockIdx.x; tile < ntiles; tile += gridDim.x)
    {
        for (u32 i = threadIdx.x; i < P; i += blockDim.x) s_cnt[i] = 0;
        __syncthreads();
        u64 g = tile * SCATTER_BLOCK + threadIdx.x;
        ChunkInfo ci; bool have = locate_chunk(rv, g, k, ci);
        // phase A: tile histogram
        if (have) foreach_kmer_in_chunk(rv, ci, k, stride, [&](u64 x, u32, int) { atomicAdd(&s_cnt[part_of(x, P)], 1u); });
        __syncthreads();
        // exclusive scan of s_cnt -> s_off (P <= 4096: serial per-thread blocks + warp 0 scan)
        {
            // each thread scans a contiguous strip of ceil(P/blockDim) entries
            u32 per = (P + blockDim.x - 1) / blockDim.x;
            u32 b = threadIdx.x * per, e = min(b + per, P);
            u32 sum = 0;
            for (u32 i = b; i < e; ++i) sum += s_cnt[i];
            // block exclusive scan of `sum` via shared memory (reuse s_off[0..blockDim) temporarily is unsafe; use warp shuffles)
            __shared__ u32 s_warp[SCATTER_BLOCK / 32];
            u32 lane = threadIdx.x & 31, w = threadIdx.x >> 5;
            u32 incl = sum;
            for (int o = 1; o < 32; o <<= 1) { u32 t = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= (u32)o) incl += t; }
            if (lane == 31) s_warp[w] = incl;
            __syncthreads();
            if (w == 0)
            {
                u32 v = lane < SCATTER_BLOCK / 32 ? s_warp[lane] : 0;
                u32 iv = v;
                for (int o = 1; o < 32; o <<= 1) { u32 t = __shfl_up_sync(0xffffffffu, iv, o); if (lane >= (u32)o) iv += t; }
                if (lane < SCATTER_BLOCK / 32) s_warp[lane] = iv - v;
                if (lane == SCATTER_BLOCK / 32 - 1) s_total = iv;
            }
            __syncthreads();
            u32 run = s_warp[w] + incl - sum;
            for (u32 i = b; i < e; ++i) { s_off[i] = run; run += s_cnt[i]; }
            if (threadIdx.x == 0) s_off[P] = s_total;
        }
        __syncthreads();
        // reserve global space per partition; reset s_cnt as fill cursors
        for (u32 i = threadIdx.x; i < P; i += blockDim.x)
        {
            u32 c = s_cnt[i];
            s_gbase[i] = c ? atomicAdd(&gcursor[i], (u64)c) : 0;
            s_cnt[i] = 0;
        }
        __syncthreads();
        // phase B: place k-mers sorted by partition
        if (have) foreach_kmer_in_chunk(rv, ci, k, stride, [&](u64 x, u32, int) {
            u32 p = part_of(x, P); u32 r = atomicAdd(&s_cnt[p], 1u); s_sorted[s_off[p] + r] = x; });
        __syncthreads();
        // phase C: coalesced runs, one warp per partition
        {
            u32 lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = blockDim.x >> 5;
            for (u32 p = w; p < P; p += nw)
            {
                u32 b = s_off[p], e = s_off[p + 1];
                u64 dst = s_gbase[p];
                for (u32 i = b + lane; i < e; i += 32) out[dst + (i - b)] = s_sorted[i];
            }
        }
        __syncthreads();
    }
}

// ------------------------------------------------------------------------------------------
// reliable k-mer -> column id
__global__ void k_lookup_build(const u64 *__restrict__ keys, const u32 *__restrict__ cnts, u32 R, Slot *__restrict__ tab, u64 mask)
{
    u32 i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= R) return;
    u64 x = keys[i];
    u64 s = slot_hash(x) & mask;
    while (true)
    {
        u64 prev = atomicCAS(&tab[s].key, EMPTY_KEY, x);
        if (prev == EMPTY_KEY) { tab[s].cnt = cnts[i]; tab[s].aux = i; return; }
        s = (s + 1) & mask;
    }
}

__device__ __forceinline__ bool lookup(const Slot *__restrict__ tab, u64 mask, u64 x, u32 &col)
{
    u64 s = slot_hash(x) & mask;
    while (true)
    {
        ulonglong2 v = __ldg(reinterpret_cast<const ulonglong2*>(tab + s));
        if (v.x == x) { col = (u32)(v.y >> 32); return true; }
        if (v.x == EMPTY_KEY) return false;
        s = (s + 1) & mask;
    }
}

// Sweep 2.  Every instance of a reliable k-mer becomes one triple: key = (local read << col_bits | column), val = pos.
// A thread first marks its hits (bitmask over its chunk), the warp reserves output space with ONE atomic,
// then the hits are recomputed from the staged bases and written.
__global__ void __launch_bounds__(256) k_emit_seeds(ReadsView rv, int k, int stride, const Slot *__restrict__ tab, u64 mask,
                                                    u64 *__restrict__ out_key, u32 *__restrict__ out_pos, u64 *__restrict__ cursor, u64 cap, int col_bits)
{
    u64 step = (u64)gridDim.x * blockDim.x;
    u64 rounds = (rv.nchunks + step - 1) / step;
    int lane = threadIdx.x & 31;
    for (u64 it = 0; it < rounds; ++it)
    {
        u64 g = it * step + (u64)blockIdx.x * blockDim.x + threadIdx.x;
        ChunkInfo ci; bool have = locate_chunk(rv, g, k, ci);
        u32 hits = 0;
        if (have) foreach_kmer_in_chunk(rv, ci, k, stride, [&](u64 x, u32, int s) { u32 c; if (lookup(tab, mask, x, c)) hits |= 1u << s; });
        u32 n = __popc(hits);
        u32 incl = n;
        for (int o = 1; o < 32; o <<= 1) { u32 t = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += t; }
        u32 total = __shfl_sync(0xffffffffu, incl, 31);
        if (total == 0) continue;
        u64 base = 0;
        if (lane == 31) base = atomicAdd(cursor, (u64)total);
        base = __shfl_sync(0xffffffffu, base, 31) + (incl - n);
        if (n)
        {
            foreach_kmer_in_chunk(rv, ci, k, stride, [&](u64 x, u32 p, int s) {
                if (hits & (1u << s))
                {
                    u32 c = 0; lookup(tab, mask, x, c);
                    if (base < cap) { out_key[base] = ((u64)ci.read << col_bits) | c; out_pos[base] = p; }
                    base++;
                }
            });
        }
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
