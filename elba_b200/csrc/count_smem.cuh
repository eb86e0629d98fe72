// K-mer counting, version 2: two-level hash partitioning, counting in shared memory.
// (reference: get_kmer_count_map_keys / get_kmer_count_map_values, src/KmerOps.cpp:18-350)
//
// Why (profiles/r1_count_v1.md): an HBM- or L2-resident count table costs two dependent global atomics per
// k-mer instance; measured 27-48 ps per instance on B200, against a budget of 8.7 ps for 50 % of the HBM
// roofline.  Shared-memory atomics are an order of magnitude cheaper, but a table only fits an SM if it serves
// a few thousand k-mers.  So the instances are split twice by digits of ONE bijective 64-bit mix h of the
// canonical k-mer (equal k-mers <=> equal h, so h itself is what is counted; the reliable ones are un-mixed at
// the end):
//
//   k_scatter1   reads -> canonical k-mer -> h -> level-1 partition (P1 <= 4096, ~1 M instances each).
//                The reference's per-owner buckets + Alltoallv pack (KmerOps.cpp:99-151); with several GPUs a
//                partition's owner is a rank and these buffers are what crosses NVLink.
//   k_scatter2   one level-1 partition -> P2 sub-buckets of at most BUCKET_CAP instances (scratch, L2-sized)
//   k_count_buckets  one CTA per sub-bucket: open-addressing table in shared memory (ATOMS.CAS.64 + ATOMS.ADD),
//                then one scan that appends {h, count} of the reliable k-mers (LOWER <= count <= UPPER).
//
// Both scatters stage one tile in shared memory ordered by destination, so global writes are runs.
// Capacities are optimistic (hash partitions are Poisson-tight); an overflow only raises a flag and the host
// redoes that step exactly (level 1: histogram first; level 2: the global-table kernel of kmer_count.cuh).
#pragma once
#include "common.cuh"

namespace elba {

// ---- the bijective mix (slot_hash of common.cuh) and its inverse --------------------------------
__host__ __device__ __forceinline__ u64 mix64(u64 x)
{
    x ^= x >> 32; x *= 0x9E3779B97F4A7C15ull; x ^= x >> 29; x *= 0xBF58476D1CE4E5B9ull; x ^= x >> 32;
    return x;
}
__host__ __device__ __forceinline__ u64 unmix64(u64 x)
{
    x ^= x >> 32; x *= 0x96de1b173f119089ull; x ^= (x >> 29) ^ (x >> 58); x *= 0xf1de83e19937733dull; x ^= x >> 32;
    return x;
}
// mix64(EMPTY_KEY): no canonical k-mer is all ones (common.cuh), so no counted h equals this
static constexpr u64 EMPTY_H = 0x9cebc8ff07279667ull;

// digits of h: level-1 partition, level-2 sub-bucket (both arbitrary radix), table slot
__device__ __forceinline__ u32 part1(u64 h, u32 P1) { return __umulhi((u32)(h >> 32), P1); }
__device__ __forceinline__ u32 part2(u64 h, u32 P1, u32 P2) { return __umulhi((u32)(h >> 32) * P1, P2); }

static constexpr u32 BUCKET_SLOTS = 8192;          // shared-memory table of one sub-bucket
static constexpr u32 BUCKET_CAP = 6144;            // instances per sub-bucket <= 0.75 * slots: the table can never fill
static constexpr u32 MAX_P1 = 4096;
static constexpr u32 MAX_P2 = 2048;

// ---- level 1 -----------------------------------------------------------------------------------
static constexpr int S1_THREADS = 256;
static constexpr int S1_TILE = S1_THREADS * CHUNK;  // 8192 k-mer instances per tile

__device__ __forceinline__ u32 block_exclusive_scan_256(u32 v, u32 *s_warp /*[9]*/)
{
    const u32 lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    u32 incl = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { u32 t = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= (u32)o) incl += t; }
    if (lane == 31) s_warp[w] = incl;
    __syncthreads();
    if (w == 0)
    {
        u32 x = lane < 8 ? s_warp[lane] : 0, ix = x;
#pragma unroll
        for (int o = 1; o < 8; o <<= 1) { u32 t = __shfl_up_sync(0xffffffffu, ix, o); if (lane >= (u32)o) ix += t; }
        if (lane < 8) s_warp[lane] = ix - x;
        if (lane == 7) s_warp[8] = ix;
    }
    __syncthreads();
    return s_warp[w] + incl - v;
}

// out: partition p occupies [part_start[p], part_start[p+1]) — or, UNIFORM, [p * cap1, (p + 1) * cap1);
// part_cnt[p] running fill (zeroed by the host).
// flags[0] |= 1 when a partition region overflows (the host then lays the regions out from an exact histogram).
template <bool UNIFORM>
__global__ void __launch_bounds__(S1_THREADS, 2) k_scatter1(ReadsView rv, int k, int stride, u32 P1, u32 cap1,
                                                            const u64 *__restrict__ part_start, u32 *__restrict__ part_cnt,
                                                            u64 *__restrict__ out, u32 *__restrict__ flags)
{
    extern __shared__ __align__(16) unsigned char s_raw[];
    u64 *s_sorted = reinterpret_cast<u64*>(s_raw);                 // [S1_TILE]
    u32 *s_off = reinterpret_cast<u32*>(s_sorted + S1_TILE);      // [P1] tile count, then exclusive offset
    u32 *s_delta = s_off + P1;                                    // [P1] (fill of the partition before this tile) - offset
    __shared__ u32 s_warp[9];

    const u32 tid = threadIdx.x;
    const u64 ntiles = (rv.nchunks + S1_THREADS - 1) / S1_THREADS;
    constexpr int PER = MAX_P1 / S1_THREADS;                      // partitions a thread owns: tid, tid+256, ...
    for (u64 tile = blockIdx.x; tile < ntiles; tile += gridDim.x)
    {
        for (u32 i = tid; i < P1; i += S1_THREADS) s_off[i] = 0;
        __syncthreads();
        u64 hk[CHUNK]; u32 rk[CHUNK / 2]; u32 vmask = 0;
#pragma unroll
        for (int s = 0; s < CHUNK / 2; ++s) rk[s] = 0;
        ChunkInfo ci;
        if (locate_chunk(rv, tile * S1_THREADS + tid, k, ci))
            foreach_kmer_in_chunk(rv, ci, k, stride, [&](u64 x, u32, int s) {
                u64 h = mix64(x);
                u32 r = atomicAdd(&s_off[part1(h, P1)], 1u);
                hk[s] = h; rk[s >> 1] |= r << ((s & 1) * 16); vmask |= 1u << s; });
        __syncthreads();
        // offsets (any order of the partitions inside the tile is a valid packing; the copy-out recomputes p from h).
        // The global reservations of a thread's partitions go out back to back, nothing waits in between.
        u32 c[PER], base[PER]; u32 sum = 0;
#pragma unroll
        for (int j = 0; j < PER; ++j) { u32 p = tid + j * S1_THREADS; c[j] = p < P1 ? s_off[p] : 0u; sum += c[j]; }
#pragma unroll
        for (int j = 0; j < PER; ++j) base[j] = c[j] ? atomicAdd(&part_cnt[tid + j * S1_THREADS], c[j]) : 0u;
        u32 run = block_exclusive_scan_256(sum, s_warp);
#pragma unroll
        for (int j = 0; j < PER; ++j)
        {
            u32 p = tid + j * S1_THREADS;
            if (p < P1) { s_off[p] = run; s_delta[p] = base[j] - run; run += c[j]; }
        }
        const u32 total = s_warp[8];
        __syncthreads();
#pragma unroll
        for (int s = 0; s < CHUNK; ++s)
            if (vmask & (1u << s)) { u64 h = hk[s]; s_sorted[s_off[part1(h, P1)] + ((rk[s >> 1] >> ((s & 1) * 16)) & 0xFFFFu)] = h; }
        __syncthreads();
#pragma unroll 4
        for (u32 i = tid; i < total; i += S1_THREADS)
        {
            u64 h = s_sorted[i];
            u32 p = part1(h, P1);
            u32 d = s_delta[p] + i;
            if (UNIFORM)
            {
                if (d < cap1) out[(u64)p * cap1 + d] = h; else atomicOr(flags, 1u);
            }
            else
            {
                u64 ps = __ldg(part_start + p), pe = __ldg(part_start + p + 1);
                if (ps + d < pe) out[ps + d] = h; else atomicOr(flags, 1u);
            }
        }
        __syncthreads();
    }
}

// exact level-1 histogram (only after an overflow of the optimistic layout)
__global__ void __launch_bounds__(256) k_hist1(ReadsView rv, int k, int stride, u32 P1, u64 *__restrict__ ghist)
{
    extern __shared__ u32 s_hist[];
    for (u32 i = threadIdx.x; i < P1; i += blockDim.x) s_hist[i] = 0;
    __syncthreads();
    u64 step = (u64)gridDim.x * blockDim.x;
    for (u64 g = (u64)blockIdx.x * blockDim.x + threadIdx.x; g < rv.nchunks; g += step)
    {
        ChunkInfo ci;
        if (!locate_chunk(rv, g, k, ci)) continue;
        foreach_kmer_in_chunk(rv, ci, k, stride, [&](u64 x, u32, int) { atomicAdd(&s_hist[part1(mix64(x), P1)], 1u); });
    }
    __syncthreads();
    for (u32 i = threadIdx.x; i < P1; i += blockDim.x) if (s_hist[i]) atomicAdd(&ghist[i], (u64)s_hist[i]);
}

// ---- level 2 -----------------------------------------------------------------------------------
// Where the instances of the level-1 partitions are: W slabs (one per source rank; W = 1 on one GPU), slab j holds
// partition p at in[j * slab_stride + part_start[p]] with cnt[j * P + p] instances.
struct PartInput
{
    const u64 *in; u32 W; u64 slab_stride; const u64 *part_start; const u32 *cnt; u32 P;   // P = partitions per slab (local)
};
// Per local partition (host-built): total instances, sub-bucket count, prefix sums of tiles and sub-buckets.
struct PartPlan
{
    const u32 *n;             // [P]
    const u32 *p2;            // [P]
    const u32 *tile_start;    // [P+1] prefix of ceil(n / S2_TILE)
    const u32 *bucket_start;  // [P+1] prefix of p2
};

static constexpr int S2_THREADS = 256;
static constexpr int S2_PER = 16;
static constexpr int S2_TILE = S2_THREADS * S2_PER;     // 4096

// largest q in [lo, hi) with a[q] <= v   (a ascending, a[lo] <= v)
__device__ __forceinline__ u32 upper_seg(const u32 *__restrict__ a, u32 lo, u32 hi, u32 v)
{
    while (hi - lo > 1) { u32 mid = (lo + hi) >> 1; if (__ldg(a + mid) <= v) lo = mid; else hi = mid; }
    return lo;
}

// partitions [g0, g1) -> scratch: sub-bucket b (global numbering bucket_start[p] + q) of the group lives at
// scratch[(b - bucket_start[g0]) * BUCKET_CAP ...], fill count bfill[b].  Instances beyond BUCKET_CAP (repeat-rich
// buckets: the sizes are compound-Poisson, not Poisson) go to the overflow list; k_count_buckets moves the rest of
// such a bucket there too and the host counts the list with the global-table kernel.
struct Overflow { u64 *list; u64 *cursor; u64 cap; };

template <bool ONE_SLAB>
__global__ void __launch_bounds__(S2_THREADS) k_scatter2(PartInput pi, PartPlan pl, u32 P1_total, u32 g0, u32 g1, u32 p2max,
                                                         u64 *__restrict__ scratch, u32 *__restrict__ bfill, Overflow ovf)
{
    extern __shared__ __align__(16) unsigned char s_raw[];        // S2_TILE * 8 + 2 * 4 * p2max bytes
    u64 *s_sorted = reinterpret_cast<u64*>(s_raw);
    u32 *s_off = reinterpret_cast<u32*>(s_sorted + S2_TILE);
    u32 *s_delta = s_off + p2max;
    __shared__ u32 s_warp[9];
    const u32 tid = threadIdx.x;
    const u32 t0 = __ldg(pl.tile_start + g0), t1 = __ldg(pl.tile_start + g1);
    const u32 b0 = __ldg(pl.bucket_start + g0);
    constexpr int PER = MAX_P2 / S2_THREADS;
    for (u32 w = t0 + blockIdx.x; w < t1; w += gridDim.x)
    {
        const u32 p = upper_seg(pl.tile_start, g0, g1, w);
        const u32 n = __ldg(pl.n + p), P2 = __ldg(pl.p2 + p);
        const u32 first = (w - __ldg(pl.tile_start + p)) * S2_TILE;
        const u32 bs = __ldg(pl.bucket_start + p);
        for (u32 i = tid; i < P2; i += S2_THREADS) s_off[i] = 0;
        __syncthreads();
        u64 hk[S2_PER]; u32 rk[S2_PER / 2]; u32 vmask = 0;
#pragma unroll
        for (int s = 0; s < S2_PER / 2; ++s) rk[s] = 0;
        const u64 *src = pi.in + __ldg(pi.part_start + p);
        const u32 left = n > first ? n - first : 0u;                // instances of this tile and after
        if (ONE_SLAB)
        {
            const u64 *q = src + first + tid;
#pragma unroll
            for (int s = 0; s < S2_PER; ++s)
                if ((u32)(s * S2_THREADS) + tid < left) { hk[s] = __ldcs(q + s * S2_THREADS); vmask |= 1u << s; }
        }
        else
        {
#pragma unroll
            for (int s = 0; s < S2_PER; ++s)
            {
                u32 i = first + s * S2_THREADS + tid;
                if (i < n)
                {
                    u32 j = 0, cc = __ldg(pi.cnt + p);
                    while (i >= cc) { i -= cc; ++j; cc = __ldg(pi.cnt + (size_t)j * pi.P + p); }
                    hk[s] = __ldcs(src + (size_t)j * pi.slab_stride + i);
                    vmask |= 1u << s;
                }
            }
        }
#pragma unroll
        for (int s = 0; s < S2_PER; ++s)
            if (vmask & (1u << s)) { u32 r = atomicAdd(&s_off[part2(hk[s], P1_total, P2)], 1u); rk[s >> 1] |= r << ((s & 1) * 16); }
        __syncthreads();
        u32 c[PER], base[PER]; u32 sum = 0;
#pragma unroll
        for (int j = 0; j < PER; ++j) { u32 q = tid + j * S2_THREADS; c[j] = q < P2 ? s_off[q] : 0u; sum += c[j]; }
#pragma unroll
        for (int j = 0; j < PER; ++j) base[j] = c[j] ? atomicAdd(&bfill[bs + tid + j * S2_THREADS], c[j]) : 0u;
        u32 run = block_exclusive_scan_256(sum, s_warp);
#pragma unroll
        for (int j = 0; j < PER; ++j)
        {
            u32 q = tid + j * S2_THREADS;
            if (q < P2) { s_off[q] = run; s_delta[q] = base[j] - run; run += c[j]; }
        }
        const u32 total = s_warp[8];
        __syncthreads();
#pragma unroll
        for (int s = 0; s < S2_PER; ++s)
            if (vmask & (1u << s)) s_sorted[s_off[part2(hk[s], P1_total, P2)] + ((rk[s >> 1] >> ((s & 1) * 16)) & 0xFFFFu)] = hk[s];
        __syncthreads();
        u64 *dst = scratch + (u64)(bs - b0) * BUCKET_CAP;
#pragma unroll 4
        for (u32 i = tid; i < total; i += S2_THREADS)
        {
            u64 h = s_sorted[i];
            u32 q = part2(h, P1_total, P2);
            u32 d = s_delta[q] + i;
            if (d < BUCKET_CAP) dst[q * BUCKET_CAP + d] = h;
            else { u64 o = atomicAdd(ovf.cursor, 1ull); if (o < ovf.cap) ovf.list[o] = h; }
        }
        __syncthreads();
    }
}

// ---- counting ----------------------------------------------------------------------------------
static constexpr int CB_THREADS = 512;
static constexpr int CB_PER = BUCKET_CAP / CB_THREADS;            // 12 instances per thread at most
static constexpr int CB_SLOTS_PER = BUCKET_SLOTS / CB_THREADS;    // 16 table slots per thread

// counters: [0] reliable cursor, [1] sum of reliable counts, [2] distinct
__global__ void __launch_bounds__(CB_THREADS, 2) k_count_buckets(PartPlan pl, u32 g0, u32 g1, const u64 *__restrict__ scratch,
                                                                 const u32 *__restrict__ bfill, Overflow ovf,
                                                                 u32 lower, u32 upper, u64 *__restrict__ out_h, u32 *__restrict__ out_cnt,
                                                                 u64 *__restrict__ counters, u64 cap)
{
    extern __shared__ __align__(16) unsigned char s_raw[];
    u64 *s_key = reinterpret_cast<u64*>(s_raw);                   // [BUCKET_SLOTS]
    u32 *s_cnt = reinterpret_cast<u32*>(s_key + BUCKET_SLOTS);    // [BUCKET_SLOTS]
    __shared__ u32 s_warp[CB_THREADS / 32 + 1];
    __shared__ u64 s_base;
    const u32 tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
    const u32 b0 = __ldg(pl.bucket_start + g0), b1 = __ldg(pl.bucket_start + g1);
    u32 my_distinct = 0; u64 my_sum = 0;
    for (u32 b = b0 + blockIdx.x; b < b1; b += gridDim.x)
    {
        const u32 nraw = __ldg(bfill + b);
        const u64 *src = scratch + (u64)(b - b0) * BUCKET_CAP;
        if (nraw > BUCKET_CAP)                                     // uniform across the CTA: the whole bucket joins its overflow
        {
            if (tid == 0) s_base = atomicAdd(ovf.cursor, (u64)BUCKET_CAP);
            __syncthreads();
            for (u32 i = tid; i < BUCKET_CAP; i += CB_THREADS) { u64 o = s_base + i; if (o < ovf.cap) ovf.list[o] = src[i]; }
            __syncthreads();
            continue;
        }
        const u32 n = nraw;
        // the instances are requested first, the table is cleared while they fly
        u64 x[CB_PER];
#pragma unroll
        for (int j = 0; j < CB_PER; ++j) { u32 i = j * CB_THREADS + tid; x[j] = i < n ? __ldcs(src + i) : EMPTY_H; }
#pragma unroll
        for (int j = 0; j < CB_SLOTS_PER / 2; ++j)
        {
            ulonglong2 e; e.x = EMPTY_H; e.y = EMPTY_H;
            reinterpret_cast<ulonglong2*>(s_key)[j * CB_THREADS + tid] = e;
        }
#pragma unroll
        for (int j = 0; j < CB_SLOTS_PER / 4; ++j) reinterpret_cast<uint4*>(s_cnt)[j * CB_THREADS + tid] = make_uint4(0, 0, 0, 0);
        __syncthreads();
#pragma unroll
        for (int j = 0; j < CB_PER; ++j)
        {
            const u64 h = x[j];
            if (h == EMPTY_H) continue;
            u32 s = (u32)h & (BUCKET_SLOTS - 1), step = 0;
            while (true)                                           // triangular probing: every slot once, short tails
            {
                u64 cur = reinterpret_cast<volatile u64*>(s_key)[s];
                if (cur == h) break;
                if (cur == EMPTY_H)
                {
                    u64 prev = atomicCAS(&s_key[s], EMPTY_H, h);
                    if (prev == EMPTY_H || prev == h) break;
                }
                s = (s + ++step) & (BUCKET_SLOTS - 1);
            }
            atomicAdd(&s_cnt[s], 1u);
        }
        __syncthreads();
        // reliable k-mers of this bucket: count, reserve once per CTA, write
        u32 rel = 0, nrel = 0;
#pragma unroll
        for (int j = 0; j < CB_SLOTS_PER; ++j)
        {
            u32 c = s_cnt[j * CB_THREADS + tid];
            if (c) { ++my_distinct; if (c >= lower && c <= upper) { rel |= 1u << j; ++nrel; my_sum += c; } }
        }
        u32 incl = nrel;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { u32 t = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= (u32)o) incl += t; }
        if (lane == 31) s_warp[w] = incl;
        __syncthreads();
        if (w == 0)
        {
            u32 v = lane < CB_THREADS / 32 ? s_warp[lane] : 0, iv = v;
#pragma unroll
            for (int o = 1; o < CB_THREADS / 32; o <<= 1) { u32 t = __shfl_up_sync(0xffffffffu, iv, o); if (lane >= (u32)o) iv += t; }
            if (lane < CB_THREADS / 32) s_warp[lane] = iv - v;
            if (lane == CB_THREADS / 32 - 1) s_base = iv ? atomicAdd(&counters[0], (u64)iv) : 0ull;
        }
        __syncthreads();
        if (nrel)
        {
            u64 o = s_base + s_warp[w] + incl - nrel;
#pragma unroll
            for (int j = 0; j < CB_SLOTS_PER; ++j)
                if (rel & (1u << j)) { if (o < cap) { out_h[o] = s_key[j * CB_THREADS + tid]; out_cnt[o] = s_cnt[j * CB_THREADS + tid]; } ++o; }
        }
        __syncthreads();
    }
    // statistics: one atomic pair per warp
    for (int o = 16; o; o >>= 1) { my_distinct += __shfl_xor_sync(0xffffffffu, my_distinct, o); my_sum += __shfl_xor_sync(0xffffffffu, my_sum, o); }
    if (lane == 0) { if (my_distinct) atomicAdd(&counters[2], (u64)my_distinct); if (my_sum) atomicAdd(&counters[1], my_sum); }
}

// reliable list: h -> canonical k-mer
__global__ void k_unmix(u64 *__restrict__ v, u64 n)
{
    u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) v[i] = unmix64(v[i]);
}

} // namespace elba
