// K-mer counting, version 3 (k >= 20): one scatter of SUPER-K-MERS into minimizer buckets, counting in shared memory.
// (reference: get_kmer_count_map_keys / get_kmer_count_map_values, src/KmerOps.cpp:18-350)
//
// Why (profiles/r1_v4_launches_celegans40x.summary.txt): the two-level hash partition of count_smem.cuh moves every
// k-mer instance twice as an 8-byte word and is bound by instruction issue, not HBM: ~96 + ~60 + ~80 thread
// instructions per instance in k_scatter1 / k_scatter2 / k_count_buckets (30 + 26 + 30 ms for 4.0 G instances).
// Consecutive k-mers of a read overlap in k-1 bases, so the unit that is moved here is a run of consecutive k-mers
// that share their MINIMIZER (the smallest hashed canonical m-mer inside the k-mer, m = k - W + 1 <= 16):
//
//   k_skm_scatter   thread = 32 consecutive window starts of one read.  The 32 + W - 1 canonical m-mers are cut out of
//                   the 2-bit words with funnel shifts (no rolling chain), hashed with one IMAD, and the sliding
//                   minimum over W of them comes from log2(W) + 1 rounds of pairwise mins in registers.  A run of
//                   k-mers with one minimizer becomes ONE 16-byte record {up to 61 bases, n - 1} appended to the
//                   bucket of that minimizer: one global atomic and one 16-byte store per ~7 instances instead of a
//                   staged 8-byte store per instance.  A k-mer and its reverse complement hold the same canonical
//                   m-mers, so every instance of a canonical k-mer lands in the same bucket.
//   k_skm_count     one CTA per bucket (~2400 instances, at most BUCKET_CAP): instances are dealt to the threads in
//                   equal consecutive ranges (prefix sum of the records' n, one binary search per thread), each k-mer
//                   is cut out of its record, canonicalised, mixed (h = mix64) and counted in the 8192-slot
//                   shared-memory table: the CAS of a group of instances is issued back to back before any result is
//                   looked at (no load-then-CAS chain, no divergent loop on the common path).  Reliable {h, count}
//                   are appended exactly as k_count_buckets does.
//   k_skm_count_global   the exact fallback (global table of kmer_count.cuh) for records of buckets that overflowed
//                   their record capacity or BUCKET_CAP instances (skewed minimizers, repeats).
//
// Level 2 of the old scheme does not exist: the minimizer space (4^m / 2, m >= 13) is fine enough to cut buckets of a few
// thousand instances in one pass.  That is not true for k < 20 (m would be too short for large genomes or W too small
// to compress), where the hash path of count_smem.cuh stays.
#pragma once
#include "common.cuh"
#include "count_smem.cuh"
#include "kmer_count.cuh"

namespace elba {

// 16-byte record: x = bases 0..31 of the run (base 0 at bits 63..62), y = bases 32..60 left-aligned | (n - 1) in the
// low 5 bits.  A run of n k-mers holds n + k - 1 <= 61 bases, hence n <= min(32, 62 - k).
typedef ulonglong2 SkmRec;

__host__ __device__ __forceinline__ u32 skm_nmax(int k) { return (u32)(62 - k < 32 ? 62 - k : 32); }

// (k) -> minimizer length m and window W = k - m + 1; false: use the hash path
inline bool skm_geometry(int k, int &m, int &W)
{
    if (k == 32) { W = 17; m = 16; return true; }
    if (k >= 28) { W = 16; m = k - 15; return true; }
    if (k >= 24) { W = 12; m = k - 11; return true; }
    if (k >= 20) { W = 8;  m = k - 7;  return true; }
    return false;
}

// reverse complement of 16 bases in one word
__device__ __forceinline__ u32 revcomp32(u32 x)
{
    u32 z = __brev(~x);
    return ((z >> 1) & 0x55555555u) | ((z & 0x55555555u) << 1);
}

// bucket of a minimizer value.  The minimum of W hashed values is concentrated near zero: re-mix before scaling.
__device__ __forceinline__ u32 skm_bucket(u32 v, u32 NB)
{
    v ^= v >> 15; v *= 0x2C1B3C6Du; v ^= v >> 12; v *= 0x297A2D39u; v ^= v >> 15;
    return __umulhi(v, NB);
}

// Where records go.  Bucket b (global numbering, NB of them) owns slab[b * rcap ...]; fill[b] counts every record
// offered to it (records beyond rcap go to the overflow list, and k_skm_count then sends the rest of that bucket
// there too, so that all instances of a k-mer are counted in one place).
struct RecSink
{
    SkmRec *slab; u32 *fill; u32 rcap; u32 NB;
    SkmRec *ovf; u64 *ovf_cursor; u64 *ovf_inst; u64 ovf_cap;
};

static constexpr int SK_THREADS = 256;

template <int W>
__global__ void __launch_bounds__(SK_THREADS, 4) k_skm_scatter(ReadsView rv, int k, int m, u32 nmax, RecSink sink)
{
    static_assert(W >= 1 && W <= 17, "the m-mers of a chunk must fit the 64 loaded bases");
    __shared__ u32 s_mn[CHUNK * SK_THREADS];          // minimizer of window start s of this thread's chunk: [s][tid]
    constexpr int NM = CHUNK + W - 1;                 // m-mers a chunk looks at
    constexpr int J = (W >= 16) ? 4 : (W >= 8) ? 3 : (W >= 4) ? 2 : (W >= 2) ? 1 : 0;
    constexpr int D = W - (1 << J);                   // window W = two windows of 2^J, D apart
    const u32 tid = threadIdx.x;
    const u32 maskL = (m >= 16) ? 0xFFFFFFFFu : ~(0xFFFFFFFFu >> (2 * m));
    const u32 vs = 2u * (u32)(16 - m);
    const u64 step = (u64)gridDim.x * SK_THREADS;
    for (u64 g = (u64)blockIdx.x * SK_THREADS + tid; g < rv.nchunks; g += step)
    {
        ChunkInfo ci;
        locate_chunk(rv, g, k, ci);
        const u64 a = __ldg(rv.off + ci.read) + (ci.p0 >> 2);
        u64 w0, w1;
        load_bases64(rv.buf, a, w0, w1);
        // forward words T (base 16 i .. 16 i + 15 in T[i]) and the reverse complement of the 64-base window, shifted
        // left by 16 - m bases so that every m-mer's twin sits at a compile-time offset
        u32 T[5] = { (u32)(w0 >> 32), (u32)w0, (u32)(w1 >> 32), (u32)w1, 0u };
        u32 Rw[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) Rw[j] = revcomp32(T[3 - j]);
        u32 V[5];
#pragma unroll
        for (int j = 0; j < 3; ++j) V[j] = __funnelshift_l(Rw[j + 1], Rw[j], vs);
        V[3] = Rw[3] << vs; V[4] = 0u;
        u32 v[NM];
#pragma unroll
        for (int q = 0; q < NM; ++q)
        {
            const u32 f = __funnelshift_l(T[(q >> 4) + 1], T[q >> 4], 2 * (q & 15)) & maskL;
            const int e = 48 - q;
            const u32 r = __funnelshift_l(V[(e >> 4) + 1], V[e >> 4], 2 * (e & 15)) & maskL;
            v[q] = min(f, r) * 0x9E3779B1u + 0x7F4A7C15u;       // order of the canonical m-mers: one IMAD
        }
        // sliding minimum over W: windows of 2, 4, .. 2^J, then two of them
#pragma unroll
        for (int st = 1; st < (1 << J); st <<= 1)
        {
#pragma unroll
            for (int i = 0; i + st < NM; ++i) v[i] = min(v[i], v[i + st]);
        }
        u32 bmask = 1u;
        u32 prev = min(v[0], v[D]);
        s_mn[tid] = prev;
#pragma unroll
        for (int s = 1; s < CHUNK; ++s)
        {
            const u32 x = min(v[s], v[s + D]);
            s_mn[s * SK_THREADS + tid] = x;
            bmask |= (x != prev ? 1u : 0u) << s;
            prev = x;
        }
        const u32 nk = ci.nk;
        bmask &= (nk >= 32u) ? 0xFFFFFFFFu : ((1u << nk) - 1u);
        // one record per run of equal minimizers
        while (bmask)
        {
            const u32 s0 = __ffs(bmask) - 1;
            bmask &= bmask - 1;
            u32 n = (bmask ? (u32)__ffs(bmask) - 1u : nk) - s0;
            if (n > nmax) { n = nmax; bmask |= 1u << (s0 + n); }
            const u32 b = skm_bucket(s_mn[s0 * SK_THREADS + tid], sink.NB);
            const u32 sh = 2 * s0;
            SkmRec rec;
            rec.x = sh ? ((w0 << sh) | (w1 >> (64 - sh))) : w0;
            rec.y = ((w1 << sh) & ~31ull) | (u64)(n - 1);
            const u32 slot = atomicAdd(sink.fill + b, 1u);
            if (slot < sink.rcap) sink.slab[(u64)b * sink.rcap + slot] = rec;
            else
            {
                const u64 o = atomicAdd(sink.ovf_cursor, 1ull);
                atomicAdd(sink.ovf_inst, (u64)n);
                if (o < sink.ovf_cap) sink.ovf[o] = rec;
            }
        }
    }
}

// ---- counting ----------------------------------------------------------------------------------
// k-mer j of a record, left-aligned
__device__ __forceinline__ u64 skm_kmer(const SkmRec &rec, u32 j, u64 kmask)
{
    const u32 sh = 2 * j;
    const u64 f = sh ? ((rec.x << sh) | (rec.y >> (64 - sh))) : rec.x;
    return f & kmask;
}
// GetRep (src/Kmer.cpp:200-205) of a left-aligned k-mer: min(forward, twin)
__device__ __forceinline__ u64 canonical_of(u64 fwd, int lsh)
{
    const u64 rc = (((u64)revcomp32((u32)fwd) << 32) | revcomp32((u32)(fwd >> 32))) << lsh;
    return fwd < rc ? fwd : rc;
}

// Records of the buckets this GPU counts: W slabs (one per source GPU; W = 1 on one GPU).  Slab j holds bucket b at
// base[j * slab_stride + b * rcap ...] with fill[j * fill_stride + b] records offered.
struct RecSlabs { const SkmRec *base; u64 slab_stride; const u32 *fill; u32 fill_stride; u32 W; u32 rcap; };
struct RecOverflow { SkmRec *list; u64 *cursor; u64 *inst; u64 cap; };

static constexpr int SC_THREADS = 512;
static constexpr int SC_PER = BUCKET_CAP / SC_THREADS;       // 12 instances per thread at most
static constexpr int SC_GROUP = 6;                           // instances whose CAS are in flight together
static constexpr u32 SC_MAXREC = 2048;                       // records of one bucket indexed in shared memory
static constexpr int SC_RPT = SC_MAXREC / SC_THREADS;
static constexpr int SC_SLOTS_PER = BUCKET_SLOTS / SC_THREADS;
static constexpr u32 SC_MAXW = 64;
static constexpr size_t SC_SMEM = (sizeof(u64) + sizeof(u32)) * BUCKET_SLOTS + sizeof(u32) * (SC_MAXREC + 1);

template <int NWARPS>
__device__ __forceinline__ u32 block_exclusive_scan(u32 v, u32 *s_warp /*[NWARPS + 1]*/)
{
    const u32 lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    u32 incl = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { u32 t = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= (u32)o) incl += t; }
    if (lane == 31) s_warp[w] = incl;
    __syncthreads();
    if (w == 0)
    {
        u32 x = lane < NWARPS ? s_warp[lane] : 0, ix = x;
#pragma unroll
        for (int o = 1; o < NWARPS; o <<= 1) { u32 t = __shfl_up_sync(0xffffffffu, ix, o); if (lane >= (u32)o) ix += t; }
        if (lane < NWARPS) s_warp[lane] = ix - x;
        if (lane == NWARPS - 1) s_warp[NWARPS] = ix;
    }
    __syncthreads();
    return s_warp[w] + incl - v;
}

__device__ __forceinline__ const SkmRec *skm_rec_ptr(const RecSlabs &in, u32 b, u32 r, const u32 *s_cum)
{
    u32 j = 0;
    while (r >= s_cum[j + 1]) ++j;
    return in.base + (u64)j * in.slab_stride + (u64)b * in.rcap + (r - s_cum[j]);
}

// counters: [0] reliable cursor, [1] sum of reliable counts, [2] distinct
__global__ void __launch_bounds__(SC_THREADS, 2) k_skm_count(RecSlabs in, u32 nb, int k, RecOverflow ovf, u32 lower, u32 upper,
                                                             u64 *__restrict__ out_h, u32 *__restrict__ out_cnt,
                                                             u64 *__restrict__ counters, u64 cap)
{
    extern __shared__ __align__(16) unsigned char s_raw[];
    u64 *s_key = reinterpret_cast<u64*>(s_raw);                       // [BUCKET_SLOTS]
    u32 *s_cnt = reinterpret_cast<u32*>(s_key + BUCKET_SLOTS);        // [BUCKET_SLOTS]
    u32 *s_start = s_cnt + BUCKET_SLOTS;                              // [SC_MAXREC + 1] first instance of record r
    __shared__ u32 s_warp[SC_THREADS / 32 + 1];
    __shared__ u32 s_cum[SC_MAXW + 1];
    __shared__ u32 s_taint;
    __shared__ u64 s_base;
    const u32 tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
    const int lsh = 2 * (32 - k);
    const u64 kmask = (k == 32) ? ~0ull : (~0ull << lsh);
    u32 my_distinct = 0; u64 my_sum = 0;
    for (u32 b = blockIdx.x; b < nb; b += gridDim.x)
    {
        if (tid == 0)
        {
            u32 cum = 0, taint = 0;
            s_cum[0] = 0;
            for (u32 j = 0; j < in.W; ++j)
            {
                const u32 f = __ldg(in.fill + (size_t)j * in.fill_stride + b);
                taint |= f > in.rcap;
                cum += min(f, in.rcap);
                s_cum[j + 1] = cum;
            }
            s_taint = taint | (cum > SC_MAXREC);
        }
        __syncthreads();
        const u32 nrec = s_cum[in.W];
        bool spill = s_taint != 0;
        u32 total = 0;
        if (!spill)
        {
            // instances per record -> first instance of every record
            u32 nn[SC_RPT]; u32 sum = 0;
#pragma unroll
            for (int i = 0; i < SC_RPT; ++i)
            {
                const u32 r = tid * SC_RPT + i;
                nn[i] = r < nrec ? ((u32)__ldg(&skm_rec_ptr(in, b, r, s_cum)->y) & 31u) + 1u : 0u;
                sum += nn[i];
            }
            u32 run = block_exclusive_scan<SC_THREADS / 32>(sum, s_warp);
#pragma unroll
            for (int i = 0; i < SC_RPT; ++i) { const u32 r = tid * SC_RPT + i; if (r <= nrec) s_start[r] = run; run += nn[i]; }
            total = s_warp[SC_THREADS / 32];
            spill = total > BUCKET_CAP;
        }
        if (spill)                                                     // uniform across the CTA
        {
            if (tid == 0) s_base = atomicAdd(ovf.cursor, (u64)nrec);
            __syncthreads();
            u32 ninst = 0;
            for (u32 r = tid; r < nrec; r += SC_THREADS)
            {
                const SkmRec rec = __ldg(skm_rec_ptr(in, b, r, s_cum));
                ninst += ((u32)rec.y & 31u) + 1u;
                const u64 o = s_base + r;
                if (o < ovf.cap) ovf.list[o] = rec;
            }
            for (int o = 16; o; o >>= 1) ninst += __shfl_xor_sync(0xffffffffu, ninst, o);
            if (lane == 0 && ninst) atomicAdd(ovf.inst, (u64)ninst);
            __syncthreads();
            continue;
        }
        // clear the table
#pragma unroll
        for (int j = 0; j < SC_SLOTS_PER / 2; ++j)
        {
            ulonglong2 e; e.x = EMPTY_H; e.y = EMPTY_H;
            reinterpret_cast<ulonglong2*>(s_key)[j * SC_THREADS + tid] = e;
        }
#pragma unroll
        for (int j = 0; j < SC_SLOTS_PER / 4; ++j) reinterpret_cast<uint4*>(s_cnt)[j * SC_THREADS + tid] = make_uint4(0, 0, 0, 0);
        __syncthreads();
        // this thread's instances: [i0, i0 + nv), consecutive, starting inside record r at k-mer j
        const u32 c = (total + SC_THREADS - 1) / SC_THREADS;
        const u32 i0 = tid * c;
        const u32 nv = i0 < total ? min(c, total - i0) : 0u;
        if (nv)
        {
            u32 lo = 0, hi = nrec;                                     // s_start[lo] <= i0 < s_start[hi]
            while (hi - lo > 1) { const u32 mid = (lo + hi) >> 1; if (s_start[mid] <= i0) lo = mid; else hi = mid; }
            u32 r = lo, j = i0 - s_start[lo];
            SkmRec rec = __ldg(skm_rec_ptr(in, b, r, s_cum));
            SkmRec nxt = rec;
            if (r + 1 < nrec) nxt = __ldg(skm_rec_ptr(in, b, r + 1, s_cum));
            u32 n = ((u32)rec.y & 31u) + 1u;
#pragma unroll
            for (int g = 0; g < SC_PER; g += SC_GROUP)
            {
                u64 H[SC_GROUP]; u32 S[SC_GROUP]; u64 P[SC_GROUP];
#pragma unroll
                for (int q = 0; q < SC_GROUP; ++q)
                {
                    H[q] = EMPTY_H;
                    if ((u32)(g + q) < nv)
                    {
                        if (j == n)
                        {
                            ++r; j = 0; rec = nxt; n = ((u32)rec.y & 31u) + 1u;
                            if (r + 1 < nrec) nxt = __ldg(skm_rec_ptr(in, b, r + 1, s_cum));
                        }
                        H[q] = mix64(canonical_of(skm_kmer(rec, j, kmask), lsh));
                        ++j;
                    }
                }
#pragma unroll
                for (int q = 0; q < SC_GROUP; ++q)
                    if (H[q] != EMPTY_H) { S[q] = (u32)H[q] & (BUCKET_SLOTS - 1); P[q] = atomicCAS(&s_key[S[q]], EMPTY_H, H[q]); }
#pragma unroll
                for (int q = 0; q < SC_GROUP; ++q)
                    if (H[q] != EMPTY_H)
                    {
                        u32 s = S[q], stepp = 0; u64 pv = P[q];
                        while (pv != EMPTY_H && pv != H[q])            // triangular probing: every slot once; total <= 0.75 * slots
                        {
                            s = (s + ++stepp) & (BUCKET_SLOTS - 1);
                            pv = atomicCAS(&s_key[s], EMPTY_H, H[q]);
                        }
                        atomicAdd(&s_cnt[s], 1u);
                    }
            }
        }
        __syncthreads();
        // reliable k-mers of this bucket: count, reserve once per CTA, write
        u32 rel = 0, nrel = 0;
#pragma unroll
        for (int j = 0; j < SC_SLOTS_PER; ++j)
        {
            const u32 cc = s_cnt[j * SC_THREADS + tid];
            if (cc) { ++my_distinct; if (cc >= lower && cc <= upper) { rel |= 1u << j; ++nrel; my_sum += cc; } }
        }
        const u32 excl = block_exclusive_scan<SC_THREADS / 32>(nrel, s_warp);
        if (tid == 0) { const u32 tot = s_warp[SC_THREADS / 32]; s_base = tot ? atomicAdd(&counters[0], (u64)tot) : 0ull; }
        __syncthreads();
        if (nrel)
        {
            u64 o = s_base + excl;
#pragma unroll
            for (int j = 0; j < SC_SLOTS_PER; ++j)
                if (rel & (1u << j)) { if (o < cap) { out_h[o] = s_key[j * SC_THREADS + tid]; out_cnt[o] = s_cnt[j * SC_THREADS + tid]; } ++o; }
        }
        __syncthreads();
    }
    for (int o = 16; o; o >>= 1) { my_distinct += __shfl_xor_sync(0xffffffffu, my_distinct, o); my_sum += __shfl_xor_sync(0xffffffffu, my_sum, o); }
    if (lane == 0) { if (my_distinct) atomicAdd(&counters[2], (u64)my_distinct); if (my_sum) atomicAdd(&counters[1], my_sum); }
    (void)w;
}

// exact fallback: the k-mers of a list of records into the global table of kmer_count.cuh
__global__ void __launch_bounds__(256) k_skm_count_global(const SkmRec *__restrict__ list, u64 nrec, int k, TableRef T,
                                                          u32 *__restrict__ err, u64 *__restrict__ distinct)
{
    const int lsh = 2 * (32 - k);
    const u64 kmask = (k == 32) ? ~0ull : (~0ull << lsh);
    const u64 step = (u64)gridDim.x * blockDim.x;
    const u64 rounds = (nrec + step - 1) / step;
    u32 nd = 0;
    for (u64 it = 0; it < rounds; ++it)
    {
        const u64 i = it * step + (u64)blockIdx.x * blockDim.x + threadIdx.x;
        if (i < nrec)
        {
            const SkmRec rec = list[i];
            const u32 n = ((u32)rec.y & 31u) + 1u;
            for (u32 j = 0; j < n; ++j) nd += table_insert(T, mix64(canonical_of(skm_kmer(rec, j, kmask), lsh)), 1u, err);
        }
    }
    tally(distinct, nd);
}

} // namespace elba
