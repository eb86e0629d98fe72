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
//                   k-mers with one minimizer becomes ONE 32-byte record {up to 61 bases, n - 1, read, pos, offset of
//                   the run inside its bucket} appended to the bucket of that minimizer: one 64-bit global atomic and
//                   one 256-bit store of a whole sector per ~7 instances instead of a staged 8-byte store per
//                   instance.  A k-mer and its reverse complement hold the same canonical m-mers, so every instance
//                   of a canonical k-mer lands in the same bucket.  The record carries what BOTH exchanges of the
//                   reference carry (k-mer; k-mer + read + pos, KmerOps.cpp:105,224).
//   k_skm_count4    (skm_count.cuh) one CTA per bucket (~2300 instances, at most 6144): the bucket's slab comes into shared
//                   memory by one bulk copy; every thread takes an equal range of consecutive instances, cuts each k-mer out
//                   of its record, canonicalises it and counts it in a 2048-slot shared-memory table of the bucket's DISTINCT
//                   k-mers.  Reliable {k-mer, count} go to a list; every instance of a reliable k-mer becomes a seed
//                   {list index, pos, read}: pass 2 of the reference (KmerOps.cpp:283-318) without a second sweep.
//   k_skm4_*_global the exact fallback (global table of kmer_count.cuh) for buckets that overflowed their record capacity,
//                   6144 instances or 3/4 of the table (skewed minimizers, repeats, noisy reads).
//
// Level 2 of the old scheme does not exist: the minimizer space (4^m / 2, m >= 13) is fine enough to cut buckets of a few
// thousand instances in one pass.  That is not true for k < 20 (m would be too short for large genomes or W too small
// to compress), where the hash path of count_smem.cuh stays.
#pragma once
#include "common.cuh"
#include "count_smem.cuh"
#include "kmer_count.cuh"

namespace elba {

// 32-byte record = one DRAM sector, written by ONE 256-bit store (STG.E.ENL2.256): x = bases 0..31 of the run (base 0 at
// bits 63..62), y = bases 32..60 left-aligned | (n - 1) in the low 5 bits, meta = (local read << 32) | position of the
// run's first window start (what pass 2 of the reference carries per instance, src/KmerOps.cpp:224-262).  A run of n
// k-mers holds n + k - 1 <= 61 bases, hence n <= min(32, 62 - k).
// Measured (profiles/r1_v6_scatter_records.md): with 16-byte records plus a separate 8-byte meta array the scatter is
// bound by partially written sectors (L2 fills them from DRAM and writes them back more than once: 3.5 GB read +
// 6.7 GB written for 3.6 GB of payload); a whole aligned sector per record needs no fill and is written once.
// spare = (first instance of the run inside its bucket << 8) | n: both come out of the ONE 64-bit atomic that reserves the
// record's slot (fill word: records offered in the low half, instances offered in the high half), so slot order and
// instance order agree and the count kernel needs no prefix sum over the records.
struct __align__(32) SkmRec { u64 x, y, meta, spare; };
struct SkmBases { u64 x, y; };

__device__ __forceinline__ void skm_store(SkmRec *p, u64 x, u64 y, u64 meta, u64 spare)
{
    asm volatile("st.global.v4.u64 [%0], {%1, %2, %3, %4};" :: "l"(p), "l"(x), "l"(y), "l"(meta), "l"(spare) : "memory");
}
__device__ __forceinline__ SkmBases skm_load_bases(const SkmRec *p)
{
    const ulonglong2 v = __ldg(reinterpret_cast<const ulonglong2*>(p));
    SkmBases b; b.x = v.x; b.y = v.y; return b;
}
__device__ __forceinline__ SkmRec skm_load(const SkmRec *p)
{
    SkmRec r;
    asm volatile("ld.global.nc.v4.u64 {%0, %1, %2, %3}, [%4];" : "=l"(r.x), "=l"(r.y), "=l"(r.meta), "=l"(r.spare) : "l"(p));
    return r;
}

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
// Two digits of the mixed value: the OWNER (GPU) of the bucket, then the bucket among the nb_own buckets of that owner;
// global bucket = owner * nb_own + local.  On one GPU the owner digit is always 0.
__device__ __forceinline__ u32 skm_mix(u32 v)
{
    v ^= v >> 15; v *= 0x2C1B3C6Du; v ^= v >> 12; v *= 0x297A2D39u; v ^= v >> 15;
    return v;
}

static constexpr int SK_MAXW = 16;                    // GPUs the peer-memory record exchange addresses

// Where records go.  Every GPU parses ITS OWN reads.  An owner's slab has one region per source GPU (its own region holds
// one sub-slab of rcap records per bucket, the others are packed by k_skm_forward), so that the slot reservation stays a LOCAL atomic: fill[global bucket] counts what THIS GPU
// offered (records in the low half, instances in the high half).  Records of this GPU's own buckets go straight into its
// slab; records of other GPUs' buckets are staged locally in the same shape (stage[global bucket][rcap]) and then pushed
// into the owners' slabs through peer memory by k_skm_forward: bucket after bucket, only the filled part, in address order.
// (Storing every 32-byte record straight into peer memory was measured first: 2.5 s instead of 9 ms on two B200 --
// sector-random stores into a 25 GB window of imported memory do not go through NVLink at any useful rate.)
// Together the two kernels are the personalised all-to-all of the reference (src/KmerOps.cpp:151 and :274, both exchanges
// in one record), 4.6 B per instance on the wire, with no packing pass and no host-visible counts.
// Records beyond rcap go to the owner's overflow list (a remote atomic, rare), and the owner then counts that whole bucket
// with the global-table fallback.
struct RecSink
{
    SkmRec *slab; SkmRec *stage; u64 *fill; u32 rcap, nb_own, nsrc, me, read_base;
    SkmRec *ovf[SK_MAXW]; u64 *ovf_ctr[SK_MAXW];      // owner's overflow list and its counters: [0] records, [1] instances
    u64 ovf_cap;
};

static constexpr int SK_THREADS = 256;
static constexpr int SK_NR = 2;                       // records whose bucket reservations are in flight together

// contig: CTA c walks the chunks [c * iters * SK_THREADS, (c + 1) * iters * SK_THREADS) in order, so a thread's next chunk
// is SK_THREADS chunks further in the same or the next read: the read is found once by binary search and then followed.
// Otherwise the grid strides over the chunks together and every chunk is located by its own binary search.
template <int W, int NR>
__global__ void __launch_bounds__(SK_THREADS, 4) k_skm_scatter(ReadsView rv, int k, int m, u32 nmax, RecSink sink, u64 iters, int contig, u64 g_begin, u64 g_end)
{
    static_assert(W >= 1 && W <= 17, "the m-mers of a chunk must fit the 64 loaded bases");
    __shared__ u32 s_mn[CHUNK * SK_THREADS];          // minimizer of window start s of this thread's chunk: [s][tid]
    constexpr int NM = CHUNK + W - 1;                 // m-mers a chunk looks at
    constexpr int J = (W >= 16) ? 4 : (W >= 8) ? 3 : (W >= 4) ? 2 : (W >= 2) ? 1 : 0;
    constexpr int D = W - (1 << J);                   // window W = two windows of 2^J, D apart
    const u32 tid = threadIdx.x;
    const u32 maskL = (m >= 16) ? 0xFFFFFFFFu : ~(0xFFFFFFFFu >> (2 * m));
    const u32 vs = 2u * (u32)(16 - m);
    // this launch handles the chunks [g_begin, g_end): all of them, or the reads of one slice of an upload still in flight
    u64 g = g_begin + (contig ? (u64)blockIdx.x * iters * SK_THREADS + tid : (u64)blockIdx.x * SK_THREADS + tid);
    const u64 gstep = contig ? (u64)SK_THREADS : (u64)gridDim.x * SK_THREADS;
    if (g >= g_end) return;
    u32 read = find_read(rv.chunk_start, rv.n, g);
    u64 cs0 = __ldg(rv.chunk_start + read), cs1 = __ldg(rv.chunk_start + read + 1);
    u64 roff = __ldg(rv.off + read);
    u32 rnk = __ldg(rv.len + read) - (u32)k + 1;
    for (u64 it = 0; it < iters && g < g_end; ++it, g += gstep)
    {
        if (g >= cs1)
        {
            if (contig) { do { ++read; cs1 = __ldg(rv.chunk_start + read + 1); } while (g >= cs1); }     // reads without chunks are stepped over
            else { read = find_read(rv.chunk_start, rv.n, g); cs1 = __ldg(rv.chunk_start + read + 1); }
            cs0 = __ldg(rv.chunk_start + read);
            roff = __ldg(rv.off + read);
            rnk = __ldg(rv.len + read) - (u32)k + 1;
        }
        const u32 p0 = (u32)(g - cs0) * CHUNK;
        const u32 nk = min((u32)CHUNK, rnk - p0);
        const u64 a = roff + (p0 >> 2);
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
        bmask &= (nk >= 32u) ? 0xFFFFFFFFu : ((1u << nk) - 1u);
        const u64 meta0 = ((u64)(read + sink.read_base) << 32) | p0;
        // one record per run of equal minimizers; the bucket reservations of SK_NR records are issued back to back
        while (bmask)
        {
            u32 b[NR], own[NR], n[NR], s0[NR]; u64 fw[NR]; bool ok[NR]; SkmBases rec[NR];
#pragma unroll
            for (int i = 0; i < NR; ++i)
            {
                ok[i] = bmask != 0;
                if (ok[i])
                {
                    s0[i] = __ffs(bmask) - 1;
                    bmask &= bmask - 1;
                    n[i] = (bmask ? (u32)__ffs(bmask) - 1u : nk) - s0[i];
                    if (n[i] > nmax) { n[i] = nmax; bmask |= 1u << (s0[i] + n[i]); }
                    {
                        const u64 t = (u64)skm_mix(s_mn[s0[i] * SK_THREADS + tid]) * sink.nsrc;
                        own[i] = (u32)(t >> 32);
                        b[i] = __umulhi((u32)t, sink.nb_own);           // bucket among the owner's
                    }
                    const u32 sh = 2 * s0[i];
                    rec[i].x = sh ? ((w0 << sh) | (w1 >> (64 - sh))) : w0;
                    rec[i].y = ((w1 << sh) & ~31ull) | (u64)(n[i] - 1);
                }
            }
#pragma unroll
            for (int i = 0; i < NR; ++i) if (ok[i]) fw[i] = atomicAdd(sink.fill + ((u64)own[i] * sink.nb_own + b[i]), ((u64)n[i] << 32) | 1ull);
#pragma unroll
            for (int i = 0; i < NR; ++i)
                if (ok[i])
                {
                    const u32 slot = (u32)fw[i];
                    const u64 spare = ((fw[i] >> 32) << 8) | n[i];
                    if (slot < sink.rcap)
                    {
                        SkmRec *dst = own[i] == sink.me ? sink.slab + (((u64)sink.me * sink.nb_own + b[i]) * sink.rcap + slot)
                                                        : sink.stage + (((u64)own[i] * sink.nb_own + b[i]) * sink.rcap + slot);
                        skm_store(dst, rec[i].x, rec[i].y, meta0 + s0[i], spare);
                    }
                    else
                    {
                        u64 *ctr = sink.ovf_ctr[own[i]];
                        const u64 o = atomicAdd(ctr, 1ull);
                        atomicAdd(ctr + 1, (u64)n[i]);
                        if (o < sink.ovf_cap) skm_store(sink.ovf[own[i]] + o, rec[i].x, rec[i].y, meta0 + s0[i], spare);
                    }
                }
        }
    }
}

// Several GPUs: the staged records of the other GPUs' buckets go into the owners' slabs (slab[r] = rank r's slab mapped into
// this process, CUDA IPC).  One warp per (owner, bucket): min(records offered, rcap) records = one contiguous run of 32-byte
// records on both sides, moved as 16-byte pieces; the buckets of one owner are visited in address order.
struct RecForward { SkmRec *slab[SK_MAXW]; const SkmRec *stage; const u64 *fill; const u64 *off; u32 rcap, nb_own, nsrc, me; };

// records this GPU offered to global bucket g that the owner will read: min(offered, rcap) (the rest went to the overflow list)
__global__ void k_skm_forward_counts(const u64 *__restrict__ fill, u64 nbg, u32 rcap, u64 *__restrict__ cnt)
{
    const u64 g = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    if (g <= nbg) cnt[g] = g < nbg ? (u64)min((u32)fill[g], rcap) : 0ull;
}

__global__ void __launch_bounds__(256) k_skm_forward(RecForward f)
{
    // The owner's slab has one REGION per source GPU (region s = slab + s * nb_own * rcap).  The source packs the records of
    // the owner's buckets into its region one bucket after the other (off = exclusive scan of the counts above): what crosses
    // NVLink towards one owner is ONE contiguous stream, written by neighbouring warps at neighbouring addresses.  Neighbouring
    // groups of warps serve different owners, starting behind this GPU in rank order, so that every GPU receives from all
    // the others at once.  (Measured on the way, profiles/r2_v8_*, r2_v17_*: 32-byte stores straight from the scatter:
    // 3.8 GB/s; strided 1.3 KB runs per bucket: 73 GB/s per GPU on eight B200.)
    const u32 lane = threadIdx.x & 31;
    const u64 nwarps = ((u64)gridDim.x * blockDim.x) >> 5;
    const u32 peers = f.nsrc - 1;
    constexpr u32 RUN = 64;                                       // consecutive buckets of one owner a group of warps takes
    const u64 nrun = ((u64)f.nb_own + RUN - 1) / RUN;
    const u64 total = nrun * RUN * peers;                         // (run, owner step, bucket in run)
    for (u64 g = ((u64)blockIdx.x * blockDim.x + threadIdx.x) >> 5; g < total; g += nwarps)
    {
        const u64 run = g / ((u64)RUN * peers); const u32 rem = (u32)(g - run * RUN * peers);
        const u32 step = rem / RUN; const u64 bl = run * RUN + (rem - step * RUN);
        if (bl >= f.nb_own) continue;
        u32 own = f.me + 1 + step; if (own >= f.nsrc) own -= f.nsrc;
        const u64 gb = (u64)own * f.nb_own + bl;                  // global bucket
        const u32 n = min((u32)__ldg(f.fill + gb), f.rcap);
        const uint4 *__restrict__ src = reinterpret_cast<const uint4*>(f.stage + gb * f.rcap);
        const u64 o = __ldg(f.off + gb) - __ldg(f.off + (u64)own * f.nb_own);       // records before this bucket in the region
        uint4 *__restrict__ dst = reinterpret_cast<uint4*>(f.slab[own] + ((u64)f.me * f.nb_own * f.rcap + o));
        for (u32 i = lane; i < 2 * n; i += 32) dst[i] = __ldcs(src + i);
    }
    __threadfence_system();           // the records are in the owners' memory before this kernel counts as done
}

// ---- counting ----------------------------------------------------------------------------------
// k-mer j of a record, left-aligned
__device__ __forceinline__ u64 skm_kmer(const SkmBases &rec, u32 j, u64 kmask)
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

static constexpr u32 SC_NOPAD = 0xFFFFFFFFu;

} // namespace elba
