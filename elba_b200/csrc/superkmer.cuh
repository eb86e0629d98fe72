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
//   k_skm_count     one CTA per bucket (~2400 instances, at most 6144): every thread takes an equal range of consecutive
//                   instances (offsets come from the scatter's fill word: no prefix sum), cuts each k-mer out of its
//                   record, canonicalises it, mixes it (h = mix64) and counts it in the 8192-slot shared-memory table.
//                   Reliable {h, count} are appended by the instance that claimed the slot; every instance of a
//                   reliable k-mer appends its {k-mer, pos, read} to the seed list: pass 2 of the reference
//                   (KmerOps.cpp:283-318) without a second sweep over the reads.
//   k_skm_count_global / k_skm_emit_global   the exact fallback (global table of kmer_count.cuh) for records of buckets
//                   that overflowed their record capacity or 6144 instances (skewed minimizers, repeats).
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
__device__ __forceinline__ u32 skm_bucket(u32 v, u32 NB)
{
    v ^= v >> 15; v *= 0x2C1B3C6Du; v ^= v >> 12; v *= 0x297A2D39u; v ^= v >> 15;
    return __umulhi(v, NB);
}

// Where records go.  Bucket b (NB of them) owns slab[b * rcap ...]; fill[b] counts every record offered to it (records
// beyond rcap go to the overflow list, and k_skm_count then sends the rest of that bucket there too, so that all
// instances of a k-mer are counted in one place).
// Several GPUs: every GPU parses ALL reads (the 2-bit arena is all-gathered: 0.25 B per base, against 4.6 B per instance
// for the records) and keeps only the records of the buckets it owns, [b_lo, b_lo + b_cnt) of NB; no record crosses
// NVLink.  read_base makes the read ids of the records global.
struct RecSink
{
    SkmRec *slab; u64 *fill; u32 rcap; u32 NB; u32 b_lo, b_cnt, read_base;
    SkmRec *ovf; u64 *ovf_cursor; u64 *ovf_inst; u64 ovf_cap;
};

static constexpr int SK_THREADS = 256;
static constexpr int SK_NR = 2;                       // records whose bucket reservations are in flight together

// contig: CTA c walks the chunks [c * iters * SK_THREADS, (c + 1) * iters * SK_THREADS) in order, so a thread's next chunk
// is SK_THREADS chunks further in the same or the next read: the read is found once by binary search and then followed.
// Otherwise the grid strides over the chunks together and every chunk is located by its own binary search.
template <int W, int NR>
__global__ void __launch_bounds__(SK_THREADS, 4) k_skm_scatter(ReadsView rv, int k, int m, u32 nmax, RecSink sink, u64 iters, int contig)
{
    static_assert(W >= 1 && W <= 17, "the m-mers of a chunk must fit the 64 loaded bases");
    __shared__ u32 s_mn[CHUNK * SK_THREADS];          // minimizer of window start s of this thread's chunk: [s][tid]
    constexpr int NM = CHUNK + W - 1;                 // m-mers a chunk looks at
    constexpr int J = (W >= 16) ? 4 : (W >= 8) ? 3 : (W >= 4) ? 2 : (W >= 2) ? 1 : 0;
    constexpr int D = W - (1 << J);                   // window W = two windows of 2^J, D apart
    const u32 tid = threadIdx.x;
    const u32 maskL = (m >= 16) ? 0xFFFFFFFFu : ~(0xFFFFFFFFu >> (2 * m));
    const u32 vs = 2u * (u32)(16 - m);
    u64 g = contig ? (u64)blockIdx.x * iters * SK_THREADS + tid : (u64)blockIdx.x * SK_THREADS + tid;
    const u64 gstep = contig ? (u64)SK_THREADS : (u64)gridDim.x * SK_THREADS;
    if (g >= rv.nchunks) return;
    u32 read = find_read(rv.chunk_start, rv.n, g);
    u64 cs0 = __ldg(rv.chunk_start + read), cs1 = __ldg(rv.chunk_start + read + 1);
    u64 roff = __ldg(rv.off + read);
    u32 rnk = __ldg(rv.len + read) - (u32)k + 1;
    for (u64 it = 0; it < iters && g < rv.nchunks; ++it, g += gstep)
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
            u32 b[NR], n[NR], s0[NR]; u64 fw[NR]; bool ok[NR]; SkmBases rec[NR];
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
                    b[i] = skm_bucket(s_mn[s0[i] * SK_THREADS + tid], sink.NB) - sink.b_lo;
                    ok[i] = b[i] < sink.b_cnt;                         // another GPU's bucket: nothing to write
                    const u32 sh = 2 * s0[i];
                    rec[i].x = sh ? ((w0 << sh) | (w1 >> (64 - sh))) : w0;
                    rec[i].y = ((w1 << sh) & ~31ull) | (u64)(n[i] - 1);
                }
            }
#pragma unroll
            for (int i = 0; i < NR; ++i) if (ok[i]) fw[i] = atomicAdd(sink.fill + b[i], ((u64)n[i] << 32) | 1ull);
#pragma unroll
            for (int i = 0; i < NR; ++i)
                if (ok[i])
                {
                    const u32 slot = (u32)fw[i];
                    const u64 spare = ((fw[i] >> 32) << 8) | n[i];
                    if (slot < sink.rcap) skm_store(sink.slab + ((u64)b[i] * sink.rcap + slot), rec[i].x, rec[i].y, meta0 + s0[i], spare);
                    else
                    {
                        const u64 o = atomicAdd(sink.ovf_cursor, 1ull);
                        atomicAdd(sink.ovf_inst, (u64)n[i]);
                        if (o < sink.ovf_cap) skm_store(sink.ovf + o, rec[i].x, rec[i].y, meta0 + s0[i], spare);
                    }
                }
        }
    }
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

// Records of the buckets this GPU counts: bucket b holds min(records offered, rcap) records at slab[b * rcap ...];
// fill[b] = (instances offered << 32) | records offered.
struct RecSlabs { const SkmRec *slab; const u64 *fill; u32 rcap; };
struct RecOverflow { SkmRec *list; u64 *cursor; u64 *inst; u64 cap; };
// Where the instances of reliable k-mers go (pass 2 of the reference, KmerOps.cpp:283-318, fused into counting):
// {canonical k-mer, pos, local read} per instance.  Like the reliable list it is handed out in chunks that a CTA fills
// on its own (one global atomic per SEED_CHUNK entries, off the critical path); unused entries are holes
// (k-mer = EMPTY_KEY / h = EMPTY_H).  *cursor ends as the number of entries handed out even if cap was too small.
struct SeedSink { Candidate *out; u64 *cursor; u64 cap; };

static constexpr u32 SC_MAXREC = 2048;                       // records of one bucket a CTA indexes
static constexpr u32 SC_OWNER = 0x8000u;                     // bit of an instance's slot code: this thread claimed the slot
static constexpr u32 SEED_CHUNK = 16384, REL_CHUNK = 16384;  // >= bucket capacity: one bucket always fits the rest of a fresh chunk
static constexpr u32 SC_NOPAD = 0xFFFFFFFFu;
// geometry of the bucket kernel: SLOTS-slot table, at most SLOTS * 3 / 4 instances per bucket (the table can never fill)
__host__ __device__ constexpr u32 skm_bucket_cap(u32 slots) { return slots / 4 * 3; }
__host__ __device__ constexpr size_t skm_count_smem(u32 slots, u32 threads) { return (sizeof(u64) + sizeof(u32)) * slots + sizeof(u32) * threads; }

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

// counters: [0] reliable-list entries handed out (with holes), [1] sum of reliable counts, [2] distinct, [8] reliable k-mers.
// One CTA per bucket, THREE barriers per bucket (ncu of the previous version: 36 % of the warp time was spent at its
// seven barriers, behind two block scans and a global reservation per bucket).
//   * The bucket's instance total and every record's first instance come from the scatter's fill word, so the c
//     consecutive instances of thread t start in the record that covers instance t * c: the record's owner writes
//     that into s_first[t].  No prefix sum.
//   * Every thread counts its instances in the shared-memory table and remembers, per instance, the slot it ended in
//     and whether its CAS claimed that slot.  The claiming instance is the k-mer's one representative: it appends
//     {h, count} to the reliable list, so the table is never scanned; with EMIT every instance whose slot holds a
//     reliable count also appends its {k-mer, pos, read} to the seed list.
//   * Both lists are written into CTA-private chunks (offsets by shared-memory atomics); the fill word of the bucket
//     after the next and the records of the next bucket are requested while this bucket is counted.
template <int THREADS, int SLOTS, int MINB, bool EMIT>
__global__ void __launch_bounds__(THREADS, MINB) k_skm_count(RecSlabs in, u32 nb, int k, RecOverflow ovf, u32 lower, u32 upper,
                                                          u64 *__restrict__ out_h, u32 *__restrict__ out_cnt,
                                                          u64 *__restrict__ counters, u64 cap, SeedSink seeds)
{
    constexpr u32 CAP = skm_bucket_cap(SLOTS);
    constexpr int PER = CAP / THREADS;                        // instances per thread at most
    constexpr int RPT = SC_MAXREC / THREADS;
    constexpr int KEY_V = SLOTS / 2 / THREADS, CNT_V = SLOTS / 4 / THREADS;
    static_assert(CAP % THREADS == 0 && PER % 2 == 0 && SC_MAXREC % THREADS == 0 && CNT_V >= 1 && (SLOTS & (SLOTS - 1)) == 0, "geometry");
    static_assert(SLOTS <= SC_OWNER && CAP < 65536 && CAP <= SEED_CHUNK && CAP <= REL_CHUNK, "slot codes and per-bucket tallies are 16 bits; a bucket fits a chunk");
    extern __shared__ __align__(16) unsigned char s_raw[];
    u64 *s_key = reinterpret_cast<u64*>(s_raw);                       // [SLOTS]
    u32 *s_cnt = reinterpret_cast<u32*>(s_key + SLOTS);        // [SLOTS]
    u32 *s_first = s_cnt + SLOTS;                              // [THREADS] (record << 5 | k-mer in it) of thread t's first instance
    __shared__ u64 s_spill_base, s_rel_base, s_seed_base, s_pad_rel_base, s_pad_seed_base;
    __shared__ u32 s_rel_used, s_seed_used, s_pad_rel_from, s_pad_seed_from;
    const u32 tid = threadIdx.x, lane = tid & 31;
    const u32 G = gridDim.x;
    const int lsh = 2 * (32 - k);
    const u64 kmask = (k == 32) ? ~0ull : (~0ull << lsh);
    u32 my_distinct = 0, my_rel = 0; u64 my_sum = 0;
    if (tid == 0) { s_rel_used = REL_CHUNK; s_seed_used = SEED_CHUNK; s_rel_base = 0; s_seed_base = 0; }     // no chunk yet
    // software pipeline over this CTA's buckets: fill word two buckets ahead, round-0 record words one bucket ahead
    u32 b = blockIdx.x;
    u64 fw_cur = b < nb ? __ldg(in.fill + b) : 0ull;
    u64 fw_nxt = (u64)b + G < nb ? __ldg(in.fill + b + G) : 0ull;
    u64 sp_cur = (b < nb && tid < min((u32)fw_cur, in.rcap)) ? __ldg(&in.slab[(u64)b * in.rcap + tid].spare) : 0ull;
    u64 fw_nn = 0, sp_nxt = 0;
    __syncthreads();
    for (; b < nb; b += G, fw_cur = fw_nxt, fw_nxt = fw_nn, sp_cur = sp_nxt)
    {
        fw_nn = (u64)b + 2ull * G < nb ? __ldg(in.fill + b + 2u * G) : 0ull;
        sp_nxt = ((u64)b + G < nb && tid < min((u32)fw_nxt, in.rcap)) ? __ldg(&in.slab[(u64)(b + G) * in.rcap + tid].spare) : 0ull;
        const u32 f = (u32)fw_cur, total = (u32)(fw_cur >> 32);
        const u32 nrec = min(f, in.rcap);
        const SkmRec *__restrict__ recs = in.slab + (u64)b * in.rcap;
        const bool spill = f > in.rcap || nrec > SC_MAXREC || total > CAP;      // uniform across the CTA
        if (spill)
        {
            if (tid == 0) s_spill_base = atomicAdd(ovf.cursor, (u64)nrec);
            __syncthreads();
            u32 ninst = 0;
            for (u32 r = tid; r < nrec; r += THREADS)
            {
                const SkmRec rec = skm_load(recs + r);
                ninst += ((u32)rec.y & 31u) + 1u;
                const u64 o = s_spill_base + r;
                if (o < ovf.cap) skm_store(ovf.list + o, rec.x, rec.y, rec.meta, rec.spare);
            }
            for (int o = 16; o; o >>= 1) ninst += __shfl_xor_sync(0xffffffffu, ninst, o);
            if (lane == 0 && ninst) atomicAdd(ovf.inst, (u64)ninst);
            __syncthreads();
            continue;
        }
        // room for this bucket in the CTA's output chunks (worst case: every instance reliable and distinct)
        if (tid == 0)
        {
            u32 pr = SC_NOPAD, ps = SC_NOPAD;
            if (s_rel_used + total > REL_CHUNK)
            {
                s_pad_rel_base = s_rel_base; pr = s_rel_used;
                s_rel_base = atomicAdd(&counters[0], (u64)REL_CHUNK); s_rel_used = 0;
            }
            if (EMIT && s_seed_used + total > SEED_CHUNK)
            {
                s_pad_seed_base = s_seed_base; ps = s_seed_used;
                s_seed_base = atomicAdd(seeds.cursor, (u64)SEED_CHUNK); s_seed_used = 0;
            }
            s_pad_rel_from = pr; s_pad_seed_from = ps;
        }
        // clear the table (the previous bucket ended with a barrier)
#pragma unroll
        for (int j = 0; j < KEY_V; ++j)
        {
            ulonglong2 e; e.x = EMPTY_H; e.y = EMPTY_H;
            reinterpret_cast<ulonglong2*>(s_key)[j * THREADS + tid] = e;
        }
#pragma unroll
        for (int j = 0; j < CNT_V; ++j) reinterpret_cast<uint4*>(s_cnt)[j * THREADS + tid] = make_uint4(0, 0, 0, 0);
        // every thread takes c consecutive instances; the record that holds instance t * c tells thread t where to start
        const u32 c = (total + THREADS - 1) / THREADS;                 // uniform across the CTA
        {
            const u32 inv = c > 1 ? 0xFFFFFFFFu / c + 1u : 0u;         // ceil(2^32 / c): x / c == umulhi(x, inv) for x * c < 2^32
#pragma unroll
            for (int i = 0; i < RPT; ++i)
            {
                if ((u32)(i * THREADS) >= nrec) break;                 // uniform
                const u32 r = i * THREADS + tid;
                if (r < nrec)
                {
                    const u64 sp = i == 0 ? sp_cur : __ldg(&recs[r].spare);
                    const u32 a = (u32)(sp >> 8), e = a + ((u32)sp & 0xFFu) - 1u;
                    const u32 t_lo = c > 1 ? __umulhi(a + c - 1, inv) : a, t_hi = c > 1 ? __umulhi(e, inv) : e;
                    for (u32 t = t_lo; t <= t_hi; ++t) s_first[t] = (r << 5) | (t * c - a);
                }
            }
        }
        __syncthreads();                                               // (1) table cleared, s_first and the chunk bases set
        {
            const u32 pr = s_pad_rel_from, ps = s_pad_seed_from;       // a chunk was closed: its unused tail becomes holes
            if (pr != SC_NOPAD)
                for (u32 i = pr + tid; i < REL_CHUNK; i += THREADS) { const u64 o = s_pad_rel_base + i; if (o < cap) out_h[o] = EMPTY_H; }
            if (EMIT && ps != SC_NOPAD)
            {
                Candidate hole; hole.kmer = EMPTY_KEY; hole.pos = 0; hole.read = 0;
                for (u32 i = ps + tid; i < SEED_CHUNK; i += THREADS) { const u64 o = s_pad_seed_base + i; if (o < seeds.cap) seeds.out[o] = hole; }
            }
        }
        const u32 i0 = tid * c;
        const u32 nv = i0 < total ? min(c, total - i0) : 0u;
        u32 r0 = 0, j0 = 0;
        u32 code[PER / 2];                                             // two 16-bit slot codes per word
        if (nv) { const u32 fs = s_first[tid]; r0 = fs >> 5; j0 = fs & 31u; }
        {
            u32 r = r0, j = j0, n = 0;
            SkmBases rec, nxt;
            rec.x = rec.y = 0; nxt = rec;
            if (nv)
            {
                rec = skm_load_bases(recs + r); nxt = rec;
                if (r + 1 < nrec) nxt = skm_load_bases(recs + r + 1);
                n = ((u32)rec.y & 31u) + 1u;
            }
#pragma unroll
            for (int g = 0; g < PER; g += 2)
            {
                if ((u32)g >= c) break;                                // uniform
                u64 H[2]; u32 S[2]; u64 P[2];
#pragma unroll
                for (int q = 0; q < 2; ++q)
                {
                    H[q] = EMPTY_H;
                    if ((u32)(g + q) < nv)
                    {
                        if (j == n)
                        {
                            ++r; j = 0; rec = nxt; n = ((u32)rec.y & 31u) + 1u;
                            if (r + 1 < nrec) nxt = skm_load_bases(recs + r + 1);
                        }
                        H[q] = mix64(canonical_of(skm_kmer(rec, j, kmask), lsh));
                        ++j;
                    }
                }
#pragma unroll
                for (int q = 0; q < 2; ++q)
                    if (H[q] != EMPTY_H) { S[q] = (u32)H[q] & (SLOTS - 1); P[q] = atomicCAS(&s_key[S[q]], EMPTY_H, H[q]); }
                u32 cw = 0;
#pragma unroll
                for (int q = 0; q < 2; ++q)
                    if (H[q] != EMPTY_H)
                    {
                        u32 s = S[q], stepp = 0; u64 pv = P[q];
                        while (pv != EMPTY_H && pv != H[q])            // triangular probing: every slot once; total <= 0.75 * slots
                        {
                            s = (s + ++stepp) & (SLOTS - 1);
                            pv = atomicCAS(&s_key[s], EMPTY_H, H[q]);
                        }
                        atomicAdd(&s_cnt[s], 1u);
                        cw |= (s | (pv == EMPTY_H ? SC_OWNER : 0u)) << (16 * q);
                    }
                code[g / 2] = cw;
            }
        }
        __syncthreads();                                               // (2) every count is final
        // tallies of this thread: k-mers it claimed (distinct), reliable ones among them, instances of reliable k-mers
        u32 tally2 = 0;
#pragma unroll
        for (int g = 0; g < PER; ++g)
        {
            if ((u32)g >= c) break;
            if ((u32)g < nv)
            {
                const u32 cd = (code[g / 2] >> (16 * (g & 1))) & 0xFFFFu;
                const u32 cc = s_cnt[cd & (SLOTS - 1)];
                const bool own = (cd & SC_OWNER) != 0, rel = cc >= lower && cc <= upper;
                my_distinct += own ? 1u : 0u;
                if (rel) tally2 += 0x10000u + (own ? 1u : 0u);
            }
        }
        if (tally2)
        {
            my_sum += tally2 >> 16; my_rel += tally2 & 0xFFFFu;
            u64 o_rel = s_rel_base + ((tally2 & 0xFFFFu) ? atomicAdd(&s_rel_used, tally2 & 0xFFFFu) : 0u);
            u64 o_seed = EMIT ? s_seed_base + atomicAdd(&s_seed_used, tally2 >> 16) : 0ull;
            u32 r = r0, j = j0;
            SkmBases rec = skm_load_bases(recs + r);
            u32 n = ((u32)rec.y & 31u) + 1u;
#pragma unroll
            for (int g = 0; g < PER; ++g)
            {
                if ((u32)g >= c) break;
                if ((u32)g < nv)
                {
                    if (j == n) { ++r; j = 0; rec = skm_load_bases(recs + r); n = ((u32)rec.y & 31u) + 1u; }
                    const u32 cd = (code[g / 2] >> (16 * (g & 1))) & 0xFFFFu;
                    const u32 s = cd & (SLOTS - 1);
                    const u32 cc = s_cnt[s];
                    if (cc >= lower && cc <= upper)
                    {
                        if (cd & SC_OWNER) { if (o_rel < cap) { out_h[o_rel] = s_key[s]; out_cnt[o_rel] = cc; } ++o_rel; }
                        if (EMIT)
                        {
                            if (o_seed < seeds.cap)
                            {
                                const u64 mt = __ldg(&recs[r].meta);
                                Candidate cnd; cnd.kmer = canonical_of(skm_kmer(rec, j, kmask), lsh); cnd.pos = (u32)mt + j; cnd.read = (u32)(mt >> 32);
                                seeds.out[o_seed] = cnd;
                            }
                            ++o_seed;
                        }
                    }
                    ++j;
                }
            }
        }
        __syncthreads();                                               // (3) the table and the chunk state are reused by the next bucket
    }
    // the unused tails of the CTA's last chunks are holes
    __syncthreads();
    for (u32 i = s_rel_used + tid; i < REL_CHUNK; i += THREADS) { const u64 o = s_rel_base + i; if (o < cap) out_h[o] = EMPTY_H; }
    if (EMIT)
    {
        Candidate hole; hole.kmer = EMPTY_KEY; hole.pos = 0; hole.read = 0;
        for (u32 i = s_seed_used + tid; i < SEED_CHUNK; i += THREADS) { const u64 o = s_seed_base + i; if (o < seeds.cap) seeds.out[o] = hole; }
    }
    for (int o = 16; o; o >>= 1)
    {
        my_distinct += __shfl_xor_sync(0xffffffffu, my_distinct, o); my_rel += __shfl_xor_sync(0xffffffffu, my_rel, o);
        my_sum += __shfl_xor_sync(0xffffffffu, my_sum, o);
    }
    if (lane == 0)
    {
        if (my_distinct) atomicAdd(&counters[2], (u64)my_distinct);
        if (my_sum) atomicAdd(&counters[1], my_sum);
        if (my_rel) atomicAdd(&counters[8], (u64)my_rel);
    }
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
            const SkmBases rec = skm_load_bases(list + i);
            const u32 n = ((u32)rec.y & 31u) + 1u;
            for (u32 j = 0; j < n; ++j) nd += table_insert(T, mix64(canonical_of(skm_kmer(rec, j, kmask), lsh)), 1u, err);
        }
    }
    tally(distinct, nd);
}

// pass 2 of the same fallback, before the table is collected and reset: every instance of the listed records whose
// k-mer ended with a reliable count appends its {k-mer, pos, read} to the seed list (rare path: one atomic per hit)
__global__ void __launch_bounds__(256) k_skm_emit_global(const SkmRec *__restrict__ list, u64 nrec, int k,
                                                         TableRef T, u32 lower, u32 upper, SeedSink seeds)
{
    const int lsh = 2 * (32 - k);
    const u64 kmask = (k == 32) ? ~0ull : (~0ull << lsh);
    const u64 step = (u64)gridDim.x * blockDim.x;
    for (u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x; i < nrec; i += step)
    {
        const SkmBases rec = skm_load_bases(list + i);
        const u64 mt = __ldg(&list[i].meta);
        const u32 n = ((u32)rec.y & 31u) + 1u;
        for (u32 j = 0; j < n; ++j)
        {
            const u64 x = canonical_of(skm_kmer(rec, j, kmask), lsh);
            const u64 h = mix64(x);
            u32 s = slot_of(h, T.slots);
            u32 cc = 0;
            for (u32 probes = 0; probes <= MAX_PROBES; ++probes)
            {
                const ulonglong2 v = __ldcg(reinterpret_cast<const ulonglong2*>(T.tab + s));
                if (v.x == h) { cc = (u32)v.y; break; }
                if (v.x == EMPTY_H) break;
                s = (s + 1 == T.slots) ? 0 : s + 1;
            }
            if (cc >= lower && cc <= upper)
            {
                const u64 o = atomicAdd(seeds.cursor, 1ull);
                if (o < seeds.cap) { Candidate cnd; cnd.kmer = x; cnd.pos = (u32)mt + j; cnd.read = (u32)(mt >> 32); seeds.out[o] = cnd; }
            }
        }
    }
}

} // namespace elba
