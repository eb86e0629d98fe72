// Counting the super-k-mer buckets (k >= 20), version 4: one bucket = one bulk copy + one small table.
// (reference: the unordered_map + Bloom of both counting passes and pass 2's kmermap.find, src/KmerOps.cpp:156-187,283-340)
//
// What the previous version (k_skm_count of round 1) was bound by (profiles/r1_v7_*, profiles/r2_ubench_smem_atomics.txt):
// 49 ms for 4.0 G instances = 0.56 shared-memory atomics per cycle per SM, an order of magnitude below what the SM sustains
// (6-12 lane-ops per cycle for the CAS.64 + ADD.32 pair at 1024-2048 threads); 250 thread instructions per instance; two
// buckets in flight per SM (a 96 KB table each, sized for the bucket's INSTANCES) with three or four exposed global-memory
// round trips per bucket (spare words, record words, record words again, meta words).
//
// This version:
//   * the table holds the bucket's DISTINCT k-mers: SLOTS = 2048 slots of {canonical k-mer (the key itself: no mix, no
//     un-mix), count16 | list index16} = 24 KB, so four CTAs (1024 threads) are resident per SM;
//   * a bucket's records are ONE contiguous slab: an elected thread brings it into shared memory with one
//     cp.async.bulk (UBLKCP) completing on an mbarrier, the slab of the CTA's next bucket is prefetched into L2
//     (cp.async.bulk.prefetch.L2) while this one is counted; every later access to a record is an LDS;
//   * a thread takes c consecutive instances; it finds its first record by a binary search over the records' first-instance
//     fields (in shared memory, no barrier), then walks records;
//   * after the counts are final, a SLOT pass (8 slots per thread) appends the reliable {k-mer, count} to the CTA's private
//     chunk of the reliable list and leaves the entry's list index in the slot; the INSTANCE pass then turns every
//     instance whose slot is reliable into a seed {list index, pos, read}: pass 2 of the reference fused, and the seed
//     already names its k-mer by an index, so build_A needs no k-mer -> column hash table, only perm[list index];
//   * a table that gets too full (more distinct k-mers than 3/4 of the slots: noisy reads) raises a flag; the bucket is
//     then handed, whole, to the exact global-table fallback.
#pragma once
#include "common.cuh"
#include "superkmer.cuh"

namespace elba {

// seed of pass 2: the k-mer is named by its index in the (holey) reliable list; id = SEED_HOLE marks an unused entry
struct __align__(16) Seed { u32 id, pos, read, pad; };
static constexpr u32 SEED_HOLE = 0xFFFFFFFFu;
struct SeedSink2 { Seed *out; u64 *cursor; u64 cap; };

// Records of the buckets this GPU counts.  The slab has one region per source GPU (one GPU: one source): region s starts at
// slab + s * nb * rcap.  The GPU's OWN region holds bucket b at b * rcap (the scatter's slots); the region of another source
// is packed: bucket b at off[s * nb + b] - off[s * nb] (exclusive scan of min(records offered, rcap)).  Either way the bucket's
// min(records offered by s, rcap) records lie contiguously in the order of their slots, and
// fill[s * nb + b] = (instances offered << 32) | records offered by source s (the scatter's reservation word, superkmer.cuh).
// plan[b] = the same summed over the sources, records = PLAN_SPILL if a source was offered more than rcap (one GPU: plan == fill).
struct RecSlabs
{
    const SkmRec *slab; const u64 *fill; const u64 *plan; const u64 *off; u32 rcap, nsrc, nb, me;
    __device__ __forceinline__ const SkmRec *bucket(u32 s, u64 b) const
    {
        const SkmRec *region = slab + (u64)s * nb * rcap;
        return (s == me) ? region + b * rcap : region + (__ldg(off + (u64)s * nb + b) - __ldg(off + (u64)s * nb));
    }
};
struct RecOverflow { SkmRec *list; u64 *cursor; u64 *inst; u64 cap; };
static constexpr u32 PLAN_SPILL = 0xFFFFFFFFu;

// several GPUs: plan[b] from the W fill words the sources sent
__global__ void k_skm_plan(const u64 *__restrict__ fill, u32 nb, u32 nsrc, u32 rcap, u64 *__restrict__ plan)
{
    const u32 b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= nb) return;
    u64 rec = 0, inst = 0; bool spill = false;
    for (u32 s = 0; s < nsrc; ++s) { const u64 fw = fill[(u64)s * nb + b]; const u32 f = (u32)fw; spill |= f > rcap; rec += f; inst += fw >> 32; }
    if (rec >= PLAN_SPILL || inst >= (1ull << 32)) spill = true;
    plan[b] = spill ? ((u64)(u32)min(inst, (u64)0xFFFFFFFFull) << 32) | PLAN_SPILL : (inst << 32) | rec;
}

static constexpr u32 SK4_CHUNK = 16384;          // entries of a CTA-private chunk of the reliable list / of the seed list
static constexpr u32 SK4_MAXPROBE = 48;          // probes after which a table counts as too full

__host__ __device__ constexpr u32 sk4_cap(u32 threads) { return threads * 24u; }                  // instances of one bucket (24 per thread)
__host__ __device__ constexpr size_t sk4_smem(u32 slots, u32 rmax, u32 pool, u32 retry, u32 threads) { return (size_t)rmax * 32 + (size_t)slots * 12 + (size_t)pool * 4 + (size_t)(threads / 32) * retry * 4 + 64; }

__device__ __forceinline__ u32 smem_u32(const void *p) { return (u32)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(u64 *bar, u32 count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(smem_u32(bar)), "r"(count) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(u64 *bar, u32 bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(u64 *bar, u32 parity)
{
    asm volatile("{\n\t.reg .pred p;\n\tWAIT_%=:\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t@p bra DONE_%=;\n\tbra WAIT_%=;\n\tDONE_%=:\n\t}"
                 :: "r"(smem_u32(bar)), "r"(parity) : "memory");
}
// global -> shared, bytes a multiple of 16, both addresses 16-byte aligned; completes on the mbarrier
__device__ __forceinline__ void bulk_g2s(void *dst, const void *src, u32 bytes, u64 *bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 :: "r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void bulk_prefetch_l2(const void *src, u32 bytes)
{
    asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" :: "l"(src), "r"(bytes) : "memory");
}
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// shared memory by its 32-bit window address (one register, no generic-pointer arithmetic in the inner loop)
__device__ __forceinline__ u64 lds_u64(u32 a) { u64 v; asm volatile("ld.shared.u64 %0, [%1];" : "=l"(v) : "r"(a) : "memory"); return v; }
__device__ __forceinline__ u32 lds_u32(u32 a) { u32 v; asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(a) : "memory"); return v; }
__device__ __forceinline__ void lds_v2u64(u32 a, u64 &x, u64 &y) { asm volatile("ld.shared.v2.u64 {%0, %1}, [%2];" : "=l"(x), "=l"(y) : "r"(a) : "memory"); }
__device__ __forceinline__ void sts_u32(u32 a, u32 v) { asm volatile("st.shared.u32 [%0], %1;" :: "r"(a), "r"(v) : "memory"); }
__device__ __forceinline__ u64 atoms_cas_u64(u32 a, u64 cmp, u64 val) { u64 o; asm volatile("atom.shared.cas.b64 %0, [%1], %2, %3;" : "=l"(o) : "r"(a), "l"(cmp), "l"(val) : "memory"); return o; }
__device__ __forceinline__ u32 atoms_add_u32(u32 a, u32 v) { u32 o; asm volatile("atom.shared.add.u32 %0, [%1], %2;" : "=r"(o) : "r"(a), "r"(v) : "memory"); return o; }
__device__ __forceinline__ u32 lanemask_lt() { u32 m; asm("mov.u32 %0, %%lanemask_lt;" : "=r"(m)); return m; }

// slot of a canonical k-mer inside a bucket's table: the k-mers of a bucket share their minimizer, so both words are mixed
template <int SLOTS>
__device__ __forceinline__ u32 sk4_slot(u64 x)
{
    constexpr int LG = (SLOTS == 1024) ? 10 : (SLOTS == 2048) ? 11 : (SLOTS == 4096) ? 12 : 13;
    const u32 t = ((u32)(x >> 32) * 0x9E3779B1u) ^ ((u32)x * 0x85EBCA77u);
    return (t * 0xC2B2AE3Du) >> (32 - LG);
}

// counters: [0] reliable-list entries handed out (chunks, with holes), [1] sum of the reliable counts, [2] distinct,
//           [8] reliable k-mers; seeds.cursor: seed entries handed out; ovf.cursor / ovf.inst: records / instances spilled
//
// Shared memory of one CTA: RMAX records (bulk-copied slab) | SLOTS keys | SLOTS {count16 | list index16} | POOL references.
// The POOL holds one reference {slot, record, k-mer in record} for every instance that arrived at its slot while the slot's
// count was still below `upper`: only those can be instances of a reliable k-mer (a k-mer with more than `upper` instances
// is dropped whatever its later instances do), so pass 2 looks at the pool (a few hundred entries) instead of walking all
// instances again, and no per-instance state lives in registers: the instance loop is a rolled loop of two instances.
template <int THREADS, int SLOTS, int RMAX, int POOL, int RETRY, int MINB>
__global__ void __launch_bounds__(THREADS, MINB) k_skm_count4(RecSlabs in, u32 nb, int k, RecOverflow ovf, u32 lower, u32 upper,
                                                           u64 *__restrict__ out_kmer, u32 *__restrict__ out_cnt,
                                                           u64 *__restrict__ counters, u64 cap, SeedSink2 seeds)
{
    constexpr u32 CAP = sk4_cap(THREADS);
    constexpr int SPT = SLOTS / THREADS;
    static_assert(SLOTS % THREADS == 0 && (SLOTS & (SLOTS - 1)) == 0 && SLOTS <= 65536 && CAP < 65536 && CAP <= SK4_CHUNK && RMAX <= 1024 && POOL <= 65536, "geometry");
    extern __shared__ __align__(128) unsigned char s_raw4[];
    SkmRec *s_rec = reinterpret_cast<SkmRec*>(s_raw4);                             // [RMAX]
    u64 *s_key = reinterpret_cast<u64*>(s_raw4 + (size_t)RMAX * 32);               // [SLOTS]
    u32 *s_cnt = reinterpret_cast<u32*>(s_key + SLOTS);                            // [SLOTS] count (low 16) | list index inside the chunk (high 16)
    u32 *s_pool = s_cnt + SLOTS;                                                   // [POOL] slot << 16 | record << 5 | k-mer in the record
    u32 *s_retry = s_pool + POOL;                                                  // [warps][RETRY] record << 5 | k-mer in the record
    u64 *s_bar = reinterpret_cast<u64*>(s_retry + (THREADS / 32) * RETRY);
    __shared__ u64 s_spill_base, s_rel_base, s_seed_base, s_pad_rel_base, s_pad_seed_base;
    __shared__ u32 s_rel_used, s_seed_used, s_pad_rel_from, s_pad_seed_from, s_full, s_pool_n;
    const u32 tid = threadIdx.x, lane = tid & 31;
    const u32 G = gridDim.x;
    const int lsh = 2 * (32 - k);
    const u64 kmask = (k == 32) ? ~0ull : (~0ull << lsh);
    const u32 a_rec = smem_u32(s_rec), a_key = smem_u32(s_key), a_cnt = smem_u32(s_cnt), a_pool = smem_u32(s_pool), a_pn = smem_u32(&s_pool_n), a_rl = smem_u32(s_retry);
    const u32 lt_mask = lanemask_lt();
    u32 my_distinct = 0, my_rel = 0; u64 my_sum = 0;
    __shared__ u32 s_src_rec[SK_MAXW + 1], s_src_inst[SK_MAXW + 1];               // prefix sums of the sources' records / instances in this bucket
    u32 b = blockIdx.x;
    u64 fw_cur = b < nb ? __ldg(in.plan + b) : 0ull;
    u64 fw_nxt = (u64)b + G < nb ? __ldg(in.plan + b + G) : 0ull;
    if (tid == 0)
    {
        s_rel_used = SK4_CHUNK; s_seed_used = SK4_CHUNK; s_rel_base = 0; s_seed_base = 0; s_full = 0; s_pool_n = 0;      // no chunk yet
        mbar_init(s_bar, 1);
    }
    __syncthreads();
    // does the CTA count the bucket with this plan word in shared memory?  (uniform)
    const u32 fmax = in.nsrc == 1 ? min(in.rcap, (u32)RMAX) : (u32)RMAX;
    auto fits = [&](u64 fw) { const u32 f = (u32)fw, tot = (u32)(fw >> 32); return f != 0 && f <= fmax && tot <= CAP; };
    // thread 0: bring bucket bb (plan word fw) into shared memory, one bulk copy per source, all completing on the mbarrier
    auto fetch = [&](u64 bb, u64 fw)
    {
        mbar_expect_tx(s_bar, (u32)fw * 32u);
        if (in.nsrc == 1)
        {
            s_src_rec[0] = 0; s_src_inst[0] = 0; s_src_rec[1] = (u32)fw; s_src_inst[1] = (u32)(fw >> 32);
            bulk_g2s(s_rec, in.bucket(0, bb), (u32)fw * 32u, s_bar);
        }
        else
        {
            u32 nr = 0, ni = 0;
            for (u32 sidx = 0; sidx < in.nsrc; ++sidx)
            {
                const u64 w = __ldg(in.fill + (u64)sidx * in.nb + bb);
                s_src_rec[sidx] = nr; s_src_inst[sidx] = ni;
                if ((u32)w) bulk_g2s(s_rec + nr, in.bucket(sidx, bb), min((u32)w, in.rcap) * 32u, s_bar);
                nr += (u32)w; ni += (u32)(w >> 32);
            }
            s_src_rec[in.nsrc] = nr; s_src_inst[in.nsrc] = ni;
        }
    };
    auto prefetch = [&](u64 bb, u64 fw)
    {
        if (in.nsrc == 1) { bulk_prefetch_l2(in.bucket(0, bb), (u32)fw * 32u); return; }
        for (u32 sidx = 0; sidx < in.nsrc; ++sidx)
        {
            const u32 fs = (u32)__ldg(in.fill + (u64)sidx * in.nb + bb);
            if (fs) bulk_prefetch_l2(in.bucket(sidx, bb), min(fs, in.rcap) * 32u);
        }
    };
    if (tid == 0)
    {
        if (fits(fw_cur)) fetch(b, fw_cur);
        if (fits(fw_nxt)) prefetch((u64)b + G, fw_nxt);
    }
    u32 parity = 0;
#pragma unroll 1
    for (; b < nb; b += G)
    {
        const u64 fw_nn = (u64)b + 2ull * G < nb ? __ldg(in.plan + b + 2u * G) : 0ull;
        const u32 f = (u32)fw_cur, total = (u32)(fw_cur >> 32);
        const bool here = fits(fw_cur);
        bool late_spill = false;
        if (here)
        {
            // room for this bucket in the CTA's chunks (worst case: every slot of the table reliable / every instance a seed)
            if (tid == 0)
            {
                u32 pr = SC_NOPAD, ps = SC_NOPAD;
                if (s_rel_used + min(total, (u32)SLOTS) > SK4_CHUNK)
                {
                    s_pad_rel_base = s_rel_base; pr = s_rel_used;
                    s_rel_base = atomicAdd(&counters[0], (u64)SK4_CHUNK); s_rel_used = 0;
                }
                if (s_seed_used + min(total, (u32)POOL) > SK4_CHUNK)
                {
                    s_pad_seed_base = s_seed_base; ps = s_seed_used;
                    s_seed_base = atomicAdd(seeds.cursor, (u64)SK4_CHUNK); s_seed_used = 0;
                }
                s_pad_rel_from = pr; s_pad_seed_from = ps;
            }
            // clear the table (the previous bucket ended with a barrier)
#pragma unroll
            for (int j = 0; j < SLOTS / 2 / THREADS; ++j)
            {
                ulonglong2 e; e.x = EMPTY_KEY; e.y = EMPTY_KEY;
                reinterpret_cast<ulonglong2*>(s_key)[j * THREADS + tid] = e;
            }
#pragma unroll
            for (int j = 0; j < SLOTS / 4 / THREADS; ++j) reinterpret_cast<uint4*>(s_cnt)[j * THREADS + tid] = make_uint4(0, 0, 0, 0);
            mbar_wait(s_bar, parity); parity ^= 1u;                        // the records are in shared memory
            __syncthreads();                                               // (1) table cleared, chunk bases set
            {
                const u32 pr = s_pad_rel_from, ps = s_pad_seed_from;       // a chunk was closed: its unused tail becomes holes (rare)
                if (pr != SC_NOPAD)
                {
#pragma unroll 1
                    for (u32 i = pr + tid; i < SK4_CHUNK; i += THREADS) { const u64 o = s_pad_rel_base + i; if (o < cap) out_kmer[o] = EMPTY_KEY; }
                }
                if (ps != SC_NOPAD)
                {
                    Seed hole; hole.id = SEED_HOLE; hole.pos = 0; hole.read = 0; hole.pad = 0;
#pragma unroll 1
                    for (u32 i = ps + tid; i < SK4_CHUNK; i += THREADS) { const u64 o = s_pad_seed_base + i; if (o < seeds.cap) seeds.out[o] = hole; }
                }
            }
            // my c consecutive instances start in the record that covers instance tid * c
            const u32 c = (total + THREADS - 1) / THREADS;                 // uniform
            const u32 i0 = tid * c;
            const u32 nv = i0 < total ? min(c, total - i0) : 0u;
            u32 ra = a_rec, j = 0, n = 0; u64 x = 0, y = 0;               // ra: shared address of my current record
            if (nv)
            {
                u32 src = 0;                                               // the source whose records hold instance i0
                while (i0 >= s_src_inst[src + 1]) ++src;
                const u32 il = i0 - s_src_inst[src];                       // instance inside that source's sub-slab
                u32 lo = s_src_rec[src], hi = s_src_rec[src + 1];          // first(lo) <= il < first(hi)
#pragma unroll 1
                while (hi - lo > 1) { const u32 mid = (lo + hi) >> 1; if ((u32)(lds_u64(a_rec + mid * 32u + 24u) >> 8) <= il) lo = mid; else hi = mid; }
                ra = a_rec + lo * 32u;
                lds_v2u64(ra, x, y);
                n = ((u32)y & 31u) + 1u; j = il - (u32)(lds_u64(ra + 24u) >> 8);
            }
            // Main loop: every lane takes one CAS at the home slot of its k-mer, the warp stays converged.  Lanes that find
            // another k-mer there (a fifth of them at these load factors) do NOT probe on the spot -- that loop would run for
            // a handful of lanes while the rest of the warp waits, every iteration -- but leave a reference in the warp's
            // private retry list; the list is worked off after the loop with all lanes busy.
            u32 nretry = 0;                                                // entries in this warp's list (uniform over the warp)
            const u32 a_retry = a_rl + (tid >> 5) * (RETRY * 4u);
            auto insert_probing = [&](u64 key, u32 sa) -> u32              // probes from the slot after sa; returns the slot address
            {
                u32 stepp = 0; u64 pv;
#pragma unroll 1
                do
                {
                    if (++stepp > SK4_MAXPROBE) { s_full = 1; break; }
                    sa = a_key + (((sa - a_key) + stepp * 8u) & (SLOTS * 8u - 8u));      // triangular probing: every slot once
                    pv = atoms_cas_u64(sa, EMPTY_KEY, key);
                } while (pv != EMPTY_KEY && pv != key);
                return sa;
            };
            auto count_and_pool = [&](u32 sa, u32 ref)
            {
                const u32 slot = (sa - a_key) >> 3;
                const u32 old = atoms_add_u32(a_cnt + slot * 4u, 1u);
                if (old < upper)                                           // can still belong to a reliable k-mer: remember the instance
                {
                    const u32 p = atoms_add_u32(a_pn, 1u);                // ptxas turns the same-address atomic of a warp into one reservation
                    if (p < (u32)POOL) sts_u32(a_pool + p * 4u, (slot << 16) | ref);
                }
            };
#pragma unroll 1
            for (u32 g = 0; g < c; g += 2)                                 // c is uniform: the warp stays converged at the loop head
            {
                u64 K[2], P[2]; u32 S[2], REF[2]; bool have[2];
#pragma unroll
                for (int q = 0; q < 2; ++q)
                {
                    have[q] = g + q < nv;
                    if (have[q] && j == n) { ra += 32u; j = 0; lds_v2u64(ra, x, y); n = ((u32)y & 31u) + 1u; }
                    const u32 sh = 2 * j;
                    const u64 fwd = (sh ? ((x << sh) | (y >> (64 - sh))) : x) & kmask;
                    K[q] = canonical_of(fwd, lsh);
                    S[q] = a_key + sk4_slot<SLOTS>(K[q]) * 8u;             // shared address of the home slot
                    REF[q] = (ra - a_rec) | j;                             // record * 32 | k-mer in the record
                    j += have[q] ? 1u : 0u;
                }
#pragma unroll
                for (int q = 0; q < 2; ++q) P[q] = have[q] ? atoms_cas_u64(S[q], EMPTY_KEY, K[q]) : K[q];
#pragma unroll
                for (int q = 0; q < 2; ++q)
                {
                    const bool coll = P[q] != EMPTY_KEY && P[q] != K[q];  // the home slot holds another k-mer
                    const unsigned m = __ballot_sync(0xffffffffu, coll);
                    bool later = false;
                    if (m)
                    {
                        const u32 pos = nretry + __popc(m & lt_mask);
                        later = coll && pos < (u32)RETRY;
                        if (later) sts_u32(a_retry + pos * 4u, REF[q]);
                        nretry += __popc(m);
                        if (coll && !later) S[q] = insert_probing(K[q], S[q]);         // the list is full (rare): probe here
                        __syncwarp();
                    }
                    if (have[q] && !later) count_and_pool(S[q], REF[q]);
                }
                __syncwarp();
            }
            // the retry list: one entry per lane, the k-mer is cut out of its record again
            nretry = min(nretry, (u32)RETRY);
            __syncwarp();
#pragma unroll 1
            for (u32 i = 0; i < nretry; i += 32)
            {
                const bool mine = i + lane < nretry;
                if (mine)
                {
                    const u32 ref = lds_u32(a_retry + (i + lane) * 4u);
                    u64 rx, ry;
                    lds_v2u64(a_rec + (ref & ~31u), rx, ry);
                    const u32 sh = 2 * (ref & 31u);
                    const u64 key = canonical_of((sh ? ((rx << sh) | (ry >> (64 - sh))) : rx) & kmask, lsh);
                    const u32 sa = insert_probing(key, a_key + sk4_slot<SLOTS>(key) * 8u);
                    count_and_pool(sa, ref);
                }
                __syncwarp();
            }
            __syncthreads();                                               // (2) every count is final
            const u32 npool = s_pool_n;
            late_spill = s_full != 0 || npool > (u32)POOL;                 // uniform (the flags are reset after barrier (4))
            if (!late_spill)
            {
                // slot pass: the reliable k-mers of this bucket go to the list; their list index stays in the slot
#pragma unroll
                for (int i = 0; i < SPT; ++i)
                {
                    const u32 s = i * THREADS + tid;
                    const u32 cc = s_cnt[s];
                    if (cc)
                    {
                        ++my_distinct;
                        if (cc >= lower && cc <= upper)
                        {
                            const u32 idx = atomicAdd(&s_rel_used, 1u);
                            const u64 o = s_rel_base + idx;
                            if (o < cap) { out_kmer[o] = s_key[s]; out_cnt[o] = cc; }
                            s_cnt[s] = cc | (idx << 16);
                            ++my_rel; my_sum += cc;
                        }
                    }
                }
                __syncthreads();                                           // (3) list indices visible
                // pool pass: every instance of a reliable k-mer becomes a seed (pass 2 of the reference, KmerOps.cpp:283-318)
                const u32 rel_lo = (u32)s_rel_base;                        // list indices are < 2^32 (checked on the host)
#pragma unroll 1
                for (u32 p = tid; p < npool; p += THREADS)
                {
                    const u32 e = s_pool[p];
                    const u32 w = s_cnt[e >> 16], cc = w & 0xFFFFu;
                    if (cc >= lower && cc <= upper)
                    {
                        const u64 o = s_seed_base + atomicAdd(&s_seed_used, 1u);
                        if (o < seeds.cap)
                        {
                            const u64 mt = s_rec[(e >> 5) & 0x7FFu].meta;
                            Seed sd; sd.id = rel_lo + (w >> 16); sd.pos = (u32)mt + (e & 31u); sd.read = (u32)(mt >> 32); sd.pad = 0;
                            seeds.out[o] = sd;
                        }
                    }
                }
            }
        }
        if (!here || late_spill)
        {
            // the bucket goes to the exact fallback, whole: too many records / instances, or too many distinct k-mers
#pragma unroll 1
            for (u32 sidx = 0; sidx < in.nsrc; ++sidx)
            {
                const u32 nrec = min((u32)__ldg(in.fill + (u64)sidx * in.nb + b), in.rcap);
                if (nrec == 0) continue;                                   // uniform
                if (tid == 0) s_spill_base = atomicAdd(ovf.cursor, (u64)nrec);
                __syncthreads();
                const SkmRec *__restrict__ recs = in.bucket(sidx, b);
                u32 ninst = 0;
#pragma unroll 1
                for (u32 rr = tid; rr < nrec; rr += THREADS)
                {
                    const SkmRec rec = skm_load(recs + rr);
                    ninst += ((u32)rec.y & 31u) + 1u;
                    const u64 o = s_spill_base + rr;
                    if (o < ovf.cap) skm_store(ovf.list + o, rec.x, rec.y, rec.meta, rec.spare);
                }
                for (int o = 16; o; o >>= 1) ninst += __shfl_xor_sync(0xffffffffu, ninst, o);
                if (lane == 0 && ninst) atomicAdd(ovf.inst, (u64)ninst);
                __syncthreads();
            }
        }
        __syncthreads();                                                   // (4) table, records, pool and chunk state are free
        fw_cur = fw_nxt; fw_nxt = fw_nn;
        if (tid == 0)
        {
            s_full = 0; s_pool_n = 0;
            const u64 bn = (u64)b + G;
            if (bn < nb && fits(fw_cur)) { fence_async_smem(); fetch(bn, fw_cur); }
            if (bn + G < nb && fits(fw_nxt)) prefetch(bn + G, fw_nxt);
        }
    }
    // the unused tails of the CTA's last chunks are holes
    __syncthreads();
#pragma unroll 1
    for (u32 i = s_rel_used + tid; i < SK4_CHUNK; i += THREADS) { const u64 o = s_rel_base + i; if (o < cap) out_kmer[o] = EMPTY_KEY; }
    {
        Seed hole; hole.id = SEED_HOLE; hole.pos = 0; hole.read = 0; hole.pad = 0;
#pragma unroll 1
        for (u32 i = s_seed_used + tid; i < SK4_CHUNK; i += THREADS) { const u64 o = s_seed_base + i; if (o < seeds.cap) seeds.out[o] = hole; }
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

// ---- the exact fallback for spilled buckets: global table of kmer_count.cuh keyed by the canonical k-mer itself ---------
__device__ __forceinline__ u32 sk4_gslot(u64 x, u32 slots) { return __umulhi((u32)mix64(x), slots); }

__global__ void __launch_bounds__(256) k_skm4_count_global(const SkmRec *__restrict__ list, const u64 *__restrict__ nrec_p, u64 list_cap, int k, TableRef T,
                                                           u32 *__restrict__ err, u64 *__restrict__ distinct)
{
    const int lsh = 2 * (32 - k);
    const u64 kmask = (k == 32) ? ~0ull : (~0ull << lsh);
    const u64 nrec = min(*nrec_p, list_cap);
    const u64 step = (u64)gridDim.x * blockDim.x;
    u32 nd = 0;
    for (u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x; i < nrec; i += step)
    {
        const SkmBases rec = skm_load_bases(list + i);
        const u32 n = ((u32)rec.y & 31u) + 1u;
        for (u32 j = 0; j < n; ++j)
        {
            const u64 x = canonical_of(skm_kmer(rec, j, kmask), lsh);
            u32 s = sk4_gslot(x, T.slots);
            u64 prev = atomicCAS(&T.tab[s].key, EMPTY_KEY, x);
            u32 probes = 0;
            while (prev != EMPTY_KEY && prev != x)
            {
                if (++probes > MAX_PROBES) { atomicOr(err, 1u); break; }
                s = (s + 1 == T.slots) ? 0 : s + 1;
                prev = atomicCAS(&T.tab[s].key, EMPTY_KEY, x);
            }
            atomicAdd(&T.tab[s].cnt, 1u);
            nd += prev == EMPTY_KEY;
        }
    }
    for (int o = 16; o; o >>= 1) nd += __shfl_xor_sync(0xffffffffu, nd, o);
    if ((threadIdx.x & 31) == 0 && nd) atomicAdd(distinct, (u64)nd);
}

// the reliable k-mers of the fallback table go to the list one by one; the list index stays in the slot's aux word.
// counters as in k_skm_count4.  `slots` is the table's capacity; untouched slots hold EMPTY_KEY.
__global__ void __launch_bounds__(256) k_skm4_collect_global(Slot *__restrict__ tab, u32 slots, u32 lower, u32 upper,
                                                             u64 *__restrict__ out_kmer, u32 *__restrict__ out_cnt, u64 *__restrict__ counters, u64 cap)
{
    const u32 step = gridDim.x * blockDim.x;
    u32 nrel = 0; u64 sum = 0;
    for (u32 i = blockIdx.x * blockDim.x + threadIdx.x; i < slots; i += step)
    {
        const ulonglong2 v = __ldcg(reinterpret_cast<const ulonglong2*>(tab + i));
        const u32 cnt = (u32)v.y;
        if (v.x != EMPTY_KEY && cnt >= lower && cnt <= upper)
        {
            const u64 o = atomicAdd(&counters[0], 1ull);
            if (o < cap) { out_kmer[o] = v.x; out_cnt[o] = cnt; }
            tab[i].aux = (u32)o;
            ++nrel; sum += cnt;
        }
    }
    for (int o = 16; o; o >>= 1) { nrel += __shfl_xor_sync(0xffffffffu, nrel, o); sum += __shfl_xor_sync(0xffffffffu, sum, o); }
    if ((threadIdx.x & 31) == 0 && nrel) { atomicAdd(&counters[8], (u64)nrel); atomicAdd(&counters[1], sum); }
}

// pass 2 of the fallback: every instance of the listed records whose k-mer ended reliable becomes a seed
__global__ void __launch_bounds__(256) k_skm4_emit_global(const SkmRec *__restrict__ list, const u64 *__restrict__ nrec_p, u64 list_cap, int k,
                                                          TableRef T, u32 lower, u32 upper, SeedSink2 seeds)
{
    const int lsh = 2 * (32 - k);
    const u64 kmask = (k == 32) ? ~0ull : (~0ull << lsh);
    const u64 nrec = min(*nrec_p, list_cap);
    const u64 step = (u64)gridDim.x * blockDim.x;
    for (u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x; i < nrec; i += step)
    {
        const SkmBases rec = skm_load_bases(list + i);
        const u64 mt = __ldg(&list[i].meta);
        const u32 n = ((u32)rec.y & 31u) + 1u;
        for (u32 j = 0; j < n; ++j)
        {
            const u64 x = canonical_of(skm_kmer(rec, j, kmask), lsh);
            u32 s = sk4_gslot(x, T.slots);
            u32 cc = 0, id = 0;
            for (u32 probes = 0; probes <= MAX_PROBES; ++probes)
            {
                const ulonglong2 v = __ldcg(reinterpret_cast<const ulonglong2*>(T.tab + s));
                if (v.x == x) { cc = (u32)v.y; id = (u32)(v.y >> 32); break; }
                if (v.x == EMPTY_KEY) break;
                s = (s + 1 == T.slots) ? 0 : s + 1;
            }
            if (cc >= lower && cc <= upper)
            {
                const u64 o = atomicAdd(seeds.cursor, 1ull);
                if (o < seeds.cap) { Seed sd; sd.id = id; sd.pos = (u32)mt + j; sd.read = (u32)(mt >> 32); sd.pad = 0; seeds.out[o] = sd; }
            }
        }
    }
}

// ---- column ids without a hash table --------------------------------------------------------------------
// the reliable list (with holes) was sorted by k-mer value with its own index as payload: rank r holds list entry idx[r].
// perm[list index] = column id; the counts follow their k-mers.
__global__ void k_iota_u32(u32 *__restrict__ v, u64 n)
{
    const u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) v[i] = (u32)i;
}
__global__ void k_rank_finish(const u32 *__restrict__ idx_sorted, const u32 *__restrict__ cnt_list, u64 R, u32 *__restrict__ perm, u32 *__restrict__ cnt_sorted)
{
    const u64 r = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= R) return;
    const u32 i = idx_sorted[r];
    perm[i] = (u32)r;
    cnt_sorted[r] = cnt_list[i];
}

// several GPUs: every GPU sorted the reliable k-mers it owns; all[off[r] ...] is rank r's sorted run (all-gathered).  The
// column id of my k-mer = its rank in my run + the number of smaller k-mers in every other run (k-mers are distinct
// across the runs: a k-mer has one owner).  Consecutive threads search neighbouring keys: the probes hit L1 / L2.
struct RankRuns { u64 off[SK_MAXW]; u64 n[SK_MAXW]; u32 nruns, me; };
static constexpr int RG_KEYS = 1024;     // keys of my run one CTA ranks
__global__ void __launch_bounds__(256) k_rank_global(const u64 *__restrict__ mine, const u32 *__restrict__ idx_sorted, const u32 *__restrict__ cnt_list, u64 n_mine,
                                                     const u64 *__restrict__ all, RankRuns runs, u32 *__restrict__ perm, u32 *__restrict__ cnt_sorted, u32 *__restrict__ gid)
{
    // my run is sorted: the keys of this CTA fall into one narrow range of every other run; two long searches per run and CTA
    // bound that range, the per-key searches then stay inside it (a few KB: L1)
    __shared__ u64 s_lo[SK_MAXW], s_hi[SK_MAXW];
    const u64 b0 = (u64)blockIdx.x * RG_KEYS;
    if (b0 >= n_mine) return;
    const u64 b1 = min(b0 + (u64)RG_KEYS, n_mine);
    if (threadIdx.x < runs.nruns && threadIdx.x != runs.me)
    {
        const u64 *__restrict__ run = all + runs.off[threadIdx.x];
        const u64 kf = mine[b0], kl = mine[b1 - 1];
        u64 lo = 0, hi = runs.n[threadIdx.x];
        while (lo < hi) { const u64 mid = (lo + hi) >> 1; if (__ldg(run + mid) < kf) lo = mid + 1; else hi = mid; }
        s_lo[threadIdx.x] = lo;
        hi = runs.n[threadIdx.x];
        u64 l2 = lo;
        while (l2 < hi) { const u64 mid = (l2 + hi) >> 1; if (__ldg(run + mid) < kl) l2 = mid + 1; else hi = mid; }
        s_hi[threadIdx.x] = l2;
    }
    __syncthreads();
    for (u64 r = b0 + threadIdx.x; r < b1; r += blockDim.x)
    {
        const u64 key = mine[r];
        u64 g = r;
        for (u32 q = 0; q < runs.nruns; ++q)
        {
            if (q == runs.me) continue;
            const u64 *__restrict__ run = all + runs.off[q];
            u64 lo = s_lo[q], hi = s_hi[q];
            while (lo < hi) { const u64 mid = (lo + hi) >> 1; if (__ldg(run + mid) < key) lo = mid + 1; else hi = mid; }
            g += lo;
        }
        const u32 i = idx_sorted[r];
        perm[i] = (u32)g; gid[r] = (u32)g; cnt_sorted[r] = cnt_list[i];
    }
}

__global__ void k_place_by_id(const u64 *__restrict__ key, const u32 *__restrict__ cnt, const u32 *__restrict__ gid, u64 n, u64 *__restrict__ okey, u32 *__restrict__ ocnt)
{
    const u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) { const u32 g = gid[i]; okey[g] = key[i]; ocnt[g] = cnt[i]; }
}

// seeds -> sort keys (read << col_bits | column) and positions; holes get `hole_key` (sorted behind every entry)
__global__ void __launch_bounds__(256) k_seed_keys(const Seed *__restrict__ seeds, u64 n, const u32 *__restrict__ perm, u32 id_base, int col_bits, u64 hole_key,
                                                   u64 *__restrict__ out_key, u32 *__restrict__ out_pos, u64 *__restrict__ nvalid)
{
    const u64 step = (u64)gridDim.x * blockDim.x;
    u32 mine = 0;
    for (u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += step)
    {
        const uint4 s = __ldcs(reinterpret_cast<const uint4*>(seeds + i));
        u64 key = hole_key; u32 pos = 0;
        if (s.x != SEED_HOLE) { key = ((u64)s.z << col_bits) | __ldg(perm + id_base + s.x); pos = s.y; ++mine; }
        out_key[i] = key; out_pos[i] = pos;
    }
    for (int o = 16; o; o >>= 1) mine += __shfl_xor_sync(0xffffffffu, mine, o);
    if ((threadIdx.x & 31) == 0 && mine) atomicAdd(nvalid, (u64)mine);
}

// ---- several GPUs: the k-mer owners send (global read, column, pos) to the read owners through peer memory ----------
// key[r] / pos[r] / cnt[r]: rank r's receive window, mapped here; region [source][cap].  first[r]: the first global read id of
// rank r's block (first[nranks] = one past the last read).  cursor: local, one per destination.
struct RouteSink { u64 *key[SK_MAXW]; u32 *pos[SK_MAXW]; u64 *cnt[SK_MAXW]; u64 first[SK_MAXW + 1]; u64 *cursor; u64 cap; u32 nranks, me; };

__global__ void __launch_bounds__(256) k_seed_route(const Seed *__restrict__ seeds, u64 n, const u32 *__restrict__ perm, RouteSink rs)
{
    __shared__ u32 s_cnt[SK_MAXW];
    __shared__ u64 s_base[SK_MAXW];
    const u64 tile = (u64)blockDim.x * 4;
    for (u64 t0 = (u64)blockIdx.x * tile; t0 < n; t0 += (u64)gridDim.x * tile)
    {
        if (threadIdx.x < SK_MAXW) s_cnt[threadIdx.x] = 0;
        __syncthreads();
        uint4 sd[4]; u32 dst[4], slot[4];
#pragma unroll
        for (int i = 0; i < 4; ++i)
        {
            const u64 idx = t0 + (u64)i * blockDim.x + threadIdx.x;
            sd[i] = make_uint4(SEED_HOLE, 0, 0, 0);
            if (idx < n) sd[i] = __ldcs(reinterpret_cast<const uint4*>(seeds + idx));
            dst[i] = 0xFFFFFFFFu; slot[i] = 0;
            if (sd[i].x != SEED_HOLE)
            {
                u32 d = 0;
                while (d + 1 < rs.nranks && (u64)sd[i].z >= rs.first[d + 1]) ++d;      // owner of the read: the blocks are consecutive
                dst[i] = d;
                slot[i] = atomicAdd(&s_cnt[d], 1u);
            }
        }
        __syncthreads();
        if (threadIdx.x < rs.nranks) s_base[threadIdx.x] = s_cnt[threadIdx.x] ? atomicAdd(rs.cursor + threadIdx.x, (u64)s_cnt[threadIdx.x]) : 0ull;
        __syncthreads();
#pragma unroll
        for (int i = 0; i < 4; ++i)
            if (dst[i] != 0xFFFFFFFFu)
            {
                const u64 o = s_base[dst[i]] + slot[i];
                if (o < rs.cap)
                {
                    const u64 at = (u64)rs.me * rs.cap + o;
                    rs.key[dst[i]][at] = ((u64)sd[i].z << 32) | __ldg(perm + sd[i].x);
                    rs.pos[dst[i]][at] = sd[i].y;
                }
            }
        __syncthreads();
    }
    __threadfence_system();
}
// after the route kernel: every destination learns how many triples this source stored for it
__global__ void k_route_publish(RouteSink rs)
{
    if (threadIdx.x < rs.nranks) rs.cnt[threadIdx.x][rs.me] = rs.cursor[threadIdx.x];
    __threadfence_system();
}
// the received triples of one source -> sort keys with LOCAL read ids (read << col_bits | column)
__global__ void k_route_unpack(const u64 *__restrict__ key, const u32 *__restrict__ pos, u64 n, u64 read0, int col_bits, u64 *__restrict__ out_key, u32 *__restrict__ out_pos)
{
    const u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const u64 kk = key[i];
    out_key[i] = (((kk >> 32) - read0) << col_bits) | (kk & 0xFFFFFFFFull);
    out_pos[i] = pos[i];
}

} // namespace elba
