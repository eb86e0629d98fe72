/*
 * TEST INFRASTRUCTURE ONLY (oracle/).  Nothing under elba_b200/ may include,
 * link or execute this; only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs load the library built from it.
 *
 * oracle/_ref/libelba_ref_k<K>_l<L>_u<U>.so = the REFERENCE'S OWN sources,
 * compiled unmodified from where they lie under /root/reference
 * (src/{DnaSeq,DnaBuffer,HashFuncs,Bloom,HyperLogLog,Logger,KmerOps,
 * SharedSeeds}.cpp + include/), against oracle/stubs/{mpi.h,CombBLAS/} and
 * this C-ABI shim.  It is what pins oracle/elba_oracle.cpp.
 *
 * What is genuinely the reference here: every k-mer, hash, HyperLogLog, Bloom,
 * DnaSeq/DnaBuffer operation, both counting passes of KmerOps.cpp, the triple
 * generation of create_kmer_matrix and the SharedSeeds semiring + Prune call.
 * What is a restatement: MPI (threads, stubs/mpi.h) and the five CombBLAS
 * entry points (stubs/CombBLAS/CombBLAS.h — unpinned external dependency).
 *
 * K, LOWER, UPPER are compile-time macros in the reference
 * (include/compiletime.h:7-22), hence one library per (K,L,U).
 */
#include <cstring>
#include <cstdint>
#include <cstddef>
#include <numeric>
#include <algorithm>
#include <iomanip>
#include <cmath>
#include <thread>
#include <chrono>
#include <mpi.h>
#include "common.h"             /* pulls the stubs + every std header first */
#define private public          /* HyperLogLog::registers, Bloom::bf are private */
#include "HyperLogLog.hpp"
#include "Bloom.hpp"
#undef private
#include "KmerOps.hpp"
#include "SharedSeeds.hpp"
#include "HashFuncs.hpp"
#include "Logger.hpp"
#include "DnaSeq.hpp"
#include "DnaBuffer.hpp"
#include "Overlap.hpp"
#include "XDropAligner.hpp"
#include "FastaIndex.hpp"
#include "TransitiveReduction.hpp"
#include <cstring>
#include <numeric>
#include <algorithm>
#include <iomanip>
#include <cmath>
#include <thread>
#include <chrono>

/*
 * src/KmerOps.cpp keeps its Bloom filter in a file-static pointer
 * (KmerOps.cpp:12-13).  Ranks are threads here, so that one static must be
 * per-thread; every header the file includes is already included above
 * (guards), so the macro below touches that single declaration only.
 */
#ifdef REF_IS_SHIM
/* the "shim" flavour (oracle/Makefile target `shim`): the same driver sequence, but the five functions come from
 * elba_b200/host/elba_fe_shim.cpp + libelba_fe.so instead of the reference's KmerOps.cpp / SharedSeeds.cpp: this is
 * how the drop-in boundary is EXECUTED (tests/test_gpu_shim.py), with the reference's headers and the same stubs. */
#include REF_KMEROPS_CPP   /* = "<repo>/elba_b200/host/elba_fe_shim.cpp" */
#else
#define static static thread_local
#include REF_KMEROPS_CPP   /* = "<reference>/src/KmerOps.cpp", set by oracle/Makefile */
#undef static
#endif

namespace {

struct RefResult
{
    int nranks = 1;
    /* reliable k-mers in the reference's own column-id order */
    std::vector<uint64_t> kmer;
    std::vector<int32_t> count;
    std::vector<int64_t> reads;      /* R x UPPER, arrival order */
    std::vector<uint32_t> pos;       /* R x UPPER */
    std::vector<uint64_t> keys_after_pass1;   /* per rank: map size after pass 1 */
    /* A */
    int64_t a_m = 0, a_n = 0;
    std::vector<int64_t> a_rowptr, a_col;
    std::vector<uint32_t> a_val;
    /* B after prune */
    std::vector<int64_t> b_rowptr, b_col;
    std::vector<int32_t> b_num;
    std::vector<uint32_t> b_seeds;   /* nnz x 4: s0.q s0.t s1.q s1.t */
    int64_t b_nnz_preprune = 0;
    double secs[6] = {0,0,0,0,0,0};  /* keys, values, matrix, transpose, spgemm, total */
};

double now() { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); }

/* contiguous, base-balanced blocks: src/FastaIndex.cpp:47-94 */
std::vector<uint64_t> partition_reads(const uint64_t *lens, uint64_t n, int P)
{
    std::vector<uint64_t> start(P + 1, n);
    double tot = 0; for (uint64_t i = 0; i < n; ++i) tot += (double)lens[i];
    double avg = tot / P;
    uint64_t id = 0;
    for (int r = 0; r < P - 1; ++r)
    {
        start[r] = id;
        double sofar = 0;
        while (id < n && sofar + (double)lens[id] < avg) { sofar += (double)lens[id]; id++; }
    }
    start[P-1] = id; start[P] = n;
    if (P == 1) start[0] = 0;
    return start;
}

} // namespace

extern "C" {

void ref_params(int *k, int *l, int *u) { *k = KMER_SIZE; *l = LOWER_KMER_FREQ; *u = UPPER_KMER_FREQ; }

/* DnaSeq::compress, src/DnaSeq.cpp:7-29 */
void ref_pack(const char *s, uint64_t len, uint8_t *mem) { DnaSeq seq(s, len, mem); (void)seq; }

/* Kmer ctor/GetTwin/GetRep/GetHash on a k-character ASCII k-mer */
void ref_kmer_info(const char *s, uint64_t *fwd, uint64_t *twin, uint64_t *rep, uint64_t *hash)
{
    TKmer f(s), t = f.GetTwin(), r = f.GetRep();
    f.CopyDataInto(fwd); t.CopyDataInto(twin); r.CopyDataInto(rep);
    *hash = r.GetHash();
}

uint64_t ref_hash(uint64_t kmer) { TKmer m((const void*)&kmer); return m.GetHash(); }
int ref_owner(uint64_t kmer, int nprocs) { TKmer m((const void*)&kmer); return GetKmerOwner(m, nprocs); }

/* TKmer::GetRepKmers over one packed read, src/Kmer.cpp:236-242 */
uint64_t ref_rep_kmers(const uint8_t *packed, uint64_t len, uint64_t *out)
{
    DnaSeq seq(len, const_cast<uint8_t*>(packed));
    if (len < KMER_SIZE) return 0;
    auto v = TKmer::GetRepKmers(seq);
    for (size_t i = 0; i < v.size(); ++i) v[i].CopyDataInto(out + i);
    return v.size();
}

/* HyperLogLog exactly as get_kmer_count_map_keys drives it: KmerOps.cpp:45-47 */
double ref_hll(const uint8_t *packed, const uint64_t *lens, uint64_t nreads, uint8_t *regs_out)
{
    size_t bufsize = 0; for (uint64_t i = 0; i < nreads; ++i) bufsize += DnaSeq::bytesneeded(lens[i]);
    uint8_t *buf = new uint8_t[bufsize ? bufsize : 1]; std::memcpy(buf, packed, bufsize);
    std::vector<size_t> l(lens, lens + nreads);
    DnaBuffer dna(bufsize, nreads, buf, l.data());
    HyperLogLog hll;
    KmerEstimateHandler estimator(hll);
    ForeachKmer(dna, estimator);
    if (regs_out) std::memcpy(regs_out, hll.registers.data(), 4096);
    return hll.estimate();
}

/* Bloom */
void *ref_bloom_new(int64_t entries, double err) { return new Bloom(entries, err); }
void ref_bloom_free(void *b) { delete (Bloom*)b; }
int64_t ref_bloom_bits(void *b) { return ((Bloom*)b)->bits; }
int ref_bloom_hashes(void *b) { return ((Bloom*)b)->hashes; }
int ref_bloom_check(void *b, uint64_t kmer) { return ((Bloom*)b)->Check(&kmer, 8); }
int ref_bloom_add(void *b, uint64_t kmer) { return ((Bloom*)b)->Add(&kmer, 8); }
void ref_bloom_copy(void *b, uint8_t *out) { std::memcpy(out, ((Bloom*)b)->bf, ((Bloom*)b)->bytes); }

/*
 * The hot path exactly as src/main.cpp:191-282 sequences it, on `nranks`
 * thread-ranks: get_kmer_count_map_keys -> get_kmer_count_map_values ->
 * create_kmer_matrix -> copy + Transpose -> create_seed_matrix.
 */
static void *run_impl(const uint8_t *packed, const uint64_t *lens, uint64_t nreads, int nranks, const char *fasta)
{
    RefResult *res = new RefResult; res->nranks = nranks;
    std::vector<uint64_t> start, byteoff(nreads + 1, 0);
    if (!fasta)
    {
        start = partition_reads(lens, nreads, nranks);
        for (uint64_t i = 0; i < nreads; ++i) byteoff[i+1] = byteoff[i] + DnaSeq::bytesneeded(lens[i]);
    }

    fake_mpi::World world(nranks);
    fake_mpi::g_world = &world;
    res->keys_after_pass1.assign(nranks, 0);

    struct PerRank { std::vector<uint64_t> kmer; std::vector<int32_t> count; std::vector<int64_t> reads; std::vector<uint32_t> pos; };
    std::vector<PerRank> per(nranks);
    std::vector<double> t(nranks * 6, 0.0);

    auto body = [&](int rank)
    {
        fake_mpi::t_rank = rank;
        auto commgrid = std::make_shared<CommGrid>(MPI_COMM_WORLD, 0, 0);
        /* the reads: a block of the caller's arena, or (src/main.cpp:130-139) FastaIndex index(fasta, commgrid); index.getmydna() */
        std::unique_ptr<FastaIndex> index;
        if (fasta) index = std::make_unique<FastaIndex>(fasta, commgrid);
        auto from_arena = [&]()
        {
            uint64_t r0 = start[rank], r1 = start[rank+1];
            size_t bufsize = byteoff[r1] - byteoff[r0];
            uint8_t *buf = new uint8_t[bufsize ? bufsize : 1]; std::memcpy(buf, packed + byteoff[r0], bufsize);
            std::vector<size_t> l(lens + r0, lens + r1);
            return DnaBuffer(bufsize, r1 - r0, buf, l.data());
        };
        DnaBuffer mydna = fasta ? index->getmydna() : from_arena();
        double *tt = &t[rank * 6];

        MPI_Barrier(MPI_COMM_WORLD); double t0 = now(), ta = t0;
        auto kmermap = get_kmer_count_map_keys(mydna, commgrid);
        MPI_Barrier(MPI_COMM_WORLD); tt[0] = now() - ta; ta = now();
        res->keys_after_pass1[rank] = kmermap->size();
        get_kmer_count_map_values(mydna, *kmermap, commgrid);
        MPI_Barrier(MPI_COMM_WORLD); tt[1] = now() - ta;

        /* snapshot the map in its own iteration order == the reference's column-id order */
        PerRank& pr = per[rank];
        for (auto it = kmermap->cbegin(); it != kmermap->cend(); ++it)
        {
            uint64_t x; it->first.CopyDataInto(&x);
            pr.kmer.push_back(x); pr.count.push_back(std::get<2>(it->second));
            for (int j = 0; j < UPPER_KMER_FREQ; ++j) { pr.reads.push_back(std::get<0>(it->second)[j]); pr.pos.push_back(std::get<1>(it->second)[j]); }
        }

        MPI_Barrier(MPI_COMM_WORLD); ta = now();
        auto A = create_kmer_matrix(mydna, *kmermap, commgrid);
        MPI_Barrier(MPI_COMM_WORLD); tt[2] = now() - ta; ta = now();
        kmermap.reset();
        auto AT = std::make_unique<CT<PosInRead>::PSpParMat>(*A);
        AT->Transpose();
        MPI_Barrier(MPI_COMM_WORLD); tt[3] = now() - ta; ta = now();

        /* create_seed_matrix (src/SharedSeeds.cpp:4-10), split only to read nnz before Prune */
        auto Bpre = Mult_AnXBn_DoubleBuff<SharedSeeds::Semiring, SharedSeeds, CT<SharedSeeds>::PSpDCCols>(*A, *AT);
        int64_t pre = Bpre.getnnz();
        auto B = create_seed_matrix(*A, *AT);
        MPI_Barrier(MPI_COMM_WORLD); tt[4] = (now() - ta) / 2.0; tt[5] = now() - t0 - tt[4];

        if (rank == 0)
        {
            res->a_m = A->getnrow(); res->a_n = A->getncol();
            res->a_rowptr.assign(A->st->rowptr.begin(), A->st->rowptr.end());
            res->a_col.assign(A->st->col.begin(), A->st->col.end());
            res->a_val.assign(A->st->val.begin(), A->st->val.end());
            res->b_nnz_preprune = pre;
            res->b_rowptr.assign(B->st->rowptr.begin(), B->st->rowptr.end());
            res->b_col.assign(B->st->col.begin(), B->st->col.end());
            for (const SharedSeeds& s : B->st->val)
            {
                res->b_num.push_back(s.getnumshared());
                res->b_seeds.push_back(std::get<0>(s.getseeds()[0])); res->b_seeds.push_back(std::get<1>(s.getseeds()[0]));
                res->b_seeds.push_back(std::get<0>(s.getseeds()[1])); res->b_seeds.push_back(std::get<1>(s.getseeds()[1]));
            }
        }
        MPI_Barrier(MPI_COMM_WORLD);
    };

    if (nranks == 1) body(0);
    else
    {
        std::vector<std::thread> th;
        for (int r = 0; r < nranks; ++r) th.emplace_back(body, r);
        for (auto& x : th) x.join();
    }
    fake_mpi::g_world = nullptr; fake_mpi::t_rank = 0;

    for (int r = 0; r < nranks; ++r)
    {
        res->kmer.insert(res->kmer.end(), per[r].kmer.begin(), per[r].kmer.end());
        res->count.insert(res->count.end(), per[r].count.begin(), per[r].count.end());
        res->reads.insert(res->reads.end(), per[r].reads.begin(), per[r].reads.end());
        res->pos.insert(res->pos.end(), per[r].pos.begin(), per[r].pos.end());
    }
    for (int s = 0; s < 6; ++s) { double mx = 0; for (int r = 0; r < nranks; ++r) mx = std::max(mx, t[r*6+s]); res->secs[s] = mx; }
    return res;
}

void *ref_run(const uint8_t *packed, const uint64_t *lens, uint64_t nreads, int nranks) { return run_impl(packed, lens, nreads, nranks, nullptr); }
/* the same from a FASTA file + its .fai: FastaIndex -> getmydna -> the five functions (src/main.cpp:130-139,191-282) */
void *ref_run_fasta(const char *fasta, int nranks) { return run_impl(nullptr, nullptr, 0, nranks, fasta); }

void ref_free(void *h) { delete (RefResult*)h; }
void ref_sizes(void *h, int64_t *out /*[8]*/)
{
    RefResult *r = (RefResult*)h;
    out[0] = r->kmer.size(); out[1] = r->a_m; out[2] = r->a_n; out[3] = r->a_col.size();
    out[4] = r->b_col.size(); out[5] = r->b_nnz_preprune; out[6] = 0; for (auto v : r->keys_after_pass1) out[6] += v; out[7] = r->nranks;
}
void ref_secs(void *h, double *out /*[6]*/) { std::memcpy(out, ((RefResult*)h)->secs, sizeof(double) * 6); }
void ref_get_kmers(void *h, uint64_t *kmer, int32_t *count, int64_t *reads, uint32_t *pos)
{
    RefResult *r = (RefResult*)h;
    std::memcpy(kmer, r->kmer.data(), r->kmer.size() * 8); std::memcpy(count, r->count.data(), r->count.size() * 4);
    std::memcpy(reads, r->reads.data(), r->reads.size() * 8); std::memcpy(pos, r->pos.data(), r->pos.size() * 4);
}
void ref_get_A(void *h, int64_t *rowptr, int64_t *col, uint32_t *val)
{
    RefResult *r = (RefResult*)h;
    std::memcpy(rowptr, r->a_rowptr.data(), r->a_rowptr.size() * 8); std::memcpy(col, r->a_col.data(), r->a_col.size() * 8);
    std::memcpy(val, r->a_val.data(), r->a_val.size() * 4);
}
void ref_get_B(void *h, int64_t *rowptr, int64_t *col, int32_t *num, uint32_t *seeds)
{
    RefResult *r = (RefResult*)h;
    std::memcpy(rowptr, r->b_rowptr.data(), r->b_rowptr.size() * 8); std::memcpy(col, r->b_col.data(), r->b_col.size() * 8);
    std::memcpy(num, r->b_num.data(), r->b_num.size() * 4); std::memcpy(seeds, r->b_seeds.data(), r->b_seeds.size() * 4);
}

/*
 * The consumer of B: Overlap(len, seed).extend_overlap(seqQ, seqT, ...) exactly as
 * src/PairwiseAlignment.cpp:82-91 drives it — xdrop_aligner + classify_alignment (src/XDropAligner.cpp) and the
 * direction / suffix fields (src/Overlap.cpp:20-73).  13 ints per pair, the order of oracle/xdrop_oracle.cpp.
 */
void ref_xdrop_batch(const uint8_t *packed, const uint64_t *off, const uint64_t *lens, const int64_t *rows, const int64_t *cols,
                     const uint32_t *seedq, const uint32_t *seedt, uint64_t npairs, int mat, int mis, int gap, int drop, int32_t *out)
{
    for (uint64_t p = 0; p < npairs; ++p)
    {
        const size_t lq = lens[rows[p]], lt = lens[cols[p]];
        DnaSeq seqQ(lq, const_cast<uint8_t*>(packed + off[rows[p]])), seqT(lt, const_cast<uint8_t*>(packed + off[cols[p]]));
        Overlap o(std::make_tuple((PosInRead)lq, (PosInRead)lt), std::make_tuple((PosInRead)seedq[p], (PosInRead)seedt[p]));
        o.extend_overlap(seqQ, seqT, mat, mis, gap, drop);
        int32_t *r = out + 13 * p;
        r[0] = (int32_t)std::get<0>(o.beg); r[1] = (int32_t)std::get<0>(o.end); r[2] = (int32_t)std::get<1>(o.beg); r[3] = (int32_t)std::get<1>(o.end);
        r[4] = o.score; r[5] = o.rc; r[6] = o.passed; r[7] = o.containedQ; r[8] = o.containedT;
        r[9] = o.direction; r[10] = o.directionT; r[11] = o.suffix; r[12] = o.suffixT;
    }
}

/*
 * The step before the hot path: FastaIndex(fasta, commgrid) + FastaIndex::getmydna()
 * (src/FastaIndex.cpp:98-176,191-290, compiled unmodified) on `nranks` thread-ranks.  Per rank: the .fai records it was
 * handed (len, pos, bases), and the DnaBuffer arena it parsed.
 */
struct RefFasta { std::vector<std::vector<uint64_t>> rec; std::vector<std::vector<uint8_t>> buf; std::vector<int64_t> displ; };

void *ref_fasta_run(const char *path, int nranks)
{
    RefFasta *res = new RefFasta; res->rec.resize(nranks); res->buf.resize(nranks); res->displ.assign(nranks + 1, 0);
    fake_mpi::World world(nranks);
    fake_mpi::g_world = &world;
    auto body = [&](int rank)
    {
        fake_mpi::t_rank = rank;
        auto commgrid = std::make_shared<CommGrid>(MPI_COMM_WORLD, 0, 0);
        FastaIndex index(path, commgrid);
        DnaBuffer mydna = index.getmydna();
        for (const auto &r : index.getmyrecords()) { res->rec[rank].push_back(r.len); res->rec[rank].push_back(r.pos); res->rec[rank].push_back(r.bases); }
        const size_t nb = mydna.getrangebufsize(0, mydna.size());
        res->buf[rank].assign(mydna.getbufoffset(0), mydna.getbufoffset(0) + nb);
        res->displ[rank] = (int64_t)index.getmyreaddispl();
        if (rank == 0) res->displ[nranks] = (int64_t)index.gettotrecords();
        MPI_Barrier(MPI_COMM_WORLD);
    };
    if (nranks == 1) body(0);
    else
    {
        std::vector<std::thread> th;
        for (int r = 0; r < nranks; ++r) th.emplace_back(body, r);
        for (auto& x : th) x.join();
    }
    fake_mpi::g_world = nullptr; fake_mpi::t_rank = 0;
    return res;
}
void ref_fasta_free(void *h) { delete (RefFasta*)h; }
/* displ[nranks + 1] = first global read id of every rank, then the total; bytes[nranks] = arena bytes of every rank */
void ref_fasta_sizes(void *h, int64_t *displ, int64_t *bytes)
{
    RefFasta *r = (RefFasta*)h;
    for (size_t i = 0; i < r->displ.size(); ++i) displ[i] = r->displ[i];
    for (size_t i = 0; i < r->buf.size(); ++i) bytes[i] = (int64_t)r->buf[i].size();
}
void ref_fasta_get(void *h, int rank, uint64_t *rec /* 3 per read */, uint8_t *packed)
{
    RefFasta *r = (RefFasta*)h;
    std::memcpy(rec, r->rec[rank].data(), r->rec[rank].size() * 8);
    std::memcpy(packed, r->buf[rank].data(), r->buf[rank].size());
}

/*
 * Transitive reduction of the overlap graph: TransitiveReduction(R), src/TransitiveReduction.cpp:3-92,
 * compiled unmodified.  R = nnz triples (row, col, Overlap) of which only direction, directionT, suffix, suffixT matter to the
 * algorithm (include/TransitiveReduction.hpp:19-110); the string graph S comes back as triples with the same four fields.
 */
struct RefTR { std::vector<int64_t> row, col; std::vector<int32_t> f; /* 4 per entry: direction, directionT, suffix, suffixT */ };

void *ref_transitive_reduction(int64_t n, uint64_t nnz, const int64_t *rows, const int64_t *cols, const int32_t *fields /* 4 per entry */)
{
    RefTR *res = new RefTR;
    auto commgrid = std::make_shared<CommGrid>(MPI_COMM_WORLD, 0, 0);
    std::vector<int64_t> r(rows, rows + nnz), c(cols, cols + nnz);
    std::vector<Overlap> v; v.reserve(nnz);
    for (uint64_t e = 0; e < nnz; ++e)
    {
        Overlap o;
        o.direction = (int8_t)fields[4 * e]; o.directionT = (int8_t)fields[4 * e + 1]; o.suffix = fields[4 * e + 2]; o.suffixT = fields[4 * e + 3];
        o.passed = true;
        v.push_back(o);
    }
    CT<int64_t>::PDistVec drows(r, commgrid), dcols(c, commgrid);
    CT<Overlap>::PDistVec dvals(v, commgrid);
    CT<Overlap>::PSpParMat R(n, n, drows, dcols, dvals, false);       /* as src/PairwiseAlignment.cpp:97-103 builds it */
    auto S = TransitiveReduction(R);
    const auto &st = *S->st;
    for (int64_t i = 0; i < st.m; ++i)
        for (int64_t p = st.rowptr[i]; p < st.rowptr[i + 1]; ++p)
        {
            const Overlap &o = st.val[p];
            res->row.push_back(i); res->col.push_back(st.col[p]);
            res->f.push_back(o.direction); res->f.push_back(o.directionT); res->f.push_back(o.suffix); res->f.push_back(o.suffixT);
        }
    return res;
}
uint64_t ref_tr_size(void *h) { return ((RefTR*)h)->row.size(); }
void ref_tr_get(void *h, int64_t *row, int64_t *col, int32_t *fields)
{
    RefTR *r = (RefTR*)h;
    std::memcpy(row, r->row.data(), 8 * r->row.size()); std::memcpy(col, r->col.data(), 8 * r->col.size()); std::memcpy(fields, r->f.data(), 4 * r->f.size());
}
void ref_tr_free(void *h) { delete (RefTR*)h; }

} // extern "C"
