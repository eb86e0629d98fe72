/*
 * elba_fe_shim.cpp — the reference's five driver functions for the overlap-detection front end,
 * re-implemented on top of the C ABI of libelba_fe.so (include/elba_fe.h).
 *
 * Build it INSTEAD OF src/KmerOps.cpp and src/SharedSeeds.cpp of PASSIONLab/ELBA (Makefile:36-55 OBJECTS) and link
 * -lelba_fe: src/main.cpp:191-282 then compiles and runs unchanged.  It includes the reference's own
 * headers and uses only their public accessors; it contains no CUDA and no algorithm — every result comes from the
 * library.  Signatures replaced (reference file:line):
 *
 *   get_kmer_count_map_keys     include/KmerOps.hpp:27-28     src/KmerOps.cpp:18-204
 *   get_kmer_count_map_values   include/KmerOps.hpp:30        src/KmerOps.cpp:206-350
 *   create_kmer_matrix          include/KmerOps.hpp:24-25     src/KmerOps.cpp:361-401
 *   GetKmerOwner                include/KmerOps.hpp:31        src/KmerOps.cpp:352-359
 *   create_seed_matrix          include/SharedSeeds.hpp:98-99 src/SharedSeeds.cpp:4-10
 *
 * and, for the step before the path, FastaIndex::getmydna (include/FastaIndex.hpp:31,
 * src/FastaIndex.cpp:191-290) as elba_fe_getmydna(index); with -DELBA_FE_SHIM_WRAP_GETMYDNA and the linker option
 * -Wl,--wrap=_ZNK10FastaIndex8getmydnaEv the reference's own call site (src/main.cpp:139) reaches it unchanged.
 *
 * State between the calls lives in one elba_fe_ctx per rank, found through the address of the object the driver
 * passes back in (the KmerCountMap, then A).  The driver's `kmermap.reset()` (main.cpp:266) and `A.reset()` (:284)
 * therefore need no change; the context is destroyed after create_seed_matrix.
 *
 * Errors: the reference asserts / MPI_Aborts; so does this (fe_check).
 * MPI ranks: one rank drives one GPU (device = local rank modulo device count).  With more than one rank the
 * k-mer exchange and the panel exchange run inside the library over NCCL (elba_fe_comm_*, bootstrapped here by
 * broadcasting the ncclUniqueId over MPI).
 */
#include "KmerOps.hpp"
#include "SharedSeeds.hpp"
#include "FastaIndex.hpp"
#ifdef ELBA_FE_SHIM_TR
#include "TransitiveReduction.hpp"    /* also replaces src/TransitiveReduction.cpp (first version: every rank reduces the whole graph) */
#endif
#ifdef ELBA_FE_SHIM_ALIGN
#include "PairwiseAlignment.hpp"      /* also replaces src/PairwiseAlignment.cpp (first version, one rank) */
#endif
#include "elba_fe.h"

#include <cstdio>
#include <cstdlib>
#include <map>
#include <mutex>
#include <vector>

/* The (rows, cols, vals) SpParMat constructor merges duplicate coordinates with maximum<NT> (x < y ? y : x) when
 * SumDuplicates is false (src/KmerOps.cpp:400 relies on it for A; include/Overlap.hpp:76-78 defines operator< so that
 * src/PairwiseAlignment.cpp:97-103 can build R the same way).  B has no duplicate coordinates, so the order is never
 * consulted; it only has to exist for the constructor to instantiate. */
static inline bool operator<(const SharedSeeds& a, const SharedSeeds& b) { return a.getnumshared() < b.getnumshared(); }

namespace
{

struct FeState
{
    elba_fe_ctx *ctx = nullptr;
    int64_t readoffset = 0, totreads = 0;
    std::shared_ptr<CommGrid> grid;
};

std::map<const void*, FeState> g_states;      /* keyed by the KmerCountMap, later re-keyed by A */
std::mutex g_mu;

void fe_check(int rc, elba_fe_ctx *ctx, const char *what, MPI_Comm comm)
{
    if (rc == 0) return;
    std::fprintf(stderr, "elba_fe: %s failed (%d): %s\n", what, rc, elba_fe_last_error(ctx));
    MPI_Abort(comm, rc);
    std::abort();
}

FeState take_state(const void *key)
{
    std::lock_guard<std::mutex> lk(g_mu);
    auto it = g_states.find(key);
    assert(it != g_states.end() && "elba_fe shim: object was not produced by this library's functions");
    FeState s = it->second;
    g_states.erase(it);
    return s;
}

FeState& peek_state(const void *key)
{
    std::lock_guard<std::mutex> lk(g_mu);
    auto it = g_states.find(key);
    assert(it != g_states.end() && "elba_fe shim: object was not produced by this library's functions");
    return it->second;
}

} // namespace

/* src/KmerOps.cpp:352-359, kept because include/KmerOps.hpp declares it (the library partitions by its own hash) */
int GetKmerOwner(const TKmer& kmer, int nprocs)
{
    uint64_t myhash = kmer.GetHash();
    double range = static_cast<double>(myhash) * static_cast<double>(nprocs);
    size_t owner = range / std::numeric_limits<uint64_t>::max();
    assert(owner >= 0 && owner < static_cast<int>(nprocs));
    return static_cast<int>(owner);
}

namespace
{

/* one context per rank: parameters from the compile-time macros, the rank's GPU, the NCCL communicator over the grid */
FeState make_state(std::shared_ptr<CommGrid> commgrid, bool with_comm = true)
{
    MPI_Comm comm = commgrid->GetWorld();
    int myrank = commgrid->GetRank(), nprocs = commgrid->GetSize();
    elba_fe_config cfg;
    elba_fe_default_config(&cfg);
    cfg.k = KMER_SIZE; cfg.lower = LOWER_KMER_FREQ; cfg.upper = UPPER_KMER_FREQ; cfg.stride = 1; cfg.seed_count = 2;
    if (const char *d = std::getenv("ELBA_FE_DEVICE")) cfg.device = std::atoi(d);
    else
    {
        /* one rank drives one GPU: rank inside its node, modulo the devices the node shows (any launcher) */
        MPI_Comm node; int local = 0;
        MPI_Comm_split_type(comm, MPI_COMM_TYPE_SHARED, myrank, MPI_INFO_NULL, &node);
        MPI_Comm_rank(node, &local);
        MPI_Comm_free(&node);
        int ndev = elba_fe_device_count();
        cfg.device = ndev > 0 ? local % ndev : 0;
    }
    FeState st; st.grid = commgrid;
    fe_check(elba_fe_create(&cfg, &st.ctx), nullptr, "elba_fe_create", comm);
    if (nprocs > 1 && with_comm)
    {
        /* one NCCL communicator over the grid's ranks; the id travels over MPI */
        elba_fe_comm_id id;
        if (myrank == 0) fe_check(elba_fe_comm_get_id(&id), st.ctx, "elba_fe_comm_get_id", comm);
        MPI_Bcast(&id, (int)sizeof id, MPI_BYTE, 0, comm);
        fe_check(elba_fe_comm_init(st.ctx, &id, myrank, nprocs), st.ctx, "elba_fe_comm_init", comm);
    }
    return st;
}

/* reads that elba_fe_getmydna left resident on the GPU, keyed by the arena of the DnaBuffer it returned */
std::map<const void*, FeState> g_ingested;

}

/*
 * FastaIndex::getmydna (src/FastaIndex.cpp:191-290) with the parse on the device: the rank reads its chunk of the file with
 * the same collective MPI-IO call and hands it, with its .fai records, to elba_fe_ingest_fasta; the packed arena comes back
 * into a DnaBuffer for the stages that read it on the host (src/main.cpp:150,289) and STAYS resident for the counting that
 * follows (get_kmer_count_map_keys below finds it by the arena's address: no second upload).
 * Binding: `DnaBuffer mydna = elba_fe_getmydna(index);` at src/main.cpp:139, or link with
 * -Wl,--wrap=_ZNK10FastaIndex8getmydnaEv.
 */
DnaBuffer elba_fe_getmydna(const FastaIndex& index)
{
    auto commgrid = index.getcommgrid();
    MPI_Comm comm = commgrid->GetWorld();
    FeState st = make_state(commgrid);
    const auto& myrecords = index.getmyrecords();
    static_assert(sizeof(FastaIndex::Record) == 3 * sizeof(uint64_t), "FastaIndex::Record is {size_t len, pos, bases}");

    MPI_Offset startpos = 0, endpos = 0, filesize = 0;
    MPI_File fh;
    MPI_File_open(comm, index.get_fasta_fname().c_str(), MPI_MODE_RDONLY, MPI_INFO_NULL, &fh);
    MPI_File_get_size(fh, &filesize);
    if (!myrecords.empty())
    {
        startpos = myrecords.front().pos;
        endpos = myrecords.back().pos + myrecords.back().len + (myrecords.back().len / myrecords.back().bases);
        if (endpos > filesize) endpos = filesize;
    }
    MPI_Offset readbufsize = endpos - startpos;
    std::unique_ptr<char[]> readbuf(new char[readbufsize > 0 ? readbufsize : 1]);
    MPI_FILE_READ_AT_ALL(fh, startpos, &readbuf[0], readbufsize, MPI_CHAR, MPI_STATUS_IGNORE);
    MPI_File_close(&fh);

    st.readoffset = (int64_t)index.getmyreaddispl(); st.totreads = (int64_t)index.gettotrecords();
    fe_check(elba_fe_ingest_fasta(st.ctx, &readbuf[0], (uint64_t)readbufsize, (uint64_t)startpos,
                                  reinterpret_cast<const uint64_t*>(myrecords.data()), myrecords.size(), st.readoffset),
             st.ctx, "elba_fe_ingest_fasta", comm);
    uint64_t numreads = 0, bufsize = 0;
    fe_check(elba_fe_reads_size(st.ctx, &numreads, &bufsize), st.ctx, "elba_fe_reads_size", comm);
    uint8_t *buf = new uint8_t[bufsize ? bufsize : 1];           /* owned by the DnaBuffer (delete[] in its destructor) */
    fe_check(elba_fe_get_reads(st.ctx, buf, nullptr, nullptr), st.ctx, "elba_fe_get_reads", comm);
    auto readlens = index.getmyreadlens();
    {
        std::lock_guard<std::mutex> lk(g_mu);
        g_ingested[buf] = st;
    }
    return DnaBuffer(bufsize, numreads, buf, readlens.data());
}

#ifdef ELBA_FE_SHIM_WRAP_GETMYDNA
/* ld --wrap: every reference to FastaIndex::getmydna() const in the other objects resolves here (same calling convention:
 * hidden result pointer, then `this`) */
extern "C" DnaBuffer __wrap__ZNK10FastaIndex8getmydnaEv(const FastaIndex *self) { return elba_fe_getmydna(*self); }
#endif

std::unique_ptr<KmerCountMap>
get_kmer_count_map_keys(const DnaBuffer& myreads, std::shared_ptr<CommGrid> commgrid)
{
    static_assert(KMER_SIZE <= ELBA_FE_MAX_KMER_SIZE, "libelba_fe counts k-mers of one 64-bit word (TKmer = Kmer<1>)");
    MPI_Comm comm = commgrid->GetWorld();
    int myrank = commgrid->GetRank();
    size_t numreads = myreads.size();
    const uint8_t *base = numreads ? myreads.getbufoffset(0) : nullptr;

    FeState st; bool resident = false;
    {
        std::lock_guard<std::mutex> lk(g_mu);
        auto it = g_ingested.find(base);
        if (base && it != g_ingested.end()) { st = it->second; g_ingested.erase(it); resident = true; }
    }
    if (!resident) st = make_state(commgrid);

    /* global read ids: the MPI_Exscan of src/KmerOps.cpp:215-216 */
    int64_t mine = (int64_t)numreads, before = 0, total = mine;
    MPI_Exscan(&mine, &before, 1, MPI_INT64_T, MPI_SUM, comm);
    if (myrank == 0) before = 0;
    MPI_Allreduce(MPI_IN_PLACE, &total, 1, MPI_INT64_T, MPI_SUM, comm);
    st.readoffset = before; st.totreads = total;

    if (!resident)
    {
        /* the arena through DnaBuffer's public accessors only (include/DnaBuffer.hpp:18-23) */
        std::vector<uint64_t> off(numreads), len(numreads);
        for (size_t i = 0; i < numreads; ++i)
        {
            off[i] = (uint64_t)(myreads.getbufoffset(i) - base);
            len[i] = (uint64_t)myreads[i].size();
        }
        uint64_t nbytes = numreads ? off[numreads - 1] + (uint64_t)myreads[numreads - 1].numbytes() : 0;
        fe_check(elba_fe_upload_reads(st.ctx, base, nbytes, off.data(), len.data(), numreads, before), st.ctx, "elba_fe_upload_reads", comm);
    }

    auto kmermap = std::make_unique<KmerCountMap>();
    {
        std::lock_guard<std::mutex> lk(g_mu);
        g_states[kmermap.get()] = st;
    }
    return kmermap;
}

void get_kmer_count_map_values(const DnaBuffer& myreads, KmerCountMap& kmermap, std::shared_ptr<CommGrid> commgrid)
{
    (void)myreads;
    FeState& st = peek_state(&kmermap);
    MPI_Comm comm = commgrid->GetWorld();
    fe_check(elba_fe_count(st.ctx), st.ctx, "elba_fe_count", comm);
    fe_check(elba_fe_build_A(st.ctx), st.ctx, "elba_fe_build_A", comm);

    /* Fill the map the driver holds: its size and the per-k-mer counts are logged (src/main.cpp:449-485) and it is what
     * create_kmer_matrix receives.  Entries: this rank's share of the reliable k-mers, their exact instance counts, and
     * the (read, position) pairs of A's column (one per read: duplicates inside a read were already merged to the
     * largest position, the rule the SpParMat constructor applies in the reference, src/KmerOps.cpp:400). */
    elba_fe_sizes_t sz;
    fe_check(elba_fe_sizes(st.ctx, &sz), st.ctx, "elba_fe_sizes", comm);
    std::vector<uint64_t> kmers(sz.reliable); std::vector<uint32_t> counts(sz.reliable);
    fe_check(elba_fe_get_kmers(st.ctx, kmers.data(), counts.data()), st.ctx, "elba_fe_get_kmers", comm);
    std::vector<int64_t> colptr(sz.reliable + 1); std::vector<uint32_t> rows(sz.nnzA), pos(sz.nnzA);
    fe_check(elba_fe_get_AT(st.ctx, colptr.data(), rows.data(), pos.data()), st.ctx, "elba_fe_get_AT", comm);
    int myrank = commgrid->GetRank(), nprocs = commgrid->GetSize();
    kmermap.reserve(sz.reliable / nprocs + 1);
    for (uint64_t c = 0; c < sz.reliable; ++c)
    {
        if ((int)(c % (uint64_t)nprocs) != myrank) continue;      /* the reliable list is replicated; each rank shows its slice */
        TKmer kmer((const void*)&kmers[c]);
        KmerCountEntry& e = kmermap[kmer];
        READIDS& readids = std::get<0>(e); POSITIONS& positions = std::get<1>(e);
        int n = 0;
        for (int64_t q = colptr[c]; q < colptr[c + 1] && n < UPPER_KMER_FREQ; ++q, ++n) { readids[n] = (ReadId)rows[q] + st.readoffset; positions[n] = pos[q]; }
        std::get<2>(e) = (int)counts[c];
    }
}

std::unique_ptr<CT<PosInRead>::PSpParMat>
create_kmer_matrix(const DnaBuffer& myreads, const KmerCountMap& kmermap, std::shared_ptr<CommGrid> commgrid)
{
    (void)myreads;
    FeState st = take_state(&kmermap);
    MPI_Comm comm = commgrid->GetWorld();
    elba_fe_sizes_t sz;
    fe_check(elba_fe_sizes(st.ctx, &sz), st.ctx, "elba_fe_sizes", comm);

    /* this rank's rows of A as global triples, handed to the same constructor the reference uses (KmerOps.cpp:396-400) */
    std::vector<int64_t> rowptr(sz.nreads + 1); std::vector<uint32_t> col(sz.nnzA), pos(sz.nnzA);
    fe_check(elba_fe_get_A(st.ctx, rowptr.data(), col.data(), pos.data()), st.ctx, "elba_fe_get_A", comm);
    std::vector<int64_t> local_rowids(sz.nnzA), local_colids(sz.nnzA);
    std::vector<PosInRead> local_positions(pos.begin(), pos.end());
    for (uint64_t r = 0; r < sz.nreads; ++r)
        for (int64_t q = rowptr[r]; q < rowptr[r + 1]; ++q) { local_rowids[q] = (int64_t)r + st.readoffset; local_colids[q] = (int64_t)col[q]; }

    CT<int64_t>::PDistVec drows(local_rowids, commgrid);
    CT<int64_t>::PDistVec dcols(local_colids, commgrid);
    CT<PosInRead>::PDistVec dvals(local_positions, commgrid);
    auto A = std::make_unique<CT<PosInRead>::PSpParMat>(st.totreads, (int64_t)sz.reliable, drows, dcols, dvals, false);
    {
        std::lock_guard<std::mutex> lk(g_mu);
        g_states[A.get()] = st;
    }
    return A;
}

std::unique_ptr<CT<SharedSeeds>::PSpParMat>
create_seed_matrix(CT<PosInRead>::PSpParMat& A, CT<PosInRead>::PSpParMat& AT)
{
    (void)AT;                                   /* the library built A and its transpose together (elba_fe_build_A) */
    FeState st = take_state(&A);
    MPI_Comm comm = st.grid->GetWorld();
    fe_check(elba_fe_spgemm(st.ctx), st.ctx, "elba_fe_spgemm", comm);
    elba_fe_sizes_t sz;
    fe_check(elba_fe_sizes(st.ctx, &sz), st.ctx, "elba_fe_sizes", comm);

    std::vector<int64_t> rows(sz.nnzB), cols(sz.nnzB); std::vector<int32_t> num(sz.nnzB); std::vector<uint32_t> seeds(4 * sz.nnzB);
    fe_check(elba_fe_get_B_triples(st.ctx, rows.data(), cols.data(), num.data(), seeds.data()), st.ctx, "elba_fe_get_B_triples", comm);
#ifndef ELBA_FE_SHIM_ALIGN
    elba_fe_destroy(st.ctx);
#endif

    /* field-wise: std::tuple's memory order is not its template order */
    std::vector<SharedSeeds> vals; vals.reserve(sz.nnzB);
    for (uint64_t e = 0; e < sz.nnzB; ++e)
        vals.emplace_back(std::make_tuple((PosInRead)seeds[4 * e], (PosInRead)seeds[4 * e + 1]),
                          std::make_tuple((PosInRead)seeds[4 * e + 2], (PosInRead)seeds[4 * e + 3]), (int)num[e]);

#ifdef ELBA_FE_SHIM_SPTUPLES
    /* real CombBLAS: SharedSeeds has no operator<, so the (rows, cols, vals) constructor (which merges duplicates with
     * maximum<NT>) does not instantiate; build the local block from tuples exactly as CombBLAS returns SpGEMM results.
     * The triples are this rank's block in local indices. */
    int64_t nloc_r = 0, nloc_c = 0, roff = 0, coff = 0;
    elba_fe_block_extent(st.totreads, st.grid->GetGridRows(), st.grid->GetRankInProcCol(), &roff, &nloc_r);
    elba_fe_block_extent(st.totreads, st.grid->GetGridCols(), st.grid->GetRankInProcRow(), &coff, &nloc_c);
    auto *tuples = new std::tuple<int64_t, int64_t, SharedSeeds>[sz.nnzB];
    for (uint64_t e = 0; e < sz.nnzB; ++e) tuples[e] = std::make_tuple(rows[e] - roff, cols[e] - coff, vals[e]);
    combblas::SpTuples<int64_t, SharedSeeds> spt((int64_t)sz.nnzB, nloc_r, nloc_c, tuples, false);
    auto *dcsc = new CT<SharedSeeds>::PSpDCCols(spt, false);
    auto B = std::make_unique<CT<SharedSeeds>::PSpParMat>(dcsc, st.grid);
#else
    CT<int64_t>::PDistVec drows(rows, st.grid);
    CT<int64_t>::PDistVec dcols(cols, st.grid);
    CT<SharedSeeds>::PDistVec dvals(vals, st.grid);
    auto B = std::make_unique<CT<SharedSeeds>::PSpParMat>(st.totreads, st.totreads, drows, dcols, dvals, false);
#endif
#ifdef ELBA_FE_SHIM_ALIGN
    {
        std::lock_guard<std::mutex> lk(g_mu);
        g_states[B.get()] = st;                 /* B is still on the device: PairwiseAlignment below aligns it there */
    }
#endif
    return B;
}

#ifdef ELBA_FE_SHIM_ALIGN
/*
 * PairwiseAlignment (include/PairwiseAlignment.hpp:9-10, src/PairwiseAlignment.cpp:5-106) on the device: the nonzeros of
 * the upper triangle of B, each extended from seeds[0] by the X-drop aligner (elba_fe_align).  The reads and B never left
 * the GPU, so dfd's row / column buffers are not touched.  The result goes through the same triple constructor as in
 * the reference (:97-103); the Overlap values are filled field-wise from the 13 ints per pair.  One rank in this version.
 */
std::unique_ptr<CT<Overlap>::PSpParMat>
PairwiseAlignment(DistributedFastaData& dfd, CT<SharedSeeds>::PSpParMat& Bmat, int mat, int mis, int gap, int dropoff)
{
    FeState st = take_state(&Bmat);
    MPI_Comm comm = st.grid->GetWorld();
    uint64_t np = 0;
    fe_check(elba_fe_align(st.ctx, mat, mis, gap, dropoff, &np), st.ctx, "elba_fe_align", comm);
    std::vector<int64_t> rows(np), cols(np); std::vector<int32_t> f((size_t)ELBA_FE_ALIGN_FIELDS * np);
    fe_check(elba_fe_get_alignments(st.ctx, rows.data(), cols.data(), f.data()), st.ctx, "elba_fe_get_alignments", comm);
    /* read lengths of the aligned pairs: from the library's copy of the reads (dfd keeps the same lengths in its index) */
    elba_fe_destroy(st.ctx);

    auto& index = dfd.getindex();
    const auto& records = index.getmyrecords();     /* one rank: every read is local */
    std::vector<Overlap> overlaps; overlaps.reserve(np);
    for (uint64_t p = 0; p < np; ++p)
    {
        const int32_t *r = &f[(size_t)ELBA_FE_ALIGN_FIELDS * p];
        std::tuple<PosInRead, PosInRead> len((PosInRead)records[rows[p]].len, (PosInRead)records[cols[p]].len);
        Overlap o(len, std::make_tuple((PosInRead)0, (PosInRead)0));
        std::get<0>(o.beg) = (PosInRead)r[0]; std::get<0>(o.end) = (PosInRead)r[1]; std::get<1>(o.beg) = (PosInRead)r[2]; std::get<1>(o.end) = (PosInRead)r[3];
        o.score = r[4]; o.rc = r[5] != 0; o.passed = r[6] != 0; o.containedQ = r[7] != 0; o.containedT = r[8] != 0;
        o.direction = (int8_t)r[9]; o.directionT = (int8_t)r[10]; o.suffix = r[11]; o.suffixT = r[12];
        overlaps.push_back(o);
    }
    CT<int64_t>::PDistVec drows(rows, st.grid);
    CT<int64_t>::PDistVec dcols(cols, st.grid);
    CT<Overlap>::PDistVec dvals(overlaps, st.grid);
    return std::make_unique<CT<Overlap>::PSpParMat>(st.totreads, st.totreads, drows, dcols, dvals, false);
}
#endif

#ifdef ELBA_FE_SHIM_TR
/*
 * TransitiveReduction (include/TransitiveReduction.hpp:17, src/TransitiveReduction.cpp:3-92) on the device
 * (elba_fe_transitive_reduction).  The overlap matrix is small next to everything before it, so every rank gathers the
 * triples of all blocks (walked exactly as src/PairwiseAlignment.cpp:16-33 walks B), reduces the whole graph on its GPU
 * (replicas, no device collective) and hands the constructor the entries of S that stem from ITS OWN triples; the
 * (rows, cols, vals) constructor redistributes them (src/PairwiseAlignment.cpp:97-103 builds R the same way).  The
 * Overlap payload of a mirror-image entry is Overlap::Transpose of the source entry (include/Overlap.hpp:46-74).
 */
Overlap opmin(const Overlap& e1, const Overlap& e2)       /* src/TransitiveReduction.cpp:94-102: declared in the header, kept for other users */
{
    Overlap e = Overlap();
    for (int i = 0; i < 4; ++i) e.suffix_paths[i] = std::min(e1.suffix_paths[i], e2.suffix_paths[i]);
    return e;
}

std::unique_ptr<CT<Overlap>::PSpParMat> TransitiveReduction(CT<Overlap>::PSpParMat R)
{
    auto commgrid = R.getcommgrid();
    MPI_Comm comm = commgrid->GetWorld();
    int myrank = commgrid->GetRank(), nprocs = commgrid->GetSize();
    const int64_t nrow = R.getnrow(), ncol = R.getncol();

    int64_t roff = 0, nr = 0, coff = 0, nc = 0;
    elba_fe_block_extent(nrow, commgrid->GetGridRows(), commgrid->GetRankInProcCol(), &roff, &nr);
    elba_fe_block_extent(ncol, commgrid->GetGridCols(), commgrid->GetRankInProcRow(), &coff, &nc);
    std::vector<int64_t> rows, cols; std::vector<int32_t> f; std::vector<const Overlap*> mine;
    auto dcsc = R.seqptr()->GetDCSC();
    if (dcsc != nullptr)
        for (int64_t i = 0; i < dcsc->nzc; ++i)
            for (int64_t j = dcsc->cp[i]; j < dcsc->cp[i+1]; ++j)
            {
                const Overlap& o = dcsc->numx[j];
                rows.push_back(dcsc->ir[j] + roff); cols.push_back(dcsc->jc[i] + coff);
                f.push_back(o.direction); f.push_back(o.directionT); f.push_back(o.suffix); f.push_back(o.suffixT);
                mine.push_back(&o);
            }

    /* the whole graph on every rank */
    std::vector<int> cnt(nprocs, 0), dis(nprocs + 1, 0);
    int mycnt = (int)rows.size();
    std::vector<int64_t> arows, acols; std::vector<int32_t> af;
    if (nprocs > 1)
    {
        MPI_Allgather(&mycnt, 1, MPI_INT, cnt.data(), 1, MPI_INT, comm);
        for (int r = 0; r < nprocs; ++r) dis[r + 1] = dis[r] + cnt[r];
        arows.resize(dis[nprocs]); acols.resize(dis[nprocs]); af.resize(4 * (size_t)dis[nprocs]);
        std::vector<int> cnt4(nprocs), dis4(nprocs);
        for (int r = 0; r < nprocs; ++r) { cnt4[r] = 4 * cnt[r]; dis4[r] = 4 * dis[r]; }
        MPI_Allgatherv(rows.data(), mycnt, MPI_INT64_T, arows.data(), cnt.data(), dis.data(), MPI_INT64_T, comm);
        MPI_Allgatherv(cols.data(), mycnt, MPI_INT64_T, acols.data(), cnt.data(), dis.data(), MPI_INT64_T, comm);
        MPI_Allgatherv(f.data(), 4 * mycnt, MPI_INT, af.data(), cnt4.data(), dis4.data(), MPI_INT, comm);
    }
    else { cnt[0] = mycnt; dis[1] = mycnt; arows = rows; acols = cols; af = f; }

    FeState st = make_state(commgrid, false);
    uint64_t ns = 0;
    fe_check(elba_fe_transitive_reduction(st.ctx, arows.data(), acols.data(), af.data(), arows.size(), nrow, FUZZ, &ns), st.ctx, "elba_fe_transitive_reduction", comm);
    std::vector<int64_t> srow(ns), scol(ns); std::vector<int32_t> sf(4 * ns); std::vector<uint64_t> ssrc(ns); std::vector<uint8_t> str(ns);
    fe_check(elba_fe_get_string_graph(st.ctx, srow.data(), scol.data(), sf.data(), ssrc.data(), str.data()), st.ctx, "elba_fe_get_string_graph", comm);
    elba_fe_destroy(st.ctx);

    std::vector<int64_t> orow, ocol; std::vector<Overlap> oval;
    const uint64_t lo = (uint64_t)dis[myrank], hi = (uint64_t)dis[myrank + 1];
    for (uint64_t e = 0; e < ns; ++e)
    {
        if (ssrc[e] < lo || ssrc[e] >= hi) continue;          /* another rank holds that entry's payload */
        const Overlap& o = *mine[ssrc[e] - lo];
        orow.push_back(srow[e]); ocol.push_back(scol[e]);
        oval.push_back(str[e] ? Overlap::Transpose()(o) : o);
    }
    CT<int64_t>::PDistVec drows(orow, commgrid);
    CT<int64_t>::PDistVec dcols(ocol, commgrid);
    CT<Overlap>::PDistVec dvals(oval, commgrid);
    return std::make_unique<CT<Overlap>::PSpParMat>(nrow, ncol, drows, dcols, dvals, false);
}
#endif
