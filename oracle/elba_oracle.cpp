/*
 * TEST INFRASTRUCTURE ONLY (oracle/).  Nothing under elba_b200/ may include,
 * link or execute this; only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs load the library built from it.
 *
 * CPU restatement of ELBA's overlap-detection front end
 *   get_kmer_count_map_keys  -> get_kmer_count_map_values -> create_kmer_matrix
 *   -> Transpose -> create_seed_matrix        (/root/reference/src/main.cpp:191-282)
 * written from the reference's observable behaviour, not from its code.
 * Every function cites the reference lines it follows.
 *
 * PARITY PINNING: this file is checked (tests/test_oracle_vs_reference.py, run
 * in the build container where /root/reference exists, and through the
 * committed fixtures in tests/golden/ everywhere else) against
 *   (1) the golden k-mer / hash / Bloom / HyperLogLog vectors of SURVEY.md §8c,
 *   (2) oracle/_ref = the reference's own KmerOps.cpp, SharedSeeds.cpp, Kmer,
 *       HashFuncs, Bloom, HyperLogLog, DnaSeq, DnaBuffer compiled unmodified
 *       (oracle/ref_wrap.cpp) and run on reads.fa / example_medium.
 * Tier 1 (k-mer stream, reliable set + counts, (read,pos) sets, A after
 * max-pos dedupe, pattern(B), numshared, prune) is pinned bit-exact.
 * Tier 2 (WHICH two seeds a nonzero keeps) is "parity unpinned": it is decided
 * inside CombBLAS, which is an unpinned external dependency absent from
 * /root/reference.  The canonical rule used here and by the CUDA path:
 *   column id of a reliable k-mer = rank of its 64-bit value (ascending);
 *   products of one output nonzero are folded in ascending column id with
 *   SharedSeeds::Semiring::add  =>  seeds[0] = pair of the smallest shared
 *   column, seeds[1] = pair of the largest.
 */
#include <cstdint>
#include <cstring>
#include <cstdlib>
#include <cmath>
#include <vector>
#include <algorithm>
#include <numeric>
#include <limits>
#include <unordered_map>
#include <chrono>
#include <thread>

namespace {

/* ---- 2-bit reads: src/DnaSeq.cpp:48-54 (base i at bits 6-2*(i%4) of byte i/4) ---- */
inline unsigned base_at(const uint8_t *mem, uint64_t i) { return (mem[i >> 2] >> (6 - 2 * (i & 3))) & 3u; }

/* include/DnaSeq.hpp:136-154: ACGT/acgt -> 0..3, N/n -> 0, everything else undefined (4) */
inline int char_code(char c)
{
    switch (c) { case 'A': case 'a': case 'N': case 'n': return 0; case 'C': case 'c': return 1;
                 case 'G': case 'g': return 2; case 'T': case 't': return 3; default: return 4; }
}

/* ---- k-mer value: include/Kmer.hpp:95-97 + src/Kmer.cpp:66-84: base j at bits 2*(31-j), low bits zero ---- */
/* reverse complement, left-aligned: what GetTwin (src/Kmer.cpp:167-198) computes with its tetramer table */
inline uint64_t twin_of(uint64_t x, int k)
{
    uint64_t y = ~x;                                                          /* complement every 2-bit code */
    y = ((y >> 2) & 0x3333333333333333ULL) | ((y & 0x3333333333333333ULL) << 2);   /* reverse the 32 codes */
    y = ((y >> 4) & 0x0F0F0F0F0F0F0F0FULL) | ((y & 0x0F0F0F0F0F0F0F0FULL) << 4);
    y = __builtin_bswap64(y);
    return y << (2 * (32 - k));                                               /* drop the complemented padding */
}
inline uint64_t rep_of(uint64_t x, int k) { uint64_t t = twin_of(x, k); return t < x ? t : x; }   /* src/Kmer.cpp:200-205 */

/* ---- MurmurHash3 x64_128 (public algorithm, A. Appleby) as src/HashFuncs.cpp:40-117 instantiates it ---- */
inline uint64_t rotl(uint64_t v, int r) { return (v << r) | (v >> (64 - r)); }
inline uint64_t fmix(uint64_t v)
{
    v ^= v >> 33; v *= 0xff51afd7ed558ccdULL; v ^= v >> 33; v *= 0xc4ceb9fe1a85ec53ULL; v ^= v >> 33; return v;
}
void murmur128(const uint8_t *data, uint32_t len, uint32_t seed, uint64_t out[2])
{
    const uint64_t C1 = 0x87c37b91114253d5ULL, C2 = 0x4cf5ad432745937fULL;
    uint64_t h1 = seed, h2 = seed;
    uint32_t nb = len / 16;
    for (uint32_t i = 0; i < nb; ++i)
    {
        uint64_t a, b; std::memcpy(&a, data + 16 * i, 8); std::memcpy(&b, data + 16 * i + 8, 8);
        a *= C1; a = rotl(a, 31); a *= C2; h1 ^= a; h1 = rotl(h1, 27) + h2; h1 = h1 * 5 + 0x52dce729;
        b *= C2; b = rotl(b, 33); b *= C1; h2 ^= b; h2 = rotl(h2, 31) + h1; h2 = h2 * 5 + 0x38495ab5;
    }
    const uint8_t *tail = data + 16 * nb;
    uint32_t rem = len & 15;
    uint64_t a = 0, b = 0;
    for (uint32_t i = rem; i > 8; --i) b ^= (uint64_t)tail[i - 1] << (8 * (i - 9));
    if (rem > 8) { b *= C2; b = rotl(b, 33); b *= C1; h2 ^= b; }
    for (uint32_t i = (rem < 8 ? rem : 8); i > 0; --i) a ^= (uint64_t)tail[i - 1] << (8 * (i - 1));
    if (rem > 0) { a *= C1; a = rotl(a, 31); a *= C2; h1 ^= a; }
    h1 ^= len; h2 ^= len; h1 += h2; h2 += h1; h1 = fmix(h1); h2 = fmix(h2); h1 += h2; h2 += h1;
    out[0] = h1; out[1] = h2;
}
/* Kmer::GetHash, src/Kmer.cpp:207-213 -> murmurhash3_64, src/HashFuncs.cpp:231-236: seed 313, first word */
inline uint64_t kmer_hash(uint64_t x) { uint64_t o[2]; murmur128((const uint8_t*)&x, 8, 313, o); return o[0]; }
/* murmurhash3(key,len,seed): low 32 bits of the first word, src/HashFuncs.cpp:245-250 */
inline uint32_t murmur32(const void *p, uint32_t len, uint32_t seed) { uint64_t o[2]; murmur128((const uint8_t*)p, len, seed, o); return (uint32_t)o[0]; }

struct Reads { const uint8_t *buf; const uint64_t *off; const uint64_t *len; uint64_t n; };

/* ForeachKmer (include/KmerOps.hpp:106-136) + GetRepKmers (src/Kmer.cpp:215-242): all len-k+1 windows, pos = window start */
template <class F> void foreach_kmer(const uint8_t *mem, uint64_t len, int k, F f)
{
    if (len < (uint64_t)k) return;
    uint64_t cur = 0;
    for (int j = 0; j < k; ++j) cur |= (uint64_t)base_at(mem, j) << (2 * (31 - j));            /* set_kmer, src/Kmer.cpp:66-84 */
    f(rep_of(cur, k), (uint64_t)0);
    for (uint64_t p = 1; p + k <= len; ++p)
    {
        cur = (cur << 2) | ((uint64_t)base_at(mem, p + k - 1) << (2 * (32 - k)));              /* GetExtension, src/Kmer.cpp:149-165 */
        f(rep_of(cur, k), p);
    }
}

/* ---- HyperLogLog, src/HyperLogLog.cpp ---- */
struct Hll
{
    static const int BITS = 12, SIZE = 1 << 12;
    uint8_t reg[SIZE];
    Hll() { std::memset(reg, 0, sizeof reg); }
    /* add(), :40-50 — hashes the k-byte ASCII string of the canonical k-mer (include/KmerOps.hpp:64-68) */
    void add_kmer(uint64_t x, int k)
    {
        char s[32];
        for (int i = 0; i < k; ++i) s[i] = "ACGT"[(x >> (2 * (31 - i))) & 3];
        uint64_t o[2]; murmur128((const uint8_t*)s, (uint32_t)k, 313, o);
        uint64_t h = o[0];
        uint32_t idx = (uint32_t)(h >> (64 - BITS));
        uint64_t w = h << BITS;                         /* rho(), :12-23: leading-zero run of the remaining 52 bits, capped */
        uint8_t v = 1;
        while (v <= 64 - BITS && !(w & (1ULL << 63))) { v++; w <<= 1; }
        if (v > reg[idx]) reg[idx] = v;
    }
    /* estimate(), :52-76 */
    double estimate() const
    {
        double alpha_mm = (0.7213 / (1.0 + 1.079 / SIZE)) * SIZE * SIZE;
        double sum = 0.0;
        for (int i = 0; i < SIZE; ++i) sum += 1.0 / (double)(1 << reg[i]);
        double est = alpha_mm / sum;
        if (est <= 2.5 * SIZE)
        {
            uint32_t zeros = 0;
            for (int i = 0; i < SIZE; ++i) zeros += (reg[i] == 0);
            if (zeros) est = SIZE * std::log((double)SIZE / zeros);
        }
        return est;
    }
};

/* ---- Bloom, src/Bloom.cpp ---- */
struct Bloom
{
    int64_t bits, bytes; int hashes; std::vector<uint8_t> bf;
    Bloom(int64_t entries, double err)       /* :6-27 */
    {
        double bpe = -(std::log(err) / 0.480453013918201);
        bits = (int64_t)((double)entries * bpe);
        bytes = bits / 8 + !!(bits % 8);
        hashes = (int)std::ceil(0.693147180559945 * bpe);
        bf.assign(bytes, 0);
    }
    static void ab(uint64_t x, uint64_t &a, uint64_t &b)   /* :44-73: four chained murmurs over the 8 raw k-mer bytes */
    {
        uint32_t a1 = murmur32(&x, 8, 0x9747b28c), a2 = murmur32(&x, 8, a1), b1 = murmur32(&x, 8, a2), b2 = murmur32(&x, 8, b1);
        a = ((uint64_t)a1 << 32) | a2; b = ((uint64_t)b1 << 32) | b2;
    }
    bool check_add(uint64_t x, bool add)
    {
        uint64_t a, b; ab(x, a, b);
        int hits = 0;
        for (uint32_t i = 0; i < (uint32_t)hashes; ++i)
        {
            uint64_t p = (a + i * b) % (uint64_t)bits;
            uint8_t m = (uint8_t)(1u << (p % 8));
            if (bf[p >> 3] & m) hits++; else if (add) bf[p >> 3] |= m;
        }
        return hits == hashes;
    }
};

struct Result
{
    uint64_t N = 0, M = 0, D = 0, R = 0, nnzA_pre = 0, F = 0, nnzB_pre = 0;
    std::vector<uint64_t> rel_kmer;        /* ascending == column id order */
    std::vector<uint32_t> rel_count;       /* instance counts in [L,U] */
    std::vector<int64_t> a_rowptr; std::vector<uint32_t> a_col, a_pos;      /* A, CSR, cols ascending, max pos kept */
    std::vector<int64_t> at_colptr; std::vector<uint32_t> at_row, at_pos;   /* same entries by column */
    std::vector<int64_t> b_rowptr; std::vector<uint32_t> b_col; std::vector<int32_t> b_num; std::vector<uint32_t> b_seeds;
    double secs[4] = {0,0,0,0};            /* count, build A, spgemm, total */
};

int g_threads = 1;
int g_stride = 1;      /* -s of the legacy command line (README.md:85): only window starts p with p % stride == 0 are visited; the reference hard-wires 1 */
template <class F> void foreach_kmer_s(const uint8_t *mem, uint64_t len, int k, F f)
{
    if (g_stride <= 1) { foreach_kmer(mem, len, k, f); return; }
    foreach_kmer(mem, len, k, [&](uint64_t x, uint64_t p) { if (p % (uint64_t)g_stride == 0) f(x, p); });
}
double now() { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); }

/*
 * Counting.  With LOWER >= 2 the reference's Bloom-gated two-pass scheme
 * (src/KmerOps.cpp:156-187 pass 1, :283-340 pass 2) yields exactly
 * { x : LOWER <= count(x) <= UPPER } with exact instance counts (SURVEY.md §8a;
 * verified against oracle/_ref).  Restated here as sort + run-length.
 */
void count_kmers(const Reads &rd, int k, int lower, int upper, Result &res)
{
    const uint64_t st = g_stride < 1 ? 1 : (uint64_t)g_stride;
    uint64_t M = 0;
    for (uint64_t r = 0; r < rd.n; ++r) if (rd.len[r] >= (uint64_t)k) M += (rd.len[r] - k + 1 + st - 1) / st;
    res.M = M; res.D = 0;
    /* value ranges by the top byte: sorted range after sorted range is the ascending order; ranges sort in parallel */
    const int NRANGE = 256;
    std::vector<std::vector<uint64_t>> part(NRANGE);
    for (auto &v : part) v.reserve(M / NRANGE + M / (4 * NRANGE) + 16);
    for (uint64_t r = 0; r < rd.n; ++r)
        foreach_kmer_s(rd.buf + rd.off[r], rd.len[r], k, [&](uint64_t x, uint64_t) { part[x >> 56].push_back(x); });
    int nt = g_threads < 1 ? 1 : g_threads;
    auto body = [&](int t) { for (int q = t; q < NRANGE; q += nt) std::sort(part[q].begin(), part[q].end()); };
    if (nt == 1) body(0);
    else { std::vector<std::thread> th; for (int t = 0; t < nt; ++t) th.emplace_back(body, t); for (auto &x : th) x.join(); }
    for (int q = 0; q < NRANGE; ++q)
    {
        const std::vector<uint64_t> &all = part[q];
        for (size_t i = 0; i < all.size(); )
        {
            size_t j = i + 1; while (j < all.size() && all[j] == all[i]) ++j;
            res.D++;
            if (j - i >= (size_t)lower && j - i <= (size_t)upper) { res.rel_kmer.push_back(all[i]); res.rel_count.push_back((uint32_t)(j - i)); }
            i = j;
        }
        std::vector<uint64_t>().swap(part[q]);
    }
    res.R = res.rel_kmer.size();
}

/*
 * A: one (read, column, pos) per instance of a reliable k-mer (src/KmerOps.cpp:377-394);
 * duplicates of (read,column) are merged by the CombBLAS constructor with
 * SumDuplicates=false -> maximum -> the largest position survives (:396-400).
 */
void build_A(const Reads &rd, int k, Result &res)
{
    std::unordered_map<uint64_t, uint32_t> id; id.reserve(res.R * 2);
    for (uint32_t c = 0; c < res.R; ++c) id.emplace(res.rel_kmer[c], c);
    res.a_rowptr.assign(rd.n + 1, 0);
    int nt = g_threads < 1 ? 1 : g_threads;
    struct Part { std::vector<uint32_t> col, pos; uint64_t pre = 0; };
    std::vector<Part> parts(nt);
    std::vector<int64_t> rownnz(rd.n, 0);
    auto body = [&](int t)
    {
        Part &pt = parts[t];
        std::vector<std::pair<uint32_t, uint32_t>> row;
        for (uint64_t r = rd.n * t / nt; r < rd.n * (t + 1) / nt; ++r)
        {
            row.clear();
            foreach_kmer_s(rd.buf + rd.off[r], rd.len[r], k, [&](uint64_t x, uint64_t p) {
                auto it = id.find(x); if (it != id.end()) row.emplace_back(it->second, (uint32_t)p); });
            pt.pre += row.size();
            std::sort(row.begin(), row.end());
            for (size_t i = 0; i < row.size(); ++i)
                if (i + 1 == row.size() || row[i + 1].first != row[i].first) { pt.col.push_back(row[i].first); pt.pos.push_back(row[i].second); rownnz[r]++; }
        }
    };
    if (nt == 1) body(0);
    else { std::vector<std::thread> th; for (int t = 0; t < nt; ++t) th.emplace_back(body, t); for (auto &x : th) x.join(); }
    for (uint64_t r = 0; r < rd.n; ++r) res.a_rowptr[r + 1] = res.a_rowptr[r] + rownnz[r];
    for (auto &pt : parts)
    {
        res.nnzA_pre += pt.pre;
        res.a_col.insert(res.a_col.end(), pt.col.begin(), pt.col.end());
        res.a_pos.insert(res.a_pos.end(), pt.pos.begin(), pt.pos.end());
    }
    /* transpose (src/main.cpp:272-273): same entries grouped by column, rows ascending */
    res.at_colptr.assign(res.R + 1, 0);
    for (uint32_t c : res.a_col) res.at_colptr[c + 1]++;
    std::partial_sum(res.at_colptr.begin(), res.at_colptr.end(), res.at_colptr.begin());
    res.at_row.resize(res.a_col.size()); res.at_pos.resize(res.a_col.size());
    std::vector<int64_t> cur(res.at_colptr.begin(), res.at_colptr.end() - 1);
    for (uint64_t r = 0; r < rd.n; ++r)
        for (int64_t p = res.a_rowptr[r]; p < res.a_rowptr[r + 1]; ++p) { int64_t q = cur[res.a_col[p]]++; res.at_row[q] = (uint32_t)r; res.at_pos[q] = res.a_pos[p]; }
}

/*
 * B = A (x) A^T under SharedSeeds::Semiring (include/SharedSeeds.hpp:36-58), then
 * Prune(numshared <= 1) (src/SharedSeeds.cpp:8).  multiply(a,b) = {(a,b), 1};
 * add(l,r) = {l.seeds[0], r.seeds[0], l.n + r.n}; left fold in ascending column id.
 */
void spgemm(uint64_t N, Result &res)
{
    res.b_rowptr.assign(N + 1, 0);
    int nt = g_threads < 1 ? 1 : g_threads;
    struct Part { std::vector<uint32_t> col; std::vector<int32_t> num; std::vector<uint32_t> seeds; uint64_t F = 0, pre = 0; };
    std::vector<Part> parts(nt);
    std::vector<int64_t> rownnz(N, 0);
    auto body = [&](int t, int T)
    {
        Part &pt = parts[t];
        uint64_t lo = N * t / T, hi = N * (t + 1) / T;
        std::vector<int32_t> slot(N, -1);
        std::vector<uint32_t> touched, s0q, s0t, s1q, s1t; std::vector<int32_t> n;
        for (uint64_t i = lo; i < hi; ++i)
        {
            touched.clear(); s0q.clear(); s0t.clear(); s1q.clear(); s1t.clear(); n.clear();
            for (int64_t p = res.a_rowptr[i]; p < res.a_rowptr[i + 1]; ++p)
            {
                uint32_t c = res.a_col[p], pi = res.a_pos[p];
                for (int64_t q = res.at_colptr[c]; q < res.at_colptr[c + 1]; ++q)
                {
                    uint32_t j = res.at_row[q], pj = res.at_pos[q];
                    pt.F++;
                    int32_t s = slot[j];
                    if (s < 0) { slot[j] = (int32_t)touched.size(); touched.push_back(j); s0q.push_back(pi); s0t.push_back(pj); s1q.push_back(0); s1t.push_back(0); n.push_back(1); }
                    else { s1q[s] = pi; s1t[s] = pj; n[s]++; }
                }
            }
            pt.pre += touched.size();
            std::vector<uint32_t> order(touched.begin(), touched.end());
            std::sort(order.begin(), order.end());
            for (uint32_t j : order)
            {
                int32_t s = slot[j];
                if (n[s] > 1) { pt.col.push_back(j); pt.num.push_back(n[s]); pt.seeds.push_back(s0q[s]); pt.seeds.push_back(s0t[s]); pt.seeds.push_back(s1q[s]); pt.seeds.push_back(s1t[s]); rownnz[i]++; }
            }
            for (uint32_t j : touched) slot[j] = -1;
        }
    };
    if (nt == 1) body(0, 1);
    else { std::vector<std::thread> th; for (int t = 0; t < nt; ++t) th.emplace_back(body, t, nt); for (auto &x : th) x.join(); }
    for (uint64_t i = 0; i < N; ++i) res.b_rowptr[i + 1] = res.b_rowptr[i] + rownnz[i];
    for (auto &pt : parts)
    {
        res.F += pt.F; res.nnzB_pre += pt.pre;
        res.b_col.insert(res.b_col.end(), pt.col.begin(), pt.col.end());
        res.b_num.insert(res.b_num.end(), pt.num.begin(), pt.num.end());
        res.b_seeds.insert(res.b_seeds.end(), pt.seeds.begin(), pt.seeds.end());
    }
}

} // namespace

extern "C" {

/* DnaSeq::compress, src/DnaSeq.cpp:7-29.  Returns 0, or -1 on a non-nucleotide character. */
int eo_pack(const char *s, uint64_t len, uint8_t *mem)
{
    uint64_t nbytes = (len + 3) / 4;
    for (uint64_t b = 0; b < nbytes; ++b)
    {
        uint8_t byte = 0;
        for (int i = 0; i < 4 && 4 * b + i < len; ++i)
        {
            int c = char_code(s[4 * b + i]); if (c > 3) return -1;
            byte |= (uint8_t)(c << (6 - 2 * i));
        }
        mem[b] = byte;
    }
    return 0;
}

/* FASTA ingest: FastaIndex::getmydna, src/FastaIndex.cpp:254-278, + DnaSeq::compress, src/DnaSeq.cpp:7-29.
 * chunk = the bytes [chunk_pos, chunk_pos + chunk_bytes) of the FASTA file; rec = the .fai records of the reads, 3 words
 * each: len, pos (file offset of the first base), bases (bases per line; every line is followed by ONE separator byte,
 * :265-266).  Base i of a read sits at pos + i + i / bases.  A character outside ACGTN/acgtn carries code 4
 * (include/DnaSeq.hpp:136-154) and `uint8_t shift = code << (6 - 2 * i)` (src/DnaSeq.cpp:20-22) then spills into the
 * previous base of the byte: reproduced bit for bit.  Returns the arena size, or -1 if a record leaves the chunk. */
int64_t eo_fasta_pack(const char *chunk, uint64_t chunk_bytes, uint64_t chunk_pos, const uint64_t *rec, uint64_t nreads, uint8_t *packed)
{
    uint64_t head = 0;
    for (uint64_t r = 0; r < nreads; ++r)
    {
        const uint64_t len = rec[3 * r], pos = rec[3 * r + 1], bases = rec[3 * r + 2];
        const uint64_t nbytes = (len + 3) / 4;
        if (len)
        {
            if (bases == 0 || pos < chunk_pos) return -1;
            const uint64_t last = pos + (len - 1) + (len - 1) / bases;
            if (last - chunk_pos >= chunk_bytes) return -1;
        }
        for (uint64_t b = 0; b < nbytes; ++b)
        {
            uint8_t byte = 0;
            for (int i = 0; i < 4 && 4 * b + i < len; ++i)
            {
                const uint64_t j = 4 * b + i;
                const int c = char_code(chunk[pos - chunk_pos + j + j / bases]);
                byte |= (uint8_t)(c << (6 - 2 * i));
            }
            packed[head + b] = byte;
        }
        head += nbytes;
    }
    return (int64_t)head;
}

/* Transitive reduction of the overlap graph: TransitiveReduction(R), src/TransitiveReduction.cpp:3-92, with the
 * functors of include/TransitiveReduction.hpp:19-110 and Overlap::Transpose / arrows (include/Overlap.hpp:36-74).
 * Input: nnz triples (row, col, {direction, directionT, suffix, suffixT}) with distinct coordinates (the matrix
 * PairwiseAlignment builds, upper triangular in the reference).  What the reference's loop computes:
 *   R  += transpose(R) with query/target fields swapped (:16-20); an entry both hold keeps R's value (operator+ returns lhs)
 *   N   = R (x) R under MinPlusSR (:49): N(i,j).suffix_paths[2 t1 + h2] = min over k of R(i,k).suffix + R(k,j).suffix, over the
 *         k whose arrows chain (t2 != h1; entries with direction -1 have no arrows)
 *   I   = { (i,j) in R and N : direction != -1 and suffix + FUZZ >= N.suffix_paths[direction] } (:60-62), made symmetric (:70-74)
 *   T  += I (:76).  T starts as ONE explicit entry at (0,0) (:27-28: nrow copies of the coordinate (0,0)).
 *   The second round multiplies P = N, whose values are default Overlaps (direction -1: no arrows), so it adds nothing and
 *   the loop ends (:43-82) - restated as the single squaring it is.
 *   S   = the entries of R that T does not hold (:86, notB), minus direction -1 (:88).
 * Output (row-major, columns ascending): row, col, the four fields as they stand at that coordinate, src = index of the input
 * triple the entry came from, transposed = 1 if it is that triple's mirror image.  Returns the number of entries. */
uint64_t eo_transitive_reduction(int64_t n, uint64_t nnz, const int64_t *rows, const int64_t *cols, const int32_t *fields, int32_t fuzz,
                                 int64_t *orow, int64_t *ocol, int32_t *ofields, uint64_t *osrc, uint8_t *otrans)
{
    struct E { int64_t r, c; int32_t dir, dirT, suf, sufT; uint64_t src; uint8_t tr; };
    std::vector<E> e; e.reserve(2 * nnz);
    for (uint64_t i = 0; i < nnz; ++i)
    {
        const int32_t *f = fields + 4 * i;
        e.push_back({rows[i], cols[i], f[0], f[1], f[2], f[3], i, 0});
        e.push_back({cols[i], rows[i], f[1], f[0], f[3], f[2], i, 1});       /* Overlap::Transpose */
    }
    std::sort(e.begin(), e.end(), [](const E& a, const E& b) { if (a.r != b.r) return a.r < b.r; if (a.c != b.c) return a.c < b.c; return a.tr < b.tr; });
    std::vector<E> R; R.reserve(e.size());
    for (size_t i = 0; i < e.size(); ++i) if (i == 0 || e[i].r != e[i-1].r || e[i].c != e[i-1].c) R.push_back(e[i]);       /* R's own entry wins */
    std::vector<uint64_t> rp((size_t)n + 1, 0);
    for (const E& x : R) rp[(size_t)x.r + 1]++;
    for (int64_t i = 0; i < n; ++i) rp[i + 1] += rp[i];
    auto find = [&](int64_t r, int64_t c) -> int64_t
    {
        uint64_t lo = rp[r], hi = rp[r + 1];
        while (lo < hi) { uint64_t mid = (lo + hi) / 2; if (R[mid].c < c) lo = mid + 1; else hi = mid; }
        return (lo < rp[r + 1] && R[lo].c == c) ? (int64_t)lo : -1;
    };
    const int INF = std::numeric_limits<int>::max();
    std::vector<uint8_t> I(R.size(), 0);
    for (int64_t i = 0; i < n; ++i)
        for (uint64_t p = rp[i]; p < rp[i + 1]; ++p)
        {
            const E& x = R[p];
            if (x.dir == -1) continue;                                        /* GreaterThanSR */
            int best = INF;                                                   /* N(i, x.c).suffix_paths[x.dir] */
            for (uint64_t a = rp[i]; a < rp[i + 1]; ++a)
            {
                const E& e1 = R[a];
                if (e1.dir == -1) continue;
                const int t1 = (e1.dir >> 1) & 1, h1 = e1.dir & 1;
                const int64_t b = find(e1.c, x.c);
                if (b < 0) continue;
                const E& e2 = R[b];
                if (e2.dir == -1) continue;
                const int t2 = (e2.dir >> 1) & 1, h2 = e2.dir & 1;
                if (t2 == h1) continue;
                if (2 * t1 + h2 != x.dir) continue;
                best = std::min(best, e1.suf + e2.suf);
            }
            if (best != INF && x.suf + fuzz >= best) I[p] = 1;
        }
    std::vector<uint8_t> T(I);
    for (int64_t i = 0; i < n; ++i)
        for (uint64_t p = rp[i]; p < rp[i + 1]; ++p)
            if (I[p]) { int64_t q = find(R[p].c, i); if (q >= 0) T[q] = 1; }     /* I += transpose(I); only coordinates R holds matter for :86 */
    uint64_t out = 0;
    for (size_t p = 0; p < R.size(); ++p)
    {
        const E& x = R[p];
        if (T[p] || (x.r == 0 && x.c == 0) || x.dir == -1) continue;
        if (orow) { orow[out] = x.r; ocol[out] = x.c; ofields[4 * out] = x.dir; ofields[4 * out + 1] = x.dirT; ofields[4 * out + 2] = x.suf; ofields[4 * out + 3] = x.sufT;
                    osrc[out] = x.src; otrans[out] = x.tr; }
        ++out;
    }
    return out;
}

/* ASCII k-mer -> forward value, twin, canonical, hash (Kmer ctor :86-107, GetTwin, GetRep, GetHash) */
void eo_kmer_info(const char *s, int k, uint64_t *fwd, uint64_t *twin, uint64_t *rep, uint64_t *hash)
{
    uint64_t x = 0;
    for (int i = 0; i < k; ++i) x |= (uint64_t)char_code(s[i]) << (2 * (31 - i));
    *fwd = x; *twin = twin_of(x, k); *rep = rep_of(x, k); *hash = kmer_hash(*rep);
}
uint64_t eo_hash(uint64_t kmer) { return kmer_hash(kmer); }
/* GetKmerOwner, src/KmerOps.cpp:352-359 */
int eo_owner(uint64_t kmer, int nprocs)
{
    double range = (double)kmer_hash(kmer) * (double)nprocs;
    return (int)(size_t)(range / (double)UINT64_MAX);
}

uint64_t eo_rep_kmers(const uint8_t *packed, uint64_t len, int k, uint64_t *out)
{
    uint64_t n = 0;
    foreach_kmer(packed, len, k, [&](uint64_t x, uint64_t) { out[n++] = x; });
    return n;
}

/* HLL registers + estimate over all reads, as KmerOps.cpp:45-47 drives it */
double eo_hll(const uint8_t *buf, const uint64_t *off, const uint64_t *len, uint64_t n, int k, uint8_t *regs_out)
{
    Hll h;
    for (uint64_t r = 0; r < n; ++r) foreach_kmer(buf + off[r], len[r], k, [&](uint64_t x, uint64_t) { h.add_kmer(x, k); });
    if (regs_out) std::memcpy(regs_out, h.reg, Hll::SIZE);
    return h.estimate();
}
double eo_hll_estimate(const uint8_t *regs) { Hll h; std::memcpy(h.reg, regs, Hll::SIZE); return h.estimate(); }
void eo_hll_add(uint8_t *regs, const uint64_t *kmers, uint64_t n, int k)
{
    Hll h; std::memcpy(h.reg, regs, Hll::SIZE);
    for (uint64_t i = 0; i < n; ++i) h.add_kmer(kmers[i], k);
    std::memcpy(regs, h.reg, Hll::SIZE);
}

void eo_bloom_size(int64_t entries, double err, int64_t *bits, int *hashes)
{
    double bpe = -(std::log(err) / 0.480453013918201);
    *bits = (int64_t)((double)entries * bpe); *hashes = (int)std::ceil(0.693147180559945 * bpe);
}
void eo_bloom_ab(uint64_t kmer, uint64_t *a, uint64_t *b) { Bloom::ab(kmer, *a, *b); }
/* add every k-mer (Bloom::Add) to a zeroed filter of `entries` at `err`; out must hold bits/8 (+1) bytes */
void eo_bloom_fill(int64_t entries, double err, const uint64_t *kmers, uint64_t n, uint8_t *out)
{
    Bloom b(entries, err);
    for (uint64_t i = 0; i < n; ++i) b.check_add(kmers[i], true);
    std::memcpy(out, b.bf.data(), b.bytes);
}
/*
 * Pass 1 exactly as src/KmerOps.cpp:156-187 on one rank: returns the number of
 * keys in the map after the Bloom-gated first pass (arrival order = read order).
 */
uint64_t eo_pass1_keys(const uint8_t *buf, const uint64_t *off, const uint64_t *len, uint64_t n, int k, int64_t entries)
{
    Bloom bm(entries, 0.05);
    std::unordered_map<uint64_t, int> map;
    for (uint64_t r = 0; r < n; ++r)
        foreach_kmer(buf + off[r], len[r], k, [&](uint64_t x, uint64_t) {
            if (bm.check_add(x, false)) map.emplace(x, 0); else bm.check_add(x, true); });
    return map.size();
}

void *eo_run(const uint8_t *buf, const uint64_t *off, const uint64_t *len, uint64_t n, int k, int lower, int upper, int stop_after /*0 all,1 count,2 A*/)
{
    Result *res = new Result; res->N = n;
    Reads rd{buf, off, len, n};
    double t0 = now(), t = t0;
    count_kmers(rd, k, lower, upper, *res); res->secs[0] = now() - t; t = now();
    if (stop_after != 1) { build_A(rd, k, *res); res->secs[1] = now() - t; t = now(); }
    if (stop_after == 0) { spgemm(n, *res); res->secs[2] = now() - t; }
    res->secs[3] = now() - t0;
    return res;
}
void eo_free(void *h) { delete (Result*)h; }
void eo_set_threads(int t) { g_threads = t; }
void eo_set_stride(int s) { g_stride = s < 1 ? 1 : s; }
void eo_sizes(void *h, uint64_t *o /*[10]*/)
{
    Result *r = (Result*)h;
    o[0] = r->N; o[1] = r->M; o[2] = r->D; o[3] = r->R; o[4] = r->nnzA_pre; o[5] = r->a_col.size(); o[6] = r->F; o[7] = r->nnzB_pre; o[8] = r->b_col.size(); o[9] = 0;
}
void eo_secs(void *h, double *o) { std::memcpy(o, ((Result*)h)->secs, sizeof(double) * 4); }
void eo_get_kmers(void *h, uint64_t *kmer, uint32_t *count)
{
    Result *r = (Result*)h; std::memcpy(kmer, r->rel_kmer.data(), r->R * 8); std::memcpy(count, r->rel_count.data(), r->R * 4);
}
void eo_get_A(void *h, int64_t *rowptr, uint32_t *col, uint32_t *pos)
{
    Result *r = (Result*)h; std::memcpy(rowptr, r->a_rowptr.data(), r->a_rowptr.size() * 8);
    std::memcpy(col, r->a_col.data(), r->a_col.size() * 4); std::memcpy(pos, r->a_pos.data(), r->a_pos.size() * 4);
}
void eo_get_AT(void *h, int64_t *colptr, uint32_t *row, uint32_t *pos)
{
    Result *r = (Result*)h; std::memcpy(colptr, r->at_colptr.data(), r->at_colptr.size() * 8);
    std::memcpy(row, r->at_row.data(), r->at_row.size() * 4); std::memcpy(pos, r->at_pos.data(), r->at_pos.size() * 4);
}
void eo_get_B(void *h, int64_t *rowptr, uint32_t *col, int32_t *num, uint32_t *seeds)
{
    Result *r = (Result*)h; std::memcpy(rowptr, r->b_rowptr.data(), r->b_rowptr.size() * 8);
    std::memcpy(col, r->b_col.data(), r->b_col.size() * 4); std::memcpy(num, r->b_num.data(), r->b_num.size() * 4);
    std::memcpy(seeds, r->b_seeds.data(), r->b_seeds.size() * 4);
}

} // extern "C"
