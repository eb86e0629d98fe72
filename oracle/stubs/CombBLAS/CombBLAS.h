/*
 * TEST INFRASTRUCTURE ONLY (oracle/).  Nothing under elba_b200/ may include,
 * link or execute this.
 *
 * CombBLAS is an un-vendored, UN-PINNED external dependency of the reference
 * (/root/reference/usage.txt:5-6 clones the default branch of
 * github.com/PASSIONLab/CombBLAS; Makefile:21-24 expects ./CombBLAS; it is
 * not a submodule, .gitmodules:1-6) and is absent from this image.  This
 * header RESTATES the published semantics of exactly the five CombBLAS entry
 * points the hot path reaches, so that the reference's own KmerOps.cpp and
 * SharedSeeds.cpp compile unmodified and run:
 *
 *   CommGrid(comm,0,0)                         src/main.cpp:86
 *   FullyDistVec<IT,NT>(std::vector, grid)     src/KmerOps.cpp:396-398
 *   SpParMat(m,n,rows,cols,vals,false)         src/KmerOps.cpp:400
 *        duplicates (same row,col) are merged with maximum<NT> when
 *        SumDuplicates == false (CombBLAS SpParMat.cpp, SparseCommon +
 *        SpTuples::RemoveDuplicates) -> the LARGEST value survives.
 *   SpParMat copy-ctor + Transpose()           src/main.cpp:272-273
 *   Mult_AnXBn_DoubleBuff<SR,NUO,DER>(A,B)     src/SharedSeeds.cpp:7
 *        C(i,j) = SR::add-fold over k of SR::multiply(A(i,k), B(k,j)).
 *        The ORDER of that fold inside CombBLAS (heap vs hash accumulator,
 *        double-buffer halves, multiway merge) is not defined by anything in
 *        /root/reference and CombBLAS is unpinned, so the fold here is the
 *        documented canonical one: ascending k, left fold.  Pattern and
 *        numshared do not depend on it (SharedSeeds.hpp:41-46 is associative
 *        and commutative in numshared); seeds[] do -> "parity unpinned" for
 *        seeds, see DESIGN.md.
 *   Prune(pred)                                src/SharedSeeds.cpp:8
 *
 * and, for src/TransitiveReduction.cpp, what that file names:
 *   getcommgrid(), FullyDistVec(grid, n, init), SpParMat(m, n, rows, cols, scalar value, false), Apply(unary op),
 *   operator+= (union of the patterns; where both hold an entry the values are combined with NT's operator+,
 *   CombBLAS Dcsc::operator+=), the converting copy SpParMat<NT> -> SpParMat<NNT> (T += I adds a bool matrix to an int
 *   one), operator==, and EWiseApply<RETT, RETDER>(A, B, op, notB, defaultBVal): with notB == false the result holds
 *   op(A(i,j), B(i,j)) on the INTERSECTION of the patterns; with notB == true it holds op(A(i,j), defaultBVal) on the
 *   entries of A that B does NOT hold (CombBLAS ParFriends.h / Friends.h EWiseApply, restated from its documentation).
 *
 * "Ranks" are the threads of oracle/stubs/mpi.h; a distributed matrix is one
 * shared, immutable global store that every rank's handle points at.
 */
#ifndef ELBA_ORACLE_STUB_COMBBLAS_H
#define ELBA_ORACLE_STUB_COMBBLAS_H

#include <mpi.h>
#include <memory>
#include <vector>
#include <tuple>
#include <string>
#include <sstream>
#include <iostream>
#include <iomanip>
#include <unordered_map>
#include <array>
#include <algorithm>
#include <numeric>
#include <cassert>
#include <cmath>

/* the reference's functors name these unqualified (include/TransitiveReduction.hpp:19-77); CombBLAS' headers make them visible */
using std::unary_function;
using std::binary_function;

namespace combblas {

class CommGrid
{
public:
    CommGrid(MPI_Comm world, int, int) : world(world) {}
    int GetRank() const { return fake_mpi::t_rank; }
    int GetSize() const { return fake_mpi::nranks(); }
    MPI_Comm GetWorld() const { return world; }
    int GetGridRows() const { return (int)std::lround(std::sqrt((double)GetSize())); }
    int GetGridCols() const { return (int)std::lround(std::sqrt((double)GetSize())); }
    int GetRankInProcRow() const { int q = (int)std::lround(std::sqrt((double)GetSize())); return GetRank() % q; }
    int GetRankInProcCol() const { int q = (int)std::lround(std::sqrt((double)GetSize())); return GetRank() / q; }
private:
    MPI_Comm world;
};

template <class IT, class NT> class SpCCols {};
template <class IT, class NT> class SpDCCols {};

template <class IT, class NT>
class FullyDistVec
{
public:
    FullyDistVec(const std::vector<NT>& v, std::shared_ptr<CommGrid> g) : arr(v), grid(g) {}
    /* a vector of globallen copies of initval, spread over the ranks: here rank 0 holds all of it */
    FullyDistVec(std::shared_ptr<CommGrid> g, IT globallen, NT initval) : arr(fake_mpi::t_rank == 0 ? (size_t)globallen : 0, initval), grid(g) {}
    std::vector<NT> arr;
    std::shared_ptr<CommGrid> grid;
};

/* global store: CSR by row, columns ascending within a row */
template <class IT, class NT>
struct Store
{
    IT m = 0, n = 0;
    std::vector<IT> rowptr;      /* m+1 */
    std::vector<IT> col;
    std::vector<NT> val;
};

template <class IT, class NT, class DER>
class SpParMat
{
public:
    typedef Store<IT,NT> store_t;
    std::shared_ptr<store_t> st;

    SpParMat() {}
    SpParMat(std::shared_ptr<store_t> s) : st(s) {}

    SpParMat(IT m, IT n, const FullyDistVec<IT,IT>& rows, const FullyDistVec<IT,IT>& cols, const FullyDistVec<IT,NT>& vals, bool SumDuplicates)
    {
        using namespace fake_mpi;
        assert(!SumDuplicates); (void)SumDuplicates;
        int P = nranks();
        std::shared_ptr<store_t> mine;
        if (P > 1) { g_world->p0[t_rank] = &rows.arr; g_world->p1[t_rank] = &cols.arr; g_world->p2[t_rank] = &vals.arr; barrier(); }
        if (t_rank == 0)
        {
            std::vector<std::tuple<IT,IT,NT>> t;
            for (int r = 0; r < P; ++r)
            {
                const std::vector<IT>& R = P > 1 ? *(const std::vector<IT>*)g_world->p0[r] : rows.arr;
                const std::vector<IT>& C = P > 1 ? *(const std::vector<IT>*)g_world->p1[r] : cols.arr;
                const std::vector<NT>& V = P > 1 ? *(const std::vector<NT>*)g_world->p2[r] : vals.arr;
                for (size_t i = 0; i < R.size(); ++i) t.emplace_back(R[i], C[i], V[i]);
            }
            std::sort(t.begin(), t.end(), [](const auto& a, const auto& b) {
                if (std::get<0>(a) != std::get<0>(b)) return std::get<0>(a) < std::get<0>(b);
                return std::get<1>(a) < std::get<1>(b); });
            mine = std::make_shared<store_t>();
            mine->m = m; mine->n = n; mine->rowptr.assign(m + 1, 0);
            for (size_t i = 0; i < t.size(); )
            {
                size_t j = i; NT best = std::get<2>(t[i]);
                /* maximum<NT>: a < b ? b : a */
                for (++j; j < t.size() && std::get<0>(t[j]) == std::get<0>(t[i]) && std::get<1>(t[j]) == std::get<1>(t[i]); ++j)
                    best = (best < std::get<2>(t[j])) ? std::get<2>(t[j]) : best;
                mine->col.push_back(std::get<1>(t[i])); mine->val.push_back(best);
                mine->rowptr[std::get<0>(t[i]) + 1]++;
                i = j;
            }
            std::partial_sum(mine->rowptr.begin(), mine->rowptr.end(), mine->rowptr.begin());
        }
        share(mine);
    }

    SpParMat(const SpParMat& o) : st(o.st) {} /* stores are immutable: sharing == copying */
    SpParMat& operator=(const SpParMat& o) { st = o.st; return *this; }

    /* every listed (row, col) holds `val` (duplicates collapse) */
    SpParMat(IT m, IT n, const FullyDistVec<IT,IT>& rows, const FullyDistVec<IT,IT>& cols, const NT& val, bool SumDuplicates)
        : SpParMat(m, n, rows, cols, FullyDistVec<IT,NT>(std::vector<NT>(rows.arr.size(), val), rows.grid), SumDuplicates) {}

    std::shared_ptr<CommGrid> getcommgrid() const { return std::make_shared<CommGrid>(MPI_COMM_WORLD, 0, 0); }

    /* values replaced by op(value) */
    template <class Op>
    void Apply(Op op)
    {
        using namespace fake_mpi;
        std::shared_ptr<store_t> mine;
        if (nranks() > 1) barrier();
        if (t_rank == 0)
        {
            mine = std::make_shared<store_t>(*st);
            for (auto& v : mine->val) { NT x = v; v = op(x); }
        }
        share(mine);
    }

    /* union of the patterns; an entry both hold becomes lhs + rhs */
    SpParMat& operator+=(const SpParMat& rhs)
    {
        using namespace fake_mpi;
        std::shared_ptr<store_t> mine;
        if (nranks() > 1) barrier();
        if (t_rank == 0)
        {
            const store_t &a = *st, &b = *rhs.st;
            assert(a.m == b.m && a.n == b.n);
            mine = std::make_shared<store_t>();
            mine->m = a.m; mine->n = a.n; mine->rowptr.assign(a.m + 1, 0);
            for (IT r = 0; r < a.m; ++r)
            {
                IT p = a.rowptr[r], pe = a.rowptr[r+1], q = b.rowptr[r], qe = b.rowptr[r+1];
                while (p < pe || q < qe)
                {
                    if (q >= qe || (p < pe && a.col[p] < b.col[q])) { mine->col.push_back(a.col[p]); mine->val.push_back(a.val[p]); ++p; }
                    else if (p >= pe || b.col[q] < a.col[p]) { mine->col.push_back(b.col[q]); mine->val.push_back(b.val[q]); ++q; }
                    else { mine->col.push_back(a.col[p]); mine->val.push_back((NT)(a.val[p] + b.val[q])); ++p; ++q; }
                    mine->rowptr[r+1]++;
                }
            }
            std::partial_sum(mine->rowptr.begin(), mine->rowptr.end(), mine->rowptr.begin());
        }
        share(mine);
        return *this;
    }

    /* the same pattern with the values converted (CombBLAS: template conversion operator of SpParMat) */
    template <class NNT, class NDER>
    operator SpParMat<IT,NNT,NDER>() const
    {
        using namespace fake_mpi;
        typedef Store<IT,NNT> nstore;
        std::shared_ptr<nstore> mine;
        if (nranks() > 1) barrier();
        if (t_rank == 0)
        {
            mine = std::make_shared<nstore>();
            mine->m = st->m; mine->n = st->n; mine->rowptr = st->rowptr; mine->col = st->col;
            mine->val.reserve(st->val.size());
            for (const auto& v : st->val) mine->val.push_back(static_cast<NNT>(v));
        }
        SpParMat<IT,NNT,NDER> C;
        C.share(mine);
        return C;
    }

    /* the local block, column-major and doubly compressed, as src/PairwiseAlignment.cpp:16-33 walks it.  One rank only: the
     * "block" is the whole matrix (the stores are global here). */
    struct Dcsc { IT nzc = 0; std::vector<IT> cp, jc, ir; std::vector<NT> numx; };
    struct Seq
    {
        std::shared_ptr<Dcsc> d;
        Dcsc *GetDCSC() const { return d && d->nzc ? d.get() : nullptr; }
        IT getnnz() const { return d ? (IT)d->ir.size() : 0; }
    };
    /* CombBLAS hands out the matrix's own local block: it must live as long as the matrix does (callers keep raw Dcsc pointers) */
    mutable std::shared_ptr<Seq> seq_cache; mutable const void *seq_of = nullptr;
    std::shared_ptr<Seq> seqptr() const
    {
        assert(fake_mpi::nranks() == 1 && "the stand-in's seqptr() serves one rank");
        if (seq_cache && seq_of == (const void*)st.get()) return seq_cache;
        auto s = std::make_shared<Seq>(); s->d = std::make_shared<Dcsc>();
        seq_cache = s; seq_of = (const void*)st.get();
        std::vector<std::tuple<IT,IT,IT>> t;      /* (col, row, position) */
        for (IT r = 0; r < st->m; ++r) for (IT p = st->rowptr[r]; p < st->rowptr[r+1]; ++p) t.emplace_back(st->col[p], r, p);
        std::sort(t.begin(), t.end());
        s->d->cp.push_back(0);
        for (size_t i = 0; i < t.size(); ++i)
        {
            if (i == 0 || std::get<0>(t[i]) != std::get<0>(t[i-1])) { if (i) s->d->cp.push_back((IT)i); s->d->jc.push_back(std::get<0>(t[i])); }
            s->d->ir.push_back(std::get<1>(t[i])); s->d->numx.push_back(st->val[std::get<2>(t[i])]);
        }
        s->d->nzc = (IT)s->d->jc.size();
        if (!t.empty()) s->d->cp.push_back((IT)t.size());
        return s;
    }

    bool operator==(const SpParMat& rhs) const
    {
        return st->m == rhs.st->m && st->n == rhs.st->n && st->rowptr == rhs.st->rowptr && st->col == rhs.st->col && st->val == rhs.st->val;
    }

    IT getnrow() const { return st->m; }
    IT getncol() const { return st->n; }
    IT getnnz() const { return (IT)st->col.size(); }

    void Transpose()
    {
        using namespace fake_mpi;
        std::shared_ptr<store_t> mine;
        if (nranks() > 1) barrier();
        if (t_rank == 0)
        {
            mine = std::make_shared<store_t>();
            mine->m = st->n; mine->n = st->m; mine->rowptr.assign(st->n + 1, 0);
            for (IT c : st->col) mine->rowptr[c + 1]++;
            std::partial_sum(mine->rowptr.begin(), mine->rowptr.end(), mine->rowptr.begin());
            mine->col.resize(st->col.size()); mine->val.resize(st->col.size());
            std::vector<IT> cur(mine->rowptr.begin(), mine->rowptr.end() - 1);
            for (IT r = 0; r < st->m; ++r)
                for (IT p = st->rowptr[r]; p < st->rowptr[r+1]; ++p)
                { IT q = cur[st->col[p]]++; mine->col[q] = r; mine->val[q] = st->val[p]; }
        }
        share(mine);
    }

    template <class Pred>
    void Prune(Pred pred, bool inPlace = true)
    {
        using namespace fake_mpi;
        (void)inPlace;
        std::shared_ptr<store_t> mine;
        if (nranks() > 1) barrier();
        if (t_rank == 0)
        {
            mine = std::make_shared<store_t>();
            mine->m = st->m; mine->n = st->n; mine->rowptr.assign(st->m + 1, 0);
            for (IT r = 0; r < st->m; ++r)
                for (IT p = st->rowptr[r]; p < st->rowptr[r+1]; ++p)
                    if (!pred(st->val[p])) { mine->col.push_back(st->col[p]); mine->val.push_back(st->val[p]); mine->rowptr[r+1]++; }
            std::partial_sum(mine->rowptr.begin(), mine->rowptr.end(), mine->rowptr.begin());
        }
        share(mine);
    }

    /* rank 0 built `mine`; every rank adopts it */
    void share(std::shared_ptr<store_t>& mine)
    {
        using namespace fake_mpi;
        if (nranks() == 1) { st = mine; return; }
        if (t_rank == 0) g_world->p0[0] = &mine;
        barrier();
        st = *(const std::shared_ptr<store_t>*)g_world->p0[0];
        barrier();
    }
};

/*
 * Row-wise Gustavson over the shared stores; rank r computes a contiguous
 * block of output rows, rank 0 concatenates.  Fold: ascending k, left.
 */
template <typename SR, typename NUO, typename UDERO, typename IU, typename NU1, typename NU2, typename UDERA, typename UDERB>
SpParMat<IU,NUO,UDERO> Mult_AnXBn_DoubleBuff(SpParMat<IU,NU1,UDERA>& A, SpParMat<IU,NU2,UDERB>& B, bool clearA = false, bool clearB = false)
{
    using namespace fake_mpi;
    (void)clearA; (void)clearB;
    typedef Store<IU,NUO> ostore;
    const auto& a = *A.st; const auto& b = *B.st;
    assert(a.n == b.m);
    int P = nranks();
    IU lo = a.m * (IU)t_rank / P, hi = a.m * (IU)(t_rank + 1) / P;

    ostore part; part.m = hi - lo; part.n = b.n; part.rowptr.assign(part.m + 1, 0);
    std::vector<IU> slot(b.n, -1), touched;
    std::vector<NUO> acc;
    for (IU i = lo; i < hi; ++i)
    {
        touched.clear(); acc.clear();
        for (IU p = a.rowptr[i]; p < a.rowptr[i+1]; ++p)
        {
            IU k = a.col[p];
            for (IU q = b.rowptr[k]; q < b.rowptr[k+1]; ++q)
            {
                IU j = b.col[q];
                NUO prod = SR::multiply(a.val[p], b.val[q]);
                if (slot[j] < 0) { slot[j] = (IU)acc.size(); acc.push_back(prod); touched.push_back(j); }
                else acc[slot[j]] = SR::add(acc[slot[j]], prod);
            }
        }
        std::sort(touched.begin(), touched.end());
        for (IU j : touched) { part.col.push_back(j); part.val.push_back(acc[slot[j]]); }
        part.rowptr[i - lo + 1] = (IU)touched.size();
        for (IU j : touched) slot[j] = -1;
    }

    std::shared_ptr<ostore> mine;
    if (P > 1) { g_world->p1[t_rank] = &part; barrier(); }
    if (t_rank == 0)
    {
        mine = std::make_shared<ostore>();
        mine->m = a.m; mine->n = b.n; mine->rowptr.assign(1, 0);
        for (int r = 0; r < P; ++r)
        {
            const ostore& pr = P > 1 ? *(const ostore*)g_world->p1[r] : part;
            for (IU i = 0; i < pr.m; ++i) mine->rowptr.push_back(mine->rowptr.back() + pr.rowptr[i+1]);
            mine->col.insert(mine->col.end(), pr.col.begin(), pr.col.end());
            mine->val.insert(mine->val.end(), pr.val.begin(), pr.val.end());
        }
    }
    SpParMat<IU,NUO,UDERO> C;
    C.share(mine);
    return C;
}

/* element-wise op over two matrices of the same shape: see the header comment for the two modes */
template <typename RETT, typename RETDER, typename IU, typename NU1, typename NU2, typename UDERA, typename UDERB, typename BinOp>
SpParMat<IU,RETT,RETDER> EWiseApply(const SpParMat<IU,NU1,UDERA>& A, const SpParMat<IU,NU2,UDERB>& B, BinOp op, bool notB, const NU2& defaultBVal)
{
    using namespace fake_mpi;
    typedef Store<IU,RETT> ostore;
    std::shared_ptr<ostore> mine;
    if (nranks() > 1) barrier();
    if (t_rank == 0)
    {
        const auto &a = *A.st; const auto &b = *B.st;
        assert(a.m == b.m && a.n == b.n);
        mine = std::make_shared<ostore>();
        mine->m = a.m; mine->n = a.n; mine->rowptr.assign(a.m + 1, 0);
        for (IU r = 0; r < a.m; ++r)
        {
            IU q = b.rowptr[r], qe = b.rowptr[r+1];
            for (IU p = a.rowptr[r]; p < a.rowptr[r+1]; ++p)
            {
                while (q < qe && b.col[q] < a.col[p]) ++q;
                const bool both = q < qe && b.col[q] == a.col[p];
                NU1 x = a.val[p];
                if (!notB && both) { NU2 y = b.val[q]; mine->col.push_back(a.col[p]); mine->val.push_back((RETT)op(x, y)); mine->rowptr[r+1]++; }
                else if (notB && !both) { NU2 y = defaultBVal; mine->col.push_back(a.col[p]); mine->val.push_back((RETT)op(x, y)); mine->rowptr[r+1]++; }
            }
        }
        std::partial_sum(mine->rowptr.begin(), mine->rowptr.end(), mine->rowptr.begin());
    }
    SpParMat<IU,RETT,RETDER> C;
    C.share(mine);
    return C;
}

} // namespace combblas

#endif
