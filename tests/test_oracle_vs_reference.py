"""The oracle against oracle/_ref (the reference's own sources, compiled unmodified) run live.  Needs the prebuilt
oracle/_ref libraries (built where /root/reference exists; they travel with the snapshot)."""
import numpy as np
import pytest

from oracle import oracle as O

needs_ref = pytest.mark.skipif(not O.ref_available(17, 2, 8), reason="oracle/_ref not built (reference tree absent)")


@needs_ref
def test_rep_kmer_stream_vs_reference(fixtures):
    dna = fixtures("reads_fa")
    for i in (0, 3, 100, 226):
        o, n = int(dna.offsets[i]), int(dna.lengths[i])
        for k in (17, 31):
            lo, up = (2, 8) if k == 17 else (2, 4)
            assert np.array_equal(O.rep_kmers(dna.buf[o:], n, k), O.ref_rep_kmers(dna.buf[o:], n, k, lo, up))


@needs_ref
@pytest.mark.parametrize("nranks", [1, 4])
def test_whole_path_vs_reference(fixtures, nranks):
    """Tier 1 bit-exact; the reference's column ids (unordered_map iteration order, rank-dependent) are mapped to
    the canonical ones before comparing."""
    dna = fixtures("reads_fa")
    ref = O.ref_run(dna, 17, 2, 8, nranks=nranks)
    r = O.run(dna, 17, 2, 8)
    order = np.argsort(ref.kmers)
    assert np.array_equal(ref.kmers[order], r.kmers) and np.array_equal(ref.counts[order], r.counts.astype(np.int32))
    rank = np.empty(ref.R, np.int64)
    rank[order] = np.arange(ref.R)
    rows = np.repeat(np.arange(r.N), np.diff(ref.a_rowptr))
    acol = rank[ref.a_col]
    ka = np.lexsort((acol, rows))
    assert np.array_equal(ref.a_rowptr, r.a_rowptr) and np.array_equal(acol[ka], r.a_col) and np.array_equal(ref.a_val[ka], r.a_pos)
    assert np.array_equal(ref.b_rowptr, r.b_rowptr) and np.array_equal(ref.b_col, r.b_col) and np.array_equal(ref.b_num, r.b_num)
    assert ref.nnzB_pre == r.nnzB_pre
    # per reliable k-mer: the same multiset of (read, pos) as the reference's READIDS/POSITIONS arrays
    got = {}
    for c in range(r.R):
        b, e = r.at_colptr[c], r.at_colptr[c + 1]
        got[int(r.kmers[c])] = set(zip(r.at_row[b:e].tolist(), r.at_pos[b:e].tolist()))
    for i in range(0, ref.R, 97):
        n = int(ref.counts[i])
        want = {}
        for rd, ps in zip(ref.reads[i, :n].tolist(), ref.pos[i, :n].tolist()):
            want[rd] = max(want.get(rd, -1), ps)          # (read, column) duplicates keep the largest position
        assert got[int(ref.kmers[i])] == set(want.items())
