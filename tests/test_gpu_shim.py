"""The drop-in boundary EXECUTED: the reference's driver sequence (src/main.cpp:191-282: get_kmer_count_map_keys ->
get_kmer_count_map_values -> create_kmer_matrix -> copy + Transpose -> create_seed_matrix) over the product's shim
(elba_b200/host/elba_fe_shim.cpp) + libelba_fe.so, compiled against the reference's own headers and the runnable MPI /
CombBLAS stand-ins of oracle/stubs (oracle/Makefile target `shim`; built where /root/reference exists, travels prebuilt).
What comes out of the reference-side objects (KmerCountMap, SpParMat A, SpParMat B) must equal what the reference's own
KmerOps.cpp / SharedSeeds.cpp put there (the committed digests of tests/golden/golden.json) and the oracle."""
import os

import numpy as np
import pytest

from common import digest

pytestmark = pytest.mark.gpu


def _tier1(r):
    """The KmerCountMap in k-mer order; A's entries row-major with columns ascending.  Unlike the reference's own run
    (column id = position in the map's iteration order) the shim hands the SpParMat constructor the library's column ids,
    which already are the canonical ones (rank of the k-mer value)."""
    order = np.argsort(r.kmers, kind="stable")
    rows = np.repeat(np.arange(r.N), np.diff(r.a_rowptr))
    ka = np.lexsort((r.a_col, rows))
    return r.kmers[order], r.counts[order].astype(np.uint32), r.a_col[ka].astype(np.uint32), r.a_val[ka].astype(np.uint32)


@pytest.mark.parametrize("key", ["reads_fa_k17_l2_u8", "example_medium_k17_l2_u8", "reads_fa_k31_l2_u4", "example_medium_k31_l2_u4"])
def test_driver_sequence_through_the_shim(fixtures, golden, key):
    from oracle import oracle as O
    g = golden["configs"][key]
    if not os.path.exists(O.shim_path(g["k"], g["lower"], g["upper"])):
        pytest.skip("oracle/_ref/libelba_shim_* not built (needs the reference headers)")
    dna = fixtures(g["fixture"])
    r = O.shim_run(dna, g["k"], g["lower"], g["upper"])
    ref = O.run(dna, g["k"], g["lower"], g["upper"])
    kmers, counts, acol, aval = _tier1(r)
    # the KmerCountMap the driver holds (main.cpp:449-485 logs it; create_kmer_matrix receives it)
    assert np.array_equal(kmers, ref.kmers) and np.array_equal(counts, ref.counts)
    assert digest(kmers, counts) == g["digests"]["kmers"]
    # READIDS / POSITIONS of every entry: the (read, position) set of the k-mer (after the max-position merge)
    for i in range(0, r.R, 53):
        c = int(np.searchsorted(ref.kmers, r.kmers[i]))
        b, e = ref.at_colptr[c], ref.at_colptr[c + 1]
        n = min(int(e - b), g["upper"])
        assert set(zip(r.reads[i, :n].tolist(), r.pos[i, :n].tolist())) <= set(zip(ref.at_row[b:e].tolist(), ref.at_pos[b:e].tolist()))
        assert int(r.counts[i]) == int(ref.counts[c])
    # A as the SpParMat the reference constructor built from the shim's triples
    assert np.array_equal(r.a_rowptr, ref.a_rowptr) and np.array_equal(acol, ref.a_col) and np.array_equal(aval, ref.a_pos)
    assert digest(r.a_rowptr.astype(np.int64), acol, aval) == g["digests"]["A"]
    # B as the SpParMat<SharedSeeds> handed to PairwiseAlignment
    assert r.nnzB == g["nnzB"] and r.nnzB_pre == g["nnzB_pre"]
    assert np.array_equal(r.b_rowptr, ref.b_rowptr) and np.array_equal(r.b_col, ref.b_col) and np.array_equal(r.b_num, ref.b_num)
    assert digest(r.b_rowptr.astype(np.int64), r.b_col.astype(np.uint32), r.b_num.astype(np.int32)) == g["digests"]["B"]
    assert np.array_equal(r.b_seeds, ref.b_seeds)


def test_driver_sequence_from_the_fasta_file_through_the_shim(fixtures, golden, tmp_path):
    """src/main.cpp:130-139 + :191-282 unchanged: FastaIndex index(fasta, commgrid); DnaBuffer mydna = index.getmydna(); then
    the five functions.  In the shim library getmydna is the product's (ld --wrap -> elba_fe_getmydna -> elba_fe_ingest_fasta):
    the FASTA is parsed on the device, the DnaBuffer the driver holds equals the reference's, and counting finds the reads
    already resident.  Results == the golden digests of the reference's own run on the same reads."""
    from elba_b200 import fasta as F
    from oracle import oracle as O
    g = golden["configs"]["reads_fa_k17_l2_u8"]
    if not os.path.exists(O.shim_path(g["k"], g["lower"], g["upper"])):
        pytest.skip("oracle/_ref/libelba_shim_* not built (needs the reference headers)")
    dna = fixtures(g["fixture"])
    path = str(tmp_path / "reads.fa")
    F.write_fasta(path, [dna.read_ascii(i) for i in range(dna.size())], 80)
    r = O.shim_run(None, g["k"], g["lower"], g["upper"], fasta=path)
    kmers, counts, acol, aval = _tier1(r)
    assert digest(kmers, counts) == g["digests"]["kmers"]
    assert digest(r.a_rowptr.astype(np.int64), acol, aval) == g["digests"]["A"]
    assert r.nnzB == g["nnzB"] and r.nnzB_pre == g["nnzB_pre"]
    assert digest(r.b_rowptr.astype(np.int64), r.b_col.astype(np.uint32), r.b_num.astype(np.int32)) == g["digests"]["B"]


def test_transitive_reduction_through_the_shim(fixtures):
    """TransitiveReduction(R) (include/TransitiveReduction.hpp:17) with src/TransitiveReduction.cpp replaced by the shim's version
    (elba_fe_transitive_reduction on the device): R is built by the reference-side triple constructor, walked through
    seqptr()->GetDCSC() as the reference's consumer does, and S comes back as an SpParMat<Overlap> - equal, entry for entry
    and field for field, to what the reference's own file returns on the CPU."""
    from oracle import oracle as O
    from tr_inputs import overlap_graph, random_graph
    klu = (17, 2, 8)
    if not os.path.exists(O.shim_path(*klu)):
        pytest.skip("oracle/_ref/libelba_shim_* not built (needs the reference headers)")
    dna = fixtures("reads_fa")
    n, rows, cols, f = overlap_graph(dna, 17, 2, 8)
    got = O.ref_transitive_reduction(n, rows, cols, f, klu, shim=True)
    want = O.transitive_reduction(n, rows, cols, f)
    assert all(np.array_equal(a, b) for a, b in zip(got, want[:3]))
    if O.ref_available(*klu):
        ref = O.ref_transitive_reduction(n, rows, cols, f, klu)
        assert all(np.array_equal(a, b) for a, b in zip(got, ref))
    rng = np.random.default_rng(3)
    for trial in range(10):
        n = int(rng.integers(2, 30))
        rows, cols, f = random_graph(rng, n, density=0.5)
        got = O.ref_transitive_reduction(n, rows, cols, f, klu, shim=True)
        want = O.transitive_reduction(n, rows, cols, f)
        assert all(np.array_equal(a, b) for a, b in zip(got, want[:3])), trial
