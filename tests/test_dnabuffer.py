"""Input marshalling: the DnaBuffer layout (src/DnaBuffer.cpp:5-29, src/DnaSeq.cpp:7-54) and the synthetic generator."""
import numpy as np

from elba_b200.dnabuffer import DnaBuffer
from oracle import oracle as O


def test_pack_matches_dnaseq_compress():
    rng = np.random.default_rng(0)
    for n in (0, 1, 3, 4, 5, 17, 64, 1001):
        s = "".join("ACGTacgtNn"[c] for c in rng.integers(0, 10, n))
        d = DnaBuffer.from_strings([s])
        assert np.array_equal(d.buf, O.pack(s))
        assert d.read_ascii(0) == s.upper().replace("N", "A")


def test_reads_start_on_byte_boundaries():
    d = DnaBuffer.from_strings(["ACGTA", "C", "", "GGGGGGGGG"])
    assert list(d.offsets) == [0, 2, 3, 3] and d.getbufsize() == 6
    assert d.num_kmers(3) == 3 + 0 + 0 + 7
    s = d.slice(1, 4)
    assert s.size() == 3 and s.read_ascii(2) == "GGGGGGGGG" and int(s.offsets[0]) == 0


def test_fixture_shapes(fixtures):
    d = fixtures("reads_fa")
    assert d.size() == 227 and d.total_bases() == 3324900 and d.num_kmers(17) == 3321268
    m = fixtures("example_medium")
    assert m.size() == 1989 and m.total_bases() == 28904367 and m.num_kmers(17) == 28872543


def test_synthetic_reads_are_deterministic_and_overlap():
    from elba_b200.synth import make_dnabuffer
    a = make_dnabuffer(50_000, 120, 4000, 400, 0.1, seed=9)
    b = make_dnabuffer(50_000, 120, 4000, 400, 0.1, seed=9)
    assert np.array_equal(a.buf, b.buf) and np.array_equal(a.lengths, b.lengths)
    r = O.run(a, 17, 2, 8)
    assert r.R > 1000 and r.nnzB > 120          # 9.6x coverage: reads do share reliable k-mers
