"""Transitive reduction, CPU side: the oracle's restatement (oracle/elba_oracle.cpp::eo_transitive_reduction)
against the REFERENCE'S OWN src/TransitiveReduction.cpp + include/TransitiveReduction.hpp + include/Overlap.hpp (oracle/_ref,
compiled unmodified; CombBLAS' operations restated in oracle/stubs - parity with a real CombBLAS build is unpinned), and
against the committed digests of that run."""
import json
import os

import numpy as np
import pytest

from common import digest
from tr_inputs import overlap_graph, random_graph
from oracle import oracle as O

KLU = (17, 2, 8)
needs_ref = pytest.mark.skipif(not O.ref_available(*KLU), reason="oracle/_ref not built")
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "golden_tr.json")


@needs_ref
def test_restatement_equals_reference_on_random_graphs():
    rng = np.random.default_rng(7)
    for trial in range(120):
        n = int(rng.integers(1, 28))
        rows, cols, f = random_graph(rng, n, density=float(rng.uniform(0, 0.9)), upper_only=trial % 3 != 0)
        a = O.ref_transitive_reduction(n, rows, cols, f, KLU)
        b = O.transitive_reduction(n, rows, cols, f)
        assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1]) and np.array_equal(a[2], b[2]), (trial, n, len(rows))


def test_chain_with_a_transitive_edge():
    """A -> B -> C on the forward strand plus the transitive A -> C: the long edge goes, in both orientations."""
    rows, cols = np.array([0, 1, 0]), np.array([1, 2, 2])
    f = np.array([[1, 2, 500, 400], [1, 2, 600, 300], [1, 2, 1100, 700]], np.int32)
    r, c, of, src, tr = O.transitive_reduction(3, rows, cols, f)
    assert list(zip(r.tolist(), c.tolist())) == [(0, 1), (1, 0), (1, 2), (2, 1)]
    assert of.tolist() == [[1, 2, 500, 400], [2, 1, 400, 500], [1, 2, 600, 300], [2, 1, 300, 600]]
    assert src.tolist() == [0, 0, 1, 1] and tr.tolist() == [0, 1, 0, 1]
    # beyond FUZZ the long edge is no longer explained by the two short ones (in either orientation)
    f[2, 2], f[2, 3] = 500 + 600 - 1001, 300 + 400 - 1001
    assert len(O.transitive_reduction(3, rows, cols, f)[0]) == 6


@needs_ref
def test_overlap_graph_of_the_reference_fixture(fixtures):
    dna = fixtures("reads_fa")
    n, rows, cols, f = overlap_graph(dna, 17, 2, 8)
    assert len(rows) > 300
    a = O.ref_transitive_reduction(n, rows, cols, f, KLU)
    b = O.transitive_reduction(n, rows, cols, f)
    assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1]) and np.array_equal(a[2], b[2])
    assert len(b[0]) < 2 * len(rows), "some edges are transitive"
    # S is symmetric
    s = set(zip(b[0].tolist(), b[1].tolist()))
    assert all((c, r) in s for r, c in s)


def test_golden_digest_of_the_reference_run(fixtures):
    """tests/golden/golden_tr.json: digests of S as the reference's own TransitiveReduction.cpp produced it (make_golden_tr.py)."""
    g = json.load(open(GOLD))
    for key, want in g.items():
        dna = fixtures(want["fixture"])
        n, rows, cols, f = overlap_graph(dna, want["k"], want["lower"], want["upper"])
        assert digest(rows, cols, f) == want["input_digest"], key
        r, c, of, _, _ = O.transitive_reduction(n, rows, cols, f)
        assert len(r) == want["nnzS"] and digest(r, c, of) == want["digest"], key
