"""The oracle of the NEXT hot-path row (X-drop seed-and-extend of B's nonzeros, the stage that consumes
the overlap matrix) against the reference: golden digests made by the reference's own XDropAligner.cpp + Overlap.cpp
(tests/golden/make_golden_xdrop.py), and the two run side by side where oracle/_ref is present.  CPU only: the CUDA kernel
for this row (elba_b200/csrc/xdrop.cuh) is checked against exactly this in tests/test_gpu_xdrop.py."""
import json
import os

import numpy as np
import pytest

from common import digest
from oracle import oracle as O

HERE = os.path.dirname(os.path.abspath(__file__))
GOLD = json.load(open(os.path.join(HERE, "golden", "golden_xdrop.json")))
needs_ref = pytest.mark.skipif(not O.ref_available(17, 2, 8), reason="oracle/_ref not built (reference tree absent)")


@pytest.mark.parametrize("key", sorted(GOLD))
def test_xdrop_oracle_matches_reference_digests(fixtures, key):
    g = GOLD[key]
    dna = fixtures(g["fixture"])
    r = O.run(dna, g["k"], g["lower"], g["upper"])
    rows, cols, sq, st = O.alignment_pairs(r.b_rowptr, r.b_col, r.b_seeds)
    assert len(rows) == g["pairs"] and digest(rows, cols, sq, st) == g["pairs_digest"]
    res = O.xdrop(dna, g["k"], rows, cols, sq, st, g["mat"], g["mis"], g["gap"], g["dropoff"])
    assert int(res[:, 6].sum()) == g["passed"] and int(res[:, 4].astype(np.int64).sum()) == g["score_sum"]
    assert int(res[:, 7].sum()) == g["containedQ"] and int(res[:, 8].sum()) == g["containedT"] and int(res[:, 5].sum()) == g["rc"]
    assert digest(res) == g["digest"]


@needs_ref
def test_xdrop_oracle_vs_reference_live_on_synthetic_reads():
    """CLR-like reads (15 % errors: wide bands, many pruned cells) and HiFi-like reads, random seeds included: pairs that do
    not share the seed k-mer must come back as the reference's 'no alignment' (score -1)."""
    from elba_b200.synth import make_dnabuffer
    for err, k, lo, up, drop in ((0.12, 17, 2, 8, 15), (0.12, 17, 2, 8, 40), (0.01, 31, 2, 4, 15)):
        dna = make_dnabuffer(genome_len=60_000, n_reads=160, mean_len=6000, sd_len=900, err=err, seed=11)
        r = O.run(dna, k, lo, up)
        rows, cols, sq, st = O.alignment_pairs(r.b_rowptr, r.b_col, r.b_seeds)
        assert len(rows) > 200
        rng = np.random.default_rng(3)
        # spoil a tenth of the seeds
        bad = rng.choice(len(rows), len(rows) // 10, replace=False)
        sq = sq.copy(); sq[bad] = rng.integers(0, 500, len(bad)).astype(np.uint32)
        a = O.xdrop(dna, k, rows, cols, sq, st, 1, -1, -1, drop)
        b = O.ref_xdrop(dna, k, lo, up, rows, cols, sq, st, 1, -1, -1, drop)
        assert np.array_equal(a, b), (err, k, drop)
        assert (a[bad, 4] == -1).sum() > len(bad) // 2
