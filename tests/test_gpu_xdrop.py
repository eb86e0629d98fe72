"""The first CUDA version of the NEXT hot-path row: X-drop seed-and-extend of B's nonzeros
(elba_b200/csrc/xdrop.cuh, through the C ABI: elba_fe_align / elba_fe_get_alignments) against the CPU oracle
(oracle/xdrop_oracle.cpp) and the digests the reference's own XDropAligner.cpp + Overlap.cpp produced
(tests/golden/golden_xdrop.json).  Bit-exact: 13 integer fields per aligned pair."""
import json
import os

import numpy as np
import pytest

from common import digest

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))
GOLD = json.load(open(os.path.join(HERE, "golden", "golden_xdrop.json")))


def _align(dna, k, lo, up, scoring):
    from elba_b200 import frontend
    ctx = frontend.Context(frontend.Params(k=k, lower=lo, upper=up))
    ctx.upload(dna)
    ctx.run()
    rows, cols, out = ctx.align(*scoring)
    B = ctx.B()
    tm = ctx.timings()
    ctx.close()
    return rows, cols, out, B, tm


@pytest.mark.parametrize("key", sorted(GOLD))
def test_xdrop_matches_reference_digests(fixtures, key):
    from oracle import oracle as O
    g = GOLD[key]
    dna = fixtures(g["fixture"])
    rows, cols, out, B, tm = _align(dna, g["k"], g["lower"], g["upper"], (g["mat"], g["mis"], g["gap"], g["dropoff"]))
    brp, bcol, bnum, bseeds = B
    er, ec, sq, st = O.alignment_pairs(brp, bcol, bseeds)
    assert np.array_equal(rows, er) and np.array_equal(cols, ec), "aligned pairs: strict upper triangle of B in row-major order"
    assert len(rows) == g["pairs"] and digest(er, ec, sq, st) == g["pairs_digest"]
    assert int(out[:, 6].sum()) == g["passed"] and int(out[:, 4].astype(np.int64).sum()) == g["score_sum"]
    assert digest(out) == g["digest"], "13 fields per pair vs the reference's own aligner"
    assert tm["align_ms"] > 0


def test_xdrop_vs_oracle_on_synthetic_reads():
    """CLR-like (wide bands, heavy pruning) and HiFi-like reads, several scoring schemes."""
    from elba_b200.synth import make_dnabuffer
    from oracle import oracle as O
    for err, k, lo, up, sc in ((0.12, 17, 2, 8, (1, -1, -1, 15)), (0.12, 17, 2, 8, (1, -1, -1, 49)), (0.01, 31, 2, 4, (1, -1, -1, 15)),
                               (0.05, 21, 2, 6, (2, -3, -2, 25)), (0.01, 32, 2, 4, (1, -2, -1, 7))):
        dna = make_dnabuffer(genome_len=60_000, n_reads=160, mean_len=6000, sd_len=900, err=err, seed=11)
        rows, cols, out, B, _ = _align(dna, k, lo, up, sc)
        brp, bcol, bnum, bseeds = B
        er, ec, sq, st = O.alignment_pairs(brp, bcol, bseeds)
        assert np.array_equal(rows, er) and np.array_equal(cols, ec) and len(rows) > 100
        want = O.xdrop(dna, k, er, ec, sq, st, *sc)
        bad = np.where((out != want).any(axis=1))[0]
        assert len(bad) == 0, (err, k, sc, len(bad), out[bad[:3]].tolist(), want[bad[:3]].tolist())


def test_xdrop_edge_cases():
    """No nonzeros, reads shorter than k, seeds at the very ends of the reads (one extension direction has nothing to do)."""
    from elba_b200.dnabuffer import DnaBuffer
    from oracle import oracle as O
    rng = np.random.default_rng(4)
    g = "".join("ACGT"[c] for c in rng.integers(0, 4, 5000))
    rcomp = lambda s: s[::-1].translate(str.maketrans("ACGT", "TGCA"))
    seqs = [g[0:1500], g[1000:2600], rcomp(g[1200:2900]), g[2500:2540], g[2500:2560], "ACGT", "", g[0:1500], g[1480:3000], rcomp(g[0:700])]
    dna = DnaBuffer.from_strings(seqs)
    for k in (17, 31):
        rows, cols, out, B, _ = _align(dna, k, 2, 8, (1, -1, -1, 15))
        brp, bcol, bnum, bseeds = B
        er, ec, sq, st = O.alignment_pairs(brp, bcol, bseeds)
        assert np.array_equal(rows, er) and np.array_equal(cols, ec) and len(rows) >= 4
        assert np.array_equal(out, O.xdrop(dna, k, er, ec, sq, st, 1, -1, -1, 15))
    empty = DnaBuffer.from_strings(["ACGTACGTACGTACGTACGTAAA", "TTTTGGGGCCCCAAAATTTTGGGC"])
    rows, cols, out, B, _ = _align(empty, 17, 2, 8, (1, -1, -1, 15))
    assert len(rows) == 0 and out.shape == (0, 13)
