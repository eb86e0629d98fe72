"""Transitive reduction on the device (elba_fe_transitive_reduction, transitive.cuh) through the C ABI, bit-exact
against the oracle's restatement (itself pinned against the reference's own TransitiveReduction.cpp, tests/test_transitive_host.py)
and against the committed digests of the reference run."""
import json
import os

import numpy as np
import pytest

from common import digest
from tr_inputs import overlap_graph, random_graph

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "golden_tr.json")


def _ctx():
    from elba_b200 import frontend
    return frontend.Context(frontend.Params(k=17, lower=2, upper=8, device=0))


def _same(got, want, what):
    for a, b, name in zip(got, want, ("row", "col", "fields", "src", "transposed")):
        assert np.array_equal(a, b), (what, name, len(a), len(b))


def test_random_graphs_vs_oracle():
    from oracle import oracle as O
    rng = np.random.default_rng(11)
    ctx = _ctx()
    for trial in range(80):
        n = int(rng.integers(1, 40))
        rows, cols, f = random_graph(rng, n, density=float(rng.uniform(0, 0.9)), upper_only=trial % 3 != 0)
        _same(ctx.transitive_reduction(n, rows, cols, f), O.transitive_reduction(n, rows, cols, f), (trial, n, len(rows)))
    # a larger one: 3000 reads, ~40 overlaps per read
    n = 3000
    i = rng.integers(0, n, 60000); j = np.minimum(n - 1, i + rng.integers(1, 60, 60000))
    keep = i < j
    key = np.unique(i[keep] * n + j[keep])
    rows, cols = key // n, key % n
    m = len(rows)
    f = np.stack([rng.integers(-1, 4, m), rng.integers(-1, 4, m), rng.integers(0, 5000, m), rng.integers(0, 5000, m)], 1).astype(np.int32)
    _same(ctx.transitive_reduction(n, rows, cols, f), O.transitive_reduction(n, rows, cols, f), "large")
    assert ctx.timings()["transitive_ms"] > 0
    ctx.close()


def test_fixture_overlap_graphs_and_golden_digests(fixtures):
    from oracle import oracle as O
    g = json.load(open(GOLD))
    ctx = _ctx()
    for key, want in g.items():
        dna = fixtures(want["fixture"])
        n, rows, cols, f = overlap_graph(dna, want["k"], want["lower"], want["upper"])
        got = ctx.transitive_reduction(n, rows, cols, f)
        _same(got, O.transitive_reduction(n, rows, cols, f), key)
        assert len(got[0]) == want["nnzS"] and digest(got[0], got[1], got[2]) == want["digest"], key
    ctx.close()


def test_edge_cases_and_errors():
    from elba_b200 import frontend
    ctx = _ctx()
    e = np.zeros(0, np.int64)
    assert len(ctx.transitive_reduction(5, e, e, np.zeros((0, 4), np.int32))[0]) == 0
    # an explicit entry at (0, 0) never survives (T starts with one there, src/TransitiveReduction.cpp:27-28,86); other diagonal entries do
    r, c, f, src, tr = ctx.transitive_reduction(3, np.array([0, 1]), np.array([0, 1]), np.array([[0, 0, 5, 5], [0, 0, 7, 7]], np.int32))
    assert r.tolist() == [1] and c.tolist() == [1] and tr.tolist() == [0]
    with pytest.raises(frontend.FrontEndError):
        ctx.transitive_reduction(3, np.array([0]), np.array([3]), np.array([[1, 2, 5, 5]], np.int32))      # read id out of range
    with pytest.raises(frontend.FrontEndError):
        ctx.transitive_reduction(3, np.array([0]), np.array([1]), np.array([[4, 2, 5, 5]], np.int32))      # direction outside -1..3
    ctx.close()
