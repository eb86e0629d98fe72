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


@pytest.mark.reference
def test_shim_host_logic_with_a_mock_abi(fixtures, tmp_path):
    """The drop-in TransitiveReduction of elba_b200/host/elba_fe_shim.cpp, compiled against the reference's headers and linked
    against a mock C ABI backed by the oracle (tests/host_shim/tr_mock_harness.cpp): the host logic around the device call."""
    import subprocess
    ref = "/root/reference"
    if not os.path.exists(os.path.join(ref, "include", "TransitiveReduction.hpp")):
        pytest.skip("reference tree not present")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    exe = str(tmp_path / "tr_harness")
    cmd = ["g++", "-O1", "-std=c++17", "-pthread", "-Wno-deprecated", "-w", "-DELBA_FE_SHIM_TR", "-DKMER_SIZE=17", "-DLOWER_KMER_FREQ=2", "-DUPPER_KMER_FREQ=8", "-DLOG_LEVEL=0",
           f'-DSHIM_CPP="{root}/elba_b200/host/elba_fe_shim.cpp"', f'-DORACLE_SO="{root}/oracle/libelba_oracle.so"',
           "-I", os.path.join(root, "oracle", "stubs"), "-I", os.path.join(ref, "include"), "-I", os.path.join(ref, "src"), "-I", os.path.join(root, "include"),
           "-o", exe, os.path.join(root, "tests", "host_shim", "tr_mock_harness.cpp")] + [os.path.join(ref, "src", f) for f in
           ("Overlap.cpp", "XDropAligner.cpp", "DnaSeq.cpp", "DnaBuffer.cpp", "FastaIndex.cpp", "Logger.cpp", "HashFuncs.cpp")] + ["-ldl"]
    stubs = str(tmp_path / "abi_stubs.o")
    p = subprocess.run(["gcc", "-w", "-c", os.path.join(root, "tests", "host_shim", "abi_stubs.c"), "-o", stubs], capture_output=True, text=True)
    assert p.returncode == 0, p.stderr[-3000:]
    p = subprocess.run(cmd[:-1] + [stubs, "-ldl"], capture_output=True, text=True)
    assert p.returncode == 0, p.stderr[-3000:]
    rng = np.random.default_rng(5)
    cases = [(n,) + random_graph(rng, n, 0.5) for n in rng.integers(2, 30, 12).tolist()] + [overlap_graph(fixtures("reads_fa"), 17, 2, 8)]
    for n, rows, cols, f in cases:
        inp = str(tmp_path / "in.txt")
        with open(inp, "w") as fo:
            fo.write(f"{n} {len(rows)}\n")
            for r, c, x in zip(rows, cols, f):
                fo.write(f"{r} {c} {x[0]} {x[1]} {x[2]} {x[3]}\n")
        q = subprocess.run([exe, inp], capture_output=True, text=True)
        assert q.returncode == 0, q.stderr[-2000:]
        got = np.array([[int(v) for v in l.split()] for l in q.stdout.strip().splitlines()], dtype=np.int64).reshape(-1, 10)
        r, c, of, src, tr = O.transitive_reduction(n, rows, cols, f)
        src = src.astype(np.int64)
        assert np.array_equal(got[:, 0], r) and np.array_equal(got[:, 1], c) and np.array_equal(got[:, 2:6], of)
        # the rest of the payload: the harness gives entry e the lengths (1000 + e, 2000 + e) and containedQ; mirror images carry them swapped
        assert np.array_equal(got[:, 6], np.where(tr == 0, 1000 + src, 2000 + src)) and np.array_equal(got[:, 7], np.where(tr == 0, 2000 + src, 1000 + src))
        assert np.array_equal(got[:, 8], (tr == 0).astype(np.int64)) and np.array_equal(got[:, 9], (tr == 1).astype(np.int64))
