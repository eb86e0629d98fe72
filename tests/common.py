"""Helpers shared by the parity tests."""
import hashlib

import numpy as np


def digest(*arrays):
    h = hashlib.sha256()
    for a in arrays:
        h.update(np.ascontiguousarray(a).tobytes())
    return h.hexdigest()


def oracle_digests(r):
    """Tier-1 digests of an OracleResult, same recipe as tests/golden/make_golden.py::tier1_from_ref."""
    return dict(kmers=digest(r.kmers, r.counts.astype(np.uint32)),
                A=digest(r.a_rowptr.astype(np.int64), r.a_col.astype(np.uint32), r.a_pos.astype(np.uint32)),
                B=digest(r.b_rowptr.astype(np.int64), r.b_col.astype(np.uint32), r.b_num.astype(np.int32)))


def revcomp_codes(c):
    return (3 - c)[::-1]


def check_seeds_valid(dna, k, b_rowptr, b_col, b_seeds, max_checks=20000, seed=0):
    """The reference's own property test (test.py:40-65; XDropAligner.cpp:239-254): the k-mer at begQ in read Q
    equals the k-mer at begT in read T or its reverse complement, for both stored seeds."""
    rows = np.repeat(np.arange(len(b_rowptr) - 1), np.diff(b_rowptr))
    nnz = len(b_col)
    idx = np.arange(nnz)
    if nnz > max_checks:
        idx = np.random.default_rng(seed).choice(nnz, max_checks, replace=False)
    cache = {}

    def codes(i):
        if i not in cache:
            cache[i] = dna.read_codes(int(i))
        return cache[i]
    bad = 0
    for e in idx:
        q, t = int(rows[e]), int(b_col[e])
        cq, ct = codes(q), codes(t)
        for s in (0, 1):
            bq, bt = int(b_seeds[e, 2 * s]), int(b_seeds[e, 2 * s + 1])
            a, b = cq[bq:bq + k], ct[bt:bt + k]
            if len(a) != k or len(b) != k or not (np.array_equal(a, b) or np.array_equal(a, revcomp_codes(b))):
                bad += 1
    return bad
