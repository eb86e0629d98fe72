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


# ---- the device-side result digests (elba_b200/csrc/digest.cuh) restated in numpy: sums mod 2^64 of one mix per entry ----
_U = np.uint64
_C1, _C2, _C3 = _U(0x9E3779B97F4A7C15), _U(0xC2B2AE3D27D4EB4F), _U(0x165667B19E3779F9)


def _mix64(x):
    x = x.astype(np.uint64, copy=True)
    with np.errstate(over="ignore"):
        x ^= x >> _U(32); x *= _U(0x9E3779B97F4A7C15); x ^= x >> _U(29); x *= _U(0xBF58476D1CE4E5B9); x ^= x >> _U(32)
    return x


def _hex(v):
    return f"{int(v):016x}"


def result_digests(kmers, counts, a_rowptr, a_col, a_pos, b_rowptr, b_col, b_num, b_seeds, row0=0, col0=0):
    """What elba_fe_digests returns for these results (global ids: row0 / col0 are added to local ones)."""
    with np.errstate(over="ignore"):
        dk = _mix64(kmers.astype(np.uint64) ^ _mix64(counts.astype(np.uint64) + _C1)).sum(dtype=np.uint64)
        rows = np.repeat(np.arange(len(a_rowptr) - 1, dtype=np.uint64), np.diff(a_rowptr)) + _U(row0)
        da = _mix64(_mix64((rows << _U(32)) | a_col.astype(np.uint64)) ^ (a_pos.astype(np.uint64) + _C2)).sum(dtype=np.uint64)
        rows = np.repeat(np.arange(len(b_rowptr) - 1, dtype=np.uint64), np.diff(b_rowptr)) + _U(row0)
        cols = b_col.astype(np.uint64) + _U(col0)
        rc = _mix64((rows << _U(32)) | cols)
        db = _mix64(rc + b_num.astype(np.uint32).astype(np.uint64) * _C3).sum(dtype=np.uint64)
        s = np.asarray(b_seeds).reshape(-1, 4).astype(np.uint64)
        ds = _mix64(_mix64(rc ^ ((s[:, 0] << _U(32)) | s[:, 1])) ^ (((s[:, 2] << _U(32)) | s[:, 3]) + _C1)).sum(dtype=np.uint64)
    return dict(kmers=_hex(dk), A=_hex(da), B=_hex(db), seeds=_hex(ds))


def oracle_result_digests(r):
    return result_digests(r.kmers, r.counts, r.a_rowptr, r.a_col, r.a_pos, r.b_rowptr, r.b_col, r.b_num, r.b_seeds)
