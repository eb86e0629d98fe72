"""ctypes bindings for the CPU oracle.  TEST INFRASTRUCTURE ONLY.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import this module.  Nothing under ``elba_b200/`` does.

Two libraries:

* ``libelba_oracle.so`` (``elba_oracle.cpp``) - the restatement; built anywhere by ``make -C oracle``.
* ``_ref/libelba_ref_k<K>_l<L>_u<U>.so`` (``ref_wrap.cpp``) - the reference's own sources; can only be
  BUILT where ``/root/reference`` exists, but the built files travel with the repo snapshot.
"""
from __future__ import annotations

import ctypes
import os
import subprocess
from dataclasses import dataclass, field

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
_c = ctypes
_u64, _i64, _i32, _dbl, _vp = _c.c_uint64, _c.c_int64, _c.c_int, _c.c_double, _c.c_void_p


def _p(a: np.ndarray):
    return a.ctypes.data_as(_vp)


def build(ref: bool = True) -> None:
    """Compile the oracle (and oracle/_ref when the reference tree is present)."""
    target = "all" if ref else os.path.join(HERE, "libelba_oracle.so")
    subprocess.run(["make", "-s", "-C", HERE, target], check=True)


_lib = None


def lib():
    global _lib
    if _lib is None:
        path = os.path.join(HERE, "libelba_oracle.so")
        if not os.path.exists(path):
            build(ref=False)
        L = _c.CDLL(path)
        L.eo_run.restype = _vp
        L.eo_hll.restype = _dbl
        L.eo_hll_estimate.restype = _dbl
        L.eo_hash.restype = _u64
        L.eo_rep_kmers.restype = _u64
        L.eo_pass1_keys.restype = _u64
        _lib = L
    return _lib


@dataclass
class OracleResult:
    N: int = 0
    M: int = 0
    D: int = 0
    R: int = 0
    nnzA_pre: int = 0
    nnzA: int = 0
    F: int = 0
    nnzB_pre: int = 0
    nnzB: int = 0
    kmers: np.ndarray = None
    counts: np.ndarray = None
    a_rowptr: np.ndarray = None
    a_col: np.ndarray = None
    a_pos: np.ndarray = None
    at_colptr: np.ndarray = None
    at_row: np.ndarray = None
    at_pos: np.ndarray = None
    b_rowptr: np.ndarray = None
    b_col: np.ndarray = None
    b_num: np.ndarray = None
    b_seeds: np.ndarray = None
    secs: dict = field(default_factory=dict)


def pack(seq: str) -> np.ndarray:
    out = np.zeros((len(seq) + 3) // 4, np.uint8)
    rc = lib().eo_pack(seq.encode("ascii"), _u64(len(seq)), _p(out))
    if rc != 0:
        raise ValueError("non-nucleotide character")
    return out


def kmer_info(s: str):
    k = len(s)
    v = [_u64() for _ in range(4)]
    lib().eo_kmer_info(s.encode("ascii"), k, *[_c.byref(x) for x in v])
    return tuple(int(x.value) for x in v)  # fwd, twin, rep, hash(rep)


def kmer_hash(x: int) -> int:
    return int(lib().eo_hash(_u64(x)))


def owner(x: int, nprocs: int) -> int:
    return int(lib().eo_owner(_u64(x), nprocs))


def rep_kmers(packed: np.ndarray, length: int, k: int) -> np.ndarray:
    out = np.zeros(max(length - k + 1, 0), np.uint64)
    n = lib().eo_rep_kmers(_p(np.ascontiguousarray(packed)), _u64(length), k, _p(out))
    return out[:n]


def all_rep_kmers(dna, k: int) -> np.ndarray:
    """Concatenated canonical k-mer stream of every read, read order then position order."""
    parts = [rep_kmers(dna.buf[int(dna.offsets[i]):], int(dna.lengths[i]), k) for i in range(dna.size())]
    return np.concatenate(parts) if parts else np.zeros(0, np.uint64)


def hll(dna, k: int):
    regs = np.zeros(4096, np.uint8)
    est = lib().eo_hll(_p(dna.buf), _p(dna.offsets), _p(dna.lengths), _u64(dna.size()), k, _p(regs))
    return float(est), regs


def hll_estimate(regs: np.ndarray) -> float:
    return float(lib().eo_hll_estimate(_p(np.ascontiguousarray(regs, np.uint8))))


def bloom_size(entries: int, err: float = 0.05):
    bits, hashes = _i64(), _i32()
    lib().eo_bloom_size(_i64(entries), _dbl(err), _c.byref(bits), _c.byref(hashes))
    return int(bits.value), int(hashes.value)


def bloom_ab(x: int):
    a, b = _u64(), _u64()
    lib().eo_bloom_ab(_u64(x), _c.byref(a), _c.byref(b))
    return int(a.value), int(b.value)


def bloom_fill(entries: int, err: float, kmers: np.ndarray) -> np.ndarray:
    bits, _ = bloom_size(entries, err)
    out = np.zeros(bits // 8 + (1 if bits % 8 else 0), np.uint8)
    kmers = np.ascontiguousarray(kmers, np.uint64)
    lib().eo_bloom_fill(_i64(entries), _dbl(err), _p(kmers), _u64(len(kmers)), _p(out))
    return out


def pass1_keys(dna, k: int, entries: int) -> int:
    return int(lib().eo_pass1_keys(_p(dna.buf), _p(dna.offsets), _p(dna.lengths), _u64(dna.size()), k, _i64(entries)))


def run(dna, k: int, lower: int, upper: int, stop_after: int = 0, threads: int = 1, stride: int = 1) -> OracleResult:
    """The whole hot path on the CPU.  stop_after: 0 = through B, 1 = counting only, 2 = through A.
    stride: the legacy -s flag (README.md:85); the reference itself hard-wires 1."""
    L = lib()
    L.eo_set_threads(int(threads))
    L.eo_set_stride(int(stride))
    h = _vp(L.eo_run(_p(dna.buf), _p(dna.offsets), _p(dna.lengths), _u64(dna.size()), k, lower, upper, stop_after))
    try:
        sz = np.zeros(10, np.uint64)
        L.eo_sizes(h, _p(sz))
        r = OracleResult(*[int(x) for x in sz[:9]])
        secs = np.zeros(4)
        L.eo_secs(h, _p(secs))
        r.secs = dict(count=secs[0], build_A=secs[1], spgemm=secs[2], total=secs[3])
        r.kmers = np.zeros(r.R, np.uint64)
        r.counts = np.zeros(r.R, np.uint32)
        L.eo_get_kmers(h, _p(r.kmers), _p(r.counts))
        if stop_after != 1:
            r.a_rowptr = np.zeros(r.N + 1, np.int64)
            r.a_col = np.zeros(r.nnzA, np.uint32)
            r.a_pos = np.zeros(r.nnzA, np.uint32)
            L.eo_get_A(h, _p(r.a_rowptr), _p(r.a_col), _p(r.a_pos))
            r.at_colptr = np.zeros(r.R + 1, np.int64)
            r.at_row = np.zeros(r.nnzA, np.uint32)
            r.at_pos = np.zeros(r.nnzA, np.uint32)
            L.eo_get_AT(h, _p(r.at_colptr), _p(r.at_row), _p(r.at_pos))
        if stop_after == 0:
            r.b_rowptr = np.zeros(r.N + 1, np.int64)
            r.b_col = np.zeros(r.nnzB, np.uint32)
            r.b_num = np.zeros(r.nnzB, np.int32)
            r.b_seeds = np.zeros((r.nnzB, 4), np.uint32)
            L.eo_get_B(h, _p(r.b_rowptr), _p(r.b_col), _p(r.b_num), _p(r.b_seeds))
        return r
    finally:
        L.eo_free(h)


# ---------------------------------------------------------------------------
# oracle/_ref : the reference's own sources
# ---------------------------------------------------------------------------
_ref_libs = {}


def ref_path(k: int, lower: int, upper: int) -> str:
    return os.path.join(HERE, "_ref", f"libelba_ref_k{k}_l{lower}_u{upper}.so")


def ref_available(k: int, lower: int, upper: int) -> bool:
    return os.path.exists(ref_path(k, lower, upper))


def ref_lib(k: int, lower: int, upper: int):
    key = (k, lower, upper)
    if key not in _ref_libs:
        path = ref_path(*key)
        if not os.path.exists(path):
            if not os.path.exists("/root/reference/src/KmerOps.cpp"):
                raise FileNotFoundError(path + " (reference tree absent: cannot build it here)")
            subprocess.run(["make", "-s", "-C", HERE, "ref", f"KLU={k}_{lower}_{upper}"], check=True)
        L = _c.CDLL(path)
        L.ref_run.restype = _vp
        L.ref_hll.restype = _dbl
        L.ref_hash.restype = _u64
        L.ref_rep_kmers.restype = _u64
        L.ref_bloom_new.restype = _vp
        L.ref_bloom_bits.restype = _i64
        _ref_libs[key] = L
    return _ref_libs[key]


@dataclass
class RefResult:
    R: int = 0
    N: int = 0
    nnzA: int = 0
    nnzB: int = 0
    nnzB_pre: int = 0
    keys_after_pass1: int = 0
    nranks: int = 1
    kmers: np.ndarray = None      # reference column-id order
    counts: np.ndarray = None
    reads: np.ndarray = None      # R x U
    pos: np.ndarray = None        # R x U
    a_rowptr: np.ndarray = None
    a_col: np.ndarray = None
    a_val: np.ndarray = None
    b_rowptr: np.ndarray = None
    b_col: np.ndarray = None
    b_num: np.ndarray = None
    b_seeds: np.ndarray = None
    secs: dict = field(default_factory=dict)


def ref_run(dna, k: int, lower: int, upper: int, nranks: int = 1, fetch: bool = True, fasta: str = None) -> RefResult:
    """Run the reference's own KmerOps/SharedSeeds code on `nranks` thread-ranks; with `fasta` the reads come from the
    reference's own FastaIndex + getmydna on that file (src/main.cpp:130-139) instead of `dna`."""
    L = ref_lib(k, lower, upper)
    if fasta is not None:
        L.ref_run_fasta.restype = _vp
        L.ref_run_fasta.argtypes = [_c.c_char_p, _c.c_int]
        h = _vp(L.ref_run_fasta(fasta.encode(), nranks))
    else:
        h = _vp(L.ref_run(_p(dna.buf), _p(dna.lengths), _u64(dna.size()), nranks))
    try:
        sz = np.zeros(8, np.int64)
        L.ref_sizes(h, _p(sz))
        r = RefResult(R=int(sz[0]), N=int(sz[1]), nnzA=int(sz[3]), nnzB=int(sz[4]), nnzB_pre=int(sz[5]),
                      keys_after_pass1=int(sz[6]), nranks=int(sz[7]))
        secs = np.zeros(6)
        L.ref_secs(h, _p(secs))
        r.secs = dict(keys=secs[0], values=secs[1], matrix=secs[2], transpose=secs[3], spgemm=secs[4], total=secs[5])
        if fetch:
            r.kmers = np.zeros(r.R, np.uint64)
            r.counts = np.zeros(r.R, np.int32)
            r.reads = np.zeros((r.R, upper), np.int64)
            r.pos = np.zeros((r.R, upper), np.uint32)
            L.ref_get_kmers(h, _p(r.kmers), _p(r.counts), _p(r.reads), _p(r.pos))
            r.a_rowptr = np.zeros(r.N + 1, np.int64)
            r.a_col = np.zeros(r.nnzA, np.int64)
            r.a_val = np.zeros(r.nnzA, np.uint32)
            L.ref_get_A(h, _p(r.a_rowptr), _p(r.a_col), _p(r.a_val))
            r.b_rowptr = np.zeros(r.N + 1, np.int64)
            r.b_col = np.zeros(r.nnzB, np.int64)
            r.b_num = np.zeros(r.nnzB, np.int32)
            r.b_seeds = np.zeros((r.nnzB, 4), np.uint32)
            L.ref_get_B(h, _p(r.b_rowptr), _p(r.b_col), _p(r.b_num), _p(r.b_seeds))
        return r
    finally:
        L.ref_free(h)


def shim_path(k: int, lower: int, upper: int) -> str:
    return os.path.join(HERE, "_ref", f"libelba_shim_k{k}_l{lower}_u{upper}.so")


def shim_run(dna, k: int, lower: int, upper: int, fasta: str = None) -> RefResult:
    """The reference's driver sequence (src/main.cpp:191-282, restated in ref_wrap.cpp::ref_run) over the PRODUCT's drop-in
    shim (elba_b200/host/elba_fe_shim.cpp) and libelba_fe.so, with the reference's own headers: executes the boundary.
    Needs a B200; the library is built where /root/reference exists (oracle/Makefile target `shim`) and travels prebuilt.
    With `fasta` the sequence starts at FastaIndex(fasta).getmydna(), which the shim library's link (--wrap) sends to
    elba_fe_getmydna: the parse runs on the device and the reads stay resident for the counting."""
    key = ("shim", k, lower, upper)
    if key not in _ref_libs:
        L = _c.CDLL(shim_path(k, lower, upper))
        L.ref_run.restype = _vp
        _ref_libs[key] = L
    saved = _ref_libs.get((k, lower, upper))
    _ref_libs[(k, lower, upper)] = _ref_libs[key]
    try:
        return ref_run(dna, k, lower, upper, nranks=1, fasta=fasta)
    finally:
        if saved is None:
            del _ref_libs[(k, lower, upper)]
        else:
            _ref_libs[(k, lower, upper)] = saved


def ref_kmer_info(s: str, lower: int = 2, upper: int = 8):
    L = ref_lib(len(s), lower, upper)
    v = [_u64() for _ in range(4)]
    L.ref_kmer_info(s.encode("ascii"), *[_c.byref(x) for x in v])
    return tuple(int(x.value) for x in v)


def ref_hll(dna, k: int, lower: int = 2, upper: int = 8):
    L = ref_lib(k, lower, upper)
    regs = np.zeros(4096, np.uint8)
    est = L.ref_hll(_p(dna.buf), _p(dna.lengths), _u64(dna.size()), _p(regs))
    return float(est), regs


def ref_rep_kmers(packed: np.ndarray, length: int, k: int, lower: int = 2, upper: int = 8) -> np.ndarray:
    L = ref_lib(k, lower, upper)
    out = np.zeros(max(length - k + 1, 0), np.uint64)
    n = L.ref_rep_kmers(_p(np.ascontiguousarray(packed)), _u64(length), _p(out))
    return out[:n]


def ref_bloom_fill(entries: int, err: float, kmers: np.ndarray, k: int = 17, lower: int = 2, upper: int = 8):
    """Bloom::Add every k-mer into the reference's own filter; returns (bits, hashes, bytes)."""
    L = ref_lib(k, lower, upper)
    b = _vp(L.ref_bloom_new(_i64(entries), _dbl(err)))
    try:
        for x in kmers:
            L.ref_bloom_add(b, _u64(int(x)))
        bits = int(L.ref_bloom_bits(b))
        out = np.zeros(bits // 8 + (1 if bits % 8 else 0), np.uint8)
        L.ref_bloom_copy(b, _p(out))
        return bits, int(L.ref_bloom_hashes(b)), out
    finally:
        L.ref_bloom_free(b)


# ---- the consumer of B: X-drop seed-and-extend of every aligned pair ---------------------------
XDROP_FIELDS = ("begQ", "endQ", "begT", "endT", "score", "rc", "passed", "containedQ", "containedT", "direction", "directionT", "suffix", "suffixT")


def alignment_pairs(b_rowptr: np.ndarray, b_col: np.ndarray, b_seeds: np.ndarray):
    """The nonzeros of B the reference aligns on one rank (src/PairwiseAlignment.cpp:52: strict upper triangle) with the
    seed it uses (seeds[0], :90): rows, cols, seed position in the row read, seed position in the column read."""
    rows = np.repeat(np.arange(len(b_rowptr) - 1, dtype=np.int64), np.diff(b_rowptr))
    cols = np.asarray(b_col, np.int64)
    keep = rows < cols
    s = np.asarray(b_seeds).reshape(-1, 4)
    return rows[keep], cols[keep], np.ascontiguousarray(s[keep, 0], np.uint32), np.ascontiguousarray(s[keep, 1], np.uint32)


def _byte_offsets(dna) -> np.ndarray:
    return np.ascontiguousarray(dna.offsets, np.uint64)


def xdrop(dna, k: int, rows, cols, seedq, seedt, mat: int = 1, mis: int = -1, gap: int = -1, dropoff: int = 15) -> np.ndarray:
    """oracle/xdrop_oracle.cpp: [npairs, 13] int32 in XDROP_FIELDS order (defaults: src/main.cpp:53-56)."""
    L = lib()
    n = len(rows)
    out = np.zeros((n, len(XDROP_FIELDS)), np.int32)
    assert L.elba_oracle_xdrop_fields() == len(XDROP_FIELDS)
    L.elba_oracle_xdrop_batch(_p(dna.buf), _p(_byte_offsets(dna)), _p(np.ascontiguousarray(dna.lengths, np.uint64)), _i32(k),
                              _p(np.ascontiguousarray(rows, np.int64)), _p(np.ascontiguousarray(cols, np.int64)),
                              _p(np.ascontiguousarray(seedq, np.uint32)), _p(np.ascontiguousarray(seedt, np.uint32)), _u64(n),
                              _i32(mat), _i32(mis), _i32(gap), _i32(dropoff), _p(out))
    return out


def ref_xdrop(dna, k: int, lower: int, upper: int, rows, cols, seedq, seedt, mat: int = 1, mis: int = -1, gap: int = -1, dropoff: int = 15) -> np.ndarray:
    """The reference's own XDropAligner.cpp + Overlap.cpp (oracle/_ref) on the same pairs."""
    L = ref_lib(k, lower, upper)
    n = len(rows)
    out = np.zeros((n, len(XDROP_FIELDS)), np.int32)
    L.ref_xdrop_batch(_p(dna.buf), _p(_byte_offsets(dna)), _p(np.ascontiguousarray(dna.lengths, np.uint64)),
                      _p(np.ascontiguousarray(rows, np.int64)), _p(np.ascontiguousarray(cols, np.int64)),
                      _p(np.ascontiguousarray(seedq, np.uint32)), _p(np.ascontiguousarray(seedt, np.uint32)), _u64(n),
                      _i32(mat), _i32(mis), _i32(gap), _i32(dropoff), _p(out))
    return out


# ---- FASTA ingest (src/FastaIndex.cpp:98-176,191-290, src/DnaSeq.cpp:7-29) -------------------------------
def fasta_pack(chunk: bytes, chunk_pos: int, rec: np.ndarray) -> np.ndarray:
    """The restatement: rec = (nreads, 3) uint64 .fai records (len, pos, bases) -> the packed arena of those reads."""
    rec = np.ascontiguousarray(rec, dtype=np.uint64).reshape(-1, 3)
    out = np.zeros(int(((rec[:, 0] + np.uint64(3)) // np.uint64(4)).sum()) if len(rec) else 0, np.uint8)
    L = lib()
    L.eo_fasta_pack.restype = _i64
    L.eo_fasta_pack.argtypes = [_c.c_char_p, _u64, _u64, _vp, _u64, _vp]
    n = L.eo_fasta_pack(chunk, len(chunk), chunk_pos, _p(rec), len(rec), _p(out))
    if n < 0:
        raise ValueError("a record leaves the chunk")
    assert n == out.size
    return out


def ref_fasta(path: str, nranks: int = 1, klu=(17, 2, 8)):
    """The reference's own FastaIndex ctor + getmydna on `nranks` thread-ranks.  Returns (displ [nranks + 1],
    [(records (n, 3) uint64, packed arena uint8)] per rank)."""
    L = ref_lib(*klu)
    L.ref_fasta_run.restype = _vp
    L.ref_fasta_run.argtypes = [_c.c_char_p, _c.c_int]
    L.ref_fasta_sizes.argtypes = [_vp, _vp, _vp]
    L.ref_fasta_get.argtypes = [_vp, _c.c_int, _vp, _vp]
    L.ref_fasta_free.argtypes = [_vp]
    h = L.ref_fasta_run(path.encode(), nranks)
    try:
        displ = np.zeros(nranks + 1, np.int64)
        nbytes = np.zeros(nranks, np.int64)
        L.ref_fasta_sizes(h, _p(displ), _p(nbytes))
        out = []
        for r in range(nranks):
            rec = np.zeros((int(displ[r + 1] - displ[r]) if r + 1 < nranks else int(displ[nranks] - displ[r]), 3), np.uint64)
            buf = np.zeros(int(nbytes[r]), np.uint8)
            L.ref_fasta_get(h, r, _p(rec), _p(buf))
            out.append((rec, buf))
        return displ, out
    finally:
        L.ref_fasta_free(h)


# ---- transitive reduction of the overlap graph (src/TransitiveReduction.cpp:3-92) -----------------------
TR_FIELDS = ("direction", "directionT", "suffix", "suffixT")


def transitive_reduction(n: int, rows, cols, fields, fuzz: int = 1000):
    """The restatement: R as triples (row, col, [direction, directionT, suffix, suffixT]) -> the string graph S as
    (row, col, fields[nnzS, 4], src, transposed), row-major.  FUZZ = 1000: include/TransitiveReduction.hpp:15."""
    rows, cols = np.ascontiguousarray(rows, np.int64), np.ascontiguousarray(cols, np.int64)
    fields = np.ascontiguousarray(fields, np.int32).reshape(-1, 4)
    L = lib()
    L.eo_transitive_reduction.restype = _u64
    L.eo_transitive_reduction.argtypes = [_i64, _u64, _vp, _vp, _vp, _i32, _vp, _vp, _vp, _vp, _vp]
    cap = 2 * len(rows)
    orow, ocol, of = np.zeros(cap, np.int64), np.zeros(cap, np.int64), np.zeros((cap, 4), np.int32)
    osrc, otr = np.zeros(cap, np.uint64), np.zeros(cap, np.uint8)
    m = L.eo_transitive_reduction(n, len(rows), _p(rows), _p(cols), _p(fields), fuzz, _p(orow), _p(ocol), _p(of), _p(osrc), _p(otr))
    return orow[:m], ocol[:m], of[:m], osrc[:m], otr[:m]


def ref_transitive_reduction(n: int, rows, cols, fields, klu=(17, 2, 8), shim: bool = False):
    """The reference's own TransitiveReduction.cpp (oracle/_ref; CombBLAS restated in oracle/stubs): (row, col, fields[nnzS, 4]).
    shim=True: the same driver code over the PRODUCT's drop-in TransitiveReduction (elba_b200/host/elba_fe_shim.cpp ->
    elba_fe_transitive_reduction; needs a B200)."""
    rows, cols = np.ascontiguousarray(rows, np.int64), np.ascontiguousarray(cols, np.int64)
    fields = np.ascontiguousarray(fields, np.int32).reshape(-1, 4)
    if shim:
        key = ("shim",) + tuple(klu)
        if key not in _ref_libs:
            _ref_libs[key] = _c.CDLL(shim_path(*klu))
        L = _ref_libs[key]
    else:
        L = ref_lib(*klu)
    L.ref_transitive_reduction.restype = _vp
    L.ref_transitive_reduction.argtypes = [_i64, _u64, _vp, _vp, _vp]
    L.ref_tr_size.restype = _u64
    L.ref_tr_size.argtypes = [_vp]
    L.ref_tr_get.argtypes = [_vp, _vp, _vp, _vp]
    L.ref_tr_free.argtypes = [_vp]
    h = L.ref_transitive_reduction(n, len(rows), _p(rows), _p(cols), _p(fields))
    try:
        m = L.ref_tr_size(h)
        orow, ocol, of = np.zeros(m, np.int64), np.zeros(m, np.int64), np.zeros((m, 4), np.int32)
        L.ref_tr_get(h, _p(orow), _p(ocol), _p(of))
        return orow, ocol, of
    finally:
        L.ref_tr_free(h)
