"""Parity tests proper: the CUDA path, called through the C ABI, against the CPU oracle on the same inputs,
and against the committed golden digests produced by the reference's own code (tests/golden/golden.json).

Bar: bit-exact (integer / byte / index work).  Seeds follow the canonical rule both sides implement
(DESIGN.md "Seed rule"; the reference leaves the choice to CombBLAS internals -> parity unpinned for seeds,
pinned for everything else) and every retained seed must pass the reference's own validity property (test.py).
"""
import numpy as np
import pytest

from common import check_seeds_valid, digest, oracle_digests

pytestmark = pytest.mark.gpu


def _run_cuda(dna, k, lower, upper, parts=0, stride=1, seed_count=2):
    from elba_b200 import frontend
    p = frontend.Params(k=k, lower=lower, upper=upper, stride=stride, seed_count=seed_count, device=0, num_partitions=parts)
    kmermap, A, B = frontend.overlap_front_end(dna, p)
    ctx = kmermap.ctx
    out = dict(sizes=ctx.sizes(), kmers=kmermap.items(), A=A.csr(), AT=A.transpose().csr(), B=B.csr(), timings=ctx.timings())
    ctx.close()
    return out


def _compare(out, ref, what=""):
    s = out["sizes"]
    assert s["num_kmers"] == ref.M, what
    assert s["distinct"] == ref.D, what
    assert s["reliable"] == ref.R, what
    assert s["nnzA_pre"] == ref.nnzA_pre, what
    assert s["nnzA"] == ref.nnzA, what
    assert s["products"] == ref.F, what
    assert s["nnzB_pre"] == ref.nnzB_pre, what
    assert s["nnzB"] == ref.nnzB, what
    kmers, counts = out["kmers"]
    assert np.array_equal(kmers, ref.kmers), what + " reliable k-mer set"
    assert np.array_equal(counts, ref.counts), what + " counts"
    rp, col, pos = out["A"]
    assert np.array_equal(rp, ref.a_rowptr) and np.array_equal(col, ref.a_col) and np.array_equal(pos, ref.a_pos), what + " A"
    cp, row, tpos = out["AT"]
    assert np.array_equal(cp, ref.at_colptr) and np.array_equal(row, ref.at_row) and np.array_equal(tpos, ref.at_pos), what + " A^T"
    brp, bcol, bnum, bseeds = out["B"]
    assert np.array_equal(brp, ref.b_rowptr) and np.array_equal(bcol, ref.b_col), what + " pattern of B"
    assert np.array_equal(bnum, ref.b_num), what + " numshared"
    assert np.array_equal(bseeds, ref.b_seeds), what + " seeds"


def test_kmer_stream_matches_oracle(fixtures):
    """Kernel 1 (2-bit parse / roll / canonicalise) == TKmer::GetRepKmers stream (src/Kmer.cpp:215-242)."""
    from elba_b200 import frontend
    from oracle import oracle as O
    dna = fixtures("reads_fa").slice(0, 40)
    for k in (17, 31, 15, 32, 5):
        ctx = frontend.Context(frontend.Params(k=k, lower=2, upper=8))
        ctx.upload(dna)
        got = ctx.kmer_stream()
        ctx.close()
        assert np.array_equal(got, O.all_rep_kmers(dna, k)), f"k={k}"


def test_ragged_and_empty_inputs():
    """Reads shorter than k keep their row but contribute nothing (include/KmerOps.hpp:118-119); empty input."""
    from elba_b200.dnabuffer import DnaBuffer
    from oracle import oracle as O
    rng = np.random.default_rng(5)
    genome = "".join("ACGT"[c] for c in rng.integers(0, 4, 3000))
    lens = [0, 1, 16, 17, 18, 31, 32, 33, 63, 64, 65, 127, 128, 129, 500, 1000, 2, 700, 17, 900]
    seqs = [genome[(37 * i) % 1500:(37 * i) % 1500 + l] for i, l in enumerate(lens)]
    dna = DnaBuffer.from_strings(seqs)
    for parts in (1, 3):
        out = _run_cuda(dna, 17, 2, 8, parts=parts)
        _compare(out, O.run(dna, 17, 2, 8), f"ragged parts={parts}")
    empty = DnaBuffer.from_strings([])
    out = _run_cuda(empty, 17, 2, 8)
    assert out["sizes"]["nnzB"] == 0 and out["sizes"]["reliable"] == 0
    short = DnaBuffer.from_strings(["ACGT", "A", ""])
    out = _run_cuda(short, 17, 2, 8)
    assert out["sizes"]["num_kmers"] == 0 and list(out["B"][0]) == [0, 0, 0, 0]


@pytest.mark.parametrize("key", ["reads_fa_k17_l2_u8", "reads_fa_first135_k17_l2_u8", "reads_fa_k31_l2_u4", "reads_fa_k31_l15_u35",
                                 "example_medium_k17_l2_u8", "example_medium_k31_l2_u4"])
@pytest.mark.parametrize("parts", [0, 1])
def test_reference_fixtures(fixtures, golden, key, parts):
    """BASELINE configs 1 (stand-in: first 135 reads of reads.fa) and 2 (example_medium k=17) plus the other
    (k,L,U) rows of SURVEY.md §8, against the oracle AND the digests of the reference's own run."""
    from oracle import oracle as O
    g = golden["configs"][key]
    dna = fixtures(g["fixture"])
    out = _run_cuda(dna, g["k"], g["lower"], g["upper"], parts=parts)
    ref = O.run(dna, g["k"], g["lower"], g["upper"])
    _compare(out, ref, key)
    s = out["sizes"]
    for name in ("R", "nnzA", "nnzB_pre", "nnzB"):
        assert s[{"R": "reliable"}.get(name, name)] == g[name], (key, name)
    kmers, counts = out["kmers"]
    rp, col, pos = out["A"]
    brp, bcol, bnum, bseeds = out["B"]
    assert digest(kmers, counts) == g["digests"]["kmers"]
    assert digest(rp, col, pos) == g["digests"]["A"]
    assert digest(brp, bcol, bnum) == g["digests"]["B"]
    rows = np.repeat(np.arange(len(brp) - 1), np.diff(brp))
    assert int((rows == bcol).sum()) == g["diag"] and int((rows < bcol).sum()) == g["strict_upper"]
    assert check_seeds_valid(dna, g["k"], brp, bcol, bseeds, max_checks=3000) == 0


def test_example_medium_default_bounds(fixtures, golden):
    """example_medium at the reference's build defaults K=31 L=15 U=35 (Makefile:1-3): F = 636 M products."""
    from oracle import oracle as O
    g = golden["configs"]["example_medium_k31_l15_u35"]
    dna = fixtures("example_medium")
    out = _run_cuda(dna, 31, 15, 35)
    ref = O.run(dna, 31, 15, 35, threads=8)
    _compare(out, ref, "example_medium 31/15/35")
    brp, bcol, bnum, _ = out["B"]
    assert digest(brp, bcol, bnum) == g["digests"]["B"]
    assert out["sizes"]["products"] == 636200657


@pytest.mark.parametrize("shape", [dict(genome_len=150_000, n_reads=500, mean_len=8000, sd_len=1000, err=0.15, k=17, lower=2, upper=8),
                                   dict(genome_len=300_000, n_reads=400, mean_len=14000, sd_len=1000, err=0.01, k=31, lower=2, upper=4),
                                   dict(genome_len=120_000, n_reads=600, mean_len=7000, sd_len=1000, err=0.15, k=17, lower=2, upper=4)])
def test_synthetic_shapes_small(shape):
    """Configs 3-5 (E. coli CLR, C. elegans HiFi, human CLR) at sizes the oracle finishes in seconds."""
    from elba_b200.synth import make_dnabuffer
    from oracle import oracle as O
    shape = dict(shape)
    k, lo, up = shape.pop("k"), shape.pop("lower"), shape.pop("upper")
    dna = make_dnabuffer(seed=11, repeat_frac=0.05, n_frac=0.0001, **shape)
    ref = O.run(dna, k, lo, up)
    for parts in (0, 1, 7):
        _compare(_run_cuda(dna, k, lo, up, parts=parts), ref, f"{shape} parts={parts}")


def test_heavy_rows_and_hot_kmers():
    """A repeat family shared by many reads: hot k-mers far above UPPER (count saturation) and rows whose
    distinct-column count overflows the shared-memory SpGEMM table (global-memory fallback)."""
    from elba_b200.dnabuffer import DnaBuffer
    from oracle import oracle as O
    rng = np.random.default_rng(3)
    core = rng.integers(0, 4, 150)
    seqs = []
    for i in range(1800):
        flank_l, flank_r = rng.integers(0, 4, 60), rng.integers(0, 4, 60)
        c = core.copy()
        # every read carries 3 private substitutions so that some core k-mers stay <= UPPER among subsets
        c[rng.integers(0, 150, 3)] = rng.integers(0, 4, 3)
        seqs.append("".join("ACGT"[x] for x in np.concatenate([flank_l, c, flank_r])))
    seqs.append("A" * 5000)                       # poly-A: one k-mer, thousands of instances
    dna = DnaBuffer.from_strings(seqs)
    ref = O.run(dna, 17, 2, 4000)
    out = _run_cuda(dna, 17, 2, 4000)
    _compare(out, ref, "heavy rows")
    assert ref.nnzB > 1800 * 1600, "test must overflow the 2048-slot table"


def test_spgemm_in_rounds(fixtures):
    """The products of A (x) A^T are generated as 16-byte tuples grouped by output row; when they do not fit the budget the rows
    are multiplied in consecutive ranges (ELBA_FE_TUPLE_MB).  Same bits, also for rows that overflow the shared-memory table."""
    import os
    from elba_b200.dnabuffer import DnaBuffer
    from oracle import oracle as O
    old = os.environ.get("ELBA_FE_TUPLE_MB")
    os.environ["ELBA_FE_TUPLE_MB"] = "1"
    try:
        dna = fixtures("reads_fa")
        _compare(_run_cuda(dna, 17, 2, 8), O.run(dna, 17, 2, 8), "rounds, reads.fa")
        rng = np.random.default_rng(3)
        core = rng.integers(0, 4, 120)
        seqs = []
        for i in range(700):
            c = core.copy()
            c[rng.integers(0, 120, 3)] = rng.integers(0, 4, 3)
            seqs.append("".join("ACGT"[x] for x in np.concatenate([rng.integers(0, 4, 50), c, rng.integers(0, 4, 50)])))
        dna = DnaBuffer.from_strings(seqs)
        _compare(_run_cuda(dna, 17, 2, 2000), O.run(dna, 17, 2, 2000), "rounds, heavy rows")
    finally:
        if old is None:
            os.environ.pop("ELBA_FE_TUPLE_MB", None)
        else:
            os.environ["ELBA_FE_TUPLE_MB"] = old


def test_skewed_partitions():
    """Heavy hitters: poly-A reads put > 1 M instances of ONE k-mer into one level-1 partition.  The optimistic
    partition layout overflows (exact-histogram layout), then that partition's sub-bucket overflows (global-table
    recount).  Results must not change."""
    from elba_b200.dnabuffer import DnaBuffer
    from elba_b200.synth import make_dnabuffer
    from oracle import oracle as O
    base = make_dnabuffer(genome_len=40_000, n_reads=120, mean_len=5000, sd_len=500, err=0.08, seed=21)
    seqs = [base.read_ascii(i) for i in range(base.size())] + ["A" * 20000] * 60 + ["AC" * 6000] * 3
    dna = DnaBuffer.from_strings(seqs)
    ref = O.run(dna, 17, 2, 8)
    for parts in (0, 8, 64):
        out = _run_cuda(dna, 17, 2, 8, parts=parts)
        _compare(out, ref, f"skew parts={parts}")
        assert out["sizes"]["overflow_instances"] >= 1_000_000


def test_stride_and_single_seed():
    from elba_b200.synth import make_dnabuffer
    from oracle import oracle as O
    dna = make_dnabuffer(genome_len=50_000, n_reads=150, mean_len=5000, sd_len=500, err=0.05, seed=4)
    ref = O.run(dna, 17, 2, 8)
    out = _run_cuda(dna, 17, 2, 8, seed_count=1)
    brp, bcol, bnum, bseeds = out["B"]
    assert np.array_equal(bnum, ref.b_num) and np.array_equal(bseeds[:, :2], ref.b_seeds[:, :2]) and not bseeds[:, 2:].any()
    # stride 3: only window starts p % 3 == 0 are visited (legacy -s flag; README.md:80-98)
    out3 = _run_cuda(dna, 17, 2, 8, stride=3)
    M3 = int(sum((max(int(l) - 17 + 1, 0) + 2) // 3 for l in dna.lengths))
    assert out3["sizes"]["num_kmers"] == M3
    _, _, pos = out3["A"]
    assert (pos % 3 == 0).all()


def test_hll_and_bloom_bit_exact(fixtures, golden):
    """a9 / a10: the reference's sizing sketches on the device (src/HyperLogLog.cpp, src/Bloom.cpp)."""
    from elba_b200 import frontend
    from oracle import oracle as O
    dna = fixtures("reads_fa")
    for k in (17, 31):
        ctx = frontend.Context(frontend.Params(k=k, lower=2, upper=8))
        ctx.upload(dna)
        est, regs = ctx.hll()
        g = golden["values"][f"hll_reads_fa_k{k}"]
        assert digest(regs) == g["registers_sha256"]
        assert est == g["estimate"]                       # 283,869.98... for k=17
        oest, oregs = O.hll(dna, k)
        assert np.array_equal(regs, oregs) and est == oest
        if k == 17:
            entries = int(np.ceil(est))
            bits, hashes, bf = ctx.bloom(entries, 0.05)
            assert (bits, hashes) == (1769993, 5)
            obf = O.bloom_fill(entries, 0.05, O.all_rep_kmers(dna, k))
            assert np.array_equal(bf, obf)
        ctx.close()


def test_error_behaviour():
    from elba_b200 import frontend
    with pytest.raises(frontend.FrontEndError):
        frontend.Context(frontend.Params(k=33, lower=2, upper=8))
    with pytest.raises(frontend.FrontEndError):
        frontend.Context(frontend.Params(k=17, lower=1, upper=8))
    with pytest.raises(frontend.FrontEndError):
        frontend.Context(frontend.Params(k=17, lower=5, upper=4))
    ctx = frontend.Context(frontend.Params(k=17, lower=2, upper=8))
    with pytest.raises(frontend.FrontEndError):
        ctx.count()                                       # no reads yet
    ctx.close()
