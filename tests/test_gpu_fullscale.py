"""Parity at the sizes the benchmark times (BASELINE.json configs 3 and 4), and the device-side result digests that let
bench.py prove what it timed.

* elba_fe_digests == the same arithmetic in numpy over the oracle's results (tests/common.py).
* config 3 (synthetic E. coli 30X CLR, 16,890 reads, k=17 U=8) IN FULL against the CPU oracle.
* config 4 (C. elegans 40X HiFi shape, k=31 U=4) at 5 % scale against the CPU oracle, and at FULL scale through the
  reference's own seed-validity property (test.py:40-65) on a 20 k sample plus the committed digests of
  tests/golden/bench_digests.json (what bench.py compares its timed result with).
* stride != 1 (the legacy -s flag) against the oracle.
"""
import json
import os

import numpy as np
import pytest

from common import check_seeds_valid, oracle_result_digests, result_digests

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _synthetic(shape: str, scale: float = 1.0):
    """The benchmark's own input (bench.py load_workload): generated on the device, returned as device tensors + DnaBuffer."""
    import torch
    from elba_b200 import synth
    s = synth.SHAPES[shape]
    genome, reads = int(s["genome"] * scale), int(s["reads"] * scale)
    dev = torch.device("cuda", 0)
    buf, off, lens = synth.make_reads_block(genome, reads, s["mean"], s["sd"], s["err"], 313, dev, 0, reads)
    torch.cuda.synchronize()          # the library runs on its own stream: the generated reads must be complete
    return (buf, off, lens), s


def _run_device(tensors, k, lo, up, stride=1):
    from elba_b200 import frontend
    buf, off, lens = tensors
    ctx = frontend.Context(frontend.Params(k=k, lower=lo, upper=up, stride=stride, device=0))
    ctx.set_reads_device(buf.data_ptr(), buf.numel(), off.data_ptr(), lens.data_ptr(), lens.numel(), 0)
    ctx.run()
    return ctx


def _compare_all(ctx, ref, what):
    s = ctx.sizes()
    got = (s["num_kmers"], s["distinct"], s["reliable"], s["nnzA_pre"], s["nnzA"], s["products"], s["nnzB_pre"], s["nnzB"])
    want = (ref.M, ref.D, ref.R, ref.nnzA_pre, ref.nnzA, ref.F, ref.nnzB_pre, ref.nnzB)
    assert got == want, (what, got, want)
    kmers, counts = ctx.kmers()
    assert np.array_equal(kmers, ref.kmers) and np.array_equal(counts, ref.counts), what + ": reliable k-mers"
    rp, col, pos = ctx.A()
    assert np.array_equal(rp, ref.a_rowptr) and np.array_equal(col, ref.a_col) and np.array_equal(pos, ref.a_pos), what + ": A"
    brp, bcol, bnum, bseeds = ctx.B()
    assert np.array_equal(brp, ref.b_rowptr) and np.array_equal(bcol, ref.b_col) and np.array_equal(bnum, ref.b_num), what + ": B"
    assert np.array_equal(bseeds, ref.b_seeds), what + ": seeds"
    assert ctx.digests() == oracle_result_digests(ref), what + ": device digests"


def test_digests_match_numpy_restatement(fixtures):
    from elba_b200 import frontend
    from oracle import oracle as O
    for name, k, lo, up in (("reads_fa", 17, 2, 8), ("reads_fa", 31, 2, 4)):
        dna = fixtures(name)
        ctx = frontend.Context(frontend.Params(k=k, lower=lo, upper=up))
        ctx.upload(dna)
        ctx.run()
        ref = O.run(dna, k, lo, up)
        assert ctx.digests() == oracle_result_digests(ref)
        # and from the downloaded arrays themselves
        kmers, counts = ctx.kmers()
        assert ctx.digests() == result_digests(kmers, counts, *ctx.A(), *ctx.B())
        ctx.close()


def test_config3_ecoli30x_full_vs_oracle():
    from elba_b200 import synth
    from oracle import oracle as O
    tensors, s = _synthetic("ecoli30x_clr")
    ctx = _run_device(tensors, s["k"], s["lower"], s["upper"])
    dna = synth.to_dnabuffer(*tensors)
    ref = O.run(dna, s["k"], s["lower"], s["upper"], threads=min(16, os.cpu_count() or 1))
    _compare_all(ctx, ref, "config 3 full")
    brp, bcol, _, bseeds = ctx.B()
    assert check_seeds_valid(dna, s["k"], brp, bcol, bseeds, max_checks=5000) == 0
    ctx.close()


def test_config4_celegans_5pct_vs_oracle():
    from elba_b200 import synth
    from oracle import oracle as O
    tensors, s = _synthetic("celegans40x_hifi", 0.05)
    ctx = _run_device(tensors, s["k"], s["lower"], s["upper"])
    dna = synth.to_dnabuffer(*tensors)
    ref = O.run(dna, s["k"], s["lower"], s["upper"], threads=min(16, os.cpu_count() or 1))
    _compare_all(ctx, ref, "config 4 at 5 %")
    ctx.close()


def test_config4_celegans_full_seeds_valid_and_golden_digests():
    from elba_b200 import synth
    tensors, s = _synthetic("celegans40x_hifi")
    ctx = _run_device(tensors, s["k"], s["lower"], s["upper"])
    sz = ctx.sizes()
    assert sz["nreads"] == s["reads"]
    with open(os.path.join(ROOT, "tests", "golden", "bench_digests.json")) as f:
        gold = json.load(f)["celegans40x_hifi@1"]
    assert ctx.digests() == {a: gold[a] for a in ("kmers", "A", "B", "seeds")}
    for a in ("reliable", "nnzA", "products", "nnzB"):
        assert sz[a] == gold["sizes"][a], a
    brp, bcol, bnum, bseeds = ctx.B()
    ctx.close()
    assert (bnum >= 2).all()
    dna = synth.to_dnabuffer(*tensors)
    assert check_seeds_valid(dna, s["k"], brp, bcol, bseeds, max_checks=20000) == 0


def test_sliced_upload_from_host_gives_the_same_bits():
    """A large arena uploaded from host memory travels in slices on a second stream and the scatter follows slice by slice
    (elba_fe_upload_reads); the result must equal the run on reads already resident in HBM.  Both counting paths."""
    import torch
    from elba_b200 import frontend, synth
    tensors, s = _synthetic("celegans40x_hifi", 0.08)
    ctx = _run_device(tensors, s["k"], s["lower"], s["upper"])
    want, want_sizes = ctx.digests(), ctx.sizes()
    ctx.close()
    dna = synth.to_dnabuffer(*tensors)
    assert dna.buf.nbytes >= 64 << 20, "the test must be large enough for the sliced upload"
    hbuf, hoff, hlen = (torch.from_numpy(a).pin_memory() for a in (dna.buf, dna.offsets.astype(np.int64), dna.lengths.astype(np.int64)))
    for env in ({}, {"ELBA_FE_COUNT_PATH": "hash"}):
        old = {a: os.environ.get(a) for a in env}
        os.environ.update(env)
        try:
            ctx = frontend.Context(frontend.Params(k=s["k"], lower=s["lower"], upper=s["upper"], device=0))
            for _ in range(2):                      # twice: the second upload overwrites the arena of the first pass
                ctx.upload_raw(hbuf.data_ptr(), hbuf.numel(), hoff.data_ptr(), hlen.data_ptr(), dna.size(), 0)
                ctx.run()
                assert ctx.digests() == want, env
            got = ctx.sizes()
            assert all(got[a] == want_sizes[a] for a in ("num_kmers", "distinct", "reliable", "nnzA", "products", "nnzB")), env
            ctx.close()
        finally:
            for a, b in old.items():
                if b is None:
                    os.environ.pop(a, None)
                else:
                    os.environ[a] = b


@pytest.mark.parametrize("k,stride", [(17, 3), (31, 2), (21, 5)])
def test_stride_vs_oracle(k, stride):
    """-s: only window starts p with p % stride == 0 (README.md:85); the reference hard-wires 1, the oracle restates the rule."""
    from elba_b200 import frontend
    from elba_b200.synth import make_dnabuffer
    from oracle import oracle as O
    dna = make_dnabuffer(genome_len=60_000, n_reads=180, mean_len=5000, sd_len=600, err=0.04, seed=9)
    ref = O.run(dna, k, 2, 8, stride=stride)
    ctx = frontend.Context(frontend.Params(k=k, lower=2, upper=8, stride=stride))
    ctx.upload(dna)
    ctx.run()
    _compare_all(ctx, ref, f"k={k} stride={stride}")
    ctx.close()
