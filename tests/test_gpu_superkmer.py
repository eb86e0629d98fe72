"""The super-k-mer counting path (elba_b200/csrc/superkmer.cuh, k >= 20) against the CPU oracle: every minimizer
geometry, both overflow fallbacks (record capacity, BUCKET_CAP instances), skewed minimizers, and the hash path
forced on the same inputs (both paths must give the same bits)."""
import os

import numpy as np
import pytest

from test_gpu_parity import _compare, _run_cuda

pytestmark = pytest.mark.gpu


class _env:
    def __init__(self, **kw):
        self.kw, self.old = kw, {}

    def __enter__(self):
        for a, b in self.kw.items():
            self.old[a] = os.environ.get(a)
            os.environ[a] = str(b)

    def __exit__(self, *exc):
        for a, b in self.old.items():
            if b is None:
                os.environ.pop(a, None)
            else:
                os.environ[a] = b


@pytest.fixture(scope="module")
def hifi():
    from elba_b200.synth import make_dnabuffer
    return make_dnabuffer(genome_len=200_000, n_reads=300, mean_len=12000, sd_len=1500, err=0.01, seed=17, repeat_frac=0.03)


@pytest.mark.parametrize("k", [20, 21, 23, 24, 25, 27, 28, 29, 31, 32])
def test_every_minimizer_geometry(hifi, k):
    """W = 8 (k 20-23), 12 (24-27), 16 (28-31), 17 (k = 32); m = k - W + 1 in 13..16."""
    from oracle import oracle as O
    ref = O.run(hifi, k, 2, 6)
    out = _run_cuda(hifi, k, 2, 6)
    _compare(out, ref, f"skm k={k}")
    assert out["sizes"]["table_slots"] == 2048 and out["sizes"]["partitions"] > 100      # the bucket path ran


def test_hash_path_forced_gives_the_same_bits(hifi):
    from oracle import oracle as O
    ref = O.run(hifi, 31, 2, 4)
    with _env(ELBA_FE_COUNT_PATH="hash"):
        out = _run_cuda(hifi, 31, 2, 4)
    _compare(out, ref, "hash path k=31")
    out2 = _run_cuda(hifi, 31, 2, 4)
    _compare(out2, ref, "skm path k=31")
    assert out2["sizes"]["partitions"] != out["sizes"]["partitions"]


def test_record_capacity_overflow(hifi):
    """Slack 1.0: about half of the buckets are offered more records than they hold -> those buckets are counted whole by the
    global-table kernel; nothing may change."""
    from oracle import oracle as O
    ref = O.run(hifi, 31, 2, 4)
    with _env(ELBA_FE_SKM_SLACK="1.0"):
        out = _run_cuda(hifi, 31, 2, 4)
    _compare(out, ref, "record overflow")
    assert out["sizes"]["overflow_instances"] > 100_000


def test_instance_overflow(hifi):
    """Mean fill = BUCKET_CAP: half of the buckets exceed BUCKET_CAP instances (spilled by k_skm_count)."""
    from oracle import oracle as O
    ref = O.run(hifi, 29, 2, 4)
    with _env(ELBA_FE_SKM_MEAN="6144", ELBA_FE_SKM_SLACK="4.0"):
        out = _run_cuda(hifi, 29, 2, 4)
    _compare(out, ref, "instance overflow")
    assert out["sizes"]["overflow_instances"] > 100_000


def test_skewed_minimizers():
    """Poly-A and dinucleotide reads: one minimizer, runs of the maximum record length, one bucket far over capacity."""
    from elba_b200.dnabuffer import DnaBuffer
    from elba_b200.synth import make_dnabuffer
    from oracle import oracle as O
    base = make_dnabuffer(genome_len=40_000, n_reads=120, mean_len=5000, sd_len=500, err=0.02, seed=21)
    seqs = [base.read_ascii(i) for i in range(base.size())] + ["A" * 20000] * 30 + ["AC" * 6000] * 3 + ["T" * 3000] * 2
    dna = DnaBuffer.from_strings(seqs)
    for k in (31, 32, 21):
        ref = O.run(dna, k, 2, 8)
        out = _run_cuda(dna, k, 2, 8, parts=16)
        _compare(out, ref, f"skewed minimizers k={k}")
        assert out["sizes"]["overflow_instances"] >= 500_000


def test_ragged_reads_k31():
    """Reads around k and around the 32-start chunk size; reads shorter than k contribute nothing."""
    from elba_b200.dnabuffer import DnaBuffer
    from oracle import oracle as O
    rng = np.random.default_rng(9)
    genome = "".join("ACGT"[c] for c in rng.integers(0, 4, 4000))
    lens = [0, 1, 30, 31, 32, 33, 61, 62, 63, 64, 65, 93, 94, 95, 127, 128, 129, 500, 1000, 2, 700, 31, 900, 62, 3000]
    seqs = [genome[(37 * i) % 900:(37 * i) % 900 + l] for i, l in enumerate(lens)]
    dna = DnaBuffer.from_strings(seqs * 3)
    for k in (31, 32, 24):
        ref = O.run(dna, k, 2, 8)
        out = _run_cuda(dna, k, 2, 8, parts=5)
        _compare(out, ref, f"ragged k={k}")


def test_fused_seed_list(hifi):
    """Pass 2 is fused into counting: build_A sees exactly nnzA_pre seeds plus the holes of the chunked list (no filter, no false positives)."""
    from oracle import oracle as O
    ref = O.run(hifi, 31, 2, 4)
    out = _run_cuda(hifi, 31, 2, 4)
    _compare(out, ref, "fused")
    assert ref.nnzA_pre <= out["sizes"]["candidates"] <= ref.nnzA_pre + 592 * 16384


def test_table_too_full_spills_the_bucket():
    """Noisy reads (15 % errors): nearly every k-mer instance is distinct, so buckets sized for HiFi data hold more distinct
    k-mers than 3/4 of the 2048-slot table: the kernel notices while inserting and hands those buckets, whole, to the exact
    global-table fallback.  Nothing may change."""
    from elba_b200.synth import make_dnabuffer
    from oracle import oracle as O
    dna = make_dnabuffer(genome_len=150_000, n_reads=250, mean_len=9000, sd_len=1000, err=0.15, seed=23)
    for k, mean in ((31, 3500), (25, 2400), (21, 6000)):
        ref = O.run(dna, k, 2, 8)
        assert ref.D > 0.8 * ref.M
        with _env(ELBA_FE_SKM_MEAN=mean, ELBA_FE_SKM_SLACK="4.0"):
            out = _run_cuda(dna, k, 2, 8)
        _compare(out, ref, f"full tables k={k}")
        assert out["sizes"]["overflow_instances"] > 200_000
    # with the default geometry (sized for a distinct ratio of 0.3 on the first pass) the result is the same
    _compare(_run_cuda(dna, 31, 2, 8), O.run(dna, 31, 2, 8), "noisy reads, default geometry")


def test_fused_seed_list_resize_and_overflow_buckets(hifi):
    """A seed list that is too small is resized to its exact size and the count redone; buckets that spill to the
    global table emit their seeds from there (k_skm_emit_global).  Most instances are reliable here (L=2, U=60)."""
    from oracle import oracle as O
    ref = O.run(hifi, 31, 2, 60)
    assert ref.nnzA_pre > 1_000_000
    with _env(ELBA_FE_SEED_CAP="1000"):
        out = _run_cuda(hifi, 31, 2, 60)
    _compare(out, ref, "seed list resize")
    with _env(ELBA_FE_SKM_SLACK="1.0", ELBA_FE_SEED_CAP="5000"):
        out = _run_cuda(hifi, 31, 2, 60)
    _compare(out, ref, "seed list resize + record overflow")
    assert out["sizes"]["overflow_instances"] > 100_000
