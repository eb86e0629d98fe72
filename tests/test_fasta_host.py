"""FASTA ingest, CPU side: the oracle's restatement (oracle/elba_oracle.cpp::eo_fasta_pack) against the REFERENCE'S
OWN FastaIndex.cpp + DnaSeq.cpp + DnaBuffer.cpp (oracle/_ref, compiled unmodified, MPI-IO over pread), and the host mirror
elba_b200/fasta.py (index parse, getpartition, chunk extents) against the same."""
import os

import numpy as np
import pytest

from elba_b200 import fasta as F
from elba_b200.dnabuffer import DnaBuffer
from oracle import oracle as O

KLU = (17, 2, 8)
needs_ref = pytest.mark.skipif(not O.ref_available(*KLU), reason="oracle/_ref not built")


def synth_reads(rng, n, lo, hi, alphabet="ACGT"):
    letters = np.frombuffer(alphabet.encode(), dtype=np.uint8)
    return ["".join(map(chr, letters[rng.integers(0, len(letters), int(l))])) for l in rng.integers(lo, hi, n)]


def oracle_ingest(path, rank, nranks):
    idx = F.FastaIndex(path, rank, nranks)
    start, end = idx.chunk_extent()
    raw = open(path, "rb").read()[start:end]
    return idx, O.fasta_pack(raw, start, idx.getmyrecords())


@needs_ref
@pytest.mark.parametrize("width", [1, 3, 4, 7, 60, 80, 100000])
def test_restatement_equals_reference_fastaindex(tmp_path, width):
    rng = np.random.default_rng(width)
    seqs = synth_reads(rng, 40, 1, 900, "ACGTacgtNn") + ["A", "AC", "ACG", "ACGT", "ACGTA"]
    path = str(tmp_path / "r.fa")
    F.write_fasta(path, seqs, width)
    for P in (1, 3):
        displ, parts = O.ref_fasta(path, P, KLU)
        for r in range(P):
            idx, got = oracle_ingest(path, r, P)
            rec, buf = parts[r]
            assert idx.getmyreaddispl() == displ[r] and idx.gettotrecords() == displ[P]
            assert np.array_equal(idx.getmyrecords(), rec), "records / partition"
            assert np.array_equal(got, buf), f"arena differs from the reference (width {width}, rank {r} of {P})"


@needs_ref
def test_characters_outside_the_table_spill_like_the_reference(tmp_path):
    """include/DnaSeq.hpp:136-154 gives code 4 to anything but ACGTN; src/DnaSeq.cpp:20-22 then ORs uint8_t(4 << (6 - 2i)) into the byte."""
    seqs = ["ACGTRYKM", "XACGT", "AXCGT", "ACXGT", "ACGXT", "acgu-*.t", "NNNNRNNNN"]
    path = str(tmp_path / "odd.fa")
    F.write_fasta(path, seqs, 5)
    _, parts = O.ref_fasta(path, 1, KLU)
    _, got = oracle_ingest(path, 0, 1)
    assert np.array_equal(got, parts[0][1])


def test_restatement_on_the_reference_fixture(fixtures, tmp_path):
    """tests/golden/reads_fa.npz IS the reference's reads.fa as its own DnaSeq packed it (tests/golden/make_golden.py): written
    back as FASTA text and ingested, the arena must come out byte for byte."""
    dna = fixtures("reads_fa").slice(0, 60)
    seqs = [dna.read_ascii(i) for i in range(dna.size())]
    path = str(tmp_path / "fix.fa")
    F.write_fasta(path, seqs, 80)
    for P in (1, 4):
        bufs = [oracle_ingest(path, r, P)[1] for r in range(P)]
        assert np.array_equal(np.concatenate(bufs), dna.buf)


def test_host_mirror_partition_and_owner(tmp_path):
    rng = np.random.default_rng(5)
    seqs = synth_reads(rng, 57, 50, 400)
    path = str(tmp_path / "p.fa")
    F.write_fasta(path, seqs, 61)
    for P in (1, 2, 5):
        idx = [F.FastaIndex(path, r, P) for r in range(P)]
        assert sum(i.getmyreadcount() for i in idx) == 57 and idx[0].gettotrecords() == 57
        for g in range(57):
            o = idx[0].getreadowner(g)
            assert idx[o].getmyreaddispl() <= g < idx[o].getmyreaddispl() + idx[o].getmyreadcount()
        # consecutive chunks do not overlap, and every record lies inside its rank's chunk
        for i in idx:
            s, e = i.chunk_extent()
            rec = i.getmyrecords().astype(np.int64)
            last = rec[:, 1] + (rec[:, 0] - 1) + (rec[:, 0] - 1) // rec[:, 2]
            assert (rec[:, 1] >= s).all() and (last < e).all()
    names, rec = F.parse_faidx(path + ".fai")
    assert names[:3] == ["1", "2", "3"] and rec.shape == (57, 3) and (rec[:, 2] == 61).all()


def test_a_record_behind_the_chunk_is_refused():
    with pytest.raises(ValueError):
        O.fasta_pack(b"ACGT\nAC", 0, np.array([[10, 0, 4]], dtype=np.uint64))
