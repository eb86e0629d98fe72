"""FASTA ingest on the device (elba_fe_ingest_fasta, fasta_ingest.cuh) and the DCSC hand-off of B (§8f-3:
elba_fe_get_B_dcsc, dcsc.cuh), through the C ABI, bit-exact against the oracle's restatement, against the reference's own
FastaIndex.cpp where oracle/_ref is present, and against the committed fixture of the reference's reads."""
import os

import numpy as np
import pytest

from test_fasta_host import KLU, oracle_ingest, synth_reads

pytestmark = pytest.mark.gpu


def _ctx(k=17, lower=2, upper=8):
    from elba_b200 import frontend
    return frontend.Context(frontend.Params(k=k, lower=lower, upper=upper, device=0))


@pytest.mark.parametrize("width", [1, 2, 3, 4, 5, 7, 60, 80, 100000])
def test_ingest_equals_oracle_and_reference(tmp_path, width):
    from elba_b200 import fasta as F
    from oracle import oracle as O
    rng = np.random.default_rng(100 + width)
    # short reads, reads around the 2048-base work item and its multiples, lower case, N
    seqs = (synth_reads(rng, 60, 1, 700, "ACGTacgtNn") + ["A", "AC", "ACG", "ACGT", "ACGTA"] +
            synth_reads(rng, 6, 2040, 2056) + synth_reads(rng, 4, 4090, 4110) + synth_reads(rng, 3, 20000, 30000))
    order = rng.permutation(len(seqs))
    seqs = [seqs[i] for i in order]
    path = str(tmp_path / "r.fa")
    F.write_fasta(path, seqs, width)
    ctx = _ctx()
    for P in (1, 3):
        ref = O.ref_fasta(path, P, KLU)[1] if O.ref_available(*KLU) else None
        for r in range(P):
            idx, want = oracle_ingest(path, r, P)
            dna = idx.getmydna(ctx)
            assert np.array_equal(dna.lengths, idx.getmyreadlens())
            assert np.array_equal(dna.buf, want), f"arena differs from the oracle (width {width}, rank {r} of {P})"
            if ref is not None:
                assert np.array_equal(dna.buf, ref[r][1]), "arena differs from the reference's FastaIndex::getmydna"
            off = np.concatenate([[0], np.cumsum((dna.lengths + np.uint64(3)) // np.uint64(4))[:-1]]).astype(np.uint64)
            assert np.array_equal(dna.offsets, off)
    ctx.close()


def test_characters_outside_the_table(tmp_path):
    from elba_b200 import fasta as F
    seqs = ["ACGTRYKM", "XACGT", "AXCGT", "ACXGT", "ACGXT", "acgu-*.t", "NNNNRNNNN"]
    path = str(tmp_path / "odd.fa")
    F.write_fasta(path, seqs, 5)
    ctx = _ctx()
    idx, want = oracle_ingest(path, 0, 1)
    assert np.array_equal(idx.getmydna(ctx).buf, want)
    ctx.close()


def test_ingested_fixture_runs_the_path_like_the_uploaded_one(fixtures, tmp_path, golden):
    """The reference's own reads (tests/golden/reads_fa.npz) written back as FASTA: the ingested arena equals the reference's
    DnaBuffer byte for byte, and the path run on it gives the golden digests of the reference run."""
    from elba_b200 import fasta as F, frontend
    from common import digest
    dna = fixtures("reads_fa")
    path = str(tmp_path / "reads.fa")
    F.write_fasta(path, [dna.read_ascii(i) for i in range(dna.size())], 80)
    ctx = _ctx(17, 2, 8)
    got = F.FastaIndex(path).getmydna(ctx)
    assert np.array_equal(got.buf, dna.buf) and np.array_equal(got.lengths, dna.lengths)
    ctx.count(); ctx.build_A(); ctx.spgemm()
    s = ctx.sizes()
    ctx2 = _ctx(17, 2, 8)
    ctx2.upload(dna); ctx2.run()
    assert ctx.digests() == ctx2.digests()
    assert all(s[a] == ctx2.sizes()[a] for a in ("reliable", "nnzA", "products", "nnzB"))
    ctx.close(); ctx2.close()


def test_sliced_ingest_of_a_large_chunk(tmp_path):
    """>= 64 MB of FASTA text: the chunk crosses PCIe in slices and every slice is packed as it arrives."""
    from elba_b200 import fasta as F
    from elba_b200.dnabuffer import DnaBuffer
    rng = np.random.default_rng(9)
    n, L = 5000, 15000
    codes = rng.integers(0, 4, n * L, dtype=np.uint8)
    lens = rng.integers(L - 2000, L, n)
    text = np.frombuffer(b"ACGT", dtype=np.uint8)[codes]
    width = 80
    # write the file with numpy (5000 x 15 kb = 75 MB): header, then the bases in lines of 80
    rec, parts, pos, o = [], [], 0, 0
    for i in range(n):
        head = f">{i + 1}\n".encode()
        l = int(lens[i])
        body = text[o:o + l]; o += L
        full = l // width
        lines = np.full((full, width + 1), 10, np.uint8)
        lines[:, :width] = body[:full * width].reshape(full, width)
        tail = body[full * width:]
        parts += [np.frombuffer(head, np.uint8), lines.ravel(), tail, np.array([10], np.uint8) if len(tail) else np.zeros(0, np.uint8)]
        pos += len(head)
        rec.append((l, pos, width))
        pos += l + (l + width - 1) // width
    raw = np.concatenate(parts)
    assert raw.size >= (64 << 20)
    rec = np.array(rec, dtype=np.uint64)
    want_codes = np.concatenate([codes[i * L:i * L + int(lens[i])] for i in range(n)])
    want = DnaBuffer.from_codes(want_codes, lens.astype(np.int64))
    ctx = _ctx()
    ctx.ingest_fasta(raw, 0, rec, 0)
    got = ctx.reads()
    assert np.array_equal(got.buf, want.buf) and np.array_equal(got.lengths, want.lengths)
    t = ctx.timings()
    print(f"[fasta] {raw.size / 1e6:.0f} MB of FASTA -> {got.buf.size / 1e6:.0f} MB arena in {t['upload_ms']:.2f} ms (H2D + pack)")
    ctx.close()


def test_error_behaviour():
    from elba_b200 import frontend
    ctx = _ctx()
    with pytest.raises(frontend.FrontEndError):
        ctx.ingest_fasta(b"ACGT\nAC", 0, np.array([[10, 0, 4]], dtype=np.uint64))        # ends behind the chunk
    with pytest.raises(frontend.FrontEndError):
        ctx.ingest_fasta(b"ACGT\nAC", 3, np.array([[2, 0, 4]], dtype=np.uint64))         # starts before the chunk
    with pytest.raises(frontend.FrontEndError):
        ctx.ingest_fasta(b"ACGT\nAC", 0, np.array([[2, 0, 0]], dtype=np.uint64))         # zero bases per line
    ctx.ingest_fasta(b"", 0, np.zeros((0, 3), np.uint64))                                # no reads: fine
    assert ctx.reads().size() == 0
    ctx.close()


# ---- B by column, doubly compressed -----------------------------------------------------------------------------------
@pytest.mark.parametrize("fixture,k,lo,up", [("reads_fa", 17, 2, 8), ("reads_fa_first135", 31, 2, 4)])
def test_B_dcsc_is_the_column_major_form_of_B(fixtures, fixture, k, lo, up):
    import scipy.sparse as sp
    from elba_b200 import frontend
    dna = fixtures(fixture)
    ctx = _ctx(k, lo, up)
    ctx.upload(dna); ctx.run()
    rp, col, num, seeds = ctx.B()
    jc, cp, ir, dnum, dseeds = ctx.B_dcsc()
    n, nnz = dna.size(), len(col)
    rows = np.repeat(np.arange(n), np.diff(rp))
    order = np.lexsort((rows, col))                    # by column, rows ascending inside a column
    assert np.array_equal(ir, rows[order]) and np.array_equal(dnum, num[order]) and np.array_equal(dseeds, seeds[order])
    cols_sorted = col[order]
    heads = np.flatnonzero(np.r_[True, cols_sorted[1:] != cols_sorted[:-1]]) if nnz else np.zeros(0, np.int64)
    assert np.array_equal(jc, cols_sorted[heads]) and np.array_equal(cp, np.r_[heads, nnz])
    # the same through scipy: CSC of the pattern with numshared as values
    m = sp.csr_matrix((num, col, rp), shape=(n, n)).tocsc()
    m.sort_indices()
    nz = np.flatnonzero(np.diff(m.indptr))
    assert np.array_equal(jc, nz) and np.array_equal(ir, m.indices) and np.array_equal(dnum, m.data)
    ctx.close()
