"""Throughput of the device FASTA ingest (elba_fe_ingest_fasta) on a synthetic FASTA of the C. elegans HiFi read shape.  Tuning aid.

    python tools/ingest_probe.py [--mb 1000] [--passes 4]

Prints the library's own CUDA-event time of upload + pack (H2D of the raw text in slices on the second stream, k_fasta_pack per
slice) from pinned host memory, the bytes moved, and a check of the arena against numpy."""
import argparse, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
from elba_b200 import frontend
from elba_b200.dnabuffer import DnaBuffer

ap = argparse.ArgumentParser()
ap.add_argument("--mb", type=int, default=400)
ap.add_argument("--passes", type=int, default=4)
ap.add_argument("--width", type=int, default=80)
a = ap.parse_args()
rng = np.random.default_rng(313)
W = a.width
n = max(1, a.mb * 1_000_000 // 14_700)
lens = np.clip(rng.normal(14550, 1000, n).astype(np.int64), 1000, None)
t0 = time.time()
codes = rng.integers(0, 4, int(lens.sum()), dtype=np.uint8)
text = np.frombuffer(b"ACGT", dtype=np.uint8)[codes]
parts, rec, pos, o = [], [], 0, 0
for i in range(n):
    head = np.frombuffer(f">{i + 1}\n".encode(), np.uint8)
    l = int(lens[i]); full = l // W
    body = text[o:o + l]; o += l
    lines = np.full((full, W + 1), 10, np.uint8)
    lines[:, :W] = body[:full * W].reshape(full, W)
    tail = body[full * W:]
    parts += [head, lines.ravel(), tail] + ([np.array([10], np.uint8)] if len(tail) else [])
    pos += len(head)
    rec.append((l, pos, W))
    pos += l + (l + W - 1) // W
raw = torch.from_numpy(np.concatenate(parts)).pin_memory()
rec = np.array(rec, dtype=np.uint64)
want = DnaBuffer.from_codes(codes, lens)
print(f"[ingest] {n} reads, {raw.numel() / 1e6:.0f} MB of FASTA text (line width {W}), built in {time.time() - t0:.1f}s", flush=True)
ctx = frontend.Context(frontend.Params(k=31, lower=2, upper=4))
best = None
for p in range(a.passes):
    t1 = time.perf_counter()
    ctx.ingest_fasta(raw.numpy(), 0, rec, 0)
    wall = (time.perf_counter() - t1) * 1e3
    ms = ctx.timings()["upload_ms"]
    best = ms if best is None else min(best, ms)
    print(f"[ingest] pass {p}: device {ms:.2f} ms (H2D + pack), call wall {wall:.2f} ms", flush=True)
got = ctx.reads()
ok = np.array_equal(got.buf, want.buf) and np.array_equal(got.lengths, want.lengths.astype(np.uint64))
print(f"[ingest] best {best:.2f} ms = {raw.numel() / best / 1e6:.1f} GB/s of FASTA text, {int(lens.sum()) / best / 1e6:.1f} Gbases/s; arena {'== numpy' if ok else 'DIFFERS'}", flush=True)
# the same reads as a packed arena through elba_fe_upload_reads, for scale
hb = torch.from_numpy(want.buf).pin_memory(); ho = torch.from_numpy(want.offsets.astype(np.int64)).pin_memory(); hl = torch.from_numpy(want.lengths.astype(np.int64)).pin_memory()
for p in range(2):
    ctx.upload_raw(hb.data_ptr(), hb.numel(), ho.data_ptr(), hl.data_ptr(), n, 0)
    ctx.synchronize()
    print(f"[ingest] upload of the packed arena ({hb.numel() / 1e6:.0f} MB): device {ctx.timings()['upload_ms']:.2f} ms", flush=True)
ctx.close()
sys.exit(0 if ok else 1)
