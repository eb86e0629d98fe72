#!/usr/bin/env python
"""Writes gpurun_out/bench_digests.json (to be committed as tests/golden/bench_digests.json): for every benchmark workload the
digest of its input and the digests of its results.  Where the CPU oracle can run the workload (config 2, config 3 in full,
config 4 at 5 %) the entry is written only if the oracle's numpy digests equal the device's; full config 4 is pinned by the
one-GPU run (and by N = 1 == N = 8 in bench.py, tests/test_gpu_fullscale.py checks the seeds' validity).  Run on a B200."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch  # noqa: E402
import bench  # noqa: E402
from common import oracle_result_digests  # noqa: E402
from elba_b200 import frontend, synth  # noqa: E402
from oracle import oracle as O  # noqa: E402

out = {}
dev = torch.device("cuda", 0)
for name, scale, use_oracle in (("example_medium_k17", 1.0, True), ("ecoli30x_clr", 1.0, True), ("celegans40x_hifi", 0.05, True), ("celegans40x_hifi", 1.0, False)):
    buf, off, lens, k, lo, up, total, r0 = bench.load_workload(name, 0, 1, dev, scale)
    din = synth.input_digest_parts(buf, lens, 0, 0)
    ctx = frontend.Context(frontend.Params(k=k, lower=lo, upper=up, device=0))
    ctx.set_reads_device(buf.data_ptr(), buf.numel(), off.data_ptr(), lens.data_ptr(), lens.numel(), 0)
    ctx.run()
    d, s = ctx.digests(), ctx.sizes()
    ctx.close()
    pinned = "one-GPU run of this library (tools/make_bench_golden.py); the oracle cannot run this size"
    if use_oracle:
        ref = O.run(synth.to_dnabuffer(buf, off, lens), k, lo, up, threads=min(16, os.cpu_count() or 1))
        assert d == oracle_result_digests(ref), (name, scale, d, oracle_result_digests(ref))
        pinned = "CPU oracle (oracle/elba_oracle.cpp, pinned to the reference's own sources) == device"
    out[f"{name}@{scale:g}"] = dict(input_digest="".join(f"{x:016x}" for x in din), **d, sizes={a: s[a] for a in ("nreads", "num_kmers", "distinct", "reliable", "nnzA_pre", "nnzA", "products", "nnzB_pre", "nnzB")}, pinned_by=pinned)
    print(name, scale, out[f"{name}@{scale:g}"], flush=True)
    del buf, off, lens
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
with open(os.path.join(ROOT, "gpurun_out", "bench_digests.json"), "w") as f:
    json.dump(out, f, indent=1)
