"""Generate the committed parity fixtures.  Run HERE (build container), where /root/reference exists:

    python tests/golden/make_golden.py

It (1) packs the reference's own FASTA fixtures into the DnaBuffer layout (data, not source),
(2) runs the REFERENCE'S OWN code (oracle/_ref: KmerOps.cpp, SharedSeeds.cpp, Kmer, HashFuncs, Bloom,
HyperLogLog compiled unmodified) on them and records sizes + SHA-256 digests of every Tier-1 result,
(3) records the value-level golden vectors of SURVEY.md §8c as produced by the reference code.

The GPU box has no /root/reference; tests there use only what this script wrote.
"""
import hashlib
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
from elba_b200.dnabuffer import DnaBuffer  # noqa: E402
from oracle import oracle as O  # noqa: E402

REF = "/root/reference"


def digest(*arrays):
    h = hashlib.sha256()
    for a in arrays:
        h.update(np.ascontiguousarray(a).tobytes())
    return h.hexdigest()


def tier1_from_ref(r: "O.RefResult"):
    """Reference results re-expressed with canonical column ids (rank of the k-mer value)."""
    order = np.argsort(r.kmers, kind="stable")
    rank = np.empty(r.R, np.int64)
    rank[order] = np.arange(r.R)
    kmers = r.kmers[order]
    counts = r.counts[order].astype(np.uint32)
    rows = np.repeat(np.arange(r.N), np.diff(r.a_rowptr))
    acol = rank[r.a_col]
    ka = np.lexsort((acol, rows))
    return dict(
        kmers=digest(kmers, counts),
        A=digest(r.a_rowptr.astype(np.int64), acol[ka].astype(np.uint32), r.a_val[ka].astype(np.uint32)),
        B=digest(r.b_rowptr.astype(np.int64), r.b_col.astype(np.uint32), r.b_num.astype(np.int32)),
    )


def main():
    out = {"values": {}, "configs": {}}
    # ---- fixtures -------------------------------------------------------------------------------------
    reads = DnaBuffer.from_fasta(os.path.join(REF, "reads.fa"))
    reads.save(os.path.join(HERE, "reads_fa.npz"))
    medium = DnaBuffer.from_fasta(os.path.join(REF, "example_medium", "reads.fa"))
    medium.save(os.path.join(HERE, "example_medium.npz"))
    sub135 = reads.slice(0, 135)
    data = {"reads_fa": reads, "example_medium": medium, "reads_fa_first135": sub135}

    # ---- value-level goldens from the reference's own Kmer / HashFuncs / Bloom / HyperLogLog -------------
    vals = out["values"]
    vals["kmers"] = {}
    for s in ["ACGTACGTACGTACGTA", "TTTTTTTTTTTTTTTTT", "GATTACAGATTACAGAT", "ACGTACGTACGTACGTACGTACGTACGTACG",
              "GATTACAGATTACAGATTACAGATTACAGAT", "ACGTTGCAACGTTGC"]:
        fwd, twin, rep, h = O.ref_kmer_info(s, 2, 8)
        L = O.ref_lib(len(s), 2, 8)
        vals["kmers"][s] = dict(fwd=hex(fwd), twin=hex(twin), rep=hex(rep), hash=hex(h),
                                owner8=int(L.ref_owner(O._u64(rep), 8)), owner4=int(L.ref_owner(O._u64(rep), 4)))
    rng = np.random.default_rng(313)
    some = rng.integers(0, 2**63, 64, dtype=np.uint64) << np.uint64(1)
    some &= ~np.uint64((1 << 30) - 1)   # valid k=17 values
    L17 = O.ref_lib(17, 2, 8)
    vals["hash64"] = {hex(int(x)): hex(int(L17.ref_hash(O._u64(int(x))))) for x in some[:16]}
    bits, hashes, bf = O.ref_bloom_fill(1000, 0.05, some)
    vals["bloom"] = dict(entries=1000, error=0.05, bits=bits, hashes=hashes, keys=[hex(int(x)) for x in some], bf_sha256=digest(bf))
    for name in ("reads_fa",):
        for k in (17, 31):
            est, regs = O.ref_hll(data[name], k, 2, 8 if k == 17 else 4)
            vals[f"hll_{name}_k{k}"] = dict(estimate=est, registers_sha256=digest(regs))
    est, regs = O.ref_hll(medium, 17, 2, 8)
    vals["hll_example_medium_k17"] = dict(estimate=est, registers_sha256=digest(regs))

    # ---- whole-path goldens from the reference's own KmerOps.cpp / SharedSeeds.cpp ----------------------
    runs = [("reads_fa", 17, 2, 8), ("reads_fa", 31, 15, 35), ("reads_fa", 31, 2, 4), ("reads_fa_first135", 17, 2, 8),
            ("example_medium", 17, 2, 8), ("example_medium", 31, 2, 4), ("example_medium", 31, 15, 35)]
    for name, k, lo, up in runs:
        d = data[name]
        r = O.ref_run(d, k, lo, up, nranks=1)
        r4 = O.ref_run(d, k, lo, up, nranks=4, fetch=True)
        t1 = tier1_from_ref(r)
        assert t1 == tier1_from_ref(r4), "reference Tier-1 results differ between np=1 and np=4"
        rows = np.repeat(np.arange(r.N), np.diff(r.b_rowptr))
        key = f"{name}_k{k}_l{lo}_u{up}"
        out["configs"][key] = dict(
            fixture=name, k=k, lower=lo, upper=up, N=d.size(), M=d.num_kmers(k), R=r.R, nnzA=r.nnzA, nnzB_pre=r.nnzB_pre, nnzB=r.nnzB,
            diag=int((rows == r.b_col).sum()), strict_upper=int((rows < r.b_col).sum()), numshared_sum=int(r.b_num.sum()),
            keys_after_pass1=r.keys_after_pass1, digests=t1, ref_secs_np1=r.secs)
        print(key, out["configs"][key], flush=True)
    with open(os.path.join(HERE, "golden.json"), "w") as f:
        json.dump(out, f, indent=1, sort_keys=True)


if __name__ == "__main__":
    main()
