"""Golden digests of the string graph S = TransitiveReduction(R) as the reference's own src/TransitiveReduction.cpp computes it
(oracle/_ref; needs /root/reference to build).  Run from the repo root:  python tests/golden/make_golden_tr.py"""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from common import digest
from tr_inputs import overlap_graph
from elba_b200.dnabuffer import DnaBuffer
from oracle import oracle as O

out = {}
for fixture, k, lo, up in (("reads_fa", 17, 2, 8), ("reads_fa", 31, 2, 4)):
    dna = DnaBuffer.load(os.path.join(ROOT, "tests", "golden", fixture + ".npz"))
    n, rows, cols, f = overlap_graph(dna, k, lo, up)
    r, c, of = O.ref_transitive_reduction(n, rows, cols, f, (17, 2, 8))
    out[f"{fixture}_k{k}_l{lo}_u{up}"] = dict(fixture=fixture, k=k, lower=lo, upper=up, nnzR=int(len(rows)), nnzS=int(len(r)),
                                             input_digest=digest(rows, cols, f), digest=digest(r, c, of))
    print(fixture, k, lo, up, "R", len(rows), "S", len(r))
json.dump(out, open(os.path.join(ROOT, "tests", "golden", "golden_tr.json"), "w"), indent=1)
