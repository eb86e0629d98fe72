"""Golden digests of the alignment stage, produced by the REFERENCE'S OWN XDropAligner.cpp + Overlap.cpp
(oracle/_ref, compiled unmodified).  Run where /root/reference exists:

    python tests/golden/make_golden_xdrop.py          -> tests/golden/golden_xdrop.json

For every fixture / parameter set: B comes from the pinned oracle of the front end (canonical seed rule), the
aligned pairs are the strict upper triangle with seeds[0] (src/PairwiseAlignment.cpp:52,90), and the 13 result fields per
pair (oracle.XDROP_FIELDS) are hashed."""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from common import digest  # noqa: E402
from elba_b200.dnabuffer import DnaBuffer  # noqa: E402
from oracle import oracle as O  # noqa: E402

CASES = [  # fixture, k, lower, upper, (mat, mis, gap, dropoff)
    ("reads_fa", 17, 2, 8, (1, -1, -1, 15)),          # the defaults of src/main.cpp:53-56
    ("reads_fa", 17, 2, 8, (1, -1, -1, 5)),           # README smoke test: --xa 5
    ("reads_fa", 17, 2, 8, (2, -3, -2, 30)),
    ("reads_fa", 31, 2, 4, (1, -1, -1, 15)),
    ("example_medium", 17, 2, 8, (1, -1, -1, 15)),
]


def main():
    out = {}
    for fx, k, lo, up, sc in CASES:
        dna = DnaBuffer.load(os.path.join(HERE, fx + ".npz"))
        r = O.run(dna, k, lo, up)
        rows, cols, sq, st = O.alignment_pairs(r.b_rowptr, r.b_col, r.b_seeds)
        res = O.ref_xdrop(dna, k, lo, up, rows, cols, sq, st, *sc)
        key = f"{fx}_k{k}_l{lo}_u{up}_m{sc[0]}_x{sc[1]}_g{sc[2]}_d{sc[3]}"
        out[key] = dict(fixture=fx, k=k, lower=lo, upper=up, mat=sc[0], mis=sc[1], gap=sc[2], dropoff=sc[3], pairs=int(len(rows)),
                        passed=int(res[:, 6].sum()), containedQ=int(res[:, 7].sum()), containedT=int(res[:, 8].sum()), rc=int(res[:, 5].sum()),
                        score_sum=int(res[:, 4].astype(np.int64).sum()), digest=digest(res), pairs_digest=digest(rows, cols, sq, st))
        print(key, out[key]["pairs"], out[key]["passed"], out[key]["score_sum"], flush=True)
    with open(os.path.join(HERE, "golden_xdrop.json"), "w") as f:
        json.dump(out, f, indent=1, sort_keys=True)


if __name__ == "__main__":
    main()
