"""Inputs for the transitive-reduction tests: random overlap graphs, and the overlap graph of a fixture as the reference's driver
would hand it to TransitiveReduction (src/main.cpp:300-313: alignments of B's upper triangle, !passed pruned, contained reads
removed), computed with the CPU oracle."""
import numpy as np

PASSED, CONT_Q, CONT_T, DIR, DIR_T, SUF, SUF_T = 6, 7, 8, 9, 10, 11, 12


def random_graph(rng, n, density=0.3, upper_only=True, with_invalid=True):
    pairs = [(i, j) for i in range(n) for j in (range(i + 1, n) if upper_only else range(n))]
    m = int(round(density * len(pairs)))
    sel = rng.choice(len(pairs), size=m, replace=False) if m else np.zeros(0, np.int64)
    rows = np.array([pairs[s][0] for s in sel], np.int64)
    cols = np.array([pairs[s][1] for s in sel], np.int64)
    lo = -1 if with_invalid else 0
    f = np.stack([rng.integers(lo, 4, m), rng.integers(lo, 4, m), rng.integers(-50, 3000, m), rng.integers(-50, 3000, m)], 1).astype(np.int32)
    return rows, cols, f.reshape(-1, 4)


def overlap_graph(dna, k, lower, upper, scoring=(1, -1, -1, 15)):
    """(nreads, rows, cols, fields[nnz, 4]) of the pruned overlap matrix R of `dna`."""
    from oracle import oracle as O
    ref = O.run(dna, k, lower, upper)
    er, ec, sq, st = O.alignment_pairs(ref.b_rowptr, ref.b_col, ref.b_seeds)
    out = O.xdrop(dna, k, er, ec, sq, st, *scoring)
    passed = out[:, PASSED] != 0
    contained = set(er[passed & (out[:, CONT_Q] != 0)].tolist()) | set(ec[passed & (out[:, CONT_T] != 0)].tolist())
    keep = passed & np.array([r not in contained and c not in contained for r, c in zip(er.tolist(), ec.tolist())], dtype=bool)
    return dna.size(), er[keep], ec[keep], out[keep][:, [DIR, DIR_T, SUF, SUF_T]].astype(np.int32)
