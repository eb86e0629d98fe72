"""Step-by-step diagnosis on a GPU box (not a test): prints where the CUDA path first departs from the oracle."""
import os, sys, time, traceback
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from elba_b200 import frontend
from elba_b200.dnabuffer import DnaBuffer
from elba_b200.synth import make_dnabuffer
from oracle import oracle as O


def first_diff(a, b, name):
    a, b = np.asarray(a), np.asarray(b)
    if a.shape != b.shape:
        print(f"   {name}: shape {a.shape} vs {b.shape}")
        n = min(len(a), len(b))
        a, b = a[:n], b[:n]
    d = np.nonzero((a != b).reshape(len(a), -1).any(axis=1))[0]
    if len(d):
        print(f"   {name}: {len(d)} diffs, first at {d[0]}: got {a[d[0]]} want {b[d[0]]}")
    else:
        print(f"   {name}: equal")
    return len(d) == 0


def case(name, dna, k, lo, up, parts):
    print(f"== {name} k={k} L={lo} U={up} parts={parts} N={dna.size()} M={dna.num_kmers(k)}", flush=True)
    ref = O.run(dna, k, lo, up, threads=8)
    try:
        ctx = frontend.Context(frontend.Params(k=k, lower=lo, upper=up, num_partitions=parts))
        t = time.time(); ctx.upload(dna)
        ks = ctx.kmer_stream() if dna.num_kmers(k) < 5_000_000 else None
        if ks is not None:
            first_diff(ks, O.all_rep_kmers(dna, k), "kmer stream")
        ctx.count(); s = ctx.sizes(); print("   count:", s, "oracle D/R/pre:", ref.D, ref.R, ref.nnzA_pre, flush=True)
        km, cn = ctx.kmers(); first_diff(km, ref.kmers, "kmers"); first_diff(cn, ref.counts, "counts")
        ctx.build_A(); s = ctx.sizes(); print("   A:", s["nnzA"], "F", s["products"], "oracle:", ref.nnzA, ref.F, flush=True)
        rp, col, pos = ctx.A(); first_diff(rp, ref.a_rowptr, "a_rowptr"); first_diff(col, ref.a_col, "a_col"); first_diff(pos, ref.a_pos, "a_pos")
        cp, row, tp = ctx.AT(); first_diff(cp, ref.at_colptr, "at_colptr"); first_diff(row, ref.at_row, "at_row"); first_diff(tp, ref.at_pos, "at_pos")
        ctx.spgemm(); s = ctx.sizes(); print("   B:", s["nnzB_pre"], s["nnzB"], "oracle:", ref.nnzB_pre, ref.nnzB, flush=True)
        brp, bc, bn, bs = ctx.B(); first_diff(brp, ref.b_rowptr, "b_rowptr"); first_diff(bc, ref.b_col, "b_col"); first_diff(bn, ref.b_num, "b_num"); first_diff(bs, ref.b_seeds, "b_seeds")
        print("   timings:", ctx.timings(), "wall", round(time.time() - t, 3), flush=True)
        ctx.close()
    except Exception as e:
        print("   EXCEPTION:", e); traceback.print_exc()


if __name__ == "__main__":
    small = make_dnabuffer(60_000, 200, 6000, 800, 0.12, seed=7)
    case("small", small, 17, 2, 8, 1)
    case("small", small, 17, 2, 8, 4)
    case("small k31", small, 31, 2, 4, 0)
    reads = DnaBuffer.load(os.path.join(ROOT, "tests/golden/reads_fa.npz"))
    case("reads_fa", reads, 17, 2, 8, 0)
    case("reads_fa", reads, 31, 15, 35, 0)
    med = DnaBuffer.load(os.path.join(ROOT, "tests/golden/example_medium.npz"))
    case("example_medium", med, 17, 2, 8, 0)
    case("example_medium", med, 17, 2, 8, 1)
    case("example_medium", med, 31, 15, 35, 0)
