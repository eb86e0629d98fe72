"""One rank of the multi-GPU parity check (launched by torchrun from test_gpu_multi.py, or by hand):
every rank runs the front end on its block of reads; rank 0 merges the blocks of B and compares everything
with the CPU oracle run on ALL reads."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from elba_b200 import frontend, distributed as D  # noqa: E402
from elba_b200.dnabuffer import DnaBuffer  # noqa: E402


def main():
    fixture, k, lo, up = sys.argv[1], int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4])
    grid = tuple(int(x) for x in sys.argv[5].split("x")) if len(sys.argv) > 5 and sys.argv[5] != "-" else None
    parts = int(sys.argv[6]) if len(sys.argv) > 6 else 0
    base = int(sys.argv[7]) if len(sys.argv) > 7 else 0      # global id of the first read of rank 0 (the ids need not start at 0)
    do_align = len(sys.argv) > 8 and sys.argv[8] == "align"  # also run the X-drop alignment of every rank's block of B
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    if fixture.startswith("synth:"):
        from elba_b200.synth import make_dnabuffer
        g, n, m, e = fixture[6:].split(",")
        dna = make_dnabuffer(genome_len=int(g), n_reads=int(n), mean_len=int(m), sd_len=int(m) // 8, err=float(e), seed=5, repeat_frac=0.05)
    else:
        dna = DnaBuffer.load(os.path.join(ROOT, "tests", "golden", fixture + ".npz"))
    mine, first = D.local_reads(dna, rank, world)
    ctx = frontend.Context(frontend.Params(k=k, lower=lo, upper=up, device=local, num_partitions=parts))
    D.bootstrap_comm(ctx, dist, device=torch.device("cuda", local), grid=grid)
    ctx.upload(mine, first + base)
    ctx.run()
    info = ctx.comm_info()
    sizes, gsizes = ctx.sizes(), ctx.sizes_global()
    kmers, counts = ctx.kmers()
    arp, acol, apos = ctx.A()
    trip = ctx.B_triples()
    digests = ctx.digests()
    trip = (trip[0] - base, trip[1] - base, trip[2], trip[3])
    info = dict(info, row0=info["row0"] - base, col0=info["col0"] - base)
    aligned = None
    if do_align:
        ar, ac, af = ctx.align(1, -1, -1, 15)
        aligned = (ar - base, ac - base, af)
    payload = dict(rank=rank, info=info, sizes=sizes, digests=digests, kmers=kmers, counts=counts, A=(arp, acol, apos), first=first, n=mine.size(), B=trip,
                   aligned=aligned, timings=ctx.timings())
    gathered = [None] * world if rank == 0 else None
    dist.gather_object(payload, gathered, dst=0)
    ok = True
    if rank == 0:
        from oracle import oracle as O
        ref = O.run(dna, k, lo, up, threads=8)
        try:
            assert gsizes["nreads"] == dna.size() and gsizes["num_kmers"] == ref.M and gsizes["distinct"] == ref.D
            assert gsizes["reliable"] == ref.R and gsizes["nnzA_pre"] == ref.nnzA_pre and gsizes["nnzA"] == ref.nnzA
            assert gsizes["products"] == ref.F, (gsizes["products"], ref.F)
            assert gsizes["nnzB_pre"] == ref.nnzB_pre and gsizes["nnzB"] == ref.nnzB
            for g in gathered:                      # the reliable list is replicated
                assert np.array_equal(g["kmers"], ref.kmers) and np.array_equal(g["counts"], ref.counts)
            # A: row blocks in rank order
            rp = np.concatenate([[0]] + [g["A"][0][1:] + sum(len(h["A"][1]) for h in gathered[:i]) for i, g in enumerate(gathered)])
            assert np.array_equal(rp, ref.a_rowptr)
            assert np.array_equal(np.concatenate([g["A"][1] for g in gathered]), ref.a_col)
            assert np.array_equal(np.concatenate([g["A"][2] for g in gathered]), ref.a_pos)
            # B: blocks of the grid
            pr, pc = gathered[0]["info"]["grid_rows"], gathered[0]["info"]["grid_cols"]
            for g in gathered:
                i, j = g["rank"] // pc, g["rank"] % pc
                assert (g["info"]["row0"], g["info"]["nrows"]) == D.block_extent(dna.size(), pr, i)
                assert (g["info"]["col0"], g["info"]["ncols"]) == D.block_extent(dna.size(), pc, j)
                r, c = g["B"][0], g["B"][1]
                if len(r):
                    assert r.min() >= g["info"]["row0"] and r.max() < g["info"]["row0"] + g["info"]["nrows"]
                    assert c.min() >= g["info"]["col0"] and c.max() < g["info"]["col0"] + g["info"]["ncols"]
            brp, bcol, bnum, bseeds = D.merge_B_blocks([g["B"] for g in gathered], dna.size())
            assert np.array_equal(brp, ref.b_rowptr) and np.array_equal(bcol, ref.b_col), "pattern of B"
            assert np.array_equal(bnum, ref.b_num), "numshared"
            assert np.array_equal(bseeds, ref.b_seeds), "seeds"
            if do_align:
                # every unordered pair of distinct reads of B's pattern is aligned exactly once over the ranks; the 13 fields of
                # each aligned (row, column) equal the oracle's for that orientation and the seed B holds at (row, column)
                rows = np.repeat(np.arange(dna.size()), np.diff(ref.b_rowptr))
                seed_of = {(int(r), int(c)): (int(s[0]), int(s[1])) for r, c, s in zip(rows, ref.b_col, ref.b_seeds)}
                ar = np.concatenate([g["aligned"][0] for g in gathered]); ac = np.concatenate([g["aligned"][1] for g in gathered])
                af = np.concatenate([g["aligned"][2] for g in gathered])
                pairs = sorted((min(int(r), int(c)), max(int(r), int(c))) for r, c in zip(ar, ac))
                want_pairs = sorted((int(r), int(c)) for r, c in zip(rows, ref.b_col) if r < c)
                assert pairs == want_pairs, ("aligned pairs", len(pairs), len(want_pairs))
                if pr != pc:
                    assert (ar < ac).all(), "non-square grid: the global upper triangle"
                sq = np.array([seed_of[(int(r), int(c))][0] for r, c in zip(ar, ac)], dtype=np.uint32)
                st = np.array([seed_of[(int(r), int(c))][1] for r, c in zip(ar, ac)], dtype=np.uint32)
                want = O.xdrop(dna, k, ar, ac, sq, st, 1, -1, -1, 15)
                bad = np.where((af != want).any(axis=1))[0]
                assert len(bad) == 0, ("alignment fields", len(bad), af[bad[:3]].tolist(), want[bad[:3]].tolist())
                print(f"MULTI-GPU ALIGNMENT OK pairs={len(pairs)} per rank={[len(g['aligned'][0]) for g in gathered]}", flush=True)
            if base == 0:
                from common import oracle_result_digests
                want = oracle_result_digests(ref)
                for g in gathered:
                    assert g["digests"] == want, ("device digests", g["rank"], g["digests"], want)
            print(f"MULTI-GPU PARITY OK world={world} grid={pr}x{pc} N={dna.size()} R={ref.R} nnzB={ref.nnzB} "
                  f"exchange_ms={[round(g['timings']['exchange_ms'], 3) for g in gathered]}", flush=True)
        except AssertionError as e:
            import traceback
            traceback.print_exc()
            print("MULTI-GPU PARITY FAILED", e, flush=True)
            ok = False
    flag = torch.tensor([1 if ok else 0], device="cuda")
    dist.broadcast(flag, src=0)
    ctx.close()
    dist.destroy_process_group()
    sys.exit(0 if int(flag.item()) else 1)


if __name__ == "__main__":
    main()
