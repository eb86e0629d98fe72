"""Host-side logic of the multi-GPU path on CPU: world_size-2 gloo processes.
Covers the read partition rule (src/FastaIndex.cpp:47-94), the CombBLAS block extents
(src/DistributedFastaData.cpp:21-29) both in Python and through the C ABI (no GPU call), shipping a
communicator id between ranks, and re-assembling blocks of B.  The data path itself (NCCL) is covered
by tests/test_gpu_multi.py on GPUs."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_partition_reads_matches_reference_rule():
    from elba_b200 import distributed as D
    rng = np.random.default_rng(0)
    for n, parts in [(227, 4), (1989, 9), (50, 16), (7, 2)]:
        lens = rng.integers(500, 20000, n)
        b = D.partition_reads(lens, parts)
        assert b[0][0] == 0 and b[-1][1] == n and all(b[i][1] == b[i + 1][0] for i in range(parts - 1))
        avg = lens.sum() / parts
        for lo, hi in b[:-1]:
            assert lens[lo:hi].sum() < avg                                   # never reaches the average ...
            assert hi == n or lens[lo:hi].sum() + lens[hi] >= avg            # ... and the next read would


def test_block_extent_python_and_c_abi_agree():
    import ctypes as C
    from elba_b200 import distributed as D, frontend
    L = frontend.load_library()
    L.elba_fe_block_extent.restype = None
    for n in (0, 1, 7, 227, 1989, 275699):
        for parts in (1, 2, 3, 4, 8):
            tot = 0
            for i in range(parts):
                o, l = C.c_int64(), C.c_int64()
                L.elba_fe_block_extent(C.c_int64(n), C.c_int(parts), C.c_int(i), C.byref(o), C.byref(l))
                assert (o.value, l.value) == D.block_extent(n, parts, i)
                tot += l.value
            assert tot == n


def test_default_grids():
    from elba_b200 import distributed as D
    assert [D.default_grid(w) for w in (1, 2, 4, 8, 16, 6)] == [(1, 1), (1, 2), (2, 2), (2, 4), (4, 4), (2, 3)]


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    sys.path.insert(0, ROOT)
    import torch
    import torch.distributed as dist
    from elba_b200 import distributed as D
    from elba_b200.dnabuffer import DnaBuffer
    dist.init_process_group("gloo", rank=rank, world_size=world)

    class FakeCtx:                      # records what bootstrap_comm hands the library
        got = None

        @staticmethod
        def comm_get_id():
            return bytes(range(128))

        def comm_init(self, comm_id, rank, nranks, grid=None):
            FakeCtx.got = (bytes(comm_id), rank, nranks, grid)
    ctx = FakeCtx()
    D.bootstrap_comm(ctx, dist, grid=(1, 2))
    dna = DnaBuffer.load(os.path.join(ROOT, "tests", "golden", "reads_fa.npz"))
    mine, first = D.local_reads(dna, rank, world)
    # every read in exactly one rank, in order
    t = torch.tensor([first, mine.size(), int(mine.lengths.sum())], dtype=torch.int64)
    allt = [torch.zeros(3, dtype=torch.int64) for _ in range(world)]
    dist.all_gather(allt, t)
    # blocks of a fake B: rank r owns column block r of a 1 x world grid
    n = dna.size()
    c0, nc = D.block_extent(n, world, rank)
    rows = np.arange(n, dtype=np.int64)
    cols = (c0 + rows % max(nc, 1)).astype(np.int64)
    blk = (rows, cols, np.full(n, rank + 2, np.int32), np.tile(np.arange(4, dtype=np.uint32), (n, 1)))
    gathered = [None] * world
    dist.all_gather_object(gathered, blk)
    rp, col, num, seeds = D.merge_B_blocks(gathered, n)
    q.put((rank, FakeCtx.got, [x.tolist() for x in allt], rp.tolist()[-1], bool((np.diff(rp) == world).all()), bool(all((np.diff(col[rp[i]:rp[i + 1]]) > 0).all() for i in range(n)))))
    dist.destroy_process_group()


def test_gloo_world2_bootstrap_partition_and_merge():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    world, port = 2, 29731
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=180) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, got, allt, nnz, rows_ok, cols_sorted in res:
        assert got == (bytes(range(128)), rank, world, (1, 2))              # same id on every rank, own rank
        firsts, counts = [a[0] for a in allt], [a[1] for a in allt]
        assert firsts[0] == 0 and firsts[1] == counts[0] and sum(counts) == 227
        assert nnz == 227 * world and rows_ok and cols_sorted
