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


def _fasta_worker(rank, world, port, path, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    sys.path.insert(0, ROOT)
    import torch
    import torch.distributed as dist
    from elba_b200 import fasta as F
    from oracle import oracle as O
    dist.init_process_group("gloo", rank=rank, world_size=world)
    # what every rank of the multi-GPU ingest does on the host: its own index block, its own chunk of the file
    idx = F.FastaIndex(path, rank, world)
    start, end = idx.chunk_extent()
    chunk = idx.read_chunk()
    arena = O.fasta_pack(chunk.tobytes(), start, idx.getmyrecords())        # the checker stands in for the device call here
    t = torch.tensor([idx.getmyreaddispl(), idx.getmyreadcount(), start, end, arena.size], dtype=torch.int64)
    allt = [torch.zeros(5, dtype=torch.int64) for _ in range(world)]
    dist.all_gather(allt, t)
    parts = [None] * world
    dist.all_gather_object(parts, arena)
    q.put((rank, [x.tolist() for x in allt], np.concatenate(parts)))
    dist.destroy_process_group()


def test_gloo_world2_fasta_index_blocks_and_chunks(fixtures, tmp_path):
    """Two ranks ingest their blocks of one FASTA file (src/FastaIndex.cpp:98-176,191-290): consecutive read blocks, chunks that
    do not overlap, and the arenas of the ranks, put end to end, are the arena of the whole file."""
    import torch.multiprocessing as mp
    from elba_b200 import fasta as F
    dna = fixtures("reads_fa").slice(0, 90)
    path = str(tmp_path / "reads.fa")
    F.write_fasta(path, [dna.read_ascii(i) for i in range(dna.size())], 70)
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    world, port = 2, 29741
    procs = [ctx.Process(target=_fasta_worker, args=(r, world, port, path, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted((r for r in (q.get(timeout=180) for _ in range(world))), key=lambda x: x[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, allt, whole in res:
        displ, count, start, end, nbytes = zip(*allt)
        assert displ[0] == 0 and displ[1] == count[0] and sum(count) == 90
        assert start[0] < end[0] <= start[1] < end[1] <= os.path.getsize(path)
        assert sum(nbytes) == dna.buf.size and np.array_equal(whole, dna.buf)
