"""Host-side helpers for running the front end on several GPUs, one process per GPU.

The data path between GPUs lives in the CUDA library (NCCL over NVLink, ``elba_fe_comm_*``); this module only
does what the reference's driver does around it: hand every rank a contiguous block of reads
(``FastaIndex::getpartition``, src/FastaIndex.cpp:47-94), ship the communicator id, and put the blocks of B back
together for checks.  ``torch.distributed`` (any backend: nccl on GPUs, gloo in the CPU tests) is the messenger.
"""
from __future__ import annotations

import numpy as np

from .dnabuffer import DnaBuffer


def partition_reads(lengths: np.ndarray, nparts: int):
    """Contiguous read ranges balanced by bases, the reference's rule (FastaIndex::getpartition,
    src/FastaIndex.cpp:47-94): every part but the last takes reads while the next one keeps it BELOW
    total / nparts bases; the last part takes the rest.  Returns [(lo, hi)] * nparts."""
    n = int(len(lengths))
    lens = np.asarray(lengths, dtype=np.int64)
    avg = float(lens.sum()) / nparts if nparts else 0.0
    bounds, readid = [], 0
    for _ in range(nparts - 1):
        sofar, start = 0, readid
        while readid < n and sofar + int(lens[readid]) < avg:
            sofar += int(lens[readid])
            readid += 1
        bounds.append((start, readid))
    bounds.append((readid, n))
    return bounds


def block_extent(n: int, parts: int, idx: int):
    """CombBLAS block distribution (src/DistributedFastaData.cpp:21-29): n // parts each, remainder to the last."""
    per = n // parts
    off = per * idx
    return off, (n - off if idx == parts - 1 else per)


def default_grid(nranks: int):
    rows = 1
    r = 1
    while r * r <= nranks:
        if nranks % r == 0:
            rows = r
        r += 1
    return rows, nranks // rows


def bootstrap_comm(ctx, dist, device=None, grid=None):
    """Collective: rank 0 creates the communicator id, everybody joins.  `dist` is an initialised torch.distributed."""
    import torch
    rank, world = dist.get_rank(), dist.get_world_size()
    if world == 1:
        ctx.comm_init(bytes(128), 0, 1)
        return
    t = torch.zeros(128, dtype=torch.uint8, device=device if device is not None else "cpu")
    if rank == 0:
        raw = type(ctx).comm_get_id()
        t.copy_(torch.frombuffer(bytearray(raw), dtype=torch.uint8))
    dist.broadcast(t, src=0)
    ctx.comm_init(bytes(t.cpu().numpy().tobytes()), rank, world, grid=grid)


def local_reads(dna: DnaBuffer, rank: int, world: int):
    """This rank's block of reads and the global id of its first read."""
    lo, hi = partition_reads(dna.lengths, world)[rank]
    return dna.slice(lo, hi), lo


def merge_B_blocks(blocks, nreads: int):
    """blocks: list of (row, col, numshared, seeds) global triples, one per rank -> CSR over all reads, columns ascending."""
    row = np.concatenate([b[0] for b in blocks]) if blocks else np.zeros(0, np.int64)
    col = np.concatenate([b[1] for b in blocks]) if blocks else np.zeros(0, np.int64)
    num = np.concatenate([b[2] for b in blocks]) if blocks else np.zeros(0, np.int32)
    seeds = np.concatenate([b[3] for b in blocks]) if blocks else np.zeros((0, 4), np.uint32)
    order = np.lexsort((col, row))
    row, col, num, seeds = row[order], col[order], num[order], seeds[order]
    rowptr = np.zeros(nreads + 1, np.int64)
    np.add.at(rowptr, row + 1, 1)
    return np.cumsum(rowptr), col.astype(np.uint32), num, seeds
