"""Host side of the FASTA ingest: the reference's ``FastaIndex`` (include/FastaIndex.hpp, src/FastaIndex.cpp) with
``getmydna`` running on the device (``elba_fe_ingest_fasta``).

Same names and meaning as the reference class: the ``.fai`` next to the FASTA is parsed into records
``(len, pos, bases)`` (``get_faidx_record``, src/FastaIndex.cpp:15-24: the first four whitespace-separated fields of a
line, the name first), the reads are split over the ranks by ``getpartition`` (:47-94), and a rank's chunk of the file is
``[myrecords.front().pos, myrecords.back().pos + len + len / bases)`` clipped to the file size (:221-223).  What differs:
the chunk is handed to the GPU as it is; no per-record line copy and no ``DnaSeq::compress`` on the host.
"""
from __future__ import annotations

import os

import numpy as np

from .distributed import partition_reads
from .dnabuffer import DnaBuffer


def parse_faidx(path: str):
    """(names, records[n, 3] uint64) of a FASTA index file (src/FastaIndex.cpp:15-24,108-114)."""
    names, rec = [], []
    with open(path, "r") as f:
        for line in f:
            t = line.split()
            if len(t) < 4:
                continue            # std::getline + operator>> on a blank line leaves garbage in the reference; skipped here
            names.append(t[0])
            rec.append((int(t[1]), int(t[2]), int(t[3])))
    return names, np.array(rec, dtype=np.uint64).reshape(-1, 3)


class FastaIndex:
    """``FastaIndex(fasta_fname, commgrid)`` with (rank, nranks) in place of the communicator grid."""

    def __init__(self, fasta_fname: str, rank: int = 0, nranks: int = 1):
        self.fasta_fname = fasta_fname
        self.rank, self.nranks = rank, nranks
        self.rootnames, self.rootrecords = parse_faidx(self.get_faidx_fname())
        bounds = partition_reads(self.rootrecords[:, 0].astype(np.int64), nranks)
        self.readcounts = [hi - lo for lo, hi in bounds]
        self.readdispls = [lo for lo, _ in bounds] + [len(self.rootrecords)]
        lo, hi = bounds[rank]
        self.myrecords = np.ascontiguousarray(self.rootrecords[lo:hi])

    def get_fasta_fname(self) -> str:
        return self.fasta_fname

    def get_faidx_fname(self) -> str:
        return self.fasta_fname + ".fai"

    def gettotrecords(self) -> int:
        return self.readdispls[-1]

    def getreadcount(self, i: int) -> int:
        return self.readcounts[i]

    def getreaddispl(self, i: int) -> int:
        return self.readdispls[i]

    def getmyreadcount(self) -> int:
        return self.readcounts[self.rank]

    def getmyreaddispl(self) -> int:
        return self.readdispls[self.rank]

    def getreadowner(self, i: int) -> int:
        """src/FastaIndex.cpp:26-45: the rank whose block holds global read i."""
        return int(np.searchsorted(np.asarray(self.readdispls), i, side="right")) - 1

    def getmyreadlens(self) -> np.ndarray:
        return self.myrecords[:, 0].copy()

    def getmyrecords(self) -> np.ndarray:
        return self.myrecords

    def chunk_extent(self):
        """[startpos, endpos) of this rank's reads in the file (src/FastaIndex.cpp:221-223)."""
        if len(self.myrecords) == 0:
            return 0, 0
        first, last = self.myrecords[0], self.myrecords[-1]
        start = int(first[1])
        end = int(last[1]) + int(last[0]) + (int(last[0]) // int(last[2]) if int(last[2]) else 0)
        return start, min(end, os.path.getsize(self.fasta_fname))

    def read_chunk(self) -> np.ndarray:
        start, end = self.chunk_extent()
        return np.fromfile(self.fasta_fname, dtype=np.uint8, count=end - start, offset=start)

    def getmydna(self, ctx, fetch: bool = True):
        """The rank's reads parsed ON THE DEVICE into ctx (left resident there, ready for ctx.count()); with `fetch` the
        DnaBuffer also comes back to the host, as the reference's getmydna returns it (src/FastaIndex.cpp:191-290)."""
        start, _ = self.chunk_extent()
        ctx.ingest_fasta(self.read_chunk(), start, self.myrecords, self.getmyreaddispl())
        return ctx.reads() if fetch else None


def write_fasta(path: str, seqs, width: int = 80, names=None) -> None:
    """A FASTA file and its .fai (samtools faidx layout: name, length, offset, bases per line, bytes per line) for tests and tools."""
    with open(path, "w", newline="\n") as f, open(path + ".fai", "w", newline="\n") as fi:
        pos = 0
        for i, s in enumerate(seqs):
            name = names[i] if names else str(i + 1)
            head = ">" + name + "\n"
            f.write(head)
            pos += len(head)
            fi.write(f"{name}\t{len(s)}\t{pos}\t{width}\t{width + 1}\n")
            for j in range(0, len(s), width):
                f.write(s[j:j + width] + "\n")
            pos += len(s) + (len(s) + width - 1) // width
