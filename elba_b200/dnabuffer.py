"""Host-side mirror of the reference's read container.

``DnaBuffer`` here is laid out byte-for-byte like the reference's
(`/root/reference/src/DnaBuffer.cpp:5-29`, `src/DnaSeq.cpp:7-29`): one contiguous
arena, every read starts on a byte boundary, 4 bases per byte, base *i* of a read
at bits ``6-2*(i%4)`` of byte ``i//4``; A=0 C=1 G=2 T=3, N/n -> 0 (A)
(`include/DnaSeq.hpp:136-154`).  It is the INPUT of the hot path — the C ABI
(`include/elba_fe.h`) takes exactly ``(buf, byte_offsets, lengths)``.

Nothing here computes k-mers; this is input marshalling only.
"""
from __future__ import annotations

import numpy as np

# include/DnaSeq.hpp:136-154: everything that is not ACGTacgtNn is code 4 (undefined)
_CODETAB = np.full(256, 4, dtype=np.uint8)
for _ch, _code in (("A", 0), ("C", 1), ("G", 2), ("T", 3), ("N", 0)):
    _CODETAB[ord(_ch)] = _code
    _CODETAB[ord(_ch.lower())] = _code


class DnaBuffer:
    """Packed reads: ``buf`` (uint8 arena), ``offsets`` (uint64 byte offset per read),
    ``lengths`` (uint64 bases per read)."""

    def __init__(self, buf: np.ndarray, offsets: np.ndarray, lengths: np.ndarray):
        self.buf = np.ascontiguousarray(buf, dtype=np.uint8)
        self.offsets = np.ascontiguousarray(offsets, dtype=np.uint64)
        self.lengths = np.ascontiguousarray(lengths, dtype=np.uint64)
        assert self.offsets.shape == self.lengths.shape

    # -- reference accessors (include/DnaBuffer.hpp:18-23) -----------------
    def size(self) -> int:
        return int(self.lengths.shape[0])

    def getbufsize(self) -> int:
        return int(self.buf.shape[0])

    def __len__(self) -> int:
        return self.size()

    def num_kmers(self, k: int) -> int:
        """M = sum over reads of max(0, len-k+1) (include/KmerOps.hpp:118-119)."""
        l = self.lengths.astype(np.int64)
        return int(np.maximum(l - k + 1, 0).sum())

    def total_bases(self) -> int:
        return int(self.lengths.sum())

    def read_codes(self, i: int) -> np.ndarray:
        """2-bit codes of read i (DnaSeq::operator[], src/DnaSeq.cpp:48-54)."""
        n = int(self.lengths[i])
        o = int(self.offsets[i])
        b = self.buf[o:o + (n + 3) // 4]
        codes = np.stack([(b >> 6) & 3, (b >> 4) & 3, (b >> 2) & 3, b & 3], axis=1).reshape(-1)
        return codes[:n]

    def read_ascii(self, i: int) -> str:
        return "".join("ACGT"[c] for c in self.read_codes(i))

    def slice(self, lo: int, hi: int) -> "DnaBuffer":
        """Reads [lo, hi) as their own buffer (contiguous block, as FastaIndex hands ranks)."""
        if hi <= lo:
            return DnaBuffer(np.zeros(0, np.uint8), np.zeros(0, np.uint64), np.zeros(0, np.uint64))
        b0 = int(self.offsets[lo])
        b1 = int(self.offsets[hi - 1]) + (int(self.lengths[hi - 1]) + 3) // 4
        return DnaBuffer(self.buf[b0:b1].copy(), self.offsets[lo:hi] - np.uint64(b0), self.lengths[lo:hi].copy())

    # -- construction ------------------------------------------------------
    @staticmethod
    def from_codes(codes: np.ndarray, lengths: np.ndarray) -> "DnaBuffer":
        """Pack a flat array of 2-bit codes (uint8 in 0..3), reads concatenated, into the arena."""
        lengths = np.asarray(lengths, dtype=np.int64)
        nbytes = (lengths + 3) // 4
        offsets = np.zeros(len(lengths), dtype=np.int64)
        if len(lengths):
            offsets[1:] = np.cumsum(nbytes)[:-1]
        total = int(nbytes.sum())
        # position of every base inside the padded (4 bases / byte) stream
        starts = np.zeros(len(lengths), dtype=np.int64)
        if len(lengths):
            starts[1:] = np.cumsum(lengths)[:-1]
        padded = np.zeros(total * 4, dtype=np.uint8)
        if codes.size:
            read_of = np.repeat(np.arange(len(lengths), dtype=np.int64), lengths)
            dst = offsets[read_of] * 4 + (np.arange(codes.size, dtype=np.int64) - starts[read_of])
            padded[dst] = codes
        q = padded.reshape(-1, 4)
        buf = (q[:, 0] << 6) | (q[:, 1] << 4) | (q[:, 2] << 2) | q[:, 3]
        return DnaBuffer(buf.astype(np.uint8), offsets.astype(np.uint64), lengths.astype(np.uint64))

    @staticmethod
    def from_strings(seqs) -> "DnaBuffer":
        lengths = np.array([len(s) for s in seqs], dtype=np.int64)
        raw = np.frombuffer("".join(seqs).encode("ascii"), dtype=np.uint8)
        codes = _CODETAB[raw]
        if (codes > 3).any():
            raise ValueError("non-nucleotide character in input (include/DnaSeq.hpp:42-43: undefined behaviour in the reference)")
        return DnaBuffer.from_codes(codes, lengths)

    @staticmethod
    def from_fasta(path: str) -> "DnaBuffer":
        """Read a (possibly line-wrapped) FASTA file (what FastaIndex::getmydna produces at np=1,
        src/FastaIndex.cpp:254-278)."""
        seqs, cur = [], []
        with open(path, "r") as f:
            for line in f:
                if line.startswith(">"):
                    if cur:
                        seqs.append("".join(cur))
                        cur = []
                else:
                    cur.append(line.strip())
        if cur:
            seqs.append("".join(cur))
        return DnaBuffer.from_strings(seqs)

    # -- tiny on-disk format for committed fixtures --------------------------
    def save(self, path: str) -> None:
        np.savez(path, buf=self.buf, lengths=self.lengths)

    @staticmethod
    def load(path: str) -> "DnaBuffer":
        z = np.load(path)
        lengths = z["lengths"].astype(np.int64)
        nbytes = (lengths + 3) // 4
        offsets = np.zeros(len(lengths), dtype=np.int64)
        if len(lengths):
            offsets[1:] = np.cumsum(nbytes)[:-1]
        return DnaBuffer(z["buf"], offsets.astype(np.uint64), lengths.astype(np.uint64))
