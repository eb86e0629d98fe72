"""Synthetic long-read sets in the reference's DnaBuffer layout (input generation only, no k-mer work).

Follows the reference's simulator `runs/simfor.py:1-135` (random genome, reads of normally distributed
length at uniform positions, random strand) and adds the i.i.d. error model SURVEY.md §8(d) asks for
(the reference simulator is error-free, which would put every k-mer above UPPER): per emitted base
an insertion with probability e/3, otherwise the next genome base (after skipping one base with
probability e/3 = deletion) substituted with probability e/3.

Written with torch tensor ops so the same code generates the small CPU test sets and, on `cuda`, the
multi-gigabase benchmark shapes in seconds.  Deterministic for a given (seed, device type).
"""
from __future__ import annotations

import numpy as np
import torch

from .dnabuffer import DnaBuffer

# BASELINE.json configs 3-5: genome length, reads, error rate, k, lower, upper
SHAPES = {
    "ecoli30x_clr": dict(genome=4_641_652, reads=16_890, mean=8244, sd=1000, err=0.15, k=17, lower=2, upper=8),
    "celegans40x_hifi": dict(genome=100_286_401, reads=275_699, mean=14_550, sd=1000, err=0.01, k=31, lower=2, upper=4),
    "human10x_clr": dict(genome=3_100_000_000, reads=4_421_593, mean=7011, sd=1000, err=0.15, k=17, lower=2, upper=4),
}


def random_genome(length: int, seed: int, device="cpu") -> torch.Tensor:
    g = torch.Generator(device=device)
    g.manual_seed(seed)
    return torch.randint(0, 4, (length,), dtype=torch.uint8, device=device, generator=g)


def sample_reads(genome: torch.Tensor, n_reads: int, mean_len: float, sd_len: float, err: float, seed: int,
                 min_len: int = 1000, n_frac: float = 0.0):
    """Returns (codes uint8 [total bases], lengths int64 [n_reads]) on genome.device."""
    dev = genome.device
    g = torch.Generator(device=dev)
    g.manual_seed(seed)
    Lg = genome.numel()
    lens = torch.normal(float(mean_len), float(sd_len), (n_reads,), generator=g, device=dev).round().to(torch.int64)
    lens = lens.clamp_(min=min(min_len, max(1, int(mean_len) // 2)), max=max(1, Lg // 2))
    span = (lens.to(torch.float64) * 1.2 + 16).to(torch.int64)               # room for deletions
    starts = (torch.rand(n_reads, generator=g, device=dev, dtype=torch.float64) * (Lg - span).clamp(min=1).to(torch.float64)).to(torch.int64)
    strand = torch.randint(0, 2, (n_reads,), generator=g, device=dev, dtype=torch.int64)
    T = int(lens.sum().item())
    read_start = torch.zeros(n_reads + 1, dtype=torch.int64, device=dev)
    read_start[1:] = torch.cumsum(lens, 0)
    read_of = torch.repeat_interleave(torch.arange(n_reads, device=dev), lens, output_size=T)
    q = err / 3.0
    u = torch.rand(T, generator=g, device=dev)
    is_ins = u < q
    is_del = (u >= q) & (u < 2 * q)
    is_sub = (u >= 2 * q) & (u < 3 * q)
    adv = (~is_ins).to(torch.int64) + is_del.to(torch.int64)
    csum = torch.cumsum(adv, 0)
    base0 = (csum - adv)[read_start[:-1]]                       # exclusive prefix at each read start
    gpos = starts[read_of] + (csum - base0[read_of]) - 1
    gpos.clamp_(min=0, max=Lg - 1)
    codes = genome[gpos].to(torch.int64)
    rnd = torch.randint(0, 4, (T,), generator=g, device=dev, dtype=torch.int64)
    codes = torch.where(is_ins, rnd, codes)
    codes = torch.where(is_sub, (codes + 1 + rnd % 3) % 4, codes)
    # reverse strand: reverse complement of the read
    local = torch.arange(T, device=dev) - read_start[read_of]
    src = torch.where(strand[read_of] == 1, read_start[read_of] + lens[read_of] - 1 - local, torch.arange(T, device=dev))
    codes = torch.where(strand[read_of] == 1, 3 - codes[src], codes)
    if n_frac > 0:   # an N in the FASTA becomes A (include/DnaSeq.hpp:136-154)
        codes = torch.where(torch.rand(T, generator=g, device=dev) < n_frac, torch.zeros_like(codes), codes)
    return codes.to(torch.uint8), lens


def pack_reads(codes: torch.Tensor, lens: torch.Tensor):
    """2-bit pack into the DnaBuffer arena (src/DnaSeq.cpp:7-29): returns (buf uint8, byte offsets int64, lens int64)."""
    dev = codes.device
    n = lens.numel()
    nbytes = (lens + 3) // 4
    off = torch.zeros(n + 1, dtype=torch.int64, device=dev)
    off[1:] = torch.cumsum(nbytes, 0)
    total = int(off[-1].item())
    T = codes.numel()
    read_start = torch.zeros(n + 1, dtype=torch.int64, device=dev)
    read_start[1:] = torch.cumsum(lens, 0)
    read_of = torch.repeat_interleave(torch.arange(n, device=dev), lens, output_size=T)
    dst = off[read_of] * 4 + (torch.arange(T, device=dev) - read_start[read_of])
    padded = torch.zeros(total * 4, dtype=torch.uint8, device=dev)
    padded[dst] = codes
    qd = padded.view(-1, 4)
    buf = (qd[:, 0] << 6) | (qd[:, 1] << 4) | (qd[:, 2] << 2) | qd[:, 3]
    return buf.contiguous(), off[:-1].contiguous(), lens.contiguous()


def make_reads(genome_len: int, n_reads: int, mean_len: float, sd_len: float, err: float, seed: int = 313, device="cpu",
               batch_reads: int = 0, repeat_frac: float = 0.0, n_frac: float = 0.0):
    """Packed synthetic reads as torch tensors on `device`: (buf uint8, offsets uint64-as-int64, lens int64)."""
    genome = random_genome(genome_len, seed, device)
    if repeat_frac > 0:     # duplicate a slice of the genome to exercise the > UPPER path
        seg = int(genome_len * repeat_frac)
        genome[genome_len - seg:] = genome[:seg]
    if batch_reads <= 0:
        batch_reads = n_reads
    bufs, offs, lens_all, base = [], [], [], 0
    for b0 in range(0, n_reads, batch_reads):
        nb = min(batch_reads, n_reads - b0)
        codes, lens = sample_reads(genome, nb, mean_len, sd_len, err, seed * 1000003 + b0 + 1, n_frac=n_frac)
        buf, off, lens = pack_reads(codes, lens)
        del codes
        bufs.append(buf); offs.append(off + base); lens_all.append(lens)
        base += buf.numel()
    return torch.cat(bufs), torch.cat(offs), torch.cat(lens_all)


GLOBAL_BATCH = 16384       # reads per generation batch: batch b holds the global reads [b * GLOBAL_BATCH, (b + 1) * GLOBAL_BATCH)


def make_reads_block(genome_len: int, n_reads: int, mean_len: float, sd_len: float, err: float, seed: int, device, r0: int, r1: int):
    """The global reads [r0, r1) of the set make_reads(genome_len, n_reads, ..., batch_reads=GLOBAL_BATCH) would produce:
    batches are seeded by their GLOBAL first read and always generated whole, then sliced, so every read is bit-identical
    no matter how the set is split over GPUs.  Returns (buf, byte offsets relative to the block, lens)."""
    genome = random_genome(genome_len, seed, device)
    bufs, offs, lens_all, base = [], [], [], 0
    for b0 in range((r0 // GLOBAL_BATCH) * GLOBAL_BATCH, r1, GLOBAL_BATCH):
        nb = min(GLOBAL_BATCH, n_reads - b0)
        codes, lens = sample_reads(genome, nb, mean_len, sd_len, err, seed * 1000003 + b0 + 1)
        buf, off, lens = pack_reads(codes, lens)
        del codes
        lo, hi = max(r0, b0) - b0, min(r1, b0 + nb) - b0
        if lo > 0 or hi < nb:
            byte_lo = int(off[lo].item())
            byte_hi = int(off[hi].item()) if hi < nb else buf.numel()
            buf, off, lens = buf[byte_lo:byte_hi].clone(), off[lo:hi] - byte_lo, lens[lo:hi]
        bufs.append(buf); offs.append(off + base); lens_all.append(lens)
        base += buf.numel()
    del genome
    if not bufs:
        z = torch.zeros(0, dtype=torch.int64, device=device)
        return torch.zeros(0, dtype=torch.uint8, device=device), z, z.clone()
    return torch.cat(bufs), torch.cat(offs), torch.cat(lens_all)


def input_digest_parts(buf: torch.Tensor, lens: torch.Tensor, byte_base: int, read_base: int):
    """Two wrap-around int64 sums that identify the input independently of how it is split: every arena byte weighted by
    a mix of its GLOBAL byte offset, every read length weighted by a mix of its GLOBAL read id.  Summed over the ranks
    (mod 2^64) they give the same value for any number of GPUs."""
    dev = buf.device
    total = torch.zeros((), dtype=torch.int64, device=dev)
    step = 1 << 25
    for c0 in range(0, buf.numel(), step):
        b = buf[c0:c0 + step].to(torch.int64)
        idx = torch.arange(c0, c0 + b.numel(), device=dev, dtype=torch.int64) + byte_base
        w = (idx * -7046029254386353131) ^ (idx >> 11)          # 0x9E3779B97F4A7C15 as int64
        total += ((b + 1) * w).sum()
    ids = torch.arange(lens.numel(), device=dev, dtype=torch.int64) + read_base
    lsum = ((ids * -4417276706812531889) ^ (ids >> 7)) * (lens.to(torch.int64) + 1)
    return int(total.item()) & 0xFFFFFFFFFFFFFFFF, int(lsum.sum().item()) & 0xFFFFFFFFFFFFFFFF


def to_dnabuffer(buf: torch.Tensor, off: torch.Tensor, lens: torch.Tensor) -> DnaBuffer:
    return DnaBuffer(buf.cpu().numpy(), off.cpu().numpy().astype(np.uint64), lens.cpu().numpy().astype(np.uint64))


def make_dnabuffer(genome_len: int, n_reads: int, mean_len: float, sd_len: float, err: float, seed: int = 313, **kw) -> DnaBuffer:
    return to_dnabuffer(*make_reads(genome_len, n_reads, mean_len, sd_len, err, seed, device="cpu", **kw))
