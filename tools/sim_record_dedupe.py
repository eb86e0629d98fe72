"""How much would counting whole super-k-mer records before expanding them save?  

Host-side simulation, numpy only: reads of the C. elegans 40X HiFi shape (scaled genome), MAXIMAL super-k-mers (runs of
consecutive k-mers sharing their minimizer, cut at nmax k-mers counted from the run's start, NOT at a per-read chunk
grid), each record put on its canonical strand.  Reports the share of k-mer instances that sit in a record seen more than
once, i.e. the instances that a record-level pre-count would insert with a multiplicity instead of one by one."""
import sys, os, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from elba_b200.synth import make_dnabuffer

k, W = 31, 16
m = k - W + 1
nmax = 62 - k
genome, cov, mean, err = 300_000, 40, 14550, 0.01
dna = make_dnabuffer(genome_len=genome, n_reads=int(genome * cov / mean), mean_len=mean, sd_len=1000, err=err, seed=313)
rng_mask = (1 << (2 * m)) - 1
recs = {}
tot_inst = 0
tot_rec = 0
t0 = time.time()
for r in range(dna.size()):
    c = dna.read_codes(r).astype(np.int64)
    n = len(c)
    if n < k:
        continue
    # canonical m-mers and their hashed order
    f = np.zeros(n - m + 1, np.int64)
    for i in range(m):
        f = (f << 2) | c[i:n - m + 1 + i]
    rc = np.zeros(n - m + 1, np.int64)
    cc = 3 - c
    for i in range(m):
        rc = rc | (cc[i:n - m + 1 + i] << (2 * i))
    can = np.minimum(f, rc)
    h = ((can * 0x9E3779B1 + 0x7F4A7C15) & 0xFFFFFFFF)
    # minimizer of every k-mer: min over W consecutive m-mers
    nk = n - k + 1
    win = np.lib.stride_tricks.sliding_window_view(h, W)[:nk]
    mn = win.min(axis=1)
    cut = np.flatnonzero(np.r_[True, mn[1:] != mn[:-1]])
    ends = np.r_[cut[1:], nk]
    s = "".join("ACGT"[x] for x in c)
    comp = str.maketrans("ACGT", "TGCA")
    for a, b in zip(cut, ends):
        p = a
        while p < b:                      # cut at nmax k-mers, counted from the run's start
            q = min(p + nmax, b)
            seq = s[p:q + k - 1]
            rcs = seq[::-1].translate(comp)
            key = min(seq, rcs)
            e = recs.get(key)
            if e is None:
                recs[key] = [1, q - p]
            else:
                e[0] += 1
            tot_inst += q - p
            tot_rec += 1
            p = q
dup_inst = sum(v[1] * v[0] for v in recs.values() if v[0] > 1)
distinct_inst = sum(v[1] for v in recs.values())
print(f"reads {dna.size()}  instances {tot_inst}  records {tot_rec} ({tot_inst / tot_rec:.2f} k-mers each)  distinct records {len(recs)}")
print(f"instances inside records seen more than once: {100 * dup_inst / tot_inst:.1f} %")
print(f"k-mer insertions if every distinct record is expanded once (with its multiplicity): {distinct_inst} = {100 * distinct_inst / tot_inst:.1f} % of today's")
print(f"[{time.time() - t0:.0f} s]")
