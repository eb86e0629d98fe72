"""Per-bucket statistics of the super-k-mer path on the C. elegans 40X HiFi shape (host simulation, numpy): how many
instances, DISTINCT k-mers, records and distinct records a minimizer bucket holds when the mean fill is 2457 instances.
Evidence for sizing the shared-memory table by distinct k-mers instead of by instances."""
import sys, os, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from elba_b200.synth import make_dnabuffer

k, W = 31, 16
m = k - W + 1
genome, cov, mean, err = 400_000, 40, 14550, 0.01
dna = make_dnabuffer(genome_len=genome, n_reads=int(genome * cov / mean), mean_len=mean, sd_len=1000, err=err, seed=313)
t0 = time.time()
all_b, all_k = [], []
for r in range(dna.size()):
    c = dna.read_codes(r).astype(np.uint64)
    n = len(c)
    if n < k:
        continue
    def pack(width):
        f = np.zeros(n - width + 1, np.uint64)
        rc = np.zeros(n - width + 1, np.uint64)
        cc = np.uint64(3) - c
        for i in range(width):
            f = (f << np.uint64(2)) | c[i:n - width + 1 + i]
            rc = rc | (cc[i:n - width + 1 + i] << np.uint64(2 * i))
        return np.minimum(f, rc)
    can_m = pack(m)
    h = (can_m * np.uint64(0x9E3779B1) + np.uint64(0x7F4A7C15)) & np.uint64(0xFFFFFFFF)
    nk = n - k + 1
    mn = np.lib.stride_tricks.sliding_window_view(h, W)[:nk].min(axis=1)
    all_b.append(mn)
    all_k.append(pack(k))
mn = np.concatenate(all_b); km = np.concatenate(all_k)
M = len(mn)
NB = max(1, M // 2457)
v = mn.astype(np.uint64)
v ^= v >> np.uint64(15); v = (v * np.uint64(0x2C1B3C6D)) & np.uint64(0xFFFFFFFF); v ^= v >> np.uint64(12)
v = (v * np.uint64(0x297A2D39)) & np.uint64(0xFFFFFFFF); v ^= v >> np.uint64(15)
b = ((v * np.uint64(NB)) >> np.uint64(32)).astype(np.int64)
inst = np.bincount(b, minlength=NB)
order = np.lexsort((km, b))
bs, ks = b[order], km[order]
new = np.r_[True, (bs[1:] != bs[:-1]) | (ks[1:] != ks[:-1])]
distinct = np.bincount(bs[new], minlength=NB)
q = lambda a: [int(np.percentile(a, p)) for p in (50, 90, 99, 99.9, 100)]
print(f"M = {M} instances, {NB} buckets (mean {M / NB:.0f} instances)   [{time.time() - t0:.0f} s]")
print(f"instances per bucket       p50/p90/p99/p99.9/max: {q(inst)}")
print(f"distinct k-mers per bucket p50/p90/p99/p99.9/max: {q(distinct)}   mean {distinct.mean():.0f}   D/M = {distinct.sum() / M:.3f}")
for slots in (1024, 2048, 4096):
    cap = slots * 3 // 4
    print(f"  a {slots}-slot table (<= {cap} distinct k-mers) holds {100 * (distinct <= cap).mean():.2f} % of the buckets, "
          f"{100 * inst[distinct <= cap].sum() / M:.2f} % of the instances")
