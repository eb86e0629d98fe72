"""One warm-up + one profiled pass of the hot path (for ncu).  Not a benchmark: numbers under a profiler are never bench values."""
import argparse, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from bench import load_workload
from elba_b200 import frontend

ap = argparse.ArgumentParser()
ap.add_argument("--workload", default="ecoli30x_clr")
ap.add_argument("--scale", type=float, default=1.0)
ap.add_argument("--partitions", type=int, default=0)
ap.add_argument("--passes", type=int, default=2)
ap.add_argument("--align", action="store_true")
a = ap.parse_args()
dev = torch.device("cuda", 0)
buf, off, lens, k, lo, up, total, r0 = load_workload(a.workload, 0, 1, dev, a.scale)
ctx = frontend.Context(frontend.Params(k=k, lower=lo, upper=up, num_partitions=a.partitions))
for i in range(a.passes):
    ctx.set_reads_device(buf.data_ptr(), buf.numel(), off.data_ptr(), lens.data_ptr(), lens.numel(), 0)
    ctx.run()
if a.align:
    ctx.align()
print(ctx.sizes(), ctx.timings())
