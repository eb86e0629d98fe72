"""Run the hot path of one workload under several environment settings in ONE process (one box, one read set) and print
the library's own per-phase CUDA-event timings for each.  Tuning aid, not a benchmark.

    python tools/tune_step.py --workload celegans40x_hifi --passes 3 "ELBA_FE_FUSE=0" "ELBA_FE_SKM_THREADS=256" ...

Each positional argument is a comma-separated list of NAME=VALUE pairs applied for that variant; "" is the default."""
import argparse, json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from bench import load_workload
from elba_b200 import frontend

ap = argparse.ArgumentParser()
ap.add_argument("--workload", default="celegans40x_hifi")
ap.add_argument("--scale", type=float, default=1.0)
ap.add_argument("--passes", type=int, default=3)
ap.add_argument("--align", action="store_true", help="also run the X-drop alignment of B's nonzeros once per variant and print its time")
ap.add_argument("variants", nargs="*", default=[""])
a = ap.parse_args()
dev = torch.device("cuda", 0)
t0 = time.time()
buf, off, lens, k, lo, up, total, r0 = load_workload(a.workload, 0, 1, dev, a.scale)
torch.cuda.synchronize()
print(f"[tune] workload {a.workload} x{a.scale}: {lens.numel()} reads, {buf.numel()} packed bytes, made in {time.time() - t0:.1f}s", flush=True)
keys = ("count_ms", "partition_ms", "count_kernel_ms", "build_ms", "lookup_ms", "spgemm_ms", "spgemm_kernel_ms")
first = None
for v in a.variants:
    env = dict(x.split("=", 1) for x in v.split(",") if x)
    old = {n: os.environ.get(n) for n in env}
    os.environ.update(env)
    try:
        ctx = frontend.Context(frontend.Params(k=k, lower=lo, upper=up))
        best = None
        for i in range(a.passes):
            ctx.set_reads_device(buf.data_ptr(), buf.numel(), off.data_ptr(), lens.data_ptr(), lens.numel(), 0)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            torch.cuda.synchronize()
            t1 = time.perf_counter()
            ctx.run()
            ctx.synchronize()
            wall = (time.perf_counter() - t1) * 1e3
            tm = ctx.timings()
            tm["wall_ms"] = wall
            if best is None or wall < best["wall_ms"]:
                best = tm
        sz = ctx.sizes()
        sig = (sz["reliable"], sz["nnzA"], sz["products"], sz["nnzB"])
        if first is None:
            first = sig
        print(f"[tune] {v or 'default':40s} wall {best['wall_ms']:8.2f} ms  " + "  ".join(f"{n[:-3]} {best.get(n, 0):7.2f}" for n in keys)
              + f"  launches {best.get('kernel_launches')}  buckets {sz['partitions']} spilled instances {sz['overflow_instances']}  sizes {sig} {'==' if sig == first else '!= FIRST VARIANT'}", flush=True)
        if a.align:
            t1 = time.perf_counter()
            rows, cols, out = ctx.align()
            wall = (time.perf_counter() - t1) * 1e3
            tm = ctx.timings()
            print(f"[tune] {v or 'default':40s} align: {len(rows)} pairs, device {tm['align_ms']:.1f} ms (wall incl. D2H {wall:.1f} ms), "
                  f"{len(rows) / max(tm['align_ms'], 1e-9) * 1e3:.0f} pairs/s, passed {int(out[:, 6].sum())}, mean score {out[:, 4].mean() if len(rows) else 0:.1f}", flush=True)
        ctx.close()
    except Exception as ex:  # keep going: the other variants still tell something
        print(f"[tune] {v or 'default':40s} FAILED: {ex}", flush=True)
    finally:
        for n, o in old.items():
            if o is None:
                os.environ.pop(n, None)
            else:
                os.environ[n] = o
