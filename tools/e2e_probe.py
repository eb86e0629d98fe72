"""Where the end-to-end step spends its time: host wall clock and device events around upload / run / get_B (tuning aid)."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from bench import load_workload
from elba_b200 import frontend
import ctypes as C
dev = torch.device("cuda", 0)
buf, off, lens, k, lo, up, total, r0 = load_workload("celegans40x_hifi", 0, 1, dev, float(os.environ.get("SCALE", "1.0")))
hbuf, hoff, hlen = (t.cpu().pin_memory() for t in (buf, off, lens))
torch.cuda.synchronize()
for own_stream in (False, True):
    ctx = frontend.Context(frontend.Params(k=k, lower=lo, upper=up))
    if not own_stream:
        ctx.set_stream(torch.cuda.current_stream().cuda_stream)
    for it in range(4):
        torch.cuda.synchronize(); t0 = time.perf_counter()
        ctx.upload_raw(hbuf.data_ptr(), hbuf.numel(), hoff.data_ptr(), hlen.data_ptr(), lens.numel(), 0)
        t1 = time.perf_counter()
        ctx.run()
        t2 = time.perf_counter()
        s = ctx.sizes(); n = s["nnzB"]
        rp = torch.empty(s["nreads"] + 1, dtype=torch.int64).pin_memory(); col = torch.empty(n, dtype=torch.int32).pin_memory()
        num = torch.empty(n, dtype=torch.int32).pin_memory(); seeds = torch.empty(4 * n, dtype=torch.int32).pin_memory()
        t3 = time.perf_counter()
        ctx._ck(ctx.L.elba_fe_get_B(ctx.h, C.c_void_p(rp.data_ptr()), C.c_void_p(col.data_ptr()), C.c_void_p(num.data_ptr()), C.c_void_p(seeds.data_ptr())))
        t4 = time.perf_counter()
        tm = ctx.timings()
        print(f"own_stream={own_stream} it={it}: upload call {1e3*(t1-t0):.2f} ms, run {1e3*(t2-t1):.2f}, alloc {1e3*(t3-t2):.2f}, get_B {1e3*(t4-t3):.2f}; device: upload {tm['upload_ms']:.2f} count {tm['count_ms']:.2f} "
              f"scatter {tm['partition_ms']:.2f} build {tm['build_ms']:.2f} spgemm {tm['spgemm_ms']:.2f} download {tm['download_ms']:.2f}", flush=True)
    ctx.close()
