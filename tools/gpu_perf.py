"""Throughput probe: device-resident batches, kernel timings from the library's CUDA events."""
import os, sys, time
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import lineprofiles_b200 as klp
from lineprofiles_b200 import synthetic as syn


def run(gt, model, B, prec, n_bins=1000, log=False, reps=2, opts=None):
    for k, v in (opts or {}).items():
        gt.set_option(k, v)
    E = syn.energy_grid_log(n_bins) if log else syn.energy_grid_linear(n_bins)
    p = syn.random_params(model, B, seed=1)
    dp = torch.from_numpy(p).cuda()
    dE = torch.from_numpy(E).cuda()
    out = torch.empty((B, n_bins), dtype=torch.float64, device="cuda")
    spec = torch.from_numpy(syn.powerlaw_continuum(E)).cuda() if model.startswith("kconv") else None
    st = torch.cuda.current_stream().cuda_stream
    gt.set_timing(True)
    best = None
    for r in range(reps + 1):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        gt.eval_batch_device(model, dp, dE, n_bins, out, B, d_spectrum=spec, precision=prec, stream=st)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        tm = gt.get_timing()
        if r > 0 and (best is None or dt < best[0]):
            best = (dt, tm)
    dt, tm = best
    gt.set_timing(False)
    s = " ".join(f"{k}={v['ms']:.1f}ms/{v['launches']}" for k, v in tm.items())
    print(f"{model:8s} {prec} B={B} n={n_bins} opts={opts} : {B/dt:10.0f} spectra/s ({dt*1e3:.1f} ms) | {s}", flush=True)
    return B / dt


if __name__ == "__main__":
    tab = syn.make_table()
    gt = klp.Table.from_arrays(tab, 0)
    B = int(sys.argv[1]) if len(sys.argv) > 1 else 20000
    print("fma peak GFLOP/s f32", gt.fma_peak_gflops("f32"), "f64", gt.fma_peak_gflops("f64"))
    for nt in (256, 512, 1024):
        run(gt, "kline5", B, "f32", opts={"quad_threads": nt})
    for nt in (256, 512, 1024):
        run(gt, "kline5", B // 2, "f64", opts={"quad_threads": nt})
    gt.set_option("quad_threads", 0)
    run(gt, "kline", B, "f32")
    run(gt, "kconv", 2000, "f32", n_bins=4096, log=True)
    run(gt, "kconv10", 2000, "f32", n_bins=4096, log=True)
    run(gt, "kline5", 1, "f32")
    run(gt, "kline5", 16, "f32")
    run(gt, "kconv5", 1, "f32", n_bins=4096, log=True)
