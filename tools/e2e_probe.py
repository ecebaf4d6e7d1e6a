"""End-to-end (host buffers) throughput of Table.eval_batch against the chunk size of the host pipeline."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import lineprofiles_b200 as klp
from lineprofiles_b200 import synthetic as syn
gt = klp.Table.from_arrays(syn.make_table(), 0)
E = syn.energy_grid_linear(1000)
B = 100000
ps = [syn.random_params("kline5", B, seed=1000 + s) for s in range(3)]
out = np.empty((B, 1000), np.float64)
tails = [int(x) for x in sys.argv[1].split(",")] if len(sys.argv) > 1 else [2048]
for chunk, tail in [(c, t) for t in tails for c in (64, 128, 256, 512)]:
    gt.set_option("stage_mb", chunk)
    gt.set_option("stage_tail", tail)
    gt.eval_batch("kline5", ps[0][:4096], E, precision="f32")
    gt.eval_batch("kline5", ps[0], E, precision="f32", out=out)
    t0 = time.perf_counter()
    for p in ps[1:]:
        gt.eval_batch("kline5", p, E, precision="f32", out=out)
    dt = (time.perf_counter() - t0) / 2
    print(f"stage_mb {chunk:6d} stage_tail {tail:6d}: {B/dt:9.0f} spectra/s end to end", flush=True)
