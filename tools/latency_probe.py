"""Single-call latency of the exported XSPEC symbols and of small batches (the reference's own use: one evaluation per call)."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import lineprofiles_b200 as klp
from lineprofiles_b200 import synthetic as syn

tab = syn.make_table()
gt = klp.Table.from_arrays(tab, 0)
E = syn.energy_grid_linear(1000)
El = syn.energy_grid_log(4096)
spec = syn.powerlaw_continuum(El)
for tm in (0,):   # (historical loop over team modes)
    for model, B, en, sp in (("kline", 1, E, None), ("kline5", 1, E, None), ("kline5", 8, E, None), ("kline5", 64, E, None),
                             ("kline5", 296, E, None), ("kline5", 1000, E, None), ("kconv5", 1, El, spec)):
        p = syn.random_params(model, B, seed=3)
        if model == "kline":
            p[0] = syn.CONFIG1_PARAMS
        for _ in range(3):
            gt.eval_batch(model, p, en, spectrum=sp, precision="f32")
        n = 20
        t0 = time.perf_counter()
        for _ in range(n):
            gt.eval_batch(model, p, en, spectrum=sp, precision="f32")
        dt = (time.perf_counter() - t0) / n
        print(f"{model:7s} B={B:5d}: {dt*1e3:8.3f} ms per call ({B/dt:9.0f} spectra/s)", flush=True)
