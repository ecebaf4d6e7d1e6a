"""Per-kernel device time of small evaluations (library CUDA events), host-API latency, and the exported-symbol latency."""
import os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import lineprofiles_b200 as klp
from lineprofiles_b200 import synthetic as syn
gt = klp.Table.from_arrays(syn.make_table(), 0)
E = syn.energy_grid_linear(1000)
dE = torch.from_numpy(E).cuda()
st = torch.cuda.current_stream().cuda_stream
for coop in (1, 0):
    gt.set_option("coop", coop)
    for B in (1, 8, 37, 64, 148, 296, 1000):
        p = syn.random_params("kline5", B, seed=3)
        dp = torch.from_numpy(p).cuda()
        out = torch.empty((B, 1000), dtype=torch.float64, device="cuda")
        gt.set_timing(True)
        for _ in range(4):
            gt.eval_batch_device("kline5", dp, dE, 1000, out, B, precision="f32", stream=st)
            torch.cuda.synchronize()
            tm = gt.get_timing()
        gt.set_timing(False)
        n = 50
        t0 = time.perf_counter()
        for _ in range(n):
            gt.eval_batch("kline5", p, E, precision="f32")
        dt = (time.perf_counter() - t0) / n
        print(f"coop={coop} B={B:5d}", {k: round(v["ms"] * 1e3, 1) for k, v in tm.items()}, f"us | host API {dt*1e6:8.1f} us per call | {gt.last_launch_info()}", flush=True)
gt.set_option("coop", 1)
# the exported XSPEC symbols (what XSPEC calls): kline on the 1000-bin grid, kconv on 4096 log bins
klp.set_default_table(gt)
El = syn.energy_grid_log(4096)
spec = syn.powerlaw_continuum(El)
flux = np.zeros(1000)
for name, fn, en, par, buf in (("kline", klp.kline, E, syn.CONFIG1_PARAMS[0], flux), ("kconv", klp.kconv, El, syn.random_params("kconv", 1, seed=3)[0], None)):
    for _ in range(20):
        fn(en, par, buf if buf is not None else spec.copy())
    n = 300
    bufs = [spec.copy() for _ in range(n)] if buf is None else None
    t0 = time.perf_counter()
    for i in range(n):
        fn(en, par, buf if buf is not None else bufs[i])
    dt = (time.perf_counter() - t0) / n
    print(f"exported {name}: {dt*1e6:.1f} us per call", flush=True)
klp.set_default_table(None)
