"""Does the order of the sets inside a launch matter?  Same sets, random order vs sorted by inclination (a cost proxy)."""
import os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import lineprofiles_b200 as klp
from lineprofiles_b200 import synthetic as syn
gt = klp.Table.from_arrays(syn.make_table(), 0)
E = syn.energy_grid_linear(1000)
dE = torch.from_numpy(E).cuda()
st = torch.cuda.current_stream().cuda_stream
for B in (2000, 8388, 16384, 50000):
    p = syn.random_params("kline5", B, seed=11)
    out = torch.empty((B, 1000), dtype=torch.float64, device="cuda")
    for name, pp in (("random", p), ("incl descending", p[np.argsort(-p[:, 1])]), ("incl ascending", p[np.argsort(p[:, 1])])):
        dp = torch.from_numpy(np.ascontiguousarray(pp)).cuda()
        gt.set_timing(True)
        best = 1e9
        for _ in range(3):
            gt.eval_batch_device("kline5", dp, dE, 1000, out, B, precision="f32", stream=st)
            torch.cuda.synchronize()
            best = min(best, gt.get_timing()["quad"]["ms"])
        print(f"B={B:6d} {name:16s}: {best:8.2f} ms  {B / best:8.1f} k spectra/s", flush=True)
