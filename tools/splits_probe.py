"""k_quad device time against the number of CTAs sharing a set (option splits), for small and medium batches."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import lineprofiles_b200 as klp
from lineprofiles_b200 import synthetic as syn
gt = klp.Table.from_arrays(syn.make_table(), 0)
E = syn.energy_grid_linear(1000)
dE = torch.from_numpy(E).cuda()
st = torch.cuda.current_stream().cuda_stream
for B in (16, 64, 148, 296, 592, 1000, 2000, 4000):
    p = syn.random_params("kline5", B, seed=3)
    dp = torch.from_numpy(p).cuda()
    out = torch.empty((B, 1000), dtype=torch.float64, device="cuda")
    res = []
    for splits in (0, 1, 2, 3, 4, 6, 8, 12, 16, 32, 64):
        gt.set_option("splits", splits)
        gt.set_timing(True)
        best = 1e9
        for _ in range(3):
            gt.eval_batch_device("kline5", dp, dE, 1000, out, B, precision="f32", stream=st)
            torch.cuda.synchronize()
            best = min(best, gt.get_timing()["quad"]["ms"])
        res.append((splits, round(best * 1e3)))
    print(B, res, flush=True)
