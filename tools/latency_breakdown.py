"""Per-kernel device time of one small evaluation (library CUDA events)."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import lineprofiles_b200 as klp
from lineprofiles_b200 import synthetic as syn
gt = klp.Table.from_arrays(syn.make_table(), 0)
E = syn.energy_grid_linear(1000)
dE = torch.from_numpy(E).cuda()
for B in (1, 8, 64, 148, 296, 592, 1000, 2960):
    p = syn.random_params("kline5", B, seed=3)
    dp = torch.from_numpy(p).cuda()
    out = torch.empty((B, 1000), dtype=torch.float64, device="cuda")
    st = torch.cuda.current_stream().cuda_stream
    gt.set_timing(True)
    for _ in range(4):
        gt.eval_batch_device("kline5", dp, dE, 1000, out, B, precision="f32", stream=st)
        torch.cuda.synchronize()
        tm = gt.get_timing()
    print(B, {k: round(v["ms"] * 1e3, 1) for k, v in tm.items()}, "us")
