"""One short device-resident evaluation (for ncu)."""
import os, sys
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import lineprofiles_b200 as klp
from lineprofiles_b200 import synthetic as syn

model = sys.argv[1] if len(sys.argv) > 1 else "kline5"
prec = sys.argv[2] if len(sys.argv) > 2 else "f32"
B = int(sys.argv[3]) if len(sys.argv) > 3 else 1184
n_bins = int(sys.argv[4]) if len(sys.argv) > 4 else 1000
log = model.startswith("kconv")
tab = syn.make_table()
gt = klp.Table.from_arrays(tab, 0)
for kv in os.environ.get("KLP_OPTIONS", "").split(","):   # e.g. KLP_OPTIONS=conv_mode=1,deposit_ring=1
    if kv:
        gt.set_option(kv.split("=")[0], int(kv.split("=")[1]))
E = syn.energy_grid_log(n_bins) if log else syn.energy_grid_linear(n_bins)
p = syn.random_params(model, B, seed=1)
dp, dE = torch.from_numpy(p).cuda(), torch.from_numpy(E).cuda()
out = torch.empty((B, n_bins), dtype=torch.float64, device="cuda")
spec = torch.from_numpy(syn.powerlaw_continuum(E)).cuda() if log else None
for _ in range(2):
    gt.eval_batch_device(model, dp, dE, n_bins, out, B, d_spectrum=spec, precision=prec, stream=torch.cuda.current_stream().cuda_stream)
torch.cuda.synchronize()
print("ok", float(out.sum()))
