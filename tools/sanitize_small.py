"""Small all-kernel workload for compute-sanitizer: every model, both precisions, split and unsplit CTAs, scratch path."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import lineprofiles_b200 as klp
from lineprofiles_b200 import synthetic as syn

gt = klp.Table.from_arrays(syn.make_table(6, 5, 40, 30), 0)
E, El = syn.energy_grid_linear(100), syn.energy_grid_log(300)
spec = syn.powerlaw_continuum(El)
tot = 0.0
for prec in ("f32", "f64"):
    for model in ("kline", "kline5", "kconv", "kconv5", "kconv10"):
        conv = model.startswith("kconv")
        p = syn.random_params(model, 3, seed=5)
        out = gt.eval_batch(model, p, El if conv else E, spectrum=spec if conv else None, precision=prec)
        tot += float(np.nansum(out))
gt.set_option("splits", 1)
tot += float(gt.eval_batch("kline5", syn.random_params("kline5", 3, seed=6), E).sum())
big = klp.Table.from_arrays(syn.make_table(3, 3, 700, 48), 0)   # frame does not fit shared memory -> scratch path
tot += float(big.eval_batch("kline", syn.random_params("kline", 2, seed=7), E).sum())
print("ok", tot)
