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
# deposit ring (consumer warps), with and without discarding the consumed lines; both precisions
for disc in (1, 0):
    gt.set_option("deposit_ring", 1); gt.set_option("ring_discard", disc)
    for prec in ("f32", "f64"):
        tot += float(gt.eval_batch("kline5", syn.random_params("kline5", 3, seed=8), E, precision=prec).sum())
gt.set_option("deposit_ring", 0); gt.set_option("splits", 0)
# cumulative-profile convolution, per-source pairs in shared memory and through L1; per-set spectra
for ss in (1, 0):
    gt.set_option("conv_mode", 1); gt.set_option("conv_src_smem", ss)
    tot += float(np.nansum(gt.eval_batch("kconv5", syn.random_params("kconv5", 3, seed=9), El, spectrum=np.tile(spec, (3, 1)))))
gt.set_option("conv_mode", 0); gt.set_option("conv_src_smem", 1)
# host pipeline with several staging chunks and a short final one
gt.set_option("stage_mb", 1); gt.set_option("stage_tail", 5)
tot += float(gt.eval_batch("kline", syn.random_params("kline", 3000, seed=10)[:2700], E).sum())
gt.set_option("stage_mb", 256); gt.set_option("stage_tail", 4096)
big = klp.Table.from_arrays(syn.make_table(3, 3, 700, 48), 0)   # frame does not fit shared memory -> scratch path
tot += float(big.eval_batch("kline", syn.random_params("kline", 2, seed=7), E).sum())
print("ok", tot)
