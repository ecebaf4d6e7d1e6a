"""kconv A/B of the library selected by KLP_LIBRARY: 2000 sets on the 4096-bin log grid (BASELINE config 3 grid), both
convolution modes, per-kernel milliseconds from the library's CUDA events; plus a checksum so that two builds can be compared."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import numpy as np
from gpu_perf import run
import lineprofiles_b200 as klp
from lineprofiles_b200 import synthetic as syn

gt = klp.Table.from_arrays(syn.make_table(), 0)
print("library", klp.api.library_path())
E = syn.energy_grid_log(4096)
spec = syn.powerlaw_continuum(E)
p = syn.random_params("kconv10", 64, seed=5)
for mode in (0, 1):
    gt.set_option("conv_mode", mode)
    r = gt.eval_batch("kconv10", p, E, spectrum=spec, precision="f32")
    print(f"mode {mode} checksum {float(np.sum(r)):.17g} sumsq {float(np.sum(r * r)):.17g}")
for mode in (0, 1):
    run(gt, "kconv", 2000, "f32", n_bins=4096, log=True, opts={"conv_mode": mode}, reps=3)
try:
    gt.set_option("conv_mode", 1)
    gt.set_option("conv_src_smem", 0)
    r0 = gt.eval_batch("kconv10", p, E, spectrum=spec, precision="f32")
    gt.set_option("conv_src_smem", 1)
    r1 = gt.eval_batch("kconv10", p, E, spectrum=spec, precision="f32")
    print("conv_src_smem 1 == 0:", np.array_equal(r0, r1))
    El = syn.energy_grid_linear(1000)
    sl = np.tile(syn.powerlaw_continuum(El), (64, 1)) * np.linspace(1, 2, 64)[:, None]
    pl = syn.random_params("kconv", 64, seed=6)
    gt.set_option("conv_src_smem", 0)
    q0 = gt.eval_batch("kconv", pl, El, spectrum=sl, precision="f64")
    gt.set_option("conv_src_smem", 1)
    q1 = gt.eval_batch("kconv", pl, El, spectrum=sl, precision="f64")
    print("per-set spectra, 1000 linear bins, conv_src_smem 1 == 0:", np.array_equal(q0, q1))
    for ss in (0, 1):
        run(gt, "kconv", 2000, "f32", n_bins=4096, log=True, opts={"conv_mode": 1, "conv_src_smem": ss}, reps=3)
        run(gt, "kconv", 5000, "f32", n_bins=1000, log=False, opts={"conv_mode": 1, "conv_src_smem": ss}, reps=3)
    gt.set_option("conv_src_smem", 0)
except klp.KlpError as e:
    print("no conv_src_smem option in this build:", e)
run(gt, "kline5", 100000, "f32", reps=2)
