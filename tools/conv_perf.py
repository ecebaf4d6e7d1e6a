"""kconv throughput per convolution kernel (option conv_mode: 0 = k_conv_exact, 1 = k_conv_cum, 2 / 3 = the round-1 kernels
k_conv / k_conv_fast), and agreement between them: the two reference-order kernels must agree bit for bit."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import numpy as np
from gpu_perf import run
import lineprofiles_b200 as klp
from lineprofiles_b200 import synthetic as syn

gt = klp.Table.from_arrays(syn.make_table(), 0)
B = int(sys.argv[1]) if len(sys.argv) > 1 else 5000
E = syn.energy_grid_log(4096)
spec = syn.powerlaw_continuum(E)
p = syn.random_params("kconv10", 64, seed=5)
res = {}
for mode in (0, 2, 1, 3):
    gt.set_option("conv_mode", mode)
    res[mode] = gt.eval_batch("kconv10", p, E, spectrum=spec, precision="f32")
print("exact new == exact old:", np.array_equal(res[0], res[2]))
sc = np.max(np.abs(res[0]), axis=1, keepdims=True)
for m in (1, 3):
    print(f"mode {m} vs exact: max rel (floor 1e-12 peak) {np.max(np.abs(res[m] - res[0]) / (np.abs(res[0]) + 1e-12 * sc)):.3e}")
El = syn.energy_grid_linear(1000)
sl = syn.powerlaw_continuum(El)
pl = syn.random_params("kconv", 64, seed=6)
r = {}
for mode in (0, 2):
    gt.set_option("conv_mode", mode)
    r[mode] = gt.eval_batch("kconv", pl, El, spectrum=sl, precision="f32")
print("linear 1000-bin grid, exact new == exact old:", np.array_equal(r[0], r[2]))
for mode in (0, 2, 1, 3):
    run(gt, "kconv", B, "f32", n_bins=4096, log=True, opts={"conv_mode": mode})
for mode in (0, 2, 1, 3):
    run(gt, "kconv10", B, "f32", n_bins=4096, log=True, opts={"conv_mode": mode})
for mode in (0, 2, 1, 3):
    run(gt, "kconv", B, "f32", n_bins=1000, log=False, opts={"conv_mode": mode})
for mode in (0, 2, 1, 3):
    run(gt, "kconv", 1, "f32", n_bins=4096, log=True, opts={"conv_mode": mode}, reps=5)
