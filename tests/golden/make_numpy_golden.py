"""Generate tests/golden/numpy_golden.npz: the f32 spectra of the parameter sets of oracle_golden.npz, computed by the
independently written numpy restatement (oracle/numpy_check.py) -- NOT by the C++ oracle.  tests/test_oracle.py requires the
two fixture files (and the live oracle) to agree bit for bit; the CUDA path is compared with oracle_golden.npz on the GPU box.
    python tests/golden/make_numpy_golden.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from lineprofiles_b200 import synthetic as syn  # noqa: E402
from oracle import numpy_check as nc  # noqa: E402

F = np.float32


def spectrum(tab, model, p, energy, spec):
    """One parameter set through the numpy restatement; parameter layouts of xspec-wrapper.zig:218-310."""
    if model == "kline":      # {a, incl, eline, rmin, rmax, alpha}
        fine, g = nc.fine_flux(tab, p[0], p[1], p[2], p[3], p[4], nc.emissivity_fn(p[5], F=F), energy[0], energy[-1], F=F)
        return nc.rebin(fine, g, energy, p[2], F=F)
    if model == "kline5":     # {a, incl, eline, rmin, rmax, rcut, alpha, e1..e5}
        fine, g = nc.fine_flux(tab, p[0], p[1], p[2], p[3], p[4], nc.emissivity_fn(p[6], p[7:], p[5], F=F), energy[0], energy[-1], F=F)
        return nc.rebin(fine, g, energy, p[2], F=F)
    # kconv {a, incl, rmin, rmax, alpha}; kconv5 / kconv10 {a, incl, rmin, rmax, rcut, alpha, e1..eN}
    eps = nc.emissivity_fn(p[4], F=F) if model == "kconv" else nc.emissivity_fn(p[5], p[6:], p[4], F=F)
    cg = nc.conv_grid()
    fine, g = nc.fine_flux(tab, p[0], p[1], 1.0, p[2], p[3], eps, cg[0], cg[-1], F=F)
    return nc.convolve(cg, nc.rebin(fine, g, cg, 1.0, F=F), energy, spec)


def main():
    here = os.path.dirname(os.path.abspath(__file__))
    gold = np.load(os.path.join(here, "oracle_golden.npz"))
    tab = syn.make_table()
    out = {}
    for model in ("kline", "kline5", "kconv", "kconv5", "kconv10"):
        conv = model.startswith("kconv")
        energy = gold["energy_conv"] if conv else gold["energy_line"]
        out[f"out_{model}_f32"] = np.stack([spectrum(tab, model, p, energy, gold["spectrum"]) for p in gold[f"params_{model}"]])
        print(model, "done", flush=True)
    np.savez_compressed(os.path.join(here, "numpy_golden.npz"), **out)
    print("wrote numpy_golden.npz")


if __name__ == "__main__":
    main()
