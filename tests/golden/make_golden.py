"""Generate tests/golden/oracle_golden.npz.

The reference (Zig) cannot be executed in this environment, so these vectors come from the CPU
oracle (oracle/klp_oracle.cpp) on the synthetic 30x20x150x30 table; they pin the oracle against
regressions and travel to the GPU box, where the CUDA path is compared with them directly.
    python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from lineprofiles_b200 import synthetic as syn  # noqa: E402
from oracle import oracle_py as orc  # noqa: E402


def main():
    orc.build()
    ot = orc.OracleTable(syn.make_table())
    e_line = syn.energy_grid_linear(1000)
    e_conv = syn.energy_grid_log(384)
    spec = syn.powerlaw_continuum(e_conv)
    out = {"energy_line": e_line, "energy_conv": e_conv, "spectrum": spec}
    for model in ("kline", "kline5", "kconv", "kconv5", "kconv10"):
        p = syn.random_params(model, 3, seed=424242)
        if model == "kline":
            p[0] = syn.CONFIG1_PARAMS[0]
        out[f"params_{model}"] = p
        conv = model.startswith("kconv")
        for prec in ("f32", "f64"):
            r = ot.eval_batch(model, p, e_conv if conv else e_line, spectrum=spec if conv else None, precision=prec,
                              faithful_cost=True)
            out[f"out_{model}_{prec}"] = r["out"]
    np.savez_compressed(os.path.join(os.path.dirname(os.path.abspath(__file__)), "oracle_golden.npz"), **out)
    print("wrote oracle_golden.npz")


if __name__ == "__main__":
    main()
