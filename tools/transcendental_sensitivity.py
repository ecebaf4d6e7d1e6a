#!/usr/bin/env python
"""How far can a spectrum move when the transcendentals differ from ours by a few ulps?

The reference evaluates `std.math.pow(f32, ..)`, `@cos` and `std.math.log10` with Zig's own f32
implementations (emissivity.zig:12,45,52-53,67-75; xspec-wrapper.zig:224-226), which are not
correctly rounded and cannot be run here.  Oracle and CUDA path both evaluate them in f64 and round
to T.  This tool re-runs the oracle (CPU only) under perturbed policies -- every transcendental
result moved by +k / -k / pseudo-random [-k, k] ulps of T, the C library's T-precision functions,
and pow restated the way Zig / Go structure it (square-and-multiply + exp(yf*log x) in T) -- and
reports the largest per-bin relative change of the final spectrum,
    max_i |y'_i - y_i| / (|y_i| + 1e-12 max|y|),
on BASELINE configs 1-3.  Output: a markdown table and a JSON file.
"""
from __future__ import annotations

import argparse
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from lineprofiles_b200 import synthetic as syn  # noqa: E402
from oracle import oracle_py as orc  # noqa: E402


def changes(y1, y0):
    """(gate, robust, peak): max per-bin |dy| / (|y| + floor * max|y|) with floor 1e-12 (the parity-gate metric) and
    1e-3 (bins that carry at least a thousandth of the peak), and max |dy| / max|y|."""
    scale = np.max(np.abs(y0), axis=-1, keepdims=True)
    d = np.abs(y1 - y0)
    return (float(np.max(d / (np.abs(y0) + 1e-12 * scale))), float(np.max(d / (np.abs(y0) + 1e-3 * scale))),
            float(np.max(d / scale)))


def workloads(n2: int, n3: int):
    e_lin = syn.energy_grid_linear(1000)
    e_log = syn.energy_grid_log(4096)
    w = []
    w.append(("config 1: kline single (a=0.998, 30 deg, 6.4 keV, ISCO..400, alpha=3), 1000 bins", "kline",
              np.array([[0.998, 30.0, 6.4, float(syn.r_isco(0.998)), 400.0, 3.0]]), e_lin, None))
    w.append((f"config 2: kline5, {n2} random sets (seed 20260), 1000 bins", "kline5",
              syn.random_params("kline5", n2, seed=20260), e_lin, None))
    w.append((f"config 3: kconv, {n3} random sets (seed 20261), 4096 log bins, E^-2 continuum", "kconv",
              syn.random_params("kconv", n3, seed=20261), e_log, syn.powerlaw_continuum(e_log)))
    return w


ALL = ("pow", "cos", "log10")
POLICIES = [("random", 1, ALL), ("random", 2, ALL), ("random", 4, ALL), ("plus", 1, ALL), ("minus", 1, ALL),
            ("random", 4, ("pow", "log10")), ("random", 1, ("cos",)), ("random", 4, ("cos",)),
            ("libm_T", 0, ALL), ("zig_style_pow", 0, ALL)]


def label(m, k, f):
    s = m + (f" {k}" if k else "")
    return s if f == ALL else s + " [" + ",".join(f) + "]"


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n2", type=int, default=64)
    ap.add_argument("--n3", type=int, default=8)
    ap.add_argument("--threads", type=int, default=os.cpu_count() or 1)
    ap.add_argument("--out", default=os.path.join(ROOT, "profiles", "r02_transcendental_sensitivity.json"))
    args = ap.parse_args()
    ot = orc.OracleTable(syn.make_table())
    rows = []
    for name, model, p, energy, spec in workloads(args.n2, args.n3):
        for prec in ("f32", "f64"):
            base = ot.eval_batch(model, p, energy, spectrum=spec, precision=prec, threads=args.threads)["out"]
            for mode, k, funcs in POLICIES:
                with orc.transcendental_policy(mode, k, funcs):
                    y = ot.eval_batch(model, p, energy, spectrum=spec, precision=prec, threads=args.threads)["out"]
                g, r, pk = changes(y, base)
                rows.append({"workload": name, "precision": prec, "policy": label(mode, k, funcs),
                             "max_rel_change_gate_metric": g, "max_rel_change_bins_above_1e-3_of_peak": r,
                             "max_change_over_peak": pk})
                print(rows[-1], flush=True)
    json.dump(rows, open(args.out, "w"), indent=1)
    # markdown: gate metric / bins above 1e-3 of the peak / relative to the peak
    labels = [label(*p) for p in POLICIES]
    print("\n| workload | T | " + " | ".join(labels) + " |")
    print("|---|---|" + "---|" * len(labels))
    keys = []
    for r in rows:
        if (r["workload"], r["precision"]) not in keys:
            keys.append((r["workload"], r["precision"]))
    for wl, prec in keys:
        cells = []
        for lb in labels:
            r = next(r for r in rows if r["workload"] == wl and r["precision"] == prec and r["policy"] == lb)
            cells.append(f'{r["max_rel_change_gate_metric"]:.0e} / {r["max_rel_change_bins_above_1e-3_of_peak"]:.0e} / '
                         f'{r["max_change_over_peak"]:.0e}')
        print(f"| {wl.split(':')[0]} | {prec} | " + " | ".join(cells) + " |")


if __name__ == "__main__":
    main()
