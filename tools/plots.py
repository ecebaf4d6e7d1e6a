#!/usr/bin/env python
"""The reference's plot script (src/plots.zig:109-176) through the exported XSPEC symbols of the B200 library:
  additive-example  kline   {0.998, 70, 6.4, 3.0, 0.0, 50.0}                                on 300 edges 0.1-12 keV
  additive5-example kline5  {0.998, 80, 6.4, 1.0, 50.0, 3.0, 1.0, 2.0, 4.0, 5.0, 2.0, ...}  on 1000 edges
  convolve-example  kconv   {0.998, 40, 3.0, 0.0, 50.0} applied to two delta lines (bins 150 and 200), then normalised
(the parameter vectors are the reference's, passed as they are -- including their orderings).  Writes one SVG and one CSV
per example (no plotting library is needed: the SVG is a poly-line).  Needs a CUDA device and a table: by default the
synthetic Kerr-shaped one, or --table PATH (kerr-transfer-functions.fits / .klpt).

usage: python tools/plots.py [--table PATH] [--out DIR]"""
from __future__ import annotations

import argparse
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import lineprofiles_b200 as klp  # noqa: E402
from lineprofiles_b200 import synthetic as syn  # noqa: E402


def range_iterator(lo: float, hi: float, n: int) -> np.ndarray:
    """utils.RangeIterator(f64): cumulative-add linspace (utils.zig:121-156)."""
    out = np.empty(n)
    delta = (hi - lo) / (n - 1)
    cur = lo
    for i in range(n):
        out[i] = cur
        cur += delta
    return out


def write_svg(path: str, x, y, title: str, w: int = 720, h: int = 360, pad: int = 45) -> None:
    x, y = np.asarray(x, float), np.nan_to_num(np.asarray(y, float))
    x0, x1, y1 = x.min(), x.max(), max(float(y.max()), 1e-300)
    px = pad + (x - x0) / (x1 - x0) * (w - 2 * pad)
    py = h - pad - y / y1 * (h - 2 * pad)
    pts = " ".join(f"{a:.1f},{b:.1f}" for a, b in zip(px, py))
    with open(path, "w") as f:
        f.write(f'<svg xmlns="http://www.w3.org/2000/svg" width="{w}" height="{h}" font-family="sans-serif" font-size="11">\n'
                f'<rect width="{w}" height="{h}" fill="white"/>\n<text x="{pad}" y="18">{title}</text>\n'
                f'<line x1="{pad}" y1="{h - pad}" x2="{w - pad}" y2="{h - pad}" stroke="black"/>\n'
                f'<line x1="{pad}" y1="{pad}" x2="{pad}" y2="{h - pad}" stroke="black"/>\n'
                f'<text x="{pad}" y="{h - pad + 16}">{x0:g} keV</text><text x="{w - pad - 40}" y="{h - pad + 16}">{x1:g} keV</text>\n'
                f'<text x="4" y="{pad + 4}">{y1:.3g}</text>\n'
                f'<polyline fill="none" stroke="#1f4e9c" stroke-width="1.2" points="{pts}"/>\n</svg>\n')


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--table", default=None)
    ap.add_argument("--out", default="plots")
    ap.add_argument("--device", type=int, default=0)
    a = ap.parse_args()
    os.makedirs(a.out, exist_ok=True)
    table = klp.Table.load(a.table, a.device) if a.table else klp.Table.from_arrays(syn.make_table(), a.device)
    klp.set_default_table(table)
    try:
        jobs = []
        e = range_iterator(0.1, 12.0, 300)
        jobs.append(("additive-example", e, klp.kline(e, [0.998, 70, 6.4, 3.0, 0.0, 50.0])))       # (the reference's vector, as it is)
        klp.legacy_reset()
        jobs.append(("additive-example-config1", e, klp.kline(e, syn.CONFIG1_PARAMS[0])))          # a = 0.998, 30 deg, ISCO..400, alpha = 3
        klp.legacy_reset()
        e = range_iterator(0.1, 12.0, 1000)
        jobs.append(("additive5-example", e, klp.kline5(e, [0.998, 80, 6.4, 1.0, 50.0, 3.0, 1.0, 2.0, 4.0, 5.0, 2.0, 0.0])))
        klp.legacy_reset()
        e = range_iterator(0.1, 12.0, 300)
        flux = np.zeros(299)
        flux[150] = flux[200] = 1.0
        klp.kconv(e, [0.998, 40, 3.0, 0.0, 50.0], flux)
        tot = flux.sum()
        jobs.append(("convolve-example", e, flux / tot if tot else flux))     # utils.normalize (the reference's vector: rmax = 0 -> empty)
        klp.legacy_reset()
        flux = np.zeros(299)
        flux[150] = flux[200] = 1.0
        klp.kconv(e, [0.998, 40, float(syn.r_isco(0.998)), 400.0, 3.0], flux)   # the same two delta lines, today's parameter order
        tot = flux.sum()
        jobs.append(("convolve-example-isco-400", e, flux / tot if tot else flux))
        for name, e, y in jobs:
            np.savetxt(os.path.join(a.out, name + ".csv"), np.column_stack([e[:-1], y]), delimiter=",", header="energy_keV,flux")
            write_svg(os.path.join(a.out, name + ".svg"), e[:-1], y, name)
            if np.all(np.isnan(y)):
                # e.g. the reference's `additive-example` vector read with today's parameter order has rmax = 0: no radius is
                # integrated, the normalisation divides by zero (Q10) -- the reference returns the same NaNs
                print(f"{name}: {y.size} bins, ALL NaN (empty radial range -> sum 0 -> NaN, as in the reference) -> {a.out}/{name}.svg")
            else:
                print(f"{name}: {y.size} bins, sum {np.nansum(y):.6g}, peak at {e[int(np.nanargmax(y))]:.3f} keV -> {a.out}/{name}.svg")
    finally:
        klp.set_default_table(None)
        klp.legacy_reset()
        table.close()


if __name__ == "__main__":
    main()
