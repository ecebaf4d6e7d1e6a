#!/usr/bin/env python
"""Random-parameter stress + timing harness, after the reference's src/fuzz.zig:63-118: time one kline and one
kconv call through the exported symbols, then hammer kconv5 with random parameters (same ranges as
fuzz.zig:100-111) -- here in batches, checking what the reference only eyeballs: finiteness and the
Sum = 1 normalisation of the underlying profile.  Usage: python tools/fuzz.py [n_batches] [batch]"""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import lineprofiles_b200 as klp
from lineprofiles_b200 import synthetic as syn


def main():
    n_batches = int(sys.argv[1]) if len(sys.argv) > 1 else 10
    B = int(sys.argv[2]) if len(sys.argv) > 2 else 4096
    gt = klp.Table.from_arrays(syn.make_table(), 0)
    klp.set_default_table(gt)
    E = syn.range_iterator(0.1, 12.0, 301)           # TestSetup.init(allocator, 0.1, 12.0, 300)
    for name, fn, p in (("1. Additive", klp.kline, [0.998, 40, 6.4, 3.0, 1.0, 50.0]),
                        ("2. Convolutional", klp.kconv, [0.998, 40, 3.0, 1.0, 50.0])):
        flux = np.zeros(300)
        flux[150] = flux[200] = 1
        klp.legacy_reset()
        t0 = time.perf_counter()
        fn(E, np.array(p), flux)
        print(f" {name}\n   elapsed time: {1e6 * (time.perf_counter() - t0):.0f}us")
    print(" 3. Convolutional Five (batched)")
    rng = np.random.default_rng(0)
    spec = np.zeros(300)
    spec[150] = spec[200] = 1
    bad = 0
    for it in range(n_batches):
        u = rng.random((B, 11))
        p = np.stack([u[:, 0] * 0.998, u[:, 1] * 90, u[:, 2] + 1.0, u[:, 3] * 500, u[:, 4] * 5.0, u[:, 5] * 5.0] +
                     [u[:, k] * 5.0 for k in range(6, 11)], axis=1)
        t0 = time.perf_counter()
        out = gt.eval_batch("kconv5", p, E, spectrum=spec)
        dt = time.perf_counter() - t0
        _, prof = gt.eval_debug("kconv5", p, E)
        nan_sets = np.isnan(prof).any(axis=1)               # rmax < rmin etc. -> empty profile -> NaN (Q10), as the reference
        ok = np.isfinite(out[~nan_sets]).all() and np.allclose(prof[~nan_sets].sum(axis=1), 1.0, rtol=1e-9)
        bad += int(not ok)
        print(f"   batch {it}: {B / dt:9.0f} evaluations/s, empty-profile sets {int(nan_sets.sum())}, {'ok' if ok else 'FAILED'}")
    klp.set_default_table(None)
    sys.exit(1 if bad else 0)


if __name__ == "__main__":
    main()
