"""Diagnostic: run every model in both precisions on the GPU, compare with the oracle, print timings."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import lineprofiles_b200 as klp
from lineprofiles_b200 import synthetic as syn
from oracle import oracle_py as orc


def relerr(got, want):
    scale = np.max(np.abs(want), axis=-1, keepdims=True)
    return np.abs(got - want) / (np.abs(want) + 1e-12 * scale)


def main():
    nb = int(sys.argv[1]) if len(sys.argv) > 1 else 6
    tab = syn.make_table()
    gt = klp.Table.from_arrays(tab, 0)
    ot = orc.OracleTable(tab)
    E = syn.energy_grid_linear(1000)
    El = syn.energy_grid_log(512)
    spec = syn.powerlaw_continuum(El)
    for prec in ("f32", "f64"):
        for model in ("kline", "kline5", "kconv", "kconv5", "kconv10"):
            p = syn.random_params(model, nb, seed=100)
            if model == "kline":
                p[0] = syn.CONFIG1_PARAMS[0]
            conv = model.startswith("kconv")
            en = El if conv else E
            sp = spec if conv else None
            t0 = time.time()
            got = gt.eval_batch(model, p, en, spectrum=sp, precision=prec)
            tg = time.time() - t0
            t0 = time.time()
            ref = ot.eval_batch(model, p, en, spectrum=sp, precision=prec, threads=8, want_fine=True, want_profile=conv)
            tc = time.time() - t0
            fine, prof = gt.eval_debug(model, p, en, precision=prec)
            fe = np.array_equal(fine, ref["fine"])
            nfd = int(np.sum(fine != ref["fine"]))
            err = relerr(got, ref["out"])
            msg = f"{prec} {model:8s} B={nb} gpu {tg*1e3:8.1f} ms cpu {tc*1e3:8.1f} ms | fine bit-equal {fe} (ndiff {nfd}) | out max rel {np.nanmax(err):.3e} bit-equal {np.array_equal(got, ref['out'], equal_nan=True)}"
            if conv:
                msg += f" | profile bit-equal {np.array_equal(prof, ref['profile'], equal_nan=True)}"
            print(msg, flush=True)
            if not fe:
                s, b = np.argwhere(fine != ref["fine"])[0]
                print("   first fine diff set", s, "bin", b, fine[s, b], ref["fine"][s, b], "params", p[s])


if __name__ == "__main__":
    main()
