"""ctypes front-end of the CPU oracle (oracle/klp_oracle.cpp).  TEST INFRASTRUCTURE ONLY:
imported by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "libklp_oracle.so")

MODELS = {"kline": 0, "kline5": 1, "kconv": 2, "kconv5": 3, "kconv10": 4}
PREC = {"f32": 0, "f64": 1}
N_FINE = 1999
N_CONV = 1999


def build(force: bool = False) -> str:
    src = os.path.join(_HERE, "klp_oracle.cpp")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-s"])
    return _SO


_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_SO):
            build()
        L = C.CDLL(_SO)
        dp, fp, zp = C.POINTER(C.c_double), C.POINTER(C.c_float), C.POINTER(C.c_size_t)
        L.orc_table_create.restype = C.c_void_p
        L.orc_table_create.argtypes = [fp, C.c_int, fp, C.c_int, C.c_int, C.c_int, fp, fp, fp, fp, fp]
        L.orc_table_destroy.argtypes = [C.c_void_p]
        L.orc_table_index.restype = C.c_size_t
        L.orc_table_index.argtypes = [C.c_void_p, C.c_size_t, C.c_size_t]
        L.orc_find_tables.argtypes = [C.c_void_p, zp, zp]
        L.orc_find_parameter_indices.argtypes = [C.c_void_p, C.c_float, C.c_float, zp]
        L.orc_parameter_weight.restype = C.c_int
        L.orc_parameter_weight.argtypes = [C.c_void_p, C.c_int, C.c_float, C.c_size_t, fp]
        L.orc_range_f32.argtypes = [C.c_float, C.c_float, C.c_int, fp, fp]
        L.orc_range_f64.argtypes = [C.c_double, C.c_double, C.c_int, dp, dp]
        L.orc_gstar_grid.argtypes = [C.c_int, C.c_int, dp]
        L.orc_conv_grid.argtypes = [dp]
        L.orc_base_limits.argtypes = [C.c_int, dp, dp]
        L.orc_emissivity.argtypes = [C.c_int, C.c_int, dp, C.c_double, C.c_double, dp, C.c_int, dp]
        L.orc_blend_frame.argtypes = [C.c_void_p, C.c_int, C.c_double, C.c_double, dp, dp, dp, dp, dp]
        L.orc_convolve.argtypes = [dp, C.c_int, dp, dp, C.c_int, dp, dp, C.c_int]
        L.orc_eval_batch.restype = C.c_int
        L.orc_eval_batch.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_long, dp, dp, C.c_int, dp, C.c_int, dp,
                                     C.c_int, C.c_int, C.POINTER(C.c_uint64), dp, dp, C.c_int, dp]
        L.orc_set_transcendental_policy.argtypes = [C.c_int, C.c_int, C.c_int]
        L.orc_set_transcendental_policy.restype = None
        L.orc_pow_probe.argtypes = [C.c_int, C.c_double, C.c_double]
        L.orc_pow_probe.restype = C.c_double
        _lib = L
    return _lib


TRANSCENDENTAL_POLICIES = {"default": 0, "plus": 1, "minus": 2, "random": 3, "libm_T": 4, "zig_style_pow": 5}


class transcendental_policy:
    """Context manager: evaluate cos / pow / log10 under another policy (sensitivity study only; see the
    policy comment in klp_oracle.cpp).  The default -- f64 libm rounded to T -- is what the CUDA path shares."""

    FUNCS = {"pow": 1, "cos": 2, "log10": 4}

    def __init__(self, mode: str, k_ulps: int = 0, funcs=("pow", "cos", "log10")):
        self.mode, self.k = TRANSCENDENTAL_POLICIES[mode], int(k_ulps)
        self.mask = sum(self.FUNCS[f] for f in funcs)

    def __enter__(self):
        lib().orc_set_transcendental_policy(self.mode, self.k, self.mask)
        return self

    def __exit__(self, *exc):
        lib().orc_set_transcendental_policy(0, 0, 7)
        return False


def pow_probe(x, y, precision="f32"):
    return float(lib().orc_pow_probe(PREC[precision], float(x), float(y)))


def _dp(a):
    return a.ctypes.data_as(C.POINTER(C.c_double)) if a is not None else None


def _fp(a):
    return a.ctypes.data_as(C.POINTER(C.c_float))


class OracleTable:
    def __init__(self, t: dict):
        self._keep = {k: np.ascontiguousarray(t[k], np.float32) for k in
                      ("spins", "mus", "radii", "gmin", "gmax", "upper", "lower")}
        k = self._keep
        self.n_a, self.n_mu, self.n_r, self.n_g = int(t["n_a"]), int(t["n_mu"]), int(t["n_r"]), int(t["n_g"])
        self.h = lib().orc_table_create(_fp(k["spins"]), self.n_a, _fp(k["mus"]), self.n_mu, self.n_r, self.n_g,
                                        _fp(k["radii"]), _fp(k["gmin"]), _fp(k["gmax"]), _fp(k["upper"]),
                                        _fp(k["lower"]))

    def __del__(self):
        try:
            if self.h:
                lib().orc_table_destroy(self.h)
                self.h = None
        except Exception:
            pass

    # -- index helpers (line-profile.zig:60-93,196-231)
    def table_index(self, ia, imu):
        return int(lib().orc_table_index(self.h, ia, imu))

    def find_tables(self, pi):
        a = (C.c_size_t * 2)(*pi)
        o = (C.c_size_t * 4)()
        lib().orc_find_tables(self.h, a, o)
        return list(o)

    def find_parameter_indices(self, a, mu):
        o = (C.c_size_t * 2)()
        lib().orc_find_parameter_indices(self.h, a, mu, o)
        return list(o)

    def parameter_weight(self, which, v, index):
        w = C.c_float()
        has = lib().orc_parameter_weight(self.h, which, v, index, C.byref(w))
        return w.value if has else None

    def blend_frame(self, a, incl_deg, precision="f64"):
        n_r, n_g = self.n_r, self.n_g
        out = [np.empty(n_r), np.empty(n_r), np.empty(n_r), np.empty(n_r * n_g), np.empty(n_r * n_g)]
        lib().orc_blend_frame(self.h, PREC[precision], a, incl_deg, *[_dp(o) for o in out])
        return dict(radii=out[0], gmin=out[1], gmax=out[2], upper=out[3].reshape(n_r, n_g),
                    lower=out[4].reshape(n_r, n_g))

    def eval_batch(self, model, params, energy, spectrum=None, precision="f32", threads=1, faithful_cost=False,
                   want_stats=False, want_fine=False, want_profile=False, ratchet=False, lims=None):
        m = MODELS[model] if isinstance(model, str) else int(model)
        params = np.ascontiguousarray(params, np.float64)
        if params.ndim == 1:
            params = params[None, :]
        B = params.shape[0]
        assert params.shape[1] == lib().orc_model_nparams(m), (params.shape, m)
        energy = np.ascontiguousarray(energy, np.float64)
        n = energy.size - 1
        per_set = 0
        if spectrum is not None:
            spectrum = np.ascontiguousarray(spectrum, np.float64)
            per_set = int(spectrum.ndim == 2)
            assert spectrum.shape[-1] == n
        out = np.empty((B, n), np.float64)
        stats = (C.c_uint64 * 7)() if want_stats else None
        fine = np.empty((B, N_FINE), np.float64) if want_fine else None
        prof = np.empty((B, N_CONV - 1), np.float64) if want_profile else None
        lim = np.array(lims if lims is not None else [0.0, 0.0], np.float64)
        rc = lib().orc_eval_batch(self.h, m, PREC[precision], B, _dp(params), _dp(energy), n, _dp(spectrum), per_set,
                                  _dp(out), int(threads), int(bool(faithful_cost)), stats, _dp(fine), _dp(prof),
                                  int(bool(ratchet)), _dp(lim))
        if rc != 0:
            raise RuntimeError(f"orc_eval_batch failed: {rc}")
        res = {"out": out}
        if want_stats:
            names = ("bin_visits", "active_bins", "gauss_bins", "edge_evals", "integrand_evals", "conv_pairs",
                     "conv_terms")
            res["stats"] = dict(zip(names, (int(v) for v in stats)))
        if want_fine:
            res["fine"] = fine
        if want_profile:
            res["profile"] = prof
        if ratchet:
            res["lims"] = lim
        return res


def range_iterator(lo, hi, n, precision="f64"):
    if precision == "f32":
        out = np.empty(n, np.float32)
        d = C.c_float()
        lib().orc_range_f32(lo, hi, n, _fp(out), C.byref(d))
    else:
        out = np.empty(n, np.float64)
        d = C.c_double()
        lib().orc_range_f64(lo, hi, n, _dp(out), C.byref(d))
    return out, d.value


def gstar_grid(n=30, precision="f32"):
    out = np.empty(n, np.float64)
    lib().orc_gstar_grid(PREC[precision], n, _dp(out))
    return out


def conv_grid():
    out = np.empty(N_CONV, np.float64)
    lib().orc_conv_grid(_dp(out))
    return out


def base_limits(precision="f32"):
    lo, hi = C.c_double(), C.c_double()
    lib().orc_base_limits(PREC[precision], C.byref(lo), C.byref(hi))
    return lo.value, hi.value


def emissivity(r, alpha, weights=None, rcut=50.0, precision="f32"):
    r = np.ascontiguousarray(r, np.float64)
    out = np.empty_like(r)
    if weights is None:
        lib().orc_emissivity(PREC[precision], 1, None, rcut, alpha, _dp(r), r.size, _dp(out))
    else:
        w = np.ascontiguousarray(weights, np.float64)
        lib().orc_emissivity(PREC[precision], w.size, _dp(w), rcut, alpha, _dp(r), r.size, _dp(out))
    return out


def convolve(g, a, x, b, faithful_cost=False):
    g, a, x, b = (np.ascontiguousarray(v, np.float64) for v in (g, a, x, b))
    out = np.empty(b.size, np.float64)
    lib().orc_convolve(_dp(out), b.size, _dp(g), _dp(a), a.size, _dp(x), _dp(b), int(bool(faithful_cost)))
    return out
