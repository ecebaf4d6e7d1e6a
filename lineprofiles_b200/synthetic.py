"""Deterministic synthetic inputs for the kline/kconv hot path.

The real `kerr-transfer-functions.fits` (reference io.zig:53-87) cannot be fetched here, so the
benchmark and the parity tests run on a "Kerr-shaped" table built from closed forms: same schema
(spins x cos-inclinations frames, each with descending radii, gmin/gmax per radius and the two
Cunningham branches sampled on n_g g* knots), same a-major / mu-minor frame order
(line-profile.zig:60-70), no RNG.

The spin grid is chosen so that the reference's table-dependent known answers (main.zig:19-30:
n_mu = 20, {0.92, 0.4} -> indices {24, 8}) also hold on the default 30 x 20 table.
"""
from __future__ import annotations

import numpy as np

H = 5e-4  # utils.zig:219

MODELS = {"kline": 0, "kline5": 1, "kconv": 2, "kconv5": 3, "kconv10": 4}
NPARAMS = {"kline": 6, "kline5": 12, "kconv": 5, "kconv5": 11, "kconv10": 16}  # lmodel.dat


def r_isco(a):
    """Bardeen, Press & Teukolsky (1972) prograde ISCO radius."""
    a = np.asarray(a, dtype=np.float64)
    z1 = 1 + np.cbrt(1 - a * a) * (np.cbrt(1 + a) + np.cbrt(1 - a))
    z2 = np.sqrt(3 * a * a + z1 * z1)
    return 3 + z2 - np.sqrt((3 - z1) * (3 + z1 + 2 * z2))


def spin_grid(n_a: int) -> np.ndarray:
    k = np.arange(n_a, dtype=np.float64)
    return (0.998 * (1 - (1 - k / (n_a - 1)) ** 1.5)).astype(np.float32)


def mu_grid(n_mu: int) -> np.ndarray:
    return np.linspace(np.cos(np.deg2rad(89.0)), np.cos(np.deg2rad(1.0)), n_mu).astype(np.float32)


def make_table(n_a: int = 30, n_mu: int = 20, n_r: int = 150, n_g: int = 30) -> dict:
    """Return the table as a dict of float32 arrays in the library's frame layout.

    radii/gmin/gmax: [F, n_r]; upper/lower: [F, n_r, n_g]; F = n_a*n_mu, frame = ia*n_mu + imu.
    """
    spins = spin_grid(n_a)
    mus = mu_grid(n_mu)
    F = n_a * n_mu
    radii = np.empty((F, n_r), np.float32)
    gmin = np.empty((F, n_r), np.float32)
    gmax = np.empty((F, n_r), np.float32)
    upper = np.empty((F, n_r, n_g), np.float32)
    lower = np.empty((F, n_r, n_g), np.float32)
    gs = H + np.arange(n_g, dtype=np.float64) * ((1 - 2 * H) / (n_g - 1))
    up_g = 1 + 0.3 * np.sin(np.pi * gs)
    lo_g = 0.8 * (1 + 0.3 * np.cos(np.pi * gs))
    j = np.arange(n_r, dtype=np.float64) / (n_r - 1)
    for ia in range(n_a):
        a = float(spins[ia])
        rin = 1.001 * float(r_isco(a))
        r = 1000.0 * (rin / 1000.0) ** j  # descending, geometric
        x = r ** 1.5
        omega = 1.0 / (x + a)
        ut = (1 + a / x) / np.sqrt(1 - 3 / r + 2 * a / x)
        b = r + 1
        sq = np.sqrt(r)
        for im in range(n_mu):
            mu = float(mus[im])
            sini = np.sqrt(max(0.0, 1 - mu * mu))
            gx = 1.0 / (ut * (1 - omega * b * sini))
            gn = 1.0 / (ut * (1 + omega * b * sini))
            gx = np.minimum(gx, 1.8)
            gn = np.minimum(gn, 0.98 * gx)
            f = ia * n_mu + im
            radii[f] = r
            gmin[f] = gn
            gmax[f] = gx
            upper[f] = (up_g[None, :] * ((1 + mu) * sq)[:, None])
            lower[f] = (lo_g[None, :] * ((1 + mu) * sq)[:, None])
    return dict(spins=spins, mus=mus, n_a=n_a, n_mu=n_mu, n_r=n_r, n_g=n_g,
                radii=radii, gmin=gmin, gmax=gmax, upper=upper, lower=lower)


def table_nbytes(t: dict) -> int:
    return sum(t[k].nbytes for k in ("radii", "gmin", "gmax", "upper", "lower"))


def range_iterator(lo: float, hi: float, n: int, dtype=np.float64) -> np.ndarray:
    """Cumulative-add linspace, reference utils.zig:121-156 (Q6)."""
    dt = np.dtype(dtype).type
    delta = dt((dt(hi) - dt(lo)) / dt(n - 1))
    out = np.empty(n, dtype)
    cur = dt(lo)
    for i in range(n):
        out[i] = cur
        cur = dt(cur + delta)
    return out


def energy_grid_linear(n_bins: int = 1000, lo: float = 0.1, hi: float = 10.0) -> np.ndarray:
    """Config-1/2 grid: RangeIterator(f64)(lo, hi, n_bins+1), as exampleDomain (xspec-wrapper.zig:548-551)."""
    return range_iterator(lo, hi, n_bins + 1, np.float64)


def energy_grid_log(n_bins: int = 4096, lo: float = 0.1, hi: float = 10.0) -> np.ndarray:
    """Config-3/4 grid: E_k = lo * (hi/lo)^(k/n_bins), computed directly in f64."""
    k = np.arange(n_bins + 1, dtype=np.float64)
    return lo * (hi / lo) ** (k / n_bins)


def powerlaw_continuum(energy: np.ndarray, gamma: float = 2.0) -> np.ndarray:
    """photons/bin of E^-gamma: for gamma = 2, b_i = 1/E_i - 1/E_{i+1}."""
    e = np.asarray(energy, np.float64)
    if gamma == 1.0:
        return np.log(e[1:] / e[:-1])
    p = 1.0 - gamma
    return (e[1:] ** p - e[:-1] ** p) / p


_SEEDS = {"kline": 20259, "kline5": 20260, "kconv": 20261, "kconv5": 20263, "kconv10": 20262}


def random_params(model: str, B: int, seed: int | None = None) -> np.ndarray:
    """Random parameter sets [B, P] (f64) with the marginals of SURVEY.md §8d (after fuzz.zig:100-111)."""
    rng = np.random.default_rng(_SEEDS[model] if seed is None else seed)
    a = rng.uniform(0.0, 0.998, B)
    incl = rng.uniform(5.0, 85.0, B)
    eline = rng.uniform(6.0, 7.0, B)
    rmin = rng.uniform(1.0, 10.0, B)
    rmax = rng.uniform(20.0, 400.0, B)
    rcut = rng.uniform(5.0, 100.0, B)
    alpha = rng.uniform(0.0, 5.0, B)
    nemis = {"kline": 0, "kline5": 5, "kconv": 0, "kconv5": 5, "kconv10": 10}[model]
    e = rng.uniform(0.0, 5.0, (B, max(nemis, 1)))
    cols = [a, incl]
    if model in ("kline", "kline5"):
        cols.append(eline)
    cols += [rmin, rmax]
    if nemis:
        cols += [rcut, alpha] + [e[:, i] for i in range(nemis)]
    else:
        cols.append(alpha)
    p = np.stack(cols, axis=1).astype(np.float64)
    assert p.shape[1] == NPARAMS[model]
    return np.ascontiguousarray(p)


CONFIG1_PARAMS = np.array([[0.998, 30.0, 6.4, float(r_isco(0.998)), 400.0, 3.0]], np.float64)
