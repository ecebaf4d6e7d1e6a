"""Independent numpy re-derivation of the kline / kline5 fine-grid flux, the rebin and the convolution.  TEST
INFRASTRUCTURE ONLY.  Every function takes the element type: np.float64 (the gate mode) or np.float32 -- the reference's own
type (xspec-wrapper.zig:61), where numpy's element-wise IEEE arithmetic reproduces the reference's roundings operation by
operation, so that the comparison with the C++ oracle in its faithful f32 mode is BIT FOR BIT (cos / pow / log10 are
evaluated in f64 and rounded, the oracle's stated policy).

Written from the reference source (file:line cited below), not from oracle/klp_oracle.cpp, and
structured differently on purpose: vectorised over the 1999 fine bins per radius with masks instead
of the scalar control flow of the reference.  tests/test_oracle.py compares it with the C++ oracle in
T = f64 mode; agreement of two independently written restatements is what stands in for the golden
spectra the reference does not have.
"""
from __future__ import annotations

import numpy as np

H = 5e-4                      # utils.zig:219
N_FINE, N_RAD = 2000, 1500    # xspec-wrapper.zig:15-16
GX = np.array([9.4910791234275852452618968404809e-01, 7.415311855993944398638647732811e-01,
               4.0584515137739716690660641207707e-01])               # Integrator.zig:5-10
GW = np.array([1.2948496616886969327061143267787e-01, 2.797053914892766679014677714229e-01,
               3.8183005050511894495036977548818e-01, 4.1795918367346938775510204081658e-01])


import math as _math

# Transcendentals: the same policy as the C++ oracle and the CUDA path (oracle/klp_oracle.cpp header): evaluated in f64 by
# the C library and rounded to the element type (Zig's own f32 implementations cannot be run here).


def _pow(F, x, y):
    return F(_math.pow(float(x), float(y)))


def _cos(F, x):
    return F(_math.cos(float(x)))


def _log10(F, x):
    return F(_math.log10(float(x)))


def cumulative_linspace(lo, hi, n, F=np.float64):
    """RangeIterator (utils.zig:121-156): current += delta, in the element type."""
    lo, hi = F(lo), F(hi)
    delta = (hi - lo) / F(n - 1)
    return np.cumsum(np.concatenate([[lo], np.full(n - 1, delta, F)]).astype(F), dtype=F)   # a sequential sum in F


def first_geq(arr, v):
    idx = np.nonzero(arr >= v)[0]
    return int(idx[0]) if idx.size else None


def blend(tab, a, mu, F=np.float64):
    """line-profile.zig:116-231 (Q5, Q14, Q15)."""
    spins, mus, n_mu = tab["spins"].astype(F), tab["mus"].astype(F), tab["n_mu"]
    a, mu = F(a), F(mu)
    i0 = first_geq(spins, a) or 0
    i1 = first_geq(mus, mu) or 0

    def frame(ia, im):
        f = ia * n_mu + im
        return {k: tab[k][f].astype(F) for k in ("radii", "gmin", "gmax", "upper", "lower")}

    def lerp(A, B, w):   # (y1 - y0) * w + y0, utils.zig:43-51
        return {k: (B[k] - A[k]) * w + A[k] for k in A}

    j0, j1 = max(i0 - 1, 0), max(i1 - 1, 0)
    if i0 != 0:
        w0 = (a - spins[i0 - 1]) / (spins[i0] - spins[i0 - 1])
        c0, c1 = lerp(frame(j0, i1), frame(i0, i1), w0), lerp(frame(j0, j1), frame(i0, j1), w0)
    else:
        c0, c1 = frame(j0, i1), frame(j0, j1)
    if i1 != 0:
        w1 = (mu - mus[i1 - 1]) / (mus[i1] - mus[i1 - 1])
        return lerp(c1, c0, w1)
    return c1


def emissivity_fn(alpha, knots=None, rcut=None, F=np.float64):
    """emissivity.zig:4-78 with init(weights, 1.0, rcut, alpha) (Q8); the power law is constructed with -alpha
    (xspec-wrapper.zig:352,401)."""
    alpha = F(alpha)
    if knots is None:
        return lambda r: _pow(F, r, -alpha)
    knots = np.asarray(knots).astype(F)
    n = len(knots)
    logr = cumulative_linspace(_log10(F, 1.0), _log10(F, F(rcut)), n + 1, F)[1:]
    logc = np.empty(n, F)
    logc[n - 1] = logr[n - 1] * (knots[n - 1] - alpha)
    for i in range(n - 2, -1, -1):
        logc[i] = logc[i + 1] + logr[i] * (knots[i] - knots[i + 1])
    radii = np.array([_pow(F, 10.0, v) for v in logr], F)
    coeffs = np.array([_pow(F, 10.0, v) for v in logc], F)

    def eps(r):
        i = first_geq(radii, r)
        return _pow(F, r, -alpha) if i is None else coeffs[i] * _pow(F, r, -knots[i])
    return eps


def fine_flux(tab, a, incl_deg, eline, rmin, rmax, eps, e_first, e_last, lim=(None, None), F=np.float64):
    """flux_cache after itf.integrate (xspec-wrapper.zig:147-160, transfer-functions.zig:274-384) in the element type F."""
    n_g = tab["n_g"]
    mu = _cos(F, F(incl_deg) * 0.017453292519943295769236907684886127)
    fr = blend(tab, a, mu, F)
    gs = cumulative_linspace(H, 1.0 - H, n_g, F)                        # utils.zig:219-232
    g = cumulative_linspace(e_first, e_last, N_FINE, F) * (F(1) / F(eline))    # refine_grid :117-132
    # radial grid :149-153 with the fresh setup() limits (:84)
    base = (F(1) / cumulative_linspace(F(1) / F(50), F(1) / F(1), N_RAD, F))[::-1]
    lo = max(base[0] if lim[0] is None else F(lim[0]), F(rmin))
    hi = min(F(rmax), base[-1] if lim[1] is None else F(lim[1]))
    r_grid = (F(1) / cumulative_linspace(F(1) / hi, F(1) / lo, N_RAD, F))[::-1]
    radii = fr["radii"]
    flux = np.zeros(N_FINE - 1, F)
    gmin = gmax = F(0)
    sqrtH = 0.022360679774997896964091736687313

    def integrand(x, cl, cu, gmin, gmax):
        gstar = (x - gmin) / (gmax - gmin)
        idx = np.searchsorted(gs, gstar, side="left")          # first knot >= gstar
        idx = np.where(idx >= n_g, n_g - 1, idx)                 # none -> last (:207-211)
        lowk = idx <= 1                                          # Q4
        i1 = np.where(lowk, 1, idx)
        fac = (gstar - gs[i1 - 1]) / (gs[i1] - gs[i1 - 1])
        lower = np.where(lowk, cl[0], fac * (cl[i1] - cl[i1 - 1]) + cl[i1 - 1])
        upper = np.where(lowk, cu[0], fac * (cu[i1] - cu[i1 - 1]) + cu[i1 - 1])
        with np.errstate(invalid="ignore", divide="ignore"):
            denom = np.sqrt(gstar * (1 - gstar)) * (gmax - gmin)
            return (upper + lower) * x * x * x / denom

    for i, r in enumerate(r_grid):
        if r < radii[-1] or r > radii[0]:
            continue                                             # Q13
        loc = np.nonzero(radii <= r)[0]
        loc = int(loc[0]) if loc.size else None
        if loc is None:
            cl, cu = fr["lower"][-1], fr["upper"][-1]
        elif loc <= 1:
            cl, cu = fr["lower"][0], fr["upper"][0]             # Q4: gmin/gmax stay stale
        else:
            f = (min(max(r, radii[-1]), radii[0]) - radii[loc - 1]) / (radii[loc] - radii[loc - 1])
            cl = (fr["lower"][loc] - fr["lower"][loc - 1]) * f + fr["lower"][loc - 1]
            cu = (fr["upper"][loc] - fr["upper"][loc - 1]) * f + fr["upper"][loc - 1]
            gmax = (fr["gmax"][loc] - fr["gmax"][loc - 1]) * f + fr["gmax"][loc - 1]
            gmin = (fr["gmin"][loc] - fr["gmin"][loc - 1]) * f + fr["gmin"][loc - 1]
        dr = (r_grid[1] - r_grid[0]) if i == 0 else (r - r_grid[i - 1])   # Q11
        weight = dr * r * eps(r) * np.pi
        glow = np.maximum(gmin, np.minimum(g[:-1], gmax))
        ghigh = np.maximum(gmin, np.minimum(g[1:], gmax))
        act = np.nonzero(glow != ghigh)[0]
        if act.size == 0:
            continue
        gl, gh = glow[act].copy(), ghigh[act].copy()
        dg = gmax - gmin
        with np.errstate(invalid="ignore", divide="ignore"):
            sl, sh = (gl - gmin) / dg, (gh - gmin) / dg
        val = np.zeros(act.size, F)
        gauss = np.ones(act.size, bool)

        def edge(lim, lim_star):                                  # Q17
            k = dg * lim_star + gmin
            with np.errstate(invalid="ignore"):
                return 2 * np.abs(np.sqrt(k) - np.sqrt(lim)) * integrand(k, cl, cu, gmin, gmax) * sqrtH * dg

        lo_edge = sl < H
        both = lo_edge & (sh > H)
        only = lo_edge & ~(sh > H)
        if both.any():
            val[both] += edge(gl[both], np.full(both.sum(), H, F))
            gl[both] = dg * F(H) + gmin
        if only.any():
            val[only] += edge(gl[only], sh[only])
            gauss[only] = False
        hi_edge = gauss & (sh > 1 - H)
        both = hi_edge & (sl < 1 - H)
        only = hi_edge & ~(sl < 1 - H)
        if both.any():
            val[both] += edge(gh[both], np.full(both.sum(), 1 - H, F))
            gh[both] = dg * F(1 - H) + gmin
        if only.any():
            val[only] += edge(gh[only], sl[only])
            gauss[only] = False
        if gauss.any():
            a_, b_ = gl[gauss], gh[gauss]
            scale = (b_ - a_) / 2
            s = np.zeros(a_.size, F)
            for k in range(3):
                xp = (float(GX[k]) + 1.0) * scale + a_            # (xi + 1) is a comptime constant, Integrator.zig:50-51
                xm = (-float(GX[k]) + 1.0) * scale + a_
                s = s + (integrand(xp, cl, cu, gmin, gmax) + integrand(xm, cl, cu, gmin, gmax)) * (float(GW[k]) * scale)
            s = s + integrand(scale + a_, cl, cu, gmin, gmax) * float(GW[3])     # Q1
            val[gauss] += s
        val = val * weight
        val[~np.isfinite(val)] = 0                                # Q12
        flux[act] += val
    return flux, g


def rebin(flux, g, energy, eline, F=np.float64):
    """xspec-wrapper.zig:162-182 (Q3, Q10): partial sums in F, the comparison and everything after in f64, with the
    reciprocal of eline rounded to F first."""
    inorm = float(F(1) / F(eline))
    out = np.zeros(energy.size - 1)
    j = 0
    for i in range(out.size):
        s = F(0)
        while j < flux.size and float(g[j]) < energy[i] * inorm:
            s = s + flux[j]
            j += 1
        out[i] = float(s) / (0.5 * (energy[i + 1] + energy[i]))
    total = 0.0
    for v in out:                      # total_flux += flux[i], sequentially
        total += float(v)
    return out / total


def conv_grid():
    """convolve_setup (xspec-wrapper.zig:98-105): RangeIterator(f64)(1e-3, 2.0, trunc((2 - 1e-3) / 1e-3) = 1999) (Q16)."""
    return cumulative_linspace(1e-3, 2.0, 1999)


def convolve(g, a, x, b):
    """convolution.zig:7-102 written as masked array arithmetic: for every source bin i one [n_out, n_window] overlap
    matrix instead of the reference's scalar scan with `continue` / `break` (g ascending: the bins a scan would skip or
    stop at are exactly those with bin_high < low, resp. every bin from the first bin_low > high on)."""
    pos = np.nonzero(a > 0)[0]
    ws = (pos[0] - 1 if pos[0] > 0 else pos[0]) if pos.size else 0                    # :17-21
    we = (pos[-1] + 1 if pos[-1] < a.size - 1 else pos[-1]) if pos.size else 0        # :22-29
    g_min, g_max = g[ws], g[we]
    bl, bh = g[ws:we], g[ws + 1:we + 1]                                               # window bins [ws, we)
    bw = bh - bl
    out = np.zeros(b.size)
    for i in range(b.size):
        avg = 2.0 / (x[i + 1] + x[i])
        low, high = (x[:-1] * avg)[:, None], (x[1:] * avg)[:, None]                  # [n_out, 1]
        inside = ~((high < g_min) | (low > g_max))                                    # :47
        BL, BH = bl[None, :], bh[None, :]
        hit = ~(BH < low)                                                             # not `continue`d ...
        stop = BL > high                                                              # ... and before the `break`
        hit &= np.cumsum(stop, axis=1) == 0
        c2 = (BL < low) & (BH > high)
        c3 = (BL < low) & (BH < high)
        c4 = (BL > low) & (BH > high)
        num = np.where(c2, high - low, np.where(c3, BH - low, np.where(c4, high - BL, 0.0)))
        overlap = np.where(c2 | c3 | c4, num / bw[None, :], 1.0)                      # Q9: exact coincidences count whole bins
        terms = np.where(hit, a[ws:we][None, :] * bw[None, :] * overlap, 0.0)
        weight = np.cumsum(terms, axis=1)[:, -1] if terms.shape[1] else np.zeros(x.size - 1)   # sequential, as `weight +=`
        out += np.where(inside[:, 0], weight * b[i], 0.0)
    return out
