"""Caller-side batching adapter: an emcee / XSPEC-`chain` style ensemble sampler whose
walkers are evaluated in ONE batched model call per half-step instead of one C call per walker.

The reference is driven one parameter set per call (XSPEC fit steps, numerical derivatives, chain walkers,
xspec-wrapper.zig:425-530 callers); the batch entry `klp_eval_batch` only pays off if callers hand it many
sets at once -- this is the smallest useful such caller: a Goodman & Weare (2010) stretch-move sampler.
"""
from __future__ import annotations

from typing import Callable, Optional, Sequence

import numpy as np


def gaussian_loglike(model: np.ndarray, data: np.ndarray, sigma: np.ndarray) -> np.ndarray:
    return -0.5 * np.sum(((model - data) / sigma) ** 2, axis=-1)


class EnsembleSampler:
    """Affine-invariant ensemble sampler with batched likelihood evaluation.

    evaluate(params[B, P]) -> spectra[B, n]   (e.g. lambda p: table.eval_batch("kline5", p, energy))
    free: indices of the sampled parameters; the others stay at `start`'s values.
    bounds: (lo[P], hi[P]) hard limits (lmodel.dat ranges); proposals outside get -inf without a model call.
    """

    def __init__(self, evaluate: Callable[[np.ndarray], np.ndarray], data: np.ndarray, sigma: np.ndarray,
                 free: Sequence[int], bounds, amplitude: Optional[Callable[[np.ndarray], np.ndarray]] = None,
                 seed: int = 0, stretch: float = 2.0):
        self.evaluate, self.data, self.sigma = evaluate, np.asarray(data, float), np.asarray(sigma, float)
        self.free = np.asarray(free, int)
        self.lo, self.hi = (np.asarray(b, float) for b in bounds)
        self.amplitude = amplitude
        self.rng = np.random.default_rng(seed)
        self.a = stretch
        self.n_model_calls = 0
        self.n_sets = 0

    def logprob(self, p: np.ndarray) -> np.ndarray:
        ok = np.all((p >= self.lo) & (p <= self.hi), axis=1)
        out = np.full(p.shape[0], -np.inf)
        if ok.any():
            m = self.evaluate(p[ok])          # ONE batched call for all in-bounds walkers
            self.n_model_calls += 1
            self.n_sets += int(ok.sum())
            if self.amplitude is not None:
                m = m * self.amplitude(p[ok])[:, None]
            ll = gaussian_loglike(m, self.data, self.sigma)
            out[ok] = np.where(np.isfinite(ll), ll, -np.inf)
        return out

    def run(self, start: np.ndarray, n_steps: int):
        """start: [W, P] initial walkers (W even).  Returns (chain [n_steps, W, P], logprob [n_steps, W])."""
        p = np.array(start, float)
        W, nd = p.shape[0], self.free.size
        assert W % 2 == 0 and W >= 2 * nd
        lp = self.logprob(p)
        chain = np.empty((n_steps, W, p.shape[1]))
        lps = np.empty((n_steps, W))
        halves = (np.arange(W // 2), np.arange(W // 2, W))
        for t in range(n_steps):
            for h in (0, 1):
                me, other = halves[h], halves[1 - h]
                z = ((self.a - 1) * self.rng.random(me.size) + 1) ** 2 / self.a
                partner = p[self.rng.choice(other, me.size)]
                prop = p[me].copy()
                prop[:, self.free] = partner[:, self.free] + z[:, None] * (p[me][:, self.free] - partner[:, self.free])
                lq = self.logprob(prop)
                accept = np.log(self.rng.random(me.size)) < (nd - 1) * np.log(z) + lq - lp[me]
                p[me[accept]] = prop[accept]
                lp[me[accept]] = lq[accept]
            chain[t], lps[t] = p, lp
        return chain, lps
