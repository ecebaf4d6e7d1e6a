"""The batched-walker adapter: CPU test with a closed-form evaluator, GPU test recovering kline parameters."""
import numpy as np
import pytest

from lineprofiles_b200 import synthetic as syn
from lineprofiles_b200.walkers import EnsembleSampler


def test_sampler_recovers_gaussian_mean_with_batched_calls():
    x = np.linspace(-1, 1, 40)
    truth = np.array([0.3, 2.0])

    def evaluate(p):
        return p[:, 1:2] * np.exp(-0.5 * ((x[None, :] - p[:, 0:1]) / 0.2) ** 2)

    rng = np.random.default_rng(1)
    data = evaluate(truth[None])[0] + rng.normal(0, 0.05, x.size)
    s = EnsembleSampler(evaluate, data, np.full(x.size, 0.05), free=[0, 1], bounds=([-1, 0], [1, 5]), seed=2)
    start = truth + rng.normal(0, 0.1, (32, 2))
    chain, lp = s.run(start, 300)
    est = chain[100:].reshape(-1, 2).mean(axis=0)
    assert np.all(np.abs(est - truth) < 0.05)
    assert s.n_model_calls == 1 + 2 * 300            # one batched call per half-step, not one per walker
    assert s.n_sets > 15 * s.n_model_calls


@pytest.mark.gpu
def test_sampler_on_gpu_recovers_inclination(gpu_table):
    E = syn.energy_grid_linear(200)
    truth = np.array([0.9, 35.0, 6.4, 2.0, 300.0, 3.0])
    data = gpu_table.eval_batch("kline", truth[None], E)[0]
    sigma = np.full(200, 2e-4)
    lo = np.array([-0.998, 0, 0, 1, 1, 0.0])
    hi = np.array([0.998, 90, 10, 100, 1000, 10.0])
    s = EnsembleSampler(lambda p: gpu_table.eval_batch("kline", p, E), data, sigma, free=[1, 5], bounds=(lo, hi), seed=3)
    rng = np.random.default_rng(4)
    start = np.tile(truth, (64, 1))
    start[:, 1] += rng.normal(0, 3.0, 64)
    start[:, 5] += rng.normal(0, 0.3, 64)
    chain, lp = s.run(start, 60)
    est = np.median(chain[30:].reshape(-1, 6), axis=0)
    assert abs(est[1] - 35.0) < 1.5 and abs(est[5] - 3.0) < 0.3
    assert s.n_model_calls == 1 + 2 * 60
