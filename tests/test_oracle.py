"""CPU tests of the oracle (oracle/klp_oracle.cpp) against every known answer the reference holds for
this path, plus property tests.  The reference asserts no spectrum value anywhere
(xspec-wrapper.zig:553-585), so flux parity is pinned by oracle-generated golden vectors
(tests/golden/make_golden.py) and by an independent numpy re-derivation (oracle/numpy_check.py)."""
import os

import numpy as np
import pytest

from lineprofiles_b200 import synthetic as syn
from oracle import oracle_py as orc

GOLD = os.path.join(os.path.dirname(__file__), "golden", "oracle_golden.npz")


# ---- known answers held by the reference ---------------------------------------------------------
def test_index_lookups_main_zig(oracle_table):
    """main.zig:19-30 "index-lookups" (the synthetic table is 30 x 20 like the shipped one is n_mu = 20)."""
    ot = oracle_table
    assert ot.table_index(27, 12) == 552          # main.zig:19
    assert ot.table_index(0, 12) == 12            # main.zig:20
    assert ot.find_tables([27, 12]) == [532, 552, 531, 551]   # main.zig:22-23
    assert ot.find_tables([0, 1]) == [1, 1, 0, 0]             # main.zig:25-26
    assert ot.find_parameter_indices(0.92, 0.4) == [24, 8]    # main.zig:28-30
    w1 = ot.parameter_weight(0, 0.92, 24)
    w2 = ot.parameter_weight(1, 0.4, 8)
    assert 0 < w1 <= 1 and 0 < w2 <= 1
    assert ot.parameter_weight(0, 0.0, 0) is None             # line-profile.zig:203


def test_range_iterator_utils_zig():
    """utils.zig:158-171 "range-iterator"."""
    v, delta = orc.range_iterator(8, 32, 4, "f32")
    assert delta == 8
    assert len(v) == 4 and v[0] == 8 and v[-1] == 32


def test_python_range_iterator_matches_oracle():
    for lo, hi, n in ((0.1, 10.0, 1001), (1e-3, 2.0, 1999)):
        v, _ = orc.range_iterator(lo, hi, n, "f64")
        assert np.array_equal(v, syn.range_iterator(lo, hi, n, np.float64))
    v, _ = orc.range_iterator(0.1, 10.0, 2000, "f32")
    assert np.array_equal(v, syn.range_iterator(0.1, 10.0, 2000, np.float32))


def test_fixed_grids():
    g = orc.conv_grid()          # xspec-wrapper.zig:98-105 (Q16): 1999 points 1e-3 ... 2.0
    assert g.size == 1999 and g[0] == 1e-3 and abs(g[-1] - 2.0) < 1e-12
    assert np.all(np.diff(g) > 0)
    gs = orc.gstar_grid(30, "f32")   # utils.zig:219-232
    assert gs[0] == np.float32(5e-4) and abs(gs[-1] - (1 - 5e-4)) < 1e-5
    lo, hi = orc.base_limits("f32")  # inverse_grid(f32, 1.0, 50.0, 1500), xspec-wrapper.zig:84
    assert abs(lo - 1.0) < 1e-4 and abs(hi - 50.0) < 1e-4


# ---- emissivity (emissivity.zig) -------------------------------------------------------------------
def test_emissivity_power_law():
    r = np.array([1.5, 3.0, 10.0, 45.0])
    e = orc.emissivity(r, 3.0, precision="f64")
    np.testing.assert_allclose(e, r ** -3.0, rtol=1e-14)


@pytest.mark.parametrize("n", [5, 10])
def test_emissivity_knots_continuous(n):
    rng = np.random.default_rng(3)
    w = rng.uniform(0, 5, n)
    rcut, alpha = 37.0, 2.5
    knots = 10 ** (np.arange(1, n + 1) * np.log10(rcut) / n)   # Q8: log-spaced between 1.0 and rcut
    for k in knots:
        below = orc.emissivity(np.array([k * (1 - 1e-9)]), alpha, w, rcut, "f64")[0]
        above = orc.emissivity(np.array([k * (1 + 1e-9)]), alpha, w, rcut, "f64")[0]
        assert abs(below - above) <= 1e-6 * abs(below)
    far = np.array([rcut * 1.5, 400.0])
    np.testing.assert_allclose(orc.emissivity(far, alpha, w, rcut, "f64"), far ** -alpha, rtol=1e-12)
    # innermost segment has slope -e1
    r = np.array([1.01, 1.02])
    e = orc.emissivity(r, alpha, w, rcut, "f64")
    assert abs(np.log(e[1] / e[0]) / np.log(r[1] / r[0]) + w[0]) < 1e-9


def test_emissivity_reference_unit_case():
    """emissivity.zig:80-88 only checks that LinInterpEmissivity(f32, 4).init([5,1,3,2], 1, 50, 3) constructs."""
    e = orc.emissivity(np.array([2.0, 10.0, 60.0]), 3.0, np.array([5.0, 1.0, 3.0, 2.0, 2.0]), 50.0, "f32")
    assert np.all(np.isfinite(e)) and np.all(e > 0)


# ---- frame blend (line-profile.zig:116-231) --------------------------------------------------------
def test_blend_at_grid_node_returns_that_frame(table_arrays, oracle_table):
    t = table_arrays
    ia, im = 7, 11
    incl = np.rad2deg(np.arccos(np.float64(t["mus"][im])))
    fr = oracle_table.blend_frame(float(t["spins"][ia]), incl, "f64")
    f = ia * t["n_mu"] + im
    np.testing.assert_allclose(fr["radii"], t["radii"][f], rtol=2e-6)
    np.testing.assert_allclose(fr["upper"], t["upper"][f], rtol=2e-5)


def test_blend_out_of_range_snaps_to_first_frame(table_arrays, oracle_table):
    """Q5: a parameter above the table maximum finds no >= entry and falls back to index 0."""
    hi = oracle_table.blend_frame(1.5, 40.0, "f32")
    lo = oracle_table.blend_frame(-0.5, 40.0, "f32")
    for k in hi:
        assert np.array_equal(hi[k], lo[k])


# ---- whole-model properties -------------------------------------------------------------------------
@pytest.mark.parametrize("model", ["kline", "kline5"])
@pytest.mark.parametrize("prec", ["f32", "f64"])
def test_line_models_normalised_and_cost_modes_identical(oracle_table, model, prec):
    E = syn.energy_grid_linear(300)
    p = syn.random_params(model, 3, seed=11)
    a = oracle_table.eval_batch(model, p, E, precision=prec, faithful_cost=False, want_fine=True)
    b = oracle_table.eval_batch(model, p, E, precision=prec, faithful_cost=True, want_fine=True)
    assert np.array_equal(a["out"], b["out"]) and np.array_equal(a["fine"], b["fine"])
    np.testing.assert_allclose(a["out"].sum(axis=1), 1.0, rtol=1e-12)   # Q10
    assert np.all(a["out"] >= 0)


def test_config1_statistics(oracle_table):
    """BASELINE config 1; the evaluation counts are the A_F inputs of SURVEY.md §8d."""
    E = syn.energy_grid_linear(1000)
    r = oracle_table.eval_batch("kline", syn.CONFIG1_PARAMS, E, precision="f32", faithful_cost=True, want_stats=True)
    st = r["stats"]
    assert st["bin_visits"] == 1498 * 1999          # 2 of the 1500 radii lie below the frame's inner edge (Q13)
    assert 3.5e6 < st["integrand_evals"] < 4.5e6
    assert st["integrand_evals"] == 7 * st["gauss_bins"] + st["edge_evals"]
    out = r["out"][0]
    assert abs(out.sum() - 1) < 1e-12
    # the line peaks just below the rest energy 6.4 keV (Q3 shifts it up by one 9.9 eV bin)
    assert 6.0 < 0.5 * (E[np.argmax(out)] + E[np.argmax(out) + 1]) < 6.8


def test_precisions_agree_loosely(oracle_table):
    E = syn.energy_grid_linear(500)
    p = syn.random_params("kline5", 4, seed=5)
    a = oracle_table.eval_batch("kline5", p, E, precision="f32")["out"]
    b = oracle_table.eval_batch("kline5", p, E, precision="f64")["out"]
    m = b > 1e-6
    assert np.median(np.abs(a[m] - b[m]) / b[m]) < 5e-3


def test_conv_cost_modes_identical_and_delta_input(oracle_table):
    E = syn.energy_grid_log(200)
    p = syn.random_params("kconv5", 2, seed=9)
    spec = np.zeros(200)
    spec[120] = 1.0
    a = oracle_table.eval_batch("kconv5", p, E, spectrum=spec, precision="f32", faithful_cost=False, want_profile=True)
    b = oracle_table.eval_batch("kconv5", p, E, spectrum=spec, precision="f32", faithful_cost=True, want_profile=True)
    assert np.array_equal(a["out"], b["out"]) and np.array_equal(a["profile"], b["profile"])
    np.testing.assert_allclose(a["profile"].sum(axis=1), 1.0, rtol=1e-12)
    # a delta input reproduces the line profile around the source bin; Q9: scaled by the ~1e-3 conv bin width
    s = a["out"].sum(axis=1)
    assert np.all(s > 0.5e-3) and np.all(s < 1.5e-3)


def test_convolve_standalone_matches_direct_sum():
    g = orc.conv_grid()
    rng = np.random.default_rng(1)
    a = np.zeros(1998)
    a[400:1300] = rng.uniform(0, 1, 900)
    x = syn.energy_grid_log(64, 0.5, 5.0)
    b = rng.uniform(0, 1, 64)
    got = orc.convolve(g, a, x, b, faithful_cost=True)
    assert np.array_equal(got, orc.convolve(g, a, x, b, faithful_cost=False))
    # independent evaluation of convolution.zig:7-102 with numpy interval arithmetic (generic position)
    want = np.zeros(64)
    nz = np.nonzero(a > 0)[0]
    ws, we = max(nz[0] - 1, 0), min(nz[-1] + 1, 1997)
    for i in range(64):
        avg = 2 / (x[i + 1] + x[i])
        for j in range(64):
            lo, hi = x[j] * avg, x[j + 1] * avg
            if hi < g[ws] or lo > g[we]:
                continue
            ov = np.clip(np.minimum(g[ws + 1:we + 1], hi) - np.maximum(g[ws:we], lo), 0, None)
            want[j] += np.sum(a[ws:we] * ov) * b[i]
    np.testing.assert_allclose(got, want, rtol=1e-10, atol=1e-18)


def test_ratchet_q2(oracle_table):
    """Q2: the radial limits only shrink across calls of one process."""
    E = syn.energy_grid_linear(100)
    p = np.array([[0.5, 40, 6.4, 5.0, 20.0, 3.0], [0.5, 40, 6.4, 1.0, 400.0, 3.0]])
    lo, hi = orc.base_limits("f32")
    r = oracle_table.eval_batch("kline", p, E, precision="f32", ratchet=True, lims=[lo, hi])
    fresh = oracle_table.eval_batch("kline", p, E, precision="f32")
    assert abs(r["lims"][0] - 5.0) < 1e-3 and abs(r["lims"][1] - 20.0) < 1e-3
    assert np.array_equal(r["out"][0], fresh["out"][0])
    assert not np.array_equal(r["out"][1], fresh["out"][1])   # second call is confined to [5, 20]


def test_reference_smoke_vectors_do_not_crash(oracle_table):
    """xspec-wrapper.zig:553-585 "smoke test all": same parameter vectors (incl. the stale kline order)."""
    dom, _ = orc.range_iterator(0.1, 10.0, 100, "f64")
    for v in ([0.998, 60, 6.4, 3.0, 1.0, 400.0], [0.998, 0.1, 6.4, 3.0, 1.0, 400.0], [0.998, 89.0, 6.4, 3.0, 1.0, 400.0],
              [0.998, 60.0, 6.4, 0.0, 1.0, 400.0], [0.998, 60.0, 6.4, 0.0, 0.0, 400.0]):
        oracle_table.eval_batch("kline", np.array([v]), dom, precision="f32")
    flux = np.ones(99)
    oracle_table.eval_batch("kconv5", np.array([[0.998, 60.0, 1.0, 400.0, 100.0, 3.0, 3.0, 3.0, 3.0, 3.0, 3.0]]), dom,
                            spectrum=flux, precision="f32")
    oracle_table.eval_batch("kconv10", np.array([[0.998, 60.0, 1.0, 400.0, 50.0, 3.0] + [3.0] * 10]), dom,
                            spectrum=flux, precision="f32")


# ---- golden vectors ---------------------------------------------------------------------------------
def test_golden_vectors(oracle_table):
    """Oracle output pinned at commit time (generated by tests/golden/make_golden.py on this table)."""
    g = np.load(GOLD)
    for model in ("kline", "kline5", "kconv", "kconv5", "kconv10"):
        for prec in ("f32", "f64"):
            conv = model.startswith("kconv")
            E = g["energy_conv"] if conv else g["energy_line"]
            spec = g["spectrum"] if conv else None
            out = oracle_table.eval_batch(model, g[f"params_{model}"], E, spectrum=spec, precision=prec)["out"]
            want = g[f"out_{model}_{prec}"]
            np.testing.assert_allclose(out, want, rtol=1e-11, atol=1e-300, equal_nan=True)


# ---- independent re-derivation ----------------------------------------------------------------------
def test_numpy_rederivation_agrees_with_oracle(table_arrays, oracle_table):
    """oracle/numpy_check.py (vectorised, written separately from the reference source) vs the C++ oracle, f64."""
    from oracle import numpy_check as nc
    E = syn.energy_grid_linear(1000)
    p = syn.CONFIG1_PARAMS[0]
    ref = oracle_table.eval_batch("kline", p[None], E, precision="f64", want_fine=True)
    fine, g = nc.fine_flux(table_arrays, p[0], p[1], p[2], p[3], p[4], nc.emissivity_fn(p[5]), E[0], E[-1])
    assert np.array_equal(fine, ref["fine"][0])          # same IEEE operations in the same order: bit-identical
    assert np.array_equal(nc.rebin(fine, g, E, p[2]), ref["out"][0])
    # a kline5 set (knot emissivity) at a high inclination
    q = syn.random_params("kline5", 1, seed=77)[0]
    q[1] = 72.0
    ref = oracle_table.eval_batch("kline5", q[None], E, precision="f64", want_fine=True)
    fine, g = nc.fine_flux(table_arrays, q[0], q[1], q[2], q[3], q[4], nc.emissivity_fn(q[6], q[7:], q[5]), E[0], E[-1])
    assert np.array_equal(fine, ref["fine"][0])


def test_numpy_convolution_agrees_with_oracle(oracle_table):
    """oracle/numpy_check.convolve (masked array arithmetic, written from convolution.zig) vs the C++ oracle's kconv:
    the profile on the fixed 1999-point grid (Q16) convolved with a continuum on linear, log and ragged output grids."""
    from oracle import numpy_check as nc
    rng = np.random.default_rng(3)
    grids = [syn.energy_grid_linear(60), syn.energy_grid_log(96), np.sort(rng.uniform(0.2, 9.0, 41))]
    p = syn.random_params("kconv", 2, seed=21)
    p[1, 1] = 75.0
    for E in grids:
        spec = rng.uniform(0.1, 1.0, E.size - 1)
        ref = oracle_table.eval_batch("kconv", p, E, spectrum=spec, precision="f64", want_profile=True)
        for s in range(2):
            got = nc.convolve(nc.conv_grid(), ref["profile"][s], E, spec)
            assert np.array_equal(got, ref["out"][s])


def test_numpy_kconv_chain_agrees_with_oracle(table_arrays, oracle_table):
    """The whole kconv path in numpy (profile on the fixed grid with eline = 1, xspec-wrapper.zig:185-205, then convolve)
    vs the C++ oracle, f64, one set."""
    from oracle import numpy_check as nc
    E = syn.energy_grid_log(64)
    spec = syn.powerlaw_continuum(E)
    p = syn.random_params("kconv", 1, seed=5)[0]   # {a, incl, rmin, rmax, alpha}
    ref = oracle_table.eval_batch("kconv", p[None], E, spectrum=spec, precision="f64", want_profile=True)
    cg = nc.conv_grid()
    fine, g = nc.fine_flux(table_arrays, p[0], p[1], 1.0, p[2], p[3], nc.emissivity_fn(p[4]), cg[0], cg[-1])
    prof = nc.rebin(fine, g, cg, 1.0)
    assert np.array_equal(prof, ref["profile"][0])
    assert np.array_equal(nc.convolve(cg, prof, E, spec), ref["out"][0])


def test_numpy_kconv10_chain_agrees_with_oracle(table_arrays, oracle_table):
    """kconv10 (ten emissivity knots, xspec-wrapper.zig:294-310 parameter layout) through the numpy derivation, f64."""
    from oracle import numpy_check as nc
    E = syn.energy_grid_log(48)
    spec = syn.powerlaw_continuum(E)
    p = syn.random_params("kconv10", 1, seed=9)[0]   # {a, incl, rmin, rmax, rcut, alpha, e1..e10}
    ref = oracle_table.eval_batch("kconv10", p[None], E, spectrum=spec, precision="f64", want_profile=True)
    cg = nc.conv_grid()
    fine, g = nc.fine_flux(table_arrays, p[0], p[1], 1.0, p[2], p[3], nc.emissivity_fn(p[5], p[6:], p[4]), cg[0], cg[-1])
    prof = nc.rebin(fine, g, cg, 1.0)
    assert np.array_equal(prof, ref["profile"][0])
    assert np.array_equal(nc.convolve(cg, prof, E, spec), ref["out"][0])


def test_numpy_f32_restatement_is_bit_identical_to_oracle(table_arrays, oracle_table):
    """The reference's own element type: numpy float32 arithmetic (element-wise IEEE, no contraction) reproduces the
    reference's roundings operation by operation, so the independently written, vectorised restatement and the C++ oracle
    in faithful f32 mode must agree BIT FOR BIT -- fine-grid flux and output spectrum, kline / kline5 / kconv."""
    from oracle import numpy_check as nc
    F = np.float32
    E = syn.energy_grid_linear(1000)
    p = syn.CONFIG1_PARAMS[0]
    ref = oracle_table.eval_batch("kline", p[None], E, precision="f32", want_fine=True)
    fine, g = nc.fine_flux(table_arrays, p[0], p[1], p[2], p[3], p[4], nc.emissivity_fn(p[5], F=F), E[0], E[-1], F=F)
    assert np.array_equal(fine.astype(np.float64), ref["fine"][0])
    assert np.array_equal(nc.rebin(fine, g, E, p[2], F=F), ref["out"][0])
    P = syn.random_params("kline5", 3, seed=77)
    P[0, 1], P[1, 1] = 72.0, 84.9
    ref = oracle_table.eval_batch("kline5", P, E, precision="f32", want_fine=True)
    for s, q in enumerate(P):
        fine, g = nc.fine_flux(table_arrays, q[0], q[1], q[2], q[3], q[4], nc.emissivity_fn(q[6], q[7:], q[5], F=F), E[0], E[-1], F=F)
        assert np.array_equal(fine.astype(np.float64), ref["fine"][s])
        assert np.array_equal(nc.rebin(fine, g, E, q[2], F=F), ref["out"][s])
    # kconv: f32 profile on the fixed f64 grid (refine_grid casts its end points to f32, xspec-wrapper.zig:125), f64 convolution
    El = syn.energy_grid_log(64)
    spec = syn.powerlaw_continuum(El)
    k = syn.random_params("kconv", 1, seed=5)[0]
    ref = oracle_table.eval_batch("kconv", k[None], El, spectrum=spec, precision="f32", want_profile=True)
    cg = nc.conv_grid()
    fine, g = nc.fine_flux(table_arrays, k[0], k[1], 1.0, k[2], k[3], nc.emissivity_fn(k[4], F=F), cg[0], cg[-1], F=F)
    prof = nc.rebin(fine, g, cg, 1.0, F=F)
    assert np.array_equal(prof, ref["profile"][0])
    assert np.array_equal(nc.convolve(cg, prof, El, spec), ref["out"][0])


def test_numpy_golden_equals_oracle_golden():
    """tests/golden/numpy_golden.npz (f32 spectra of all five models from the numpy restatement,
    tests/golden/make_numpy_golden.py) and oracle_golden.npz (from the C++ oracle) are bit-identical: the vectors the CUDA
    path is checked against on the GPU box have been produced twice, by two separately written implementations."""
    a = np.load(os.path.join(os.path.dirname(GOLD), "numpy_golden.npz"))
    b = np.load(GOLD)
    assert sorted(a.files) == sorted(f"out_{m}_f32" for m in ("kline", "kline5", "kconv", "kconv5", "kconv10"))
    for k in a.files:
        assert np.array_equal(a[k], b[k], equal_nan=True), k


def _same_bits(a, b):
    return bool(np.all((a == b) | (np.isnan(a) & np.isnan(b))))


def test_numpy_f32_restatement_on_edge_and_irregular_sets(table_arrays, oracle_table):
    """The parameter sets of the GPU edge tests (out-of-table spin / inclination, rmin > rmax, zero knots, rcut = 1, eline
    negative / NaN / huge / tiny / inf) through the numpy f32 restatement: bit-identical to the oracle, NaNs included."""
    import warnings
    from oracle import numpy_check as nc
    F = np.float32
    E = syn.energy_grid_linear(300)
    p = syn.random_params("kline5", 14, seed=31)
    p[0, 0] = 1.7
    p[1, 0] = -0.9
    p[2, 1] = 0.0
    p[3, 1] = 90.0
    p[4, 3:5] = (30.0, 10.0)
    p[5, 7:] = 0.0
    p[6, 5] = 1.0
    p[7, 3:5] = (1.0, 1000.0)
    p[8, 6] = 10.0
    p[9:14, 2] = (-6.4, np.nan, 1e6, 1e-3, 0.5)
    ref = oracle_table.eval_batch("kline5", p, E, precision="f32", want_fine=True)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        for s, q in enumerate(p):
            fine, g = nc.fine_flux(table_arrays, q[0], q[1], q[2], q[3], q[4], nc.emissivity_fn(q[6], q[7:], q[5], F=F), E[0], E[-1], F=F)
            assert _same_bits(fine.astype(np.float64), ref["fine"][s]), s
            assert _same_bits(nc.rebin(fine, g, E, q[2], F=F), ref["out"][s]), s


def test_numpy_f32_restatement_follows_the_ratchet(table_arrays, oracle_table):
    """Q2 through the numpy restatement: the second call of a process starts from the radial limits the first one left
    (set_radial_grid reads the CURRENT r_grid, xspec-wrapper.zig:149-153) -- bit-identical to the oracle's ratchet mode."""
    from oracle import numpy_check as nc
    F = np.float32
    E = syn.energy_grid_linear(100)
    p = np.array([[0.5, 40, 6.4, 5.0, 20.0, 3.0], [0.5, 40, 6.4, 1.0, 400.0, 3.0]])
    lo, hi = orc.base_limits("f32")
    first = oracle_table.eval_batch("kline", p[:1], E, precision="f32", ratchet=True, lims=[lo, hi])
    both = oracle_table.eval_batch("kline", p, E, precision="f32", ratchet=True, lims=[lo, hi])
    q = p[1]
    fine, g = nc.fine_flux(table_arrays, q[0], q[1], q[2], q[3], q[4], nc.emissivity_fn(q[5], F=F), E[0], E[-1],
                           lim=tuple(first["lims"]), F=F)
    assert np.array_equal(nc.rebin(fine, g, E, q[2], F=F), both["out"][1])


@pytest.mark.parametrize("n_bins", [1, 7, 2500])
def test_numpy_f32_restatement_on_ragged_grids(table_arrays, oracle_table, n_bins):
    """One-bin, short and finer-than-the-fine-grid ragged output grids (the rebin walks its index past empty output bins,
    xspec-wrapper.zig:162-182): bit-identical to the oracle."""
    from oracle import numpy_check as nc
    F = np.float32
    E = np.sort(np.random.default_rng(n_bins).uniform(0.3, 9.5, n_bins + 1))
    q = syn.random_params("kline5", 1, seed=9)[0]
    ref = oracle_table.eval_batch("kline5", q[None], E, precision="f32")["out"][0]
    fine, g = nc.fine_flux(table_arrays, q[0], q[1], q[2], q[3], q[4], nc.emissivity_fn(q[6], q[7:], q[5], F=F), E[0], E[-1], F=F)
    assert _same_bits(nc.rebin(fine, g, E, q[2], F=F), ref)


# ---- round 2: transcendental policies (sensitivity study; see klp_oracle.cpp, tools/transcendental_sensitivity.py) ----

def test_transcendental_policies_are_what_they_say():
    x, y = 7.3, -2.75
    base = orc.pow_probe(x, y, "f32")
    assert base == float(np.float32(np.float64(np.float32(x)) ** np.float64(np.float32(y))))
    with orc.transcendental_policy("plus", 3):
        assert orc.pow_probe(x, y, "f32") == float(np.nextafter(np.nextafter(np.nextafter(np.float32(base), np.float32(9)), np.float32(9)), np.float32(9)))
    with orc.transcendental_policy("minus", 1):
        assert orc.pow_probe(x, y, "f32") == float(np.nextafter(np.float32(base), np.float32(0)))
    with orc.transcendental_policy("random", 2):
        v = np.float32(orc.pow_probe(x, y, "f32"))
        assert abs(int(v.view(np.int32)) - int(np.float32(base).view(np.int32))) <= 2
    with orc.transcendental_policy("plus", 2, funcs=("cos",)):
        assert orc.pow_probe(x, y, "f32") == base          # pow is not in the mask
    assert orc.pow_probe(x, y, "f32") == base              # the context manager restores the default
    # pow the way Zig's std.math.pow / Go's math.Pow structure it (square-and-multiply + exp(yf log x)), all in f32:
    # within a few ulps of the correctly rounded value, exact for small integer powers of small integers
    rng = np.random.default_rng(3)
    with orc.transcendental_policy("zig_style_pow"):
        assert orc.pow_probe(3.0, 4.0, "f32") == 81.0 and orc.pow_probe(2.0, -3.0, "f32") == 0.125
        assert orc.pow_probe(9.0, 0.5, "f32") == 3.0 and orc.pow_probe(1.0, 123.4, "f32") == 1.0 and orc.pow_probe(5.5, 0.0, "f32") == 1.0
        worst = 0
        for _ in range(2000):
            xx, yy = float(np.float32(rng.uniform(1.0, 60.0))), float(np.float32(rng.uniform(-6.0, 6.0)))
            got = np.float32(orc.pow_probe(xx, yy, "f32"))
            ref = np.float32(np.float64(xx) ** np.float64(yy))
            worst = max(worst, abs(int(got.view(np.int32)) - int(ref.view(np.int32))))
        assert 1 <= worst <= 16, worst    # not correctly rounded (that is the point), but close


def test_transcendental_sensitivity_bounds(oracle_table):
    """What DESIGN.md section 2 states: in f64 a few ulps in cos / pow / log10 move no bin by more than 1e-9; in f32 the
    same perturbation of pow / log10 stays below 1e-5 of the peak, while ONE ulp in cos(inclination) can move isolated
    bins by percents (Q1: every fine bin that reaches the Gauss rule receives 0.418 f(centre) whatever its width, so a
    clamped sliver that appears or disappears is a full-sized term) -- parity with the Zig binary in f32 is bit-exact
    only where the three functions agree bit for bit."""
    E = syn.energy_grid_linear(1000)
    p = syn.random_params("kline5", 12, seed=20260)

    def run(prec):
        return oracle_table.eval_batch("kline5", p, E, precision=prec, threads=4)["out"]

    def peak_change(y, y0):
        return float(np.max(np.abs(y - y0) / np.max(np.abs(y0), axis=1, keepdims=True)))

    b64, b32 = run("f64"), run("f32")
    with orc.transcendental_policy("random", 4):
        assert peak_change(run("f64"), b64) < 1e-9
    with orc.transcendental_policy("random", 4, funcs=("pow", "log10")):
        assert peak_change(run("f32"), b32) < 1e-5
    with orc.transcendental_policy("zig_style_pow"):
        assert peak_change(run("f32"), b32) < 1e-5
        assert peak_change(run("f64"), b64) < 1e-12
