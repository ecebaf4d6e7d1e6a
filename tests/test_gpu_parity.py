"""GPU parity tests (run on the B200 box): the CUDA path, called through the C ABI, against the CPU oracle
on the same seeded inputs, against the committed golden vectors, and -- at BASELINE.json's full batch
sizes -- through size-independent properties.

Tolerance (BASELINE.json north_star / SURVEY.md §8d): every bin |x - y| <= 1e-6 |y| + 1e-12 max|y|.
In f32 ("reference-faithful") mode the CUDA path issues the oracle's operations in the oracle's order,
so additionally the fine-grid flux must be bit-identical in (almost) every bin; the only permitted
source of differences is the last f64 ulp of cos/pow/log10 (CUDA vs glibc).
"""
import os

import numpy as np
import pytest

import lineprofiles_b200 as klp
from lineprofiles_b200 import synthetic as syn
from lineprofiles_b200 import tableio
from oracle import oracle_py as orc

pytestmark = pytest.mark.gpu

RTOL, AFLOOR = 1e-6, 1e-12
MODELS = ("kline", "kline5", "kconv", "kconv5", "kconv10")
GOLD = os.path.join(os.path.dirname(__file__), "golden", "oracle_golden.npz")


def assert_parity(got, want, what=""):
    got, want = np.asarray(got), np.asarray(want)
    assert got.shape == want.shape
    nan_g, nan_w = np.isnan(got), np.isnan(want)
    assert np.array_equal(nan_g, nan_w), f"{what}: NaN pattern differs"
    g, w = np.where(nan_g, 0, got), np.where(nan_w, 0, want)
    scale = np.max(np.abs(w), axis=-1, keepdims=True)
    bad = np.abs(g - w) > RTOL * np.abs(w) + AFLOOR * scale
    assert not bad.any(), f"{what}: {bad.sum()} bins out of tolerance, worst rel {np.max(np.abs(g - w) / (np.abs(w) + AFLOOR * scale)):.3e}"


def grids(model):
    if model.startswith("kconv"):
        E = syn.energy_grid_log(512)
        return E, syn.powerlaw_continuum(E)
    return syn.energy_grid_linear(1000), None


def test_exact_arithmetic_sequences(gpu_table):
    """The FAST division / square-root sequences equal the IEEE intrinsics on 2^33 operands per kind."""
    m = gpu_table.arith_selftest(1 << 33, seed=20261019)
    assert m == {"f32_div_by": 0, "f64_div_by": 0, "f32_div_gen": 0, "f32_sqrt_gen": 0}, m


def test_fast_and_strict_paths_identical(gpu_table):
    """Option "fast" = 0 forces the plain IEEE-intrinsic path; results must be bit-identical."""
    E = syn.energy_grid_linear(1000)
    p = syn.random_params("kline5", 256, seed=91)
    try:
        for prec in ("f32", "f64"):
            gpu_table.set_option("fast", 1)
            a = gpu_table.eval_batch("kline5", p, E, precision=prec)
            gpu_table.set_option("fast", 0)
            b = gpu_table.eval_batch("kline5", p, E, precision=prec)
            assert np.array_equal(a, b), prec
    finally:
        gpu_table.set_option("fast", 1)


def test_config1_single_kline(gpu_table, oracle_table):
    """BASELINE config 1: kline a=0.998 incl=30 eline=6.4 rmin=ISCO rout=400 alpha=3 on the 1000-bin grid."""
    E = syn.energy_grid_linear(1000)
    for prec in ("f32", "f64"):
        got = gpu_table.eval_batch("kline", syn.CONFIG1_PARAMS, E, precision=prec)
        want = oracle_table.eval_batch("kline", syn.CONFIG1_PARAMS, E, precision=prec, faithful_cost=True)["out"]
        assert_parity(got, want, f"config1 {prec}")
        assert abs(got.sum() - 1) < 1e-12
    got32 = gpu_table.eval_batch("kline", syn.CONFIG1_PARAMS, E, precision="f32")
    want32 = oracle_table.eval_batch("kline", syn.CONFIG1_PARAMS, E, precision="f32")["out"]
    assert np.array_equal(got32, want32), "f32 mode is expected to be bit-identical on config 1"


@pytest.mark.parametrize("model", MODELS)
@pytest.mark.parametrize("prec", ["f32", "f64"])
def test_random_sets_all_models(gpu_table, oracle_table, model, prec):
    E, spec = grids(model)
    p = syn.random_params(model, 48, seed=2024)
    got = gpu_table.eval_batch(model, p, E, spectrum=spec, precision=prec)
    ref = oracle_table.eval_batch(model, p, E, spectrum=spec, precision=prec, threads=os.cpu_count() or 1, want_fine=True,
                                  want_profile=spec is not None)
    assert_parity(got, ref["out"], f"{model} {prec}")
    fine, prof = gpu_table.eval_debug(model, p, E, precision=prec)
    assert_parity(fine, ref["fine"], f"{model} {prec} fine grid")
    if prof is not None:
        assert_parity(prof, ref["profile"], f"{model} {prec} profile")
    if prec == "f32":
        frac = np.mean(fine == ref["fine"])
        assert frac > 0.999, f"f32 fine flux bit-identical in only {frac:.5f} of the bins"


def test_golden_vectors(gpu_table):
    g = np.load(GOLD)
    for model in MODELS:
        conv = model.startswith("kconv")
        for prec in ("f32", "f64"):
            got = gpu_table.eval_batch(model, g[f"params_{model}"], g["energy_conv"] if conv else g["energy_line"],
                                       spectrum=g["spectrum"] if conv else None, precision=prec)
            assert_parity(got, g[f"out_{model}_{prec}"], f"golden {model} {prec}")


def test_reference_smoke_vectors(gpu_table, oracle_table):
    """The parameter vectors of xspec-wrapper.zig:553-585 (including the stale kline order that puts
    rmin=3 > rmax=1, and rmin = rmax = 0 which makes every weight NaN -> all-NaN output, Q10)."""
    dom, _ = orc.range_iterator(0.1, 10.0, 100, "f64")
    vecs = np.array([[0.998, 60, 6.4, 3.0, 1.0, 400.0], [0.998, 0.1, 6.4, 3.0, 1.0, 400.0], [0.998, 89.0, 6.4, 3.0, 1.0, 400.0],
                     [0.998, 60.0, 6.4, 0.0, 1.0, 400.0], [0.998, 60.0, 6.4, 0.0, 0.0, 400.0]])
    for prec in ("f32", "f64"):
        got = gpu_table.eval_batch("kline", vecs, dom, precision=prec)
        want = oracle_table.eval_batch("kline", vecs, dom, precision=prec)["out"]
        assert_parity(got, want, f"smoke kline {prec}")
        flux = np.ones(99)
        for model, v in (("kconv5", [0.998, 60.0, 1.0, 400.0, 100.0, 3.0, 3.0, 3.0, 3.0, 3.0, 3.0]),
                         ("kconv10", [0.998, 60.0, 1.0, 400.0, 50.0, 3.0] + [3.0] * 10)):
            got = gpu_table.eval_batch(model, np.array([v]), dom, spectrum=flux, precision=prec)
            want = oracle_table.eval_batch(model, np.array([v]), dom, spectrum=flux, precision=prec)["out"]
            assert_parity(got, want, f"smoke {model} {prec}")


def test_edge_parameters(gpu_table, oracle_table):
    """Out-of-table spin / inclination (Q5 snap), grazing and face-on inclinations, rmin > rmax, zero knots."""
    E = syn.energy_grid_linear(400)
    base = syn.random_params("kline5", 10, seed=31)
    base[0, 0] = 1.7          # above the spin grid -> first frame
    base[1, 0] = -0.9         # below the spin grid
    base[2, 1] = 0.0          # cos = 1 above the mu grid
    base[3, 1] = 90.0         # cos ~ 0 below the mu grid
    base[4, 3:5] = (30.0, 10.0)   # rmin > rmax
    base[5, 7:] = 0.0         # all emissivity knots zero (the reference's "with zero emissivities" case)
    base[6, 5] = 1.0          # rcut = 1 -> log10(rcut) = 0, every knot radius = 1
    base[7, 3:5] = (1.0, 1000.0)
    base[8, 2] = 0.5          # eline far below the grid centre
    base[9, 6] = 10.0         # steep outer index
    for prec in ("f32", "f64"):
        got = gpu_table.eval_batch("kline5", base, E, precision=prec)
        want = oracle_table.eval_batch("kline5", base, E, precision=prec)["out"]
        assert_parity(got, want, f"edge params {prec}")


def test_irregular_fine_grids(gpu_table, oracle_table):
    """Sets whose fine grid in g is not an ascending finite sequence take the general path of the quadrature (every
    bin a candidate, NaN-aware clamps): negative eline (descending grid), NaN eline, huge eline (all bins below gmin),
    tiny eline (all above gmax, grid wider than a knot spacing)."""
    E = syn.energy_grid_linear(300)
    p = syn.random_params("kline5", 8, seed=77)
    p[0, 2] = -6.4
    p[1, 2] = np.nan
    p[2, 2] = 1e6
    p[3, 2] = 1e-3
    p[4, 2] = 0.5         # bins ~13 times wider in g: wider than a knot spacing at some radii -> no node sharing
    p[5, 2] = np.inf
    for prec in ("f32", "f64"):
        got = gpu_table.eval_batch("kline5", p, E, precision=prec)
        want = oracle_table.eval_batch("kline5", p, E, precision=prec)["out"]
        assert_parity(got, want, f"irregular grids {prec}")
    # the options that move the blended frame / change the CTA shape must not change a bit
    base = gpu_table.eval_batch("kline5", p, E, precision="f32")
    try:
        for key, val in (("frame_global", 1), ("quad_threads", 256), ("quad_threads", 512), ("quad_threads", 1024), ("deposit_ring", 1)):
            gpu_table.set_option(key, val)
            assert np.array_equal(gpu_table.eval_batch("kline5", p, E, precision="f32"), base, equal_nan=True), (key, val)
            gpu_table.set_option(key, 0)
    finally:
        gpu_table.set_option("frame_global", 0)
        gpu_table.set_option("quad_threads", 0)
        gpu_table.set_option("deposit_ring", 0)


@pytest.mark.parametrize("n_bins", [1, 2, 7, 100, 2500, 5000])
def test_ragged_energy_grids(gpu_table, oracle_table, n_bins):
    rng = np.random.default_rng(n_bins)
    E = np.sort(rng.uniform(0.3, 9.0, n_bins + 1))
    E += np.arange(n_bins + 1) * 1e-9   # strictly ascending
    p = syn.random_params("kline", 3, seed=n_bins)
    got = gpu_table.eval_batch("kline", p, E, precision="f32")
    want = oracle_table.eval_batch("kline", p, E, precision="f32")["out"]
    assert_parity(got, want, f"ragged kline n={n_bins}")
    spec = rng.uniform(0, 1, n_bins)
    pc = syn.random_params("kconv", 3, seed=n_bins)
    got = gpu_table.eval_batch("kconv", pc, E, spectrum=spec, precision="f32")
    want = oracle_table.eval_batch("kconv", pc, E, spectrum=spec, precision="f32")["out"]
    assert_parity(got, want, f"ragged kconv n={n_bins}")


def test_config3_grid_kconv_4096(gpu_table, oracle_table):
    """BASELINE config 3 grid (4096-bin log grid, shared power-law continuum), a few sets against the oracle."""
    E = syn.energy_grid_log(4096)
    spec = syn.powerlaw_continuum(E)
    p = syn.random_params("kconv", 4, seed=20261)
    got = gpu_table.eval_batch("kconv", p, E, spectrum=spec, precision="f32")
    want = oracle_table.eval_batch("kconv", p, E, spectrum=spec, precision="f32", threads=4)["out"]
    assert_parity(got, want, "kconv 4096")


def test_per_set_spectrum_equals_shared(gpu_table):
    E = syn.energy_grid_log(300)
    spec = syn.powerlaw_continuum(E)
    p = syn.random_params("kconv10", 5, seed=3)
    a = gpu_table.eval_batch("kconv10", p, E, spectrum=spec, precision="f32")
    b = gpu_table.eval_batch("kconv10", p, E, spectrum=np.tile(spec, (5, 1)), precision="f32")
    assert np.array_equal(a, b)


def test_invariance_to_batching_knobs(gpu_table):
    """Sets are independent: the result of a set must not depend on batch order, chunking, CTA splitting or CTA size."""
    E = syn.energy_grid_linear(1000)
    p = syn.random_params("kline5", 700, seed=55)
    base = gpu_table.eval_batch("kline5", p, E, precision="f32")
    perm = np.random.default_rng(0).permutation(700)
    assert np.array_equal(gpu_table.eval_batch("kline5", p[perm], E, precision="f32"), base[perm])
    try:
        for key, val in (("chunk", 97), ("splits", 5), ("quad_threads", 256), ("quad_threads", 512), ("quad_threads", 1024), ("splits", 63), ("splits", 300), ("frame_global", 1), ("deposit_ring", 1), ("order", 0)):
            gpu_table.set_option(key, val)
            assert np.array_equal(gpu_table.eval_batch("kline5", p, E, precision="f32"), base), (key, val)
            gpu_table.set_option(key, {"chunk": 131072, "splits": 0, "quad_threads": 0, "frame_global": 0, "deposit_ring": 0, "order": 1}[key])
        # the deposit ring with and without discarding the consumed lines from the L2 (a discarded line that is reused must
        # never be read before it is rewritten: repeat, in both precisions, with 256- and 1024-thread CTAs)
        base64 = gpu_table.eval_batch("kline5", p, E, precision="f64")
        gpu_table.set_option("deposit_ring", 1)
        for disc in (1, 0, 1):
            gpu_table.set_option("ring_discard", disc)
            for nt in (0, 256):
                gpu_table.set_option("quad_threads", nt)
                assert np.array_equal(gpu_table.eval_batch("kline5", p, E, precision="f32"), base), ("ring_discard", disc, nt)
                assert ("discard=1" in gpu_table.last_launch_info()) == bool(disc)
            gpu_table.set_option("quad_threads", 0)
            assert np.array_equal(gpu_table.eval_batch("kline5", p, E, precision="f64"), base64), ("ring_discard f64", disc)
    finally:
        for key, val in (("chunk", 131072), ("splits", 0), ("quad_threads", 0), ("frame_global", 0), ("deposit_ring", 0), ("order", 1), ("ring_discard", 1)):
            gpu_table.set_option(key, val)


def test_device_entry_equals_host_entry(gpu_table):
    import torch
    E = syn.energy_grid_linear(1000)
    p = syn.random_params("kline5", 300, seed=8)
    host = gpu_table.eval_batch("kline5", p, E, precision="f64")
    dp, dE = torch.from_numpy(p).cuda(), torch.from_numpy(E).cuda()
    out = torch.empty((300, 1000), dtype=torch.float64, device="cuda")
    gpu_table.eval_batch_device("kline5", dp, dE, 1000, out, 300, precision="f64", stream=torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    assert np.array_equal(out.cpu().numpy(), host)


def test_blend_kernel_matches_oracle(gpu_table, oracle_table, table_arrays):
    import torch
    rng = np.random.default_rng(12)
    ai = np.stack([rng.uniform(-0.2, 1.1, 64), rng.uniform(0, 90, 64)], axis=1)
    n_r, n_g = table_arrays["n_r"], table_arrays["n_g"]
    for prec, dt in (("f32", torch.float32), ("f64", torch.float64)):
        out = torch.empty((64, n_r * (2 * n_g + 3)), dtype=dt, device="cuda")
        gpu_table.blend_frames_device(torch.from_numpy(ai).cuda(), out, 64, precision=prec)
        torch.cuda.synchronize()
        got = out.cpu().numpy().astype(np.float64)
        for s in range(64):
            fr = oracle_table.blend_frame(ai[s, 0], ai[s, 1], prec)
            want = np.concatenate([fr["radii"], fr["gmin"], fr["gmax"], fr["upper"].ravel(), fr["lower"].ravel()])
            # f32: same operations in the same order -> bit-identical; f64: cos() differs in its last ulp
            # between CUDA and glibc, which perturbs the mu weight at the 1e-16 level
            if prec == "f32":
                assert np.mean(got[s] == want) > 0.999, (prec, s)
            assert np.allclose(got[s], want, rtol=1e-6 if prec == "f32" else 1e-12), (prec, s)


def test_big_table_rows_in_global_scratch(oracle_table):
    """A frame too large for shared memory (n_r * (2 n_g + 3) * 4 B > 227 KB) takes the per-CTA scratch path."""
    tab = syn.make_table(4, 3, 700, 48)
    gt = klp.Table.from_arrays(tab, 0)
    ot = orc.OracleTable(tab)
    E = syn.energy_grid_linear(200)
    p = syn.random_params("kline", 6, seed=1)
    assert_parity(gt.eval_batch("kline", p, E, precision="f32"), ot.eval_batch("kline", p, E, precision="f32")["out"], "scratch f32")
    assert_parity(gt.eval_batch("kline", p, E, precision="f64"), ot.eval_batch("kline", p, E, precision="f64")["out"], "scratch f64")
    gt.close()


def test_table_file_load_on_device(tmp_path, gpu_table, table_arrays):
    p = str(tmp_path / "kerr-transfer-functions.klpt")
    tableio.save_table(p, table_arrays)
    t2 = klp.Table.load(p, 0)
    E = syn.energy_grid_linear(100)
    a = t2.eval_batch("kline", syn.CONFIG1_PARAMS, E)
    assert np.array_equal(a, gpu_table.eval_batch("kline", syn.CONFIG1_PARAMS, E))
    t2.close()


def test_xspec_symbols_and_ratchet(gpu_table, oracle_table):
    """The five exported XSPEC symbols, including the reference's radial-limit ratchet (Q2) across calls."""
    klp.set_default_table(gpu_table)
    klp.set_legacy_precision("f32")
    klp.legacy_reset()
    try:
        E = syn.energy_grid_linear(100)
        calls = np.array([[0.5, 40, 6.4, 5.0, 20.0, 3.0], [0.5, 40, 6.4, 1.0, 400.0, 3.0], [0.9, 70, 6.4, 8.0, 15.0, 2.0]])
        lo, hi = orc.base_limits("f32")
        want = oracle_table.eval_batch("kline", calls, E, precision="f32", ratchet=True, lims=[lo, hi])["out"]
        for k in range(3):
            got = klp.kline(E, calls[k])
            assert_parity(got, want[k], f"ratchet call {k}")
        klp.legacy_reset()
        fresh = oracle_table.eval_batch("kline", calls[1:2], E, precision="f32")["out"][0]
        assert_parity(klp.kline(E, calls[1]), fresh, "after reset")
        # the other four symbols, fresh process each
        spec = syn.powerlaw_continuum(E)
        for name, fn in (("kline5", klp.kline5), ("kconv", klp.kconv), ("kconv5", klp.kconv5), ("kconv10", klp.kconv10)):
            klp.legacy_reset()
            p = syn.random_params(name, 1, seed=99)[0]
            conv = name.startswith("kconv")
            flux = spec.copy() if conv else None
            got = fn(E, p, flux)
            want = oracle_table.eval_batch(name, p[None], E, spectrum=spec if conv else None, precision="f32")["out"][0]
            assert_parity(got, want, name)
    finally:
        klp.set_default_table(None)
        klp.legacy_reset()


def test_full_batch_properties_config2(gpu_table, oracle_table):
    """BASELINE config 2 at full size (kline5, B = 1e5): properties that need no oracle at that size, plus an
    oracle comparison of a random subsample."""
    B = 100000
    E = syn.energy_grid_linear(1000)
    p = syn.random_params("kline5", B)
    out = gpu_table.eval_batch("kline5", p, E, precision="f32")
    assert out.shape == (B, 1000) and np.isfinite(out).all() and (out >= 0).all()
    np.testing.assert_allclose(out.sum(axis=1), 1.0, rtol=1e-12)     # Q10: every spectrum sums to one
    idx = np.random.default_rng(1).choice(B, 64, replace=False)
    want = oracle_table.eval_batch("kline5", p[idx], E, precision="f32", threads=os.cpu_count() or 1)["out"]
    assert_parity(out[idx], want, "config 2 subsample")
    # idempotence: a second evaluation of a slice reproduces the batch result bit for bit
    assert np.array_equal(gpu_table.eval_batch("kline5", p[5000:5600], E, precision="f32"), out[5000:5600])


def test_conv_linearity_config3_grid(gpu_table):
    """The convolution is linear in the input spectrum: conv(a*s1 + s2) == a*conv(s1) + conv(s2) to rounding."""
    E = syn.energy_grid_log(4096)
    s1 = syn.powerlaw_continuum(E)
    s2 = np.zeros(4096)
    s2[[1000, 2500, 3000]] = 1e-2
    p = syn.random_params("kconv", 256, seed=20261)
    c1 = gpu_table.eval_batch("kconv", p, E, spectrum=s1, precision="f32")
    c2 = gpu_table.eval_batch("kconv", p, E, spectrum=s2, precision="f32")
    c3 = gpu_table.eval_batch("kconv", p, E, spectrum=3.0 * s1 + s2, precision="f32")
    np.testing.assert_allclose(c3, 3.0 * c1 + c2, rtol=1e-9, atol=1e-18)


def test_conv_mode_cumulative_profile(gpu_table, oracle_table):
    """Option "conv_mode" = 1 (cumulative-profile formulation of convolution.zig:58-102, a different operation
    order): within 1e-9 relative + 1e-12 of the peak of the reference-order result on linear, log and ragged grids,
    shared and per-set spectra, and against the oracle at the BASELINE tolerance."""
    cases = [("kconv", syn.energy_grid_log(4096), 6), ("kconv10", syn.energy_grid_log(300), 5),
             ("kconv5", syn.energy_grid_linear(1000), 5), ("kconv", syn.energy_grid_linear(7), 3),
             ("kconv", syn.energy_grid_linear(1), 2)]
    try:
        for model, E, nb in cases:
            n = E.size - 1
            spec = syn.powerlaw_continuum(E)
            p = syn.random_params(model, nb, seed=20261 + n)
            per_set = np.tile(spec, (nb, 1)) * np.linspace(0.5, 2.0, nb)[:, None]
            for sp in (spec, per_set):
                gpu_table.set_option("conv_mode", 0)
                exact = gpu_table.eval_batch(model, p, E, spectrum=sp, precision="f32")
                gpu_table.set_option("conv_mode", 1)
                scale = np.max(np.abs(exact), axis=-1, keepdims=True)
                # both variants of the kernel: per-source {avg, spectrum} pairs staged in shared memory (default) or read from L1
                for src_smem in (1, 0):
                    gpu_table.set_option("conv_src_smem", src_smem)
                    fast = gpu_table.eval_batch(model, p, E, spectrum=sp, precision="f32")
                    err = np.abs(fast - exact) / (np.abs(exact) + 1e-3 * scale)
                    assert np.max(err) < 1e-9, (model, n, src_smem, float(np.max(err)))
                gpu_table.set_option("conv_src_smem", 1)
            want = oracle_table.eval_batch(model, p, E, spectrum=spec, precision="f32", threads=4)["out"]
            gpu_table.set_option("conv_mode", 1)
            assert_parity(gpu_table.eval_batch(model, p, E, spectrum=spec, precision="f32"), want, f"conv_mode 1 {model} n={n}")
    finally:
        gpu_table.set_option("conv_mode", 0)
        gpu_table.set_option("conv_src_smem", 1)


def test_fits_table_evaluates_like_arrays(tmp_path, gpu_table, table_arrays):
    """The FITS route (reference schema) and the array route give the same table on the device."""
    p = str(tmp_path / "kerr-transfer-functions.fits")
    tableio.save_fits(p, table_arrays)
    t2 = klp.Table.load(p, 0)
    E = syn.energy_grid_linear(200)
    q = syn.random_params("kline5", 5, seed=4)
    assert np.array_equal(t2.eval_batch("kline5", q, E), gpu_table.eval_batch("kline5", q, E))
    t2.close()


def test_setup_cache_across_calls(gpu_table):
    """The host pipeline keeps the per-call constants of the last energy grid on the device (k_setup is skipped when the
    grid, the precision and the model kind repeat): interleaved grids, precisions, model kinds and device-API calls must
    not change a result."""
    import torch
    E1, E2 = syn.energy_grid_linear(300), syn.energy_grid_log(200)
    p = syn.random_params("kline5", 5, seed=12)
    pc = syn.random_params("kconv5", 5, seed=13)
    s2 = syn.powerlaw_continuum(E2)
    s1 = syn.powerlaw_continuum(E1)
    ref = {}
    for key, fn in (("l1", lambda: gpu_table.eval_batch("kline5", p, E1, precision="f32")),
                    ("l2", lambda: gpu_table.eval_batch("kline5", p, E2, precision="f32")),
                    ("l1d", lambda: gpu_table.eval_batch("kline5", p, E1, precision="f64")),
                    ("c2", lambda: gpu_table.eval_batch("kconv5", pc, E2, spectrum=s2, precision="f32")),
                    ("c1", lambda: gpu_table.eval_batch("kconv5", pc, E1, spectrum=s1, precision="f32"))):
        gpu_table.eval_batch("kline", syn.CONFIG1_PARAMS, syn.energy_grid_linear(17))   # invalidate with another grid
        ref[key] = (fn, fn())
    order = ["l1", "l1", "c1", "c1", "l1", "l1d", "l1d", "l1", "l2", "c2", "c2", "l2", "c1", "l1"]
    for k in order:
        fn, want = ref[k]
        assert np.array_equal(fn(), want, equal_nan=True), k
    # a device-API call in between overwrites the constants: the next host call must notice
    dp, dE = torch.from_numpy(p).cuda(), torch.from_numpy(E2).cuda()
    out = torch.empty((5, 200), dtype=torch.float64, device="cuda")
    fn, want = ref["l1"]
    assert np.array_equal(fn(), want)
    gpu_table.eval_batch_device("kline5", dp, dE, 200, out, 5, precision="f32", stream=torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    assert np.array_equal(fn(), want)
    assert np.array_equal(out.cpu().numpy(), ref["l2"][1])


# ---- round 2: BASELINE configs 4 and 5 at their stated shapes, irregular caller grids, host pipeline chunks ----------

@pytest.mark.parametrize("prec", ["f32", "f64"])
def test_config4_kconv10_4096_log_bins(gpu_table, oracle_table, prec):
    """BASELINE config 4 shape: kconv10 walkers (seed 20262) on the 4096-bin log grid with the shared E^-2 continuum,
    default (reference-order) convolution, both precisions, against the oracle."""
    E = syn.energy_grid_log(4096)
    spec = syn.powerlaw_continuum(E)
    p = syn.random_params("kconv10", 8, seed=20262)
    got = gpu_table.eval_batch("kconv10", p, E, spectrum=spec, precision=prec)
    want = oracle_table.eval_batch("kconv10", p, E, spectrum=spec, precision=prec, threads=os.cpu_count() or 1)["out"]
    assert_parity(got, want, f"config 4 kconv10 4096 {prec}")
    if prec == "f32":
        assert np.array_equal(got, want), "f32: reference operation order end to end -> bit-identical"


def test_config5_gigabyte_table():
    """BASELINE config 5 shape: a table above 1 GB (96 x 64 frames of 450 radii x 48 g* knots = 1.02 GiB) resident in HBM:
    32 random kline sets against the oracle in f32 and 8 in f64, and the stand-alone blend kernel against the oracle's
    interpolate_parameters on the same table."""
    import torch
    tab = syn.make_table(96, 64, 450, 48)
    nbytes = sum(tab[k].nbytes for k in ("radii", "gmin", "gmax", "upper", "lower"))
    assert nbytes >= 1 << 30
    gt = klp.Table.from_arrays(tab, 0)
    ot = orc.OracleTable(tab)
    try:
        E = syn.energy_grid_linear(1000)
        p = syn.random_params("kline", 32, seed=20259)
        nt = os.cpu_count() or 1
        got = gt.eval_batch("kline", p, E, precision="f32")
        want = ot.eval_batch("kline", p, E, precision="f32", threads=nt)["out"]
        assert_parity(got, want, "config 5 f32")
        assert np.array_equal(got, want)
        assert_parity(gt.eval_batch("kline", p[:8], E, precision="f64"),
                      ot.eval_batch("kline", p[:8], E, precision="f64", threads=nt)["out"], "config 5 f64")
        ai = np.ascontiguousarray(p[:16, :2])
        n = tab["n_r"] * (2 * tab["n_g"] + 3)
        out = torch.empty((16, n), dtype=torch.float32, device="cuda")
        gt.blend_frames_device(torch.from_numpy(ai).cuda(), out, 16, precision="f32")
        torch.cuda.synchronize()
        g = out.cpu().numpy().astype(np.float64)
        for s in range(16):
            fr = ot.blend_frame(ai[s, 0], ai[s, 1], "f32")
            w = np.concatenate([fr["radii"], fr["gmin"], fr["gmax"], fr["upper"].ravel(), fr["lower"].ravel()])
            assert np.array_equal(g[s], w), s
    finally:
        gt.close()


def _irregular_edge_sets(n):
    rng = np.random.default_rng(5)
    asc = np.sort(rng.uniform(0.3, 9.0, n + 1))
    cases = {"descending": asc[::-1].copy(), "shuffled": rng.permutation(asc), "repeated": np.repeat(asc[: (n + 2) // 2], 2)[: n + 1].copy(),
             "with_nan": asc.copy(), "negative_start": asc - 1.0, "zero_start": np.concatenate([[0.0], asc[1:]])}
    cases["with_nan"][n // 2] = np.nan
    return cases


@pytest.mark.parametrize("prec", ["f32", "f64"])
def test_irregular_caller_edges_compute_like_the_reference(gpu_table, oracle_table, prec):
    """The reference evaluates whatever edges it is handed (its loops only scan).  Edges that are not strictly ascending
    and positive switch k_rebin / k_conv from bisection to the reference's scans: same values, NaN pattern included."""
    n = 48
    p = syn.random_params("kline5", 3, seed=31)
    pc = syn.random_params("kconv5", 3, seed=32)
    for name, E in _irregular_edge_sets(n).items():
        got = gpu_table.eval_batch("kline5", p, E, precision=prec)
        want = oracle_table.eval_batch("kline5", p, E, precision=prec)["out"]
        assert_parity(got, want, f"kline5 {name} {prec}")
        spec = np.linspace(1.0, 2.0, n)
        for mode in (0, 1):
            gpu_table.set_option("conv_mode", mode)
            try:
                got = gpu_table.eval_batch("kconv5", pc, E, spectrum=spec, precision=prec)
            finally:
                gpu_table.set_option("conv_mode", 0)
            want = oracle_table.eval_batch("kconv5", pc, E, spectrum=spec, precision=prec)["out"]
            assert_parity(got, want, f"kconv5 {name} {prec} conv_mode {mode}")


def test_host_pipeline_multiple_staging_chunks(gpu_table):
    """stage_mb = 1 forces several staging chunks through the double-buffered pinned slots (three streams, slot reuse):
    the result must equal the single-chunk evaluation bit for bit -- additive model and convolution with per-set spectra."""
    B = 700
    E = syn.energy_grid_linear(1000)
    p = syn.random_params("kline5", B, seed=41)
    one = gpu_table.eval_batch("kline5", p, E, precision="f32")
    E2 = syn.energy_grid_log(600)
    pc = syn.random_params("kconv", B, seed=42)
    spec = np.random.default_rng(43).uniform(0.5, 1.5, (B, 600))
    one_c = gpu_table.eval_batch("kconv", pc, E2, spectrum=spec, precision="f32")
    gpu_table.set_option("stage_mb", 1)
    try:
        # 131 sets per slot: five full chunks + 13 sets + the short final chunk (stage_tail, capped at a quarter slot: 32 sets)
        assert np.array_equal(gpu_table.eval_batch("kline5", p, E, precision="f32"), one)
        assert np.array_equal(gpu_table.eval_batch("kconv", pc, E2, spectrum=spec, precision="f32"), one_c)
        for tail in (0, 1, 7):   # equal chunks; one-set and ragged final chunks
            gpu_table.set_option("stage_tail", tail)
            assert np.array_equal(gpu_table.eval_batch("kline5", p, E, precision="f32"), one), tail
    finally:
        gpu_table.set_option("stage_mb", 256)
        gpu_table.set_option("stage_tail", 4096)


def test_installed_table_closed_then_xspec_symbol_zero_fills(table_arrays, capfd):
    """A table that is destroyed while still installed for the XSPEC symbols is uninstalled by the library: the next call
    takes the reference's error path (message + zero-filled output, xspec-wrapper.zig:361-372), no dangling pointer."""
    t = klp.Table.from_arrays(table_arrays, device=0)
    klp.set_default_table(t)
    E = syn.energy_grid_linear(50)
    old = os.environ.get("KLINE_PROF_DATA_DIR")
    os.environ["KLINE_PROF_DATA_DIR"] = "/nonexistent-klp-dir"
    try:
        assert klp.kline(E, syn.CONFIG1_PARAMS).sum() > 0
        t.close()
        flux = np.full(50, 7.0)
        klp.kline(E, syn.CONFIG1_PARAMS, flux)
        assert np.all(flux == 0.0)
        err = capfd.readouterr().err
        assert "kerrlineprofile: " in err and "MODEL OUTPUT SET TO ZERO" in err and "KLINE_PROF_DATA_DIR" in err
    finally:
        klp.set_default_table(None)
        klp.legacy_reset()
        if old is None:
            del os.environ["KLINE_PROF_DATA_DIR"]
        else:
            os.environ["KLINE_PROF_DATA_DIR"] = old


def test_table_with_zeroed_rows_keeps_fast_path_elsewhere(oracle_table, table_arrays):
    """The reference zero-fills radii a frame does not cover.  Exact zeros in the branches pass the operand-range check of the
    branch-free arithmetic (a zero numerator is exact in every sequence; what the blend makes of a zero next to its neighbours
    is checked per set on the blended frame), tiny non-zero values do not: only sets that blend one of THOSE frames take the
    plain-intrinsic path, and either way the values equal the oracle's."""
    tab = {k: (v.copy() if isinstance(v, np.ndarray) else v) for k, v in table_arrays.items()}
    n_mu = tab["n_mu"]
    for f in (3 * n_mu + 5, 3 * n_mu + 6, 17 * n_mu + 11):
        tab["upper"][f, :4] = 0.0
        tab["lower"][f, :4] = 0.0
    gt = klp.Table.from_arrays(tab, 0)
    ot = orc.OracleTable(tab)
    try:
        E = syn.energy_grid_linear(1000)
        p = syn.random_params("kline5", 96, seed=51)
        p[:8, 0] = tab["spins"][3] - 1e-3          # sets that blend the zeroed frames
        p[:8, 1] = np.rad2deg(np.arccos(tab["mus"][5])) - 0.3
        got = gt.eval_batch("kline5", p, E, precision="f32")
        want = ot.eval_batch("kline5", p, E, precision="f32", threads=os.cpu_count() or 1)["out"]
        assert np.array_equal(got, want, equal_nan=True)
        assert "fast_frames=600/600" in gt.last_launch_info()
    finally:
        gt.close()
    # tiny non-zero branch values (below 2^-20) still take a frame out of the branch-free arithmetic; same values either way
    tab["upper"][17 * n_mu + 11, :4] = 1e-9
    tab["lower"][3 * n_mu + 5, 2:4] = 3e-8
    gt = klp.Table.from_arrays(tab, 0)
    ot = orc.OracleTable(tab)
    try:
        got = gt.eval_batch("kline5", p, E, precision="f32")
        want = ot.eval_batch("kline5", p, E, precision="f32", threads=os.cpu_count() or 1)["out"]
        assert np.array_equal(got, want, equal_nan=True)
        assert "fast_frames=598/600" in gt.last_launch_info()
    finally:
        gt.close()


@pytest.mark.parametrize("prec", ["f32", "f64"])
def test_small_batch_cooperative_mode_is_bit_identical(gpu_table, oracle_table, prec):
    """Batches much smaller than the GPU (a single XSPEC call is the extreme) run as ONE cooperative k_quad launch: items of
    one radius, the ordered additions shared by bin group (deposits staged through shared memory), the rebin fused into the
    last CTA of the set.  Same operations in the same order: equal to the separate-kernel path bit for bit, all five models,
    1 to 9 sets, and equal to the oracle."""
    for model in MODELS:
        E, spec = grids(model)
        for B in (1, 3, 9):
            p = syn.random_params(model, B, seed=700 + B)
            got = gpu_table.eval_batch(model, p, E, spectrum=spec, precision=prec)
            assert "coop=1" in gpu_table.last_launch_info(), gpu_table.last_launch_info()
            gpu_table.set_option("coop", 0)
            try:
                ref = gpu_table.eval_batch(model, p, E, spectrum=spec, precision=prec)
                assert "coop=0" in gpu_table.last_launch_info()
            finally:
                gpu_table.set_option("coop", 1)
            assert np.array_equal(got, ref, equal_nan=True), (model, B)
            if B == 3:
                want = oracle_table.eval_batch(model, p, E, spectrum=spec, precision=prec)["out"]
                assert_parity(got, want, f"coop {model} {prec}")
    # irregular fine grids (negative / NaN eline) and unsorted caller edges through the fused rebin
    E = syn.energy_grid_linear(300)
    p = syn.random_params("kline5", 4, seed=77)
    p[0, 2], p[1, 2] = -6.4, np.nan
    for edges in (E, E[::-1].copy()):
        got = gpu_table.eval_batch("kline5", p, edges, precision=prec)
        want = oracle_table.eval_batch("kline5", p, edges, precision=prec)["out"]
        assert_parity(got, want, f"coop irregular {prec}")


def test_shared_table_handles_evaluate_concurrently(gpu_table):
    """klp_table_share: further handles on the same device frames, each with its own workspaces -- two host threads evaluate at
    the same time (a sampler's walkers) and get what the single handle gives; the frames outlive the handle that loaded them."""
    import threading
    E = syn.energy_grid_linear(1000)
    ps = [syn.random_params("kline5", 1500, seed=900 + k) for k in range(2)]
    want = [gpu_table.eval_batch("kline5", p, E, precision="f32") for p in ps]
    tab = klp.Table.from_arrays(syn.make_table(), 0)
    handles = [tab.share(), tab.share()]
    tab.close()                       # the clones keep the frames alive
    got = [None, None]

    def work(k):
        for _ in range(3):
            got[k] = handles[k].eval_batch("kline5", ps[k], E, precision="f32")

    th = [threading.Thread(target=work, args=(k,)) for k in range(2)]
    for t in th:
        t.start()
    for t in th:
        t.join()
    for k in range(2):
        assert np.array_equal(got[k], want[k])
        handles[k].close()
