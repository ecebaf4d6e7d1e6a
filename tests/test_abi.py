"""CPU tests of the C-ABI library: it loads, exports every symbol include/*.h declares, the host-side
index logic matches the reference's known answers, and evaluation refuses to run without a GPU."""
import ctypes as C
import os
import re

import numpy as np
import pytest

import lineprofiles_b200 as klp
from lineprofiles_b200 import synthetic as syn
from lineprofiles_b200 import tableio

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    text = open(os.path.join(ROOT, "include", "klineprofiles_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    names = re.findall(r"\b([A-Za-z_][A-Za-z0-9_]*)\s*\([^;{]*\)\s*;", text)
    return sorted(set(n for n in names if n.startswith("klp_") or n in klp.MODELS))


def test_library_exports_every_declared_symbol():
    lib = C.CDLL(klp.library_path())
    syms = _declared_symbols()
    assert {"kline", "kline5", "kconv", "kconv5", "kconv10", "klp_eval_batch", "klp_eval_batch_device",
            "klp_table_create", "klp_table_load"} <= set(syms)
    for s in syms:
        assert hasattr(lib, s), f"{s} declared in the header but not exported"


def test_parameter_counts_match_lmodel_dat():
    lib = klp.load_library()
    # reference xspec/lmodel.dat:1,9,23,30,43
    assert [lib.klp_model_nparams(m) for m in range(5)] == [6, 12, 5, 11, 16]
    assert lib.klp_model_nparams(9) == -1


def test_host_index_logic_known_answers(table_arrays):
    """main.zig:19-30 through the product library's host-side mirror (host-only table, no GPU needed)."""
    t = klp.Table.from_arrays(table_arrays, device=-1)
    assert (t.n_a, t.n_mu, t.n_r, t.n_g) == (30, 20, 150, 30)
    assert t.table_index(27, 12) == 552 and t.table_index(0, 12) == 12
    assert t.find_tables([27, 12]) == [532, 552, 531, 551]
    assert t.find_tables([0, 1]) == [1, 1, 0, 0]
    assert t.find_parameter_indices(0.92, 0.4) == [24, 8]
    assert t.find_parameter_indices(5.0, 5.0) == [0, 0]          # Q5
    assert t.parameter_weight(0, 0.1, 0) is None
    w = t.parameter_weight(1, 0.4, 8)
    m = table_arrays["mus"]
    assert abs(w - (np.float32(0.4) - m[7]) / (m[8] - m[7])) < 1e-6
    t.close()


def test_evaluation_refuses_without_device(table_arrays):
    t = klp.Table.from_arrays(table_arrays, device=-1)
    with pytest.raises(klp.KlpError, match="no CPU path"):
        t.eval_batch("kline", syn.CONFIG1_PARAMS, syn.energy_grid_linear(10))
    t.close()


def test_bad_arguments_are_rejected(table_arrays):
    bad = dict(table_arrays)
    bad["spins"] = table_arrays["spins"][::-1].copy()
    with pytest.raises(klp.KlpError, match="ascending"):
        klp.Table.from_arrays(bad, device=-1)
    with pytest.raises(klp.KlpError):
        klp.Table.load("/nonexistent/table.klpt", device=-1)


def test_table_file_roundtrip(tmp_path, table_arrays):
    small = syn.make_table(3, 4, 9, 30)
    p = str(tmp_path / "t.klpt")
    tableio.save_table(p, small)
    back = tableio.load_table(p)
    for k in ("spins", "mus", "radii", "gmin", "gmax", "upper", "lower"):
        assert np.array_equal(back[k], small[k])
    (tmp_path / "bad.klpt").write_bytes(b"NOTATABLE")
    with pytest.raises(klp.KlpError, match="KLPT0001"):
        klp.Table.load(str(tmp_path / "bad.klpt"), device=-1)


def test_xspec_symbols_zero_fill_on_error(tmp_path, capfd, monkeypatch):
    """xspec-wrapper.zig:361-372: failures are logged with the kerrlineprofile prefix and the output zeroed."""
    monkeypatch.setenv("KLINE_PROF_DATA_DIR", str(tmp_path))   # no table there
    klp.set_default_table(None)
    E = syn.energy_grid_linear(20)
    flux = np.full(20, 7.0)
    klp.kline(E, syn.CONFIG1_PARAMS[0], flux)
    err = capfd.readouterr().err
    assert np.all(flux == 0)
    assert "kerrlineprofile: " in err and "MODEL OUTPUT SET TO ZERO" in err and "KLINE_PROF_DATA_DIR" in err


def test_fits_table_in_reference_schema(tmp_path):
    """klp_table_load reads the reference's FITS layout (io.zig:8-87: HDU2 spins, HDU3 cos-inclinations,
    HDU4.. one BINTABLE per frame with r, gmin, gmax, upper_f[30], lower_f[30]) without cfitsio."""
    small = syn.make_table(5, 4, 11, 30)
    p = str(tmp_path / "kerr-transfer-functions.fits")
    tableio.save_fits(p, small)
    t = klp.Table.load(p, device=-1)
    assert (t.n_a, t.n_mu, t.n_r, t.n_g) == (5, 4, 11, 30)
    for a, mu in ((0.3, 0.5), (0.9, 0.1), (2.0, 2.0)):
        want = [0, 0]
        ia = np.nonzero(small["spins"] >= np.float32(a))[0]
        im = np.nonzero(small["mus"] >= np.float32(mu))[0]
        want = [int(ia[0]) if ia.size else 0, int(im[0]) if im.size else 0]
        assert t.find_parameter_indices(a, mu) == want
    t.close()
    # one frame short -> rejected
    raw = open(p, "rb").read()
    (tmp_path / "short.fits").write_bytes(raw[: len(raw) - 2880 * 2])
    with pytest.raises(klp.KlpError):
        klp.Table.load(str(tmp_path / "short.fits"), device=-1)


def test_lmodel_matches_library():
    """xspec/lmodel.dat (generated from lineprofiles_b200.xspec_models) binds c_<name> symbols the library exports
    with the parameter counts the library expects; layout as reference xspec/lmodel.dat:1-59."""
    from lineprofiles_b200 import xspec_models as xm
    lib = klp.load_library()
    text = xm.lmodel_text()   # what xspec/make_lmodel.py writes to xspec/lmodel.dat at build time
    heads = [ln.split() for ln in text.splitlines() if "c_k" in ln]
    assert [h[0] for h in heads] == ["kline", "kline5", "kconv", "kconv5", "kconv10"]
    for h in heads:
        name, npar, func, kind = h[0], int(h[1]), h[4], h[5]
        assert func == "c_" + name and hasattr(lib, name)
        assert npar == lib.klp_model_nparams(klp.MODELS[name]) == len(xm.MODELS[name][1])
        assert kind == ("con" if name.startswith("kconv") else "add")
    assert [p[0] for p in xm.MODELS["kline5"][1]][:7] == ["a", "incl", "eline", "rmin", "rmax", "rcut", "alpha"]
    assert [p[0] for p in xm.MODELS["kconv"][1]] == ["a", "incl", "rmin", "rmax", "alpha"]


def test_committed_ncu_figures_are_what_bench_reads():
    """profiles/r*_quad_traffic.json (tools/ncu_traffic.py) carry the per-spectrum figures bench.py turns into roofline.traffic,
    roofline_issue and roofline_fp32_pipe -- but only for the kernel instantiation the capture is of: bench.py compares the
    capture's kernel name with klp_last_launch_info() and prints null + the reason on a mismatch (round 1 quoted a capture of
    <float, 512, 0> for a <float, 1024, 1> run)."""
    import importlib.util
    import json
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    spec = importlib.util.spec_from_file_location("bench", os.path.join(root, "bench.py"))
    bench = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(bench)
    newest = bench.latest_profile("r*_quad_traffic.json")
    assert os.path.basename(newest) == "r02_quad_traffic.json"
    d = json.load(open(newest))
    for key in ("sets_per_launch", "dram_bytes_read", "dram_bytes_write", "warp_instructions_per_set", "fma_pipe_cycles_per_set"):
        assert float(d[key]) > 0, key
    shipped = "k_quad<float, 1024, 1> splits=1 ring=0 coop=0 fast_frames=600/600"
    assert bench.kernel_matches(d["kernel"], shipped) and int(d.get("ring", 0)) == 0
    # ... and only for the very hot loop the capture was taken from: the one-correction division changed the instruction
    # counts of the same instantiation, so the capture carries the fingerprint of its build
    assert d["hot_loop_fingerprint"] == bench.recorded_fingerprint()
    for superseded in ("r02_quad_two_corrections_traffic_capture.json", "r02_quad_493_instruction_loop_capture.json"):
        stale = json.load(open(os.path.join(root, "profiles", superseded)))
        assert stale.get("hot_loop_fingerprint") != bench.recorded_fingerprint()
    old = json.load(open(os.path.join(root, "profiles", "r01_quad_traffic.json")))
    assert not bench.kernel_matches(old["kernel"], shipped)            # the stale round-1 capture would be refused
    assert not bench.kernel_matches(d["kernel"], "k_quad<double, 512, 1> splits=1 ring=0")
    assert not bench.kernel_matches(d["kernel"], "")


def test_c99_client_links_and_gets_the_xspec_error_convention(tmp_path):
    """A plain C99 translation unit (what XSPEC's initpackage glue is) includes the header with -pedantic, links the
    library and calls `kline` without a table: message with the reference's prefix on stderr, flux zero-filled
    (xspec-wrapper.zig:361-372)."""
    import shutil
    import subprocess
    if shutil.which("gcc") is None:
        pytest.skip("no gcc")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    src = tmp_path / "client.c"
    src.write_text(
        '#include "klineprofiles_b200.h"\n#include <stdio.h>\n'
        "int main(void) {\n"
        "    double e[3] = {1.0, 2.0, 3.0}, p[6] = {0.5, 30, 6.4, 1.0, 400.0, 3.0}, f[2] = {7, 7};\n"
        '    printf("%d %d\\n", klp_model_nparams(KLP_KLINE), klp_model_nparams(KLP_KCONV10));\n'
        '    kline(e, 2, p, 0, f, 0, "");\n'
        '    printf("%g %g\\n", f[0], f[1]);\n'
        "    return 0;\n}\n")
    libdir = os.path.dirname(klp.api.library_path())
    exe = tmp_path / "client"
    cc = subprocess.run(["gcc", "-std=c99", "-Wall", "-Wextra", "-pedantic", "-Werror", "-I", os.path.join(root, "include"), str(src),
                         "-o", str(exe), "-L", libdir, "-lklineprofiles_b200", f"-Wl,-rpath,{libdir}"], capture_output=True, text=True)
    assert cc.returncode == 0, cc.stderr
    env = dict(os.environ, KLINE_PROF_DATA_DIR=str(tmp_path / "nowhere"))
    run = subprocess.run([str(exe)], capture_output=True, text=True, env=env)
    assert run.returncode == 0
    assert run.stdout.split() == ["6", "16", "0", "0"]
    assert "kerrlineprofile: " in run.stderr and "MODEL OUTPUT SET TO ZERO" in run.stderr


# ---- round 2: the FITS reader against a fixture it did not write ----------------------------------------------------

def _fixture():
    import importlib.util
    p = os.path.join(os.path.dirname(__file__), "golden", "make_fits_fixture.py")
    spec = importlib.util.spec_from_file_location("make_fits_fixture", p)
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    return m


def test_fits_fixture_is_reproducible():
    m = _fixture()
    assert open(m.PATH, "rb").read() == m.build(), "tests/golden/fixture_table.fits is stale: rerun make_fits_fixture.py"


def test_fits_reader_loads_the_hand_written_fixture():
    """io.zig:53-87 reads HDU 2 column 1 (spins), HDU 3 column 1 (cos i) and columns 1-5 of every HDU from 4 on.  The fixture
    is written from the FITS standard by tests/golden/make_fits_fixture.py (not by tableio.save_fits): '30E' vector columns
    with TDIM, D-typed scalar columns, an extra J column, EXTNAME/COMMENT/HISTORY cards, two-block frame headers."""
    m = _fixture()
    want = m.fixture_arrays()
    t = klp.Table.load(m.PATH, device=-1)
    assert (t.n_a, t.n_mu, t.n_r, t.n_g) == (want["n_a"], want["n_mu"], want["n_r"], want["n_g"])
    a, mu = t.parameters()
    assert np.array_equal(a, want["spins"]) and np.array_equal(mu, want["mus"])
    for f in range(t.n_a * t.n_mu):
        fr = t.read_frame(f)
        for k in ("radii", "gmin", "gmax", "upper", "lower"):
            assert np.array_equal(fr[k], want[k][f]), (f, k)
    assert t.find_parameter_indices(0.7, 0.5) == [2, 1] and t.find_tables([2, 1]) == [3, 5, 2, 4]
    t.close()


def test_fits_reader_rejects_corrupted_files(tmp_path):
    m = _fixture()
    good = m.build()

    def load(b, name):
        p = tmp_path / name
        p.write_bytes(b)
        return klp.Table.load(str(p), device=-1)

    load(good, "good.fits").close()
    with pytest.raises(klp.KlpError, match="truncated"):                      # cut inside the last frame's data
        load(good[:-2880 - 100], "cut.fits")
    with pytest.raises(klp.KlpError, match="NAXIS1"):                         # a column format that no longer adds up
        load(good.replace(b"TFORM5  = '30E     '", b"TFORM5  = '29E     '", 1), "width.fits")
    with pytest.raises(klp.KlpError, match="unsupported TFORM"):
        load(good.replace(b"TFORM4  = '30E     '", b"TFORM4  = '30C     '", 1), "tform.fits")
    frame_bytes = 2 * 2880 + 2880                                             # one frame HDU fewer than n_a * n_mu
    with pytest.raises(klp.KlpError, match="n_spins"):
        load(good[:-frame_bytes], "short.fits")
    with pytest.raises(klp.KlpError, match="without END"):                    # primary END card wiped out: the next HDU would be swallowed
        load(good[:2880].replace(b"END" + b" " * 77, b" " * 80) + good[2880:], "noend.fits")
    with pytest.raises(klp.KlpError, match="floating point"):                 # spins column turned into integers
        load(good.replace(b"TFORM1  = '1D      '", b"TFORM1  = '1K      '", 1), "ints.fits")
    with pytest.raises(klp.KlpError) as ei:                                   # a missing file is a different status
        klp.Table.load(str(tmp_path / "absent.fits"), device=-1)
    assert "status 5" in str(ei.value)


def test_xspec_static_archive_links_like_the_reference_package(tmp_path):
    """The XSPEC package absorbs a STATIC archive named libxsklineprofiles.a (/root/reference/Makefile:43-48,61-72).  Ours is
    built from the same object as the shared library; a plain C client links it with the line xspec/Makefile puts into the
    generated package Makefile and gets the same behaviour."""
    import shutil
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    arch = os.path.join(root, "xspec", "libxsklineprofiles.a")
    if shutil.which("g++") is None or not os.path.exists("/usr/local/cuda/lib64/libcudart_static.a"):
        pytest.skip("no g++ / static CUDA runtime")
    assert os.path.exists(arch), "run __graft_entry__.build() (or make -C xspec lib)"
    syms = subprocess.run(["nm", "-g", "--defined-only", arch], capture_output=True, text=True).stdout
    for name in ("kline", "kline5", "kconv", "kconv5", "kconv10"):
        assert f" T {name}\n" in syms, name
    src = tmp_path / "client.c"
    src.write_text('#include "klineprofiles_b200.h"\n#include <stdio.h>\nint main(void) { double e[3] = {1, 2, 3}, p[5] = {0.5, 30, 1.0, 400.0, 3.0}, '
                   'f[2] = {7, 7}; kconv(e, 2, p, 0, f, 0, ""); printf("%g %g\\n", f[0], f[1]); return 0; }\n')
    obj, exe = tmp_path / "client.o", tmp_path / "client"
    assert subprocess.run(["gcc", "-std=c99", "-c", "-I", os.path.join(root, "include"), str(src), "-o", str(obj)]).returncode == 0
    ln = subprocess.run(["g++", str(obj), "-o", str(exe), "-L", os.path.dirname(arch), "-l:libxsklineprofiles.a", "-L/usr/local/cuda/lib64",
                         "-lcudart_static", "-ldl", "-lrt", "-lpthread"], capture_output=True, text=True)
    assert ln.returncode == 0, ln.stderr
    run = subprocess.run([str(exe)], capture_output=True, text=True, env=dict(os.environ, KLINE_PROF_DATA_DIR=str(tmp_path / "nowhere")))
    assert run.returncode == 0 and run.stdout.split() == ["0", "0"]
    assert "kerrlineprofile: " in run.stderr and "MODEL OUTPUT SET TO ZERO" in run.stderr
    # the smoke script skips cleanly where HEASoft is absent
    sm = subprocess.run([os.path.join(root, "tools", "xspec_smoke.sh")], capture_output=True, text=True)
    assert sm.returncode in (0, 77), sm.stdout + sm.stderr


def test_hot_loop_fingerprint_is_the_measured_one():
    """The schedule ptxas picks for the quadrature's hot loop moves the headline by up to 2.4 % when unrelated kernel code
    changes (410.5 .. 420.1 ms per 50 000 sets for identical hot-loop source).  The fingerprint of the build that was measured
    is recorded in profiles/hot_loop_fingerprint.json; a different one means: run tools/ab_lib.py on a B200 and record again."""
    import json
    import shutil
    import subprocess
    import sys
    if shutil.which("cuobjdump") is None:
        pytest.skip("no cuobjdump")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    rec = json.load(open(os.path.join(root, "profiles", "hot_loop_fingerprint.json")))
    out = subprocess.run([sys.executable, os.path.join(root, "tools", "hot_loop_fingerprint.py")], capture_output=True, text=True).stdout
    assert json.loads(out) == rec["fingerprint"], "hot loop re-scheduled by ptxas: re-measure (tools/ab_lib.py) and run tools/hot_loop_fingerprint.py --write"


def test_one_correction_division_is_correctly_rounded(tmp_path):
    """div_by<T, true> (klp_kernels.cuh) divides by a value whose correctly rounded reciprocal is known with q0 = RN(a y),
    r = fma(-b, q0, a), q = fma(r, y, q0): three operations.  Its claim - q == RN(a / b) for every a - is checked on the CPU
    against the hardware division: every f32 mantissa of a against 48 divisors (12 just below 2 and 12 just above 1: the
    corner the proof in the kernel file leaves to enumeration), 1.4e6 f64 corner pairs, 8e7 random pairs.  (`full` mode of the
    tool: 5.6e9 checks, 0 mismatches; the GPU test test_exact_arithmetic_sequences repeats it against __fdiv_rn / __ddiv_rn.)"""
    import shutil
    import subprocess
    if shutil.which("gcc") is None:
        pytest.skip("no gcc")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    exe = tmp_path / "markstein_check"
    cc = subprocess.run(["gcc", "-O2", "-ffp-contract=off", os.path.join(root, "tools", "markstein_check.c"), "-lm", "-o", str(exe)],
                        capture_output=True, text=True)
    assert cc.returncode == 0, cc.stderr
    run = subprocess.run([str(exe), "quick"], capture_output=True, text=True)
    assert run.returncode == 0 and run.stdout.startswith("bad=0 "), run.stdout


def test_stage_chunk_plan():
    """Host-buffer pipeline: full pinned slots, then a short final chunk (its copy-out overlaps nothing); small batches are one chunk
    = one stream, one synchronisation (the XSPEC call pattern)."""
    from lineprofiles_b200.api import stage_chunk_plan as plan
    per_slot = (256 << 20) // 8000                       # 1000 bins: 33554 sets per 256 MB slot
    assert plan(100000, 1000) == [per_slot, per_slot, 100000 - 2 * per_slot - 4096, 4096]
    assert plan(100000, 1000, 128, 0) == [16777] * 5 + [100000 - 5 * 16777]
    assert plan(1, 1000) == [1] and plan(per_slot, 1000) == [per_slot]            # one chunk
    assert plan(per_slot + 4096, 1000) == [per_slot, 4096]                        # no tail when it would be the whole second chunk
    assert plan(per_slot + 4097, 1000) == [per_slot, 1, 4096]
    assert plan(700, 1000, 1, 4096) == [131] * 5 + [13, 32]                       # the tail is at most a quarter slot
    assert plan(700, 1000, 1, 7) == [131] * 5 + [38, 7]
    assert plan(100, 10 ** 6, 1) == [16] * 5 + [16, 4]                            # very wide grids: at least 16 sets per slot
    assert plan(10 ** 6, 4, 4096) == [65535] * 15 + [10 ** 6 - 15 * 65535 - 4096, 4096]   # at most 65535 sets per slot
    for B in (1, 17, 4097, 65536, 123457):
        for n in (1, 1000, 4096):
            p = plan(B, n)
            assert sum(p) == B and all(c > 0 for c in p) and max(p) == p[0]
