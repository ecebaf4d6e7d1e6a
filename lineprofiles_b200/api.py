"""ctypes binding of include/klineprofiles_b200.h.

The five module-level functions kline / kline5 / kconv / kconv5 / kconv10 call the exported XSPEC
symbols of the same names (reference src/xspec-wrapper.zig:425-530) with the reference's argument
meaning: `energy` are n+1 bin edges, `params` the model's parameter vector (xspec/lmodel.dat),
`flux` the n-bin output (additive models) or in/out spectrum (convolution models).
`Table.eval_batch` is the batched entry the reference does not have.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

MODELS = {"kline": 0, "kline5": 1, "kconv": 2, "kconv5": 3, "kconv10": 4}
NPARAMS = {"kline": 6, "kline5": 12, "kconv": 5, "kconv5": 11, "kconv10": 16}
PRECISIONS = {"f32": 0, "f64": 1}
N_FINE = 1999
N_CONV_BINS = 1998

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


class KlpError(RuntimeError):
    pass


def library_path() -> str:
    # KLP_LIBRARY: developer override to A/B-test differently compiled builds of the same library
    return os.environ.get("KLP_LIBRARY") or os.path.join(_HERE, "libklineprofiles_b200.so")


def load_library():
    """Load the CUDA shared library; fail loudly when it has not been built."""
    global _LIB
    if _LIB is not None:
        return _LIB
    path = library_path()
    if not os.path.exists(path):
        raise KlpError(f"{path} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                       "(there is no CPU fallback)")
    L = C.CDLL(path)
    dp, fp, zp, vp = C.POINTER(C.c_double), C.POINTER(C.c_float), C.POINTER(C.c_size_t), C.c_void_p
    L.klp_last_error.restype = C.c_char_p
    L.klp_model_nparams.argtypes = [C.c_int]
    L.klp_device_count.argtypes = [C.POINTER(C.c_int)]
    L.klp_table_create.argtypes = [fp, C.c_int, fp, C.c_int, C.c_int, C.c_int, fp, fp, fp, fp, fp, C.c_int, C.POINTER(vp)]
    L.klp_table_load.argtypes = [C.c_char_p, C.c_int, C.POINTER(vp)]
    L.klp_table_share.argtypes = [vp, C.POINTER(vp)]
    L.klp_table_destroy.argtypes = [vp]
    L.klp_table_destroy.restype = None
    L.klp_table_dims.argtypes = [vp, C.POINTER(C.c_int)]
    L.klp_table_parameters.argtypes = [vp, fp, fp]
    L.klp_table_read_frame.argtypes = [vp, C.c_size_t, fp]
    L.klp_table_index.argtypes = [vp, C.c_size_t, C.c_size_t]
    L.klp_table_index.restype = C.c_size_t
    L.klp_find_parameter_indices.argtypes = [vp, C.c_float, C.c_float, zp]
    L.klp_find_parameter_indices.restype = None
    L.klp_find_tables.argtypes = [vp, zp, zp]
    L.klp_find_tables.restype = None
    L.klp_parameter_weight.argtypes = [vp, C.c_int, C.c_float, C.c_size_t, fp]
    L.klp_eval_batch.argtypes = [vp, C.c_int, C.c_int, C.c_long, dp, dp, C.c_int, dp, C.c_int, dp]
    L.klp_eval_batch_device.argtypes = [vp, C.c_int, C.c_int, C.c_long, vp, vp, C.c_int, vp, C.c_int, vp, vp]
    L.klp_eval_debug.argtypes = [vp, C.c_int, C.c_int, C.c_long, dp, dp, C.c_int, dp, dp]
    L.klp_blend_frames_device.argtypes = [vp, C.c_int, C.c_long, vp, vp, vp]
    L.klp_set_timing.argtypes = [vp, C.c_int]
    L.klp_get_timing.argtypes = [vp, dp, C.POINTER(C.c_long)]
    L.klp_set_option.argtypes = [vp, C.c_char_p, C.c_long]
    L.klp_last_launch_info.argtypes = [vp, C.c_char_p, C.c_size_t]
    L.klp_get_phase_times.argtypes = [vp, dp]
    L.klp_fma_peak_gflops.argtypes = [vp, C.c_int, dp]
    L.klp_arith_selftest.argtypes = [vp, C.c_ulonglong, C.c_ulonglong, C.POINTER(C.c_ulonglong)]
    pvp = C.POINTER(vp)
    L.klp_nccl_unique_id.argtypes = [C.c_char_p]
    L.klp_nccl_version.argtypes = [C.POINTER(C.c_int)]
    L.klp_group_create_local.argtypes = [pvp, C.c_int, pvp]
    L.klp_group_create_rank.argtypes = [vp, C.c_char_p, C.c_int, C.c_int, pvp]
    L.klp_group_destroy.argtypes = [vp]
    L.klp_group_destroy.restype = None
    L.klp_group_size.argtypes = [vp, C.POINTER(C.c_int), C.POINTER(C.c_int)]
    L.klp_group_eval_batch.argtypes = [vp, C.c_int, C.c_int, C.c_long, dp, dp, C.c_int, dp, C.c_int, dp]
    L.klp_group_eval_allgather_device.argtypes = [vp, C.c_int, C.c_int, C.c_long, pvp, pvp, C.c_int, pvp, C.c_int, pvp, pvp]
    L.klp_group_synchronize.argtypes = [vp]
    L.klp_group_set_option.argtypes = [vp, C.c_char_p, C.c_long]
    L.klp_group_last_gather_ops.argtypes = [vp]
    L.klp_group_chunk_plan.argtypes = [C.c_long, C.c_long, C.c_long, C.POINTER(C.c_long), C.c_int]
    L.klp_stage_chunk_plan.argtypes = [C.c_long, C.c_int, C.c_long, C.c_long, C.POINTER(C.c_long), C.c_int]
    sig = [dp, C.c_int, dp, C.c_int, dp, dp, C.c_char_p]
    for name in MODELS:
        fn = getattr(L, name)
        fn.argtypes = sig
        fn.restype = None
    L.klp_set_default_table.argtypes = [vp]
    L.klp_set_default_table.restype = None
    L.klp_set_legacy_precision.argtypes = [C.c_int]
    L.klp_set_legacy_precision.restype = None
    L.klp_legacy_reset.restype = None
    _LIB = L
    return L


def _check(rc: int, what: str):
    if rc != 0:
        msg = load_library().klp_last_error().decode(errors="replace")
        raise KlpError(f"{what} failed (status {rc}): {msg}")


def _dp(a):
    return a.ctypes.data_as(C.POINTER(C.c_double)) if a is not None else None


def _fp(a):
    return a.ctypes.data_as(C.POINTER(C.c_float))


def _model_id(model) -> int:
    return MODELS[model] if isinstance(model, str) else int(model)


def _ptr(x) -> int:
    """Device pointer of a torch tensor / anything with data_ptr(), or a raw int."""
    if x is None:
        return 0
    if hasattr(x, "data_ptr"):
        return int(x.data_ptr())
    return int(x)


class Table:
    """A transfer-function table resident on one GPU (LineProfileTable, reference line-profile.zig:9-54)."""

    def __init__(self, handle, keep=None):
        self._h = handle
        self._keep = keep
        d = (C.c_int * 4)()
        _check(load_library().klp_table_dims(self._h, d), "klp_table_dims")
        self.n_a, self.n_mu, self.n_r, self.n_g = (int(v) for v in d)

    @classmethod
    def from_arrays(cls, t: dict, device: int = 0) -> "Table":
        L = load_library()
        k = {n: np.ascontiguousarray(t[n], np.float32) for n in ("spins", "mus", "radii", "gmin", "gmax", "upper", "lower")}
        h = C.c_void_p()
        _check(L.klp_table_create(_fp(k["spins"]), int(t["n_a"]), _fp(k["mus"]), int(t["n_mu"]), int(t["n_r"]), int(t["n_g"]),
                                  _fp(k["radii"]), _fp(k["gmin"]), _fp(k["gmax"]), _fp(k["upper"]), _fp(k["lower"]),
                                  int(device), C.byref(h)), "klp_table_create")
        return cls(h)

    @classmethod
    def load(cls, path: str, device: int = 0) -> "Table":
        h = C.c_void_p()
        _check(load_library().klp_table_load(os.fsencode(path), int(device), C.byref(h)), "klp_table_load")
        return cls(h)

    def share(self) -> "Table":
        """Another handle on the same device frames with its own workspaces (for concurrent evaluations)."""
        h = C.c_void_p()
        _check(load_library().klp_table_share(self._h, C.byref(h)), "klp_table_share")
        return Table(h)

    def close(self):
        if self._h:
            load_library().klp_table_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @property
    def handle(self):
        return self._h

    @property
    def frame_len(self) -> int:
        return self.n_r * (2 * self.n_g + 3)

    def parameters(self):
        """(spins[n_a], mus[n_mu]) as the table holds them."""
        a, m = np.empty(self.n_a, np.float32), np.empty(self.n_mu, np.float32)
        _check(load_library().klp_table_parameters(self._h, _fp(a), _fp(m)), "klp_table_parameters")
        return a, m

    def read_frame(self, frame: int) -> dict:
        """One frame as the table holds it: radii/gmin/gmax [n_r], upper/lower [n_r, n_g]."""
        buf = np.empty(self.frame_len, np.float32)
        _check(load_library().klp_table_read_frame(self._h, int(frame), _fp(buf)), "klp_table_read_frame")
        n_r, n_g = self.n_r, self.n_g
        return dict(radii=buf[:n_r].copy(), gmin=buf[n_r:2 * n_r].copy(), gmax=buf[2 * n_r:3 * n_r].copy(),
                    upper=buf[3 * n_r:3 * n_r + n_r * n_g].reshape(n_r, n_g).copy(),
                    lower=buf[3 * n_r + n_r * n_g:].reshape(n_r, n_g).copy())

    # ---- index logic (line-profile.zig:60-93,196-231)
    def table_index(self, i_a: int, i_mu: int) -> int:
        return int(load_library().klp_table_index(self._h, i_a, i_mu))

    def find_parameter_indices(self, a: float, mu: float):
        o = (C.c_size_t * 2)()
        load_library().klp_find_parameter_indices(self._h, a, mu, o)
        return list(o)

    def find_tables(self, param_indices):
        a = (C.c_size_t * 2)(*param_indices)
        o = (C.c_size_t * 4)()
        load_library().klp_find_tables(self._h, a, o)
        return list(o)

    def parameter_weight(self, which: int, value: float, index: int):
        w = C.c_float()
        has = load_library().klp_parameter_weight(self._h, which, value, index, C.byref(w))
        return w.value if has else None

    # ---- evaluation
    def eval_batch(self, model, params, energy, spectrum=None, precision: str = "f32", out=None) -> np.ndarray:
        """Evaluate B parameter sets; host numpy arrays in, [B, n] f64 out."""
        m = _model_id(model)
        params = np.ascontiguousarray(params, np.float64)
        if params.ndim == 1:
            params = params[None, :]
        B, P = params.shape
        if P != load_library().klp_model_nparams(m):
            raise KlpError(f"model {model} takes {load_library().klp_model_nparams(m)} parameters, got {P}")
        energy = np.ascontiguousarray(energy, np.float64)
        n = energy.size - 1
        per_set = 0
        if spectrum is not None:
            spectrum = np.ascontiguousarray(spectrum, np.float64)
            per_set = int(spectrum.ndim == 2)
            if spectrum.shape[-1] != n or (per_set and spectrum.shape[0] != B):
                raise KlpError("spectrum shape does not match the energy grid / batch")
        if out is None:
            out = np.empty((B, n), np.float64)
        _check(load_library().klp_eval_batch(self._h, m, PRECISIONS[precision], B, _dp(params), _dp(energy), n,
                                             _dp(spectrum), per_set, _dp(out)), "klp_eval_batch")
        return out

    def eval_batch_device(self, model, d_params, d_energy, n_flux: int, d_out, B: int, d_spectrum=None,
                          spectrum_per_set: bool = False, precision: str = "f32", stream=None) -> None:
        """Device-pointer entry (torch tensors or raw pointers); asynchronous on `stream`."""
        _check(load_library().klp_eval_batch_device(self._h, _model_id(model), PRECISIONS[precision], int(B), _ptr(d_params),
                                                    _ptr(d_energy), int(n_flux), _ptr(d_spectrum), int(spectrum_per_set),
                                                    _ptr(d_out), _ptr(stream)), "klp_eval_batch_device")

    def eval_debug(self, model, params, energy, precision: str = "f32"):
        m = _model_id(model)
        params = np.ascontiguousarray(params, np.float64)
        if params.ndim == 1:
            params = params[None, :]
        B = params.shape[0]
        energy = np.ascontiguousarray(energy, np.float64)
        fine = np.empty((B, N_FINE), np.float64)
        prof = np.empty((B, N_CONV_BINS), np.float64) if m >= 2 else None
        _check(load_library().klp_eval_debug(self._h, m, PRECISIONS[precision], B, _dp(params), _dp(energy),
                                             energy.size - 1, _dp(fine), _dp(prof)), "klp_eval_debug")
        return fine, prof

    def blend_frames_device(self, d_a_incl, d_out, B: int, precision: str = "f32", stream=None) -> None:
        _check(load_library().klp_blend_frames_device(self._h, PRECISIONS[precision], int(B), _ptr(d_a_incl), _ptr(d_out),
                                                      _ptr(stream)), "klp_blend_frames_device")

    # ---- instrumentation / tuning
    def set_timing(self, enable: bool) -> None:
        _check(load_library().klp_set_timing(self._h, int(enable)), "klp_set_timing")

    def get_timing(self):
        ms = (C.c_double * 4)()
        n = (C.c_long * 4)()
        _check(load_library().klp_get_timing(self._h, ms, n), "klp_get_timing")
        names = ("setup", "quad", "rebin", "conv")
        return {k: {"ms": float(ms[i]), "launches": int(n[i])} for i, k in enumerate(names)}

    def set_option(self, key: str, value: int) -> None:
        _check(load_library().klp_set_option(self._h, key.encode(), int(value)), "klp_set_option")

    def last_launch_info(self) -> str:
        """Instantiation and mode of the most recent quadrature launch, e.g. 'k_quad<float, 1024, 1> splits=1 ring=0 ...'."""
        buf = C.create_string_buffer(160)
        _check(load_library().klp_last_launch_info(self._h, buf, len(buf)), "klp_last_launch_info")
        return buf.value.decode()

    def phase_times(self):
        """Microseconds at which k_quad's first CTA passed its phase boundaries (option phase_clock = 1)."""
        us = (C.c_double * 10)()
        _check(load_library().klp_get_phase_times(self._h, us), "klp_get_phase_times")
        return dict(zip(("start", "unpack", "grid_blend", "plan", "bin_ranges", "stage_a", "all_arrived", "additions", "rebin", "rebin_start"), (float(v) for v in us)))

    def arith_selftest(self, n_total: int, seed: int = 1):
        m = (C.c_ulonglong * 4)()
        _check(load_library().klp_arith_selftest(self._h, seed, int(n_total), m), "klp_arith_selftest")
        return dict(zip(("f32_div_by", "f64_div_by", "f32_div_gen", "f32_sqrt_gen"), (int(v) for v in m)))

    def fma_peak_gflops(self, precision: str = "f32") -> float:
        g = C.c_double()
        _check(load_library().klp_fma_peak_gflops(self._h, PRECISIONS[precision], C.byref(g)), "klp_fma_peak_gflops")
        return g.value


# ---- several GPUs (include/klineprofiles_b200.h, klp_group_*) ---------------------------------------

def nccl_unique_id() -> bytes:
    """128-byte NCCL id (rank 0 creates it, the host distributes it to the other ranks)."""
    buf = C.create_string_buffer(128)
    _check(load_library().klp_nccl_unique_id(buf), "klp_nccl_unique_id")
    return buf.raw


def nccl_version() -> int:
    v = C.c_int()
    _check(load_library().klp_nccl_version(C.byref(v)), "klp_nccl_version")
    return v.value


def group_chunk_plan(B_local: int, gather_chunk: int = 131072, gather_tail: int = 2048):
    """Chunk sizes of the device-resident multi-GPU entry for B_local sets per rank (few big chunks + a short tail)."""
    buf = (C.c_long * 4096)()
    n = load_library().klp_group_chunk_plan(int(B_local), int(gather_chunk), int(gather_tail), buf, 4096)
    if n < 0:
        raise KlpError("klp_group_chunk_plan: bad arguments")
    return [int(buf[i]) for i in range(n)]


def stage_chunk_plan(B: int, n_flux: int, stage_mb: int = 256, stage_tail: int = 4096):
    """Staging chunks of the host-buffer pipeline (klp_eval_batch): full pinned slots, then a short final chunk."""
    buf = (C.c_long * 8192)()
    n = load_library().klp_stage_chunk_plan(int(B), int(n_flux), int(stage_mb), int(stage_tail), buf, 8192)
    if n < 0:
        raise KlpError("klp_stage_chunk_plan: bad arguments")
    return [int(buf[i]) for i in range(n)]


def _vp_array(items):
    return (C.c_void_p * len(items))(*[_ptr(x) for x in items])


class Group:
    """The GPUs that share a batch: table replicas, one per device (klp_group)."""

    def __init__(self, handle, tables):
        self._h = handle
        self._tables = list(tables)   # borrowed by the C group: keep them alive
        self.rank = None              # rank of this process in a one-process-per-GPU group
        nl, nr = C.c_int(), C.c_int()
        _check(load_library().klp_group_size(self._h, C.byref(nl), C.byref(nr)), "klp_group_size")
        self.n_local, self.nranks = nl.value, nr.value

    @classmethod
    def local(cls, tables) -> "Group":
        """One process drives len(tables) GPUs."""
        arr = (C.c_void_p * len(tables))(*[t.handle for t in tables])
        h = C.c_void_p()
        _check(load_library().klp_group_create_local(arr, len(tables), C.byref(h)), "klp_group_create_local")
        return cls(h, tables)

    @classmethod
    def from_rank(cls, table: Table, unique_id: bytes, rank: int, nranks: int) -> "Group":
        """One process per GPU: this process is `rank` of `nranks`."""
        h = C.c_void_p()
        _check(load_library().klp_group_create_rank(table.handle, unique_id, int(rank), int(nranks), C.byref(h)),
               "klp_group_create_rank")
        g = cls(h, [table])
        g.rank = int(rank)
        return g

    def close(self):
        if self._h:
            load_library().klp_group_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_option(self, key: str, value: int) -> None:
        _check(load_library().klp_group_set_option(self._h, key.encode(), int(value)), "klp_group_set_option")

    def synchronize(self) -> None:
        _check(load_library().klp_group_synchronize(self._h), "klp_group_synchronize")

    def last_gather_ops(self) -> int:
        return int(load_library().klp_group_last_gather_ops(self._h))

    def eval_batch(self, model, params, energy, spectrum=None, precision: str = "f32", out=None) -> np.ndarray:
        """Host numpy arrays in, [B, n] f64 out; the sets are sharded over the GPUs of a local group."""
        m = _model_id(model)
        params = np.ascontiguousarray(params, np.float64)
        B = params.shape[0]
        energy = np.ascontiguousarray(energy, np.float64)
        n = energy.size - 1
        per_set = 0
        if spectrum is not None:
            spectrum = np.ascontiguousarray(spectrum, np.float64)
            per_set = int(spectrum.ndim == 2)
        if out is None:
            out = np.empty((B, n), np.float64)
        _check(load_library().klp_group_eval_batch(self._h, m, PRECISIONS[precision], B, _dp(params), _dp(energy), n,
                                                   _dp(spectrum), per_set, _dp(out)), "klp_group_eval_batch")
        return out

    def eval_allgather_device(self, model, d_params, d_energy, n_flux: int, d_out_all, B_local: int, d_spectrum=None,
                              spectrum_per_set: bool = False, precision: str = "f32", streams=None) -> None:
        """Device-resident shards in, every GPU holds all nranks * B_local spectra afterwards ([rank][set][bin]).
        d_params / d_energy / d_out_all / d_spectrum / streams: one entry per local member."""
        sp = _vp_array(d_spectrum) if d_spectrum is not None else None
        st = _vp_array(streams) if streams is not None else None
        _check(load_library().klp_group_eval_allgather_device(
            self._h, _model_id(model), PRECISIONS[precision], int(B_local), _vp_array(d_params), _vp_array(d_energy), int(n_flux),
            sp, int(spectrum_per_set), _vp_array(d_out_all), st), "klp_group_eval_allgather_device")


# ---- the reference's model surface ---------------------------------------------------------------

_INSTALLED = None   # the Table installed for the XSPEC symbols is kept alive here (the library does not own it)


def set_default_table(table: Table | None) -> None:
    global _INSTALLED
    load_library().klp_set_default_table(table.handle if table is not None else None)
    _INSTALLED = table


def set_legacy_precision(precision: str) -> None:
    load_library().klp_set_legacy_precision(PRECISIONS[precision])


def legacy_reset() -> None:
    load_library().klp_legacy_reset()


def _legacy(name: str, energy, params, flux):
    L = load_library()
    energy = np.ascontiguousarray(energy, np.float64)
    params = np.ascontiguousarray(params, np.float64)
    n = energy.size - 1
    if params.size != NPARAMS[name]:
        raise KlpError(f"{name} takes {NPARAMS[name]} parameters")
    if flux is None:
        if MODELS[name] >= 2:
            raise KlpError(f"{name} convolves the spectrum passed in `flux`")
        flux = np.zeros(n, np.float64)
    if not (isinstance(flux, np.ndarray) and flux.dtype == np.float64 and flux.flags.c_contiguous and flux.size == n):
        raise KlpError("flux must be a contiguous float64 array with one entry per energy bin")
    fvar = np.zeros(1, np.float64)
    getattr(L, name)(_dp(energy), n, _dp(params), 0, _dp(flux), _dp(fvar), b"")
    return flux


def kline(energy, params, flux=None):
    return _legacy("kline", energy, params, flux)


def kline5(energy, params, flux=None):
    return _legacy("kline5", energy, params, flux)


def kconv(energy, params, flux):
    return _legacy("kconv", energy, params, flux)


def kconv5(energy, params, flux):
    return _legacy("kconv5", energy, params, flux)


def kconv10(energy, params, flux):
    return _legacy("kconv10", energy, params, flux)
