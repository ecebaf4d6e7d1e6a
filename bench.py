#!/usr/bin/env python
"""bench.py -- kline-family spectra/s on B200 (BASELINE.json metric), with roofline and CPU baseline.

  python bench.py --gpus N --steps K --warmup W            our arm (CUDA, one process per GPU)
  python bench.py --impl reference --gpus N --steps K ...  the reference algorithm on the host cores

Workload (BASELINE.json configs[1]): kline5, a batch of 1e5 random parameter points per GPU and step
(5 emissivity knots, rcut) on the 1000-bin 0.1-10 keV grid, synthetic Kerr-shaped 30x20x150x30 table.
A step = one pass of the hot path over one batch.  `value` times the device-resident path (parameters already in HBM;
for N > 1 through the C ABI's klp_group entry: chunked evaluation, the NCCL gather of one chunk overlapping the kernels
of the next); `e2e` times Table.eval_batch with host buffers (params H2D from pinned staging, every spectrum D2H).
N > 1 runs either as one process per GPU (torchrun: the driver's launch) or -- `python bench.py --gpus N` without
torchrun -- as ONE process driving N GPUs through klp_group_create_local.  At N = 1 the line also carries a bounded
`configs` block with the other BASELINE configs (1: single-call latency, 3: kconv on 4096 log bins, 5: multi-GB table).
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "kline spectra/sec (1k-bin grid)"
UNIT = "spectra/s"
N_BINS = 1000
MODEL = "kline5"


def host_cores() -> int:
    try:
        return len(os.sched_getaffinity(0))
    except AttributeError:
        return os.cpu_count() or 1


def measured_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled during the timed region."""

    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index = index
        self.rows = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self) -> dict:
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap")
        for r in self.rows:
            try:
                sm.append(float(r[0]))
                mx.append(float(r[1]))
            except Exception:
                continue
            for n, v in zip(names, r[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def run_reference(args):
    """The reference's CPU algorithm (C++ restatement, oracle/; the Zig toolchain is unavailable), all host threads."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from lineprofiles_b200 import synthetic as syn
    from oracle import oracle_py as orc
    orc.build()
    cores = host_cores()
    ot = orc.OracleTable(syn.make_table())
    energy = syn.energy_grid_linear(N_BINS)
    # bounded sample per step: probe the per-set cost, then size a step to ~4 s of wall time
    probe = syn.random_params(MODEL, cores, seed=999)
    t0 = time.perf_counter()
    ot.eval_batch(MODEL, probe, energy, precision="f32", threads=cores, faithful_cost=True)
    per_round = max(time.perf_counter() - t0, 1e-3)
    n = int(max(cores, min(4.0 / per_round, 64) * cores))
    times = []
    for s in range(args.warmup + args.steps):
        p = syn.random_params(MODEL, n, seed=5000 + s)
        t0 = time.perf_counter()
        ot.eval_batch(MODEL, p, energy, precision="f32", threads=cores, faithful_cost=True)
        dt = time.perf_counter() - t0
        if s >= args.warmup:
            times.append(dt)
    total = sum(times)
    value = n * len(times) / total
    sample = f"{n} random kline5 sets per step (same marginals/grid/table as the GPU arm), {len(times)} steps"
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1e3 * total / len(times), "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": "kline5 batch of random parameter points, 1000-bin 0.1-10 keV grid, synthetic Kerr-shaped "
                                   "table 30x20x150x30 (BASELINE.json configs[1])", "sets_per_step": n},
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port",
                             "sample": sample + "; C++ restatement of the reference algorithm in its own f32 arithmetic and "
                                                "cost structure (Zig toolchain unavailable), g++ -O3, one set per thread"},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


def cpu_baseline(syn, budget_s: float = 12.0):
    from oracle import oracle_py as orc
    orc.build()
    cores = host_cores()
    ot = orc.OracleTable(syn.make_table())
    energy = syn.energy_grid_linear(N_BINS)
    probe = syn.random_params(MODEL, cores, seed=998)
    t0 = time.perf_counter()
    ot.eval_batch(MODEL, probe, energy, precision="f32", threads=cores, faithful_cost=True)
    per_round = max(time.perf_counter() - t0, 1e-3)
    n = int(max(cores, min(budget_s / per_round, 256) * cores))
    p = syn.random_params(MODEL, n, seed=997)
    t0 = time.perf_counter()
    r = ot.eval_batch(MODEL, p, energy, precision="f32", threads=cores, faithful_cost=True, want_stats=True)
    dt = time.perf_counter() - t0
    evals_per_set = r["stats"]["integrand_evals"] / n
    return {"value": n / dt, "unit": UNIT, "cores": cores, "kind": "port",
            "sample": f"{n} random kline5 sets of the same workload in {dt:.1f} s; C++ restatement of the reference algorithm "
                      "(f32, reference cost structure; Zig toolchain unavailable), g++ -O3, one set per thread"}, evals_per_set



def cpu_baseline_conv(syn, params, energy, spec, n_sets: int = 4):
    """CPU-baseline leg of config 3: the restated reference convolution on a few of the same kconv sets, all host threads
    (reference cost structure); also returns the in-window (source, output) pairs per set that the k_conv roofline uses."""
    from oracle import oracle_py as orc
    orc.build()
    cores = host_cores()
    ot = orc.OracleTable(syn.make_table())
    n = max(n_sets, min(cores, 16))
    t0 = time.perf_counter()
    st = ot.eval_batch("kconv", params[:n], energy, spectrum=spec, precision="f32", threads=cores, faithful_cost=True, want_stats=True)["stats"]
    dt = time.perf_counter() - t0
    return {"value": n / dt, "unit": UNIT, "cores": cores, "kind": "port", "sample": f"{n} kconv sets of the same workload in {dt:.1f} s"}, st["conv_pairs"] / n


def recorded_fingerprint():
    """Hot-loop fingerprint of the build whose measurements are committed (tests/test_abi.py checks that the built library has it)."""
    p = os.path.join(os.path.dirname(os.path.abspath(__file__)), "profiles", "hot_loop_fingerprint.json")
    return json.load(open(p)).get("fingerprint") if os.path.exists(p) else None


def latest_profile(name_glob: str):
    import glob
    c = sorted(glob.glob(os.path.join(ROOT, "profiles", name_glob)))
    return c[-1] if c else None


def kernel_matches(capture_kernel: str, launch_info: str) -> bool:
    """'void k_quad<float, 1024, 1>(QuadArgs<T1>)' (ncu) vs 'k_quad<float, 1024, 1> splits=1 ring=0 ...' (klp_last_launch_info)."""
    inst = launch_info.split(" splits=")[0].replace(" ", "")
    return inst != "" and inst in capture_kernel.replace(" ", "")


def other_configs(klp, syn, gt, dev, peak_hbm, with_cpu=True, budget_s=25.0):
    """The other BASELINE configs, bounded (~25 s): 1 single-call latency through the exported `kline`; 3 kconv on the
    4096-bin log grid (both convolution modes, k_conv roofline against a DFMA microbenchmark); 5 a multi-GB table:
    fused spectra/s and the stand-alone k_blend GB/s against the measured HBM peak."""
    import torch
    t_start = time.perf_counter()
    res = {}
    st = torch.cuda.current_stream().cuda_stream

    def device_rate(table, model, dp, dE, n, B, dspec=None, reps=3):
        out = torch.empty((B, n), dtype=torch.float64, device=dev)
        table.eval_batch_device(model, dp, dE, n, out, B, d_spectrum=dspec, stream=st)
        torch.cuda.synchronize()
        table.set_timing(True)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            table.eval_batch_device(model, dp, dE, n, out, B, d_spectrum=dspec, stream=st)
        e1.record()
        torch.cuda.synchronize()
        tm = table.get_timing()
        table.set_timing(False)
        ms = e0.elapsed_time(e1) / reps
        return B / (ms * 1e-3), ms, {k: v["ms"] / reps for k, v in tm.items()}

    # ---- config 1: one kline evaluation per call through the exported XSPEC symbol (host buffers, the XSPEC call pattern)
    try:
        E1 = syn.energy_grid_linear(1000)
        klp.set_default_table(gt)
        klp.legacy_reset()
        flux = np.zeros(1000)
        for _ in range(20):
            klp.kline(E1, syn.CONFIG1_PARAMS[0], flux)
        n_calls = 300
        t0 = time.perf_counter()
        for _ in range(n_calls):
            klp.kline(E1, syn.CONFIG1_PARAMS[0], flux)
        dt = (time.perf_counter() - t0) / n_calls
        res["config1_kline_single_call"] = {"ms_per_call": dt * 1e3, "calls_per_s": 1.0 / dt, "n_calls": n_calls,
                                            "api": "exported C symbol kline() (XSPEC signature), host buffers, 1000-bin 0.1-10 keV grid"}
    except Exception as e:   # a config must not take the headline down
        res["config1_kline_single_call"] = {"error": repr(e)}
    finally:
        klp.set_default_table(None)
        klp.legacy_reset()

    # ---- config 3: kconv, 4096 log bins, shared E^-2 continuum
    try:
        E3 = syn.energy_grid_log(4096)
        spec = syn.powerlaw_continuum(E3)
        B3 = 2000
        p3 = syn.random_params("kconv", B3, seed=20261)
        dp, dE, dS = (torch.from_numpy(x).to(dev) for x in (p3, E3, spec))
        r3 = {"sets": B3, "grid": "4096 log bins 0.1-10 keV", "continuum": "shared E^-2"}
        pk64 = gt.fma_peak_gflops("f64")
        # in-window (source, output) pairs counted by the CPU restatement on a few of the same sets
        pairs = None
        if with_cpu:
            r3["cpu_baseline"], pairs = cpu_baseline_conv(syn, p3, E3, spec)
        for mode, name in ((0, "reference_order"), (1, "cumulative_profile")):
            gt.set_option("conv_mode", mode)
            try:
                rate, ms, tm = device_rate(gt, "kconv", dp, dE, 4096, B3, dS)
            finally:
                gt.set_option("conv_mode", 0)
            ach = 8.0 * pairs * B3 / (tm["conv"] * 1e-3) / 1e9 if (pairs and tm["conv"] > 0) else None
            r3[name] = {"spectra_per_s": rate, "ms_per_batch": ms, "k_quad_ms": tm["quad"], "k_conv_ms": tm["conv"], "k_rebin_ms": tm["rebin"],
                        "roofline": {"bound": "fp64_cuda_core", "kernel": "k_conv_cum<2, true>" if mode else "k_conv_exact", "achieved": ach, "peak": pk64,
                                     "unit": "GFLOP/s", "frac": ach / pk64 if ach and pk64 else None,
                                     "flops": "8 x in-window (source, output) pairs (SURVEY.md 8d A_F,conv); pairs/set counted by the oracle: %s" % ("%.4g" % pairs if pairs else "n/a"),
                                     "peak_source": "klp_fma_peak_gflops f64 (DFMA microbenchmark on this GPU)"}}
        res["config3_kconv_4096"] = r3
    except Exception as e:
        res["config3_kconv_4096"] = {"error": repr(e)}

    # ---- config 5: a multi-GB table resident in HBM
    try:
        if time.perf_counter() - t_start < budget_s * 0.5:
            dims = (120, 80, 600, 60)
            big = syn.make_table(*dims)
            nbytes = sum(big[k].nbytes for k in ("radii", "gmin", "gmax", "upper", "lower"))
            gb = klp.Table.from_arrays(big, device=dev.index)
            try:
                E5 = syn.energy_grid_linear(1000)
                B5 = 20000
                p5 = syn.random_params("kline", B5, seed=20259)
                dp, dE = torch.from_numpy(p5).to(dev), torch.from_numpy(E5).to(dev)
                rate, ms, tm = device_rate(gb, "kline", dp, dE, 1000, B5, reps=2)
                # the stand-alone blend kernel: 4 frame reads + 1 frame write per set, nothing else -- the HBM-bound stage
                Bb = 4096
                flen = dims[2] * (2 * dims[3] + 3)
                ai = torch.from_numpy(np.ascontiguousarray(syn.random_params("kline", Bb, seed=7)[:, :2])).to(dev)
                outb = torch.empty((Bb, flen), dtype=torch.float32, device=dev)
                gb.blend_frames_device(ai, outb, Bb, stream=st)
                torch.cuda.synchronize()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                best = None
                for _ in range(3):
                    e0.record()
                    gb.blend_frames_device(ai, outb, Bb, stream=st)
                    e1.record()
                    torch.cuda.synchronize()
                    t_ms = e0.elapsed_time(e1)
                    best = t_ms if best is None else min(best, t_ms)
                gbs = 5.0 * flen * 4 * Bb / (best * 1e-3) / 1e9
                res["config5_table_%dx%dx%dx%d" % dims] = {
                    "table_bytes": nbytes, "fused_kline_spectra_per_s": rate, "sets": B5, "k_quad_ms": tm["quad"],
                    "k_blend": {"bound": "hbm", "achieved": gbs, "peak": peak_hbm, "unit": "GB/s", "frac": gbs / peak_hbm, "ms": best,
                                "sets": Bb, "bytes": "4 frame reads + 1 frame write per set (f32), output %.2f GB > L2" % (flen * 4 * Bb / 1e9)}}
            finally:
                gb.close()
        else:
            res["config5"] = {"skipped": "time budget of the configs block spent"}
    except Exception as e:
        res["config5"] = {"error": repr(e)}
    # ---- a table with zero-filled rows (the reference zero-fills radii a frame does not cover): exact zeros stay inside the
    #      branch-free arithmetic (checked per set on the blended frame)
    try:
        tz = syn.make_table()
        n_mu = tz["n_mu"]
        holes = [ia * n_mu + im for ia in (3, 11, 17, 24) for im in (2, 5, 11, 16)]
        for f in holes:
            tz["upper"][f, -6:] = 0.0     # the innermost radii of the frame: the ones the quadrature actually visits
            tz["lower"][f, -6:] = 0.0
        gz = klp.Table.from_arrays(tz, device=dev.index)
        try:
            Ez = syn.energy_grid_linear(1000)
            Bz = 20000
            dp, dE = torch.from_numpy(syn.random_params("kline5", Bz, seed=31)).to(dev), torch.from_numpy(Ez).to(dev)
            rate_z, _, _ = device_rate(gz, "kline5", dp, dE, 1000, Bz, reps=2)
            rate_0, _, _ = device_rate(gt, "kline5", dp, dE, 1000, Bz, reps=2)
            res["table_with_zero_rows"] = {"frames_with_zero_rows": len(holes), "kline5_spectra_per_s": rate_z, "same_batch_on_the_clean_table": rate_0,
                                           "launch": gz.last_launch_info()}
        finally:
            gz.close()
    except Exception as e:
        res["table_with_zero_rows"] = {"error": repr(e)}
    res["seconds"] = time.perf_counter() - t_start
    return res


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--precision", default="f32", choices=["f32", "f64"])
    ap.add_argument("--batch", type=int, default=100000, help="parameter sets per GPU per step")
    ap.add_argument("--gather-chunk", type=int, default=0, help="sets per compute/gather chunk for N > 1 (0 = library default)")
    ap.add_argument("--gather-tail", type=int, default=-1, help="sets of the final, un-overlapped chunk for N > 1 (-1 = library default)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-configs", action="store_true", help="skip the bounded block with BASELINE configs 1, 3 and 5")
    ap.add_argument("--option", action="append", default=[], help="table option key=value (klp_set_option), e.g. deposit_ring=1")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    if args.impl == "reference":
        run_reference(args)
        return

    import torch
    import torch.distributed as dist

    import lineprofiles_b200 as klp
    from lineprofiles_b200 import dist as kdist
    from lineprofiles_b200 import synthetic as syn

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- this framework has no CPU path")
    # one process per GPU (torchrun) or ONE process driving --gpus N devices through the C ABI's local group
    n_local = args.gpus if (world == 1 and args.gpus > 1) else 1
    if n_local > torch.cuda.device_count():
        raise SystemExit(f"bench.py: --gpus {n_local} but only {torch.cuda.device_count()} visible")
    n_total = world * n_local
    local_devs = [local_rank] if n_local == 1 else list(range(n_local))
    torch.cuda.set_device(local_devs[0])
    dev = torch.device("cuda", local_devs[0])
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    B, K, W, prec = args.batch, args.steps, args.warmup, args.precision
    tab = syn.make_table()
    tables = [klp.Table.from_arrays(tab, device=d) for d in local_devs]
    gt = tables[0]
    for t in tables:
        for kv in args.option:
            k, v = kv.split("=")
            t.set_option(k, int(v))
    group = None
    if world > 1:
        group = kdist.rank_group(gt)
    elif n_local > 1:
        group = klp.Group.local(tables)
    if group is not None and args.gather_chunk > 0:
        group.set_option("gather_chunk", args.gather_chunk)
    if group is not None and args.gather_tail >= 0:
        group.set_option("gather_tail", args.gather_tail)
    energy = syn.energy_grid_linear(N_BINS)
    P = klp.NPARAMS[MODEL]
    # a different parameter batch every step and GPU (nothing can be reused from a previous step)
    def seed_of(s, member):
        return 1000 * (rank * n_local + member + 1) + s
    host_params = [[syn.random_params(MODEL, B, seed=seed_of(s, m)) for s in range(W + K)] for m in range(n_local)]
    devs = [torch.device("cuda", d) for d in local_devs]
    d_params = [[torch.from_numpy(p).to(devs[m]) for p in host_params[m]] for m in range(n_local)]
    d_energy = [torch.from_numpy(energy).to(d) for d in devs]
    rows = n_total * B if group is not None else B
    d_all = [torch.empty((rows, N_BINS), dtype=torch.float64, device=d) for d in devs]
    flush = [torch.empty(256 << 20, dtype=torch.uint8, device=d) for d in devs]  # > 126 MB L2
    streams = []
    for d in devs:
        with torch.cuda.device(d):
            streams.append(torch.cuda.current_stream(d).cuda_stream)

    def step(i):
        for m, d in enumerate(devs):
            with torch.cuda.device(d):
                flush[m].zero_()
        if group is None:
            gt.eval_batch_device(MODEL, d_params[0][i], d_energy[0], N_BINS, d_all[0], B, precision=prec, stream=streams[0])
        else:
            # C ABI: chunked evaluation; the NCCL gather of chunk k overlaps the kernels of chunk k+1
            group.eval_allgather_device(MODEL, [d_params[m][i] for m in range(n_local)], d_energy, N_BINS, d_all, B,
                                        precision=prec, streams=streams)

    def barrier():
        if world > 1:
            dist.barrier()
        for d in devs:
            torch.cuda.synchronize(d)

    for i in range(W):
        step(i)
    barrier()
    sampler = ClockSampler(local_devs[0])
    if rank == 0:
        sampler.start()
    for t in tables:
        t.set_timing(True)
    evs = []
    for d in devs:
        with torch.cuda.device(d):
            evs.append((torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)))
    barrier()
    for m, d in enumerate(devs):
        with torch.cuda.device(d):
            evs[m][0].record()
    for i in range(W, W + K):
        step(i)
    for m, d in enumerate(devs):
        with torch.cuda.device(d):
            evs[m][1].record()
    barrier()
    elapsed_ms = max(e0.elapsed_time(e1) for e0, e1 in evs)
    # per-kernel device times accumulated over the K timed steps (events recorded by the library on the launching stream)
    tm = gt.get_timing()
    for t in tables:
        t.set_timing(False)
    clocks = sampler.stop() if rank == 0 else None
    quad_ms, quad_launches = tm["quad"]["ms"], tm["quad"]["launches"]
    launches = sum(v["launches"] for v in tm.values()) * n_local
    if world > 1:
        t = torch.tensor([elapsed_ms], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        elapsed_ms = float(t.item())
    value = n_total * B * K / (elapsed_ms * 1e-3)
    launch_info = gt.last_launch_info()
    gather_ops = group.last_gather_ops() if group is not None else 0

    # ---- e2e through the public host API: numpy in, numpy out, copies inside the timed region, all K steps
    e2e = None
    if not args.no_e2e:
        if n_local > 1:
            Bh = n_local * B
            hp = [np.concatenate([host_params[m][s] for m in range(n_local)], 0) for s in range(W + K)]
            call = lambda p, o: group.eval_batch(MODEL, p, energy, precision=prec, out=o)
            api = "lineprofiles_b200.Group.eval_batch -> klp_group_eval_batch (one host thread + pinned pipeline per GPU, no collective)"
        else:
            Bh, hp = B, host_params[0]
            call = lambda p, o: gt.eval_batch(MODEL, p, energy, precision=prec, out=o)
            api = "lineprofiles_b200.Table.eval_batch -> klp_eval_batch (host buffers, pinned staging, 3 streams)"
        out_h = np.zeros((Bh, N_BINS), np.float64)   # touched: the timed region must not pay the first-touch page faults
        out_h.fill(0.0)
        call(hp[0], out_h)  # untimed warm-up step: sizes the pinned staging buffers
        barrier()
        t0 = time.perf_counter()
        for i in range(K):
            call(hp[W + i], out_h)
        barrier()
        dt = time.perf_counter() - t0
        if world > 1:
            t = torch.tensor([dt], dtype=torch.float64, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            dt = float(t.item())
        e2e = {"value": world * Bh * K / dt, "unit": UNIT, "h2d_bytes_per_step": Bh * P * 8 + (N_BINS + 1) * 8,
               "d2h_bytes_per_step": Bh * N_BINS * 8, "steps": K, "api": api}

    if rank == 0:
        peak, peak_src = measured_peak()
        n_r, n_g = tab["n_r"], tab["n_g"]
        bytes_per_set = 8 * P + 4 * 4 * n_r * (2 * n_g + 3) + 8 * N_BINS  # SURVEY.md §8d A_B
        sets_per_launch = B * K / max(quad_launches, 1)
        avg_launch_ms = quad_ms / max(quad_launches, 1)
        achieved = bytes_per_set * sets_per_launch / (avg_launch_ms * 1e-3) / 1e9 if avg_launch_ms > 0 else None
        # per-spectrum figures of the committed ncu --set full capture; quoted only when the capture is of the very
        # kernel instantiation this run launched
        tj, traffic, stale = None, None, None
        tp = latest_profile("r*_quad_traffic.json")
        if tp and prec == "f32":
            tj = json.load(open(tp))
            fp_now = recorded_fingerprint()
            if not (kernel_matches(tj.get("kernel", ""), launch_info) and int(tj.get("ring", 0)) == int("ring=1" in launch_info)):
                stale = f"{os.path.basename(tp)} is a capture of '{tj.get('kernel')}' but this run launched '{launch_info}'"
                tj = None
            elif tj.get("hot_loop_fingerprint") != fp_now:
                # same instantiation, different hot loop (e.g. captured before an arithmetic change): instruction counts differ
                stale = (f"{os.path.basename(tp)} was captured from a build whose hot loop is {tj.get('hot_loop_fingerprint')}, "
                         f"the measured build's is {fp_now} (profiles/hot_loop_fingerprint.json)")
                tj = None
            else:
                traffic = (tj["dram_bytes_read"] + tj["dram_bytes_write"]) * sets_per_launch / tj["sets_per_launch"]
        roofline = {"bound": "hbm", "kernel": launch_info, "achieved": achieved, "peak": peak, "unit": "GB/s",
                    "frac": (achieved / peak) if achieved else None, "traffic": traffic, "peak_source": peak_src,
                    "algorithmic_bytes_per_spectrum": bytes_per_set, "avg_launch_ms": avg_launch_ms,
                    "sets_per_launch": sets_per_launch, "launches_timed": quad_launches,
                    "note": "the quadrature is bound by the FP32 datapath of the CUDA cores, not by HBM (SURVEY.md §8d, DESIGN.md §4); see roofline_fp32_pipe, roofline_issue, roofline_compute"}
        if stale:
            roofline["traffic_note"] = "null: " + stale
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": n_total, "steps": K, "warmup": W,
                "ms_per_step": elapsed_ms / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": prec, "data": "synthetic",
                "config": {"workload": "kline5 batch of 1e5 random parameter points per GPU and step (5 emissivity knots, rcut), "
                                       "1000-bin 0.1-10 keV grid, synthetic Kerr-shaped table 30x20x150x30 "
                                       "(BASELINE.json configs[1])",
                           "sets_per_gpu_per_step": B, "precision": prec, "parallelism": f"sets sharded x{n_total}",
                           "processes": ("one per GPU (torchrun)" if world > 1 else ("one process driving %d GPUs (klp_group_create_local)" % n_local if n_local > 1 else "one")),
                           "gather": (f"klp_group_eval_allgather_device: {gather_ops} grouped NCCL operations per step on a high-priority stream, "
                                      "each overlapping the next chunk's kernels" if group is not None else "none (1 GPU)"),
                           "inputs": "parameters resident in HBM before the timed region (their H2D copy is inside `e2e`, not `value`)",
                           "l2": "256 MiB buffer zeroed between steps (L2 flush) and a fresh parameter batch every step"},
                "e2e": e2e, "gpu_launches": launches, "clocks": clocks, "roofline": roofline}
        if tj is not None and clocks and clocks.get("sm_mhz"):
            # the bound that applies: warp-instruction issue slots (4 per SM per clock); instructions per spectrum
            # from the ncu capture, rate from this run
            wi = tj["warp_instructions_per_set"]
            peak_issue = 4.0 * torch.cuda.get_device_properties(dev).multi_processor_count * clocks["sm_mhz"] * 1e6
            ach_issue = wi * sets_per_launch / (avg_launch_ms * 1e-3)
            src = os.path.relpath(tp, ROOT)
            line["roofline_issue"] = {"bound": "warp_instruction_issue", "achieved": ach_issue / 1e9, "peak": peak_issue / 1e9,
                                      "unit": "G warp-inst/s", "frac": ach_issue / peak_issue,
                                      "warp_instructions_per_spectrum": wi,
                                      "source": f"{src} (ncu smsp__inst_executed.sum, kernel {tj['kernel']}) x measured launch rate; "
                                                "peak = 4 x SMs x sm clock sampled during the run"}
            if "fma_pipe_cycles_per_set" in tj:
                # the bound that binds: cycles of the FP32 (FMA) datapath, one per sub-partition and clock
                ach_pipe = tj["fma_pipe_cycles_per_set"] * sets_per_launch / (avg_launch_ms * 1e-3)
                line["roofline_fp32_pipe"] = {"bound": "fp32_datapath_cycles", "achieved": ach_pipe / 1e9, "peak": peak_issue / 1e9,
                                              "unit": "G pipe-cycles/s", "frac": ach_pipe / peak_issue,
                                              "source": f"{src} (ncu sm__pipe_fma_cycles_active) x measured launch rate"}
            roofline["traffic_note"] = tj.get("note")
        elif stale:
            line["roofline_issue"] = None
            line["roofline_fp32_pipe"] = None
        try:
            pk = gt.fma_peak_gflops(prec)
        except Exception:
            pk = None
        cb, evals = (None, None)
        if not args.no_cpu_baseline and n_total == 1:
            cb, evals = cpu_baseline(syn)
            line["cpu_baseline"] = cb
        if evals and pk:
            flops_per_set = 22.0 * evals  # SURVEY.md §8d A_F = 22 * N_eval
            ach = flops_per_set * sets_per_launch / (avg_launch_ms * 1e-3) / 1e9
            line["roofline_compute"] = {"bound": f"{prec}_cuda_core", "achieved": ach, "peak": pk, "unit": "GFLOP/s",
                                        "frac": ach / pk, "integrand_evals_per_spectrum": evals,
                                        "peak_source": "klp_fma_peak_gflops (FMA microbenchmark on this GPU)"}
        if n_total == 1 and not args.no_configs:
            line["configs"] = other_configs(klp, syn, gt, dev, peak, with_cpu=not args.no_cpu_baseline)
        print(json.dumps(line), flush=True)
    if group is not None:
        group.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
