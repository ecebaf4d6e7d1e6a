#!/usr/bin/env python
"""bench.py -- kline-family spectra/s on B200 (BASELINE.json metric), with roofline and CPU baseline.

  python bench.py --gpus N --steps K --warmup W            our arm (CUDA, one process per GPU)
  python bench.py --impl reference --gpus N --steps K ...  the reference algorithm on the host cores

Workload (BASELINE.json configs[1]): kline5, a batch of 1e5 random parameter points per GPU and step
(5 emissivity knots, rcut) on the 1000-bin 0.1-10 keV grid, synthetic Kerr-shaped 30x20x150x30 table.
A step = one pass of the hot path over one batch.  `value` times the device-resident path (+ the
NCCL all-gather of the spectra for N > 1); `e2e` times Table.eval_batch with host buffers
(params H2D from pinned staging, every spectrum D2H).
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


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--precision", default="f32", choices=["f32", "f64"])
    ap.add_argument("--batch", type=int, default=100000, help="parameter sets per GPU per step")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    if args.impl == "reference":
        run_reference(args)
        return

    import torch
    import torch.distributed as dist

    import lineprofiles_b200 as klp
    from lineprofiles_b200 import synthetic as syn

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- this framework has no CPU path")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    B, K, W, prec = args.batch, args.steps, args.warmup, args.precision
    tab = syn.make_table()
    gt = klp.Table.from_arrays(tab, device=local_rank)
    energy = syn.energy_grid_linear(N_BINS)
    P = klp.NPARAMS[MODEL]
    # a different parameter batch every step (nothing can be reused from a previous step)
    host_params = [syn.random_params(MODEL, B, seed=1000 * (rank + 1) + s) for s in range(W + K)]
    d_params = [torch.from_numpy(p).to(dev) for p in host_params]
    d_energy = torch.from_numpy(energy).to(dev)
    d_out = torch.empty((B, N_BINS), dtype=torch.float64, device=dev)
    d_all = torch.empty((world * B, N_BINS), dtype=torch.float64, device=dev) if world > 1 else None
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)  # > 126 MB L2
    stream = torch.cuda.current_stream().cuda_stream

    def step(i):
        flush.zero_()
        gt.eval_batch_device(MODEL, d_params[i], d_energy, N_BINS, d_out, B, precision=prec, stream=stream)
        if world > 1:
            dist.all_gather_into_tensor(d_all, d_out)  # the single NCCL gather of the spectra

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for i in range(W):
        step(i)
    barrier()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    gt.set_timing(True)
    quad_ms, quad_launches, launches = 0.0, 0, 0
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    ev0.record()
    for i in range(W, W + K):
        step(i)
        if i == W + K - 1:
            pass
    ev1.record()
    barrier()
    elapsed_ms = ev0.elapsed_time(ev1)
    # per-kernel device times of the LAST timed step (events recorded by the library on this stream)
    tm = gt.get_timing()
    gt.set_timing(False)
    clocks = sampler.stop() if rank == 0 else None
    quad_ms, quad_launches = tm["quad"]["ms"], tm["quad"]["launches"]
    launches = sum(v["launches"] for v in tm.values()) * K
    if world > 1:
        t = torch.tensor([elapsed_ms], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        elapsed_ms = float(t.item())
    value = world * B * K / (elapsed_ms * 1e-3)

    # ---- e2e through the public host API: numpy in, numpy out, copies inside the timed region
    e2e = None
    if not args.no_e2e:
        out_h = np.zeros((B, N_BINS), np.float64)   # touched: the timed region must not pay the first-touch page faults
        out_h.fill(0.0)
        gt.eval_batch(MODEL, host_params[0], energy, precision=prec, out=out_h)  # untimed warm-up step: sizes the pinned staging buffers
        ke = max(1, min(K, 3))
        barrier()
        t0 = time.perf_counter()
        for i in range(ke):
            gt.eval_batch(MODEL, host_params[W + i], energy, precision=prec, out=out_h)
        barrier()
        dt = time.perf_counter() - t0
        if world > 1:
            t = torch.tensor([dt], dtype=torch.float64, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            dt = float(t.item())
        e2e = {"value": world * B * ke / dt, "unit": UNIT, "h2d_bytes_per_step": B * P * 8 + (N_BINS + 1) * 8,
               "d2h_bytes_per_step": B * N_BINS * 8, "steps": ke,
               "api": "lineprofiles_b200.Table.eval_batch -> klp_eval_batch (host buffers, pinned staging, 3 streams)"}

    if rank == 0:
        peak, peak_src = measured_peak()
        n_r, n_g = tab["n_r"], tab["n_g"]
        bytes_per_set = 8 * P + 4 * 4 * n_r * (2 * n_g + 3) + 8 * N_BINS  # SURVEY.md §8d A_B
        sets_per_launch = B / max(quad_launches, 1)
        avg_launch_ms = quad_ms / max(quad_launches, 1)
        achieved = bytes_per_set * sets_per_launch / (avg_launch_ms * 1e-3) / 1e9 if avg_launch_ms > 0 else None
        traffic = None   # dram__bytes_read.sum + dram__bytes_write.sum of one ncu --set full capture, scaled per launch
        tp = os.path.join(ROOT, "profiles", "r01_quad_traffic.json")
        if os.path.exists(tp) and prec == "f32":
            tj = json.load(open(tp))
            traffic = (tj["dram_bytes_read"] + tj["dram_bytes_write"]) * sets_per_launch / tj["sets_per_launch"]
        roofline = {"bound": "hbm", "kernel": "k_quad", "achieved": achieved, "peak": peak, "unit": "GB/s",
                    "frac": (achieved / peak) if achieved else None, "traffic": traffic, "peak_source": peak_src,
                    "algorithmic_bytes_per_spectrum": bytes_per_set, "avg_launch_ms": avg_launch_ms,
                    "sets_per_launch": sets_per_launch,
                    "note": "the quadrature is bound by the FP32 datapath of the CUDA cores, not by HBM (SURVEY.md §8d, DESIGN.md §4); see roofline_fp32_pipe, roofline_issue, roofline_compute"}
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W,
                "ms_per_step": elapsed_ms / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": prec, "data": "synthetic",
                "config": {"workload": "kline5 batch of 1e5 random parameter points per GPU and step (5 emissivity knots, rcut), "
                                       "1000-bin 0.1-10 keV grid, synthetic Kerr-shaped table 30x20x150x30 "
                                       "(BASELINE.json configs[1])",
                           "sets_per_gpu_per_step": B, "precision": prec, "parallelism": f"sets sharded x{world}",
                           "gather": "one NCCL all_gather_into_tensor of the spectra per step" if world > 1 else "none (1 GPU)",
                           "l2": "256 MiB buffer zeroed between steps (L2 flush) and a fresh parameter batch every step"},
                "e2e": e2e, "gpu_launches": launches, "clocks": clocks, "roofline": roofline}
        if traffic is not None and clocks and clocks.get("sm_mhz"):
            # the bound that applies: warp-instruction issue slots (4 per SM per clock); instructions per spectrum
            # from the ncu capture, rate from this run
            wi = tj["warp_instructions_per_set"]
            peak_issue = 4.0 * torch.cuda.get_device_properties(dev).multi_processor_count * clocks["sm_mhz"] * 1e6
            ach_issue = wi * sets_per_launch / (avg_launch_ms * 1e-3)
            line["roofline_issue"] = {"bound": "warp_instruction_issue", "achieved": ach_issue / 1e9, "peak": peak_issue / 1e9,
                                      "unit": "G warp-inst/s", "frac": ach_issue / peak_issue,
                                      "warp_instructions_per_spectrum": wi,
                                      "source": "profiles/r01_quad_traffic.json (ncu smsp__inst_executed.sum) x measured launch rate; "
                                                "peak = 4 x SMs x sm clock sampled during the run"}
            if "fma_pipe_cycles_per_set" in tj:
                # the bound that binds: cycles of the FP32 (FMA) datapath, one per sub-partition and clock
                ach_pipe = tj["fma_pipe_cycles_per_set"] * sets_per_launch / (avg_launch_ms * 1e-3)
                line["roofline_fp32_pipe"] = {"bound": "fp32_datapath_cycles", "achieved": ach_pipe / 1e9, "peak": peak_issue / 1e9,
                                              "unit": "G pipe-cycles/s", "frac": ach_pipe / peak_issue,
                                              "source": "profiles/r01_quad_traffic.json (ncu sm__pipe_fma_cycles_active) x measured launch rate"}
            roofline["traffic_note"] = tj.get("note")
        try:
            pk = gt.fma_peak_gflops(prec)
        except Exception:
            pk = None
        cb, evals = (None, None)
        if not args.no_cpu_baseline:
            cb, evals = cpu_baseline(syn)
            line["cpu_baseline"] = cb
        if evals and pk:
            flops_per_set = 22.0 * evals  # SURVEY.md §8d A_F = 22 * N_eval
            ach = flops_per_set * sets_per_launch / (avg_launch_ms * 1e-3) / 1e9
            line["roofline_compute"] = {"bound": f"{prec}_cuda_core", "achieved": ach, "peak": pk, "unit": "GFLOP/s",
                                        "frac": ach / pk, "integrand_evals_per_spectrum": evals,
                                        "peak_source": "klp_fma_peak_gflops (FMA microbenchmark on this GPU)"}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
