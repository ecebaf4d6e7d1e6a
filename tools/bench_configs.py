#!/usr/bin/env python
"""Measure the five BASELINE.json configurations.  One JSON line per config on rank 0.

  python tools/bench_configs.py [--configs 1,2,3,4,5] [--scale 1.0]
  python -m torch.distributed.run --nproc-per-node N ... tools/bench_configs.py   (configs 4, 5 shard over ranks)

Times are CUDA-event times of the device-resident path (max over ranks); the CPU figure is the oracle
(reference algorithm and cost structure, f32) on all host cores over a bounded sample.
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import time

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import lineprofiles_b200 as klp  # noqa: E402
from lineprofiles_b200 import synthetic as syn  # noqa: E402
from lineprofiles_b200.dist import shard_bounds  # noqa: E402

WORLD = int(os.environ.get("WORLD_SIZE", "1"))
RANK = int(os.environ.get("RANK", "0"))
LOCAL = int(os.environ.get("LOCAL_RANK", "0"))


def cores():
    return len(os.sched_getaffinity(0))


def emit(d):
    if RANK == 0:
        d["n_gpus"] = WORLD
        print(json.dumps(d), flush=True)


def maxrank(x):
    if WORLD == 1:
        return x
    t = torch.tensor([x], dtype=torch.float64, device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def timed(fn, reps=3, warm=1):
    for _ in range(warm):
        fn()
    best = None
    for _ in range(reps):
        if WORLD > 1:
            dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        ms = maxrank(e0.elapsed_time(e1))
        best = ms if best is None else min(best, ms)
    return best


def cpu_rate(ot, model, energy, spec, n_sample, prec="f32"):
    p = syn.random_params(model, n_sample, seed=4242)
    t0 = time.perf_counter()
    ot.eval_batch(model, p, energy, spectrum=spec, precision=prec, threads=cores(), faithful_cost=True)
    return n_sample / (time.perf_counter() - t0)


def device_eval(gt, model, p, E, n, out, spec=None, prec="f32"):
    st = torch.cuda.current_stream().cuda_stream
    gt.eval_batch_device(model, p, E, n, out, p.shape[0], d_spectrum=spec, precision=prec, stream=st)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--configs", default="1,2,3,4,5")
    ap.add_argument("--scale", type=float, default=1.0, help="scale the batch sizes (smoke runs)")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--precisions", default="f32,f64")
    args = ap.parse_args()
    cfgs = [int(c) for c in args.configs.split(",")]
    precs = args.precisions.split(",")
    torch.cuda.set_device(LOCAL)
    if WORLD > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", LOCAL))
    from oracle import oracle_py as orc
    tab = syn.make_table()
    gt = klp.Table.from_arrays(tab, LOCAL)
    ot = None if args.no_cpu or RANK != 0 else orc.OracleTable(tab)
    E1k = syn.energy_grid_linear(1000)
    E4k = syn.energy_grid_log(4096)
    spec4k = syn.powerlaw_continuum(E4k)
    dE1k, dE4k, dspec = (torch.from_numpy(a).cuda() for a in (E1k, E4k, spec4k))

    if 1 in cfgs and RANK == 0:
        # config 1: one kline evaluation through the XSPEC symbol, host buffers, wall clock
        klp.set_default_table(gt)
        res = {}
        for prec in precs:
            klp.set_legacy_precision(prec)
            klp.legacy_reset()
            flux = klp.kline(E1k, syn.CONFIG1_PARAMS[0])
            ts = []
            for _ in range(20):
                klp.legacy_reset()
                t0 = time.perf_counter()
                klp.kline(E1k, syn.CONFIG1_PARAMS[0])
                ts.append(time.perf_counter() - t0)
            res[prec] = {"ms_per_call": 1e3 * float(np.median(ts)), "sum": float(flux.sum())}
            if ot is not None:
                want = ot.eval_batch("kline", syn.CONFIG1_PARAMS, E1k, precision=prec, faithful_cost=True)["out"][0]
                res[prec]["max_rel_err_vs_oracle"] = float(np.max(np.abs(flux - want) / (np.abs(want) + 1e-12 * want.max())))
        klp.set_default_table(None)
        cpu_ms = None
        if ot is not None:
            t0 = time.perf_counter()
            ot.eval_batch("kline", syn.CONFIG1_PARAMS, E1k, precision="f32", faithful_cost=True)
            cpu_ms = 1e3 * (time.perf_counter() - t0)
        emit({"config": 1, "what": "kline single evaluation a=0.998 incl=30 eline=6.4 rmin=ISCO rout=400 alpha=3, 1000-bin grid, via the exported kline symbol",
              "gpu": res, "cpu_oracle_f32_ms_per_call_1_thread": cpu_ms})

    if 2 in cfgs:
        B = max(1024, int(100000 * args.scale))
        lo, hi = shard_bounds(B, WORLD, RANK)
        p = torch.from_numpy(syn.random_params("kline5", B)[lo:hi]).cuda()
        out = torch.empty((hi - lo, 1000), dtype=torch.float64, device="cuda")
        r = {}
        for prec in precs:
            ms = timed(lambda: device_eval(gt, "kline5", p, dE1k, 1000, out, prec=prec))
            r[prec] = {"spectra_per_s": B / (ms * 1e-3), "ms": ms}
        cpu = cpu_rate(ot, "kline5", E1k, None, 40 * cores()) if ot is not None else None
        emit({"config": 2, "what": f"kline5 batch of {B} random parameter points, 1000-bin grid", "gpu": r,
              "cpu_oracle_f32_spectra_per_s": cpu, "cpu_cores": cores()})

    if 3 in cfgs:
        B = max(256, int(10000 * args.scale))
        lo, hi = shard_bounds(B, WORLD, RANK)
        p = torch.from_numpy(syn.random_params("kconv", B)[lo:hi]).cuda()
        out = torch.empty((hi - lo, 4096), dtype=torch.float64, device="cuda")
        r = {}
        for prec in precs:
            gt.set_timing(True)
            ms = timed(lambda: device_eval(gt, "kconv", p, dE4k, 4096, out, spec=dspec, prec=prec))
            tm = gt.get_timing()
            gt.set_timing(False)
            r[prec] = {"spectra_per_s": B / (ms * 1e-3), "ms": ms, "quad_ms": tm["quad"]["ms"], "conv_ms": tm["conv"]["ms"]}
        # the cumulative-profile convolution (option conv_mode = 1: not the reference's operation order, equal to rounding)
        gt.set_option("conv_mode", 1)
        gt.set_timing(True)
        ms = timed(lambda: device_eval(gt, "kconv", p, dE4k, 4096, out, spec=dspec, prec="f32"))
        tm = gt.get_timing()
        gt.set_timing(False)
        gt.set_option("conv_mode", 0)
        r["f32_conv_mode_1"] = {"spectra_per_s": B / (ms * 1e-3), "ms": ms, "quad_ms": tm["quad"]["ms"], "conv_ms": tm["conv"]["ms"]}
        cpu = cpu_rate(ot, "kconv", E4k, spec4k, max(cores(), 16)) if ot is not None else None
        emit({"config": 3, "what": f"kconv, batch of {B}, power-law continuum on the 4096-bin log grid", "gpu": r,
              "cpu_oracle_f32_spectra_per_s": cpu, "cpu_cores": cores()})

    if 4 in cfgs:
        # kconv10 MCMC-style batch of 1e6 walkers sharded over the ranks through the C ABI's group entry
        # (klp_group_eval_allgather_device): every GPU ends up with all spectra (32.8 GB at 1e6 x 4096 x f64); the NCCL gather
        # of one chunk overlaps the kernels of the next.  On 1 GPU the 1/8 subset is run.
        from lineprofiles_b200 import dist as kdist
        B = int(1000000 * args.scale) if WORLD > 1 else int(125000 * args.scale)
        per = max(1024, B // WORLD)
        B = per * WORLD
        params = syn.random_params("kconv10", B)[RANK * per:(RANK + 1) * per]
        dp = torch.from_numpy(params).cuda()
        full = torch.empty((B, 4096), dtype=torch.float64, device="cuda")
        group = kdist.rank_group(gt) if WORLD > 1 else klp.Group.local([gt])
        st = torch.cuda.current_stream().cuda_stream

        def run():
            group.eval_allgather_device("kconv10", [dp], [dE4k], 4096, [full], per, d_spectrum=[dspec], precision="f32", streams=[st])
        ms = timed(run, reps=1, warm=0 if B >= 100000 else 1)
        ops = group.last_gather_ops()
        gt.set_option("conv_mode", 1)
        ms1 = timed(run, reps=1, warm=0)
        gt.set_option("conv_mode", 0)
        # every rank holds every spectrum: a checksum of checksums over the ranks must agree
        chk = float(full.sum().item())
        same = True
        if WORLD > 1:
            t = torch.tensor([chk, -chk], dtype=torch.float64, device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            same = bool(t[0].item() == chk and -t[1].item() == chk)
            same = bool(maxrank(0.0 if same else 1.0) == 0.0)
        emit({"config": 4, "what": f"kconv10, {B} walkers sharded x{WORLD}, 4096-bin grid, klp_group_eval_allgather_device ({ops} grouped NCCL operations, each overlapping the next chunk's kernels)",
              "gpu": {"f32": {"spectra_per_s": B / (ms * 1e-3), "ms": ms}, "f32_conv_mode_1": {"spectra_per_s": B / (ms1 * 1e-3), "ms": ms1}},
              "gathered_bytes_total": B * 4096 * 8 if WORLD > 1 else 0, "all_ranks_hold_identical_spectra": same})
        group.close()
        del full

    if 5 in cfgs:
        # table resolution sweep: fused kline throughput + stand-alone blend kernel GB/s
        for dims in ((30, 20, 150, 30), (60, 40, 300, 30), (90, 60, 450, 48), (120, 80, 600, 60)):
            n_a, n_mu, n_r, n_g = dims
            nbytes = n_a * n_mu * n_r * (2 * n_g + 3) * 4
            if nbytes * args.scale > 4e9:
                continue
            t0 = time.perf_counter()
            tb = syn.make_table(*dims)
            g5 = klp.Table.from_arrays(tb, LOCAL)
            del tb
            gen_s = time.perf_counter() - t0
            B = max(1024, int(100000 * args.scale))
            lo, hi = shard_bounds(B, WORLD, RANK)
            rng = np.random.default_rng(5)
            pk = syn.random_params("kline", B, seed=55)[lo:hi]
            p = torch.from_numpy(pk).cuda()
            out = torch.empty((hi - lo, 1000), dtype=torch.float64, device="cuda")
            ms = timed(lambda: device_eval(g5, "kline", p, dE1k, 1000, out, prec="f32"), reps=2)
            # blend: as many sets as fit ~8 GB of output
            frame_len = n_r * (2 * n_g + 3)
            nb = int(min(200000, 8e9 // (frame_len * 4)))
            ai = torch.from_numpy(np.stack([rng.uniform(0, 0.998, nb), rng.uniform(5, 85, nb)], axis=1)).cuda()
            bo = torch.empty((nb, frame_len), dtype=torch.float32, device="cuda")
            st = torch.cuda.current_stream().cuda_stream
            bms = timed(lambda: g5.blend_frames_device(ai, bo, nb, precision="f32", stream=st), reps=3)
            gbs = nb * frame_len * 4 * 5 / (bms * 1e-3) / 1e9
            emit({"config": 5, "table_dims": dims, "table_MB": nbytes / 1e6, "table_build_s": gen_s,
                  "kline_f32": {"spectra_per_s": B / (ms * 1e-3), "ms": ms, "batch": B},
                  "blend_f32": {"sets": nb, "ms": bms, "algorithmic_GBps": gbs, "bytes_per_set": frame_len * 4 * 5,
                                "frac_of_measured_hbm_6462": gbs / 6462.1}})
            g5.close()
            del bo, ai, out, p
            torch.cuda.empty_cache()
    if WORLD > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
