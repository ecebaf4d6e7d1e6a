"""Multi-GPU behind the C ABI (klp_group_*): one process driving several GPUs, and one process per GPU.

CPU part: the entry points exist, refuse host-only tables and bad arguments, NCCL is found by dlopen.
GPU part (`-m gpu`): a one-member group (runs on any GPU box, NCCL world of one), and -- when the box has at least two
GPUs (`gpurun --gpus 2`) -- a two-GPU local group: host-buffer sharding and the chunked device-resident gather, against
the single-GPU evaluation of the same sets (bit-identical: sets are independent)."""
import os

import numpy as np
import pytest

import lineprofiles_b200 as klp
from lineprofiles_b200 import synthetic as syn


def test_group_refuses_host_only_tables(table_arrays):
    t = klp.Table.from_arrays(table_arrays, device=-1)
    with pytest.raises(klp.KlpError, match="no CPU path"):
        klp.Group.local([t])
    with pytest.raises(klp.KlpError, match="no CPU path"):
        klp.Group.from_rank(t, b"\0" * 128, 0, 1)
    with pytest.raises(klp.KlpError):
        klp.Group.local([])
    t.close()


def test_nccl_is_loaded_lazily_by_dlopen():
    assert klp.nccl_version() >= 22000
    assert len(klp.nccl_unique_id()) == 128


def test_gather_chunk_plan():
    """Few big chunks (every k_quad launch has an idle tail) and a short final chunk (its gather is the only one nothing overlaps);
    small batches are one chunk = one ncclAllGather."""
    from lineprofiles_b200.api import group_chunk_plan as plan
    assert plan(100000) == [97952, 2048]
    assert plan(100000, 65535, 4096) == [65535, 30369, 4096]
    assert plan(1000000, 131072, 2048) == [131072] * 7 + [1000000 - 7 * 131072 - 2048, 2048]
    assert plan(3000) == [3000] and plan(1024) == [1024] and plan(4096) == [4096]      # B <= 2 * tail: no tail chunk
    assert plan(5000, 1000, 0) == [1000] * 5
    assert plan(2500, 131072, 300) == [2200, 300]
    assert plan(0) == []
    for B in (1, 17, 4097, 65536, 123457):
        assert sum(plan(B)) == B and all(c > 0 for c in plan(B))


def _ndev():
    import torch
    return torch.cuda.device_count() if torch.cuda.is_available() else 0


@pytest.mark.gpu
def test_single_member_group_equals_table(gpu_table):
    import torch
    E = syn.energy_grid_linear(1000)
    p = syn.random_params("kline5", 3000, seed=61)
    want = gpu_table.eval_batch("kline5", p, E, precision="f32")
    g = klp.Group.local([gpu_table])
    try:
        assert (g.n_local, g.nranks) == (1, 1)
        assert np.array_equal(g.eval_batch("kline5", p, E, precision="f32"), want)
        dp, dE = torch.from_numpy(p).cuda(), torch.from_numpy(E).cuda()
        out = torch.zeros((3000, 1000), dtype=torch.float64, device="cuda")
        g.set_option("gather_chunk", 1000)
        g.eval_allgather_device("kline5", [dp], [dE], 1000, [out], 3000, streams=[torch.cuda.current_stream().cuda_stream])
        torch.cuda.synchronize()
        assert np.array_equal(out.cpu().numpy(), want)
    finally:
        g.close()


@pytest.mark.gpu
def test_rank_group_world_of_one_through_dist_layer(gpu_table):
    """lineprofiles_b200.dist with the CUDA evaluator over NCCL at world size 1: the code the multi-process bench runs."""
    import torch
    import torch.distributed as dist
    from lineprofiles_b200 import dist as kdist
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    os.environ.setdefault("MASTER_PORT", "29531")
    own = not dist.is_initialized()
    if own:
        dist.init_process_group("nccl", rank=0, world_size=1, device_id=torch.device("cuda", 0))
    try:
        kg = kdist.rank_group(gpu_table)
        E = syn.energy_grid_log(256)
        spec = syn.powerlaw_continuum(E)
        p = syn.random_params("kconv5", 37, seed=62)
        got = kdist.evaluate_sharded_device(kg, "kconv5", torch.from_numpy(p).cuda(), torch.from_numpy(E).cuda(), 256,
                                            d_spectrum=torch.from_numpy(spec).cuda())
        torch.cuda.synchronize()
        assert np.array_equal(got.cpu().numpy(), gpu_table.eval_batch("kconv5", p, E, spectrum=spec, precision="f32"))
        # the generic sharding logic with the CUDA evaluator (one all_gather_into_tensor)
        dE = torch.from_numpy(syn.energy_grid_linear(100)).cuda()
        p2 = syn.random_params("kline", 11, seed=63)

        def evaluate(ps, out):
            gpu_table.eval_batch_device("kline", ps, dE, 100, out, ps.shape[0], stream=torch.cuda.current_stream().cuda_stream)

        full = kdist.evaluate_sharded(evaluate, torch.from_numpy(p2).cuda(), 100)
        torch.cuda.synchronize()
        assert np.array_equal(full.cpu().numpy(), gpu_table.eval_batch("kline", p2, syn.energy_grid_linear(100), precision="f32"))
        kg.close()
    finally:
        if own:
            dist.destroy_process_group()


@pytest.mark.gpu
@pytest.mark.skipif(_ndev() < 2, reason="needs two GPUs (gpurun --gpus 2)")
def test_two_gpu_local_group(table_arrays, gpu_table):
    import torch
    t1 = klp.Table.from_arrays(table_arrays, device=1)
    g = klp.Group.local([gpu_table, t1])
    try:
        E = syn.energy_grid_linear(1000)
        # host buffers: odd batch, shards of different length, no collective
        p = syn.random_params("kline5", 5001, seed=64)
        want = gpu_table.eval_batch("kline5", p, E, precision="f32")
        assert np.array_equal(g.eval_batch("kline5", p, E, precision="f32"), want)
        # device-resident: 2 x 2500 sets, chunks of 1000 -> several grouped broadcasts; then one chunk -> one all-gather
        B = 2500
        ps = [torch.from_numpy(p[:B]).to("cuda:0"), torch.from_numpy(p[B:2 * B]).to("cuda:1")]
        Es = [torch.from_numpy(E).to(f"cuda:{d}") for d in (0, 1)]
        torch.cuda.synchronize(0)
        torch.cuda.synchronize(1)
        for chunk, tail, n_ops in ((1000, 4096, 3), (131072, 4096, 1), (131072, 300, 2)):   # [1000, 1000, 500]; one all-gather; [2200, 300]
            outs = [torch.zeros((2 * B, 1000), dtype=torch.float64, device=f"cuda:{d}") for d in (0, 1)]
            torch.cuda.synchronize(0)
            torch.cuda.synchronize(1)   # the group's own streams do not wait for torch's
            g.set_option("gather_chunk", chunk)
            g.set_option("gather_tail", tail)
            g.eval_allgather_device("kline5", ps, Es, 1000, outs, B)
            g.synchronize()
            assert g.last_gather_ops() == n_ops
            for d in (0, 1):
                assert np.array_equal(outs[d].cpu().numpy(), want[: 2 * B]), (chunk, tail, d)
        # convolution model with per-set spectra
        E2 = syn.energy_grid_log(300)
        pc = syn.random_params("kconv", 600, seed=65)
        spec = np.random.default_rng(66).uniform(0.5, 1.5, (600, 300))
        wantc = gpu_table.eval_batch("kconv", pc, E2, spectrum=spec, precision="f32")
        assert np.array_equal(g.eval_batch("kconv", pc, E2, spectrum=spec, precision="f32"), wantc)
        outs = [torch.zeros((600, 300), dtype=torch.float64, device=f"cuda:{d}") for d in (0, 1)]
        g.set_option("gather_chunk", 128)
        dev_in = [torch.from_numpy(pc[:300]).to("cuda:0"), torch.from_numpy(pc[300:]).to("cuda:1"),
                  torch.from_numpy(spec[:300]).to("cuda:0"), torch.from_numpy(spec[300:]).to("cuda:1")]
        torch.cuda.synchronize(0)
        torch.cuda.synchronize(1)
        g.eval_allgather_device("kconv", dev_in[:2],
                                [torch.from_numpy(E2).to(f"cuda:{d}") for d in (0, 1)], 300, outs, 300, d_spectrum=dev_in[2:],
                                spectrum_per_set=True)
        g.synchronize()
        for d in (0, 1):
            assert np.array_equal(outs[d].cpu().numpy(), wantc)
    finally:
        g.close()
        t1.close()


def _build_c_client(tmp_path):
    import shutil
    import subprocess
    if shutil.which("gcc") is None:
        pytest.skip("no gcc")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    libdir = os.path.dirname(klp.api.library_path())
    exe = tmp_path / "group_client"
    cc = subprocess.run(["gcc", "-std=c99", "-Wall", "-Wextra", "-pedantic", "-Werror", "-I", os.path.join(root, "include"),
                         os.path.join(root, "tests", "c", "group_client.c"), "-o", str(exe), "-L", libdir, "-lklineprofiles_b200",
                         f"-Wl,-rpath,{libdir}"], capture_output=True, text=True)
    assert cc.returncode == 0, cc.stderr
    return exe


def test_c_group_client_builds_and_refuses_without_device(tmp_path, table_arrays):
    """tests/c/group_client.c is the C99 view of the multi-GPU entry points; without a CUDA device it reports 'skipped' (77)."""
    import subprocess
    from lineprofiles_b200 import tableio
    exe = _build_c_client(tmp_path)
    tableio.save_table(str(tmp_path / "t.klpt"), syn.make_table(3, 4, 12, 30))
    if _ndev() == 0:
        run = subprocess.run([str(exe), str(tmp_path / "t.klpt"), "8"], capture_output=True, text=True)
        assert run.returncode == 77, run.stdout + run.stderr


@pytest.mark.gpu
def test_c_group_client_drives_the_visible_gpus(tmp_path, table_arrays):
    """The same C program on the GPU box: one table replica per GPU (two when present), klp_group_eval_batch == klp_eval_batch."""
    import subprocess
    from lineprofiles_b200 import tableio
    exe = _build_c_client(tmp_path)
    tableio.save_table(str(tmp_path / "t.klpt"), table_arrays)
    run = subprocess.run([str(exe), str(tmp_path / "t.klpt"), "3001"], capture_output=True, text=True)
    assert run.returncode == 0, run.stdout + run.stderr
    assert "group == single GPU: yes" in run.stdout
    assert f"devices {min(_ndev(), 2)}" in run.stdout
