"""World-size-2 gloo test of the multi-GPU plumbing (lineprofiles_b200/dist.py): shard bounds and the single
all-gather.  The evaluator here is the CPU oracle (tests may use it); on the B200 box the same
function is driven with the CUDA evaluator (bench.py, tests/test_gpu_parity.py)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from lineprofiles_b200 import synthetic as syn
from lineprofiles_b200.dist import evaluate_sharded, shard_bounds


def test_shard_bounds_cover_everything():
    for n in (0, 1, 7, 8, 9, 100000):
        for w in (1, 2, 3, 8):
            blocks = [shard_bounds(n, w, r) for r in range(w)]
            assert blocks[0][0] == 0 and blocks[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(blocks, blocks[1:]))
            assert max(hi - lo for lo, hi in blocks) <= -(-n // w) if n else True


def _worker(rank, world, port, B, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from oracle import oracle_py as orc
    ot = orc.OracleTable(syn.make_table(6, 5, 40, 30))
    E = syn.energy_grid_linear(50)
    params = torch.from_numpy(syn.random_params("kline5", B, seed=77))

    def evaluate(p, out):
        out.copy_(torch.from_numpy(ot.eval_batch("kline5", p.numpy(), E, precision="f32")["out"]))

    full = evaluate_sharded(evaluate, params, 50)
    if rank == 0:
        q.put(full.numpy())
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("B", [5, 8])
def test_two_rank_gather_equals_single_process(B):
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, B, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = q.get(timeout=240)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    from oracle import oracle_py as orc
    ot = orc.OracleTable(syn.make_table(6, 5, 40, 30))
    want = ot.eval_batch("kline5", syn.random_params("kline5", B, seed=77), syn.energy_grid_linear(50), precision="f32")["out"]
    assert got.shape == (B, 50)
    assert np.array_equal(got, want)
