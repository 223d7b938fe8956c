"""N>1 path on CPU: two gloo ranks build their per-rank C5a grids (weak scaling, distinct
times per rank), 'evaluate' with a stand-in, and gather to rank 0 exactly as bench.py does."""
import os
import socket
import sys

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import bench
    from oracle import oracle
    d, t, r, z = bench.c5a_grid(rank, nr=4, nz=8, nt=2)
    p, tD, sv, rD, zD, lay = bench.derive(d, t, r, z, oracle)
    # stand-in for the GPU evaluation: a deterministic function of the inputs
    s = torch.from_numpy((tD[:, None, None] + rD[None, :, None] + zD[None, None, :]).ravel().copy())
    gathered = [torch.empty_like(s) for _ in range(world)] if rank == 0 else None
    dist.gather(s, gathered, dst=0)
    tot = torch.tensor([float(s.numel())])
    dist.all_reduce(tot, op=dist.ReduceOp.SUM)
    if rank == 0:
        np.save(out, torch.stack(gathered).numpy())
        assert tot.item() == world * 64
    dist.destroy_process_group()


def test_two_rank_weak_scaling_gather(tmp_path):
    out = str(tmp_path / "g.npy")
    port = _free_port()
    mp.spawn(_worker, args=(2, port, out), nprocs=2, join=True)
    g = np.load(out)
    assert g.shape == (2, 64)
    import bench
    from oracle import oracle
    for rank in range(2):
        d, t, r, z = bench.c5a_grid(rank, nr=4, nz=8, nt=2)
        p, tD, sv, rD, zD, lay = bench.derive(d, t, r, z, oracle)
        want = (tD[:, None, None] + rD[None, :, None] + zD[None, None, :]).ravel()
        assert np.array_equal(g[rank], want)
    assert not np.array_equal(g[0], g[1])      # ranks own distinct work


def test_flop_model_matches_survey_formula():
    import bench
    from oracle import oracle
    d, t, r, z = bench.c5a_grid(0)
    p, tD, sv, rD, zD, lay = bench.derive(d, t, r, z, oracle)
    F = bench.flops_per_point(p, lay, len(zD))
    n1 = int((lay == 1).sum()); n2 = int((lay == 2).sum()); n3 = int((lay == 3).sum())
    capz = (135 * n1 + 275 * n2 + 375 * n3) / 128
    want = 703 * 53 * (870 / 128 + capz) + 53 * (25 * 8 + 144 * 45) + 2 * (26 * 26 * 60 + 26 * 60)
    assert abs(F - want) < 1e-6 and 5e6 < F < 1.2e7
