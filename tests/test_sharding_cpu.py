"""CPU tests of the multi-GPU host logic: block partition and the end-of-run gather, exercised with
world_size = 2 over gloo (no GPU).  The per-rank compute is replaced by a deterministic stand-in - only
the plumbing is under test here; the GPU path is covered by tests -m gpu and bench.py --gpus N."""
import os
import socket

import numpy as np
import pytest


def test_shard_bounds_partition():
    from geobipy_b200.parallel import shard_bounds
    for n in (0, 1, 7, 8, 4096, 262144, 65537):
        for world in (1, 2, 3, 8):
            blocks = [shard_bounds(n, r, world) for r in range(world)]
            assert blocks[0][0] == 0 and blocks[-1][1] == n
            assert all(blocks[i][1] == blocks[i + 1][0] for i in range(world - 1))
            sizes = [b - a for a, b in blocks]
            assert max(sizes) - min(sizes) <= 1
    assert shard_bounds(262144, 3, 8) == (3 * 32768, 4 * 32768)   # BASELINE config 3


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, n_total, q):
    import torch
    import torch.distributed as dist
    from geobipy_b200.parallel import gather_to_rank0, shard_bounds
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    lo, hi = shard_bounds(n_total, rank, world)
    idx = torch.arange(lo, hi)
    local = {
        "hitmap": (idx.view(-1, 1, 1) * torch.ones((1, 4, 5), dtype=torch.int64)).to(torch.int32),
        "scalars": torch.stack([idx.double(), idx.double() ** 2], dim=1),
        # per-system error histograms of a dual-moment (time-domain) datapoint: [block, n_systems, n_err_bins]
        "rel_hist": (idx.view(-1, 1, 1) + torch.arange(2).view(1, 2, 1) * 1000 + torch.zeros((1, 1, 99), dtype=torch.int64)).to(torch.int32),
        # height histogram of a sampled sensor height (solve_z): [block, n_err_bins]
        "height_hist": (idx.view(-1, 1) * 3 + torch.zeros((1, 99), dtype=torch.int64)).to(torch.int32),
    }
    out = gather_to_rank0(local, n_total)
    if rank == 0:
        q.put({k: v.numpy() for k, v in out.items()})
    else:
        assert out is None
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("n_total", [7, 8])
def test_gather_world_size_2_gloo(n_total):
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, n_total, q)) for r in range(2)]
    for p in procs:
        p.start()
    out = q.get(timeout=120)
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    assert out["hitmap"].shape == (n_total, 4, 5)
    assert np.array_equal(out["hitmap"][:, 0, 0], np.arange(n_total))
    assert np.array_equal(out["scalars"][:, 1], np.arange(n_total) ** 2.0)
    assert out["rel_hist"].shape == (n_total, 2, 99)
    assert np.array_equal(out["rel_hist"][:, 1, 7], np.arange(n_total) + 1000)
    assert out["height_hist"].shape == (n_total, 99) and np.array_equal(out["height_hist"][:, 50], 3 * np.arange(n_total))


def test_summarise_hitmap_matches_numpy():
    import torch
    from geobipy_b200.parallel import summarise_hitmap
    rng = np.random.default_rng(0)
    h = rng.integers(0, 50, (3, 40, 11)).astype(np.int32)
    edges = np.linspace(-5.0, 3.0, 41)
    out = summarise_hitmap(torch.tensor(h), torch.tensor(edges))
    c = 0.5 * (edges[1:] + edges[:-1])
    mean = (h * c[None, :, None]).sum(axis=1) / h.sum(axis=1)
    assert np.allclose(out["mean"].numpy(), mean)
    cs = np.cumsum(h, axis=1)
    for b in range(3):
        for j in range(11):
            i = np.searchsorted(cs[b, :, j], 0.5 * cs[b, -1, j])
            assert out["p50"][b, j].item() == c[min(i, 39)]
