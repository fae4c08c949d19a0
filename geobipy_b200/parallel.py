"""Sharding of a survey's soundings over the GPUs of one node and the end-of-run collation.

Replaces the reference's MPI master/worker queue (geobipy/src/inversion/Inference3D.py:518-635): chains
are independent, so ranks take contiguous blocks of the sounding index range and never talk during
sampling; the only collective is one gather of the posterior arrays to rank 0 at the end
(``torch.distributed``: NCCL over NVLink on GPUs, gloo in the CPU tests).
"""
import numpy as np

__all__ = ["shard_bounds", "gather_to_rank0", "summarise_hitmap", "run_sharded"]


def shard_bounds(n, rank, world):
    """Contiguous block [lo, hi) of rank ``rank`` when n items are split over ``world`` ranks; block
    sizes differ by at most one (first n % world ranks get the extra item)."""
    base, extra = divmod(int(n), int(world))
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def gather_to_rank0(local, n_total, group=None):
    """Gather a dict of per-sounding tensors (leading dimension = local block) to rank 0.

    Every rank passes its block; blocks are padded to the largest block so one fixed-size
    ``dist.gather`` per array is enough.  Returns the dict of full arrays on rank 0, None elsewhere.
    Works without torch.distributed initialised (single process): returns ``local``.
    """
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return local
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    max_block = max(shard_bounds(n_total, r, world)[1] - shard_bounds(n_total, r, world)[0] for r in range(world))
    out = {} if rank == 0 else None
    for name in sorted(local):
        t = local[name]
        pad = max_block - t.shape[0]
        if pad:
            t = torch.cat([t, torch.zeros((pad,) + tuple(t.shape[1:]), dtype=t.dtype, device=t.device)], dim=0)
        t = t.contiguous()
        bufs = [torch.empty_like(t) for _ in range(world)] if rank == 0 else None
        dist.gather(t, bufs, dst=0, group=group)
        if rank == 0:
            parts = []
            for r in range(world):
                lo, hi = shard_bounds(n_total, r, world)
                parts.append(bufs[r][: hi - lo])
            out[name] = torch.cat(parts, dim=0)
    return out


def summarise_hitmap(hitmap, sigma_edges_ln, percentiles=(5.0, 50.0, 95.0)):
    """Per-depth-cell posterior summaries of hitmaps [B, n_sigma, n_depth] (torch, any device).

    Mean and percentiles of ln(sigma), following Mesh._mean / Mesh._percentile
    (geobipy/src/classes/mesh/Mesh.py:80, :173-217): the percentile is the centre of the first bin whose
    cumulative count reaches p % of the column total.  Returns dict(mean [B, n_depth], p<q> [B, n_depth]).
    Shrinks the end-of-run gather from 440 KB to ~7 KB per sounding.
    """
    import torch
    if hitmap.is_cuda:
        # hand-written kernel (gbp_summarise_hitmap): one coalesced pass over the hitmaps instead of torch's
        # cumsum / gather chain (which moves ~8x the bytes)
        from . import ops
        # sigma_edges_ln: [n_sig + 1] (shared) or [B, n_sig + 1] (per sounding: the bins are centred on each sounding's
        # best half-space); uniform bins of one common width either way (Model.set_posteriors, Model.py:666-684)
        e = sigma_edges_ln.to(hitmap.device, torch.float64)
        lo = e[..., 0] if e.dim() == 2 else e[0].expand(hitmap.shape[0])
        dx = float((e[..., -1] - e[..., 0]).reshape(-1)[0]) / (e.shape[-1] - 1)
        mean, pct = ops.summarise_hitmap(hitmap.to(torch.int32).contiguous(), lo.contiguous(), dx, percentiles)
        out = {"mean": mean}
        for i, p in enumerate(percentiles):
            out["p%g" % p] = pct[i]
        return out
    h = hitmap.to(torch.float64)
    sigma_edges_ln = sigma_edges_ln.to(torch.float64)
    centres = 0.5 * (sigma_edges_ln[..., 1:] + sigma_edges_ln[..., :-1])
    if centres.dim() == 1:
        centres = centres.unsqueeze(0).expand(h.shape[0], -1)
    tot = h.sum(dim=1).clamp_min(1.0)
    out = {"mean": (h * centres.unsqueeze(2)).sum(dim=1) / tot}
    cs = torch.cumsum(h, dim=1)
    for p in percentiles:
        target = (p / 100.0) * tot
        idx = (cs < target.unsqueeze(1)).sum(dim=1).clamp_max(h.shape[1] - 1)
        out["p%g" % p] = torch.gather(centres, 1, idx)
    return out


def run_sharded(system, opt, data, altitude, seed=0, precision=32, outputs=("hitmap", "edges_hist", "ncells_hist", "scalars"),
                max_iterations=0, gather=True):
    """One process per GPU: run this rank's block of soundings, then gather to rank 0.

    ``data`` [n_total, 2F] and ``altitude`` [n_total] are the full survey (numpy, tiny); each rank uploads
    only its block.  Sounding ``i`` always uses random stream ``(seed, i)`` whatever the number of ranks, so
    results do not depend on the sharding.
    """
    import torch
    import torch.distributed as dist
    from . import ops
    rank = dist.get_rank() if dist.is_available() and dist.is_initialized() else 0
    world = dist.get_world_size() if dist.is_available() and dist.is_initialized() else 1
    n_total = int(np.shape(data)[0])
    lo, hi = shard_bounds(n_total, rank, world)
    dev = torch.device("cuda", torch.cuda.current_device())
    d = torch.as_tensor(np.ascontiguousarray(data[lo:hi]), dtype=torch.float64).to(dev)
    a = torch.as_tensor(np.ascontiguousarray(altitude[lo:hi]), dtype=torch.float64).to(dev)
    res = ops.rjmcmc_run(system, opt, d, a, seed=seed, first_index=lo, max_iterations=max_iterations,
                         precision=precision, outputs=outputs)
    if not gather:
        return res
    return gather_to_rank0(res, n_total)
