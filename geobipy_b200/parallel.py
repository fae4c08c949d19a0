"""Sharding of a survey's soundings over the GPUs of one node and the end-of-run collation.

Replaces the reference's MPI master/worker queue (geobipy/src/inversion/Inference3D.py:518-635): chains
are independent, so ranks take contiguous blocks of the sounding index range and never talk during
sampling; the only collective is one gather of the posterior arrays to rank 0 at the end
(``torch.distributed``: NCCL over NVLink on GPUs, gloo in the CPU tests).
"""
import numpy as np

__all__ = ["shard_bounds", "Collator", "gather_to_rank0", "summarise_hitmap", "run_sharded"]


def shard_bounds(n, rank, world):
    """Contiguous block [lo, hi) of rank ``rank`` when n items are split over ``world`` ranks; block
    sizes differ by at most one (first n % world ranks get the extra item)."""
    base, extra = divmod(int(n), int(world))
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


class Collator:
    """The end-of-run collation: ONE collective for all result arrays of a rank.

    Every per-sounding array of the block is packed into one contiguous byte row per sounding ([block, row_bytes] uint8,
    fields 8-byte aligned), the rows of all ranks land in one buffer preallocated on rank 0 ([world, max_block,
    row_bytes]) with a single ``dist.gather`` (NCCL over NVLink / NVSwitch on GPUs, gloo in the CPU tests), and rank 0
    reads the arrays back as typed views of that buffer.  ``warm_up()`` runs the collective once on the real buffers so
    that the lazily created NCCL point-to-point channels exist before anything is timed (the first gather of a process
    otherwise pays ~0.2-1.4 s of channel set-up: SCALE_r01.json)."""

    def __init__(self, example, n_total, group=None):
        import torch
        import torch.distributed as dist
        self.group, self.n_total = group, int(n_total)
        self.active = dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1
        self.world = dist.get_world_size(group) if self.active else 1
        self.rank = dist.get_rank(group) if self.active else 0
        self.blocks = [shard_bounds(n_total, r, self.world) for r in range(self.world)]
        self.max_block = max(hi - lo for lo, hi in self.blocks)
        self.fields, off = [], 0
        for name in sorted(example):
            t = example[name]
            nbytes = int(np.prod(t.shape[1:], dtype=np.int64)) * t.element_size()
            self.fields.append((name, off, nbytes, t.dtype, tuple(t.shape[1:])))
            off += (nbytes + 7) & ~7
        self.row_bytes = off
        dev = next(iter(example.values())).device
        self.send = torch.zeros((self.max_block, self.row_bytes), dtype=torch.uint8, device=dev)
        self.recv = (torch.zeros((self.world, self.max_block, self.row_bytes), dtype=torch.uint8, device=dev)
                     if (self.active and self.rank == 0) else None)

    @property
    def bytes_per_rank(self):
        return self.max_block * self.row_bytes

    def _pack(self, local):
        import torch
        for name, off, nbytes, dtype, shape in self.fields:
            t = local[name].contiguous()
            n = t.shape[0]
            self.send[:n, off:off + nbytes] = t.reshape(n, -1).view(torch.uint8)

    def _collective(self):
        import torch.distributed as dist
        dist.gather(self.send, list(self.recv.unbind(0)) if self.rank == 0 else None, dst=0, group=self.group)

    def warm_up(self):
        if self.active:
            self._collective()

    def gather(self, local):
        """Full arrays (typed views of the receive buffer, block order = sounding order) on rank 0, None elsewhere."""
        import torch
        if not self.active:
            return local
        self._pack(local)
        self._collective()
        if self.rank != 0:
            return None
        out = {}
        even = all(hi - lo == self.max_block for lo, hi in self.blocks)
        for name, off, nbytes, dtype, shape in self.fields:
            if even:  # one strided view: [world * block, nbytes] -> typed
                v = self.recv[:, :, off:off + nbytes].reshape(self.world * self.max_block, nbytes)
            else:
                v = torch.cat([self.recv[r, :hi - lo, off:off + nbytes] for r, (lo, hi) in enumerate(self.blocks)], dim=0)
            out[name] = v.contiguous().view(dtype).reshape((v.shape[0],) + shape)
        return out


def gather_to_rank0(local, n_total, group=None):
    """Gather a dict of per-sounding tensors (leading dimension = local block) to rank 0 with one collective
    (``Collator``).  Returns the dict of full arrays on rank 0, None elsewhere.  Works without torch.distributed
    initialised (single process): returns ``local``.  Callers that gather repeatedly keep a ``Collator``."""
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return local
    return Collator(local, n_total, group).gather(local)


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
        if isinstance(sigma_edges_ln, np.ndarray):   # host edges: no device round trip (the caller's stream keeps running)
            eh = np.asarray(sigma_edges_ln, dtype=np.float64)
            dx = float((eh[..., -1] - eh[..., 0]).reshape(-1)[0]) / (eh.shape[-1] - 1)
            lo = torch.as_tensor(np.ascontiguousarray(eh[..., 0]).reshape(-1), device=hitmap.device)
            lo = lo if eh.ndim == 2 else lo.expand(hitmap.shape[0])
        else:
            e = sigma_edges_ln.to(hitmap.device, torch.float64)
            lo = e[..., 0] if e.dim() == 2 else e[0].expand(hitmap.shape[0])
            dx = float((e[..., -1] - e[..., 0]).reshape(-1)[0]) / (e.shape[-1] - 1)
        mean, pct = ops.summarise_hitmap(hitmap.to(torch.int32).contiguous(), lo.contiguous(), dx, percentiles)
        out = {"mean": mean}
        for i, p in enumerate(percentiles):
            out["p%g" % p] = pct[i]
        return out
    h = hitmap.to(torch.float64)
    sigma_edges_ln = torch.as_tensor(sigma_edges_ln).to(torch.float64)
    centres = 0.5 * (sigma_edges_ln[..., 1:] + sigma_edges_ln[..., :-1])
    if centres.dim() == 1:
        centres = centres.unsqueeze(0).expand(h.shape[0], -1)
    tot = h.sum(dim=1)
    out = {"mean": (h * centres.unsqueeze(2)).sum(dim=1) / tot.clamp_min(1.0)}
    cs = torch.cumsum(h, dim=1)
    frac = torch.where(tot.unsqueeze(1) > 0, cs / tot.unsqueeze(1).clamp_min(1.0), torch.zeros_like(cs))
    for p in percentiles:   # first bin whose cumulative fraction reaches p * 0.01 (fp64 both: the reference's rule, ties included)
        idx = (frac < p * 0.01).sum(dim=1).clamp_max(h.shape[1] - 1)
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
