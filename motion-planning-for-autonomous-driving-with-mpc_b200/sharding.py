"""Multi-GPU batch sharding: one process per GPU (torch.distributed), contiguous split of the batch index.

Every ego instance is an independent NLP (the reference's loop has no cross-instance state, optimizer.py:596), so the
data path needs NO collective: each rank solves `shard_range(B, rank, world)`.  Collectives are used only at the edges:
`broadcast_scenario` (constants, once) and `gather_solutions` (results to every rank / rank 0).  With backend "nccl"
these run over NVLink/NVSwitch; the CPU tests run the same code with "gloo".
"""
import numpy as np


def shard_range(B, rank, world):
    """Contiguous, balanced split: the first B % world ranks get one extra instance."""
    q, r = divmod(int(B), int(world))
    lo = rank * q + min(rank, r)
    return lo, lo + q + (1 if rank < r else 0)


def shard_sizes(B, world):
    return [shard_range(B, r, world)[1] - shard_range(B, r, world)[0] for r in range(world)]


def broadcast_scenario(arrays, src=0, device=None):
    """Broadcast a dict of float64 numpy arrays (path, orientation, weights ...) from `src` to all ranks."""
    import torch
    import torch.distributed as dist
    out = {}
    for k in sorted(arrays.keys()):
        t = torch.as_tensor(np.ascontiguousarray(arrays[k], np.float64))
        if device is not None:
            t = t.to(device)
        dist.broadcast(t, src=src)
        out[k] = t.cpu().numpy()
    return out


def gather_solutions(local, B, device=None):
    """all_gather variable-length shards: `local` is a tensor [b_local, ...]; returns [B, ...] on every rank."""
    import torch
    import torch.distributed as dist
    world = dist.get_world_size()
    sizes = shard_sizes(B, world)
    mx = max(sizes)
    pad = torch.zeros((mx,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    pad[: local.shape[0]] = local
    bufs = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(bufs, pad)
    return torch.cat([bufs[r][: sizes[r]] for r in range(world)], dim=0)


def solve_sharded(solve_fn, xref, X_init, U_init, gather=True):
    """Split [B,...] inputs by rank, call `solve_fn(xref_l, X_l, U_l) -> (U, X, status, iters)` on the local shard and
    (optionally) all_gather the solutions.  `solve_fn` is `B200Optimizer.solve_batch` on a GPU rank."""
    import torch.distributed as dist
    rank, world = dist.get_rank(), dist.get_world_size()
    B = xref.shape[0]
    lo, hi = shard_range(B, rank, world)
    U, X, status, iters = solve_fn(xref[lo:hi], X_init[lo:hi], U_init[lo:hi])
    if not gather:
        return U, X, status, iters
    return (gather_solutions(U, B), gather_solutions(X, B), gather_solutions(status, B), gather_solutions(iters, B))
