"""One-process-per-GPU plumbing over torch.distributed (NCCL on GPUs, gloo in CPU tests).

The reference's communication backend is PySpark (sc.broadcast / parallelize / map /
collect: skdist/distribute/search.py:411-436).  Here: X and y are replicated on every
rank (one broadcast from rank 0 over NVLink when only rank 0 holds the data), the
independent (candidate x fold) columns / labels / trees are dealt round-robin to ranks
(no data-path collective), and fixed-size per-column results are all-gathered at the end.
"""
import os

import numpy as np


def dist_info():
    """(rank, world_size, local_rank); (0, 1, 0) when torch.distributed is not initialised."""
    # torch is only needed when a process group exists; importing it costs seconds, so a plain
    # single-process run (torch never imported by the caller) does not pay for it
    import sys
    if "torch" not in sys.modules:
        return 0, 1, 0
    try:
        import torch.distributed as dist
    except Exception:  # pragma: no cover
        return 0, 1, 0
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(), dist.get_world_size(), int(os.environ.get("LOCAL_RANK", 0))
    return 0, 1, 0


def shard_indices(n_items, rank, world):
    """Round-robin deal: item j -> rank j % world.  Candidate-major task
    order means every rank receives a mix of hyper-parameter values, which balances the
    per-column iteration counts."""
    return np.arange(rank, n_items, world, dtype=np.int64)


def all_gather_columns(local, n_items, rank, world):
    """Inverse of shard_indices for per-item result rows.

    local: array [len(shard_indices(n_items, rank, world)), ...].  Returns the full
    [n_items, ...] array on every rank."""
    local = np.ascontiguousarray(local)
    if world == 1:
        return local
    import torch
    import torch.distributed as dist

    per = (n_items + world - 1) // world
    tail = local.shape[1:]
    pad = np.zeros((per,) + tail, dtype=local.dtype)
    pad[: local.shape[0]] = local
    backend = dist.get_backend()
    dev = torch.device("cuda", torch.cuda.current_device()) if backend == "nccl" else torch.device("cpu")
    t = torch.from_numpy(pad).to(dev)
    outs = [torch.empty_like(t) for _ in range(world)]
    dist.all_gather(outs, t)
    full = np.zeros((n_items,) + tail, dtype=local.dtype)
    for r in range(world):
        idx = shard_indices(n_items, r, world)
        full[idx] = outs[r].cpu().numpy()[: len(idx)]
    return full


def _block_size(n_items, world, block):
    return max(1, min(int(block), n_items // max(1, world)))


def shard_blocks(n_items, rank, world, order=None, block=128):
    """Block-cyclic deal: the items, taken in `order` (default 0..n-1), are cut into blocks of
    `block` (shrunk for small problems so that every rank gets work) and block b goes to rank
    b % world.  The search passes a fold-major order, so a rank receives whole groups of 128
    same-fold columns -- the unit the tensor-core kernel works on -- instead of a thin slice of every
    fold that would have to be padded to 128 slots per fold."""
    order = np.arange(n_items, dtype=np.int64) if order is None else np.asarray(order, dtype=np.int64)
    if world == 1:
        return order
    b = _block_size(n_items, world, block)
    pos = np.arange(n_items, dtype=np.int64)
    return order[(pos // b) % world == rank]


def all_gather_blocks(local, n_items, rank, world, order=None, block=128):
    """Inverse of shard_blocks for per-item result rows: the full [n_items, ...] array on every rank."""
    local = np.ascontiguousarray(local)
    if world == 1:
        if order is None:
            return local
        full = np.zeros((n_items,) + local.shape[1:], dtype=local.dtype)
        full[np.asarray(order, dtype=np.int64)] = local
        return full
    import torch
    import torch.distributed as dist

    idx = [shard_blocks(n_items, r, world, order, block) for r in range(world)]
    per = max(len(i) for i in idx)
    tail = local.shape[1:]
    pad = np.zeros((per,) + tail, dtype=local.dtype)
    pad[: local.shape[0]] = local
    backend = dist.get_backend()
    dev = torch.device("cuda", torch.cuda.current_device()) if backend == "nccl" else torch.device("cpu")
    t = torch.from_numpy(pad).to(dev)
    outs = [torch.empty_like(t) for _ in range(world)]
    dist.all_gather(outs, t)
    full = np.zeros((n_items,) + tail, dtype=local.dtype)
    for r in range(world):
        full[idx[r]] = outs[r].cpu().numpy()[: len(idx[r])]
    return full


def broadcast_array(arr, shape, dtype, src=0):
    """Replicate a host array held by rank `src` on every rank (NCCL broadcast through
    device memory on GPUs).  Ranks other than src pass arr=None."""
    rank, world, _ = dist_info()
    if world == 1:
        return arr
    import torch
    import torch.distributed as dist

    backend = dist.get_backend()
    dev = torch.device("cuda", torch.cuda.current_device()) if backend == "nccl" else torch.device("cpu")
    if rank == src:
        t = torch.from_numpy(np.ascontiguousarray(arr, dtype=dtype)).to(dev)
    else:
        t = torch.empty(tuple(shape), dtype=getattr(torch, np.dtype(dtype).name), device=dev)
    dist.broadcast(t, src=src)
    return t.cpu().numpy()
