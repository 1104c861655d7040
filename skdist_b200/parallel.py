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


def logreg_column_cost(C):
    """Relative number of lock-step rounds a logistic column is expected to need: weakly regularised
    fits (large C) run to max_iter, strongly regularised ones stop early.  Only the ordering and the
    rough ratio matter (block dealing below); 1.0 = runs to the end."""
    C = np.asarray(C, dtype=np.float64)
    return np.clip(0.15 + 0.25 * (np.log10(np.maximum(C, 1e-300)) + 4.0), 0.15, 1.0)


def _assign_blocks(block_cost, world):
    """Rank of every block.  Without costs: block b -> rank b % world.  With costs: longest
    processing time first -- blocks in descending cost (ties: lower index) to the rank with the
    least load (ties: fewer blocks, lower rank), so the ranks that must take one block more than
    the others get the blocks that finish early."""
    nb = len(block_cost)
    if world == 1:
        return np.zeros(nb, dtype=np.int64)
    if np.all(block_cost == block_cost[0]):
        return np.arange(nb, dtype=np.int64) % world
    out = np.empty(nb, dtype=np.int64)
    load = np.zeros(world)
    cnt = np.zeros(world, dtype=np.int64)
    for bi in np.argsort(-block_cost, kind="stable"):
        r = min(range(world), key=lambda q: (round(load[q], 9), cnt[q], q))
        out[bi] = r
        load[r] += block_cost[bi]
        cnt[r] += 1
    return out


def shard_blocks(n_items, rank, world, order=None, block=128, cost=None):
    """Block deal: the items, taken in `order` (default 0..n-1), are cut into blocks of `block`
    (shrunk for small problems so that every rank gets work); block b goes to rank b % world, or,
    when per-item costs are given (indexed by item id; a block costs as much as its most expensive
    item because its columns advance in lock-step), to the rank `_assign_blocks` picks.  The search
    passes a fold-major order, so a rank receives whole groups of 128 same-fold columns -- the unit
    the tensor-core kernel works on -- instead of a thin slice of every fold that would have to be
    padded to 128 slots per fold."""
    order = np.arange(n_items, dtype=np.int64) if order is None else np.asarray(order, dtype=np.int64)
    if world == 1:
        return order
    b = _block_size(n_items, world, block)
    pos = np.arange(n_items, dtype=np.int64)
    blk = pos // b
    if cost is None:
        return order[blk % world == rank]
    cost = np.asarray(cost, dtype=np.float64)
    nb = int(blk[-1]) + 1 if n_items else 0
    bcost = np.zeros(nb)
    np.maximum.at(bcost, blk, cost[order])
    return order[_assign_blocks(bcost, world)[blk] == rank]


def all_gather_blocks(local, n_items, rank, world, order=None, block=128, cost=None):
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

    idx = [shard_blocks(n_items, r, world, order, block, cost) for r in range(world)]
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


def _host_memory_available():
    try:
        import psutil
        return int(psutil.virtual_memory().available)
    except Exception:  # pragma: no cover
        return None


def all_gather_trees(local, n_items, rank, world, rebuild, mode=None, piece_bytes=256 << 20):
    """Fitted trees of a forest dealt by `shard_indices`, back on the ranks: streamed, never pickled as a whole.

    The reference collects the trees on the Spark driver (ensemble.py:319: `.collect()`).  In the SPMD form
    every rank ends with the complete forest -- but a config-4 forest (1024 trees of 372 k nodes) is 38 GB of
    node records, and one `all_gather_object` of the local list held the list, its pickle, the byte tensor,
    every other rank's bytes and the unpickled copies at once (~7x the forest per rank: enough to take a
    host down).  Here a tree travels as its raw `nodes` / `values` arrays in pieces of about `piece_bytes`
    per rank (one fixed-size all-gather per piece, sizes agreed on up front), and the receiving side rebuilds
    the estimator with `rebuild(item_index, max_depth, nodes, values)`; peak temporary memory is
    world x piece_bytes.

    mode (default: env SKDIST_B200_FOREST_GATHER or "auto"): "all" = every rank gets every tree;
    "rank0" = only rank 0 does (the reference's driver), the other ranks return their own trees and None for
    the rest; "auto" = "all" unless world copies of the forest would not fit in the host memory that is
    available (one node), then "rank0" with a warning.  Returns the list of n_items estimators."""
    if world == 1:
        return list(local)
    import torch
    import torch.distributed as dist

    mode = mode or os.environ.get("SKDIST_B200_FOREST_GATHER", "auto")
    if mode not in ("all", "rank0", "auto"):
        raise ValueError("SKDIST_B200_FOREST_GATHER must be all, rank0 or auto")
    backend = dist.get_backend()
    dev = torch.device("cuda", torch.cuda.current_device()) if backend == "nccl" else torch.device("cpu")
    idx = [shard_indices(n_items, r, world) for r in range(world)]
    per = max(len(i) for i in idx)
    # sizes of every tree of every rank: [node_count, nodes bytes, values bytes, max_depth]
    meta = np.zeros((per, 4), dtype=np.int64)
    states = []
    for j, est in enumerate(local):
        st = est.tree_.__getstate__()
        nodes = np.ascontiguousarray(st["nodes"])
        values = np.ascontiguousarray(st["values"])
        states.append((nodes, values))
        meta[j] = (st["node_count"], nodes.nbytes, values.nbytes, st["max_depth"])
    mt = torch.from_numpy(meta).to(dev)
    mall = [torch.empty_like(mt) for _ in range(world)]
    dist.all_gather(mall, mt)
    mall = [m.cpu().numpy() for m in mall]
    total = int(sum(int(m[:, 1:3].sum()) for m in mall))
    if mode == "auto":
        avail = _host_memory_available()
        # every rank of this node would hold the other ranks' trees on top of its own
        need = sum(total - int(mall[r][:, 1:3].sum()) for r in range(world))
        mode = "all"
        if avail is not None and need > 0.6 * avail:
            mode = "rank0"
            if rank == 0:
                import warnings
                warnings.warn("forest of %.1f GB: %d copies do not fit in the %.1f GB of host memory available; "
                              "only rank 0 collects every tree (SKDIST_B200_FOREST_GATHER=all overrides)"
                              % (total / 1e9, world, avail / 1e9))
        flag = torch.tensor([1 if mode == "rank0" else 0], dtype=torch.int32, device=dev)
        dist.all_reduce(flag, op=dist.ReduceOp.MAX)      # the ranks must agree (they sample free memory at different times)
        mode = "rank0" if int(flag.item()) else "all"
    node_dtype = states[0][0].dtype if states else None      # scikit-learn's NODE_DTYPE: the same on every rank
    out = [None] * n_items
    for j, est in enumerate(local):
        out[idx[rank][j]] = est
    keep = mode == "all" or rank == 0
    j0 = 0
    while j0 < per:
        # trees j0 .. j1-1 of every rank in one piece: as many as fit in piece_bytes on the fullest rank
        j1, size = j0, np.zeros(world, dtype=np.int64)
        while j1 < per:
            nxt = np.array([int(m[j1, 1] + m[j1, 2]) for m in mall], dtype=np.int64)
            if j1 > j0 and (size + nxt).max() > piece_bytes:
                break
            size += nxt
            j1 += 1
        width = int(size.max())
        buf = np.zeros(max(width, 1), dtype=np.uint8)
        o = 0
        for j in range(j0, min(j1, len(local))):
            for a in states[j]:
                b = a.view(np.uint8).reshape(-1)
                buf[o:o + b.size] = b
                o += b.size
            states[j] = None                  # the local copy lives on in local[j].tree_
        t = torch.from_numpy(buf).to(dev)
        outs = [torch.empty_like(t) for _ in range(world)]
        dist.all_gather(outs, t)
        if keep:
            for r in range(world):
                if r == rank:
                    continue
                got = outs[r].cpu().numpy()
                o = 0
                for j in range(j0, min(j1, len(idx[r]))):
                    cnt, nb, vb, depth = (int(v) for v in mall[r][j])
                    nd = _node_dtype(node_dtype)
                    nodes = got[o:o + nb].view(nd).copy()
                    o += nb
                    values = got[o:o + vb].view(np.float64).copy()
                    o += vb
                    out[idx[r][j]] = rebuild(int(idx[r][j]), depth, nodes, values.reshape(cnt, -1))
        del outs, t
        j0 = j1
    return out


def _node_dtype(local_dtype):
    if local_dtype is not None:
        return local_dtype
    from sklearn.tree._tree import NODE_DTYPE      # a rank without trees of its own
    return NODE_DTYPE


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


class _DeviceView:
    """Zero-copy torch view of a raw device buffer (CUDA array interface)."""

    def __init__(self, ptr, shape, typestr="<f4"):
        self.__cuda_array_interface__ = {"shape": tuple(shape), "typestr": typestr, "data": (int(ptr), False),
                                         "strides": None, "version": 3}


def stage_x_replicated(eng, X):
    """Stage X in the HBM of every rank.  One process: a host-to-device copy.  Several ranks over
    NCCL: rank 0 copies its host array to its GPU and the other ranks receive the staged matrix
    over NVLink in ONE broadcast (the reference's sc.broadcast(X), search.py:414-421) -- N
    concurrent host-to-device copies of the same matrix would share the host's memory bandwidth.
    Every rank must call this with the same X (SPMD), as for the sharded fits themselves."""
    rank, world, local = dist_info()
    if world == 1:
        return eng.stage_x(X)
    import torch
    import torch.distributed as dist
    if dist.get_backend() != "nccl" or not hasattr(eng, "staged_x"):
        return eng.stage_x(X)
    # this may run on a worker thread (the search overlaps staging with the cv split): torch's
    # current device is per thread, so name the rank's GPU explicitly
    with torch.cuda.device(local):
        if hasattr(eng, "stage_x_sliced") and os.environ.get("SKDIST_B200_STAGE", "sliced") == "sliced":
            return _stage_x_allgather(eng, X, rank, world, torch.device("cuda", local))
        return _stage_x_broadcast(eng, X, rank, torch.device("cuda", local))


def _stage_x_allgather(eng, X, rank, world, dev):
    """Every rank holds X on the host (SPMD): rank r copies rows [r * per, (r + 1) * per) to its GPU
    through its own PCIe link (1/N of the bytes each, concurrently) and ONE in-place NCCL all-gather
    over NVLink completes the matrix everywhere -- the reference's sc.broadcast(X) (search.py:414-421)
    without the single host-to-device copy of the whole matrix in front of it."""
    import torch
    import torch.distributed as dist
    n, d = np.asarray(X).shape
    per = (n + world - 1) // world
    row0, row1 = min(n, rank * per), min(n, (rank + 1) * per)

    def gather(ptr, ldx):
        full = torch.as_tensor(_DeviceView(ptr, (per * world, ldx)), device=dev)
        dist.all_gather_into_tensor(full, full[rank * per:(rank + 1) * per])
        torch.cuda.current_stream().synchronize()

    status = torch.zeros(1, dtype=torch.int32, device=dev)
    err = None
    try:
        eng.stage_x_sliced(X, row0, row1, per * world, gather)
    except Exception as e:      # noqa: BLE001 - a NaN anywhere in X is seen by every rank's commit; other errors are agreed on below
        err = e
        status[0] = 1
    dist.all_reduce(status, op=dist.ReduceOp.MAX)
    if err is not None:
        raise err
    if int(status.item()):
        raise ValueError("another rank could not stage X (see its error)")


def _stage_x_broadcast(eng, X, rank, dev):
    import torch
    import torch.distributed as dist
    X = np.asarray(X)
    n, d = X.shape
    # header: [status, ldx]; status != 0 -> rank 0 could not stage (e.g. NaN in X): everybody raises
    head = torch.zeros(2, dtype=torch.int64, device=dev)
    err = None
    if rank == 0:
        try:
            eng.stage_x(X)
            ptr, _, _, ldx = eng.staged_x()
            head[1] = ldx
        except Exception as e:      # noqa: BLE001 - re-raised below on every rank
            err = e
            head[0] = 1
    dist.broadcast(head, src=0)
    status, ldx = (int(v) for v in head.cpu().tolist())
    if status:
        if err is not None:
            raise err
        raise ValueError("rank 0 could not stage X (see its error)")
    if rank == 0:
        t = torch.as_tensor(_DeviceView(ptr, (n, ldx)), device=dev)
        dist.broadcast(t, src=0)
        torch.cuda.current_stream().synchronize()
    else:
        t = torch.empty((n, ldx), dtype=torch.float32, device=dev)
        dist.broadcast(t, src=0)
        torch.cuda.current_stream().synchronize()
        eng.stage_x_device(t.data_ptr(), n, d, ldx)
        del t
