"""Gene sharding across GPUs: one process per GPU, ``torch.distributed`` for the plumbing.

Genes are independent given theta, so each rank keeps a contiguous range of genes
resident on its GPU for the whole fit and the only per-evaluation exchange is an all-reduce of
``[value, grad[0..P-1], n_used, n_bad]`` (P+3 doubles).  Per-gene results are all-gathered once
per stage.  With no process group (the default), everything degenerates to one shard.

The collectives work on whatever device the process group serves (NCCL: the rank's GPU, where the
CUDA library writes its sums straight into the tensor being reduced; gloo: CPU tensors -- used by
the world_size-2 CPU tests with a stand-in local evaluator).
"""
import numpy as np


# Imported HERE, at module import, and not lazily inside a fit: something in torch's import keeps a
# reference to the importing frame, and with it the whole chain of calling frames -- a first fit that
# triggered the import from Session.__init__ had all its result arrays (and their pinned host blocks)
# kept alive by that chain until the cycle collector ran.
try:
    import torch.distributed as _torch_dist
except Exception:  # torch not importable: single process
    _torch_dist = None


def _dist():
    dist = _torch_dist
    if dist is not None and dist.is_available() and dist.is_initialized():
        return dist
    return None


def world():
    """(rank, world_size) of the default process group, (0, 1) without one."""
    dist = _dist()
    if dist is None:
        return 0, 1
    return dist.get_rank(), dist.get_world_size()


def shard_bounds(n, world_size):
    """Contiguous, balanced gene ranges: rank r owns [bounds[r], bounds[r+1])."""
    base, extra = divmod(int(n), int(world_size))
    sizes = [base + (1 if r < extra else 0) for r in range(world_size)]
    return np.concatenate([[0], np.cumsum(sizes)]).astype(np.int64)


def _backend_device():
    import torch
    dist = _dist()
    if dist is not None and dist.get_backend() == "nccl":
        return torch.device("cuda", torch.cuda.current_device())
    return torch.device("cpu")


def allreduce_sum(values):
    """Sum a small float64 vector over ranks; returns numpy.  Identity without a group."""
    dist = _dist()
    values = np.asarray(values, dtype=np.float64)
    if dist is None or dist.get_world_size() == 1:
        return values
    import torch
    t = torch.from_numpy(values.copy()).to(_backend_device())
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return t.cpu().numpy()


def allreduce_sum_tensor(tensor):
    """In-place sum of a tensor already on the backend's device (the CUDA hot path)."""
    dist = _dist()
    if dist is not None and dist.get_world_size() > 1:
        dist.all_reduce(tensor, op=dist.ReduceOp.SUM)
    return tensor


def allgather_rows(local, n_total):
    """Concatenate per-rank row blocks (shards from shard_bounds) into the full array on every rank."""
    dist = _dist()
    local = np.ascontiguousarray(local)
    if dist is None or dist.get_world_size() == 1:
        return local
    import torch
    ws = dist.get_world_size()
    bounds = shard_bounds(n_total, ws)
    sizes = np.diff(bounds)
    width = int(sizes.max())
    tail = local.shape[1:]
    padded = np.zeros((width,) + tail, dtype=local.dtype)
    padded[:local.shape[0]] = local
    dev = _backend_device()
    mine = torch.from_numpy(padded).to(dev)
    parts = [torch.empty_like(mine) for _ in range(ws)]
    dist.all_gather(parts, mine)
    out = np.empty((int(n_total),) + tail, dtype=local.dtype)
    for r in range(ws):
        out[bounds[r]:bounds[r + 1]] = parts[r].cpu().numpy()[:sizes[r]]
    return out
