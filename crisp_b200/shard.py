"""Region sharding of candidate sites across GPUs and the coordinate-order merge of their results.

The reference's only parallelism is "run several processes with different --regions" (README.md:51;
optionparser.c:31-35; bamsreader.c:147-173) and concatenate the VCFs.  Sites are independent, so the
B200 build shards a batch into contiguous genomic regions, one per rank, balanced by candidate count rather than
base pairs; there is no data-path collective.  Results come back to rank 0 in (tid, pos) order, which is the order
print_pooledvariant would have been called in by one process.
"""
from __future__ import annotations

import numpy as np

from .abi import _RESULT_FIELDS, Batch, Results


def coordinate_order(tid: np.ndarray, pos: np.ndarray) -> np.ndarray:
    return np.lexsort((pos, tid))


def partition_sites(tid: np.ndarray, pos: np.ndarray, world: int, weight: np.ndarray | None = None) -> list[np.ndarray]:
    """Split the sites into `world` contiguous genomic regions of (nearly) equal total weight.
    Returns site-index arrays, each sorted by coordinate; their concatenation is the coordinate order."""
    order = coordinate_order(tid, pos)
    n = order.size
    if weight is None:
        bounds = [(n * r) // world for r in range(world + 1)]
    else:
        w = np.asarray(weight, dtype=np.float64)[order]
        cum = np.concatenate([[0.0], np.cumsum(w)])
        targets = cum[-1] * np.arange(world + 1) / world
        bounds = np.searchsorted(cum, targets, side="left").tolist()
        bounds[0], bounds[-1] = 0, n
        for r in range(1, world + 1):
            bounds[r] = max(bounds[r], bounds[r - 1])
    return [order[bounds[r]:bounds[r + 1]] for r in range(world)]


def region_of(tid: np.ndarray, pos: np.ndarray, idx: np.ndarray) -> str:
    """The shard as a `--regions`-style string (first and last site), for logs."""
    if idx.size == 0:
        return "(empty)"
    a, b = idx[0], idx[-1]
    return f"{int(tid[a])}:{int(pos[a])}-{int(tid[b])}:{int(pos[b])}"


def shard_batch(batch: Batch, world: int, rank: int, by_reads: bool = True):
    """(sub-batch of rank, its site indices in the parent batch)."""
    weight = None
    if by_reads:
        P = batch.n_pools
        # the bytes a site costs: its reads (2 B each) or bins (4 B each), plus its per-pool counts
        if batch.read_off is not None:
            ro = batch.read_off.astype(np.int64)
            weight = 2 * (ro[P::P] - ro[:-1:P])
        else:
            bo = batch.bin_off.astype(np.int64)
            weight = 4 * (bo[P::P] - bo[:-1:P])
        if batch.ic_off is not None:
            io = batch.ic_off.astype(np.int64)
            weight = weight + 4 * (io[P::P] - io[:-1:P])
        else:
            weight = weight + 96 * P
    parts = partition_sites(batch.tid, batch.pos, world, weight)
    idx = parts[rank]
    return batch.select(idx), idx


def merge_results(parts: list[Results], indices: list[np.ndarray], n_sites: int, n_pools: int) -> Results:
    """Scatter per-rank results back to parent site order; AC/QV rows are renumbered rank by rank."""
    total_calls = sum(p.n_calls for p in parts)
    out = Results(n_sites, n_pools, max_calls=total_calls)
    out.n_calls = total_calls
    base = 0
    for res, idx in zip(parts, indices):
        for name, _, _ in _RESULT_FIELDS:
            getattr(out, name)[idx] = getattr(res, name)
        cr = res.call_row.copy()
        cr[cr >= 0] += base
        out.call_row[idx] = cr
        out.ac[base:base + res.n_calls] = res.ac[: res.n_calls]
        out.qv[base:base + res.n_calls] = res.qv[: res.n_calls]
        base += res.n_calls
    return out


def pack_results(res: Results, idx: np.ndarray) -> np.ndarray:
    """One flat byte buffer holding a rank's results and the parent indices of its sites (what travels to rank 0)."""
    idx = np.ascontiguousarray(idx, dtype=np.int64)
    head = np.array([res.n_sites, res.n_calls, res.n_pools, idx.size], dtype=np.int64)
    parts = [head.view(np.uint8), idx.view(np.uint8)]
    for name, dt, _ in _RESULT_FIELDS:
        parts.append(np.ascontiguousarray(getattr(res, name), dtype=dt).reshape(-1).view(np.uint8))
    parts.append(np.ascontiguousarray(res.ac[: res.n_calls], dtype=np.int32).reshape(-1).view(np.uint8))
    parts.append(np.ascontiguousarray(res.qv[: res.n_calls], dtype=np.float64).reshape(-1).view(np.uint8))
    return np.concatenate(parts)


def unpack_results(buf: np.ndarray):
    """Inverse of pack_results: (Results, idx)."""
    buf = np.ascontiguousarray(buf, dtype=np.uint8)
    n, n_calls, P, n_idx = (int(x) for x in buf[:32].view(np.int64))
    o = 32
    idx = buf[o:o + 8 * n_idx].view(np.int64).copy()
    o += 8 * n_idx
    res = Results(n, P, max_calls=n_calls)
    res.n_calls = n_calls
    for name, dt, w in _RESULT_FIELDS:
        nb = n * w * np.dtype(dt).itemsize
        getattr(res, name)[...] = buf[o:o + nb].view(dt).reshape(getattr(res, name).shape)
        o += nb
    nb = n_calls * P * 4
    res.ac[...] = buf[o:o + nb].view(np.int32).reshape(n_calls, P)
    o += nb
    res.qv[...] = buf[o:o + n_calls * P * 8].view(np.float64).reshape(n_calls, P)
    return res, idx


def gather_results(local: Results, idx: np.ndarray, n_sites: int, n_pools: int, group=None):
    """Collect every rank's results on rank 0 and merge them in coordinate order (control plane only — the data path has no
    collective).  Each rank's results travel as ONE flat byte tensor (two small collectives: the sizes, then the padded
    buffers), not as pickled Python objects.  Returns the merged Results on rank 0 and None elsewhere."""
    import torch
    import torch.distributed as dist

    world, rank = dist.get_world_size(group), dist.get_rank(group)
    dev = torch.device("cuda", torch.cuda.current_device()) if dist.get_backend(group) == "nccl" else torch.device("cpu")
    mine = pack_results(local, idx)
    sizes = [torch.zeros(1, dtype=torch.int64, device=dev) for _ in range(world)]
    dist.all_gather(sizes, torch.tensor([mine.size], dtype=torch.int64, device=dev), group=group)
    sizes = [int(x.item()) for x in sizes]
    cap = max(sizes)
    send = torch.zeros(cap, dtype=torch.uint8, device=dev)
    send[: mine.size] = torch.from_numpy(mine).to(dev)
    recv = [torch.empty(cap, dtype=torch.uint8, device=dev) for _ in range(world)] if rank == 0 else None
    dist.gather(send, recv, dst=0, group=group)
    if rank != 0:
        return None
    parts = [unpack_results(recv[r][: sizes[r]].cpu().numpy()) for r in range(world)]
    return merge_results([p[0] for p in parts], [p[1] for p in parts], n_sites, n_pools)
