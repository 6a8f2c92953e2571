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


def gather_results(local: Results, idx: np.ndarray, n_sites: int, n_pools: int, group=None):
    """Collect every rank's results on rank 0 (torch.distributed object gather; control plane only — the
    data path has no collective).  Returns the merged Results on rank 0 and None elsewhere."""
    import torch.distributed as dist

    world, rank = dist.get_world_size(group), dist.get_rank(group)
    payload = (local, np.asarray(idx))
    gathered = [None] * world if rank == 0 else None
    dist.gather_object(payload, gathered, dst=0, group=group)
    if rank != 0:
        return None
    return merge_results([g[0] for g in gathered], [g[1] for g in gathered], n_sites, n_pools)
