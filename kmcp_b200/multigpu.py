"""Host-side plumbing of the multi-GPU search (one process per GPU, torch.distributed).

The path shards by index BLOCK: every rank holds the blocks kmcpg_shard_plan gives it, the
read batch is broadcast (NCCL over NVLink — the only collective on the data path), every rank probes the whole
batch against its blocks, and the per-rank hit lists — disjoint by target — are concatenated on the host of
rank 0 (what `kmcp merge` does across processes, merge.go:190-256).
"""
from __future__ import annotations

import numpy as np
import torch
import torch.distributed as dist


def broadcast_batch(dev_buf: torch.Tensor, src: int = 0) -> torch.Tensor:
    """the read batch (uint8 tensor already on this rank's GPU) → every rank"""
    dist.broadcast(dev_buf, src=src)
    return dev_buf


def gather_hits(hits: np.ndarray, rank: int, world: int, group=None, dst: int = 0, order: str = "query_target"):
    """variable-size gather of 12-byte hit records on a CPU (gloo) group; rank `dst` gets all records, the others
    get None.  order: "query_target" = canonical (query, target) order; "query" = stable by query only (each
    shard's list is already sorted, shards are disjoint by target); "none" = plain concatenation"""
    raw = torch.from_numpy(np.ascontiguousarray(hits).view(np.uint8).copy())
    cnt = torch.tensor([raw.numel()], dtype=torch.int64)
    counts = [torch.zeros(1, dtype=torch.int64) for _ in range(world)]
    dist.all_gather(counts, cnt, group=group)
    if rank == dst:
        parts = [raw]
        for src in range(world):
            if src == dst:
                continue
            buf = torch.empty(int(counts[src].item()), dtype=torch.uint8)
            if buf.numel():
                dist.recv(buf, src=src, group=group)
            parts.append(buf)
        allb = torch.cat(parts).numpy()
        out = allb.view(hits.dtype)
        if order == "none":
            return out
        if order == "query":
            return out[np.argsort(out["query"], kind="stable")]
        return out[np.lexsort((out["target"], out["query"]))]
    if raw.numel():
        dist.send(raw, dst=dst, group=group)
    return None
