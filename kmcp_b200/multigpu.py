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


def gather_hits_padded(hits: np.ndarray, rank: int, world: int, device, group=None, dst: int = 0, target_base: int = 0):
    """the same gather as one padded collective on `device` (NCCL over NVLink when device is this rank's GPU; works on
    a gloo group with device "cpu" too): sizes by all_gather, then one gather of buffers padded to the largest list.
    `target_base` is added to the target column on the way (shards that number their targets locally).
    Rank `dst` gets the concatenation in rank order, the others None."""
    h32 = np.ascontiguousarray(hits).view(np.uint32).reshape(-1, 3)
    n = torch.tensor([h32.shape[0]], dtype=torch.int64, device=device)
    sizes = [torch.zeros_like(n) for _ in range(world)]
    dist.all_gather(sizes, n, group=group)
    sizes = [int(x.item()) for x in sizes]
    mx = max(max(sizes), 1)
    buf = torch.zeros((mx, 3), dtype=torch.int32, device=device)
    if h32.shape[0]:
        buf[:h32.shape[0]].copy_(torch.from_numpy(h32.view(np.int32)), non_blocking=True)
        if target_base:
            buf[:h32.shape[0], 1] += int(np.int32(np.uint32(target_base)))
    if rank == dst:
        outs = [torch.empty_like(buf) for _ in range(world)]
        dist.gather(buf, outs, dst=dst, group=group)
        allb = torch.cat([o[:sizes[i]] for i, o in enumerate(outs)]).cpu().numpy()
        return np.ascontiguousarray(allb).view(np.uint32).reshape(-1).view(hits.dtype)
    dist.gather(buf, None, dst=dst, group=group)
    return None
