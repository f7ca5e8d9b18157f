"""Multi-GPU plumbing: queries shard across ranks, the index is replicated, and the only exchange of
the path is one all-gather of the fixed-size hit records at the end (NCCL over NVLink on GPUs; the same
code runs on gloo/CPU tensors for the host-logic tests)."""
from __future__ import annotations

import numpy as np

from ._abi import HIT_DT


def shard_range(n: int, rank: int, world: int):
    """contiguous block of original queries for `rank` (keeps all frames of a query together)"""
    return n * rank // world, n * (rank + 1) // world


def shard_queries(residues: np.ndarray, offsets: np.ndarray, rank: int, world: int):
    """-> (residues of the shard, offsets rebased to 0, index of the shard's first query)"""
    b, e = shard_range(len(offsets) - 1, rank, world)
    o = offsets[b:e + 1]
    return residues[int(o[0]):int(o[-1])], (o - o[0]).astype(np.uint64), b


_REC = HIT_DT.itemsize


def all_gather_hits(hits: np.ndarray, first_query: int = 0, device=None, to_host: bool = True):
    """All ranks contribute their hit records (q_id local to the shard; `first_query` rebases them)
    and receive everybody's: one all-gather of the counts, one of the padded records.
    Returns (gathered records in rank order or None, total count)."""
    import torch
    import torch.distributed as dist

    assert hits.dtype == HIT_DT
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        out = hits.copy()
        out["q_id"] += first_query
        return (out if to_host else None), len(out)
    world = dist.get_world_size()
    nccl = dist.get_backend() == "nccl"
    if device is None:
        device = torch.device("cuda", torch.cuda.current_device()) if nccl else torch.device("cpu")
    n = len(hits)
    cnt = torch.tensor([n], dtype=torch.int64, device=device)
    if nccl:
        cnts = torch.empty(world, dtype=torch.int64, device=device)
        dist.all_gather_into_tensor(cnts, cnt)
        counts = cnts.tolist()
    else:
        lst = [torch.zeros_like(cnt) for _ in range(world)]
        dist.all_gather(lst, cnt)
        counts = [int(c.item()) for c in lst]
    mx = max(max(counts), 1)
    # records travel as raw bytes; q_id (first int32 of a record) is rebased on the device
    buf = torch.empty(mx * _REC, dtype=torch.uint8, device=device)
    if n:
        buf[: n * _REC].copy_(torch.from_numpy(np.ascontiguousarray(hits).view(np.uint8).reshape(-1)))
        if first_query:
            buf.view(torch.int32).view(-1, _REC // 4)[:n, 0] += int(first_query)
    if nccl:
        flat = torch.empty(world * mx * _REC, dtype=torch.uint8, device=device)
        dist.all_gather_into_tensor(flat, buf)
        out = list(flat.view(world, mx * _REC).unbind(0))
    else:
        out = [torch.empty_like(buf) for _ in range(world)]
        dist.all_gather(out, buf)
    total = sum(counts)
    if not to_host:
        return None, total
    parts = [o[: c * _REC].cpu().numpy().view(HIT_DT) for o, c in zip(out, counts)]
    return np.concatenate(parts) if parts else np.zeros(0, HIT_DT), total
