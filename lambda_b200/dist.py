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


class DeviceHitGather:
    """The path's only exchange, GPU to GPU: every rank contributes the records of its last search straight from the
    library's device buffer (lgpu_ctx_export_hits) and receives everybody's with ONE fixed-capacity all-gather (NCCL
    over NVLink) -- the record count travels in the first word of each rank's slot, so there is no separate size
    exchange and no host synchronisation.  The collective is issued on a side stream: the next search call starts
    while it runs.  A rank with more records than the capacity only sends its count; every rank sees that in
    finish(), all grow their slots to the same size and the exchange is repeated (rare: size the capacity with
    max_matches records per query and it cannot happen).

        g = DeviceHitGather(cap_records)          # per-rank capacity
        g.start(searcher, first_query)            # after searcher.search(); returns immediately
        ...next search...
        total = g.finish()                        # counts of all ranks (host), records stay on the device
        hits = g.records(searcher)                # optional: all records on the host, bit score / e-value filled in

    `searcher` needs export_hits(dev_ptr, cap_records, first_query) -> n (lambda_b200.Searcher).  On CPU tensors with
    the gloo backend the same code runs without streams (host-logic tests)."""
    HEADER = 16  # bytes in front of the records of a slot (count as int64 + padding; keeps the doubles aligned)

    def __init__(self, cap_records: int, device=None):
        import torch
        import torch.distributed as dist
        self.torch, self.dist = torch, dist
        self.world = dist.get_world_size() if (dist.is_available() and dist.is_initialized()) else 1
        nccl = self.world > 1 and dist.get_backend() == "nccl"
        if device is None:
            device = torch.device("cuda", torch.cuda.current_device()) if (nccl or torch.cuda.is_available()) else torch.device("cpu")
        self.device = device
        self.cuda = device.type == "cuda"
        if self.cuda:
            self.stream = torch.cuda.Stream(device=device)
            self.ev0 = torch.cuda.Event(enable_timing=True)
            self.ev1 = torch.cuda.Event(enable_timing=True)
        self._alloc(max(int(cap_records), 1))
        self.pending = False
        self.last_ms = 0.0
        self.regrown = 0

    def _alloc(self, cap):
        t = self.torch
        self.cap = int(cap)
        self.slot = self.HEADER + self.cap * _REC
        self.send = t.zeros(self.slot, dtype=t.uint8, device=self.device)
        self.recv = t.zeros(self.world * self.slot, dtype=t.uint8, device=self.device)

    def fit(self, n_local: int, slack: float = 1.25):
        """Collective, outside any timed region: size the slots for the largest per-rank record count seen so far
        (times `slack`) -- the gather moves whole slots, so a tight slot is a cheap gather.  A later overflow is still
        handled by finish()."""
        t, dist = self.torch, self.dist
        need = t.tensor([int(n_local)], dtype=t.int64, device=self.device)
        if self.world > 1:
            dist.all_reduce(need, op=dist.ReduceOp.MAX)
        if self.cuda:
            self.stream.synchronize()
        self.pending = False
        self._alloc(max(int(int(need.item()) * slack), 1024))

    def _exchange(self):
        t, dist = self.torch, self.dist
        self.send[:8].view(t.int64).fill_(self.n_local)
        if self.world == 1:
            self.recv.copy_(self.send)
        elif self.cuda:
            dist.all_gather_into_tensor(self.recv, self.send)
        else:
            dist.all_gather(list(self.recv.view(self.world, self.slot).unbind(0)), self.send)

    def start(self, searcher, first_query: int = 0):
        if self.pending and self.cuda:
            self.stream.synchronize()  # the previous gather still owns the send buffer (long finished in practice)
        self._searcher, self._first = searcher, first_query
        self.n_local = searcher.export_hits(self.send.data_ptr() + self.HEADER, self.cap, first_query)
        if self.cuda:
            with self.torch.cuda.stream(self.stream):
                self.ev0.record()
                self._exchange()
                self.ev1.record()
        else:
            self._exchange()
        self.pending = True

    def _counts(self):
        t = self.torch
        return self.recv.view(self.world, self.slot)[:, :8].contiguous().view(t.int64).reshape(-1).cpu().tolist()

    def finish(self) -> int:
        """wait for the gather; returns the total number of records of all ranks"""
        if self.cuda:
            self.stream.synchronize()
            self.last_ms = self.ev0.elapsed_time(self.ev1)
        self.pending = False
        self.counts = self._counts()
        if max(self.counts) > self.cap:  # seen by every rank alike: grow together and repeat
            self.regrown += 1
            self._alloc(max(self.counts) * 5 // 4)
            self.n_local = self._searcher.export_hits(self.send.data_ptr() + self.HEADER, self.cap, self._first)
            self._exchange()
            if self.cuda:
                self.torch.cuda.synchronize(self.device)
            self.counts = self._counts()
        return int(sum(self.counts))

    def records(self, searcher=None) -> np.ndarray:
        """all ranks' records in rank order, on the host (call after finish())"""
        slots = self.recv.view(self.world, self.slot)
        parts = [slots[r, self.HEADER:self.HEADER + c * _REC].cpu().numpy().view(HIT_DT).copy()
                 for r, c in enumerate(self.counts)]
        out = np.concatenate(parts) if parts else np.zeros(0, HIT_DT)
        return searcher.fill_scores(out) if (searcher is not None and len(out)) else out
