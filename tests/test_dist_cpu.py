"""world_size-2 gloo test of the multi-GPU host logic: query sharding + the all-gather of hit records."""
import os
import socket

import numpy as np
import torch.multiprocessing as mp

from lambda_b200._abi import HIT_DT
from lambda_b200.dist import DeviceHitGather, all_gather_hits, shard_queries, shard_range


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _fake_hits(first, n_queries, seed):
    """deterministic per-query records so that every rank can predict everybody's contribution"""
    rng = np.random.default_rng(seed)
    per = rng.integers(0, 4, n_queries)
    h = np.zeros(int(per.sum()), HIT_DT)
    h["q_id"] = np.repeat(np.arange(n_queries), per)  # local ids
    h["s_id"] = rng.integers(0, 1000, len(h))
    h["score"] = rng.integers(50, 500, len(h))
    h["bit_score"] = h["score"] * 0.4
    return h


def _worker(rank, world, port, n_queries, q):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    b, e = shard_range(n_queries, rank, world)
    mine = _fake_hits(b, e - b, seed=100 + rank)
    got, total = all_gather_hits(mine, first_query=b)
    expect = []
    for r in range(world):
        rb, re = shard_range(n_queries, r, world)
        h = _fake_hits(rb, re - rb, seed=100 + r)
        h["q_id"] += rb
        expect.append(h)
    expect = np.concatenate(expect)
    ok = total == len(expect) and len(got) == len(expect) and (got == expect).all()
    # an empty contribution must work too
    empty, t2 = all_gather_hits(np.zeros(0, HIT_DT) if rank == 0 else mine, first_query=b)
    ok = ok and t2 == (len(expect) - len(expect[expect["q_id"] < shard_range(n_queries, 0, world)[1]]))
    q.put((rank, bool(ok)))
    dist.destroy_process_group()


class _FakeSearcher:
    """stands in for lambda_b200.Searcher.export_hits: 'device' records are a numpy array, the destination a CPU tensor"""

    def __init__(self, hits):
        self.hits = hits

    def export_hits(self, dev_ptr, cap_records, first_query):
        import ctypes
        if len(self.hits) <= cap_records and len(self.hits):
            h = self.hits.copy()
            h["q_id"] += first_query
            ctypes.memmove(dev_ptr, h.ctypes.data, h.nbytes)
        return len(self.hits)


def _worker_device_gather(rank, world, port, n_queries, q):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    expect = []
    for r in range(world):
        rb, re = shard_range(n_queries, r, world)
        h = _fake_hits(rb, re - rb, seed=100 + r)
        h["q_id"] += rb
        expect.append(h)
    expect = np.concatenate(expect)
    b, e = shard_range(n_queries, rank, world)
    mine = _fake_hits(b, e - b, seed=100 + rank)
    ok = True
    for cap in (4 * n_queries, 3):  # roomy slots; slots that overflow and are grown collectively
        g = DeviceHitGather(cap)
        for _ in range(2):  # the buffers are reused step after step
            g.start(_FakeSearcher(mine), first_query=b)
            total = g.finish()
            got = g.records()
            ok = ok and total == len(expect) and len(got) == len(expect) and (got == expect).all()
        ok = ok and (g.regrown == (0 if cap > 3 else 1))
    g = DeviceHitGather(4 * n_queries)
    g.fit(len(mine))  # collective: slots for the largest contribution
    ok = ok and g.cap >= len(mine) and g.cap < 4 * n_queries + 1024
    g.start(_FakeSearcher(mine), first_query=b)
    ok = ok and g.finish() == len(expect) and (g.records() == expect).all()
    g = DeviceHitGather(8)
    g.start(_FakeSearcher(np.zeros(0, HIT_DT) if rank == 0 else mine[:5]), first_query=b)
    ok = ok and g.finish() == 5 * (world - 1) and len(g.records()) == 5 * (world - 1)
    q.put((rank, bool(ok)))
    dist.destroy_process_group()


def test_device_hit_gather_gloo_world2():
    """the fixed-capacity gather (count in the first word of every slot) incl. the collective regrow on overflow"""
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker_device_gather, args=(r, 2, port, 37, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = dict(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
    assert res == {0: True, 1: True}


def test_all_gather_hits_gloo_world2():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, 37, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = dict(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
    assert res == {0: True, 1: True}


def test_shard_queries_partition():
    offs = np.concatenate([[0], np.cumsum(np.arange(1, 12))]).astype(np.uint64)
    res = np.arange(int(offs[-1]), dtype=np.uint8)
    seen = []
    for r in range(3):
        sub, o, first = shard_queries(res, offs, r, 3)
        assert o[0] == 0 and len(sub) == int(o[-1])
        for i in range(len(o) - 1):
            seen.append((first + i, bytes(sub[int(o[i]):int(o[i + 1])])))
    assert [s[0] for s in seen] == list(range(11))
    for qi, b in seen:
        assert b == bytes(res[int(offs[qi]):int(offs[qi + 1])])
    h, t = all_gather_hits(np.zeros(2, HIT_DT), first_query=5)  # no process group: identity + rebase
    assert t == 2 and (h["q_id"] == 5).all()
