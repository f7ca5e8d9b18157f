#!/usr/bin/env python
"""Benchmark of the seed-and-extend hot path (BASELINE.json metric: query-seqs/s + GCUPS, searchp).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

Workload (BASELINE.json configs[1]): searchp, 100 000 synthetic 300-aa queries against a
5 000 000-sequence synthetic protein index (Li10 FM index, BLOSUM62, default profile), per GPU.
One "step" = one pass of the whole hot path (seeding -> merge -> DP score pass -> filter -> DP trace
pass -> hits) over the rank's 100k-query batch.  Queries shard naturally, so N GPUs = N independent
shards of 100k queries (weak scaling) + one NCCL all-gather of the hit records at the end of the step.

The index is built ONCE per box by the unmodified reference (oracle/_ref/lambda3 mkindexp -- indexing
is out of scope) and cached under $LAMBDA_B200_CACHE (default /tmp/lambda_b200_cache).

JSON line (rank 0): see the contract in the task description; extra keys: gcups_score, gcups_trace,
stage_ms, hits, funnel, parity_sample.
"""
from __future__ import annotations

import argparse
import hashlib
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
REF = os.path.join(ROOT, "oracle", "_ref", "lambda3")
CACHE = os.environ.get("LAMBDA_B200_CACHE", "/tmp/lambda_b200_cache")


def log(*a):
    print("[bench]", *a, file=sys.stderr, flush=True)


def kernel_sources_sha():
    """identifies the kernel sources a build / an ncu capture belongs to"""
    h = hashlib.sha1()
    src = os.path.join(ROOT, "lambda_b200", "csrc")
    for fn in sorted(os.listdir(src)):
        if fn.endswith((".cuh", ".cu")):
            h.update(open(os.path.join(src, fn), "rb").read())
    return h.hexdigest()[:12]


# ------------------------------------------------------------------------------------------------
# workload
# ------------------------------------------------------------------------------------------------

WORKLOADS = {
    # BASELINE.json configs[0]: the reference's own CPU-runnable case (parity-sized; not a headline bench line)
    "searchp_small": dict(domain="protein", dom=0, mk="mkindexp", search="searchp", n_seqs=50_000, n_queries=1_000,
                          qlen=150, unit="aa", what="synthetic protein index (Li10), BLOSUM62"),
    # BASELINE.json configs[1]: the configuration the headline metric is quoted on
    "searchp": dict(domain="protein", dom=0, mk="mkindexp", search="searchp", n_seqs=5_000_000, n_queries=100_000,
                    qlen=300, unit="aa", what="synthetic protein index (Li10), BLOSUM62"),
    # BASELINE.json configs[2]: n_seqs = chromosomes of 1 Mbp each (200 Mbp)
    "searchn": dict(domain="nucleotide", dom=1, mk="mkindexn", search="searchn", n_seqs=200, n_queries=1_000_000,
                    qlen=150, unit="bp", what="x1Mbp synthetic nucleotide index (dna4), match 2 / mismatch -3"),
    # BASELINE.json configs[4]: bisulfite reads vs a 50 Mbp genome (the index holds the C->T and G->A texts)
    "searchbs": dict(domain="bisulfite", dom=2, mk="mkindexbs", search="searchbs", n_seqs=50, n_queries=500_000,
                     qlen=100, unit="bp", what="x1Mbp synthetic genome, bisulfite index (dna3bs), C->T converted reads"),
    # BASELINE.json configs[3] (parity mode): real length distribution against the searchp index
    "searchp_real": dict(domain="protein", dom=0, mk="mkindexp", search="searchp", n_seqs=5_000_000, n_queries=10_000,
                         qlen=0, unit="aa", what="synthetic protein index (Li10), BLOSUM62, log-normal query lengths "
                                                  "(median 270, 50..2000)", index_of="searchp"),
}
CHROM_LEN = 1_000_000


def workload_dir(wl, n_seqs, seed):
    wl = WORKLOADS[wl].get("index_of", wl)
    key = hashlib.sha1(f"{wl}-flat-{n_seqs}-{seed}-v2".encode()).hexdigest()[:12]
    return os.path.join(CACHE, f"{wl}_{n_seqs}_{key}")


def ensure_index(wl, n_seqs, seed=1):
    """database FASTA + reference-built .lba, cached; safe against concurrent ranks"""
    from lambda_b200 import synth
    W = WORKLOADS[wl]
    d = workload_dir(wl, n_seqs, seed)
    done = os.path.join(d, "READY")
    if os.path.exists(done):
        return d
    os.makedirs(d, exist_ok=True)
    lock = os.path.join(d, "LOCK")
    try:
        fd = os.open(lock, os.O_CREAT | os.O_EXCL | os.O_WRONLY)
        os.close(fd)
    except FileExistsError:
        log("waiting for another process to build the index ...")
        t0 = time.time()
        while not os.path.exists(done):
            time.sleep(2)
            try:
                stale = time.time() - os.path.getmtime(lock) > 1800  # a killed builder leaves its lock behind
            except OSError:
                stale = False  # the builder just finished (or failed) and removed it
            if stale:
                log("stale lock: the previous builder died; taking over")
                try:
                    os.remove(lock)
                except OSError:
                    pass
                return ensure_index(wl, n_seqs, seed)
            if not os.path.exists(lock) and not os.path.exists(done):
                return ensure_index(wl, n_seqs, seed)  # the builder failed: try ourselves (and report its error again)
            if time.time() - t0 > 3600:
                raise RuntimeError("timed out waiting for the index build")
        return d
    try:
        if not os.path.exists(REF):
            raise RuntimeError(f"{REF} missing: run `python -c 'import __graft_entry__ as g; g.build()'` where "
                               "/root/reference is available (the binary travels with the repo snapshot)")
        t0 = time.time()
        if W["dom"] == 0:
            db, offs = synth.protein_db(n_seqs, seed=seed)
        else:
            db, offs = synth.nucl_db(n_seqs, CHROM_LEN, seed=seed)
        synth.write_fasta(os.path.join(d, "db.fasta"), db, offs, "S")
        np.save(os.path.join(d, "db_offsets.npy"), offs)
        log(f"database: {n_seqs} seqs, {int(offs[-1])} residues generated in {time.time() - t0:.1f}s")
        t0 = time.time()
        lba = os.path.join(d, "db.lba")
        if os.path.exists(lba):
            os.remove(lba)
        subprocess.check_call([REF, W["mk"], "-d", os.path.join(d, "db.fasta"), "-i", lba, "-v", "0",
                               "-t", str(os.cpu_count() or 1)])
        log(f"reference {W['mk']}: {time.time() - t0:.1f}s, {os.path.getsize(lba) / 1e9:.2f} GB")
        open(done, "w").write("ok\n")
    finally:
        os.remove(lock)
    return d


def make_queries(wl, d, n_queries, qlen, seed):
    """mutated windows of database sequences; returns (ascii residues, offsets)"""
    from lambda_b200 import synth
    offs = np.load(os.path.join(d, "db_offsets.npy"))
    fa = np.memmap(os.path.join(d, "db.fasta"), dtype=np.uint8, mode="r")
    # recover residue positions inside the FASTA: record i = ">S<i>\n" + seq + "\n"
    n = len(offs) - 1
    idlen = np.char.str_len(np.arange(n).astype(str)).astype(np.int64) + 3
    rec_start = np.zeros(n + 1, np.int64)
    np.cumsum(idlen + np.diff(offs) + 1, out=rec_start[1:])
    seq_start = rec_start[:-1] + idlen
    rng = np.random.default_rng(seed)
    lens = np.diff(offs)
    if qlen == 0:
        # real length distribution: every query is a mutated window of a database sequence at least as long
        want = synth._protein_lengths(rng, n_queries)
        order = np.argsort(lens)
        first = np.searchsorted(lens[order], want)  # sequences order[first:] are long enough
        pick = order[first + (rng.random(n_queries) * (len(order) - first)).astype(np.int64)]
        start = seq_start[pick] + (rng.random(n_queries) * (lens[pick] - want + 1)).astype(np.int64)
        qoffs = np.zeros(n_queries + 1, np.uint64)
        np.cumsum(want, out=qoffs[1:].view(np.int64))
        out = np.empty(int(qoffs[-1]), np.uint8)
        for i in range(n_queries):
            seq = np.array(fa[start[i]:start[i] + want[i]])
            m = rng.random(len(seq)) < rng.uniform(0.15, 0.20)
            seq[m] = synth._random_residues(rng, int(m.sum()))
            out[int(qoffs[i]):int(qoffs[i + 1])] = seq
        return out, qoffs
    elig = np.nonzero(lens >= qlen)[0]
    pick = elig[rng.integers(0, len(elig), n_queries)]
    start = seq_start[pick] + (rng.random(n_queries) * (lens[pick] - qlen + 1)).astype(np.int64)
    q = np.empty((n_queries, qlen), np.uint8)
    step = 1 << 16
    for b in range(0, n_queries, step):
        e = min(n_queries, b + step)
        q[b:e] = fa[start[b:e, None] + np.arange(qlen)[None, :]]
    if wl == "searchbs":
        conv = (q == ord("C")) & (rng.random(q.shape, dtype=np.float32) < 0.95)
        q[conv] = ord("T")
        m = rng.random(q.shape, dtype=np.float32) < 0.01
        q[m] = synth.NT[rng.integers(0, 4, int(m.sum()), dtype=np.uint8)]
        odd = np.arange(n_queries) % 2 == 1
        q[odd] = synth._COMP[q[odd][:, ::-1]]
        return q.reshape(-1), np.arange(n_queries + 1, dtype=np.uint64) * qlen
    if wl == "searchn":
        m = rng.random(q.shape, dtype=np.float32) < 0.03
        q[m] = synth.NT[rng.integers(0, 4, int(m.sum()), dtype=np.uint8)]
        odd = np.arange(n_queries) % 2 == 1
        q[odd] = synth._COMP[q[odd][:, ::-1]]
        return q.reshape(-1), np.arange(n_queries + 1, dtype=np.uint64) * qlen
    rate = rng.uniform(0.15, 0.20, n_queries)[:, None]
    m = rng.random(q.shape, dtype=np.float32) < rate
    q[m] = synth._random_residues(rng, int(m.sum()))
    ev = rng.random(q.shape, dtype=np.float32) < 0.01
    rows, cols = np.nonzero(ev)
    kinds = rng.random(len(rows)) < 0.5
    fill = synth._random_residues(rng, len(rows))
    for r, c, k, f in zip(rows, cols, kinds, fill):
        if k:
            q[r, c:-1] = q[r, c + 1:]
            q[r, -1] = f
        else:
            q[r, c + 1:] = q[r, c:-1]
            q[r, c] = f
    return q.reshape(-1), np.arange(n_queries + 1, dtype=np.uint64) * qlen


# ------------------------------------------------------------------------------------------------
# clocks
# ------------------------------------------------------------------------------------------------

class ClockSampler(threading.Thread):
    """SM clock + throttle reasons of one GPU, sampled DURING the timed region.  NVML in-process (a sample costs
    well under a millisecond, so even a 0.3 s region gets tens of samples); `nvidia-smi` only if NVML cannot be
    loaded.  The first sample is taken synchronously in start_sampling() so the handle/driver path is warm."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
    REASONS = (("hw_slowdown", 0x8), ("hw_thermal_slowdown", 0x40), ("sw_thermal_slowdown", 0x20),
               ("sw_power_cap", 0x4), ("hw_power_brake_slowdown", 0x80))

    def __init__(self, gpu):
        super().__init__(daemon=True)
        self.gpu, self.rows, self.stop_flag, self.active = self._physical_index(gpu), [], False, False
        self.nvml, self.handle, self.max_mhz, self.source = None, None, None, "nvidia-smi"
        try:
            import pynvml
            pynvml.nvmlInit()
            self.handle = pynvml.nvmlDeviceGetHandleByIndex(self.gpu)
            self.max_mhz = int(pynvml.nvmlDeviceGetMaxClockInfo(self.handle, pynvml.NVML_CLOCK_SM))
            self.nvml, self.source = pynvml, "nvml"
        except Exception as e:  # noqa: BLE001 - fall back to the command-line tool
            log(f"NVML unavailable ({e}); sampling clocks with nvidia-smi")

    @staticmethod
    def _physical_index(local):
        vis = os.environ.get("CUDA_VISIBLE_DEVICES", "")
        ids = [x.strip() for x in vis.split(",") if x.strip()]
        if local < len(ids) and ids[local].isdigit():
            return int(ids[local])
        return local

    def _sample(self):
        if self.nvml is not None:
            n = self.nvml
            mhz = int(n.nvmlDeviceGetClockInfo(self.handle, n.NVML_CLOCK_SM))
            try:
                mask = int(n.nvmlDeviceGetCurrentClocksEventReasons(self.handle))
            except Exception:  # noqa: BLE001 - older binding name
                mask = int(n.nvmlDeviceGetCurrentClocksThrottleReasons(self.handle))
            return mhz, self.max_mhz, [name for name, bit in self.REASONS if mask & bit]
        out = subprocess.run(["nvidia-smi", "-i", str(self.gpu), f"--query-gpu={self.Q}",
                              "--format=csv,noheader,nounits"], capture_output=True, text=True, timeout=10).stdout
        r = [x.strip() for x in out.strip().split(",")]
        names = ("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap")
        return int(r[0]), int(r[1]), [nm for nm, v in zip(names, r[2:6]) if v.lower().startswith("active")]

    def run(self):
        while not self.stop_flag:
            if self.active:
                try:
                    self.rows.append(self._sample())
                except Exception:  # noqa: BLE001 - a failed sample is just missing
                    pass
            time.sleep(0.01 if self.nvml is not None else 0.2)

    def begin(self):
        """call right before the timed region"""
        self.active = True

    def end(self):
        """call right after the timed region; with the slow nvidia-smi path make sure at least one sample that
        STARTED inside the region is waited for"""
        self.active = False

    def summary(self):
        self.stop_flag = True
        if self.is_alive():
            self.join(timeout=15)
        sm = [r[0] for r in self.rows]
        mx = [r[1] for r in self.rows if r[1]]
        reasons = sorted({x for r in self.rows for x in r[2]})
        return {"sm_mhz": int(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(sm), "source": self.source}


# ------------------------------------------------------------------------------------------------
# reference arm / cpu baseline
# ------------------------------------------------------------------------------------------------

def run_reference_search(wl, d, q_ascii, qoffs, n_sample, threads, tag):
    """time the unmodified reference (searchp, OpenMP) on the first n_sample queries; returns
    (queries/s over the reference's own 'Runtime total' search phase, wall seconds, hits)"""
    import re
    import tempfile
    from lambda_b200 import synth
    with tempfile.TemporaryDirectory() as tmp:
        qf = os.path.join(tmp, "q.fasta")
        synth.write_fasta(qf, q_ascii[: int(qoffs[n_sample])], qoffs[: n_sample + 1].astype(np.int64), "Q")
        out = os.path.join(tmp, "out.m8")
        t0 = time.time()
        txt = subprocess.run([REF, WORKLOADS[wl]["search"], "-q", qf, "-i", os.path.join(d, "db.lba"), "-o", out, "-t", str(threads),
                              "--version-to-outputfile", "0", "-v", "2"], check=True, capture_output=True, text=True).stdout
        wall = time.time() - t0
        m = re.search(r"Runtime total: ([0-9.eE+-]+)s", txt)
        phase = float(m.group(1)) if m else wall
        lines = open(out).read().splitlines(True)
    return n_sample / phase, wall, phase, lines


# ------------------------------------------------------------------------------------------------
# main
# ------------------------------------------------------------------------------------------------

def main():
    # libraries (NCCL's version banner, ...) may print to stdout; the contract is ONE JSON line there
    real_stdout = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)

    def emit(obj):
        real_stdout.write(json.dumps(obj) + "\n")
        real_stdout.flush()

    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="searchp", choices=sorted(WORKLOADS))
    ap.add_argument("--n-seqs", type=int, default=0, help="database sequences (0 = the workload's BASELINE size)")
    ap.add_argument("--n-queries", type=int, default=0)
    ap.add_argument("--qlen", type=int, default=0)
    ap.add_argument("--cpu-sample", type=int, default=0, help="queries for the CPU baseline (0 = auto)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"],
                    help="weak: every rank searches its own --n-queries queries; strong: the ranks split ONE batch of "
                         "--n-queries queries (the reference's only parallel axis, src/search.cpp:379-385)")
    ap.add_argument("--parity-queries", type=int, default=4000,
                    help="N > 1: every rank checks the first this-many queries of its shard against the reference binary")
    ap.add_argument("--band", type=int, default=0,
                    help="lgpu_params.window_band: 0 = the reference's rule floor(sqrt(qlen))+1 (parity mode); 16/32/64 = the "
                         "band sweep of BASELINE configs[3] (non-parity: the reference binary has no such option)")
    args = ap.parse_args()
    wl = args.workload
    W = WORKLOADS[wl]
    args.n_seqs = args.n_seqs or int(os.environ.get("LAMBDA_B200_NSEQS", W["n_seqs"]))
    args.n_queries = args.n_queries or int(os.environ.get("LAMBDA_B200_NQUERIES", W["n_queries"]))
    args.qlen = args.qlen or W["qlen"]
    qdesc = f"{args.qlen}{W['unit']}" if args.qlen else f"(50..2000){W['unit']}"

    rank = int(os.environ.get("RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))
    local_rank = int(os.environ.get("LOCAL_RANK", 0))
    if world > 1 and args.gpus != world:
        log(f"--gpus {args.gpus} != WORLD_SIZE {world}; using WORLD_SIZE")
    n_gpus = world
    strong = args.scaling == "strong" and n_gpus > 1
    workload = (f"{W['search']}: {args.n_queries}x{qdesc} synthetic queries vs {args.n_seqs}-seq {W['what']}, "
                f"default profile")
    cfg = {"workload": workload, "queries_per_gpu": args.n_queries // n_gpus if strong else args.n_queries,
           "queries_total": args.n_queries if strong else args.n_queries * n_gpus,
           "query_len": args.qlen, "index_seqs": args.n_seqs,
           "profile": "none", "sharding": (f"one batch of {args.n_queries} queries split over {n_gpus} ranks" if strong else
                                           f"queries x{n_gpus}") + ", index replicated",
           "streams_per_gpu": "library default: 1 sub-batch in flight for protein searches, 4 for nucleotide / bisulfite",
           "l2_policy": "inputs larger than L2 (index and per-step trace/DP working sets are GBs)",
           "window_band": args.band if args.band else "reference rule floor(sqrt(qlen))+1"}
    cores = os.cpu_count() or 1

    if args.impl == "reference":
        if rank != 0:
            return
        d = ensure_index(wl, args.n_seqs)
        q_ascii, qoffs = make_queries(wl, d, args.n_queries, args.qlen, seed=1000)
        n_sample = args.cpu_sample or min(args.n_queries, max(2000, 2000 * cores))
        for _ in range(args.warmup):
            run_reference_search(wl, d, q_ascii, qoffs, min(n_sample, 200), cores, "warm")
        qps, walls = [], []
        for _ in range(args.steps):
            v, wall, phase, _ = run_reference_search(wl, d, q_ascii, qoffs, n_sample, cores, "step")
            qps.append(n_sample / phase)
            walls.append(phase)
        value = n_sample * len(walls) / sum(walls)
        sample = f"first {n_sample} of the {args.n_queries} queries per step, lambda3 {wl} -t {cores}, search phase"
        emit(({"impl": "reference", "metric": f"{wl}_query_seqs_per_s", "value": value, "unit": "queries/s",
                          "n_gpus": n_gpus, "steps": args.steps, "warmup": args.warmup,
                          "ms_per_step": 1e3 * sum(walls) / len(walls), "higher_is_better": True, "scaling": args.scaling,
                          "vs_baseline": None, "dtype": "int16", "data": "synthetic", "config": cfg,
                          "cpu_baseline": {"value": value, "unit": "queries/s", "cores": cores, "kind": "reference",
                                           "sample": sample},
                          "e2e": {"value": value, "unit": "queries/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}))
        return

    import torch
    import torch.distributed as dist
    import lambda_b200
    from lambda_b200._abi import HIT_DT

    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    d = ensure_index(wl, args.n_seqs)
    if world > 1:
        dist.barrier()
    t0 = time.time()
    ix = lambda_b200.Index.load(os.path.join(d, "db.lba"), device=local_rank, keep_ids=False)

    class _SyntheticIds:  # the benchmark database is written with ids S0, S1, ... (ensure_index)
        def __getitem__(self, i):
            return f"S{i}"
    ix.subject_ids = _SyntheticIds()
    log(f"rank {rank}: index in HBM: {ix.device_bytes / 1e9:.2f} GB, load {time.time() - t0:.1f}s")
    kw = {"window_band": args.band} if args.band else {}
    s = lambda_b200.Searcher(ix, W["domain"], **kw)              # default number of sub-batches in flight
    s_serial = lambda_b200.Searcher(ix, W["domain"], streams=1, **kw)  # strictly serial: per-kernel timing / roofline
    if strong:
        # ONE batch for the whole job (the same on every rank); this rank searches its contiguous shard
        from lambda_b200.dist import shard_queries
        q_ascii, qoffs = make_queries(wl, d, args.n_queries, args.qlen, seed=1000)
        q_ascii, qoffs, first_query = shard_queries(q_ascii, qoffs, rank, world)
        q_ascii = np.ascontiguousarray(q_ascii)
    else:
        q_ascii, qoffs = make_queries(wl, d, args.n_queries, args.qlen, seed=1000 + rank)
        first_query = rank * args.n_queries
    n_local = len(qoffs) - 1
    res = lambda_b200.encode(q_ascii, W["dom"])
    h_res = torch.from_numpy(res).pin_memory()
    h_offs = torch.from_numpy(qoffs.view(np.int64)).pin_memory()
    d_res = h_res.cuda()
    d_offs = h_offs.cuda()
    pin_res, pin_offs = h_res.numpy(), h_offs.numpy().view(np.uint64)  # numpy views of the pinned buffers

    # The path's only collective: all ranks exchange their hit records GPU to GPU, straight from the library's device
    # buffer, with one fixed-capacity NCCL all-gather on a side stream (lambda_b200/dist.py DeviceHitGather).
    from lambda_b200.dist import DeviceHitGather
    gather = DeviceHitGather(max(2 * n_local, 1024)) if world > 1 else None

    def step(resident, searcher=None):
        searcher = searcher or s
        if resident:
            hits, st = searcher.search(d_res, d_offs)
        else:
            # host buffers (pinned): H2D of the queries + D2H of the hit records happen inside the call;
            # copy=False returns a view of the library's result buffer (valid until the next call)
            hits, st = searcher.search(pin_res, pin_offs, copy=False)
        if gather is not None:
            gather.start(searcher, first_query)  # returns at once; overlaps the next step's search
        return hits, st

    def timed(resident, k, searcher=None):
        """K steps.  Device time = CUDA events recorded by the library on ITS stream around every
        lgpu_search_batch call (ms_total; torch.cuda.Event only sees torch's stream) + the exposed tail of the
        last step's gather (torch events on the gather's side stream; the gathers of the earlier steps run
        underneath the next search); max over ranks.  Host wall-clock around the same region is reported next to it."""
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        acc, ms, per_step, g_ms = None, 0.0, [], []
        total = 0
        for i in range(k):
            hits, st = step(resident, searcher)
            ms += float(st["ms_total"])
            per_step.append(float(st["ms_total"]))
            acc = st.copy() if acc is None else _acc(acc, st)
            if gather is not None and i + 1 < k:
                pass  # the gather of this step is waited for by the next start()
        if gather is not None:
            tw = time.perf_counter()
            total = gather.finish()
            ms += gather.last_ms  # device time of the last gather (events on its stream): the only one not hidden
            g_ms.append(gather.last_ms)
            del tw
        else:
            total = len(hits)
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        wall = (time.perf_counter() - t0) * 1e3
        mine_ms = ms
        if world > 1:
            t = torch.tensor([ms, wall], device="cuda", dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            tmin = torch.tensor([mine_ms], device="cuda", dtype=torch.float64)
            dist.all_reduce(tmin, op=dist.ReduceOp.MIN)
            ms, wall = float(t[0]), float(t[1])
            spread = (float(tmin[0]) / k, ms / k)
        else:
            spread = (ms / k, ms / k)
        return ms, wall, acc, hits, total, {"rank_ms_per_step_min_max": spread, "gather_device_ms_last": g_ms[-1] if g_ms else 0.0}

    def _acc(a, b):
        for n in a.dtype.names:
            a[n] += b[n]
        return a

    hits_w = None
    for _ in range(args.warmup):
        hits_w, _ = step(True)
        step(True, s_serial)
    if gather is not None and hits_w is not None:
        gather.finish()
        gather.fit(len(hits_w))  # slots sized for what the ranks really produce (collective, before the timed region)
    sampler = ClockSampler(local_rank) if rank == 0 else None
    if sampler:
        sampler.start()
        sampler.begin()
    ms_res, wall_res, st_pipe, hits, total_hits, tinfo = timed(True, args.steps)
    hits = hits.copy()
    ms_e2e, wall_e2e, st_e2e, _, _, tinfo_e2e = timed(False, args.steps)
    # same steps again strictly serial (one stream): per-stage / per-kernel device times for the roofline
    ms_serial, _, st, _, _, _ = timed(True, args.steps, s_serial)
    if sampler:
        sampler.end()
    clocks = sampler.summary() if sampler else None

    nq_total = (args.n_queries if strong else args.n_queries * n_gpus) * args.steps
    value = nq_total / (ms_res * 1e-3)
    # end to end = host wall-clock around the public API call (host buffers in, host records out)
    e2e = nq_total / (wall_e2e * 1e-3)
    cells_score, cells_trace = float(st["cells_score"]), float(st["cells_trace"])
    if world > 1:
        t = torch.tensor([cells_score, cells_trace], device="cuda", dtype=torch.float64)
        dist.all_reduce(t)
        cells_score_all, cells_trace_all = float(t[0]), float(t[1])
    else:
        cells_score_all, cells_trace_all = cells_score, cells_trace
    # parity: our tabular lines against the unmodified reference run on this box, on a sample of THIS rank's queries
    parity, ref_run = None, None
    if not args.band and os.path.exists(REF) and not (args.no_cpu_baseline and n_gpus == 1):
        if n_gpus == 1:
            n_sample = args.cpu_sample or min(n_local, max(2000, 2000 * cores))
            threads = cores
        else:
            n_sample = min(n_local, args.parity_queries)
            threads = max(1, cores // world)
        qps, wall, phase, ref_lines = run_reference_search(wl, d, q_ascii, qoffs, n_sample, threads, "cpu")
        ids = [f"Q{i}" for i in range(n_sample)]
        mine = sorted(s.m8(hits[hits["q_id"] < n_sample], ids))
        same = mine == sorted(ref_lines)
        ref_run = (qps, wall, phase, n_sample, threads)
        parity = {"queries": n_sample, "reference_lines": len(ref_lines), "our_lines": len(mine), "identical": bool(same)}
        if world > 1:
            t = torch.tensor([1 if same else 0, len(ref_lines), len(mine)], device="cuda", dtype=torch.int64)
            tmin = t.clone()
            dist.all_reduce(tmin, op=dist.ReduceOp.MIN)
            dist.all_reduce(t)
            parity = {"queries_per_rank": n_sample, "ranks": world, "reference_lines": int(t[1]), "our_lines": int(t[2]),
                      "identical": bool(int(tmin[0]) == 1), "ranks_identical": int(t[0])}
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # roofline of the dominant kernel (DP score pass): integer-ALU bound, SURVEY §8(d)
    peaks_path = os.path.join(ROOT, "profiles", "INT_PEAKS.json")
    peak_gops, peak_src = None, "unmeasured"
    if os.path.exists(peaks_path):
        pk = json.load(open(peaks_path))
        # one VIADDMNMX.S16x2 warp-instruction = 32 lanes x 2 halves x (add + max) = 128 int16 ops
        peak_gops = pk["viaddmnmx_s16x2"] * 128.0
        peak_src = "profiles/INT_PEAKS.json (bin/int16_peak on this pool's B200)"
    ms_score = float(st["ms_extend_score"]) / args.steps
    gcups_score = cells_score / args.steps / (ms_score * 1e-3) / 1e9 if ms_score > 0 else 0.0
    ms_trace = float(st["ms_extend_trace"]) / args.steps
    gcups_trace = cells_trace / args.steps / (ms_trace * 1e-3) / 1e9 if ms_trace > 0 else 0.0
    achieved = gcups_score * 10.0  # 10 int16 ops per cell update (5 add + 5 max), SURVEY §8(d)
    # DRAM traffic of that kernel from the committed `ncu --set full` capture of this workload (per launch).  A capture
    # is only used if it was taken from the kernel sources this run was built from (sha of csrc/*.cuh + engine.cu).
    ksha = kernel_sources_sha()
    cap, cap_note = None, "no capture committed for this workload"
    tp = os.path.join(ROOT, "profiles", f"r2_ncu_kernels_{wl}.json")
    if os.path.exists(tp):
        try:
            cap = json.load(open(tp))
            if cap.get("kernel_sources_sha") != ksha:
                cap_note = (f"stale: {os.path.relpath(tp, ROOT)} was captured at kernel sources {cap.get('kernel_sources_sha')} "
                            f"(git {cap.get('git')}), this run is {ksha}")
                cap = None
            else:
                cap_note = f"{os.path.relpath(tp, ROOT)} (git {cap.get('git')})"
        except Exception as e:  # noqa: BLE001
            cap, cap_note = None, f"unreadable capture: {e}"

    traffic = None
    score_name = [k for k in (cap or {}).get("kernels", []) if k["name"].startswith("swDpxKernel") and not k.get("trace")]
    if score_name:
        traffic = float(score_name[0]["dram_bytes"])
    roofline = {"kernel": "swDpxKernel<T,K,PRIV,false> (DP pass 1, packed int16 DPX)", "bound": "int16-alu", "achieved": achieved,
                "peak": peak_gops, "unit": "Gop/s (int16)", "frac": (achieved / peak_gops) if peak_gops else None,
                "peak_source": peak_src, "traffic": traffic, "traffic_unit": "bytes per launch (DRAM read + write, ncu --set full)",
                "traffic_source": cap_note, "gcups": gcups_score,
                "algorithmic": "10 int16 ops per DP cell x cells of the step / DP pass-1 stage time (CUDA events)"}

    # DP pass 2 = the same packed recurrence + ONE byte per cell (H mod 256) written to HBM + the traceback.  Two views:
    # against the int16 ALU peak (10 ops per cell like pass 1: the fill is ALU bound now) and against the HBM copy peak
    # with the reference's own 1 B per cell (SQ/align/dp_profile.h:122) as algorithmic bytes.
    hbm_peak, hbm_src = 6533.2, "fallback: MEASURED_PEAKS.json value recorded in DESIGN.md"
    try:
        hbm_peak = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
        hbm_src = "MEASURED_PEAKS.json"
    except Exception:  # noqa: BLE001 - file is driver-written and may be absent
        pass
    trace_gbs = 1.0 * cells_trace / args.steps / (ms_trace * 1e-3) / 1e9 if ms_trace > 0 else 0.0
    trace_traffic = [k for k in (cap or {}).get("kernels", []) if k["name"].startswith("swDpxKernel") and k.get("trace")]
    roofline_trace = {"kernel": "swDpxKernel<T,K,true,true> + tracebackResKernel (DP pass 2: fill + traceback)", "bound": "int16-alu",
                      "achieved": gcups_trace * 10.0, "peak": peak_gops, "unit": "Gop/s (int16)",
                      "frac": (gcups_trace * 10.0 / peak_gops) if peak_gops else None, "gcups": gcups_trace,
                      "hbm": {"achieved": trace_gbs, "peak": hbm_peak, "unit": "GB/s", "frac": trace_gbs / hbm_peak,
                              "peak_source": hbm_src, "algorithmic": "1 B per DP cell (residue plane; the reference's trace "
                                                                     "matrix is 1 B per cell too) / DP pass-2 stage time"},
                      "traffic": float(trace_traffic[0]["dram_bytes"]) if trace_traffic else None,
                      "algorithmic": "10 int16 ops per DP cell x cells of the trace pass / DP pass-2 stage time (CUDA events; the "
                                     "stage also contains the classification, the traceback and the wavefront ramps)"}
    other = cap

    out = {"metric": f"{wl}_query_seqs_per_s", "value": value, "unit": "queries/s", "n_gpus": n_gpus,
           "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_res / args.steps, "wall_ms_per_step": wall_res / args.steps,
           "ms_per_step_serial_1_stream": ms_serial / args.steps,
           "higher_is_better": True, "scaling": args.scaling if n_gpus > 1 else "weak", "vs_baseline": None, "dtype": "int16",
           "data": "synthetic", "config": cfg,
           "gather": ({"what": "one fixed-capacity NCCL all-gather of the hit records per step, device to device, on a side "
                               "stream (overlaps the next step); only the last one of the timed region is exposed",
                       "device_ms_last": tinfo["gather_device_ms_last"],
                       "gather_ms_per_step": tinfo["gather_device_ms_last"] / args.steps,  # what the timed region pays per step
                       "slot_records": gather.cap,
                       "regrown": gather.regrown} if gather is not None else None),
           "rank_ms_per_step_min_max": tinfo["rank_ms_per_step_min_max"],
           "gcups": (cells_score_all + cells_trace_all) / (ms_res * 1e-3) / 1e9,
           "gcups_score_kernel": gcups_score, "gcups_trace_kernel": gcups_trace,
           "stage_ms": {k: float(st[k]) / args.steps for k in ("ms_seed", "ms_sort_merge", "ms_extend_score",
                                                               "ms_extend_trace", "ms_h2d", "ms_host", "ms_total")},
           "funnel": {k: int(st[k]) // args.steps for k in ("hits_after_seeding", "hits_failed_pre_extend",
                                                            "hits_duplicate", "hits_failed_evalue", "hits_final",
                                                            "n_extensions_score", "n_extensions_trace")},
           "hits_per_step_all_ranks": total_hits,
           "clocks": clocks, "roofline": roofline, "roofline_trace": roofline_trace,
           "e2e": {"value": e2e, "unit": "queries/s", "h2d_bytes_per_step": int(res.nbytes + qoffs.nbytes),
                   "d2h_bytes_per_step": int(len(hits) * HIT_DT.itemsize), "ms_per_step": wall_e2e / args.steps,
                   "device_ms_per_step": ms_e2e / args.steps, "timing": "host wall-clock around Searcher.search()"},
           "gpu_launches": int(st_pipe["kernel_launches"]), "kernel_sources_sha": ksha, "kernels_ncu": other}

    # CPU baseline: the unmodified reference on this box's host cores, bounded sample of the same queries
    out["parity_sample"] = parity
    if ref_run is not None and n_gpus == 1:
        qps, wall, phase, n_sample, threads = ref_run
        out["cpu_baseline"] = {"value": qps, "unit": "queries/s", "cores": threads, "kind": "reference",
                               "sample": f"first {n_sample} of the {args.n_queries} queries, lambda3 {wl} -t {threads} "
                                         f"(SSE4 build), reference's own search-phase timer {phase:.2f}s, wall {wall:.2f}s"}
    emit(out)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
