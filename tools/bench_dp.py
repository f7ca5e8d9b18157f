#!/usr/bin/env python
"""Kernel-level benchmark of the extension DP (no big index needed): N queries x W windows each against
the small golden protein database, through lgpu_extend_scores / lgpu_extend_trace.

    python tools/bench_dp.py [--queries 100000] [--windows 16] [--qlen 300] [--wlen 336] [--trace]

Prints one JSON line per repetition with the stage time measured by the library's CUDA events and GCUPS.
LAMBDA_B200_LIB=<path> selects an alternative build of the library (kernel experiments).
"""
import argparse
import gzip
import json
import os
import shutil
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--queries", type=int, default=100000)
    ap.add_argument("--windows", type=int, default=16)
    ap.add_argument("--qlen", type=int, default=300)
    ap.add_argument("--wlen", type=int, default=336)
    ap.add_argument("--reps", type=int, default=3)
    ap.add_argument("--trace", action="store_true")
    ap.add_argument("--ragged", action="store_true", help="log-normal query lengths (median qlen)")
    a = ap.parse_args()
    import lambda_b200
    from lambda_b200._abi import MATCH_DT
    gold = os.path.join(ROOT, "tests", "golden", "prot_flat")
    with tempfile.TemporaryDirectory() as tmp:
        lba = os.path.join(tmp, "db.lba")
        with gzip.open(os.path.join(gold, "db.lba.gz"), "rb") as fi, open(lba, "wb") as fo:
            shutil.copyfileobj(fi, fo)
        ix = lambda_b200.Index.load(lba, device=0)
        s = lambda_b200.Searcher(ix, "protein")
        fa = os.path.join(tmp, "db.fasta")
        with gzip.open(os.path.join(gold, "db.fasta.gz"), "rb") as fi, open(fa, "wb") as fo:
            shutil.copyfileobj(fi, fo)
        _, db, doffs = lambda_b200.read_fasta(fa)
        slen = np.diff(doffs.astype(np.int64))
        rng = np.random.default_rng(7)
        if a.ragged:
            qlens = np.clip(np.exp(rng.normal(np.log(a.qlen), 0.6, a.queries)), 50, 2000).astype(np.int64)
        else:
            qlens = np.full(a.queries, a.qlen, np.int64)
        offs = np.zeros(a.queries + 1, np.uint64)
        np.cumsum(qlens, out=offs[1:].view(np.int64))
        res = rng.integers(0, 20, int(offs[-1]), dtype=np.uint8)
        n = a.queries * a.windows
        win = np.zeros(n, MATCH_DT)
        q = np.repeat(np.arange(a.queries), a.windows)
        ql = qlens[q]
        wl = ql + 2 * (np.sqrt(ql).astype(np.int64) + 1) if a.ragged else np.full(n, a.wlen, np.int64)
        elig = np.nonzero(slen >= wl.max())[0] if not a.ragged else np.argsort(-slen)[:8]
        sj = elig[rng.integers(0, len(elig), n)]
        wl = np.minimum(wl, slen[sj])
        st0 = (rng.random(n) * (slen[sj] - wl + 1)).astype(np.int64)
        win["qry_id"], win["subj_id"] = q, sj
        win["qry_start"], win["qry_end"] = 0, ql
        win["subj_start"], win["subj_end"] = st0, st0 + wl
        fn = s.extend_trace if a.trace else s.extend_scores
        key = "ms_extend_trace" if a.trace else "ms_extend_score"
        ckey = "cells_trace" if a.trace else "cells_score"
        for rep in range(a.reps):
            out, st = fn(res, offs, win)
            ms = float(st[key])
            cells = float(st[ckey])
            print(json.dumps({"lib": os.environ.get("LAMBDA_B200_LIB", "default"), "stage": key, "alignments": n,
                              "cells": cells, "ms": ms, "gcups": cells / ms / 1e6,
                              "checksum": int(np.asarray(out["score"] if a.trace else out, np.int64).sum())}), flush=True)
        s.close()
        ix.close()


if __name__ == "__main__":
    main()
