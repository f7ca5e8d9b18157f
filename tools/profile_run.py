#!/usr/bin/env python
"""Runs the benchmark workload's search a few times on ONE stream, nothing else: the process to put under
ncu (python tools/profile_run.py [workload] [n_searches])."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402


def main():
    import torch
    import lambda_b200
    wl = sys.argv[1] if len(sys.argv) > 1 else "searchp"
    n = int(sys.argv[2]) if len(sys.argv) > 2 else 2
    W = bench.WORKLOADS[wl]
    d = bench.ensure_index(wl, W["n_seqs"])
    ix = lambda_b200.Index.load(os.path.join(d, "db.lba"), device=0, keep_ids=False)
    q_ascii, qoffs = bench.make_queries(wl, d, W["n_queries"], W["qlen"], seed=1000)
    res = lambda_b200.encode(q_ascii, W["dom"])
    d_res = torch.from_numpy(res).cuda()
    d_offs = torch.from_numpy(qoffs.view(np.int64)).cuda()
    s = lambda_b200.Searcher(ix, W["domain"], streams=1)
    for _ in range(n):
        hits, st = s.search(d_res, d_offs)
    print(len(hits), float(st["ms_total"]))


if __name__ == "__main__":
    main()
