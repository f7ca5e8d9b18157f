#!/usr/bin/env python
"""Sweeps engine knobs (env vars read at context creation) on the full searchp benchmark workload with
ONE index load:  python tools/sweep.py "STREAMS=2,DPX_OCC=32" "STREAMS=3,DPX_OCC=10" ...
Prints one JSON line per configuration (device ms per step, resident queries)."""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402


def main():
    import torch
    import lambda_b200
    wl = os.environ.get("SWEEP_WORKLOAD", "searchp")
    W = bench.WORKLOADS[wl]
    d = bench.ensure_index(wl, W["n_seqs"])
    ix = lambda_b200.Index.load(os.path.join(d, "db.lba"), device=0, keep_ids=False)
    q_ascii, qoffs = bench.make_queries(wl, d, W["n_queries"], W["qlen"], seed=1000)
    res = lambda_b200.encode(q_ascii, W["dom"])
    d_res = torch.from_numpy(res).cuda()
    d_offs = torch.from_numpy(qoffs.view(np.int64)).cuda()
    steps = int(os.environ.get("SWEEP_STEPS", 4))
    for cfg in sys.argv[1:]:
        for kv in cfg.split(","):
            k, v = kv.split("=")
            os.environ["LAMBDA_B200_" + k] = v
        s = lambda_b200.Searcher(ix, W["domain"])
        for _ in range(2):
            s.search(d_res, d_offs)
        ms, acc = 0.0, None
        for _ in range(steps):
            hits, st = s.search(d_res, d_offs)
            ms += float(st["ms_total"])
            acc = st.copy() if acc is None else acc
        print(json.dumps({"cfg": cfg, "ms_per_step": ms / steps, "hits": int(len(hits)),
                          "stage_ms_sum_over_streams": {k: float(acc[k]) for k in
                                                        ("ms_seed", "ms_extend_score", "ms_extend_trace", "ms_host")}}),
              flush=True)
        s.close()


if __name__ == "__main__":
    main()
