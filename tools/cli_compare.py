#!/usr/bin/env python
"""End-to-end wall-clock of the two command-line programs on the benchmark workload (north star:
">= 20x the reference CPU searchp wall-clock"): bin/lambda3_b200 vs the unmodified oracle/_ref/lambda3,
same .lba, same query FASTA, outputs compared line by line (as sorted multisets: the reference's record
order depends on its thread count).

    python tools/cli_compare.py [--workload searchp] [--queries N] [--gpus 1]
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402


def run(cmd):
    t0 = time.time()
    p = subprocess.run(cmd, capture_output=True, text=True)
    dt = time.time() - t0
    if p.returncode != 0:
        sys.stderr.write(p.stdout[-2000:] + p.stderr[-2000:])
        raise SystemExit(f"{cmd[0]} failed with {p.returncode}")
    for line in p.stderr.splitlines():  # LAMBDA_B200_TRACE_TIMES=1: where a cold call spends its time
        if line.startswith("[lgpu"):
            sys.stderr.write(line + "\n")
    return dt, p.stdout + p.stderr


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="searchp")
    ap.add_argument("--queries", type=int, default=0)
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--reps", type=int, default=2)
    a = ap.parse_args()
    from lambda_b200 import synth
    import numpy as np
    W = bench.WORKLOADS[a.workload]
    d = bench.ensure_index(a.workload, W["n_seqs"])
    nq = a.queries or W["n_queries"]
    q, qo = bench.make_queries(a.workload, d, nq, W["qlen"], seed=1000)
    cores = os.cpu_count() or 1
    with tempfile.TemporaryDirectory() as tmp:
        qf = os.path.join(tmp, "q.fasta")
        synth.write_fasta(qf, q, qo.astype(np.int64), "Q")
        lba = os.path.join(d, "db.lba")
        ours, ref = os.path.join(tmp, "ours.m8"), os.path.join(tmp, "ref.m8")
        t_ours, t_ref = [], []
        for _ in range(a.reps):  # first repetition warms the page cache for both
            for f in (ours,):
                if os.path.exists(f):
                    os.remove(f)  # like the reference, lambda3_b200 refuses to overwrite its output
            dt, txt = run([os.path.join(ROOT, "bin", "lambda3_b200"), W["search"], "-q", qf, "-i", lba, "-o", ours,
                           "--gpus", str(a.gpus), "-v", "2"])
            t_ours.append(dt)
            detail = [l.strip() for l in txt.splitlines() if "Runtime total" in l or "GPU " in l or "Index mapped" in l]
        for _ in range(a.reps):
            if os.path.exists(ref):
                os.remove(ref)
            dt, txt = run([bench.REF, W["search"], "-q", qf, "-i", lba, "-o", ref, "-t", str(cores),
                           "--version-to-outputfile", "0", "-v", "0"])
            t_ref.append(dt)
        lo = sorted(open(ours).read().splitlines())
        lr = sorted(open(ref).read().splitlines())
        print(json.dumps({"workload": a.workload, "queries": nq, "gpus": a.gpus, "host_cores": cores,
                          "lambda3_b200_wall_s": t_ours, "lambda3_b200_detail": detail, "reference_wall_s": t_ref,
                          "speedup_best_of": min(t_ref) / min(t_ours), "lines_ours": len(lo), "lines_reference": len(lr),
                          "identical": lo == lr}))


if __name__ == "__main__":
    main()
