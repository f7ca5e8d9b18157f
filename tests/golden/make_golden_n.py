#!/usr/bin/env python
"""Golden fixtures for queries that contain 'N' (the reference randomises every READ of an 'N'
during seeding, lambda_b200/csrc/n_random.hpp).  For the nucleotide, bisulfite and BLASTX cases the committed
q.fasta gets 'N's (4 % of the bases, plus a run of 3-6 'N's in every tenth read) -> <case>/qn.fasta, and the
unmodified reference binary produces <case>/n.<profile>.m8 + n.<profile>.funnel.json on the committed index."""
import gzip
import json
import os
import shutil
import subprocess
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
from make_golden import REF, parse_funnel  # noqa: E402

CASES = [("nucl", "searchn", ["none", "fast", "sensitive"]), ("bisulfite", "searchbs", ["none", "fast", "sensitive"]),
         ("blastx", "searchp", ["none"])]


def add_ns(src, dst, seed):
    rng = np.random.default_rng(seed)
    out = []
    k = 0
    for line in open(src):
        if line.startswith(">"):
            out.append(line)
            continue
        s = np.frombuffer(line.rstrip("\n").encode(), np.uint8).copy()
        s[rng.random(len(s)) < 0.04] = ord("N")
        if k % 10 == 3 and len(s) > 20:
            b = int(rng.integers(0, len(s) - 6))
            s[b:b + int(rng.integers(3, 7))] = ord("N")
        k += 1
        out.append(s.tobytes().decode() + "\n")
    open(dst, "w").writelines(out)


if __name__ == "__main__":
    for n, (case, cmd, profiles) in enumerate(CASES):
        src = os.path.join(HERE, case)
        add_ns(os.path.join(src, "q.fasta"), os.path.join(src, "qn.fasta"), 900 + n)
        with tempfile.TemporaryDirectory() as tmp:
            with gzip.open(os.path.join(src, "db.lba.gz"), "rb") as fi, open(os.path.join(tmp, "db.lba"), "wb") as fo:
                shutil.copyfileobj(fi, fo)
            for prof in profiles:
                o = os.path.join(tmp, prof + ".m8")
                c = [REF, cmd, "-q", os.path.join(src, "qn.fasta"), "-i", os.path.join(tmp, "db.lba"), "-o", o, "-t", "1",
                     "--version-to-outputfile", "0", "-v", "2"] + (["-p", prof] if prof != "none" else [])
                txt = subprocess.run(c, check=True, capture_output=True, text=True).stdout
                shutil.copy(o, os.path.join(src, f"n.{prof}.m8"))
                with open(os.path.join(src, f"n.{prof}.funnel.json"), "w") as f:
                    json.dump(parse_funnel(txt), f, indent=1)
                print(case, prof, sum(1 for _ in open(o)), "hits")
