#!/usr/bin/env python
"""Golden outputs of the unmodified reference for the seed-search variants outside its profiles
(src/search_algo.hpp:484-494 -> FMC search/BacktrackingWithBuffers.h): Hamming distance over the WHOLE seed,
  <case>/nohalf.m8        --seed-half-exact 0                       (phase 2: 11-mers, one mismatch anywhere)
  <case>/nohalf_d2.m8     --seed-half-exact 0 --seed-delta 2        (two mismatches anywhere)
  <case>/half_d2.m8       --seed-delta 2                            (half-exact seeds, two mismatches in the second half)
plus the funnel counters, for the committed fixtures of two cases."""
import gzip
import json
import os
import shutil
import subprocess
import sys
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
from make_golden import REF, parse_funnel  # noqa: E402

for case, cmd in (("prot_diverged", "searchp"), ("nucl", "searchn")):
    src = os.path.join(HERE, case)
    with tempfile.TemporaryDirectory() as tmp:
        with gzip.open(os.path.join(src, "db.lba.gz"), "rb") as fi, open(os.path.join(tmp, "db.lba"), "wb") as fo:
            shutil.copyfileobj(fi, fo)
        for name, extra in (("nohalf", ["--seed-half-exact", "0"]),
                            ("nohalf_d2", ["--seed-half-exact", "0", "--seed-delta", "2"]),
                            ("half_d2", ["--seed-delta", "2"])):
            if case == "nucl" and name == "nohalf_d2":
                continue
            out = os.path.join(tmp, name + ".m8")
            txt = subprocess.run([REF, cmd, "-q", os.path.join(src, "q.fasta"), "-i", os.path.join(tmp, "db.lba"), "-o", out,
                                  "-t", "1", "--version-to-outputfile", "0", "-v", "2", *extra], check=True,
                                 capture_output=True, text=True).stdout
            shutil.copy(out, os.path.join(src, name + ".m8"))
            with open(os.path.join(src, name + ".funnel.json"), "w") as f:
                json.dump(parse_funnel(txt), f, indent=1)
            print(case, name, sum(1 for _ in open(out)), "hits")
