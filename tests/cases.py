"""The golden cases (tests/golden/make_golden.py): (case dir, domain, profiles)."""
import json
import os

CASES = [
    ("prot_flat", 0, ["none", "fast", "sensitive", "pairs-default"]),
    ("prot_family", 0, ["none", "sensitive"]),
    ("prot_diverged", 0, ["none"]),
    ("nucl", 1, ["none", "fast", "sensitive"]),
]
CASE_PROFILES = [(c, d, p) for c, d, ps in CASES for p in ps]
FUNNEL = ["hits_after_seeding", "hits_failed_pre_extend", "hits_failed_evalue", "hits_failed_bitscore",
          "hits_failed_identity", "hits_duplicate", "hits_duplicate2", "hits_abundant", "hits_final", "pairs",
          "qrys_with_hit"]


def load_golden(gdir, case, profile):
    with open(os.path.join(gdir, case, profile + ".m8")) as f:
        lines = f.read().splitlines(True)
    with open(os.path.join(gdir, case, profile + ".funnel.json")) as f:
        funnel = json.load(f)
    return lines, funnel
