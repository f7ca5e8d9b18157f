"""The golden cases (tests/golden/make_golden.py): (case dir, domain, profiles)."""
import json
import os

CASES = [
    ("prot_flat", 0, ["none", "fast", "sensitive", "pairs-default"]),
    ("prot_family", 0, ["none", "sensitive"]),
    ("prot_diverged", 0, ["none"]),
    ("nucl", 1, ["none", "fast", "sensitive"]),
    ("bisulfite", 2, ["none", "fast", "sensitive"]),
    ("blastx", 0, ["none", "sensitive"]),
    ("tblastn", 0, ["none", "sensitive"]),
    ("tblastx", 0, ["none"]),
]
# cases whose queries are nucleotides searched against a protein / translated index (BLASTX, TBLASTX):
# lgpu_params.query_alph = LGPU_ALPH_DNA5, residues encoded as dna5 ranks
DNA_QUERY_CASES = {"blastx", "tblastx"}


def query_alph(case):
    return 3 if case in DNA_QUERY_CASES else 0


def query_encoding(case, domain):
    """`domain` argument for encode(): 0 = aa27 ranks, otherwise dna5 ranks"""
    return 1 if case in DNA_QUERY_CASES else domain
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


# queries containing 'N' (tests/golden/make_golden_n.py): <case>/qn.fasta, n.<profile>.m8, n.<profile>.funnel.json
N_CASES = [("nucl", 1, ["none", "fast", "sensitive"]), ("bisulfite", 2, ["none", "fast", "sensitive"]),
           ("blastx", 0, ["none"])]
N_CASE_PROFILES = [(c, d, p) for c, d, ps in N_CASES for p in ps]
