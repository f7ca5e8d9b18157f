"""Pins the CPU oracle: it must reproduce the reference binary's committed outputs exactly
(tabular lines as a multiset AND in file order, plus every funnel counter)."""
import os

import numpy as np
import pytest

import orc
from cases import CASE_PROFILES, FUNNEL, load_golden, query_alph, query_encoding


@pytest.mark.parametrize("case,domain,profile", CASE_PROFILES)
def test_oracle_reproduces_reference(golden_dir, case, domain, profile):
    o = orc.Oracle(os.path.join(golden_dir, case, "db.lba"))
    ids, data, offs = orc.read_fasta(os.path.join(golden_dir, case, "q.fasta"))
    res = orc.encode(data, query_encoding(case, domain))
    p = o.params(domain, profile)
    p.query_alph = query_alph(case)
    hits, st = o.search(p, res, offs)
    lines = o.m8(p, hits, ids)
    ref, funnel = load_golden(golden_dir, case, profile)
    assert sorted(lines) == sorted(ref)
    for k in FUNNEL:
        assert int(st[k]) == funnel[k], k
    # multi-threaded sharding must not change anything
    hits2, st2 = o.search(p, res, offs, threads=3)
    assert sorted(o.m8(p, hits2, ids)) == sorted(ref)
    o.close()


@pytest.mark.parametrize("case,domain,name,kw", [("prot_diverged", 0, "nohalf", dict(seed_half_exact=0)),
                                                 ("prot_diverged", 0, "nohalf_d2", dict(seed_half_exact=0, delta=2)),
                                                 ("nucl", 1, "nohalf", dict(seed_half_exact=0))])
def test_oracle_reproduces_reference_seed_variants(golden_dir, case, domain, name, kw):
    """Hamming distance over the whole seed (search_backtracking_with_buffers), delta 1 and 2"""
    o = orc.Oracle(os.path.join(golden_dir, case, "db.lba"))
    ids, data, offs = orc.read_fasta(os.path.join(golden_dir, case, "q.fasta"))
    res = orc.encode(data, query_encoding(case, domain))
    p = o.params(domain, "none")
    p.seed_half_exact = kw["seed_half_exact"]
    if "delta" in kw:
        p.opts.max_seed_dist = kw["delta"]
    hits, st = o.search(p, res, offs)
    ref, funnel = load_golden(golden_dir, case, name)
    assert sorted(o.m8(p, hits, ids)) == sorted(ref)
    for k in FUNNEL:
        assert int(st[k]) == funnel[k], k
    o.close()


def test_fm_primitives_against_text(golden_dir):
    """rank / locate against brute force on the text reconstructed from the stored sequences"""
    o = orc.Oracle(os.path.join(golden_dir, "prot_flat", "db.lba"))
    d = o.desc
    n_rows = int(np.ctypeslib.as_array(__import__("ctypes").cast(d.C, __import__("ctypes").POINTER(__import__("ctypes").c_uint64)), (d.sigma + 1,))[-1])
    rng = np.random.default_rng(0)
    rows = rng.integers(0, n_rows, 500).astype(np.uint64)
    subj, pos = o.locate(rows)
    assert (subj < d.n_seqs).all()
    # rank is monotone and consistent with C: rank(n_rows, s) - rank(0, s) == count(s)
    for s in range(1, d.sigma):
        lo = o.rank(np.array([0], np.uint64), np.array([s], np.uint8))[0]
        hi = o.rank(np.array([n_rows], np.uint64), np.array([s], np.uint8))[0]
        Cv = np.ctypeslib.as_array(__import__("ctypes").cast(d.C, __import__("ctypes").POINTER(__import__("ctypes").c_uint64)), (d.sigma + 1,))
        assert lo == Cv[s] and hi == Cv[s + 1]
    o.close()
