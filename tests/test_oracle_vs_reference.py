"""Pins the CPU oracle: it must reproduce the reference binary's committed outputs exactly
(tabular lines as a multiset AND in file order, plus every funnel counter)."""
import os

import numpy as np
import pytest

import orc
from cases import CASE_PROFILES, FUNNEL, N_CASE_PROFILES, load_golden, query_alph, query_encoding


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
                                                 ("nucl", 1, "nohalf", dict(seed_half_exact=0)),
                                                 ("prot_diverged", 0, "half_d2", dict(seed_half_exact=1, delta=2)),
                                                 ("nucl", 1, "half_d2", dict(seed_half_exact=1, delta=2))])
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


def test_window_band_override(golden_dir):
    """lgpu_params.window_band (the band sweep of BASELINE configs[3]): 0 keeps the reference's
    floor(sqrt(query length)) + 1 (_bandSize, src/search_misc.hpp:46-50) -- asked for explicitly it must
    reproduce the default; any other value pads the seed diagonal's window by exactly that many residues
    (src/search_algo.hpp:929-937) and the whole path stays consistent (more band -> the same or more cells)."""
    from lambda_b200._abi import MATCH_DT
    o = orc.Oracle(os.path.join(golden_dir, "prot_flat", "db.lba"))
    ids, data, offs = orc.read_fasta(os.path.join(golden_dir, "prot_flat", "q.fasta"))
    res = orc.encode(data, 0)
    lens = np.diff(offs.astype(np.int64))
    q = int(np.argmax(lens == lens[0]))  # any query; its length decides the default band
    qlen = int(lens[q])
    m = np.zeros(1, MATCH_DT)
    m["qry_id"], m["subj_id"], m["qry_start"], m["qry_end"], m["subj_start"], m["subj_end"] = q, 3, 10, 20, 200, 210
    p = o.params(0)
    w_def, _ = o.merge(p, res, offs, m)
    p.window_band = int(np.floor(np.sqrt(qlen))) + 1
    w_same, _ = o.merge(p, res, offs, m)
    assert w_def.tobytes() == w_same.tobytes()
    for band in (16, 32, 64):
        p.window_band = band
        w, _ = o.merge(p, res, offs, m)
        assert int(w["subj_start"][0]) == 190 - band and int(w["qry_end"][0]) == qlen
        assert int(w["subj_end"][0]) == 190 + qlen + band  # subject 3 is longer than that in this fixture
    # whole path: hits of the default band are reproduced with the explicit band on a same-length query subset
    same = np.nonzero(lens == qlen)[0][:8]
    sub_offs = np.zeros(len(same) + 1, np.uint64)
    sub = []
    for k, i in enumerate(same):
        sub.append(res[int(offs[i]):int(offs[i + 1])])
        sub_offs[k + 1] = sub_offs[k] + np.uint64(len(sub[-1]))
    sub = np.concatenate(sub)
    p = o.params(0)
    h0, st0 = o.search(p, sub, sub_offs)
    p.window_band = int(np.floor(np.sqrt(qlen))) + 1
    h1, st1 = o.search(p, sub, sub_offs)
    assert h0.tobytes() == h1.tobytes() and len(h0) > 0
    o.close()


@pytest.mark.parametrize("case,domain,profile", N_CASE_PROFILES)
def test_oracle_reproduces_reference_with_n_in_queries(golden_dir, case, domain, profile):
    """'N' in nucleotide queries: during seeding the reference replaces every READ of an 'N' by the next output of
    a per-view std::mt19937{0xDEADBEEF} (src/view_dna_n_to_random.hpp); restated in
    lambda_b200/csrc/n_random.hpp.  4 % 'N's + runs of 'N's; lines and every funnel counter must match."""
    o = orc.Oracle(os.path.join(golden_dir, case, "db.lba"))
    ids, data, offs = orc.read_fasta(os.path.join(golden_dir, case, "qn.fasta"))
    assert (data == ord("N")).sum() > 100
    res = orc.encode(data, query_encoding(case, domain))
    p = o.params(domain, profile)
    p.query_alph = query_alph(case)
    hits, st = o.search(p, res, offs)
    ref, funnel = load_golden(golden_dir, case, "n." + profile)
    assert sorted(o.m8(p, hits, ids)) == sorted(ref)
    for k in FUNNEL:
        assert int(st[k]) == funnel[k], k
    o.close()


def test_n_random_sequence_is_mt19937():
    """the packed table of n_random.hpp against a Mersenne twister seeded like the reference's view
    (numpy's RandomState seeds MT19937 with init_genrand like std::mt19937{seed})"""
    import ctypes as C
    rs = np.random.RandomState(0xDEADBEEF)
    want = [int(x) % 4 for x in rs.randint(0, 2**32, size=64, dtype=np.uint64)]
    lo, hi = 0xa3736c5835666461, 0xf83739b5e56c0330
    got = [((lo if k < 32 else hi) >> (2 * (k & 31))) & 3 for k in range(64)]
    assert got == want
