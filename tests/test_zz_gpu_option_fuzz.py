"""CUDA path against the oracle on random small databases with random NON-default options (the same generator and
seeds as tests/test_oracle_fuzz_cpu.py, where the oracle is held to the live reference binary): cut-offs, -n, seed
lengths / offsets / distances, --search0 off, adaptive seeding off, pre-scoring variants, BLOSUM 45 / 80, gap costs,
nucleotide scores.  Runs last (file name) because it was written after this round's GPU budget was used up: it has
not been executed on a GPU by its author."""
import os

import numpy as np
import pytest

import lambda_b200
import orc
from cases import FUNNEL
from test_oracle_fuzz_cpu import REF, random_case, random_case_other_modes, random_case_with_n

pytestmark = [pytest.mark.gpu, pytest.mark.skipif(not os.path.exists(REF), reason="reference binary (index builder) not built")]


@pytest.mark.parametrize("seed", [2, 4, 6, 7, 9, 10, 13, 14, 17, 21, 22, 23, 24, 25, 29, 30, 35, 39])
def test_cuda_path_equals_oracle_with_random_options(tmp_path, seed):
    tmp = str(tmp_path)
    dom, se, o, p, flags = random_case(seed, tmp)
    ids, data, qoffs = lambda_b200.read_queries(f"{tmp}/q.fasta")
    res = lambda_b200.encode(data, dom)
    ix = lambda_b200.Index.load(f"{tmp}/db.lba")
    kw = {k: getattr(p, k) for k in ("max_evalue", "max_matches", "min_bit_score", "id_cutoff", "adaptive_seeding",
                                     "iterative_search", "pre_scoring", "pre_scoring_thresh", "scoring_method", "gap_open",
                                     "gap_extend", "match", "mismatch")}
    kw["opts"] = (p.opts.seed_length, p.opts.max_seed_dist, p.opts.seed_offset)
    kw["opts0"] = (p.opts0.seed_length, p.opts0.max_seed_dist, p.opts0.seed_offset)
    s = lambda_b200.Searcher(ix, dom, "none", **kw)
    h_gpu, st = s.search(res, qoffs)
    h_cpu, st2 = o.search(p, res, qoffs)
    assert sorted(s.m8(h_gpu, ids)) == sorted(o.m8(p, h_cpu, ids)), flags
    for k in FUNNEL:
        assert int(st[k]) == int(st2[k]), (k, flags)
    s.close(); ix.close(); o.close()


@pytest.mark.parametrize("seed", list(range(100, 112)))
def test_cuda_path_equals_oracle_other_modes(tmp_path, seed):
    """bisulfite / BLASTX / TBLASTN / TBLASTX with random options"""
    tmp = str(tmp_path)
    dom, se, o, p, flags, enc = random_case_other_modes(seed, tmp)
    ids, data, qoffs = lambda_b200.read_queries(f"{tmp}/q.fasta")
    res = lambda_b200.encode(data, enc)
    ix = lambda_b200.Index.load(f"{tmp}/db.lba")
    kw = {k: getattr(p, k) for k in ("max_evalue", "max_matches", "id_cutoff", "adaptive_seeding", "iterative_search",
                                     "pre_scoring", "scoring_method", "query_alph")}
    kw["opts"] = (p.opts.seed_length, p.opts.max_seed_dist, p.opts.seed_offset)
    s = lambda_b200.Searcher(ix, dom, "none", **kw)
    h_gpu, st = s.search(res, qoffs)
    h_cpu, st2 = o.search(p, res, qoffs)
    assert sorted(s.m8(h_gpu, ids)) == sorted(o.m8(p, h_cpu, ids)), flags
    for k in FUNNEL:
        assert int(st[k]) == int(st2[k]), (k, flags)
    s.close(); ix.close(); o.close()


@pytest.mark.parametrize("seed", list(range(200, 208)))
def test_cuda_path_equals_oracle_with_n_and_random_seeding(tmp_path, seed):
    """'N' in nucleotide / bisulfite reads under random seeding options"""
    tmp = str(tmp_path)
    dom, se, o, p, flags = random_case_with_n(seed, tmp)
    ids, data, qoffs = lambda_b200.read_queries(f"{tmp}/q.fasta")
    res = lambda_b200.encode(data, 1)
    ix = lambda_b200.Index.load(f"{tmp}/db.lba")
    kw = {k: getattr(p, k) for k in ("max_evalue", "adaptive_seeding", "iterative_search")}
    kw["opts"] = (p.opts.seed_length, p.opts.max_seed_dist, p.opts.seed_offset)
    kw["opts0"] = (p.opts0.seed_length, p.opts0.max_seed_dist, p.opts0.seed_offset)
    s = lambda_b200.Searcher(ix, dom, "none", **kw)
    h_gpu, st = s.search(res, qoffs)
    h_cpu, st2 = o.search(p, res, qoffs)
    assert sorted(s.m8(h_gpu, ids)) == sorted(o.m8(p, h_cpu, ids)), flags
    for k in FUNNEL:
        assert int(st[k]) == int(st2[k]), (k, flags)
    s.close(); ix.close(); o.close()


@pytest.mark.parametrize("case,domain", [("prot_flat", 0), ("prot_family", 0), ("nucl", 1), ("bisulfite", 2), ("blastx", 0),
                                         ("tblastx", 0)])
def test_cuda_path_equals_oracle_pairs_sensitive(golden_dir, case, domain):
    """-p pairs-sensitive (the one profile without committed golden files) on the golden inputs"""
    from cases import query_alph, query_encoding
    path = os.path.join(golden_dir, case, "db.lba")
    ids, data, offs = lambda_b200.read_queries(os.path.join(golden_dir, case, "q.fasta"))
    res = lambda_b200.encode(data, query_encoding(case, domain))
    o = orc.Oracle(path)
    ix = lambda_b200.Index.load(path)
    s = lambda_b200.Searcher(ix, domain, "pairs-sensitive", query_alph=query_alph(case))
    p = o.params(domain, "pairs-sensitive")
    p.query_alph = query_alph(case)
    h_gpu, st = s.search(res, offs)
    h_cpu, st2 = o.search(p, res, offs)
    assert sorted(s.m8(h_gpu, ids)) == sorted(o.m8(p, h_cpu, ids))
    for k in FUNNEL:
        assert int(st[k]) == int(st2[k]), k
    s.close(); ix.close(); o.close()
