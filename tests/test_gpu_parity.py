"""Parity of the CUDA path with the CPU oracle and the reference's golden outputs (B200 only).
Everything goes through the C ABI (lambda_b200/liblambda_b200.so)."""
import os
import subprocess

import numpy as np
import pytest

import lambda_b200
import orc
from cases import CASE_PROFILES, CASES, FUNNEL, N_CASE_PROFILES, load_golden, query_alph, query_encoding
from lambda_b200 import synth
from lambda_b200._abi import HIT_DT, MATCH_DT

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.path.join(ROOT, "oracle", "_ref", "lambda3")
HIT_INT_FIELDS = ["q_id", "s_id", "q_start", "q_end", "s_start", "s_end", "q_len", "s_len", "score", "n_match",
                  "n_mismatch", "n_gap_open", "n_gap_ext", "n_positive", "aln_len", "q_frame", "s_frame"]


def _load(gdir, case, domain):
    path = os.path.join(gdir, case, "db.lba")
    ids, data, offs = lambda_b200.read_fasta(os.path.join(gdir, case, "q.fasta"))
    return path, ids, lambda_b200.encode(data, query_encoding(case, domain)), offs


def _pair(ix, o, case, domain, profile="none", **kw):
    """a Searcher and the oracle's parameter block for the same case (query alphabet included)"""
    s = lambda_b200.Searcher(ix, domain, profile, query_alph=query_alph(case), **kw)
    p = o.params(domain, profile)
    p.query_alph = query_alph(case)
    for k, v in kw.items():
        setattr(p, k, v)
    return s, p


def _sorted(m):
    return np.sort(m, order=list(m.dtype.names))


@pytest.mark.parametrize("case,domain", [(c, d) for c, d, _ in CASES])
def test_fm_rank_and_locate(golden_dir, case, domain):
    path = os.path.join(golden_dir, case, "db.lba")
    o = orc.Oracle(path)
    ix = lambda_b200.Index.load(path)
    rng = np.random.default_rng(1)
    n_rows = ix.desc["n_residues"] + 5 * ix.desc["n_seqs"]  # upper bound; clip below
    import ctypes as C
    Cv = np.ctypeslib.as_array(C.cast(o.desc.C, C.POINTER(C.c_uint64)), (o.desc.sigma + 1,))
    n_rows = int(Cv[-1])
    idx = np.concatenate([rng.integers(0, n_rows + 1, 5000), [0, n_rows, 64, 63, 128]]).astype(np.uint64)
    symb = rng.integers(0, o.desc.sigma, len(idx)).astype(np.uint8)
    assert (ix.rank(idx, symb) == o.rank(idx, symb)).all()
    rows = rng.integers(0, n_rows, 3000).astype(np.uint64)
    s1, p1 = ix.locate(rows)
    s2, p2 = o.locate(rows)
    assert (s1 == s2).all() and (p1 == p2).all()
    o.close()
    ix.close()


@pytest.mark.parametrize("mode", ["thread", "warp", "block", "spec"])
@pytest.mark.parametrize("case,domain,profile", CASE_PROFILES)
def test_seeding_matches_oracle(golden_dir, case, domain, profile, mode, monkeypatch):
    """all three seeding kernels (thread / warp / block per query) against the oracle, both phases"""
    monkeypatch.setenv("LAMBDA_B200_SEED", mode)
    path, ids, res, offs = _load(golden_dir, case, domain)
    o = orc.Oracle(path)
    ix = lambda_b200.Index.load(path)
    s, p = _pair(ix, o, case, domain, profile)
    for phase in (1, 2):
        m_gpu, st_gpu = s.seed(res, offs, phase)
        m_cpu, st_cpu = o.seed(p, res, offs, phase)
        assert len(m_gpu) == len(m_cpu)
        assert (_sorted(m_gpu) == _sorted(m_cpu)).all()
        for k in ("hits_after_seeding", "hits_failed_pre_extend"):
            assert int(st_gpu[k]) == int(st_cpu[k]), k
        # widen / sort / merge / unique on the same matches
        if len(m_cpu) and mode == "block":
            w_gpu, wst_gpu = s.merge(res, offs, m_cpu)
            w_cpu, wst_cpu = o.merge(p, res, offs, m_cpu)
            assert len(w_gpu) == len(w_cpu) and (w_gpu == w_cpu).all()  # same order: both sorted
            assert int(wst_gpu["hits_duplicate"]) == int(wst_cpu["hits_duplicate"])
    s.close(); ix.close(); o.close()


@pytest.mark.parametrize("mode", ["auto", "warp", "block", "spec", "thread"])
@pytest.mark.parametrize("case,domain", [("prot_diverged", 0), ("nucl", 1)])
def test_full_seed_hamming_search(golden_dir, case, domain, mode, monkeypatch):
    """--seed-half-exact 0: one mismatch anywhere in the seed, cursors in the level order of the reference's
    search_backtracking_with_buffers; seeding against the oracle, the whole search against the reference's file"""
    monkeypatch.setenv("LAMBDA_B200_SEED", mode)
    path, ids, res, offs = _load(golden_dir, case, domain)
    o = orc.Oracle(path)
    ix = lambda_b200.Index.load(path)
    s, p = _pair(ix, o, case, domain, seed_half_exact=0)
    m_gpu, st_gpu = s.seed(res, offs, 2)
    m_cpu, st_cpu = o.seed(p, res, offs, 2)
    assert len(m_gpu) == len(m_cpu) and (_sorted(m_gpu) == _sorted(m_cpu)).all()
    for k in ("hits_after_seeding", "hits_failed_pre_extend"):
        assert int(st_gpu[k]) == int(st_cpu[k]), k
    hits, st = s.search(res, offs)
    ref, funnel = load_golden(golden_dir, case, "nohalf")
    assert sorted(s.m8(hits, ids)) == sorted(ref)
    for k in FUNNEL:
        assert int(st[k]) == funnel[k], k
    # three mismatches are not implemented on the device: refused, not approximated
    with pytest.raises(lambda_b200.LambdaError):
        lambda_b200.Searcher(ix, domain, "none", seed_half_exact=0, opts=(11, 3, 3))
    s.close(); ix.close(); o.close()


@pytest.mark.parametrize("mode", ["auto", "warp", "block", "spec"])
@pytest.mark.parametrize("case,domain,half,name", [("prot_diverged", 0, 0, "nohalf_d2"), ("prot_diverged", 0, 1, "half_d2"),
                                                   ("nucl", 1, 1, "half_d2"), ("nucl", 1, 0, None)])
def test_two_mismatch_seeds(golden_dir, case, domain, half, name, mode, monkeypatch):
    """--seed-delta 2 on the device: the leaves of the two-mismatch search tree in the reference's production order,
    for half-exact seeds (breadth-first = lexicographic) and for Hamming distance over the whole seed (level order of
    search_backtracking_with_buffers); seeding against the oracle, the whole search against the reference's files"""
    monkeypatch.setenv("LAMBDA_B200_SEED", mode)
    path, ids, res, offs = _load(golden_dir, case, domain)
    o = orc.Oracle(path)
    ix = lambda_b200.Index.load(path)
    L, off = (11, 3) if domain == 0 else (14, 7)
    s, p = _pair(ix, o, case, domain, seed_half_exact=half, opts=(L, 2, off))
    p.opts.seed_length, p.opts.max_seed_dist, p.opts.seed_offset = L, 2, off
    m_gpu, st_gpu = s.seed(res, offs, 2)
    m_cpu, st_cpu = o.seed(p, res, offs, 2)
    assert len(m_gpu) == len(m_cpu) and (_sorted(m_gpu) == _sorted(m_cpu)).all()
    for k in ("hits_after_seeding", "hits_failed_pre_extend"):
        assert int(st_gpu[k]) == int(st_cpu[k]), k
    hits, st = s.search(res, offs)
    if name is not None:
        ref, funnel = load_golden(golden_dir, case, name)
        assert sorted(s.m8(hits, ids)) == sorted(ref)
        for k in FUNNEL:
            assert int(st[k]) == funnel[k], k
    else:
        h_cpu, st_cpu = o.search(p, res, offs)
        assert sorted(s.m8(hits, ids)) == sorted(o.m8(p, h_cpu, ids))
    s.close(); ix.close(); o.close()


@pytest.mark.parametrize("trace,prof", [("planes", "auto"), ("planes", "private"), ("planes", "shared"), ("scalar", "auto")])
@pytest.mark.parametrize("case,domain", [(c, d) for c, d, _ in CASES])
def test_extension_matches_oracle(golden_dir, case, domain, trace, prof, monkeypatch):
    """DP pass 1 scores and the full pass-2 records (coordinates + statistics) for both trace paths -- residue planes
    (default) and the scalar trace-byte kernel -- and for both ways of filling a warp: one query profile per warp
    (alignments of one query share it) and one per group (alignments of different queries side by side)"""
    monkeypatch.setenv("LAMBDA_B200_TRACE", trace)
    monkeypatch.setenv("LAMBDA_B200_PROFILE", prof)
    if prof == "private":
        monkeypatch.setenv("LAMBDA_B200_PLANE_MB", "1")  # several launch groups of residue planes
    path, ids, res, offs = _load(golden_dir, case, domain)
    o = orc.Oracle(path)
    ix = lambda_b200.Index.load(path)
    s, p = _pair(ix, o, case, domain)
    m, _ = o.seed(p, res, offs, 2)
    win, _ = o.merge(p, res, offs, m)
    # add ragged / degenerate windows: partial queries, 1-row windows, windows at subject ends
    extra = win[: min(len(win), 20)].copy()
    extra["qry_start"] = np.minimum(extra["qry_end"] - 1, 7)
    extra["subj_end"] = extra["subj_start"] + np.minimum(extra["subj_end"] - extra["subj_start"], 13)
    one = win[: min(len(win), 5)].copy()
    one["subj_end"] = one["subj_start"] + 1
    win = np.concatenate([win, extra, one]).astype(MATCH_DT)
    if domain == 2:
        # bisulfite: the same windows against the other conversion of the subject (other scoring matrix);
        # seeding never pairs these, the stage API may
        flip = win.copy()
        flip["subj_id"] ^= 1
        win = np.concatenate([win, flip]).astype(MATCH_DT)
    sc_gpu, _ = s.extend_scores(res, offs, win)
    sc_cpu, h_cpu = o.extend(p, res, offs, win, True)
    assert (sc_gpu == sc_cpu).all()
    h_gpu, _ = s.extend_trace(res, offs, win)
    for f in HIT_INT_FIELDS:
        assert (h_gpu[f] == h_cpu[f]).all(), f
    s.close(); ix.close(); o.close()


@pytest.mark.parametrize("trace,prof", [("planes", "auto"), ("planes", "private"), ("scalar", "auto")])
def test_extension_long_queries_multi_block(golden_dir, trace, prof, monkeypatch):
    """queries longer than one 32*K column block (boundary row path) and long merged windows"""
    monkeypatch.setenv("LAMBDA_B200_TRACE", trace)
    monkeypatch.setenv("LAMBDA_B200_PROFILE", prof)
    path = os.path.join(golden_dir, "prot_flat", "db.lba")
    o = orc.Oracle(path)
    ix = lambda_b200.Index.load(path)
    rng = np.random.default_rng(5)
    db, offs = synth.protein_db(500, seed=101)
    lens = np.diff(offs)
    long_ids = np.argsort(lens)[-6:]
    qs, qo = [], [0]
    for sid in long_ids:  # queries = mutated copies of the longest subjects (up to 2000 aa)
        qs.append(synth.mutate_protein(rng, db[offs[sid]:offs[sid + 1]], 0.2, 0.02))
        qo.append(qo[-1] + len(qs[-1]))
    res = lambda_b200.encode(np.concatenate(qs), 0)
    qo = np.array(qo, np.uint64)
    win = np.zeros(len(long_ids), MATCH_DT)
    win["qry_id"] = np.arange(len(long_ids))
    win["subj_id"] = long_ids
    win["qry_end"] = np.diff(qo)
    win["subj_end"] = lens[long_ids]
    s = lambda_b200.Searcher(ix, "protein")
    p = o.params(0)
    sc_gpu, _ = s.extend_scores(res, qo, win)
    sc_cpu, h_cpu = o.extend(p, res, qo, win, True)
    assert (sc_gpu == sc_cpu).all() and (sc_cpu > 500).all()
    h_gpu, _ = s.extend_trace(res, qo, win)
    for f in HIT_INT_FIELDS:
        assert (h_gpu[f] == h_cpu[f]).all(), f
    s.close(); ix.close(); o.close()


@pytest.mark.parametrize("prof", ["shared", "private"])
def test_score_kernel_all_length_classes(golden_dir, prof, monkeypatch):
    """DP pass 1 and pass 2 over every (T, K) class of the packed DPX kernels plus the scalar fallback (> 2048),
    with ragged windows (1 .. 2000 rows), related and unrelated pairs; "private": the groups of a warp carry
    alignments of DIFFERENT queries (mixed-query warps, the layout of the short-read searches)"""
    monkeypatch.setenv("LAMBDA_B200_PROFILE", prof)
    path = os.path.join(golden_dir, "prot_flat", "db.lba")
    o = orc.Oracle(path)
    ix = lambda_b200.Index.load(path)
    rng = np.random.default_rng(7)
    db, offs = synth.protein_db(500, seed=101)
    lens = np.diff(offs)
    # one length inside every class (columns = 2*T*K, kernels_dpx.cuh LGPU_DPX_CLASSES) plus the class edges
    qlens = [1, 2, 7, 33, 64, 65, 96, 100, 128, 129, 144, 150, 160, 170, 176, 190, 192, 193, 208, 220, 224, 240, 250,
             256, 257, 272, 288, 300, 304, 305, 320, 321, 336, 350, 352, 368, 384, 385, 416, 440, 448, 480, 500, 512,
             513, 640, 700, 768, 769, 896, 1000, 1024, 1025, 1280, 1500, 1536, 1537, 1792, 2000, 2048, 2049, 2300]
    qs, qo, wins = [], [0], []
    for qi, L in enumerate(qlens):
        sid = int(rng.integers(0, len(lens)))
        src = db[offs[sid]:offs[sid + 1]]
        reps = -(-L // len(src))
        seq = synth.mutate_protein(rng, np.tile(src, reps)[:L + 20], 0.15, 0.02)[:L]
        if len(seq) < L:
            seq = np.concatenate([seq, synth._random_residues(rng, L - len(seq))])
        qs.append(seq)
        qo.append(qo[-1] + L)
        for k in range(6):  # the source subject (related) + random subjects, random windows
            s2 = sid if k < 2 else int(rng.integers(0, len(lens)))
            a = int(rng.integers(0, lens[s2]))
            b = int(rng.integers(a + 1, lens[s2] + 1)) if k % 2 else int(lens[s2])
            if k == 0:
                a = 0
            wins.append((qi, s2, 0, L, a, b))
    res = lambda_b200.encode(np.concatenate(qs), 0)
    qo = np.array(qo, np.uint64)
    win = np.array(wins, dtype=MATCH_DT)
    s = lambda_b200.Searcher(ix, "protein")
    p = o.params(0)
    sc_gpu, st = s.extend_scores(res, qo, win)
    sc_cpu, _ = o.extend(p, res, qo, win, False)
    bad = np.nonzero(sc_gpu != sc_cpu)[0]
    assert len(bad) == 0, (win[bad[:5]], sc_gpu[bad[:5]], sc_cpu[bad[:5]])
    assert sc_cpu.max() > 1000
    # pass 2 over the same ragged set: every K class of the packed trace kernel + the scalar fallback
    _, h_cpu = o.extend(p, res, qo, win, True)
    h_gpu, _ = s.extend_trace(res, qo, win)
    for f in HIT_INT_FIELDS:
        bad = np.nonzero(h_gpu[f] != h_cpu[f])[0]
        assert len(bad) == 0, (f, win[bad[:3]], h_gpu[bad[:3]], h_cpu[bad[:3]])
    # other scoring schemes through the same kernels
    for kw in (dict(scoring_method=45, gap_open=-14, gap_extend=-2), dict(scoring_method=80, gap_open=-10, gap_extend=-1)):
        s2 = lambda_b200.Searcher(ix, "protein", **kw)
        p2 = o.params(0)
        for k, v in kw.items():
            setattr(p2, k, v)
        g, _ = s2.extend_scores(res, qo, win[:60])
        c, _ = o.extend(p2, res, qo, win[:60], False)
        assert (g == c).all()
        s2.close()
    s.close(); ix.close(); o.close()


@pytest.mark.parametrize("mode", ["auto", "thread", "warp", "spec", "plain"])
@pytest.mark.parametrize("case,domain,profile", CASE_PROFILES)
def test_search_reproduces_reference_output(golden_dir, case, domain, profile, mode, monkeypatch):
    if mode == "plain":  # the default kernels without the prefix table and without text-mode elongation
        monkeypatch.setenv("LAMBDA_B200_SEED_TEXT", "0")
        monkeypatch.setenv("LAMBDA_B200_SEED_PREFIX", "0")
        mode = "auto"
    monkeypatch.setenv("LAMBDA_B200_SEED", mode)
    path, ids, res, offs = _load(golden_dir, case, domain)
    ix = lambda_b200.Index.load(path)
    s = lambda_b200.Searcher(ix, domain, profile, query_alph=query_alph(case))
    hits, st = s.search(res, offs)
    ref, funnel = load_golden(golden_dir, case, profile)
    assert sorted(s.m8(hits, ids)) == sorted(ref)
    for k in FUNNEL:
        assert int(st[k]) == funnel[k], k
    assert int(st["kernel_launches"]) > 0
    # empty batch and a batch where nothing can seed
    h0, _ = s.search(np.zeros(0, np.uint8), np.zeros(1, np.uint64))
    assert len(h0) == 0
    h1, _ = s.search(np.zeros(3, np.uint8), np.array([0, 3], np.uint64))
    assert len(h1) == 0
    # idempotence: the same batch again gives the same records
    hits2, _ = s.search(res, offs)
    assert (hits2 == hits).all()
    s.close(); ix.close()


@pytest.mark.skipif(not os.path.exists(REF), reason="reference binary (oracle/_ref/lambda3) not built")
@pytest.mark.parametrize("family", [False, True])
def test_search_vs_live_reference_binary(tmp_path, family):
    """BASELINE config[0]-shaped case at reduced size, checked against the reference run on this box"""
    db, offs = synth.protein_db(8000, seed=31, family=family)
    q, qo = synth.protein_queries(db, offs, 400, 150, seed=32)
    synth.write_fasta(str(tmp_path / "db.fasta"), db, offs, "S")
    synth.write_fasta(str(tmp_path / "q.fasta"), q, qo, "Q")
    subprocess.check_call([REF, "mkindexp", "-d", str(tmp_path / "db.fasta"), "-i", str(tmp_path / "db.lba"),
                           "-v", "0"])
    subprocess.check_call([REF, "searchp", "-q", str(tmp_path / "q.fasta"), "-i", str(tmp_path / "db.lba"),
                           "-o", str(tmp_path / "ref.m8"), "--version-to-outputfile", "0", "-v", "0"])
    ix = lambda_b200.Index.load(str(tmp_path / "db.lba"))
    s = lambda_b200.Searcher(ix, "protein")
    ids, hits, st = s.search_fasta(str(tmp_path / "q.fasta"))
    ref = open(tmp_path / "ref.m8").read().splitlines(True)
    assert len(ref) >= 400
    assert sorted(s.m8(hits, ids)) == sorted(ref)
    s.close(); ix.close()


CLI = os.path.join(ROOT, "bin", "lambda3_b200")


@pytest.mark.skipif(not os.path.exists(CLI), reason="bin/lambda3_b200 not built")
@pytest.mark.parametrize("case,domain,profile", [("prot_flat", 0, "none"), ("prot_diverged", 0, "none"),
                                                 ("prot_family", 0, "sensitive"), ("nucl", 1, "none"),
                                                 ("bisulfite", 2, "none"), ("blastx", 0, "none"),
                                                 ("tblastn", 0, "sensitive"), ("tblastx", 0, "none")])
def test_cli_output_is_byte_identical_to_reference(golden_dir, tmp_path, case, domain, profile):
    """the host program keeps the lambda3 command line and reproduces the reference's -t 1 file, in order"""
    out = tmp_path / "out.m8"
    cmd = [CLI, ("searchp", "searchn", "searchbs")[domain], "-q", os.path.join(golden_dir, case, "q.fasta"), "-i",
           os.path.join(golden_dir, case, "db.lba"), "-o", str(out), "-t", "1", "--version-to-outputfile", "0", "-v", "2"]
    if profile != "none":
        cmd += ["-p", profile]
    txt = subprocess.run(cmd, check=True, capture_output=True, text=True).stdout
    ref, funnel = load_golden(golden_dir, case, profile)
    assert open(out).read() == "".join(ref)
    assert f"Number of total hits:                           {funnel['hits_final']}" in txt
    # refuses to overwrite, like the reference's create_new validator
    assert subprocess.run(cmd, capture_output=True).returncode != 0


@pytest.mark.skipif(not os.path.exists(CLI), reason="bin/lambda3_b200 not built")
@pytest.mark.parametrize("case,domain", [("prot_flat", 0), ("prot_diverged", 0), ("nucl", 1), ("bisulfite", 2),
                                         ("blastx", 0), ("tblastn", 0), ("tblastx", 0)])
def test_cli_output_columns_are_byte_identical_to_reference(golden_dir, tmp_path, case, domain):
    """--output-columns (src/search_options.hpp:224-232,710-760; SQ/blast/blast_tabular_out.h:248-560): custom
    column list incl. frames, % positives, unimplemented (n/i) and accession (n/a) columns, with .m9 comment lines"""
    from golden.make_golden_columns import COLUMNS
    cmd = [CLI, ("searchp", "searchn", "searchbs")[domain], "-q", "q.fasta", "-i", "db.lba", "-o", str(tmp_path / "cols.m9"),
           "-t", "1", "--version-to-outputfile", "0", "-v", "0", "--output-columns", COLUMNS]
    subprocess.run(cmd, check=True, capture_output=True, text=True, cwd=os.path.join(golden_dir, case))
    assert open(tmp_path / "cols.m9").read() == open(os.path.join(golden_dir, case, "cols.m9")).read()
    # unknown specifiers and the taxonomy columns are refused
    for bad in ("qseqid nonsense", "std staxids"):
        r = subprocess.run(cmd[:-1] + [bad, "-o", str(tmp_path / "bad.m8")], capture_output=True, text=True,
                           cwd=os.path.join(golden_dir, case))
        assert r.returncode != 0 and not os.path.exists(tmp_path / "bad.m8")


@pytest.mark.skipif(not os.path.exists(CLI), reason="bin/lambda3_b200 not built")
@pytest.mark.parametrize("case,domain", [("prot_flat", 0), ("nucl", 1), ("blastx", 0)])
@pytest.mark.parametrize("ext", ["fastq", "fq.gz", "fa.gz"])
def test_cli_reads_fastq_and_gzip_queries(golden_dir, tmp_path, case, domain, ext):
    """the reference reads its queries through bio::io (FASTA / FASTQ, transparently decompressed;
    src/search_algo.hpp:342-348) and `lambda3 searchp -q q.fastq.gz` reproduces the golden .m8 (checked when the
    reader was written); so must the host program"""
    import gzip
    ids, data, offs = lambda_b200.read_queries(os.path.join(golden_dir, case, "q.fasta"))
    seqs = [bytes(data[int(offs[i]):int(offs[i + 1])]).decode() for i in range(len(ids))]
    if ext.startswith("f") and "q" in ext.split(".")[0]:
        text = "".join(f"@{i}\n{s}\n+\n{'I' * len(s)}\n" for i, s in zip(ids, seqs))
    else:
        text = "".join(f">{i}\n{s}\n" for i, s in zip(ids, seqs))
    qf = tmp_path / ("q." + ext)
    with (gzip.open(qf, "wt") if ext.endswith(".gz") else open(qf, "w")) as f:
        f.write(text)
    out = tmp_path / "out.m8"
    cmd = [CLI, ("searchp", "searchn", "searchbs")[domain], "-q", str(qf), "-i", os.path.join(golden_dir, case, "db.lba"),
           "-o", str(out), "-t", "1", "--version-to-outputfile", "0", "-v", "0"]
    subprocess.run(cmd, check=True, capture_output=True, text=True)
    ref, _ = load_golden(golden_dir, case, "none")
    assert open(out).read() == "".join(ref)


@pytest.mark.skipif(not os.path.exists(CLI), reason="bin/lambda3_b200 not built")
@pytest.mark.parametrize("case,domain", [("prot_flat", 0), ("prot_diverged", 0), ("nucl", 1), ("bisulfite", 2),
                                         ("blastx", 0), ("tblastn", 0), ("tblastx", 0)])
def test_cli_m9_is_byte_identical_to_reference(golden_dir, case, domain):
    """.m9 = tabular with comment lines: program tag of all six BLAST modes, records only for queries with
    matches, footer with the record count; with and without the version string"""
    cwd = os.path.join(golden_dir, case)
    for name, extra in (("none.m9", ["--version-to-outputfile", "0"]), ("none.v1.m9", [])):
        out = os.path.join(cwd, "cli_" + name)
        if os.path.exists(out):
            os.remove(out)
        subprocess.run([CLI, ("searchp", "searchn", "searchbs")[domain], "-q", "q.fasta", "-i", "db.lba", "-o",
                        "cli_" + name, "-t", "1", "-v", "0", *extra], check=True, cwd=cwd)
        assert open(out).read() == open(os.path.join(cwd, name)).read(), name


@pytest.mark.skipif(not os.path.exists(CLI), reason="bin/lambda3_b200 not built")
@pytest.mark.parametrize("case,domain", [("prot_flat", 0), ("prot_family", 0), ("prot_diverged", 0), ("nucl", 1),
                                         ("bisulfite", 2), ("blastx", 0), ("tblastn", 0), ("tblastx", 0)])
def test_cli_sam_is_byte_identical_to_reference(golden_dir, case, domain):
    """.sam with the reference's default tags / clipping: CIGARs come from the device traceback (run-length
    operations of the gapped rows), sequences from the query frames; all six BLAST modes"""
    cwd = os.path.join(golden_dir, case)
    out = os.path.join(cwd, "cli_none.sam")
    if os.path.exists(out):
        os.remove(out)
    subprocess.run([CLI, ("searchp", "searchn", "searchbs")[domain], "-q", "q.fasta", "-i", "db.lba", "-o",
                    "cli_none.sam", "-t", "1", "-v", "0", "--version-to-outputfile", "0"], check=True, cwd=cwd)
    ours, ref = open(out).read().splitlines(), open(os.path.join(cwd, "none.sam")).read().splitlines()
    assert len(ours) == len(ref)
    for a, b in zip(ours, ref):
        assert a == b


@pytest.mark.skipif(not os.path.exists(CLI), reason="bin/lambda3_b200 not built")
@pytest.mark.parametrize("case,domain", [("prot_flat", 0), ("prot_family", 0), ("prot_diverged", 0), ("nucl", 1),
                                         ("bisulfite", 2), ("blastx", 0), ("tblastn", 0), ("tblastx", 0)])
def test_cli_m0_is_byte_identical_to_reference(golden_dir, case, domain):
    """.m0 = BLAST pairwise report: gapped rows from the device traceback, frame translation, statistics
    block, position arithmetic of all six BLAST modes"""
    cwd = os.path.join(golden_dir, case)
    out = os.path.join(cwd, "cli_none.m0")
    if os.path.exists(out):
        os.remove(out)
    subprocess.run([CLI, ("searchp", "searchn", "searchbs")[domain], "-q", "q.fasta", "-i", "db.lba", "-o",
                    "cli_none.m0", "-t", "1", "-v", "0", "--version-to-outputfile", "0"], check=True, cwd=cwd)
    ours, ref = open(out).read().splitlines(), open(os.path.join(cwd, "none.m0")).read().splitlines()
    for i, (a, b) in enumerate(zip(ours, ref)):
        assert a == b, (i, a, b)
    assert len(ours) == len(ref)


@pytest.mark.skipif(not os.path.exists(CLI), reason="bin/lambda3_b200 not built")
@pytest.mark.parametrize("case,domain", [("prot_flat", 0), ("prot_family", 0), ("nucl", 1), ("bisulfite", 2),
                                         ("blastx", 0), ("tblastn", 0), ("tblastx", 0)])
def test_cli_bam_content_is_identical_to_reference(golden_dir, case, domain):
    """.bam: the BGZF blocks may be cut and compressed differently, the uncompressed BAM stream (header,
    reference dictionary, binary records with CIGAR / 4-bit sequence / typed tags) must be the same bytes"""
    import gzip
    cwd = os.path.join(golden_dir, case)
    out = os.path.join(cwd, "cli_none.bam")
    if os.path.exists(out):
        os.remove(out)
    subprocess.run([CLI, ("searchp", "searchn", "searchbs")[domain], "-q", "q.fasta", "-i", "db.lba", "-o",
                    "cli_none.bam", "-t", "1", "-v", "0", "--version-to-outputfile", "0"], check=True, cwd=cwd)
    ours = gzip.open(out, "rb").read()
    ref = gzip.open(os.path.join(cwd, "none.bam"), "rb").read()
    assert ours[:4] == b"BAM\x01"
    if ours != ref:
        i = next(k for k in range(min(len(ours), len(ref))) if ours[k] != ref[k])
        raise AssertionError((len(ours), len(ref), i, ours[max(0, i - 40):i + 40], ref[max(0, i - 40):i + 40]))
    assert open(out, "rb").read()[-28:] == bytes.fromhex("1f8b08040000000000ff0600424302001b0003000000000000000000")


def test_cigar_runs_are_consistent(golden_dir, monkeypatch):
    """want_cigar through the API: the runs of every hit add up to its coordinates and statistics, and the
    packed-plane and the scalar trace paths emit the same runs"""
    for case, domain in (("prot_family", 0), ("nucl", 1)):
        path, ids, res, offs = _load(golden_dir, case, domain)
        ix = lambda_b200.Index.load(path)
        per_mode = {}
        for trace in ("planes", "scalar"):
            monkeypatch.setenv("LAMBDA_B200_TRACE", trace)
            s = lambda_b200.Searcher(ix, domain, "none", want_cigar=1)
            hits, st = s.search(res, offs)
            ops = s.last_cigar_ops
            assert len(hits) and ops is not None
            runs = []
            for h in hits:
                r = ops[h["cigar_off"]:h["cigar_off"] + h["cigar_len"]]
                kind, run = r & 3, r >> 2
                assert run[kind != 2].sum() == h["q_end"] - h["q_start"]
                assert run[kind != 1].sum() == h["s_end"] - h["s_start"]
                assert run.sum() == h["aln_len"] and (kind != 0).sum() == h["n_gap_open"]
                assert (kind[1:] != kind[:-1]).all()
                runs.append(r.tolist())
            per_mode[trace] = (hits.copy(), runs)
            s.close()
        assert per_mode["planes"][1] == per_mode["scalar"][1]
        monkeypatch.delenv("LAMBDA_B200_TRACE")
        # the same search without cigars returns the same records
        s2 = lambda_b200.Searcher(ix, domain, "none")
        h2, _ = s2.search(res, offs)
        for f in HIT_INT_FIELDS:
            assert (h2[f] == per_mode["planes"][0][f]).all(), f
        s2.close(); ix.close()


def test_multi_stream_split_is_invisible(golden_dir, monkeypatch):
    """large batches are cut into sub-batches running concurrently on several streams / host threads:
    hits, their order and every counter must equal the strictly serial run"""
    monkeypatch.setenv("LAMBDA_B200_MIN_SUBBATCH", "1024")
    path, ids, res, offs = _load(golden_dir, "prot_flat", 0)
    reps = 90  # 57 queries x 90 = 5130 > the 4096-query split threshold
    lens = np.diff(offs.astype(np.int64))
    big_res = np.tile(res, reps)
    big_offs = np.concatenate([[0], np.cumsum(np.tile(lens, reps))]).astype(np.uint64)
    big_ids = [f"{i}_{q}" for i in range(reps) for q in ids]
    ix = lambda_b200.Index.load(path)
    ref = None
    for streams in (1, 2, 3):
        s = lambda_b200.Searcher(ix, "protein", streams=streams)
        hits, st = s.search(big_res, big_offs)
        cur = (hits.copy(), {k: int(st[k]) for k in FUNNEL})
        if ref is None:
            ref = cur
            golden, funnel = load_golden(golden_dir, "prot_flat", "none")
            first = hits[hits["q_id"] < len(ids)]
            assert sorted(s.m8(first, ids)) == sorted(golden)
            assert cur[1]["hits_final"] == funnel["hits_final"] * reps
        else:
            assert (cur[0] == ref[0]).all() and cur[1] == ref[1]
        s.close()
    ix.close()


@pytest.mark.parametrize("band", [16, 64])
@pytest.mark.parametrize("case,domain", [("prot_flat", 0), ("prot_family", 0), ("nucl", 1)])
def test_window_band_override_matches_oracle(golden_dir, case, domain, band):
    """lgpu_params.window_band (band sweep of BASELINE configs[3]; non-parity with the reference binary, which has
    no such option): windows, hit records and funnel counters equal the oracle's with the same override"""
    path, ids, res, offs = _load(golden_dir, case, domain)
    ix = lambda_b200.Index.load(path)
    o = orc.Oracle(path)
    s, p = _pair(ix, o, case, domain, "none", window_band=band)
    for phase in (1, 2):
        m_cpu, _ = o.seed(p, res, offs, phase)
        w_gpu, _ = s.merge(res, offs, m_cpu)
        w_cpu, _ = o.merge(p, res, offs, m_cpu)
        assert len(w_gpu) == len(w_cpu) and (_sorted(w_gpu) == _sorted(w_cpu)).all(), phase
    h_gpu, st = s.search(res, offs)
    h_cpu, st2 = o.search(p, res, offs)
    assert sorted(s.m8(h_gpu, ids)) == sorted(o.m8(p, h_cpu, ids))
    for k in FUNNEL:
        assert int(st[k]) == int(st2[k]), k
    # and it is not the default: the reference-rule search sees other windows
    p0 = o.params(domain, "none")
    w0, _ = o.merge(p0, res, offs, o.seed(p0, res, offs, 1)[0])
    w1, _ = o.merge(p, res, offs, o.seed(p, res, offs, 1)[0])
    assert w0.tobytes() != w1.tobytes()
    s.close(); ix.close(); o.close()


@pytest.mark.parametrize("mode", ["auto", "thread", "warp", "block", "spec"])
@pytest.mark.parametrize("case,domain,profile", N_CASE_PROFILES)
def test_queries_with_n_reproduce_reference(golden_dir, case, domain, profile, mode, monkeypatch):
    """'N' in nucleotide queries (lambda_b200/csrc/n_random.hpp): every seeding kernel resolves the
    'N's like the reference's views::dna_n_to_random; seeds (both phases) equal the oracle's and the whole search
    reproduces the reference's golden output and funnel"""
    if mode != "auto":
        monkeypatch.setenv("LAMBDA_B200_SEED", mode)
    path = os.path.join(golden_dir, case, "db.lba")
    ids, data, offs = lambda_b200.read_queries(os.path.join(golden_dir, case, "qn.fasta"))
    res = lambda_b200.encode(data, query_encoding(case, domain))
    o = orc.Oracle(path)
    ix = lambda_b200.Index.load(path)
    s, p = _pair(ix, o, case, domain, profile)
    for phase in (1, 2):
        m_gpu, st_gpu = s.seed(res, offs, phase)
        m_cpu, st_cpu = o.seed(p, res, offs, phase)
        assert len(m_gpu) == len(m_cpu) and (_sorted(m_gpu) == _sorted(m_cpu)).all(), phase
        for k in ("hits_after_seeding", "hits_failed_pre_extend"):
            assert int(st_gpu[k]) == int(st_cpu[k]), k
    if mode != "block":  # the block-per-query kernel is a phase-2 tool for small active sets
        hits, st = s.search(res, offs)
        ref, funnel = load_golden(golden_dir, case, "n." + profile)
        assert sorted(s.m8(hits, ids)) == sorted(ref)
        for k in FUNNEL:
            assert int(st[k]) == funnel[k], k
    s.close(); ix.close(); o.close()


@pytest.mark.parametrize("case,domain,profile", [("prot_family", 0, "pairs-default"), ("prot_flat", 0, "none"),
                                                 ("nucl", 1, "sensitive"), ("bisulfite", 2, "none")])
def test_device_records_export_and_finalisation(golden_dir, case, domain, profile, monkeypatch):
    """_writeRecord runs on the device (sort / unique / top-N): order, records and every counter (incl. pairs and
    queries-with-hit) equal the oracle's host implementation; lgpu_ctx_export_hits hands the same records out of the
    device buffer (query ids rebased, doubles filled in by lgpu_hits_fill_scores) -- also when the call was cut into
    sub-batches; id_cutoff and max_matches exercise the compaction paths"""
    import torch
    path, ids, res, offs = _load(golden_dir, case, domain)
    ix = lambda_b200.Index.load(path)
    o = orc.Oracle(path)
    for kw in ({}, {"id_cutoff": 60, "max_matches": 3}, {"finalize": 0}):
        s, p = _pair(ix, o, case, domain, profile, **kw)
        h_gpu, st_gpu = s.search(res, offs)
        h_cpu, st_cpu = o.search(p, res, offs)
        if kw.get("finalize", 1):
            assert len(h_gpu) == len(h_cpu)
            for f in HIT_INT_FIELDS + ["phase", "bit_score", "evalue"]:
                assert (h_gpu[f] == h_cpu[f]).all(), (kw, f)
        else:
            assert sorted(s.m8(h_gpu, ids)) == sorted(o.m8(p, h_cpu, ids))
        for k in FUNNEL:
            assert int(st_gpu[k]) == int(st_cpu[k]), (kw, k)
        # the same records straight from the device
        buf = torch.zeros((len(h_gpu) + 1) * HIT_DT.itemsize, dtype=torch.uint8, device="cuda")
        assert s.export_hits(buf.data_ptr(), 0, 7) == len(h_gpu)  # too small a buffer: only the count
        n = s.export_hits(buf.data_ptr(), len(h_gpu) + 1, 1000)
        dev = s.fill_scores(buf[: n * HIT_DT.itemsize].cpu().numpy().view(HIT_DT).copy())
        assert n == len(h_gpu) and (dev["q_id"] == h_gpu["q_id"] + 1000).all()
        dev["q_id"] -= 1000
        assert (dev == h_gpu).all()
        s.close()
    # sub-batches: the export concatenates the workers' device buffers in query order
    monkeypatch.setenv("LAMBDA_B200_MIN_SUBBATCH", "8")
    s = lambda_b200.Searcher(ix, domain, profile, streams=3)
    h, _ = s.search(res, offs)
    buf = torch.zeros(max(len(h), 1) * HIT_DT.itemsize, dtype=torch.uint8, device="cuda")
    n = s.export_hits(buf.data_ptr(), len(h), 0)
    dev = s.fill_scores(buf[: n * HIT_DT.itemsize].cpu().numpy().view(HIT_DT).copy())
    assert n == len(h) and (dev == h).all()
    s.close(); ix.close(); o.close()


def test_queries_beyond_the_int16_column_limit_take_the_scalar_path(golden_dir):
    """32767 / 11 = 2978 columns is the longest BLOSUM62 query whose scores cannot leave the packed kernels' int16
    lanes; longer ones (and everything > 2048 columns) run on the 32-bit scalar kernels and agree with the oracle"""
    path = os.path.join(golden_dir, "prot_flat", "db.lba")
    o = orc.Oracle(path)
    ix = lambda_b200.Index.load(path)
    s = lambda_b200.Searcher(ix, "protein")
    p = o.params(0)
    lens = np.diff(np.ctypeslib.as_array(__import__("ctypes").cast(o.desc.seq_delims, __import__("ctypes").POINTER(
        __import__("ctypes").c_uint64)), (o.desc.n_seqs + 1,)).astype(np.int64))
    sid = int(np.argmax(lens))
    # queries of W (BLOSUM62 W:W = 11) against a subject window: only the stage API can pair them (no seed would)
    for n_w, ok in ((2900, True), (3100, True)):  # real subjects: scores stay small, both lengths take the scalar path
        res = lambda_b200.encode(np.frombuffer(b"W" * n_w, np.uint8), 0)
        offs = np.array([0, n_w], np.uint64)
        win = np.zeros(1, MATCH_DT)
        win["subj_id"], win["qry_end"], win["subj_end"] = sid, n_w, lens[sid]
        sc, _ = s.extend_scores(res, offs, win)
        sc_cpu, _ = o.extend(p, res, offs, win, False)
        assert (sc == sc_cpu).all()
    s.close(); ix.close(); o.close()


@pytest.mark.skipif(not os.path.exists(REF), reason="reference binary (oracle/_ref/lambda3) not built")
def test_long_self_hits_vs_live_reference_and_int16_limit(tmp_path):
    """queries beyond the packed kernels' 2048 columns against the live reference: a 3300-residue self hit whose score
    stays below 32767 must come out identical; a 3700-residue one exceeds the reference's int16 lanes and is refused"""
    rng = np.random.default_rng(77)
    db, offs = synth.protein_db(300, seed=78)
    hi = np.frombuffer(b"WCH", np.uint8)  # self scores 11, 9, 8
    longs = [hi[rng.integers(0, 3, n)] for n in (3300, 3700)]
    seqs = [db[offs[i]:offs[i + 1]] for i in range(len(offs) - 1)] + longs
    o2 = np.concatenate([[0], np.cumsum([len(x) for x in seqs])]).astype(np.int64)
    synth.write_fasta(str(tmp_path / "db.fasta"), np.concatenate(seqs), o2, "S")
    subprocess.check_call([REF, "mkindexp", "-d", str(tmp_path / "db.fasta"), "-i", str(tmp_path / "db.lba"), "-v", "0"])
    ix = lambda_b200.Index.load(str(tmp_path / "db.lba"))
    s = lambda_b200.Searcher(ix, "protein")
    # below the limit: identical to the reference
    q_ok = np.concatenate([seqs[5], longs[0]])
    qo = np.array([0, len(seqs[5]), len(q_ok)], np.int64)
    synth.write_fasta(str(tmp_path / "q.fasta"), q_ok, qo, "Q")
    subprocess.check_call([REF, "searchp", "-q", str(tmp_path / "q.fasta"), "-i", str(tmp_path / "db.lba"), "-o",
                           str(tmp_path / "ref.m8"), "--version-to-outputfile", "0", "-v", "0"])
    ids, hits, st = s.search_fasta(str(tmp_path / "q.fasta"))
    ref = open(tmp_path / "ref.m8").read().splitlines(True)
    assert sorted(s.m8(hits, ids)) == sorted(ref)
    assert hits["score"].max() > 25000 and hits["score"].max() <= 32767
    # above it: an error, not an answer
    res = lambda_b200.encode(longs[1], 0)
    with pytest.raises(lambda_b200.LambdaError) as e:
        s.search(res, np.array([0, len(res)], np.uint64))
    assert e.value.code == -4 and "32767" in str(e.value)
    # the context stays usable
    ids2, hits2, _ = s.search_fasta(str(tmp_path / "q.fasta"))
    assert (hits2 == hits).all()
    s.close(); ix.close()


def test_device_side_index_validation(golden_dir, tmp_path):
    """corruption that the host parser cannot see in O(1) -- an occurrence count that does not continue its predecessor,
    a sampled suffix-array entry naming a sequence that does not exist, a CSA super block ranking outside the sampled
    array -- is caught by the cross-checks that run on the device after the upload: LGPU_ERR_IO, not a wild read"""
    import struct
    from test_abi_cpu import _lba_layout
    src = open(os.path.join(golden_dir, "prot_flat", "db.lba"), "rb").read()
    lay = _lba_layout(src)
    sigma, block_bytes = 11, 80
    n_blocks = struct.unpack_from("<Q", src, lay["n_blocks"])[0]
    occ = lay["n_blocks"] + 8
    pos = occ + n_blocks * block_bytes
    n_super = struct.unpack_from("<Q", src, pos)[0]
    pos += 8 + n_super * sigma * 8 + (sigma + 1) * 8
    n_ssa = struct.unpack_from("<Q", src, pos)[0]
    ssa = pos + 8
    csa = ssa + n_ssa * 8 + 4 + 8  # version, super block count

    def load(data):
        path = str(tmp_path / "x.lba")
        open(path, "wb").write(bytes(data))
        ix = lambda_b200.Index.load(path)
        ix.close()

    load(src)
    cases = []
    d = bytearray(src)
    struct.pack_into("<I", d, occ + 7 * block_bytes + 4 * 3, struct.unpack_from("<I", d, occ + 7 * block_bytes + 4 * 3)[0] + 1)
    cases.append(("occ block counts", d))
    d = bytearray(src)
    struct.pack_into("<Q", d, ssa + 8 * 17, (0xffffffffffffffff << 55) & 0xffffffffffffffff | 5)  # sequence id 511 of 500
    cases.append(("sampled suffix array", d))
    d = bytearray(src)
    struct.pack_into("<Q", d, csa + 48 * 3, n_ssa + 1000)
    cases.append(("CSA bit vector", d))
    for what, data in cases:
        try:
            load(data)
        except lambda_b200.LambdaError as e:
            assert e.code == -2 and what in str(e), (what, str(e))
        else:
            raise AssertionError("accepted an index with corrupt " + what)
