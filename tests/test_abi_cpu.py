"""CPU-only checks of the boundary: the CUDA library loads, exports every symbol the header declares,
its host-side helpers agree with the reference's numbers, and compute calls fail loudly without a GPU."""
import ctypes as C
import os
import re

import numpy as np
import pytest

import lambda_b200
from lambda_b200 import api
from lambda_b200._abi import HIT_DT, Params

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _has_gpu():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


def test_library_exports_every_declared_symbol():
    lib = lambda_b200.load_library()
    hdr = open(os.path.join(ROOT, "include", "lambda_b200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    names = set(re.findall(r"\b(lgpu_[a-z0-9_]+)\s*\(", hdr))
    assert len(names) >= 20
    for n in sorted(names):
        assert hasattr(lib, n), f"{n} declared in include/lambda_b200.h but not exported"
    assert lib.lgpu_version() == 100


def test_struct_sizes_match_header():
    # compile-time layout of the PODs as the C compiler sees them
    import subprocess
    import tempfile
    src = ('#include "lambda_b200.h"\n#include <stdio.h>\nint main(){printf("%zu %zu %zu %zu %zu\\n",'
           'sizeof(lgpu_hit),sizeof(lgpu_match),sizeof(lgpu_stats),sizeof(lgpu_params),sizeof(lgpu_index_desc));}')
    with tempfile.TemporaryDirectory() as t:
        open(f"{t}/a.c", "w").write(src)
        subprocess.check_call(["gcc", "-I", os.path.join(ROOT, "include"), f"{t}/a.c", "-o", f"{t}/a"])
        out = subprocess.check_output([f"{t}/a"]).split()
    from lambda_b200._abi import MATCH_DT, STATS_DT, IndexDesc
    assert [int(x) for x in out] == [HIT_DT.itemsize, MATCH_DT.itemsize, STATS_DT.itemsize, C.sizeof(Params),
                                     C.sizeof(IndexDesc)]


def test_default_params_follow_reference_defaults():
    # src/search_options.hpp:309-337 and profiles :631-682
    p = api.default_params("protein")
    assert (p.opts0.seed_length, p.opts0.seed_offset, p.opts0.max_seed_dist) == (10, 5, 0)
    assert (p.opts.seed_length, p.opts.seed_offset, p.opts.max_seed_dist) == (11, 3, 1)
    assert (p.gap_open, p.gap_extend, p.max_matches, p.pre_scoring) == (-11, -1, 25, 2)
    assert p.pre_scoring_thresh == 2.0 and p.max_evalue == 1e-2
    n = api.default_params("nucleotide")
    assert (n.opts0.seed_length, n.opts0.seed_offset) == (14, 9) and n.pre_scoring_thresh == 1.4
    assert (n.gap_open, n.gap_extend, n.match, n.mismatch) == (-5, -2, 2, -3)
    s = api.default_params("protein", "pairs-sensitive")
    assert (s.opts0.seed_length, s.opts.seed_length, s.iterative_search, s.pre_scoring) == (9, 7, 0, 3)
    with pytest.raises(lambda_b200.LambdaError):
        api.default_params("protein", "no-such-profile")


def test_statistics_known_answers():
    lib = lambda_b200.load_library()
    p = api.default_params("protein")
    bits = C.c_double()
    assert lib.lgpu_bit_score(C.byref(p), 100, C.byref(bits)) == 0
    # (lambda * S - ln K) / ln 2 with BLOSUM62 11/1: lambda 0.267, K 0.041
    assert abs(bits.value - (0.267 * 100 - np.log(0.041)) / np.log(2)) < 1e-12
    ev = C.c_double()
    assert lib.lgpu_evalue(C.byref(p), 100, 150, 955200, C.byref(ev)) == 0
    assert 0 < ev.value < 1e-3
    mn = C.c_int32()
    assert lib.lgpu_min_raw_score(C.byref(p), 150, 955200, C.byref(mn)) == 0
    e_lo, e_hi = C.c_double(), C.c_double()
    lib.lgpu_evalue(C.byref(p), mn.value - 1, 150, 955200, C.byref(e_lo))
    lib.lgpu_evalue(C.byref(p), mn.value, 150, 955200, C.byref(e_hi))
    assert e_lo.value > p.max_evalue >= e_hi.value


def test_report_helpers_known_answers():
    """Karlin-Altschul block and substitution matrix used by the .m0 writer (SQ/blast/blast_statistics.h tables)"""
    lib = lambda_b200.load_library()
    p = api.default_params("protein")
    lam, k, h = C.c_double(), C.c_double(), C.c_double()
    assert lib.lgpu_ka_params(C.byref(p), C.byref(lam), C.byref(k), C.byref(h)) == 0
    assert (lam.value, k.value, h.value) == (0.267, 0.041, 0.14)
    m = np.zeros(32 * 32, np.int8)
    assert lib.lgpu_score_matrix(C.byref(p), C.c_void_p(m.ctypes.data)) == 0
    aa = "ABCDEFGHIJKLMNOPQRSTUVWXYZ*"
    sc = lambda a, b: int(m[aa.index(a) * 32 + aa.index(b)])
    assert (sc("W", "W"), sc("A", "A"), sc("A", "R"), sc("L", "I"), sc("C", "C")) == (11, 4, -1, 2, 9)  # BLOSUM62
    n = api.default_params("nucleotide")
    assert lib.lgpu_ka_params(C.byref(n), C.byref(lam), C.byref(k), C.byref(h)) == 0
    assert (lam.value, k.value, h.value) == (0.625, 0.41, 0.78)
    assert lib.lgpu_ka_params(None, C.byref(lam), C.byref(k), C.byref(h)) != 0


def test_m8_formatting_matches_golden_line():
    lib = lambda_b200.load_library()
    p = api.default_params("protein")
    h = np.zeros(1, HIT_DT)
    h["q_start"], h["q_end"], h["s_start"], h["s_end"] = 0, 150, 70, 220
    h["aln_len"], h["n_match"], h["n_mismatch"], h["n_gap_open"] = 151, 121, 28, 2
    h["evalue"], h["bit_score"], h["q_len"] = 6.2e-58, 216.4, 150
    buf = C.create_string_buffer(512)
    n = lib.lgpu_format_m8(C.byref(p), C.c_void_p(h.ctypes.data), b"Q0 descr", b"S1856", buf, 512)
    assert buf.raw[:n].decode() == "Q0\tS1856\t80.13\t151\t28\t2\t1\t150\t71\t220\t6e-58\t 216\n"


def test_custom_tabular_columns_match_golden_line():
    """first match line of tests/golden/tblastx/cols.m9 (made by the reference with --output-columns): TBLASTX hit in
    frames 3/3, coordinates un-translated to nucleotides, unimplemented columns print n/i, accessions n/a"""
    lib = lambda_b200.load_library()
    p = api.default_params("protein")
    h = np.zeros(1, HIT_DT)
    h["q_start"], h["q_end"], h["s_start"], h["s_end"] = 1, 68, 42, 109
    h["q_frame"], h["s_frame"], h["q_len"], h["s_len"] = 3, 3, 218, 570
    h["aln_len"], h["n_match"], h["n_mismatch"], h["n_positive"], h["score"] = 67, 60, 7, 61, 330
    h["evalue"], h["bit_score"] = 1.2e-33, 131.7
    from golden.make_golden_columns import COLUMNS
    cols = [lib.lgpu_tabular_column(c.encode()) for c in COLUMNS.split()]
    assert min(cols) >= 0 and lib.lgpu_tabular_column(b"nonsense") == -1
    arr = (C.c_uint32 * len(cols))(*cols)
    buf = C.create_string_buffer(4096)
    n = lib.lgpu_format_tabular(C.byref(p), C.c_void_p(h.ctypes.data), b"Q0", b"S104", arr, len(cols), buf, 4096)
    want = ("Q0\t218\tS104\t570\tQ0\tS104\t89.55\t67\t7\t0\t6\t206\t129\t329\t1e-33\t 132\t330\t67\t89.55\t60\t7\t61"
            "\t0\t0\t91.04\t3/3\t3\t3\tn/a\tn/a\tn/a\tn/i\tn/i\t1e-33\t 132\t6\t206\t129\t329\n")
    assert buf.raw[:n].decode() == want
    # "# Fields:" labels and the unsupported taxonomy columns
    assert lib.lgpu_tabular_column_label(0).decode().startswith("query id, subject id, % identity")
    assert lib.lgpu_tabular_column_label(lib.lgpu_tabular_column(b"ppos")).decode() == "% positives"
    assert lib.lgpu_tabular_column_name(13).decode() == "slen" and lib.lgpu_tabular_column_label(47) is None
    for name in (b"staxids", b"lcaid", b"lcataxid"):
        c = lib.lgpu_tabular_column(name)
        assert c >= 0 and not lib.lgpu_tabular_column_supported(c)
        one = (C.c_uint32 * 1)(c)
        assert lib.lgpu_format_tabular(C.byref(p), C.c_void_p(h.ctypes.data), b"Q0", b"S104", one, 1, buf, 4096) < 0
    # too small a buffer is an error, not a truncated line
    assert lib.lgpu_format_tabular(C.byref(p), C.c_void_p(h.ctypes.data), b"Q0", b"S104", arr, len(cols), buf, 20) < 0


@pytest.mark.skipif(_has_gpu(), reason="checks the no-GPU failure mode")
def test_compute_fails_loudly_without_gpu(golden_dir):
    with pytest.raises(lambda_b200.LambdaError) as e:
        lambda_b200.Index.load(os.path.join(golden_dir, "prot_flat", "db.lba"))
    assert e.value.code == -3 and "no CPU fallback" in str(e.value)


def test_lba_reader_rejects_garbage(tmp_path):
    bad = tmp_path / "x.lba"
    bad.write_bytes(b"\x01" + b"\0" * 100)
    lib = lambda_b200.load_library()
    h = C.c_void_p()
    assert lib.lgpu_lba_open(C.byref(h), str(bad).encode()) == -2
    assert b"generation" in lib.lgpu_last_error(None)
    assert lib.lgpu_lba_open(C.byref(h), b"/nonexistent.lba") == -2


def test_lba_reader_parses_golden_index(golden_dir):
    lib = lambda_b200.load_library()
    h = C.c_void_p()
    assert lib.lgpu_lba_open(C.byref(h), os.path.join(golden_dir, "prot_flat", "db.lba").encode()) == 0
    d = lib.lgpu_lba_desc(h).contents
    assert (d.sigma, d.sigma_bits, d.block_bytes, d.planes_offset) == (11, 4, 80, 48)
    assert d.n_seqs == 500 and d.sampling_rate == 5 and d.red_alph == 6
    lib.lgpu_lba_close(h)
    assert lib.lgpu_lba_open(C.byref(h), os.path.join(golden_dir, "nucl", "db.lba").encode()) == 0
    d = lib.lgpu_lba_desc(h).contents
    assert (d.sigma, d.sigma_bits, d.block_bytes, d.planes_offset) == (5, 3, 48, 24)
    lib.lgpu_lba_close(h)


def test_header_is_plain_c_and_the_example_client_runs(golden_dir, tmp_path):
    """include/lambda_b200.h must be consumable from C (the boundary is a C ABI): examples/minimal.c is compiled as strict
    C99 (-pedantic -Werror), linked against the library and run.  Without a device it must stop at lgpu_index_create with
    the no-CPU-fallback error (exit code 3); with one it searches two queries."""
    import shutil
    import subprocess
    gcc = shutil.which("gcc")
    if not gcc:
        pytest.skip("no C compiler")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    exe = str(tmp_path / "minimal")
    subprocess.run([gcc, "-std=c99", "-Wall", "-Wextra", "-pedantic", "-Werror", "-I" + os.path.join(root, "include"),
                    os.path.join(root, "examples", "minimal.c"), "-L" + os.path.join(root, "lambda_b200"), "-llambda_b200",
                    "-Wl,-rpath," + os.path.join(root, "lambda_b200"), "-o", exe], check=True, capture_output=True, text=True)
    r = subprocess.run([exe, os.path.join(golden_dir, "prot_flat", "db.lba")], capture_output=True, text=True)
    assert "lambda_b200 ABI version 100" in r.stdout and "index: 500 subjects" in r.stdout
    if _has_gpu():
        assert r.returncode == 0 and r.stdout.strip().endswith("hits")
    else:
        assert r.returncode == 3 and "no CPU fallback" in r.stderr


def test_integer_score_thresholds_are_exact_boundaries():
    """The device filter compares raw scores with one integer per query (a10 of SURVEY §8: e-value / bit-score test,
    src/search_algo.hpp:1252-1281).  For random query lengths, database sizes, cut-offs and scoring schemes the integer
    from lgpu_min_raw_score must be THE boundary of the double-precision tests: it passes, the score below it fails."""
    lib = lambda_b200.load_library()
    rng = np.random.default_rng(7)
    cases = 0
    for dom, tweaks in (("protein", [{}, {"scoring_method": 80, "gap_open": -10}, {"scoring_method": 45, "gap_open": -14, "gap_extend": -2}]),
                        ("nucleotide", [{}, {"match": 3, "mismatch": -4}])):
        for kw in tweaks:
            for _ in range(60):
                p = api.default_params(dom, **kw)
                p.max_evalue = float(rng.choice([1e-30, 1e-10, 1e-3, 1e-2, 1.0, 10.0, 1000.0]))
                p.min_bit_score = int(rng.choice([-1, -1, 30, 50, 200]))
                qlen = int(rng.integers(20, 5000))
                dblen = int(rng.choice([10_000, 955_200, 1_570_914_953, 200_000_000_000]))
                mn = C.c_int32()
                assert lib.lgpu_min_raw_score(C.byref(p), qlen, dblen, C.byref(mn)) == 0

                def passes(score):
                    e, b = C.c_double(), C.c_double()
                    lib.lgpu_evalue(C.byref(p), score, qlen, dblen, C.byref(e))
                    lib.lgpu_bit_score(C.byref(p), score, C.byref(b))
                    return e.value <= p.max_evalue and (p.min_bit_score < 0 or b.value >= p.min_bit_score)
                assert passes(mn.value), (dom, kw, qlen, dblen, p.max_evalue, p.min_bit_score, mn.value)
                if mn.value > 1:
                    assert not passes(mn.value - 1), (dom, kw, qlen, dblen, p.max_evalue, p.min_bit_score, mn.value)
                cases += 1
    assert cases == 300


def test_lba_reader_survives_truncated_and_corrupted_files(golden_dir, tmp_path):
    """the .lba parser works on an untrusted mmap: every truncation must end in LGPU_ERR_IO and random corruption of the
    length fields must never crash it (run in a child process so that a crash is a test failure, not a dead pytest)"""
    import subprocess
    import sys
    code = r"""
import ctypes as C, os, sys, numpy as np
sys.path.insert(0, %r)
import lambda_b200
lib = lambda_b200.load_library()
src = open(%r, 'rb').read()
tmp = %r
rng = np.random.default_rng(11)
vp = C.c_void_p
def try_open(data):
    path = os.path.join(tmp, 'x.lba')
    open(path, 'wb').write(data)
    h = vp()
    rc = lib.lgpu_lba_open(C.byref(h), path.encode())
    if rc == 0:
        lib.lgpu_lba_close(h)
    return rc
assert try_open(src) == 0
cuts = sorted(set([0, 1, 7, 8, 12, 13, 21, len(src) - 1] + [int(x) for x in rng.integers(0, len(src), 40)]))
for n in cuts:
    assert try_open(src[:n]) == -2, n          # LGPU_ERR_IO
assert try_open(src + b'xx') == -2              # trailing bytes
bad = 0
for _ in range(60):
    d = bytearray(src)
    pos = int(rng.integers(0, min(len(d), 4096)))   # the header region holds the vector lengths
    d[pos] ^= 1 << int(rng.integers(0, 8))
    if try_open(bytes(d)) != 0:
        bad += 1
print('ok', bad)
""" % (os.path.dirname(os.path.dirname(os.path.abspath(__file__))), os.path.join(golden_dir, "prot_flat", "db.lba"), str(tmp_path))
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True)
    assert r.returncode == 0 and r.stdout.startswith("ok"), (r.returncode, r.stdout[-500:], r.stderr[-1500:])


def _lba_layout(data):
    """offsets of the fields of a generation-0 FM .lba (csrc/lba_index.hpp parse())"""
    import struct
    pos, out = 13, {}

    def vec(name, elem):
        nonlocal pos
        n = struct.unpack_from("<Q", data, pos)[0]
        out[name] = (pos, pos + 8, n)  # (offset of the count, offset of the payload, count)
        pos += 8 + n * elem

    for name, elem in (("ids", 1), ("id_delims", 8), ("seqs", 1), ("seq_delims", 8), ("s_tax_ids", 4), ("s_tax_delims", 8),
                       ("taxon_parents", 4), ("taxon_heights", 1), ("taxon_names", 1), ("taxon_name_delims", 8)):
        vec(name, elem)
    out["occ_version"] = pos
    out["n_blocks"] = pos + 4
    return out


def test_lba_reader_rejects_crafted_inconsistencies(golden_dir, tmp_path):
    """counts that wrap when multiplied, delimiters that are not sorted or do not end at their payload, tax ids and
    parents outside the taxonomy: LGPU_ERR_IO, whatever the rest of the file looks like"""
    import struct
    lib = lambda_b200.load_library()

    def try_open(data):
        path = str(tmp_path / "x.lba")
        open(path, "wb").write(bytes(data))
        h = C.c_void_p()
        rc = lib.lgpu_lba_open(C.byref(h), path.encode())
        msg = lib.lgpu_last_error(None).decode() if rc else ""
        if rc == 0:
            lib.lgpu_lba_close(h)
        return rc, msg

    src = open(os.path.join(golden_dir, "tax", "db.lba"), "rb").read()
    lay = _lba_layout(src)
    assert try_open(src)[0] == 0
    assert lay["s_tax_ids"][2] > 0 and lay["taxon_parents"][2] > 0  # the fixture has a taxonomy

    def put64(d, off, v):
        struct.pack_into("<Q", d, off, v)

    def put32(d, off, v):
        struct.pack_into("<I", d, off, v)

    cases = {}
    d = bytearray(src); _, pay, n = lay["seq_delims"]
    a, b = struct.unpack_from("<QQ", d, pay + 8)
    put64(d, pay + 8, b); put64(d, pay + 16, a)                     # two sequence delimiters swapped
    cases["sequence delimiters are not sorted"] = d
    d = bytearray(src); _, pay, n = lay["id_delims"]
    put64(d, pay + 8 * (n - 1), struct.unpack_from("<Q", d, pay + 8 * (n - 1))[0] - 1)  # ids end before their payload
    cases["id delimiters do not match"] = d
    d = bytearray(src); _, pay, n = lay["s_tax_ids"]
    put32(d, pay, lay["taxon_parents"][2] + 5)                      # a subject's tax id beyond the taxonomy
    cases["subject tax id outside"] = d
    d = bytearray(src); _, pay, n = lay["taxon_parents"]
    put32(d, pay + 4 * (n - 1), 0x7fffffff)                         # a parent beyond the taxonomy
    cases["taxon parent outside"] = d
    d = bytearray(src); _, pay, n = lay["s_tax_delims"]
    put64(d, pay + 8 * (n - 1), lay["s_tax_ids"][2] + 3)            # tax id delimiters past their payload
    cases["taxonomy id delimiters do not match"] = d
    d = bytearray(src)
    block_bytes = 80                                                 # Li10: 11 counts + pad + 4 planes
    put64(d, lay["n_blocks"], (1 << 64) // block_bytes + 2)          # n_blocks * block_bytes wraps to a small number
    cases["truncated"] = d
    for what, data in cases.items():
        rc, msg = try_open(data)
        assert rc == -2 and what in msg, (what, rc, msg)
