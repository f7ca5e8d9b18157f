"""The five output writers of the command-line host (.m8 .m9 .m0 .sam .bam, --output-columns) checked against the
reference binary's golden files WITHOUT a GPU: the CPU oracle (test infrastructure) computes the records incl. the
gapped rows as run-length operations, `lambda3_b200 --replay-hits FILE` only formats them (no search happens in that
mode).  This pins (a) the writers and (b) the oracle's runs against the reference's SAM CIGARs / pairwise rows.
The same files are produced from the CUDA path's records in tests/test_gpu_parity.py::test_cli_*."""
import gzip
import os
import struct
import subprocess

import numpy as np
import pytest

import orc
from cases import query_alph, query_encoding
from lambda_b200._abi import HIT_DT, STATS_DT

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CLI = os.path.join(ROOT, "bin", "lambda3_b200")
pytestmark = pytest.mark.skipif(not os.path.exists(CLI), reason="bin/lambda3_b200 not built")
CASES = [("prot_flat", 0), ("prot_family", 0), ("prot_diverged", 0), ("nucl", 1), ("bisulfite", 2), ("blastx", 0),
         ("tblastn", 0), ("tblastx", 0)]
SUB = ("searchp", "searchn", "searchbs")
HOOK_ENV = dict(os.environ, LAMBDA_B200_TEST_HOOKS="1")  # --replay-hits is refused without it


def write_hit_file(path, hits, stats, ops):
    with open(path, "wb") as f:
        f.write(b"LGPUHITS" + struct.pack("<QQ", len(hits), len(ops)))
        f.write(np.asarray([stats], STATS_DT).tobytes())
        f.write(np.ascontiguousarray(hits, HIT_DT).tobytes())
        f.write(np.ascontiguousarray(ops, np.uint32).tobytes())


@pytest.fixture(scope="module")
def replay(golden_dir):
    """per case: the oracle's records (profile none, want_cigar) as a hit file inside the unpacked golden dir"""
    made = {}

    def get(case, domain, qfile="q.fasta"):
        key = (case, qfile)
        if key not in made:
            cwd = os.path.join(golden_dir, case)
            o = orc.Oracle(os.path.join(cwd, "db.lba"))
            ids, data, offs = orc.read_fasta(os.path.join(cwd, qfile))
            p = o.params(domain, "none")
            p.query_alph = query_alph(case)
            p.want_cigar = 1
            hits, st = o.search(p, orc.encode(data, query_encoding(case, domain)), offs)
            path = os.path.join(cwd, qfile + ".hits")
            write_hit_file(path, hits, st, o.last_cigar_ops)
            o.close()
            made[key] = path
        return made[key]
    return get


def run_cli(golden_dir, case, domain, hits, out, *extra, qfile="q.fasta"):
    cwd = os.path.join(golden_dir, case)
    if os.path.exists(os.path.join(cwd, out)):
        os.remove(os.path.join(cwd, out))
    subprocess.run([CLI, SUB[domain], "-q", qfile, "-i", "db.lba", "-o", out, "-t", "1", "-v", "0", "--replay-hits", hits,
                    *extra], check=True, cwd=cwd, capture_output=True, env=HOOK_ENV)
    return os.path.join(cwd, out)


@pytest.mark.parametrize("case,domain", CASES)
def test_replayed_m8_m9_sam_m0_equal_reference(golden_dir, replay, case, domain):
    hits = replay(case, domain)
    v0 = ["--version-to-outputfile", "0"]
    for name, extra in (("none.m8", v0), ("none.m9", v0), ("none.v1.m9", []), ("none.sam", v0), ("none.m0", v0)):
        ref = os.path.join(golden_dir, case, name)
        if not os.path.exists(ref):
            continue  # (prot_family has no .m9 fixture)
        out = run_cli(golden_dir, case, domain, hits, "replay_" + name, *extra)
        ours, want = open(out).read().splitlines(), open(ref).read().splitlines()
        for i, (a, b) in enumerate(zip(ours, want)):
            assert a == b, (name, i, a, b)
        assert len(ours) == len(want), name


@pytest.mark.parametrize("case,domain", [c for c in CASES if c[0] != "prot_diverged"])
def test_replayed_bam_equals_reference(golden_dir, replay, case, domain):
    out = run_cli(golden_dir, case, domain, replay(case, domain), "replay_none.bam", "--version-to-outputfile", "0")
    ours = gzip.open(out, "rb").read()
    ref = gzip.open(os.path.join(golden_dir, case, "none.bam"), "rb").read()
    assert ours[:4] == b"BAM\x01" and ours == ref


@pytest.mark.parametrize("case,domain", [c for c in CASES if c[0] != "prot_family"])
def test_replayed_output_columns_equal_reference(golden_dir, replay, case, domain):
    from golden.make_golden_columns import COLUMNS
    out = run_cli(golden_dir, case, domain, replay(case, domain), "replay_cols.m9", "--version-to-outputfile", "0",
                  "--output-columns", COLUMNS)
    assert open(out).read() == open(os.path.join(golden_dir, case, "cols.m9")).read()


def test_replay_refuses_foreign_hit_files(golden_dir, replay, tmp_path):
    hits = replay("prot_flat", 0)
    cwd = os.path.join(golden_dir, "nucl")  # other index, other queries
    r = subprocess.run([CLI, "searchn", "-q", "q.fasta", "-i", "db.lba", "-o", str(tmp_path / "x.m8"), "--replay-hits", hits],
                       cwd=cwd, capture_output=True, text=True, env=HOOK_ENV)
    bad = tmp_path / "bad.hits"
    bad.write_bytes(b"NOTHITS!" + b"\0" * 64)
    r2 = subprocess.run([CLI, "searchn", "-q", "q.fasta", "-i", "db.lba", "-o", str(tmp_path / "y.m8"), "--replay-hits", str(bad)],
                        cwd=cwd, capture_output=True, text=True, env=HOOK_ENV)
    # and without the explicit opt-in the hook does not exist
    r3 = subprocess.run([CLI, "searchn", "-q", "q.fasta", "-i", "db.lba", "-o", str(tmp_path / "z.m8"), "--replay-hits", hits],
                        cwd=cwd, capture_output=True, text=True, env={k: v for k, v in os.environ.items() if k != "LAMBDA_B200_TEST_HOOKS"})
    assert r3.returncode == 255 and "performs no search" in r3.stderr
    assert r2.returncode == 255 and "malformed hit file" in r2.stderr
    assert r.returncode in (0, 255)  # ids may happen to be in range; a mismatch must not crash


@pytest.mark.parametrize("variant", ["opt_soft.sam", "opt_tags.sam", "opt_tags.bam", "opt_soft.bam"])
@pytest.mark.parametrize("case,domain", CASES)
def test_replayed_sam_dialect_options_equal_reference(golden_dir, replay, case, domain, variant):
    """--sam-bam-clip soft, --sam-bam-seq always|never, --sam-with-refheader, --sam-bam-tags with every non-taxonomy
    tag (src/search_options.hpp:276-370, src/search_output.hpp:116-298,482-719) for all six BLAST modes"""
    from golden.make_golden_sam_options import VARIANTS
    out = run_cli(golden_dir, case, domain, replay(case, domain), "replay_" + variant, "--version-to-outputfile", "0",
                  *VARIANTS[variant])
    ref = os.path.join(golden_dir, case, variant)
    if variant.endswith(".bam"):
        ours, want = gzip.open(out, "rb").read(), gzip.open(ref, "rb").read()
        if ours != want:
            i = next(k for k in range(min(len(ours), len(want))) if ours[k] != want[k])
            raise AssertionError((len(ours), len(want), i, ours[max(0, i - 60):i + 60], want[max(0, i - 60):i + 60]))
        return
    ours, want = open(out).read().splitlines(), open(ref).read().splitlines()
    for i, (a, b) in enumerate(zip(ours, want)):
        assert a == b, (i, a, b)
    assert len(ours) == len(want)


@pytest.mark.parametrize("name", ["none.m8", "none.m9", "none.sam", "none.m0"])
def test_gzip_compressed_outputs(golden_dir, replay, name):
    """the reference accepts .m0/.m8/.m9/.sam followed by .gz (src/search_options.hpp:210-214); content must be the same"""
    out = run_cli(golden_dir, "prot_flat", 0, replay("prot_flat", 0), "replay_" + name + ".gz", "--version-to-outputfile", "0")
    assert open(out, "rb").read(2) == b"\x1f\x8b"
    assert gzip.open(out, "rt").read() == open(os.path.join(golden_dir, "prot_flat", name)).read()


@pytest.mark.parametrize("name", ["tax.m9", "tax.sam", "tax.bam"])
def test_replayed_taxonomy_columns_and_tags_equal_reference(golden_dir, replay, name):
    """staxids / lcaid / lcataxid columns and the st / ls / lt tags: the subjects' tax ids from the index, the lowest
    common ancestor of every record (_writeRecord, src/search_algo.hpp:884-909; computeLCA, src/search_misc.hpp:86-112).
    Index with taxonomy built by the reference (tests/golden/make_golden_tax.py): subjects without tax id, with two
    tax ids, names with blanks."""
    from golden.make_golden_tax import COLUMNS, TAGS
    extra = ["--output-columns", COLUMNS] if name.endswith(".m9") else ["--sam-bam-tags", TAGS]
    out = run_cli(golden_dir, "tax", 0, replay("tax", 0), "replay_" + name, "--version-to-outputfile", "0", *extra)
    ref = os.path.join(golden_dir, "tax", name)
    if name.endswith(".bam"):
        assert gzip.open(out, "rb").read() == gzip.open(ref, "rb").read()
    else:
        ours, want = open(out).read().splitlines(), open(ref).read().splitlines()
        for i, (a, b) in enumerate(zip(ours, want)):
            assert a == b, (i, a, b)
        assert len(ours) == len(want)
    assert "root" in open(os.path.join(golden_dir, "tax", "tax.m9")).read()  # the fixture has non-trivial LCAs


def test_taxonomy_columns_need_an_index_with_taxonomy(golden_dir, replay, tmp_path):
    for extra, msg in ((["--output-columns", "std staxids"], "does not contain taxonomic information"),
                       (["--output-columns", "std lcaid"], "does not contain taxonomic information")):
        r = subprocess.run([CLI, "searchp", "-q", "q.fasta", "-i", "db.lba", "-o", str(tmp_path / "x.m8"), "--replay-hits",
                            replay("prot_flat", 0), *extra], cwd=os.path.join(golden_dir, "prot_flat"), capture_output=True,
                           text=True, env=HOOK_ENV)
        assert r.returncode == 255 and msg in r.stderr


@pytest.mark.parametrize("out", ["out_full.m8", "out_full.sam", "out_full.bam", "out_full.m8.gz"])
def test_failed_writes_are_errors_not_truncated_files(golden_dir, replay, out):
    """a write error (here: the file-size limit, which behaves like a full disk) must end in a non-zero exit code and must
    not leave a truncated output file behind"""
    import resource
    import signal
    case, domain = "prot_family", 0
    hits = replay(case, domain)
    cwd = os.path.join(golden_dir, case)
    if os.path.exists(os.path.join(cwd, out)):
        os.remove(os.path.join(cwd, out))

    def limit():
        signal.signal(signal.SIGXFSZ, signal.SIG_IGN)  # write() fails with EFBIG instead of killing the process
        resource.setrlimit(resource.RLIMIT_FSIZE, (512, 512))

    r = subprocess.run([CLI, SUB[domain], "-q", "q.fasta", "-i", "db.lba", "-o", out, "-t", "1", "-v", "0", "--replay-hits", hits],
                       cwd=cwd, capture_output=True, env=HOOK_ENV, preexec_fn=limit, text=True)
    assert r.returncode != 0
    assert "error while writing" in (r.stdout + r.stderr)
    assert not os.path.exists(os.path.join(cwd, out))
    # the same command without the limit works
    path = run_cli(golden_dir, case, domain, hits, out)
    assert os.path.getsize(path) > 512
