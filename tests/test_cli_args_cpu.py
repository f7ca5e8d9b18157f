"""Command-line surface of bin/lambda3_b200 that needs no GPU: option names / validation follow the reference's
parser (src/search_options.hpp), errors end with exit code 255 (the reference returns -1) and an ERROR line."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CLI = os.path.join(ROOT, "bin", "lambda3_b200")
pytestmark = pytest.mark.skipif(not os.path.exists(CLI), reason="bin/lambda3_b200 not built")


def run(*args, cwd=None):
    r = subprocess.run([CLI, *args], capture_output=True, text=True, cwd=cwd)
    return r.returncode, r.stdout, r.stderr


@pytest.mark.parametrize("args,msg", [
    (["mkindexp", "-d", "x.fasta"], "unknown sub-command"),
    (["searchp"], "-q and -i are required"),
    (["searchp", "-q", "q.fasta", "-i", "db.lba", "--no-such-option", "1"], "unknown option"),
    (["searchp", "-q", "q.fasta", "-i", "db.lba", "-p", "turbo"], "invalid profile"),
    (["searchp", "-q", "q.fasta", "-i", "db.lba", "-o", "out.txt"], "supported output formats"),
    (["searchp", "-q", "q.fasta", "-i", "db.lba", "-o", "out.bam.gz"], ".bam.gz is not supported"),
    (["searchp", "-q", "q.fasta", "-i", "db.lba", "--pre-scoring", "0"], "is not in range [1,10]"),
    (["searchp", "-q", "q.fasta", "-i", "db.lba", "--adaptive-seeding", "maybe"], "could not be parsed as type bool"),
    (["searchp", "-q", "q.fasta", "-i", "db.lba", "--sam-bam-clip", "medium"], "is not one of [hard,soft]"),
    (["searchp", "-q", "q.fasta", "-i", "db.lba", "--sam-bam-seq", "sometimes"], "is not one of [always,uniq,never]"),
    (["searchp", "-q", "q.fasta", "-i", "db.lba", "--sam-bam-tags", "AS XX"], 'Unknown column specifier "XX"'),
    (["searchp", "-q", "q.fasta", "-i", "db.lba", "--output-columns", "qseqid nonsense"], 'Unknown column specifier "nonsense"'),
    (["searchn", "-q", "q.fasta", "-i", "db.lba", "-a", "dna5"], "--input-alphabet is a searchp option"),
    (["searchp", "-q", "q.fasta", "-i", "db.lba", "-a", "rna"], "Invalid argument to --input-alphabet"),
    (["searchp", "-q", "q.fasta", "-i", "db.lba", "-e"], "missing value for -e"),
])
def test_argument_errors(tmp_path, args, msg):
    rc, out, err = run(*args, cwd=str(tmp_path))
    assert rc == 255 and err.startswith("ERROR: ") and msg in err, (rc, err)


def test_existing_output_is_not_overwritten(tmp_path):
    (tmp_path / "out.m8").write_text("precious\n")
    rc, _, err = run("searchp", "-q", "q.fasta", "-i", "db.lba", "-o", "out.m8", cwd=str(tmp_path))
    assert rc == 255 and "already exists" in err and (tmp_path / "out.m8").read_text() == "precious\n"


def test_help_pages(tmp_path):
    rc, out, _ = run("--help")
    assert rc == 0 and "searchp|searchn|searchbs" in out and "--search0" in out and "--sam-bam-tags" in out
    rc, out, _ = run("searchp", "-q", "q", "-i", "i", "--output-columns", "help", cwd=str(tmp_path))
    assert rc == 0 and "qseqid" in out and "staxids" in out and "lcataxid" in out and "btop" not in out
    rc, out, _ = run("searchp", "-q", "q", "-i", "i", "--sam-bam-tags", "help", cwd=str(tmp_path))
    assert rc == 0 and "AS\tbit score" in out and "lt\tlowest common ancestor taxonomy ID" in out


def test_missing_files_and_no_device(tmp_path):
    import gzip
    import shutil
    rc, _, err = run("searchp", "-q", "q.fasta", "-i", "nope.lba", "-o", "o.m8", cwd=str(tmp_path))
    assert rc == 255 and err.startswith("ERROR: ")
    with gzip.open(os.path.join(ROOT, "tests", "golden", "prot_flat", "db.lba.gz"), "rb") as fi, open(tmp_path / "db.lba", "wb") as fo:
        shutil.copyfileobj(fi, fo)
    rc, _, err = run("searchp", "-q", "nope.fasta", "-i", "db.lba", "-o", "o.m8", cwd=str(tmp_path))
    assert rc == 255 and "Could not open file nope.fasta" in err
    shutil.copy(os.path.join(ROOT, "tests", "golden", "prot_flat", "q.fasta"), tmp_path / "q.fasta")
    rc, _, err = run("searchn", "-q", "q.fasta", "-i", "db.lba", "-o", "o.m8", cwd=str(tmp_path))
    assert rc == 255  # protein index for a nucleotide search (or no device: both are errors, never a silent fallback)
    try:
        import torch
        has_gpu = torch.cuda.is_available()
    except Exception:
        has_gpu = False
    if not has_gpu:
        rc, _, err = run("searchp", "-q", "q.fasta", "-i", "db.lba", "-o", "o.m8", cwd=str(tmp_path))
        assert rc == 255 and "no CPU fallback" in err and not (tmp_path / "o.m8").exists()
