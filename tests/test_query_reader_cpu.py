"""Query file formats of the command-line host (csrc/query_reader.hpp) and of the Python binding
(lambda_b200.read_queries): FASTA / FASTQ, plain or gzip-compressed, with the rules of the reference's
bio::io::seq::reader (BIO-IO format/fasta_input_handler.hpp, format/fastq_input_handler.hpp).  No GPU:
`lambda3_b200 dumpq` parses the file exactly like a search would and prints the records."""
import gzip
import os
import subprocess

import numpy as np
import pytest

import lambda_b200
from lambda_b200._abi import AA27, DNA5

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CLI = os.path.join(ROOT, "bin", "lambda3_b200")
GOLDEN = os.path.join(ROOT, "tests", "golden")
pytestmark = pytest.mark.skipif(not os.path.exists(CLI), reason="bin/lambda3_b200 not built")


def dumpq(path, *extra):
    r = subprocess.run([CLI, "dumpq", "-q", str(path), *extra], capture_output=True, text=True)
    return r.returncode, r.stdout, r.stderr


def records_of(stdout):
    lines = stdout.splitlines()
    alph = lines[0].split("\t")[1]
    return alph, [tuple(l.split("\t")) if "\t" in l else (l, "") for l in lines[1:]]


def write_variants(tmp_path, ids, seqs):
    """the same records as multi-line FASTA, FASTQ and their gzip-compressed copies"""
    fa = "".join(f">{i}\n" + "".join(s[k:k + 60] + "\n" for k in range(0, len(s), 60)) for i, s in zip(ids, seqs))
    fq = "".join(f"@{i}\n{s}\n+\n{'I' * len(s)}\n" for i, s in zip(ids, seqs))
    paths = []
    for name, text in (("q.fasta", fa), ("q.fq", fq)):
        p = tmp_path / name
        p.write_text(text)
        paths.append(p)
        with gzip.open(str(p) + ".gz", "wt") as f:
            f.write(text)
        paths.append(tmp_path / (name + ".gz"))
    return paths


@pytest.mark.parametrize("case,alph", [("prot_flat", "aminoacid"), ("nucl", "dna5"), ("blastx", "dna5")])
def test_all_formats_give_the_same_records(tmp_path, case, alph):
    ids, data, offs = lambda_b200.read_queries(os.path.join(GOLDEN, case, "q.fasta"))
    seqs = [bytes(data[int(offs[i]):int(offs[i + 1])]).decode() for i in range(len(ids))]
    rc, out, _ = dumpq(os.path.join(GOLDEN, case, "q.fasta"))
    assert rc == 0
    a0, rec0 = records_of(out)
    assert a0 == alph
    assert rec0 == list(zip(ids, seqs))  # the golden queries hold only letters of their alphabet
    for p in write_variants(tmp_path, ids, seqs):
        rc, out, err = dumpq(p)
        assert rc == 0, err
        assert records_of(out) == (a0, rec0), p
        ids2, data2, offs2 = lambda_b200.read_queries(str(p))
        assert ids2 == ids and np.array_equal(data2, data) and np.array_equal(offs2, offs), p


def test_fasta_details(tmp_path):
    """';' id lines, ids kept whole, digits and white space dropped, CRLF, lower case, U -> T, unknown -> N"""
    p = tmp_path / "x.fa"
    p.write_bytes(b">s1 some description\r\nACGT acgt\r\n12 NNU\r\n\r\n;s2\nAC\n>s3\nAXR\n\n\n")
    rc, out, _ = dumpq(p)
    assert rc == 0
    assert records_of(out) == ("dna5", [("s1 some description", "ACGTACGTNNT"), ("s2", "AC"), ("s3", "ANN")])
    ids, data, offs = lambda_b200.read_queries(str(p))
    assert ids == ["s1 some description", "s2", "s3"]
    ranks = lambda_b200.encode(data, 1)
    assert "".join(DNA5[r] for r in ranks) == "ACGTACGTNNTACANN"
    assert list(offs) == [0, 11, 13, 16]


def test_amino_acid_detection_and_override(tmp_path):
    p = tmp_path / "x.faa"
    p.write_text(">p1\nMKV*LX\nbzj\n")
    rc, out, _ = dumpq(p)
    assert records_of(out) == ("aminoacid", [("p1", "MKV*LXBZJ")])
    q = tmp_path / "y.fasta"
    q.write_text(">d1\nACGTNN\n")  # looks like DNA; -a aminoacid reads it as protein
    assert records_of(dumpq(q)[1])[0] == "dna5"
    assert records_of(dumpq(q, "-a", "aminoacid")[1]) == ("aminoacid", [("d1", "ACGTNN")])
    ids, data, _ = lambda_b200.read_queries(str(p))
    assert "".join(AA27[r] for r in lambda_b200.encode(data, 0)) == "MKV*LXBZJ"


def test_fastq_allows_empty_sequences(tmp_path):
    p = tmp_path / "x.fastq"
    p.write_text("@r1 x\nACGTN\n+\nIIIII\n@r2\n\n+r2\n\n")
    rc, out, _ = dumpq(p)
    assert rc == 0 and records_of(out) == ("dna5", [("r1 x", "ACGTN"), ("r2", "")])
    ids, data, offs = lambda_b200.read_queries(str(p))
    assert ids == ["r1 x", "r2"] and list(offs) == [0, 5, 5]


@pytest.mark.parametrize("name,content,msg", [
    ("a.txt", b">s\nACGT\n", "extension is not handled"),
    ("a.fq", b"@r1\nACGTN\n+\nIIII\n", "Size mismatch between sequence (5) and qualities (4)"),
    ("b.fq", b"@r1\nACGTN\n-\nIIIII\n", "Third FastQ record line does not begin with '+'"),
    ("c.fq", b">r1\nACGTN\n", "ID-line does not begin with '@'"),
    ("a.fa", b"ACGT\n>s\nAC\n", "Record does not begin with '>' or ';'"),
    ("b.fa", b">s1\n>s2\nAC\n", "No sequence or no valid sequence characters"),
    ("c.fa.bz2", b"BZh91AY&SY", "bzip2-compressed query files are not supported"),
])
def test_malformed_input_fails_loudly(tmp_path, name, content, msg):
    p = tmp_path / name
    p.write_bytes(content)
    rc, out, err = dumpq(p)
    assert rc == 255 and msg in err and out == ""
    if not name.endswith(".bz2"):
        with pytest.raises(ValueError):
            lambda_b200.read_queries(str(p))


def test_missing_file(tmp_path):
    rc, _, err = dumpq(tmp_path / "nope.fasta")
    assert rc == 255 and "Could not open file" in err


def test_reader_never_crashes_on_garbage(tmp_path):
    """random bytes, random cuts of valid files, corrupted gzip streams: a clean error (255) or a parse, never a crash"""
    import numpy as np
    rng = np.random.default_rng(5)
    good_fa = b"".join(b">s%d x\nACGTNNACGT\nACG\n" % i for i in range(50))
    good_fq = b"".join(b"@r%d\nACGTN\n+\nIIIII\n" % i for i in range(50))
    blobs = []
    for _ in range(15):
        blobs.append(("x.fa", rng.integers(0, 256, int(rng.integers(0, 400)), dtype=np.uint8).tobytes()))
        blobs.append(("x.fq", rng.integers(0, 256, int(rng.integers(0, 400)), dtype=np.uint8).tobytes()))
        blobs.append(("x.fa", good_fa[: int(rng.integers(0, len(good_fa)))]))
        blobs.append(("x.fq", good_fq[: int(rng.integers(0, len(good_fq)))]))
        z = bytearray(gzip.compress(good_fa))
        z[int(rng.integers(10, len(z)))] ^= 0xFF
        blobs.append(("x.fa.gz", bytes(z)))
        blobs.append(("x.fq.gz", gzip.compress(good_fq)[: int(rng.integers(1, 60))]))
    for name, data in blobs:
        p = tmp_path / name
        p.write_bytes(data)
        rc, out, err = dumpq(p, "-a", "dna5")
        assert rc in (0, 255), (name, rc, err[-300:])
        if rc == 255:
            assert err.startswith("ERROR: ")
    # a gzip stream that stops in the middle is an error, not a shorter file
    whole = gzip.compress(good_fq)
    p = tmp_path / "cut.fq.gz"
    p.write_bytes(whole[: len(whole) - 12])
    rc, out, err = dumpq(p, "-a", "dna5")
    assert rc == 255 and "cut.fq.gz" in err
