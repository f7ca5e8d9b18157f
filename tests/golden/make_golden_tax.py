#!/usr/bin/env python
"""Golden fixture with taxonomy (tests/golden/tax/): a small protein family database whose ids carry UniProt-style
accessions, an NCBI-style accession->taxid map and a tiny nodes.dmp / names.dmp tree; the unmodified reference
builds the index (mkindexp --acc-tax-map --tax-dump-dir) and writes
    tax.m9    --output-columns 'std staxids lcaid lcataxid'
    tax.sam   --sam-bam-tags 'AS NM ae ai qf st ls lt'
    tax.bam   the same tags
Subjects without a mapping, subjects with two tax ids and names with blanks are included on purpose."""
import gzip
import os
import shutil
import subprocess
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
from lambda_b200 import synth  # noqa: E402

REF = os.path.join(ROOT, "oracle", "_ref", "lambda3")
# tax id -> (parent, scientific name)
TREE = {1: (1, "root"), 10: (1, "cellular organisms"), 20: (10, "Bacteria"), 30: (10, "Eukaryota"),
        21: (20, "Escherichia coli"), 22: (20, "Bacillus subtilis"), 23: (21, "Escherichia coli K-12"),
        31: (30, "Homo sapiens"), 32: (30, "Mus musculus"), 33: (31, "Homo sapiens neanderthalensis"),
        40: (1, "Viruses"), 41: (40, "Tobacco mosaic virus")}
LEAVES = [21, 22, 23, 31, 32, 33, 41, 20]
COLUMNS = "std staxids lcaid lcataxid"
TAGS = "AS NM ae ai qf st ls lt"

if __name__ == "__main__":
    out = os.path.join(HERE, "tax")
    shutil.rmtree(out, ignore_errors=True)
    os.makedirs(out)
    rng = np.random.default_rng(151)
    # close families (8 % divergence from the family root) so that a query hits several subjects of different taxa
    roots, roffs = synth.protein_db(20, seed=152)
    seqs = [synth.mutate_protein(rng, roots[roffs[r]:roffs[r + 1]], 0.08, 0.01) for r in range(20) for _ in range(6)]
    seqs = [seqs[i] for i in rng.permutation(len(seqs))]
    offs = np.zeros(len(seqs) + 1, np.int64)
    np.cumsum([len(x) for x in seqs], out=offs[1:])
    db = np.concatenate(seqs)
    q, qo = synth.protein_queries(db, offs, 24, 140, seed=153, sub=(0.10, 0.20), indel=0.01)
    with tempfile.TemporaryDirectory() as tmp:
        acc = [f"P{i:05d}" for i in range(1, len(offs))]
        with open(f"{tmp}/db.fasta", "w") as f:
            for i, a in enumerate(acc):
                f.write(f">{a} synthetic protein {i}\n{db[offs[i]:offs[i + 1]].tobytes().decode()}\n")
        synth.write_fasta(f"{out}/q.fasta", q, qo, "Q")
        with open(f"{tmp}/map.accession2taxid", "w") as f:
            f.write("accession\taccession.version\ttaxid\tgi\n")
            for i, a in enumerate(acc):
                if i % 7 == 3:
                    continue  # no taxonomy for this subject
                f.write(f"{a}\t{a}.1\t{LEAVES[int(rng.integers(0, len(LEAVES)))]}\t0\n")
                if i % 11 == 5:  # a second tax id for the same subject
                    f.write(f"{a}\t{a}.1\t{LEAVES[int(rng.integers(0, len(LEAVES)))]}\t0\n")
        os.makedirs(f"{tmp}/taxdump")
        with open(f"{tmp}/taxdump/nodes.dmp", "w") as f:
            for t, (p, _) in TREE.items():
                f.write(f"{t}\t|\t{p}\t|\tno rank\t|\t\t|\n")
        with open(f"{tmp}/taxdump/names.dmp", "w") as f:
            for t, (_, name) in TREE.items():
                f.write(f"{t}\t|\t{name}\t|\t\t|\tscientific name\t|\n")
                f.write(f"{t}\t|\tsynonym of {name}\t|\t\t|\tsynonym\t|\n")
        subprocess.check_call([REF, "mkindexp", "-d", f"{tmp}/db.fasta", "-i", f"{tmp}/db.lba", "-v", "0", "--acc-tax-map",
                               f"{tmp}/map.accession2taxid", "--tax-dump-dir", f"{tmp}/taxdump"])
        shutil.copy(f"{out}/q.fasta", f"{tmp}/q.fasta")
        for name, extra in (("tax.m9", ["--output-columns", COLUMNS]), ("tax.sam", ["--sam-bam-tags", TAGS]),
                            ("tax.bam", ["--sam-bam-tags", TAGS]), ("none.m8", [])):
            subprocess.check_call([REF, "searchp", "-q", "q.fasta", "-i", "db.lba", "-o", name, "-t", "1", "-v", "0",
                                   "--version-to-outputfile", "0", *extra], cwd=tmp)
            shutil.copy(f"{tmp}/{name}", f"{out}/{name}")
            print(name, os.path.getsize(f"{out}/{name}"), "bytes")
        with open(f"{tmp}/db.lba", "rb") as fi, gzip.GzipFile(f"{out}/db.lba.gz", "wb", 9, mtime=0) as fo:
            shutil.copyfileobj(fi, fo)
