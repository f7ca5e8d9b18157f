#!/usr/bin/env python
"""Golden tabular files with a custom column list (`--output-columns`) from the unmodified reference binary:
<case>/cols.m9 for every case (all columns the reference's tabular writer implements except the taxonomy ones, plus
two it does not implement -> "n/i").  Run inside the case directory with `-i db.lba` (echoed in `# Database:`)."""
import gzip
import os
import shutil
import subprocess
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.path.join(os.path.dirname(os.path.dirname(HERE)), "oracle", "_ref", "lambda3")
CASES = [("prot_flat", "searchp"), ("prot_diverged", "searchp"), ("nucl", "searchn"), ("bisulfite", "searchbs"),
         ("blastx", "searchp"), ("tblastn", "searchp"), ("tblastx", "searchp")]
COLUMNS = ("qseqid qlen sseqid slen std score length pident nident mismatch positive gapopen gaps ppos frames qframe "
           "sframe qacc sacc sallacc qgi btop evalue bitscore qstart qend sstart send")

if __name__ == "__main__":
    for case, cmd in CASES:
        src = os.path.join(HERE, case)
        with tempfile.TemporaryDirectory() as tmp:
            with gzip.open(os.path.join(src, "db.lba.gz"), "rb") as fi, open(os.path.join(tmp, "db.lba"), "wb") as fo:
                shutil.copyfileobj(fi, fo)
            shutil.copy(os.path.join(src, "q.fasta"), tmp)
            subprocess.check_call([REF, cmd, "-q", "q.fasta", "-i", "db.lba", "-o", "cols.m9", "-t", "1", "-v", "0",
                                   "--version-to-outputfile", "0", "--output-columns", COLUMNS], cwd=tmp)
            shutil.copy(os.path.join(tmp, "cols.m9"), os.path.join(src, "cols.m9"))
            print(case, os.path.getsize(os.path.join(src, "cols.m9")), "bytes")
