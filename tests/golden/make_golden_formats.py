#!/usr/bin/env python
"""Golden `.m9` (BLAST tabular with comment lines) and `.sam` / `.bam` / `.m0` files from the unmodified reference binary, for
the committed index / query fixtures:  <case>/none.m9, <case>/none.sam (--version-to-outputfile 0) and
<case>/none.v1.m9 (default version string; the SAM @PG line echoes the command line, so only v0 is kept).  The reference is run inside the case directory with
`-i db.lba`, because the index path is echoed in the `# Database:` line."""
import gzip
import os
import shutil
import subprocess
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.path.join(os.path.dirname(os.path.dirname(HERE)), "oracle", "_ref", "lambda3")
CASES = [("prot_flat", "searchp"), ("prot_family", "searchp"), ("prot_diverged", "searchp"), ("nucl", "searchn"),
         ("bisulfite", "searchbs"), ("blastx", "searchp"), ("tblastn", "searchp"), ("tblastx", "searchp")]

for case, cmd in CASES:
    src = os.path.join(HERE, case)
    with tempfile.TemporaryDirectory() as tmp:
        with gzip.open(os.path.join(src, "db.lba.gz"), "rb") as fi, open(os.path.join(tmp, "db.lba"), "wb") as fo:
            shutil.copyfileobj(fi, fo)
        shutil.copy(os.path.join(src, "q.fasta"), tmp)
        for name, extra in (("none.m9", ["--version-to-outputfile", "0"]), ("none.v1.m9", []),
                            ("none.sam", ["--version-to-outputfile", "0"]), ("none.m0", ["--version-to-outputfile", "0"]),
                            ("none.bam", ["--version-to-outputfile", "0"])):
            subprocess.check_call([REF, cmd, "-q", "q.fasta", "-i", "db.lba", "-o", name, "-t", "1", "-v", "0", *extra], cwd=tmp)
            if name.endswith(".m0"):  # pairwise reports are long: stored gzipped (the test fixture unpacks *.gz)
                with open(os.path.join(tmp, name), "rb") as fi, gzip.GzipFile(os.path.join(src, name + ".gz"), "wb", 9, mtime=0) as fo:
                    shutil.copyfileobj(fi, fo)
            else:
                shutil.copy(os.path.join(tmp, name), os.path.join(src, name))
            print(case, name, os.path.getsize(os.path.join(tmp, name)), "bytes")
