#!/usr/bin/env python
"""Golden SAM / BAM files for the non-default dialect options of the reference (src/search_options.hpp:276-370):
  <case>/opt_soft.sam  --sam-bam-clip soft --sam-bam-seq always --sam-with-refheader 1 + all non-taxonomy tags
  <case>/opt_tags.sam  hard clipping, --sam-bam-seq uniq, all non-taxonomy tags
  <case>/opt_tags.bam  all non-taxonomy tags, --sam-bam-seq never
  <case>/opt_soft.bam  soft clipping, --sam-bam-seq always, all non-taxonomy tags
made by the unmodified reference binary, run inside the case directory (-t 1, --version-to-outputfile 0)."""
import gzip
import os
import shutil
import subprocess
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.path.join(os.path.dirname(os.path.dirname(HERE)), "oracle", "_ref", "lambda3")
CASES = [("prot_flat", "searchp"), ("prot_family", "searchp"), ("prot_diverged", "searchp"), ("nucl", "searchn"),
         ("bisulfite", "searchbs"), ("blastx", "searchp"), ("tblastn", "searchp"), ("tblastx", "searchp")]
ALL_TAGS = "AS OC NM IH ar ae ai ap qf qs sf"
VARIANTS = {
    "opt_soft.sam": ["--sam-bam-clip", "soft", "--sam-bam-seq", "always", "--sam-with-refheader", "1", "--sam-bam-tags", ALL_TAGS],
    "opt_tags.sam": ["--sam-bam-tags", ALL_TAGS],
    "opt_tags.bam": ["--sam-bam-tags", ALL_TAGS, "--sam-bam-seq", "never"],
    "opt_soft.bam": ["--sam-bam-clip", "soft", "--sam-bam-seq", "always", "--sam-bam-tags", ALL_TAGS],
}

if __name__ == "__main__":
    for case, cmd in CASES:
        src = os.path.join(HERE, case)
        with tempfile.TemporaryDirectory() as tmp:
            with gzip.open(os.path.join(src, "db.lba.gz"), "rb") as fi, open(os.path.join(tmp, "db.lba"), "wb") as fo:
                shutil.copyfileobj(fi, fo)
            shutil.copy(os.path.join(src, "q.fasta"), tmp)
            for name, extra in VARIANTS.items():
                subprocess.check_call([REF, cmd, "-q", "q.fasta", "-i", "db.lba", "-o", name, "-t", "1", "-v", "0",
                                       "--version-to-outputfile", "0", *extra], cwd=tmp)
                shutil.copy(os.path.join(tmp, name), os.path.join(src, name))
                print(case, name, os.path.getsize(os.path.join(src, name)), "bytes")
