#!/usr/bin/env python
"""Generate the golden fixtures under tests/golden/ with the UNMODIFIED reference binary.

Run in the build container (needs oracle/_ref/lambda3, built by `make -C oracle` from
/root/reference).  For every case it writes
    <case>/db.fasta.gz  q.fasta  db.lba.gz          inputs (index built by the reference's mkindex*)
    <case>/<profile>.m8                              the reference's tabular output (-t 1)
    <case>/<profile>.funnel.json                     the reference's hit funnel (-v 2 statistics)
The fixtures are small (a few MB in total) and committed; this script is the provenance.
"""
import gzip
import json
import os
import re
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

FUNNEL_KEYS = [("after Seeding", "hits_after_seeding"), ("failed pre-extend test", "hits_failed_pre_extend"),
               ("failed e-value test", "hits_failed_evalue"), ("failed bitScore test", "hits_failed_bitscore"),
               ("failed %-identity test", "hits_failed_identity"), ("- duplicates", "hits_duplicate"),
               ("late duplicates", "hits_duplicate2"), ("abundant", "hits_abundant")]


def parse_funnel(text):
    text = re.sub(r"\x1b\[[0-9;]*m", "", text)
    out = {}
    for label, key in FUNNEL_KEYS:
        m = re.search(re.escape(label) + r"\s+(\d+)", text)
        out[key] = int(m.group(1)) if m else None
    m = re.search(r"Number of total hits:\s+(\d+)", text)
    out["hits_final"] = int(m.group(1))
    m = re.search(r"Number of Query-Subject pairs:\s+(\d+)", text)
    out["pairs"] = int(m.group(1))
    m = re.search(r"Number of Queries with at least one valid hit:\s+(\d+)", text)
    out["qrys_with_hit"] = int(m.group(1))
    return out


def run_case(name, domain, db, offs, q, qoffs, profiles, extra=()):
    mk = {"p": "mkindexp", "n": "mkindexn", "bs": "mkindexbs"}[domain]
    se = {"p": "searchp", "n": "searchn", "bs": "searchbs"}[domain]
    out = os.path.join(HERE, name)
    shutil.rmtree(out, ignore_errors=True)
    os.makedirs(out)
    with tempfile.TemporaryDirectory() as tmp:
        synth.write_fasta(f"{tmp}/db.fasta", db, offs, "S")
        synth.write_fasta(f"{out}/q.fasta", q, qoffs, "Q")
        subprocess.check_call([REF, mk, "-d", f"{tmp}/db.fasta", "-i", f"{tmp}/db.lba", "-v", "0"])
        for prof in profiles:
            o = f"{tmp}/{prof}.m8"
            cmd = [REF, se, "-q", f"{out}/q.fasta", "-i", f"{tmp}/db.lba", "-o", o, "-t", "1",
                   "--version-to-outputfile", "0", "-v", "2", *extra]
            if prof != "none":
                cmd += ["-p", prof]
            txt = subprocess.run(cmd, check=True, capture_output=True, text=True).stdout
            shutil.copy(o, f"{out}/{prof}.m8")
            with open(f"{out}/{prof}.funnel.json", "w") as f:
                json.dump(parse_funnel(txt), f, indent=1)
            print(name, prof, sum(1 for _ in open(o)), "hits")
        for fn in ("db.fasta", "db.lba"):
            with open(f"{tmp}/{fn}", "rb") as fi, gzip.GzipFile(f"{out}/{fn}.gz", "wb", 9, mtime=0) as fo:
                shutil.copyfileobj(fi, fo)


def with_random_queries(rng_seed, q, qoffs, n_random, length):
    """append queries unrelated to the database so that phase 2 / no-hit paths are exercised"""
    rng = np.random.default_rng(rng_seed)
    extra = synth._random_residues(rng, n_random * length)
    q2 = np.concatenate([q, extra])
    qo2 = np.concatenate([qoffs, qoffs[-1] + np.arange(1, n_random + 1, dtype=np.int64) * length])
    return q2, qo2


def main_bs():
    # bisulfite: C->T converted reads (95 %), 1 % errors, odd reads reverse-complemented
    db, offs = synth.nucl_db(4, 50_000, seed=131)
    q, qo = synth.nucl_reads(db, offs, 160, 100, seed=132, bisulfite=True)
    # the last 48 reads get 8 % extra substitutions (phase 2, e-value failures), plus 8 unrelated reads
    rng = np.random.default_rng(133)
    q = q.copy().reshape(160, 100)
    m = rng.random((48, 100)) < 0.08
    tail = q[112:]
    tail[m] = synth.NT[rng.integers(0, 4, int(m.sum()))]
    q = np.concatenate([q.reshape(-1), synth.NT[rng.integers(0, 4, 8 * 100)]])
    qo = np.arange(169, dtype=np.int64) * 100
    run_case("bisulfite", "bs", db, offs, q, qo, ["none", "fast", "sensitive"])


def main_translated():
    rng = np.random.default_rng(140)
    # BLASTX: protein database, nucleotide queries coding for mutated protein windows (+ random reads)
    db, offs = synth.protein_db(300, seed=141)
    qp, qpo = synth.protein_queries(db, offs, 40, 90, seed=142, sub=(0.10, 0.25))
    qn, qno = synth.coding_nucl_seqs(rng, qp, qpo, flank=(0, 12))
    extra = synth.NT[rng.integers(0, 4, 6 * 250)]
    qn = np.concatenate([qn, extra, synth.NT[rng.integers(0, 4, 20)]])
    qno = np.concatenate([qno, qno[-1] + np.arange(1, 7, dtype=np.int64) * 250, [qno[-1] + 6 * 250 + 20]])
    run_case("blastx", "p", db, offs, qn, qno, ["none", "sensitive"])
    # TBLASTN: nucleotide database (coding sequences on both strands), protein queries
    pdb, poffs = synth.protein_db(120, seed=143)
    ndb, noffs = synth.coding_nucl_seqs(rng, pdb, poffs, flank=(0, 40))
    qp, qpo = synth.protein_queries(pdb, poffs, 40, 80, seed=144, sub=(0.10, 0.25))
    qp, qpo = with_random_queries(145, qp, qpo, 5, 80)
    run_case("tblastn", "p", ndb, noffs, qp, qpo, ["none", "sensitive"])
    # TBLASTX: the same nucleotide database, nucleotide queries
    qp, qpo = synth.protein_queries(pdb, poffs, 30, 70, seed=146, sub=(0.10, 0.20))
    qn, qno = synth.coding_nucl_seqs(rng, qp, qpo, flank=(0, 9))
    run_case("tblastx", "p", ndb, noffs, qn, qno, ["none"])


def main():
    # protein, flat database; queries: mutated windows + a few random ones + one shorter than a seed
    db, offs = synth.protein_db(500, seed=101)
    q, qo = synth.protein_queries(db, offs, 48, 120, seed=102)
    q, qo = with_random_queries(103, q, qo, 8, 120)
    q = np.concatenate([q, np.frombuffer(b"MKVLA", np.uint8)])
    qo = np.concatenate([qo, [qo[-1] + 5]])
    run_case("prot_flat", "p", db, offs, q, qo, ["none", "fast", "sensitive", "pairs-default"])

    # protein, family database (many homologs per query -> merging, top-N truncation with -n 5)
    db, offs = synth.protein_db(400, seed=111, family=True)
    q, qo = synth.protein_queries(db, offs, 30, 150, seed=112, sub=(0.25, 0.35), indel=0.02)
    run_case("prot_family", "p", db, offs, q, qo, ["none", "sensitive"])
    # strongly diverged queries: most fail phase 1 and go through phase 2 (half-exact seeds)
    q, qo = synth.protein_queries(db, offs, 40, 100, seed=113, sub=(0.30, 0.40), indel=0.02)
    run_case("prot_diverged", "p", db, offs, q, qo, ["none"])

    # nucleotide
    db, offs = synth.nucl_db(4, 50_000, seed=121)
    q, qo = synth.nucl_reads(db, offs, 120, 150, seed=122, sub=0.04)
    run_case("nucl", "n", db, offs, q, qo, ["none", "fast", "sensitive"])


if __name__ == "__main__":
    if "--bs-only" in sys.argv:
        main_bs()
    elif "--translated-only" in sys.argv:
        main_translated()
    else:
        main()
        main_bs()
        main_translated()
