"""Synthetic protein / nucleotide data for parity tests and the benchmark.

Follows SURVEY.md Appendix F: log-normal protein lengths (median ~270) over the 20 standard
residues with Robinson-Robinson-like background frequencies; flat or family-structured
databases; queries are mutated windows of database sequences so that every query has a true
hit.  All generators are seeded and never emit N / X / *.

Everything is vectorised numpy so that the 5M-sequence (1.8 G residue) benchmark database is
produced in tens of seconds.
"""
from __future__ import annotations

import numpy as np

AA = np.frombuffer(b"ARNDCQEGHILKMFPSTWYV", dtype=np.uint8)
# Robinson & Robinson (1991) background frequencies, same residue order as AA
AA_FREQ = np.array([0.07805, 0.05129, 0.04487, 0.05364, 0.01925, 0.04264, 0.06295, 0.07377,
                    0.02199, 0.05142, 0.09019, 0.05744, 0.02243, 0.03856, 0.05203, 0.07120,
                    0.05841, 0.01330, 0.03216, 0.06441])
AA_FREQ = AA_FREQ / AA_FREQ.sum()
NT = np.frombuffer(b"ACGT", dtype=np.uint8)


def _protein_lengths(rng, n, lo=50, hi=2000):
    return np.clip(rng.lognormal(5.6, 0.55, n).astype(np.int64), lo, hi)


def _residue_table():
    # 256-entry lookup table whose letter counts approximate AA_FREQ to 1/256: one random byte per
    # residue makes the 1.6 G-residue benchmark database cheap to generate
    counts = np.floor(AA_FREQ * 256).astype(int)
    rest = 256 - counts.sum()
    order = np.argsort(-(AA_FREQ * 256 - counts))
    counts[order[:rest]] += 1
    return np.repeat(AA, counts)


_RES_TABLE = _residue_table()


def _random_residues(rng, n):
    out = np.empty(n, np.uint8)
    step = 1 << 26
    for b in range(0, n, step):
        e = min(n, b + step)
        out[b:e] = _RES_TABLE[rng.integers(0, 256, e - b, dtype=np.uint8)]
    return out


def protein_db(n_seqs: int, seed: int = 1, family: bool = False):
    """Return (concat uint8 ASCII residues, offsets int64[n+1])."""
    rng = np.random.default_rng(seed)
    if not family:
        lens = _protein_lengths(rng, n_seqs)
        offs = np.zeros(n_seqs + 1, np.int64)
        np.cumsum(lens, out=offs[1:])
        return _random_residues(rng, int(offs[-1])), offs
    fam = 20
    n_roots = max(1, n_seqs // fam)
    rlens = _protein_lengths(rng, n_roots)
    roffs = np.zeros(n_roots + 1, np.int64)
    np.cumsum(rlens, out=roffs[1:])
    roots = _random_residues(rng, int(roffs[-1]))
    seqs = []
    for r in range(n_roots):
        root = roots[roffs[r]:roffs[r + 1]]
        for _ in range(fam):
            seqs.append(mutate_protein(rng, root, 0.30, 0.02))
    order = rng.permutation(len(seqs))
    seqs = [seqs[i] for i in order][:n_seqs]
    lens = np.array([len(s) for s in seqs], np.int64)
    offs = np.zeros(len(seqs) + 1, np.int64)
    np.cumsum(lens, out=offs[1:])
    return np.concatenate(seqs), offs


def mutate_protein(rng, seq, sub, indel):
    s = seq.copy()
    m = rng.random(len(s)) < sub
    s[m] = _random_residues(rng, int(m.sum()))
    if indel > 0:
        r = rng.random(len(s))
        keep = r >= indel / 2            # deletions
        ins = (r >= indel / 2) & (r < indel)  # insertions after this residue
        out = []
        last = 0
        idx = np.nonzero(ins | ~keep)[0]
        for i in idx:
            out.append(s[last:i])
            if ins[i]:
                out.append(s[i:i + 1])
                out.append(_random_residues(rng, 1))
            last = i + 1
        out.append(s[last:])
        s = np.concatenate(out) if out else s
    return s


def protein_queries(db, offs, n_q: int, length: int, seed: int = 2, sub=(0.15, 0.20), indel=0.01):
    """Mutated length-`length` windows of database sequences; returns (concat, offsets)."""
    rng = np.random.default_rng(seed)
    lens = np.diff(offs)
    elig = np.nonzero(lens >= length)[0]
    if len(elig) == 0:
        raise ValueError("no database sequence is long enough")
    pick = elig[rng.integers(0, len(elig), n_q)]
    start = offs[pick] + (rng.random(n_q) * (lens[pick] - length + 1)).astype(np.int64)
    # take a slightly larger window so deletions can be compensated, then cut to exactly `length`
    idx = start[:, None] + np.arange(length)[None, :]
    q = db[idx]                                            # (n_q, length)
    rate = rng.uniform(sub[0], sub[1], n_q)[:, None]
    m = rng.random(q.shape, dtype=np.float32) < rate
    q[m] = _random_residues(rng, int(m.sum()))
    if indel > 0:
        # indels: delete a residue (shift left, pad with a random residue at the end) or insert one
        ev = rng.random(q.shape, dtype=np.float32) < indel
        rows, cols = np.nonzero(ev)
        kinds = rng.random(len(rows)) < 0.5
        fill = _random_residues(rng, len(rows))
        for r, c, k, f in zip(rows, cols, kinds, fill):
            if k:   # deletion
                q[r, c:-1] = q[r, c + 1:]
                q[r, -1] = f
            else:   # insertion
                q[r, c + 1:] = q[r, c:-1]
                q[r, c] = f
    qoffs = np.arange(n_q + 1, dtype=np.int64) * length
    return q.reshape(-1), qoffs


def protein_queries_lengths(db, offs, n_q: int, seed: int = 3, lo=50, hi=2000):
    """Queries with the database's own (log-normal) length distribution (config[3])."""
    rng = np.random.default_rng(seed)
    lens = np.diff(offs)
    want = _protein_lengths(rng, n_q, lo, hi)
    out = []
    for L in want:
        elig = np.nonzero(lens >= L)[0]
        p = elig[rng.integers(0, len(elig))]
        s = offs[p] + rng.integers(0, lens[p] - L + 1)
        w = mutate_protein(rng, db[s:s + L], rng.uniform(0.15, 0.20), 0.01)
        out.append(w)
    qoffs = np.zeros(n_q + 1, np.int64)
    np.cumsum([len(o) for o in out], out=qoffs[1:])
    return np.concatenate(out), qoffs


def nucl_db(n_chrom: int, chrom_len: int, seed: int = 4):
    rng = np.random.default_rng(seed)
    db = NT[rng.integers(0, 4, n_chrom * chrom_len, dtype=np.uint8)]
    offs = np.arange(n_chrom + 1, dtype=np.int64) * chrom_len
    return db, offs


_COMP = np.zeros(256, np.uint8)
for a, b in zip(b"ACGT", b"TGCA"):
    _COMP[a] = b


def nucl_reads(db, offs, n_reads: int, length: int, seed: int = 5, sub=0.03, bisulfite=False):
    rng = np.random.default_rng(seed)
    lens = np.diff(offs)
    pick = rng.integers(0, len(lens), n_reads)
    start = offs[pick] + (rng.random(n_reads) * (lens[pick] - length + 1)).astype(np.int64)
    q = db[start[:, None] + np.arange(length)[None, :]]
    if bisulfite:
        conv = (q == ord("C")) & (rng.random(q.shape, dtype=np.float32) < 0.95)
        q[conv] = ord("T")
        sub = 0.01
    m = rng.random(q.shape, dtype=np.float32) < sub
    q[m] = NT[rng.integers(0, 4, int(m.sum()), dtype=np.uint8)]
    odd = np.arange(n_reads) % 2 == 1
    q[odd] = _COMP[q[odd][:, ::-1]]
    return q.reshape(-1), np.arange(n_reads + 1, dtype=np.int64) * length


# one codon per amino acid (canonical genetic code) for back-translation of synthetic proteins
_CODON = {"A": "GCT", "R": "CGT", "N": "AAT", "D": "GAT", "C": "TGT", "Q": "CAA", "E": "GAA", "G": "GGT", "H": "CAT",
          "I": "ATT", "L": "CTT", "K": "AAA", "M": "ATG", "F": "TTT", "P": "CCT", "S": "TCT", "T": "ACT", "W": "TGG",
          "Y": "TAT", "V": "GTT"}
_SYN = {"A": "GCN", "R": "CGN", "G": "GGN", "L": "CTN", "P": "CCN", "S": "TCN", "T": "ACN", "V": "GTN"}  # 4-fold sites


def back_translate(rng, prot):
    """ASCII protein -> ASCII nucleotides; 4-fold degenerate third positions are randomised."""
    tab = np.zeros((256, 3), np.uint8)
    four = np.zeros(256, bool)
    for a, c in _CODON.items():
        tab[ord(a)] = np.frombuffer(c.encode(), np.uint8)
        four[ord(a)] = a in _SYN
    out = tab[prot].copy()
    m = four[prot]
    out[m, 2] = NT[rng.integers(0, 4, int(m.sum()))]
    return out.reshape(-1)


def revcomp(nt):
    return _COMP[nt[::-1]]


def coding_nucl_seqs(rng, prots, offs, flank=(0, 30), rc_every=2):
    """every protein back-translated into its own nucleotide sequence with random flanks (so that all
    three frames occur); every `rc_every`-th one reverse-complemented"""
    seqs = []
    for i in range(len(offs) - 1):
        p = prots[offs[i]:offs[i + 1]]
        a = NT[rng.integers(0, 4, int(rng.integers(flank[0], flank[1] + 1)))]
        b = NT[rng.integers(0, 4, int(rng.integers(flank[0], flank[1] + 1)))]
        s = np.concatenate([a, back_translate(rng, p), b])
        if rc_every and i % rc_every == 1:
            s = revcomp(s)
        seqs.append(s)
    no = np.zeros(len(seqs) + 1, np.int64)
    np.cumsum([len(x) for x in seqs], out=no[1:])
    return np.concatenate(seqs), no


def write_fasta(path: str, seqs, offs, prefix: str, width: int = 0):
    """Write one record per sequence, `>prefixN` ids (no spaces), sequence on a single line."""
    n = len(offs) - 1
    lens = np.diff(offs)
    ids = [f">{prefix}{i}\n".encode() for i in range(n)]
    idlen = np.fromiter((len(x) for x in ids), np.int64, n)
    rec = idlen + lens + 1
    roffs = np.zeros(n + 1, np.int64)
    np.cumsum(rec, out=roffs[1:])
    buf = np.empty(int(roffs[-1]), np.uint8)
    # sequence bytes: one vectorised scatter
    dst = np.repeat(roffs[:-1] + idlen - offs[:-1], lens) + np.arange(int(offs[-1]), dtype=np.int64)
    buf[dst] = seqs
    buf[roffs[1:] - 1] = ord("\n")
    hdr = np.frombuffer(b"".join(ids), np.uint8)
    hdst = np.repeat(roffs[:-1] - np.concatenate(([0], np.cumsum(idlen)[:-1])), idlen) + np.arange(len(hdr), dtype=np.int64)
    buf[hdst] = hdr
    with open(path, "wb") as f:
        f.write(buf.tobytes())
