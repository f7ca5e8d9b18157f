"""ctypes binding of the CPU oracle (oracle/_build/liblambda_oracle.so) -- test infrastructure only."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "oracle", "_build", "liblambda_oracle.so")


class SearchOpts(C.Structure):
    _fields_ = [("seed_length", C.c_uint32), ("max_seed_dist", C.c_uint32), ("seed_offset", C.c_uint32)]


class Params(C.Structure):
    _fields_ = [("domain", C.c_uint32), ("opts0", SearchOpts), ("opts", SearchOpts),
                ("seed_half_exact", C.c_uint32), ("adaptive_seeding", C.c_uint32), ("iterative_search", C.c_uint32),
                ("max_matches", C.c_uint32), ("pre_scoring", C.c_int32), ("pre_scoring_thresh", C.c_double),
                ("scoring_method", C.c_int32), ("gap_open", C.c_int32), ("gap_extend", C.c_int32),
                ("match", C.c_int32), ("mismatch", C.c_int32), ("min_bit_score", C.c_int32),
                ("max_evalue", C.c_double), ("id_cutoff", C.c_int32), ("finalize", C.c_uint32)]


MATCH_DT = np.dtype([("qry_id", "<u4"), ("subj_id", "<u4"), ("qry_start", "<u4"), ("qry_end", "<u4"),
                     ("subj_start", "<u4"), ("subj_end", "<u4")])
HIT_DT = np.dtype([("q_id", "<u4"), ("s_id", "<u4"), ("q_start", "<u4"), ("q_end", "<u4"), ("s_start", "<u4"),
                   ("s_end", "<u4"), ("q_len", "<u4"), ("s_len", "<u4"), ("score", "<i4"), ("n_match", "<u4"),
                   ("n_mismatch", "<u4"), ("n_gap_open", "<u4"), ("n_gap_ext", "<u4"), ("n_positive", "<u4"),
                   ("aln_len", "<u4"), ("q_frame", "i1"), ("s_frame", "i1"), ("phase", "u1"), ("reserved", "u1"),
                   ("bit_score", "<f8"), ("evalue", "<f8")])
assert HIT_DT.itemsize == 80
STATS_DT = np.dtype([(n, "<u8") for n in
                     ("hits_after_seeding", "hits_failed_pre_extend", "hits_failed_evalue", "hits_failed_bitscore",
                      "hits_failed_identity", "hits_duplicate", "hits_duplicate2", "hits_abundant", "hits_final",
                      "pairs", "qrys_with_hit", "n_extensions_score", "n_extensions_trace", "cells_score",
                      "cells_trace", "kernel_launches")] +
                    [(n, "<f4") for n in ("ms_seed", "ms_sort_merge", "ms_extend_score", "ms_extend_trace", "ms_h2d",
                                          "ms_d2h", "ms_total", "reserved")])
assert STATS_DT.itemsize == 160


class IndexDesc(C.Structure):
    _fields_ = [("index_type", C.c_uint32), ("orig_alph", C.c_uint32), ("trans_alph", C.c_uint32),
                ("red_alph", C.c_uint32), ("sigma", C.c_uint32), ("sigma_bits", C.c_uint32),
                ("block_bytes", C.c_uint32), ("planes_offset", C.c_uint32), ("occ_blocks", C.c_void_p),
                ("n_blocks", C.c_uint64), ("super_blocks", C.c_void_p), ("n_super", C.c_uint64), ("C", C.c_void_p),
                ("ssa", C.c_void_p), ("n_ssa", C.c_uint64), ("csa_bv", C.c_void_p), ("n_csa_sb", C.c_uint64),
                ("sampling_rate", C.c_uint64), ("bits_for_position", C.c_uint64), ("seqs", C.c_void_p),
                ("n_residues", C.c_uint64), ("seq_delims", C.c_void_p), ("n_seqs", C.c_uint64), ("ids", C.c_void_p),
                ("id_delims", C.c_void_p)]


def build():
    subprocess.check_call(["make", "-s", "-f", os.path.join(ROOT, "oracle", "Makefile.oracle")])


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = C.CDLL(LIB)
        _lib.orc_open.restype = C.c_void_p
        _lib.orc_open.argtypes = [C.c_char_p]
        _lib.orc_close.argtypes = [C.c_void_p]
        _lib.orc_desc.restype = C.POINTER(IndexDesc)
        _lib.orc_desc.argtypes = [C.c_void_p]
    return _lib


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


AA27 = "ABCDEFGHIJKLMNOPQRSTUVWXYZ*"
DNA5 = "ACGNT"


def encode(seq_bytes: np.ndarray, domain: int) -> np.ndarray:
    """ASCII -> original-alphabet ranks (aa27 or dna5), unknown -> X / N like BioC++."""
    alph = AA27 if domain == 0 else DNA5
    tab = np.full(256, 23 if domain == 0 else 3, np.uint8)
    for r, ch in enumerate(alph):
        tab[ord(ch)] = r
        tab[ord(ch.lower())] = r
    if domain != 0:
        tab[ord("U")] = tab[ord("u")] = 4
    return tab[seq_bytes]


def read_fasta(path):
    ids, seqs = [], []
    cur = []
    with open(path, "rb") as f:
        for line in f:
            line = line.rstrip(b"\r\n")
            if line.startswith(b">"):
                if ids:
                    seqs.append(b"".join(cur))
                ids.append(line[1:].decode())
                cur = []
            elif line:
                cur.append(line)
    if ids:
        seqs.append(b"".join(cur))
    offs = np.zeros(len(seqs) + 1, np.uint64)
    np.cumsum([len(s) for s in seqs], out=offs[1:])
    data = np.frombuffer(b"".join(seqs), np.uint8)
    return ids, data, offs


class Oracle:
    def __init__(self, index_path: str):
        self.l = lib()
        self.h = self.l.orc_open(index_path.encode())
        if not self.h:
            raise RuntimeError("orc_open failed for " + index_path)
        self.desc = self.l.orc_desc(self.h).contents
        d = self.desc
        n = d.n_seqs
        delims = np.ctypeslib.as_array(C.cast(d.id_delims, C.POINTER(C.c_uint64)), (n + 1,))
        raw = C.string_at(d.ids, int(delims[-1]))
        self.subject_ids = [raw[int(delims[i]):int(delims[i + 1])].decode() for i in range(n)]

    def close(self):
        if self.h:
            self.l.orc_close(self.h)
            self.h = None

    def params(self, domain=0, profile="none") -> Params:
        p = Params()
        rc = self.l.orc_params_default(C.byref(p), domain, profile.encode())
        assert rc == 0
        return p

    def rank(self, idx, symb):
        idx = np.ascontiguousarray(idx, np.uint64)
        symb = np.ascontiguousarray(symb, np.uint8)
        out = np.zeros(len(idx), np.uint64)
        self.l.orc_rank(C.c_void_p(self.h), _p(idx), _p(symb), C.c_uint64(len(idx)), _p(out))
        return out

    def locate(self, rows):
        rows = np.ascontiguousarray(rows, np.uint64)
        subj = np.zeros(len(rows), np.uint64)
        pos = np.zeros(len(rows), np.uint64)
        self.l.orc_locate(C.c_void_p(self.h), _p(rows), C.c_uint64(len(rows)), _p(subj), _p(pos))
        return subj, pos

    def seed(self, p, res, offs, phase):
        st = np.zeros(1, STATS_DT)
        out = C.c_void_p()
        n = C.c_uint64()
        rc = self.l.orc_seed(C.c_void_p(self.h), C.byref(p), _p(res), _p(offs), C.c_uint64(len(offs) - 1), phase,
                             C.byref(out), C.byref(n), _p(st))
        assert rc == 0
        m = np.frombuffer(C.string_at(out, n.value * MATCH_DT.itemsize), MATCH_DT).copy() if n.value else np.zeros(0, MATCH_DT)
        return m, st[0]

    def merge(self, p, res, offs, matches):
        st = np.zeros(1, STATS_DT)
        matches = np.ascontiguousarray(matches, MATCH_DT)
        out = C.c_void_p()
        n = C.c_uint64()
        rc = self.l.orc_merge(C.c_void_p(self.h), C.byref(p), _p(res), _p(offs), C.c_uint64(len(offs) - 1),
                              _p(matches), C.c_uint64(len(matches)), C.byref(out), C.byref(n), _p(st))
        assert rc == 0
        m = np.frombuffer(C.string_at(out, n.value * MATCH_DT.itemsize), MATCH_DT).copy() if n.value else np.zeros(0, MATCH_DT)
        return m, st[0]

    def extend(self, p, res, offs, windows, with_trace):
        windows = np.ascontiguousarray(windows, MATCH_DT)
        scores = np.zeros(len(windows), np.int32)
        hits = np.zeros(len(windows), HIT_DT)
        rc = self.l.orc_extend(C.c_void_p(self.h), C.byref(p), _p(res), _p(offs), C.c_uint64(len(offs) - 1),
                               _p(windows), C.c_uint64(len(windows)), int(with_trace), _p(scores), _p(hits))
        assert rc == 0
        return scores, hits

    def search(self, p, res, offs, threads=1):
        st = np.zeros(1, STATS_DT)
        out = C.c_void_p()
        n = C.c_uint64()
        rc = self.l.orc_search(C.c_void_p(self.h), C.byref(p), _p(res), _p(offs), C.c_uint64(len(offs) - 1),
                               threads, C.byref(out), C.byref(n), _p(st))
        assert rc == 0
        hits = np.frombuffer(C.string_at(out, n.value * HIT_DT.itemsize), HIT_DT).copy() if n.value else np.zeros(0, HIT_DT)
        return hits, st[0]

    def m8(self, p, hits, query_ids):
        buf = C.create_string_buffer(4096)
        lines = []
        for h in hits:
            hh = np.array([h], HIT_DT)
            n = self.l.orc_format_m8(p.domain, _p(hh), query_ids[int(h["q_id"])].encode(),
                                     self.subject_ids[int(h["s_id"])].encode(), buf, C.c_size_t(4096))
            lines.append(buf.raw[:n].decode())
        return lines
