"""ctypes binding of the CPU oracle (oracle/_build/liblambda_oracle.so) -- test infrastructure only."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
import sys  # noqa: E402

sys.path.insert(0, ROOT)
LIB = os.path.join(ROOT, "oracle", "_build", "liblambda_oracle.so")


from lambda_b200._abi import (HIT_DT, MATCH_DT, STATS_DT, IndexDesc, Params, SearchOpts, encode,  # noqa: E402,F401
                              read_fasta)


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
        # gapped rows as run-length operations (lgpu_params.want_cigar), indexed by lgpu_hit.cigar_off / cigar_len
        ops = C.POINTER(C.c_uint32)()
        self.l.orc_last_cigar.restype = C.c_uint64
        nops = self.l.orc_last_cigar(C.c_void_p(self.h), C.byref(ops))
        self.last_cigar_ops = np.ctypeslib.as_array(ops, (nops,)).copy() if nops else np.zeros(0, np.uint32)
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
