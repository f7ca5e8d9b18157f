"""Host-side mirror of the reference's search interface on top of the C ABI.

    idx = Index.load("db.lba", device=0)            # loadDbIndexFromDisk   (src/search_algo.hpp:245)
    s   = Searcher(idx, "protein", profile="none")  # LambdaOptions + LocalDataHolder
    hits, stats = s.search(residues, offsets)       # search() + iterateMatches() + writeRecords()
    lines = s.m8(hits, query_ids)                   # BLAST tabular, SQ/blast/blast_tabular_out.h
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

from ._abi import (DOMAIN, HIT_DT, MATCH_DT, STATS_DT, Hits, IndexDesc, Params, QueryBatch, encode, read_fasta)

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "liblambda_b200.so")


class LambdaError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"lambda_b200 error {code}: {msg}")
        self.code = code


_lib = None


def load_library():
    """Load the CUDA library; fails loudly when it has not been built (no fallback exists)."""
    global _lib
    if _lib is not None:
        return _lib
    path = os.environ.get("LAMBDA_B200_LIB", LIB_PATH)  # kernel experiments: an alternative build of the same library
    if not os.path.exists(path):
        raise ImportError(f"{path} is missing: build it with `python -m lambda_b200.build` "
                          "(lambda_b200 has no CPU fallback)")
    lib = C.CDLL(path)
    vp, u64, i32 = C.c_void_p, C.c_uint64, C.c_int
    lib.lgpu_version.restype = i32
    lib.lgpu_lba_open.argtypes = [C.POINTER(vp), C.c_char_p]
    lib.lgpu_lba_desc.restype = C.POINTER(IndexDesc)
    lib.lgpu_lba_desc.argtypes = [vp]
    lib.lgpu_lba_close.argtypes = [vp]
    lib.lgpu_index_create.argtypes = [C.POINTER(vp), C.POINTER(IndexDesc), i32]
    lib.lgpu_index_destroy.argtypes = [vp]
    lib.lgpu_index_device_bytes.restype = u64
    lib.lgpu_index_device_bytes.argtypes = [vp]
    lib.lgpu_index_db_total_length.restype = u64
    lib.lgpu_index_db_total_length.argtypes = [vp]
    lib.lgpu_index_db_num_seqs.restype = u64
    lib.lgpu_index_db_num_seqs.argtypes = [vp]
    lib.lgpu_params_default.argtypes = [C.POINTER(Params), C.c_uint32, C.c_char_p]
    lib.lgpu_ctx_create.argtypes = [C.POINTER(vp), vp, C.POINTER(Params)]
    lib.lgpu_ctx_destroy.argtypes = [vp]
    lib.lgpu_ctx_set_streams.argtypes = [vp, C.c_uint32]
    lib.lgpu_last_error.restype = C.c_char_p
    lib.lgpu_last_error.argtypes = [vp]
    lib.lgpu_search_batch.argtypes = [vp, C.POINTER(QueryBatch), C.POINTER(Hits), vp]
    lib.lgpu_ctx_export_hits.argtypes = [vp, vp, u64, u64, C.POINTER(u64)]
    lib.lgpu_hits_fill_scores.argtypes = [vp, vp, u64]
    lib.lgpu_seed_batch.argtypes = [vp, C.POINTER(QueryBatch), i32, C.POINTER(vp), C.POINTER(u64), vp]
    lib.lgpu_merge_matches.argtypes = [vp, C.POINTER(QueryBatch), vp, u64, C.POINTER(vp), C.POINTER(u64), vp]
    lib.lgpu_extend_scores.argtypes = [vp, C.POINTER(QueryBatch), vp, u64, vp, vp]
    lib.lgpu_extend_trace.argtypes = [vp, C.POINTER(QueryBatch), vp, u64, vp, vp]
    lib.lgpu_fm_rank.argtypes = [vp, vp, vp, u64, vp]
    lib.lgpu_fm_locate.argtypes = [vp, vp, u64, vp, vp]
    lib.lgpu_bit_score.argtypes = [C.POINTER(Params), C.c_int32, C.POINTER(C.c_double)]
    lib.lgpu_ka_params.argtypes = [C.POINTER(Params)] + [C.POINTER(C.c_double)] * 3
    lib.lgpu_score_matrix.argtypes = [C.POINTER(Params), vp]
    lib.lgpu_evalue.argtypes = [C.POINTER(Params), C.c_int32, u64, u64, C.POINTER(C.c_double)]
    lib.lgpu_min_raw_score.argtypes = [C.POINTER(Params), u64, u64, C.POINTER(C.c_int32)]
    lib.lgpu_format_m8.argtypes = [C.POINTER(Params), vp, C.c_char_p, C.c_char_p, C.c_char_p, C.c_size_t]
    lib.lgpu_format_tabular.argtypes = [C.POINTER(Params), vp, C.c_char_p, C.c_char_p, C.POINTER(C.c_uint32), C.c_size_t,
                                        C.c_char_p, C.c_size_t]
    lib.lgpu_tabular_column.argtypes = [C.c_char_p]
    lib.lgpu_tabular_column_label.argtypes = [C.c_uint32]
    lib.lgpu_tabular_column_label.restype = C.c_char_p
    lib.lgpu_tabular_column_name.argtypes = [C.c_uint32]
    lib.lgpu_tabular_column_name.restype = C.c_char_p
    lib.lgpu_tabular_column_supported.argtypes = [C.c_uint32]
    lib.lgpu_tabular_column_implemented.argtypes = [C.c_uint32]
    _lib = lib
    return lib


def _p(a):
    return C.c_void_p(a.ctypes.data)


def _check(rc, ctx=None):
    if rc != 0:
        msg = load_library().lgpu_last_error(ctx)
        raise LambdaError(rc, msg.decode() if msg else "")


class Index:
    """Device-resident index (replaces index_file<> / GlobalDataHolder, src/shared_definitions.hpp:346)."""

    def __init__(self, handle, lba, subject_ids, desc, device):
        self._h, self._lba, self.subject_ids, self.desc, self.device = handle, lba, subject_ids, desc, device

    @classmethod
    def load(cls, path: str, device: int = 0, keep_ids: bool = True) -> "Index":
        lib = load_library()
        lba = C.c_void_p()
        _check(lib.lgpu_lba_open(C.byref(lba), os.fsencode(path)))
        d = lib.lgpu_lba_desc(lba).contents
        h = C.c_void_p()
        try:
            _check(lib.lgpu_index_create(C.byref(h), C.byref(d), device))
            ids = None
            if keep_ids:
                n = d.n_seqs
                delims = np.ctypeslib.as_array(C.cast(d.id_delims, C.POINTER(C.c_uint64)), (n + 1,))
                raw = C.string_at(d.ids, int(delims[-1]))
                ids = [raw[int(delims[i]):int(delims[i + 1])].decode() for i in range(n)]
            scal = {f: getattr(d, f) for f, t in IndexDesc._fields_ if t is not C.c_void_p}
        finally:
            lib.lgpu_lba_close(lba)  # the host mapping is no longer needed once the index is in HBM
        return cls(h, None, ids, scal, device)

    @property
    def device_bytes(self):
        return load_library().lgpu_index_device_bytes(self._h)

    @property
    def db_total_length(self):
        return load_library().lgpu_index_db_total_length(self._h)

    @property
    def n_seqs(self):
        return load_library().lgpu_index_db_num_seqs(self._h)

    def rank(self, idx, symb):
        idx = np.ascontiguousarray(idx, np.uint64)
        symb = np.ascontiguousarray(symb, np.uint8)
        out = np.zeros(len(idx), np.uint64)
        _check(load_library().lgpu_fm_rank(self._h, _p(idx), _p(symb), len(idx), _p(out)))
        return out

    def locate(self, rows):
        rows = np.ascontiguousarray(rows, np.uint64)
        subj = np.zeros(len(rows), np.uint64)
        pos = np.zeros(len(rows), np.uint64)
        _check(load_library().lgpu_fm_locate(self._h, _p(rows), len(rows), _p(subj), _p(pos)))
        return subj, pos

    def close(self):
        if self._h:
            load_library().lgpu_index_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def default_params(domain="protein", profile="none", **overrides) -> Params:
    p = Params()
    dom = DOMAIN[domain] if isinstance(domain, str) else int(domain)
    _check(load_library().lgpu_params_default(C.byref(p), dom, profile.encode()))
    for k, v in overrides.items():
        if k in ("opts0", "opts"):
            o = getattr(p, k)
            o.seed_length, o.max_seed_dist, o.seed_offset = v
        else:
            if not hasattr(p, k):
                raise AttributeError(k)
            setattr(p, k, v)
    return p


class Searcher:
    """One search context = the reference's per-thread LocalDataHolder plus the batch loop body."""

    def __init__(self, index: Index, domain="protein", profile="none", streams=None, **overrides):
        self.index = index
        self.params = default_params(domain, profile, **overrides)
        self._h = C.c_void_p()
        _check(load_library().lgpu_ctx_create(C.byref(self._h), index._h, C.byref(self.params)))
        if streams is not None:
            _check(load_library().lgpu_ctx_set_streams(self._h, int(streams)), self._h)

    def close(self):
        if self._h:
            load_library().lgpu_ctx_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # -- helpers -----------------------------------------------------------------------------
    @staticmethod
    def _batch(residues, offsets):
        if hasattr(residues, "data_ptr"):  # torch tensors already resident on the GPU
            qb = QueryBatch(residues.data_ptr(), offsets.data_ptr(), offsets.numel() - 1, 1)
            return qb, (residues, offsets)
        residues = np.ascontiguousarray(residues, np.uint8)
        offsets = np.ascontiguousarray(offsets, np.uint64)
        return QueryBatch(residues.ctypes.data, offsets.ctypes.data, len(offsets) - 1, 0), (residues, offsets)

    # -- full path -----------------------------------------------------------------------------
    def search(self, residues, offsets, copy=True):
        """residues: original-alphabet ranks (see encode()), offsets: uint64[n+1] -> (hits, stats).
        copy=False returns a view of the context's own host buffer, valid until the next call."""
        lib = load_library()
        qb, keep = self._batch(residues, offsets)
        out = Hits()
        st = np.zeros(1, STATS_DT)
        _check(lib.lgpu_search_batch(self._h, C.byref(qb), C.byref(out), _p(st)), self._h)
        self.last_cigar_ops = None
        if out.cigar_ops and out.n_cigar_ops:
            rawc = (C.c_uint32 * out.n_cigar_ops).from_address(out.cigar_ops)
            self.last_cigar_ops = np.frombuffer(rawc, np.uint32).copy()  # see lgpu_hit.cigar_off / cigar_len
        if not out.n:
            return np.zeros(0, HIT_DT), st[0]
        raw = (C.c_char * (out.n * HIT_DT.itemsize)).from_address(out.hits)
        hits = np.frombuffer(raw, HIT_DT)
        return (hits.copy() if copy else hits), st[0]

    def export_hits(self, dev_ptr: int, cap_records: int, first_query: int = 0) -> int:
        """copy the records of the last search() into a caller-owned DEVICE buffer (the send buffer of the multi-GPU
        gather), query ids rebased by first_query; returns the number of records (nothing is copied if > cap_records)"""
        n = C.c_uint64()
        _check(load_library().lgpu_ctx_export_hits(self._h, C.c_void_p(dev_ptr), cap_records, first_query, C.byref(n)),
               self._h)
        return int(n.value)

    def fill_scores(self, hits: np.ndarray) -> np.ndarray:
        """bit score and e-value of records that came off a device buffer (host arithmetic, like the reference)"""
        hits = np.ascontiguousarray(hits, HIT_DT)
        _check(load_library().lgpu_hits_fill_scores(self._h, _p(hits), len(hits)), self._h)
        return hits

    @property
    def query_is_protein(self):
        """protein queries (aa27 ranks) unless the domain or `query_alph` says nucleotides (dna5 ranks)"""
        return self.params.domain == 0 and self.params.query_alph != 3

    def search_fasta(self, path):
        ids, data, offs = read_fasta(path)
        hits, st = self.search(encode(data, 0 if self.query_is_protein else 1), offs)
        return ids, hits, st

    def m8(self, hits, query_ids):
        lib = load_library()
        buf = C.create_string_buffer(8192)
        sids = self.index.subject_ids
        lines = []
        for i in range(len(hits)):
            h = hits[i:i + 1]
            n = lib.lgpu_format_m8(C.byref(self.params), _p(h), query_ids[int(h["q_id"][0])].encode(),
                                   sids[int(h["s_id"][0])].encode(), buf, 8192)
            if n <= 0:
                raise LambdaError(n, "lgpu_format_m8 failed")
            lines.append(buf.raw[:n].decode())
        return lines

    def tabular(self, hits, query_ids, columns="std"):
        """BLAST tabular lines with the column list of `lambda3 --output-columns` (space-separated NCBI specifiers)"""
        lib = load_library()
        cols = []
        for name in columns.split():
            c = lib.lgpu_tabular_column(name.encode())
            if c < 0:
                raise LambdaError(-1, f'Unknown column specifier "{name}".')
            cols.append(c)
        arr = (C.c_uint32 * len(cols))(*cols)
        buf = C.create_string_buffer(16384)
        sids = self.index.subject_ids
        lines = []
        for i in range(len(hits)):
            h = hits[i:i + 1]
            n = lib.lgpu_format_tabular(C.byref(self.params), _p(h), query_ids[int(h["q_id"][0])].encode(),
                                        sids[int(h["s_id"][0])].encode(), arr, len(cols), buf, 16384)
            if n < 0:
                raise LambdaError(n, "unsupported column in " + columns)
            lines.append(buf.raw[:n].decode())
        return lines

    # -- stages --------------------------------------------------------------------------------
    def seed(self, residues, offsets, phase):
        lib = load_library()
        qb, keep = self._batch(residues, offsets)
        out, n = C.c_void_p(), C.c_uint64()
        st = np.zeros(1, STATS_DT)
        _check(lib.lgpu_seed_batch(self._h, C.byref(qb), phase, C.byref(out), C.byref(n), _p(st)), self._h)
        m = (np.frombuffer(C.string_at(out, n.value * MATCH_DT.itemsize), MATCH_DT).copy() if n.value
             else np.zeros(0, MATCH_DT))
        return m, st[0]

    def merge(self, residues, offsets, matches):
        lib = load_library()
        qb, keep = self._batch(residues, offsets)
        matches = np.ascontiguousarray(matches, MATCH_DT)
        out, n = C.c_void_p(), C.c_uint64()
        st = np.zeros(1, STATS_DT)
        _check(lib.lgpu_merge_matches(self._h, C.byref(qb), _p(matches), len(matches), C.byref(out), C.byref(n),
                                      _p(st)), self._h)
        m = (np.frombuffer(C.string_at(out, n.value * MATCH_DT.itemsize), MATCH_DT).copy() if n.value
             else np.zeros(0, MATCH_DT))
        return m, st[0]

    def extend_scores(self, residues, offsets, windows):
        lib = load_library()
        qb, keep = self._batch(residues, offsets)
        windows = np.ascontiguousarray(windows, MATCH_DT)
        scores = np.zeros(len(windows), np.int32)
        st = np.zeros(1, STATS_DT)
        _check(lib.lgpu_extend_scores(self._h, C.byref(qb), _p(windows), len(windows), _p(scores), _p(st)), self._h)
        return scores, st[0]

    def extend_trace(self, residues, offsets, windows):
        lib = load_library()
        qb, keep = self._batch(residues, offsets)
        windows = np.ascontiguousarray(windows, MATCH_DT)
        hits = np.zeros(len(windows), HIT_DT)
        st = np.zeros(1, STATS_DT)
        _check(lib.lgpu_extend_trace(self._h, C.byref(qb), _p(windows), len(windows), _p(hits), _p(st)), self._h)
        return hits, st[0]
