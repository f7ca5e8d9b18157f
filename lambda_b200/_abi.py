"""ctypes / numpy mirrors of the POD types in include/lambda_b200.h."""
import ctypes as C

import numpy as np


class SearchOpts(C.Structure):
    _fields_ = [("seed_length", C.c_uint32), ("max_seed_dist", C.c_uint32), ("seed_offset", C.c_uint32)]


class Params(C.Structure):
    _fields_ = [("domain", C.c_uint32), ("opts0", SearchOpts), ("opts", SearchOpts),
                ("seed_half_exact", C.c_uint32), ("adaptive_seeding", C.c_uint32), ("iterative_search", C.c_uint32),
                ("max_matches", C.c_uint32), ("pre_scoring", C.c_int32), ("pre_scoring_thresh", C.c_double),
                ("scoring_method", C.c_int32), ("gap_open", C.c_int32), ("gap_extend", C.c_int32),
                ("match", C.c_int32), ("mismatch", C.c_int32), ("min_bit_score", C.c_int32),
                ("max_evalue", C.c_double), ("id_cutoff", C.c_int32), ("finalize", C.c_uint32),
                ("query_alph", C.c_uint32), ("want_cigar", C.c_uint32), ("window_band", C.c_uint32)]


class IndexDesc(C.Structure):
    _fields_ = [("index_type", C.c_uint32), ("orig_alph", C.c_uint32), ("trans_alph", C.c_uint32),
                ("red_alph", C.c_uint32), ("sigma", C.c_uint32), ("sigma_bits", C.c_uint32),
                ("block_bytes", C.c_uint32), ("planes_offset", C.c_uint32), ("occ_blocks", C.c_void_p),
                ("n_blocks", C.c_uint64), ("super_blocks", C.c_void_p), ("n_super", C.c_uint64), ("C", C.c_void_p),
                ("ssa", C.c_void_p), ("n_ssa", C.c_uint64), ("csa_bv", C.c_void_p), ("n_csa_sb", C.c_uint64),
                ("sampling_rate", C.c_uint64), ("bits_for_position", C.c_uint64), ("seqs", C.c_void_p),
                ("n_residues", C.c_uint64), ("seq_delims", C.c_void_p), ("n_seqs", C.c_uint64), ("ids", C.c_void_p),
                ("id_delims", C.c_void_p)]


class QueryBatch(C.Structure):
    _fields_ = [("residues", C.c_void_p), ("offsets", C.c_void_p), ("n_queries", C.c_uint64),
                ("on_device", C.c_uint32)]


class Hits(C.Structure):
    _fields_ = [("hits", C.c_void_p), ("n", C.c_uint64), ("cigar_ops", C.c_void_p), ("n_cigar_ops", C.c_uint64)]


MATCH_DT = np.dtype([("qry_id", "<u4"), ("subj_id", "<u4"), ("qry_start", "<u4"), ("qry_end", "<u4"),
                     ("subj_start", "<u4"), ("subj_end", "<u4")])
HIT_DT = np.dtype([("q_id", "<u4"), ("s_id", "<u4"), ("q_start", "<u4"), ("q_end", "<u4"), ("s_start", "<u4"),
                   ("s_end", "<u4"), ("q_len", "<u4"), ("s_len", "<u4"), ("score", "<i4"), ("n_match", "<u4"),
                   ("n_mismatch", "<u4"), ("n_gap_open", "<u4"), ("n_gap_ext", "<u4"), ("n_positive", "<u4"),
                   ("aln_len", "<u4"), ("q_frame", "i1"), ("s_frame", "i1"), ("phase", "u1"), ("reserved", "u1"),
                   ("bit_score", "<f8"), ("evalue", "<f8"), ("cigar_off", "<u4"), ("cigar_len", "<u4")])
assert HIT_DT.itemsize == 88
STATS_U64 = ("hits_after_seeding", "hits_failed_pre_extend", "hits_failed_evalue", "hits_failed_bitscore",
             "hits_failed_identity", "hits_duplicate", "hits_duplicate2", "hits_abundant", "hits_final", "pairs",
             "qrys_with_hit", "n_extensions_score", "n_extensions_trace", "cells_score", "cells_trace",
             "kernel_launches")
STATS_F32 = ("ms_seed", "ms_sort_merge", "ms_extend_score", "ms_extend_trace", "ms_h2d", "ms_d2h", "ms_total",
             "ms_host")
STATS_DT = np.dtype([(n, "<u8") for n in STATS_U64] + [(n, "<f4") for n in STATS_F32])
assert STATS_DT.itemsize == 160

DOMAIN = {"protein": 0, "nucleotide": 1, "bisulfite": 2}
AA27 = "ABCDEFGHIJKLMNOPQRSTUVWXYZ*"
DNA5 = "ACGNT"


def encode(seq_bytes: np.ndarray, domain: int) -> np.ndarray:
    """ASCII -> original-alphabet ranks (aa27 / dna5); unknown characters -> X / N like BioC++
    (BIO/alphabet/aminoacid/aa27.hpp:72-92, nucleotide/dna5.hpp:88-110)."""
    if domain == 0:
        tab = np.full(256, 23, np.uint8)
        for r, ch in enumerate(AA27):
            tab[ord(ch)] = r
            tab[ord(ch.lower())] = r
    else:
        tab = np.full(256, 3, np.uint8)
        for r, ch in enumerate(DNA5):
            tab[ord(ch)] = r
            tab[ord(ch.lower())] = r
        tab[ord("U")] = tab[ord("u")] = 4
    return tab[np.ascontiguousarray(seq_bytes, np.uint8)]


_FASTA_EXT = (".fasta", ".fa", ".fna", ".ffn", ".faa", ".frn", ".fas")
_FASTQ_EXT = (".fastq", ".fq")


def read_queries(path):
    """Query file reader with the reference's rules (bio::io::seq::reader, src/search_algo.hpp:342-348; the C++ host
    does the same in csrc/query_reader.hpp): FASTA or FASTQ by file extension, gzip/BGZF detected by magic bytes.
    Returns (ids, concatenated ASCII residues uint8, offsets uint64[n+1])."""
    import gzip
    low = path.lower()
    for z in (".gz", ".bgzf"):
        if low.endswith(z):
            low = low[: -len(z)]
            break
    if low.endswith(_FASTA_EXT):
        fastq = False
    elif low.endswith(_FASTQ_EXT):
        fastq = True
    else:
        raise ValueError(f"The query file's extension is not handled: {path}")
    with open(path, "rb") as f:
        magic = f.read(2)
    opener = gzip.open if magic == b"\x1f\x8b" else open
    ids, seqs = [], []
    with opener(path, "rb") as f:
        if fastq:
            while True:
                head = f.readline()
                if not head:
                    break
                head = head.rstrip(b"\r\n")
                if not head:
                    continue
                if not head.startswith(b"@"):
                    raise ValueError("ID-line does not begin with '@'.")
                seq = f.readline().rstrip(b"\r\n")
                plus = f.readline()
                qual = f.readline().rstrip(b"\r\n")
                if not plus.startswith(b"+"):
                    raise ValueError("Third FastQ record line does not begin with '+'.")
                if len(qual) != len(seq):
                    raise ValueError(f"Size mismatch between sequence ({len(seq)}) and qualities ({len(qual)}).")
                ids.append(head[1:].decode())
                seqs.append(seq)
        else:
            drop = bytes(range(48, 58)) + b" \t\r\n\x0b\x0c"  # digits and white space
            cur = None
            for line in f:
                if line[:1] in (b">", b";"):
                    if cur is not None:
                        seqs.append(b"".join(cur))
                    ids.append(line[1:].rstrip(b"\r\n").decode())
                    cur = []
                elif cur is not None:
                    cur.append(line.translate(None, drop))
                elif line.strip():
                    raise ValueError("Record does not begin with '>' or ';'.")
            if cur is not None:
                seqs.append(b"".join(cur))
            if any(len(s) == 0 for s in seqs):
                raise ValueError("No sequence or no valid sequence characters.")
    offs = np.zeros(len(seqs) + 1, np.uint64)
    if seqs:
        np.cumsum([len(s) for s in seqs], out=offs[1:])
    data = np.frombuffer(b"".join(seqs), np.uint8)
    return ids, data, offs


read_fasta = read_queries  # older name
