"""lambda_b200 -- B200-native seed-and-extend engine behind the lambda3 searchp/searchn interface.

The product is lambda_b200/liblambda_b200.so (hand-written sm_100a CUDA kernels + a C ABI,
include/lambda_b200.h).  This package is the thin host-side binding: it loads the library with
ctypes and mirrors the reference's search interface (index + options in, BLAST-tabular hits out).
There is no CPU fallback: importing works anywhere, but every compute call needs the CUDA library
and a GPU and raises otherwise.
"""
from ._abi import DOMAIN, HIT_DT, MATCH_DT, STATS_DT, Params, encode, read_fasta, read_queries  # noqa: F401
from .api import Index, LambdaError, Searcher, load_library  # noqa: F401

__all__ = ["Index", "Searcher", "LambdaError", "load_library", "Params", "encode", "read_fasta", "read_queries", "HIT_DT",
           "MATCH_DT", "STATS_DT", "DOMAIN"]
