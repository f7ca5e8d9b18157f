"""The rules of the residue-plane traceback (lambda_b200/csrc/kernels_dpx_trace.cuh) as an executable model, checked on
the CPU against a restatement of SeqAn's trace-byte traceback on random inputs full of ties."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tools"))
import residue_traceback_model as m  # noqa: E402


def test_random_scoring_schemes_with_ties(monkeypatch):
    monkeypatch.setattr(sys, "argv", ["residue_traceback_model.py", "250"])
    assert m.main() == 0


def test_gappy_nucleotide_alignments_and_residue_wrap():
    """mutated copies with indels (many gap runs) and scores far above 256 (the residues wrap several times)"""
    rng = np.random.default_rng(3)
    M = np.full((4, 4), -3)
    np.fill_diagonal(M, 2)
    runs = 0
    for case in range(60):
        nt, nq = 400, 220
        t = rng.integers(0, 4, nt)
        s0 = int(rng.integers(0, nt - nq))
        q = t[s0:s0 + nq].copy()
        for _ in range(3):
            p = int(rng.integers(5, len(q) - 5))
            q = np.concatenate([q[:p], q[p + 2:], rng.integers(0, 4, 2)])
        mut = rng.random(len(q)) < 0.04
        q[mut] = rng.integers(0, 4, int(mut.sum()))
        S, T, best, bi, bj = m.fill(q, t, M, -7, -2)
        assert best > 256
        ref = m.traceback_ref(T, bi, bj)
        got = m.traceback_res(S & 255, q, t, M, -7, -2, best, bi, bj)
        assert ref == got, case
        runs += sum(1 for k, _ in ref[2] if k)
    assert runs > 60
