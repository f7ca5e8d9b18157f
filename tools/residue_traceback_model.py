#!/usr/bin/env python
"""Executable model of the residue-plane traceback (kernels_dpx_trace.cuh) against the reference's trace-byte
traceback (SQ/align/dp_formula_affine.h:66-126, dp_traceback_impl.h:223-474).

The fill kernel stores ONE byte per DP cell: H(i,j) mod 256.  This script checks, on random inputs with many ties
(small alphabets, small gap costs), that the path rebuilt from the residues alone equals the path SeqAn's
traceback walks on its 7-bit trace bytes:

  * the score of the current cell is tracked exactly along the path (it starts at the best score);
  * DIAG  <=>  res(i-1,j-1) == (h - M[q_i][t_j]) mod 256          (neighbouring cells differ by < 128);
  * otherwise the cell came out of a gap: the vertical gap value is E(i,j) = max_k H(i,j-k) + go + (k-1) ge, the
    exact H of the column above is recovered by chaining the residues (vertical neighbours differ by at most
    Mmax - go), and SeqAn's "stay in the gap while the extend bit is set" loop ends at the LARGEST k that attains
    the maximum; the same along the row for horizontal gaps; vertical is tried first (MAX_FROM_VERTICAL is tested
    first, dp_traceback_impl.h:401-421);
  * the end cell of a local alignment can never be the end of a gap when go < 0, so the walk starts diagonally.

    python tools/residue_traceback_model.py [n_cases]
"""
import sys

import numpy as np

DIAG, HORI, VERT, HOPEN, VOPEN, MAXH, MAXV = 1, 2, 4, 8, 16, 32, 64
NEG = -16384


def fill(q, t, M, go, ge):
    nq, nt = len(q), len(t)
    S = np.zeros((nq + 1, nt + 1), np.int64)
    T = np.zeros((nq + 1, nt + 1), np.int64)
    Sp = np.zeros(nt + 1, np.int64)
    Hp = np.full(nt + 1, NEG, np.int64)
    best, bi, bj = 0, 0, 0
    for i in range(1, nq + 1):
        v, s_up = NEG, 0
        Sc = np.zeros(nt + 1, np.int64)
        Hc = np.full(nt + 1, NEG, np.int64)
        for j in range(1, nt + 1):
            diag = Sp[j - 1] + M[q[i - 1]][t[j - 1]]
            a, b = Hp[j] + ge, Sp[j] + go
            h, tv = (a, HORI | HOPEN) if a == b else ((b, HOPEN) if a < b else (a, HORI))
            a, b = v + ge, s_up + go
            v, tvv = (a, VERT | VOPEN) if a == b else ((b, VOPEN) if a < b else (a, VERT))
            tv |= tvv
            g, t2 = (v, MAXV | MAXH) if v == h else ((h, MAXH) if v < h else (v, MAXV))
            if diag == g:
                cur, tv = diag, DIAG | tv | t2
            elif diag < g:
                cur, tv = g, tv | t2
            else:
                cur, tv = diag, DIAG | tv
            if cur <= 0:
                cur, tv = 0, 0
            Sc[j], Hc[j], T[i][j], s_up = cur, h, tv, cur
            if cur > best:
                best, bi, bj = cur, i, j
        S[i] = Sc
        Sp, Hp = Sc, Hc
    return S, T, best, bi, bj


def traceback_ref(T, bi, bj):
    """SURVEY B.4 pseudo-code; returns (i, j, runs end->begin as (kind, len))"""
    i, j = bi, bj
    tv = T[i][j]
    if tv & MAXV:
        tv &= (VERT | VOPEN | MAXV)
        last = 2
    elif tv & MAXH:
        tv &= (HORI | HOPEN | MAXH)
        last = 1
    else:
        last = 0
    run, segs = 0, []

    def switch(k):
        nonlocal last, run
        if last != k:
            if run:
                segs.append((last, run))
            last, run = k, 0

    while i > 0 and j > 0 and tv != 0:
        if tv & DIAG:
            switch(0); i -= 1; j -= 1; tv = T[i][j]; run += 1
        elif (tv & MAXV) and (tv & VERT):
            switch(2)
            while ((not tv & VOPEN) or (tv & VERT)) and j != 1:
                j -= 1; tv = T[i][j]; run += 1
            j -= 1; tv = T[i][j]; run += 1
        elif (tv & MAXV) and (tv & VOPEN):
            switch(2); j -= 1; tv = T[i][j]; run += 1
        elif (tv & MAXH) and (tv & HORI):
            switch(1)
            while ((not tv & HOPEN) or (tv & HORI)) and i != 1:
                i -= 1; tv = T[i][j]; run += 1
            i -= 1; tv = T[i][j]; run += 1
        elif (tv & MAXH) and (tv & HOPEN):
            switch(1); i -= 1; tv = T[i][j]; run += 1
        else:
            break
    if run:
        segs.append((last, run))
    return i, j, segs


def centered(d):
    return ((int(d) + 128) & 255) - 128


def traceback_res(R, q, t, M, go, ge, score, bi, bj):
    """the same walk from residues R = H mod 256 only"""
    nq, nt = len(q), len(t)

    def res(i, j):
        return 0 if i == 0 or j == 0 else int(R[i][j])

    i, j, h = bi, bj, score
    last, run, segs = 0, 0, []

    def switch(k):
        nonlocal last, run
        if last != k:
            if run:
                segs.append((last, run))
            last, run = k, 0

    def scan(vertical):
        """largest k with H(cell k steps back) + go + (k-1) ge == h, or 0"""
        n = j if vertical else i
        kmax = n if ge == 0 else min(n, (score - h + go) // (-ge) + 1)
        prev_res, cur, kbest = h & 255, h, 0
        for k in range(1, kmax + 1):
            r = res(i, j - k) if vertical else res(i - k, j)
            border = (j - k == 0) if vertical else (i - k == 0)
            cur = 0 if border else cur + centered(r - prev_res)
            prev_res = r
            if cur + go + (k - 1) * ge == h:
                kbest = k
        return kbest

    while i > 0 and j > 0 and h > 0:
        m = int(M[q[i - 1]][t[j - 1]])
        if res(i - 1, j - 1) == ((h - m) & 255):
            switch(0); i -= 1; j -= 1; h -= m; run += 1
            continue
        k = scan(True)
        if k:
            switch(2); j -= k; run += k; h = h - go - (k - 1) * ge
            continue
        k = scan(False)
        if k:
            switch(1); i -= k; run += k; h = h - go - (k - 1) * ge
            continue
        raise AssertionError("cell explained by neither the diagonal nor a gap")
    if run:
        segs.append((last, run))
    return i, j, segs


def main():
    n_cases = int(sys.argv[1]) if len(sys.argv) > 1 else 400
    rng = np.random.default_rng(7)
    bad = 0
    for case in range(n_cases):
        sigma = int(rng.integers(2, 6))
        mmax, mmin = int(rng.integers(1, 12)), -int(rng.integers(1, 6))
        M = rng.integers(mmin, mmax + 1, (sigma, sigma))
        M = np.minimum(M, M.T)
        for a in range(sigma):
            M[a][a] = int(rng.integers(1, mmax + 1))
        ge = -int(rng.integers(0 if case % 5 == 0 else 1, 4))
        go = ge - int(rng.integers(1 if ge == 0 else 0, 13))  # go <= ge, go < 0
        nq, nt = int(rng.integers(1, 60)), int(rng.integers(1, 80))
        t = rng.integers(0, sigma, nt)
        q = rng.integers(0, sigma, nq)
        if nq > 8 and nt > nq and case % 2 == 0:  # a mutated copy: long alignments with gaps
            s0 = int(rng.integers(0, nt - nq + 1))
            q = t[s0:s0 + nq].copy()
            for _ in range(int(rng.integers(0, 4))):
                p = int(rng.integers(0, nq))
                q = np.concatenate([q[:p], q[p + int(rng.integers(1, 4)):], rng.integers(0, sigma, 3)])[:nq]
            mut = rng.random(len(q)) < 0.1
            q[mut] = rng.integers(0, sigma, int(mut.sum()))
        assert 2 * (mmax - go) - mmin < 256
        S, T, best, bi, bj = fill(q, t, M, go, ge)
        if best == 0:
            continue
        ref = traceback_ref(T, bi, bj)
        got = traceback_res(S & 255, q, t, M, go, ge, best, bi, bj)
        if ref != got:
            bad += 1
            print("MISMATCH case", case, "go", go, "ge", ge, "ref", ref, "got", got)
    print(f"{n_cases} cases, {bad} mismatches")
    return 1 if bad else 0


if __name__ == "__main__":
    sys.exit(main())
