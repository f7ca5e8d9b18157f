// Traceback of DP pass 2 on residue planes: ONE byte per DP cell (the reference's trace matrix is also one byte per
// cell, SQ/align/dp_profile.h:122), but the byte is H(i,j) mod 256 instead of SeqAn's 7 trace bits -- it falls out of
// the packed fill loop (swDpxKernel<.., TRACE = true>, kernels_dpx.cuh) with half a PRMT per cell pair, and the
// traceback rebuilds exactly the decisions SeqAn's _doTraceback takes on its trace bits.
//
// Reference: _computeTraceback / _doTraceback (SQ/align/dp_traceback_impl.h:223-474), trace bits as built in
// SQ/align/dp_formula_affine.h:66-126 (CompleteTrace: ties set both bits), computeAlignmentStats
// (SQ/align/evaluate_alignment.h:215-300).  tools/residue_traceback_model.py is an executable model of the rules
// below, checked against a restatement of SeqAn's trace-byte traceback on random inputs full of ties.
//
//  * The exact score h of the current cell is known all along the path: it starts as the best score and every move
//    changes it by a known amount.
//  * Cells next to each other differ by a bounded amount: |H(i,j) - H(i,j-1)| <= Mmax - go (remove the last subject
//    residue from the best alignment ending in (i,j): a match column becomes a gap column), the same along a row,
//    and H(i,j) - M[q_i][t_j] >= H(i-1,j-1) >= H(i,j) - 2 (Mmax - go).  With 2 (Mmax - go) - Mmin < 256 (checked on
//    the host; 50 for BLOSUM62 11/1) a residue next to a cell of known score identifies its exact score.
//  * DIAGONAL is set  <=>  H(i-1,j-1) + M[q_i][t_j] == h  <=>  residue(i-1,j-1) == (h - M) mod 256.  The main loop
//    tests DIAGONAL first (dp_traceback_impl.h:390), so nothing else of the cell is needed when it holds.  32 cells of
//    the diagonal are tested per round: lane k assumes the k cells in front of it were diagonal, which makes its own
//    h a prefix sum of substitution scores; the longest prefix of lanes whose test holds is committed.
//  * Otherwise the cell was entered from a gap.  The vertical gap value is E(i,j) = max_k H(i,j-k) + go + (k-1) ge; the
//    exact H of the cells above follow from chaining their residues.  MAX_FROM_VERTICAL <=> E == h is tried first
//    (:401-413), and SeqAn's "keep going while the extend bit is set, then one more step" loop (:240-256) ends on
//    the LARGEST k that attains the maximum (extend is set as long as a longer gap ties).  The same along the row
//    for horizontal gaps (:414-421, :319-335).  A gap of k characters ending in a cell of score h started from a
//    cell of score h - go - (k-1) ge <= best score, which bounds k.
//  * The end cell (first strict maximum in column-major order, SQ/align/dp_scout_simd.h:216-229) can never be the
//    end of a gap when go < 0, so the walk always starts diagonally and PreferGapsAtEnd (:458-474) never triggers.
#pragma once

#include <cstdint>
#include <cuda_runtime.h>

#include "../../include/lambda_b200.h"
#include "kernels_dpx.cuh"
#include "kernels_extend.cuh"

namespace lgpu
{

struct TracebackResParams
{
    DevIndex                   ix;
    DevQueries                 Q;
    lgpu_match const *         tasks;
    unsigned int const *       order;    // sorted slot -> task
    unsigned long long const * keys;     // sort keys: class of a slot = key >> kDpxClassShift
    unsigned int               slotFirst; // this launch walks the sorted slots [slotFirst, slotFirst + nSlots)
    unsigned int               nSlots;
    int                        tab;      // class table the fill kernels used (kDpxTabPriv / kDpxTabTrace32)
    signed char const *        matrix;   // 2 x (32 x 32)
    int                        go, ge;
    int const *                scores;   // indexed by task
    unsigned int const *       bestCol;  // indexed by task
    unsigned int const *       planes;
    unsigned long long const * planeOff; // per sorted slot
    unsigned long long         planeOffBase; // offset of the first plane of this launch's group inside `planes`
    lgpu_hit *                 out;      // indexed by task
    // second pass (lgpu_params.want_cigar): emit the runs of the path instead of the record
    int                        emit;
    unsigned int *             cigarOps;  // run << 2 | kind, traceback order
    unsigned int const *       cigarOff;  // per slot of this launch: first op of the alignment
    unsigned int               cigarBase; // added to the offsets stored in the records
};

__device__ __forceinline__ int resCentered(unsigned int d)
{
    return static_cast<int>((d + 128u) & 255u) - 128;
}

__device__ __forceinline__ int warpInclusiveScanInt(int v, unsigned int lane)
{
#pragma unroll
    for (int off = 1; off < 32; off <<= 1)
    {
        int const n = __shfl_up_sync(0xffffffffu, v, off);
        if (lane >= static_cast<unsigned int>(off))
            v += n;
    }
    return v;
}

constexpr int kTbWarps = 4;

// one warp per alignment
__global__ void __launch_bounds__(32 * kTbWarps) tracebackResKernel(TracebackResParams P)
{
    unsigned int const lane = threadIdx.x & 31u;
    unsigned int const wid  = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (wid >= P.nSlots)
        return;
    unsigned int const slot = P.slotFirst + wid;
    unsigned int const       task = P.order[slot];
    DpxClass const           kc   = dpxClass(P.tab, static_cast<int>(P.keys[slot] >> kDpxClassShift));
    unsigned int const       T = static_cast<unsigned int>(kc.T), K = static_cast<unsigned int>(kc.K), KN = (K + 1) / 2;
    lgpu_match const         m    = P.tasks[task];
    unsigned int const       q    = m.qry_id / P.Q.F;
    unsigned int const       f    = m.qry_id % P.Q.F;
    unsigned long long const qb   = P.Q.offs[q];
    unsigned int const       qLen = static_cast<unsigned int>(P.Q.offs[q + 1] - qb);
    unsigned char const *    qs   = P.Q.trans + P.Q.F * qb + static_cast<unsigned long long>(f) * qLen + m.qry_start;
    unsigned int const       sId  = m.subj_id / P.ix.sbjFrames;
    unsigned char const *    ts   = P.ix.seqs + sbjBase(P.ix, m.subj_id) + m.subj_start;
    signed char const *      M    = P.matrix + matrixOffset(P.ix, m.subj_id);
    unsigned int const       nt   = m.subj_end - m.subj_start;
    unsigned int const *     plane = P.planes + (P.planeOff[slot] - P.planeOffBase);
    int const                go = P.go, ge = P.ge;

    // residue of H(i,j), 1 <= i <= nq, 1 <= j <= nt: strip v = (i-1)/K works on row j at step (j-1) + v
    auto res = [&](unsigned int i, unsigned int j) -> unsigned int {
        unsigned int const v = (i - 1) / K, r = (i - 1) - v * K;
        unsigned int const half = v >= T ? 1u : 0u, p = v - half * T;
        unsigned long long const s = static_cast<unsigned long long>(j - 1) + v;
        unsigned int const w = __ldg(plane + (s * KN + (r >> 1)) * T + p);
        return (((w >> (8u * (2u * (r & 1u) + half))) & 255u) - static_cast<unsigned int>(go)) & 255u; // stored: W = H + go
    };

    int const    score = P.scores[task];
    unsigned int i = (score > 0) ? P.bestCol[task] : 0u, j = 0; // no positive cell: empty alignment at (0, 0)
    if (score > 0)
    {
        // first row of the best column that holds the best score: exact H down the column from the border H(i,0) = 0
        int          carryH   = 0;
        unsigned int carryRes = 0;
        for (unsigned int base = 0; base < nt && j == 0; base += 32)
        {
            unsigned int const jj   = base + lane + 1;
            unsigned int const r    = jj <= nt ? res(i, jj) : 0u;
            unsigned int       prev = __shfl_up_sync(0xffffffffu, r, 1);
            if (lane == 0)
                prev = carryRes;
            int const          H   = carryH + warpInclusiveScanInt(jj <= nt ? resCentered(r - prev) : 0, lane);
            unsigned int const hit = __ballot_sync(0xffffffffu, jj <= nt && H == score);
            if (hit)
                j = base + static_cast<unsigned int>(__ffs(hit));
            carryH   = __shfl_sync(0xffffffffu, H, 31);
            carryRes = __shfl_sync(0xffffffffu, r, 31);
        }
    }
    unsigned int const bi = i, bj = j;
    unsigned int nMatch = 0, nMismatch = 0, nPositive = 0, nGapOpen = 0, nGapExt = 0, alnLen = 0;
    unsigned int nOps = 0;
    unsigned int * const ops = P.emit ? P.cigarOps + P.cigarOff[wid] : nullptr;

    int          h    = (j > 0) ? score : 0;
    int          last = 0; // 0 diag, 1 horizontal, 2 vertical
    unsigned int run  = 0;
    auto flush = [&]() {
        if (run)
        {
            if (ops && lane == 0)
                ops[nOps] = (run << 2) | static_cast<unsigned int>(last);
            ++nOps;
            alnLen += run;
            if (last != 0)
            {
                nGapOpen += 1;
                nGapExt += run - 1;
            }
        }
    };
    auto switchTo = [&](int k) {
        if (last != k)
        {
            flush();
            last = k;
            run  = 0;
        }
    };
    // largest k such that a gap of k characters ending in the current cell attains its score h (0: none)
    auto gapScan = [&](bool vertical) -> unsigned int {
        unsigned int kmax = vertical ? j : i;
        if (ge < 0)
        {
            int const room = score - h + go; // the cell the gap started from scores h - go - (k-1) ge <= score
            if (room < 0)
                return 0u;
            kmax = min(kmax, static_cast<unsigned int>(room / (-ge)) + 1u);
        }
        int          carryH   = h;
        unsigned int carryRes = static_cast<unsigned int>(h) & 255u;
        unsigned int kbest    = 0;
        for (unsigned int base = 0; base < kmax; base += 32)
        {
            unsigned int const k      = base + lane + 1;
            bool const         act    = k <= kmax;
            bool const         border = act && (vertical ? j - k : i - k) == 0;
            unsigned int const r      = (act && !border) ? (vertical ? res(i, j - k) : res(i - k, j)) : 0u;
            unsigned int       prev   = __shfl_up_sync(0xffffffffu, r, 1);
            if (lane == 0)
                prev = carryRes;
            int H = carryH + warpInclusiveScanInt(act ? resCentered(r - prev) : 0, lane);
            if (border)
                H = 0;
            bool const         att = act && H + go + static_cast<int>(k - 1) * ge == h;
            unsigned int const bal = __ballot_sync(0xffffffffu, att);
            if (bal)
                kbest = base + 32u - static_cast<unsigned int>(__clz(bal));
            carryH   = __shfl_sync(0xffffffffu, H, 31);
            carryRes = __shfl_sync(0xffffffffu, r, 31);
        }
        return kbest;
    };

    while (i > 0 && j > 0 && h > 0)
    {
        // ---- up to 32 diagonal steps ----
        bool const   inside = lane < i && lane < j;
        unsigned int a = 0, b = 0;
        int          mk = 0;
        if (inside)
        {
            a  = qs[i - lane - 1];
            b  = ts[j - lane - 1];
            mk = M[a * 32 + b];
        }
        int const  pNext  = h - warpInclusiveScanInt(mk, lane); // score of the diagonal predecessor of this lane's cell
        int const  pCur   = pNext + mk;
        bool const border = inside && (i - lane == 1 || j - lane == 1);
        bool       ok     = inside && pCur > 0 && pNext >= 0;
        if (ok)
            ok = border ? pNext == 0 : res(i - lane - 1, j - lane - 1) == (static_cast<unsigned int>(pNext) & 255u);
        unsigned int const okMask = __ballot_sync(0xffffffffu, ok);
        unsigned int const n      = okMask == 0xffffffffu ? 32u : static_cast<unsigned int>(__ffs(~okMask)) - 1u;
        if (n)
        {
            unsigned int const cm = n == 32 ? 0xffffffffu : (1u << n) - 1u;
            switchTo(0);
            nMatch += __popc(__ballot_sync(0xffffffffu, inside && alignedIdentical(P.ix, M, a, b)) & cm);
            nPositive += __popc(__ballot_sync(0xffffffffu, inside && mk > 0) & cm);
            run += n;
            h = n == 32 ? __shfl_sync(0xffffffffu, pNext, 31) : __shfl_sync(0xffffffffu, pCur, n);
            i -= n;
            j -= n;
            if (n == 32)
                continue;
        }
        if (!(i > 0 && j > 0 && h > 0))
            break;
        // ---- the cell was entered from a gap: vertical first, then horizontal ----
        unsigned int k = gapScan(true);
        if (k)
        {
            switchTo(2);
            j -= k;
        }
        else
        {
            k = gapScan(false);
            if (!k)
                break; // cannot happen on planes written by the fill kernel
            switchTo(1);
            i -= k;
        }
        run += k;
        h = h - go - static_cast<int>(k - 1) * ge;
    }
    flush();

    if (P.emit)
    {
        if (lane == 0)
        {
            P.out[task].cigar_off = P.cigarBase + P.cigarOff[wid];
            P.out[task].cigar_len = nOps;
        }
        return;
    }
    if (lane != 0)
        return;
    nMismatch = alnLen - nGapOpen - nGapExt - nMatch; // aligned columns that are not identical
    lgpu_hit hh;
    hh.q_id       = q;
    hh.s_id       = sId;
    hh.q_start    = m.qry_start + i;
    hh.q_end      = m.qry_start + bi;
    hh.s_start    = m.subj_start + j;
    hh.s_end      = m.subj_start + bj;
    hh.q_len      = qLen;
    hh.s_len      = static_cast<unsigned int>(P.ix.origDelims[sId + 1] - P.ix.origDelims[sId]);
    hh.score      = score;
    hh.n_match    = nMatch;
    hh.n_mismatch = nMismatch;
    hh.n_gap_open = nGapOpen;
    hh.n_gap_ext  = nGapExt;
    hh.n_positive = nPositive;
    hh.aln_len    = alnLen;
    setFrames(P.Q, P.ix, m.qry_id, m.subj_id, hh.q_frame, hh.s_frame);
    hh.phase      = 0;
    hh.reserved   = 0;
    hh.bit_score  = 0.0;
    hh.evalue     = 0.0;
    hh.cigar_off  = 0;
    hh.cigar_len  = 0;
    P.out[task]   = hh;
}

} // namespace lgpu
