// DP pass 2 on the packed-int16 DPX instructions: instead of SeqAn's 7-bit trace byte per cell the fill
// kernel stores what is needed to reconstruct that byte exactly, and the traceback rebuilds it only for
// the ~600 cells it visits.
//
// Reference: _performAlignment<withTrace=true> (src/search_algo.hpp:1296), trace bits as defined in
// SQ/align/dp_formula.h:136-243 + dp_formula_affine.h:66-126 (CompleteTrace: ties set both bits),
// _doTraceback (SQ/align/dp_traceback_impl.h:223-474), computeAlignmentStats
// (SQ/align/evaluate_alignment.h:215-300).
//
// Per cell (i = query column, j = subject row) the fill kernel stores
//     W  = H + go                       16 bit  (plane H, the register image of the wavefront)
//     dE = min(H - E(i,j), 15)           4 bit   E(i,j) = vertical gap value entering the cell
//     dF = min(H - F(i,j), 15)           4 bit   F(i,j) = horizontal gap value entering the cell
// With D = ge - go (<= 14) the reference's decisions are functions of these:
//     HORI  <=> dF(i-1,j) <= D      HOPEN <=> dF(i-1,j) >= D      (border column: open only)
//     VERT  <=> dE(i,j-1) <= D      VOPEN <=> dE(i,j-1) >= D      (border row:    open only)
//     g = max(E,F) = H - min(dE,dF);  MAXV <=> dE <= dF;  MAXH <=> dF <= dE
//     min(dE,dF) > 0  ->  H came from the diagonal alone: DIAG, MAX_FROM_* bits not set
//     min(dE,dF) == 0 ->  MAX_FROM_* bits set, DIAG <=> H(i-1,j-1) + M[q_i][t_j] == H
//     H == 0          ->  trace = 0
// The fill loop is the score kernel's (kernels_dpx.cuh) plus five packed instructions per two cells
// (two subtractions, two clips, the per-column running maximum) and two integer FMAs that pack the
// nibbles; one warp owns one alignment (T = 32: 64 strips of K columns, both int16 halves on the same
// alignment).  The end cell follows the reference's rule (first strict maximum in column-major order):
// per-column maxima in the fill kernel pick the smallest column holding the best score, the traceback
// kernel finds the first row of that column.
#pragma once

#include <cstdint>
#include <cuda_runtime.h>

#include "../../include/lambda_b200.h"
#include "kernels_dpx.cuh"
#include "kernels_extend.cuh"

namespace lgpu
{

struct DpxTraceParams
{
    DevIndex                   ix;
    DevQueries                 Q;
    lgpu_match const *         tasks;
    unsigned int const *       order;    // tasks of this class (indices into `tasks`)
    unsigned int               nTasks;   // entries in `order`
    signed char const *        matrix;   // 2 x (32 x 32)
    int                        go, ge;
    unsigned int               nCodes;   // alphabet size + 1 (last row = null)
    unsigned int               winCap;   // bytes reserved for the padded window
    unsigned int *             workCounter;
    unsigned int *             planes;   // per task: H plane then N plane, 32-bit words
    unsigned long long const * planeOff; // word offset of every task's planes (indexed by task)
    int *                      scores;   // out, indexed by task
    unsigned int *             bestCol;  // out, indexed by task: 1-based column of the end cell
};

__host__ __device__ constexpr unsigned int dpxTraceKN(int K) // nibble words per lane and step
{
    return static_cast<unsigned int>((K + 1) / 2);
}
// words of both planes of one alignment with `nt` subject rows
__host__ __device__ inline unsigned long long dpxTracePlaneWords(int K, unsigned int nt)
{
    return static_cast<unsigned long long>(nt + 63) * 32ull * (static_cast<unsigned int>(K) + dpxTraceKN(K));
}

template <int K>
__global__ void __launch_bounds__(32) swTraceDpxKernel(DpxTraceParams P)
{
    constexpr int T    = 32;
    constexpr int KW   = (K + 3) / 4;
    constexpr int KN   = (K + 1) / 2;
    constexpr int ROWW = dpxRowWords(T, K);
    constexpr int PAD  = 2 * T;

    extern __shared__ unsigned int smem[];
    unsigned int const lane      = threadIdx.x;
    unsigned int const profWords = P.nCodes * ROWW;
    unsigned int *     prof      = smem;
    unsigned char *    win       = reinterpret_cast<unsigned char *>(smem + profWords);
    unsigned int const nullCode  = P.nCodes - 1;

    unsigned int const go2   = (static_cast<unsigned int>(P.go) & 0xffffu) * 0x10001u;
    unsigned int const ge2   = (static_cast<unsigned int>(P.ge) & 0xffffu) * 0x10001u;
    unsigned int const neg2  = 0xE000E000u; // -8192: far below any real gap value, and H - E cannot overflow int16
    unsigned int const clip2 = 0x000F000Fu;

    for (;;)
    {
        unsigned int slot = 0;
        if (lane == 0)
            slot = atomicAdd(P.workCounter, 1u);
        slot = __shfl_sync(0xffffffffu, slot, 0);
        if (slot >= P.nTasks)
            break;
        unsigned int const       task = P.order[slot];
        lgpu_match const         m    = P.tasks[task];
        unsigned int const       q    = m.qry_id / P.Q.F;
        unsigned int const       f    = m.qry_id % P.Q.F;
        unsigned long long const qb   = P.Q.offs[q];
        unsigned int const       qLen = static_cast<unsigned int>(P.Q.offs[q + 1] - qb);
        unsigned char const *    qs   = P.Q.trans + P.Q.F * qb + static_cast<unsigned long long>(f) * qLen + m.qry_start;
        unsigned int const       nq   = m.qry_end - m.qry_start;
        unsigned char const *    ts   = P.ix.seqs + sbjBase(P.ix, m.subj_id) + m.subj_start;
        signed char const *      M    = P.matrix + matrixOffset(P.ix, m.subj_id);
        unsigned int const       nt   = m.subj_end - m.subj_start;
        unsigned int const       nSteps = nt + 2 * T - 1;
        unsigned int *           planeH = P.planes + P.planeOff[task];
        unsigned int *           planeN = planeH + static_cast<unsigned long long>(nSteps) * 32ull * K;

        __syncwarp();
        for (unsigned int idx = lane; idx < profWords; idx += 32)
        {
            unsigned int const c   = idx / ROWW;
            unsigned int const rem = idx % ROWW;
            unsigned int const w   = rem / (2 * T);
            unsigned int const v   = rem % (2 * T);
            unsigned int       word = 0x80808080u; // null = -128
            if (c != nullCode && w < KW)
            {
                word = 0;
#pragma unroll
                for (int b = 0; b < 4; ++b)
                {
                    unsigned int const r = w * 4 + b;
                    unsigned int const i = v * K + r;
                    int                val = -128;
                    if (r < K && i < nq)
                        val = static_cast<int>(M[qs[i] * 32 + c]) - P.go;
                    word |= (static_cast<unsigned int>(val) & 0xffu) << (8 * b);
                }
            }
            prof[idx] = word;
        }
        for (unsigned int idx = lane; idx < P.winCap; idx += 32)
        {
            int const     j = static_cast<int>(idx) - PAD;
            unsigned char c = static_cast<unsigned char>(nullCode);
            if (j >= 0 && j < static_cast<int>(nt))
                c = ts[j];
            win[idx] = c;
        }
        __syncwarp();

        unsigned int E[K], W[K], CB[K];
#pragma unroll
        for (int r = 0; r < K; ++r)
        {
            E[r]  = neg2;
            W[r]  = go2; // H = 0
            CB[r] = go2; // per-column maximum of W
        }
        unsigned int outW = go2, outF = neg2, diagIn = go2;

        unsigned int wl[KW], wh[KW];
        {
            unsigned int const cLo = win[PAD - lane];
            unsigned int const cHi = win[PAD - lane - T];
#pragma unroll
            for (int k = 0; k < KW; ++k)
            {
                wl[k] = prof[cLo * ROWW + k * 2 * T + lane];
                wh[k] = prof[cHi * ROWW + k * 2 * T + T + lane];
            }
        }
        for (unsigned int s = 0; s < nSteps; ++s)
        {
            unsigned int nl[KW], nh[KW];
            {
                unsigned int const cLo = win[PAD + s + 1 - lane];
                unsigned int const cHi = win[PAD + s + 1 - lane - T];
#pragma unroll
                for (int k = 0; k < KW; ++k)
                {
                    nl[k] = prof[cLo * ROWW + k * 2 * T + lane];
                    nh[k] = prof[cHi * ROWW + k * 2 * T + T + lane];
                }
            }
            unsigned int inW = __shfl_sync(0xffffffffu, outW, (lane - 1) & 31u);
            unsigned int inF = __shfl_sync(0xffffffffu, outF, (lane - 1) & 31u);
            if (lane == 0)
            {
                inW = prmt(go2, inW, 0x5410);
                inF = prmt(neg2, inF, 0x5410);
            }
            unsigned int diag = diagIn;
            diagIn            = inW;
            unsigned int F    = inF;
            unsigned int nib[K];
#pragma unroll
            for (int r = 0; r < K; ++r)
            {
                unsigned int const b   = r & 3;
                unsigned int const sel = ((0xCu + b) << 12) | ((4u + b) << 8) | ((8u + b) << 4) | b;
                unsigned int const sub = prmt(wl[r >> 2], wh[r >> 2], sel);
                unsigned int const t   = __viaddmax_s16x2_relu(diag, sub, E[r]);
                unsigned int const u   = __vadd2(t, go2);
                unsigned int const w   = __viaddmax_s16x2(F, go2, u);
                // what the traceback needs of this cell: H - E and H - F of the gap values that ENTER it
                unsigned int const h   = __vsub2(w, go2);
                unsigned int const dE  = __vmins2(__vsub2(h, E[r]), clip2);
                unsigned int const dF  = __vmins2(__vsub2(h, F), clip2);
                nib[r]                 = dF * 16u + dE; // both halves stay below 256: no carry between them
                F                      = __viaddmax_s16x2(F, ge2, u);
                E[r]                   = __viaddmax_s16x2(E[r], ge2, w);
                diag                   = W[r];
                W[r]                   = w;
                CB[r]                  = __vmaxs2(CB[r], w);
            }
            outW = W[K - 1];
            outF = F;
            // the register image of this step: lane p holds strip p (low half, row s - p) and strip p + 32
            // (high half, row s - p - 32)
            unsigned int * dstH = planeH + (static_cast<unsigned long long>(s) * 32u + lane) * K;
#pragma unroll
            for (int r = 0; r < K; ++r)
                dstH[r] = W[r];
            unsigned int * dstN = planeN + (static_cast<unsigned long long>(s) * 32u + lane) * KN;
#pragma unroll
            for (int w2 = 0; w2 < KN; ++w2)
            {
                // bytes: [low cell r, low cell r+1, high cell r, high cell r+1]
                unsigned int const a = nib[2 * w2];
                unsigned int const c = (2 * w2 + 1 < K) ? nib[2 * w2 + 1] : 0u;
                dstN[w2]             = c * 256u + a;
            }
#pragma unroll
            for (int k = 0; k < KW; ++k)
            {
                wl[k] = nl[k];
                wh[k] = nh[k];
            }
        }
        // best score and the smallest column that holds it (padded columns never reach the maximum)
        int          best = P.go;
        unsigned int bcol = 0xffffffffu;
#pragma unroll
        for (int r = 0; r < K; ++r)
        {
            int const          lo = static_cast<int>(static_cast<short>(CB[r] & 0xffffu));
            int const          hi = static_cast<int>(CB[r]) >> 16;
            unsigned int const cl = lane * K + r, ch = (lane + T) * K + r; // 0-based columns
            if (lo > best || (lo == best && cl < bcol)) { best = lo; bcol = cl; }
            if (hi > best || (hi == best && ch < bcol)) { best = hi; bcol = ch; }
        }
#pragma unroll
        for (int off = 16; off > 0; off >>= 1)
        {
            int const          ob = __shfl_xor_sync(0xffffffffu, best, off);
            unsigned int const oc = __shfl_xor_sync(0xffffffffu, bcol, off);
            if (ob > best || (ob == best && oc < bcol)) { best = ob; bcol = oc; }
        }
        if (lane == 0)
        {
            P.scores[task]  = best - P.go;
            P.bestCol[task] = bcol + 1;
        }
    }
}

// ---------------------------------------------------------------------------------------------
// traceback on the stored planes
// ---------------------------------------------------------------------------------------------

struct TracebackDpxParams
{
    DevIndex                   ix;
    DevQueries                 Q;
    lgpu_match const *         tasks;
    unsigned int const *       order;
    unsigned int               nTasks;
    signed char const *        matrix; // 2 x (32 x 32)
    int                        go, ge;
    unsigned char const *      kOf; // per task: columns per strip of the fill kernel that wrote its planes
    int const *                scores;
    unsigned int const *       bestCol;
    unsigned int const *       planes;
    unsigned long long const * planeOff;
    lgpu_hit *                 out; // indexed by task
    // second pass (lgpu_params.want_cigar): emit the runs of the path instead of the record
    int                        emit;
    unsigned int *             cigarOps;  // run << 2 | kind, traceback order
    unsigned int const *       cigarOff;  // per slot of `order`: first op of the alignment
    unsigned int               cigarBase; // added to the offsets stored in the records
};

__global__ void __launch_bounds__(128) tracebackDpxKernel(TracebackDpxParams P)
{
    unsigned int const t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= P.nTasks)
        return;
    unsigned int const       task = P.order[t];
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
    unsigned int const       K    = P.kOf[task], KN = (K + 1) / 2;
    unsigned int const       nSteps = nt + 63;
    unsigned int const *     planeH = P.planes + P.planeOff[task];
    unsigned int const *     planeN = planeH + static_cast<unsigned long long>(nSteps) * 32ull * K;
    int const                D      = P.ge - P.go;

    // H(i,j) and the nibbles of cell (i,j), 1-based, 1 <= i <= nq, 1 <= j <= nt
    auto cellH = [&](unsigned int i, unsigned int j) -> int {
        unsigned int const v = (i - 1) / K, r = (i - 1) % K, half = v >> 5, p = v & 31u;
        unsigned int const s = (j - 1) + v;
        unsigned int const w = planeH[(static_cast<unsigned long long>(s) * 32u + p) * K + r];
        int const          W = half ? (static_cast<int>(w) >> 16) : static_cast<int>(static_cast<short>(w & 0xffffu));
        return W - P.go;
    };
    auto cellN = [&](unsigned int i, unsigned int j) -> unsigned int {
        unsigned int const v = (i - 1) / K, r = (i - 1) % K, half = v >> 5, p = v & 31u;
        unsigned int const s = (j - 1) + v;
        unsigned int const w = planeN[(static_cast<unsigned long long>(s) * 32u + p) * KN + (r >> 1)];
        return (w >> (8u * ((r & 1u) + 2u * half))) & 0xffu; // dE | dF << 4
    };
    // SeqAn's trace byte of cell (ii, jj); 0 on the matrix border
    auto tr = [&](unsigned int ii, unsigned int jj) -> unsigned int {
        if (ii == 0 || jj == 0)
            return 0u;
        int const H = cellH(ii, jj);
        if (H <= 0)
            return 0u;
        unsigned int const n   = cellN(ii, jj);
        int const          dE = static_cast<int>(n & 15u), dF = static_cast<int>(n >> 4);
        int const          dFl = (ii > 1) ? static_cast<int>(cellN(ii - 1, jj) >> 4) : 15;
        int const          dEu = (jj > 1) ? static_cast<int>(cellN(ii, jj - 1) & 15u) : 15;
        unsigned int       tv = (dFl <= D ? T_HORI : 0u) | (dFl >= D ? T_HOPEN : 0u) | (dEu <= D ? T_VERT : 0u) |
                          (dEu >= D ? T_VOPEN : 0u);
        if (min(dE, dF) > 0)
            tv |= T_DIAG;
        else
        {
            tv |= (dE <= dF ? T_MAXV : 0u) | (dF <= dE ? T_MAXH : 0u);
            int const hd = (ii > 1 && jj > 1) ? cellH(ii - 1, jj - 1) : 0;
            if (hd + static_cast<int>(M[qs[ii - 1] * 32 + ts[jj - 1]]) == H)
                tv |= T_DIAG;
        }
        return tv;
    };

    int const    score = P.scores[task];
    unsigned int i = (score > 0) ? P.bestCol[task] : 0u, j = 0; // no positive cell: empty alignment at (0, 0)
    if (score > 0)
        for (unsigned int jj = 1; jj <= nt; ++jj) // first row of the best column that holds the best score
            if (cellH(i, jj) == score)
            {
                j = jj;
                break;
            }
    unsigned int const bi = i, bj = j;
    unsigned int nMatch = 0, nMismatch = 0, nPositive = 0, nGapOpen = 0, nGapExt = 0, alnLen = 0;
    unsigned int nOps = 0;
    unsigned int * const ops = P.emit ? P.cigarOps + P.cigarOff[t] : nullptr;

    if (score > 0 && j > 0)
    {
        unsigned int tv = tr(i, j);
        int          last; // 0 diag, 1 horizontal, 2 vertical
        if (tv & T_MAXV) { tv &= (T_VERT | T_VOPEN | T_MAXV); last = 2; }
        else if (tv & T_MAXH) { tv &= (T_HORI | T_HOPEN | T_MAXH); last = 1; }
        else last = 0;
        unsigned int run = 0;
        auto flush = [&]() {
            if (run)
            {
                if (ops)
                    ops[nOps++] = (run << 2) | static_cast<unsigned int>(last);
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
        while (i > 0 && j > 0 && tv != 0)
        {
            if (tv & T_DIAG)
            {
                switchTo(0);
                unsigned int const a = qs[i - 1], b = ts[j - 1];
                if (alignedIdentical(P.ix, M, a, b)) ++nMatch; else ++nMismatch;
                if (M[a * 32 + b] > 0) ++nPositive;
                --i; --j; tv = tr(i, j); ++run;
            }
            else if ((tv & T_MAXV) && (tv & T_VERT))
            {
                switchTo(2);
                while ((!(tv & T_VOPEN) || (tv & T_VERT)) && j != 1)
                {
                    --j; tv = tr(i, j); ++run;
                }
                --j; tv = tr(i, j); ++run;
            }
            else if ((tv & T_MAXV) && (tv & T_VOPEN))
            {
                switchTo(2);
                --j; tv = tr(i, j); ++run;
            }
            else if ((tv & T_MAXH) && (tv & T_HORI))
            {
                switchTo(1);
                while ((!(tv & T_HOPEN) || (tv & T_HORI)) && i != 1)
                {
                    --i; tv = tr(i, j); ++run;
                }
                --i; tv = tr(i, j); ++run;
            }
            else if ((tv & T_MAXH) && (tv & T_HOPEN))
            {
                switchTo(1);
                --i; tv = tr(i, j); ++run;
            }
            else
                break;
        }
        flush();
    }

    if (P.emit)
    {
        P.out[task].cigar_off = P.cigarBase + P.cigarOff[t];
        P.out[task].cigar_len = nOps;
        return;
    }
    lgpu_hit h;
    h.q_id       = q;
    h.s_id       = sId;
    h.q_start    = m.qry_start + i;
    h.q_end      = m.qry_start + bi;
    h.s_start    = m.subj_start + j;
    h.s_end      = m.subj_start + bj;
    h.q_len      = qLen;
    h.s_len      = static_cast<unsigned int>(P.ix.origDelims[sId + 1] - P.ix.origDelims[sId]);
    h.score      = score;
    h.n_match    = nMatch;
    h.n_mismatch = nMismatch;
    h.n_gap_open = nGapOpen;
    h.n_gap_ext  = nGapExt;
    h.n_positive = nPositive;
    h.aln_len    = alnLen;
    setFrames(P.Q, P.ix, m.qry_id, m.subj_id, h.q_frame, h.s_frame);
    h.phase      = 0;
    h.reserved   = 0;
    h.bit_score  = 0.0;
    h.evalue     = 0.0;
    h.cigar_off  = 0;
    h.cigar_len  = 0;
    P.out[task]  = h;
}

// classes of the packed trace kernel: T = 32, K columns per strip, 64 * K >= query length
constexpr int kNumTraceClasses = 12;
__host__ __device__ inline int dpxTraceK(int cls)
{
    constexpr int ks[kNumTraceClasses] = {1, 2, 3, 4, 5, 6, 8, 10, 12, 16, 24, 32};
    return ks[cls];
}
__host__ __device__ inline int dpxTraceClassOf(unsigned int nq)
{
    for (int c = 0; c < kNumTraceClasses; ++c)
        if (nq <= 64u * static_cast<unsigned int>(dpxTraceK(c)))
            return c;
    return kNumTraceClasses; // scalar wavefront kernel
}

} // namespace lgpu
