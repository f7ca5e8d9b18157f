// Extension DP (both passes) on Blackwell's packed-int16 DPX instructions.
//
// Same recurrence as the reference's SIMD passes (_performAlignment<withTrace>, src/search_algo.hpp:1071-1134,
// 1246, 1296; SQ/align/dp_formula_affine.h:66-126), restated so that one cell update on two int16 lanes is
//
//   W = H + go is the only stored form of H; the profile stores sub' = M[q][s] - go  (int8, -128 = null)
//   t   = max(Wdiag + sub', E, 0)            VIADDMNMX.S16x2.RELU   = H' of the cell without F
//   u   = t + go                             VIADD.16x2
//   W   = max(F + go, u)                     VIADDMNMX.S16x2
//   F'  = max(F + ge, u)                     VIADDMNMX.S16x2        (= max(F + ge, W) because go <= ge;
//                                                                    the only op on the row's dependency chain)
//   E'  = max(E + ge, W)                     VIADDMNMX.S16x2
//   best = max(best, W)                      VIMNMX.S16x2 (pairs fuse into VIMNMX3)
//
// plus one PRMT that sign-extends and interleaves the two int8 profile bytes: 6.5 half-rate ALU instructions per
// cell pair in the score pass.
//
// Work decomposition: a group of T threads (8/16/32) owns one alignment.  The query is cut into 2T strips of K
// columns; thread p holds strip p in the low int16 half and strip p + T in the high half of every register, so both
// halves are always busy on different cells of the SAME alignment.  Strip v works on subject row j = s - v at step
// s: the 2T strips form an anti-diagonal wavefront, and the only communication per step is the rotation of (W, F) of
// a strip's last column to the next strip (two shuffles).  Rows outside the window and columns past the query end
// read the "null" profile byte, which can never raise a cell above a value that already exists and leaves the
// all-zero state in front of the window untouched: no masks anywhere in the inner loop, and all groups of a warp
// simply run for the longest window among them.
//
// Who shares a warp (template parameter PRIV):
//   PRIV = false  a job = up to G = 32/T consecutive (sorted) alignments of the SAME query; the warp builds ONE
//                 query profile in shared memory and its G groups run one alignment each against it (protein
//                 searches: tens of candidate subjects per query, 28-row profiles).
//   PRIV = true   a job = G consecutive alignments of the class in window-length order, whatever their query: every
//                 group builds its own profile (nucleotide / bisulfite searches: about one alignment per read and
//                 6-row profiles -- the reference batches alignments of different queries into one vector the same
//                 way, src/search_algo.hpp:1087-1133).
//
// Pass 2 (TRACE = true) runs the same loop and additionally stores ONE byte per cell, W mod 256, in the register
// image of the wavefront (kernels_dpx_trace.cuh explains how the traceback rebuilds SeqAn's trace decisions from the
// residues), and the per-column maxima that locate the end cell.  Extra cost per cell pair: one VIMNMX (column
// maximum), half a PRMT (byte packing) and a quarter of a 32-bit store.
#pragma once

#include <cstdint>
#include <cuda_runtime.h>

#include "../../include/lambda_b200.h"
#include "kernels_fm.cuh"

namespace lgpu
{

struct DpxParams
{
    DevIndex                   ix;
    DevQueries                 Q;
    lgpu_match const *         tasks;
    unsigned int const *       order;    // task indices sorted by class, then (query, window length) or window length
    unsigned long long const * keys;     // the sort keys belonging to `order`
    unsigned int               nSorted;  // entries in order / keys
    unsigned int const *       jobs;     // PRIV = false: first sorted slot of every job of this class
    unsigned int               nJobs;    // jobs of this launch
    unsigned int               slotBase; // PRIV = true: first sorted slot of this class; job b = slots slotBase + b*G ...
    unsigned int               nSlots;   // PRIV = true: alignments of this class
    signed char const *        matrix;   // 2 x (32 x 32)
    int                        go, ge;
    unsigned int               nCodes;   // alphabet size + 1 (last row = null)
    unsigned int               winCap;   // bytes reserved per group for the padded window
    unsigned int *             workCounter;
    int *                      scores;   // out, indexed by task
    // pass 2 only
    unsigned int *             planes;   // residue planes, 32-bit words
    unsigned long long const * planeOff; // word offset of every sorted slot's plane
    unsigned long long         planeOffBase; // ... minus this (planes of one launch group start at `planes`)
    unsigned int *             bestCol;  // out, indexed by task: 1-based column of the end cell
};

__device__ __forceinline__ unsigned int prmt(unsigned int a, unsigned int b, unsigned int sel)
{
    unsigned int d;
    asm("prmt.b32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(sel));
    return d;
}

// words per profile row: ((K+3)/4) * 2T rounded up to a multiple of 32.  Shared profile: plus 8, so that rows of
// different residue codes start 8 banks apart (the groups of a warp read different rows at once).  Private profiles:
// the profiles of the groups start T banks apart instead (dpxProfStride).
__host__ __device__ constexpr int dpxRowWords(int T, int K, bool priv = false)
{
    return (((K + 3) / 4) * 2 * T + 31) / 32 * 32 + ((T == 32 || priv) ? 0 : 8);
}
__host__ __device__ constexpr unsigned int dpxProfStride(int T, int K, unsigned int nCodes)
{
    return nCodes * static_cast<unsigned int>(dpxRowWords(T, K, true)) + (T == 32 ? 0u : static_cast<unsigned int>(T));
}
// dynamic shared memory of one warp
__host__ __device__ constexpr size_t dpxSmemBytes(int T, int K, bool priv, unsigned int nCodes, unsigned int winCap)
{
    return (priv ? static_cast<size_t>(32 / T) * dpxProfStride(T, K, nCodes) : static_cast<size_t>(nCodes) * dpxRowWords(T, K)) * 4 +
           static_cast<size_t>(32 / T) * (winCap + 32);
}
// words of the residue plane of one alignment with `nt` subject rows: (nt + 2T - 1) wavefront steps, 2TK bytes each
__host__ __device__ constexpr unsigned long long dpxPlaneWords(int T, int K, unsigned int nt)
{
    return static_cast<unsigned long long>(nt + 2 * T - 1) * static_cast<unsigned int>((K + 1) / 2) * static_cast<unsigned int>(T);
}

// sort key = class << 58 | segment << 20 | window.  Shared profiles: segment = qryId << 1 | matrix selector (the
// subject parity in bisulfite mode, else 0) and window = min(nt, 2^20 - 1): all alignments of a job share one
// profile.  Private profiles: segment = 0 and window = 2^20 - 1 - min(nt, 2^20 - 1): longest windows first.
constexpr unsigned int kDpxSegShift   = 20;
constexpr unsigned int kDpxClassShift = 58; // class in the top 6 bits
#ifndef LGPU_DPX_BEST2
#define LGPU_DPX_BEST2 0
#endif
constexpr unsigned int kDpxNullWord   = 0x80808080u;
constexpr int          kDpxNullVal    = -128;

template <int T, int K, bool PRIV, bool TRACE>
__global__ void __launch_bounds__(32) swDpxKernel(DpxParams P)
{
    constexpr int G    = 32 / T;
    constexpr int KW   = (K + 3) / 4; // profile words per strip
    constexpr int KN   = (K + 1) / 2; // residue words per strip and step
    constexpr int ROWW = dpxRowWords(T, K, PRIV);
    constexpr int PAD  = 2 * T;       // null rows in front of the window
    static_assert(!TRACE || PRIV, "pass 2 runs with one profile per group");

    extern __shared__ unsigned int smem[];
    unsigned int const lane      = threadIdx.x;
    unsigned int const grp       = lane / T;
    unsigned int const gl        = lane % T;
    unsigned int const profWords = P.nCodes * ROWW;
    unsigned int const profAll   = PRIV ? G * dpxProfStride(T, K, P.nCodes) : profWords;
    unsigned int *     prof      = smem + (PRIV ? grp * dpxProfStride(T, K, P.nCodes) : 0u);
    // window buffers of the groups start 8 banks apart as well (winCap is a multiple of 128 bytes)
    unsigned char *    win       = reinterpret_cast<unsigned char *>(smem + profAll) + grp * (P.winCap + 32);
    unsigned int const nullCode  = P.nCodes - 1;

    unsigned int const go2  = (static_cast<unsigned int>(P.go) & 0xffffu) * 0x10001u;
    unsigned int const ge2  = (static_cast<unsigned int>(P.ge) & 0xffffu) * 0x10001u;
    unsigned int const neg2 = 0xC000C000u; // -16384 in both halves

    for (;;)
    {
        unsigned int job = 0;
        if (lane == 0)
            job = atomicAdd(P.workCounter, 1u);
        job = __shfl_sync(0xffffffffu, job, 0);
        if (job >= P.nJobs)
            break;

        // ---- the group's alignment (none: an all-null profile and an empty window) ----
        bool                  valid = false;
        unsigned int          task = 0, slot = 0, nq = 0, nt = 0;
        unsigned char const * qs = nullptr;
        unsigned char const * ts = nullptr;
        signed char const *   M  = P.matrix;
        if constexpr (PRIV)
        {
            unsigned int const k = job * G + grp;
            valid                = k < P.nSlots;
            slot                 = P.slotBase + k;
        }
        else
        {
            unsigned int const       slot0 = P.jobs[job];
            unsigned long long const seg   = P.keys[slot0] >> kDpxSegShift;
            slot                           = slot0 + grp;
            // all alignments of a job share the query (frame) and the query range: a job ends at the first slot
            // that differs (same rule as segFlagKernel)
            lgpu_match const m0 = P.tasks[P.order[slot0]];
            valid               = slot < P.nSorted && (P.keys[slot] >> kDpxSegShift) == seg;
            if (valid)
            {
                lgpu_match const m = P.tasks[P.order[slot]];
                valid              = m.qry_start == m0.qry_start && m.qry_end == m0.qry_end;
            }
            unsigned int const eq = __ballot_sync(0xffffffffu, valid);
            for (unsigned int g2 = 0; g2 < grp; ++g2)
                valid = valid && ((eq >> (g2 * T)) & 1u);
            unsigned int const       q    = m0.qry_id / P.Q.F;
            unsigned int const       f    = m0.qry_id % P.Q.F;
            unsigned long long const qb   = P.Q.offs[q];
            unsigned int const       qLen = static_cast<unsigned int>(P.Q.offs[q + 1] - qb);
            qs = P.Q.trans + P.Q.F * qb + static_cast<unsigned long long>(f) * qLen + m0.qry_start;
            nq = m0.qry_end - m0.qry_start;
            M  = P.matrix + matrixOffset(P.ix, m0.subj_id);
        }
        if (valid)
        {
            task               = P.order[slot];
            lgpu_match const m = P.tasks[task];
            ts                 = P.ix.seqs + sbjBase(P.ix, m.subj_id) + m.subj_start;
            nt                 = m.subj_end - m.subj_start;
            if constexpr (PRIV)
            {
                unsigned int const       q    = m.qry_id / P.Q.F;
                unsigned int const       f    = m.qry_id % P.Q.F;
                unsigned long long const qb   = P.Q.offs[q];
                unsigned int const       qLen = static_cast<unsigned int>(P.Q.offs[q + 1] - qb);
                qs = P.Q.trans + P.Q.F * qb + static_cast<unsigned long long>(f) * qLen + m.qry_start;
                nq = m.qry_end - m.qry_start;
                M  = P.matrix + matrixOffset(P.ix, m.subj_id);
            }
        }
        // the warp runs for its longest window; extra steps are null rows for the shorter ones
        unsigned int ntMax = nt;
#pragma unroll
        for (int off = 16; off > 0; off >>= 1)
            ntMax = max(ntMax, __shfl_xor_sync(0xffffffffu, ntMax, off));
        unsigned int const nSteps  = ntMax + 2 * T - 1;
        unsigned int const mySteps = valid ? nt + 2 * T - 1 : 0u; // steps whose residues belong to this group's plane

        __syncwarp();
        // ---- query profile: P[c][w][v], byte r%4 of word w = r/4 of strip v <-> column i = v*K + r ----
        for (unsigned int idx = PRIV ? gl : lane; idx < profWords; idx += PRIV ? T : 32)
        {
            unsigned int const c   = idx / ROWW;
            unsigned int const rem = idx % ROWW;
            unsigned int const w   = rem / (2 * T);
            unsigned int const v   = rem % (2 * T);
            unsigned int       word = kDpxNullWord;
            if (c != nullCode && w < KW)
            {
                word = 0;
#pragma unroll
                for (int b = 0; b < 4; ++b)
                {
                    unsigned int const r = w * 4 + b;
                    unsigned int const i = v * K + r;
                    int                val = kDpxNullVal;
                    if (r < K && i < nq)
                        val = static_cast<int>(M[qs[i] * 32 + c]) - P.go;
                    word |= (static_cast<unsigned int>(val) & 0xffu) << (8 * b);
                }
            }
            prof[idx] = word;
        }
        // ---- subject window, null-padded on both sides ----
        for (unsigned int idx = gl; idx < P.winCap; idx += T)
        {
            int const     j = static_cast<int>(idx) - PAD;
            unsigned char c = static_cast<unsigned char>(nullCode);
            if (j >= 0 && j < static_cast<int>(nt))
                c = ts[j];
            win[idx] = c;
        }
        __syncwarp();

        unsigned int E[K], W[K], CB[TRACE ? K : 1];
#pragma unroll
        for (int r = 0; r < K; ++r)
        {
            E[r] = neg2;
            W[r] = go2; // H = 0
            if constexpr (TRACE)
                CB[r] = go2; // per-column maximum of W
        }
        // LGPU_DPX_BEST2 (experiment, off): two running maxima -- a single one fuses into VIMNMX3 on the half-rate DPX
        // pipe, two independent plain VIMNMX.S16x2 can issue on the other pipe.  Measured slower (searchp DP pass 1
        // 38.2 vs 35.2 ms, profiles/r2_dp_best2.json): ptxas gives up the 10-fold unrolled software pipeline.
        unsigned int best = go2, best1 = go2;
        unsigned int outW = go2, outF = neg2, diagIn = go2;
        unsigned int * const plane = TRACE && valid ? P.planes + (P.planeOff[slot] - P.planeOffBase) : nullptr;

        // profile words of the current step (software pipelined one step ahead)
        unsigned int wl[KW], wh[KW];
        {
            unsigned int const cLo = win[PAD - gl]; // row 0 - gl  (null for gl > 0)
            unsigned int const cHi = win[PAD - gl - T];
#pragma unroll
            for (int k = 0; k < KW; ++k)
            {
                wl[k] = prof[cLo * ROWW + k * 2 * T + gl];
                wh[k] = prof[cHi * ROWW + k * 2 * T + T + gl];
            }
        }
        for (unsigned int s = 0; s < nSteps; ++s)
        {
            // prefetch the next step's operands (the window buffer is padded, so s + 1 is always in range)
            unsigned int nl[KW], nh[KW];
            {
                unsigned int const cLo = win[PAD + s + 1 - gl];
                unsigned int const cHi = win[PAD + s + 1 - gl - T];
#pragma unroll
                for (int k = 0; k < KW; ++k)
                {
                    nl[k] = prof[cLo * ROWW + k * 2 * T + gl];
                    nh[k] = prof[cHi * ROWW + k * 2 * T + T + gl];
                }
            }
            // (W, F) of the left strip's last column for the row this strip works on now
            unsigned int inW = __shfl_sync(0xffffffffu, outW, (lane - 1) & (T - 1), T);
            unsigned int inF = __shfl_sync(0xffffffffu, outF, (lane - 1) & (T - 1), T);
            if (gl == 0)
            {
                // strip 0 sees the matrix border (H = 0, no horizontal gap); strip T continues strip T-1
                inW = prmt(go2, inW, 0x5410);
                inF = prmt(neg2, inF, 0x5410);
            }
            unsigned int diag = diagIn; // W of the left strip at the previous row
            diagIn            = inW;
            unsigned int F    = inF;
#pragma unroll
            for (int r = 0; r < K; ++r)
            {
                // {lo byte, hi byte} of column r, each sign-extended to 16 bits
                unsigned int const b   = r & 3;
                unsigned int const sel = ((0xCu + b) << 12) | ((4u + b) << 8) | ((8u + b) << 4) | b;
                unsigned int const sub = prmt(wl[r >> 2], wh[r >> 2], sel);
                unsigned int const t   = __viaddmax_s16x2_relu(diag, sub, E[r]);
                unsigned int const u   = __vadd2(t, go2);
                unsigned int const w   = __viaddmax_s16x2(F, go2, u);
                F                      = __viaddmax_s16x2(F, ge2, u);
                E[r]                   = __viaddmax_s16x2(E[r], ge2, w);
                diag                   = W[r];
                W[r]                   = w;
                if constexpr (TRACE)
                    CB[r] = __vmaxs2(CB[r], w);
                else if (LGPU_DPX_BEST2 && (r & 1))
                    best1 = __vmaxs2(best1, w);
                else
                    best = __vmaxs2(best, w);
            }
            outW = W[K - 1];
            outF = F;
            if constexpr (TRACE)
            {
                // the register image of this step, one byte per cell: word w2 of lane p = columns 2*w2, 2*w2+1 of
                // strip p (bytes 0 and 2) and of strip p + T (bytes 1 and 3), rows s - p and s - p - T
                if (s < mySteps)
                {
                    unsigned int * dst = plane + (static_cast<unsigned long long>(s) * KN) * T + gl;
#pragma unroll
                    for (int w2 = 0; w2 < KN; ++w2)
                        dst[w2 * T] = prmt(W[2 * w2], (2 * w2 + 1 < K) ? W[(2 * w2 + 1 < K) ? 2 * w2 + 1 : 0] : 0u, 0x6420);
                }
            }
#pragma unroll
            for (int k = 0; k < KW; ++k)
            {
                wl[k] = nl[k];
                wh[k] = nh[k];
            }
        }
        if constexpr (TRACE)
        {
            // best score and the smallest column that holds it (padded columns never reach the maximum)
            int          bestW = P.go;
            unsigned int bcol  = 0xffffffffu;
#pragma unroll
            for (int r = 0; r < K; ++r)
            {
                int const          lo = static_cast<int>(static_cast<short>(CB[r] & 0xffffu));
                int const          hi = static_cast<int>(CB[r]) >> 16;
                unsigned int const cl = gl * K + r, ch = (gl + T) * K + r; // 0-based columns
                if (lo > bestW || (lo == bestW && cl < bcol)) { bestW = lo; bcol = cl; }
                if (hi > bestW || (hi == bestW && ch < bcol)) { bestW = hi; bcol = ch; }
            }
#pragma unroll
            for (int off = T / 2; off > 0; off >>= 1)
            {
                int const          ob = __shfl_xor_sync(0xffffffffu, bestW, off);
                unsigned int const oc = __shfl_xor_sync(0xffffffffu, bcol, off);
                if (ob > bestW || (ob == bestW && oc < bcol)) { bestW = ob; bcol = oc; }
            }
            if (valid && gl == 0)
            {
                P.scores[task]  = bestW - P.go;
                P.bestCol[task] = bcol + 1;
            }
        }
        else
        {
            // reduce over the group: both halves, all T threads
            best  = __vmaxs2(best, best1);
            int b = max(static_cast<int>(static_cast<short>(best & 0xffffu)), static_cast<int>(best) >> 16);
#pragma unroll
            for (int off = T / 2; off > 0; off >>= 1)
                b = max(b, __shfl_xor_sync(0xffffffffu, b, off));
            if (valid && gl == 0)
                P.scores[task] = b - P.go;
        }
    }
}

// ---------------------------------------------------------------------------------------------
// classification of tasks into (T, K) classes
// ---------------------------------------------------------------------------------------------

struct DpxClass
{
    int T, K;
};

// Shared-profile classes of the score pass, ascending by the columns they cover (2 * T * K).  Fine steps (16
// columns) where protein lengths concentrate, so that a query wastes few padded columns.
#define LGPU_DPX_CLASSES(X)                                                                                        \
    X(8, 4) X(8, 6) X(8, 8) X(8, 9) X(8, 10) X(8, 11) X(8, 12) X(8, 13) X(8, 14) X(8, 15) X(8, 16) X(8, 17)       \
    X(8, 18) X(8, 19) X(8, 20) X(8, 21) X(8, 22) X(8, 23) X(8, 24) X(8, 26) X(8, 28) X(8, 30) X(8, 32)            \
    X(16, 20) X(16, 24) X(16, 28) X(16, 32) X(32, 20) X(32, 24) X(32, 28) X(32, 32)

// Private-profile classes (both passes of nucleotide / bisulfite searches; read lengths 50 .. 300 fall into the
// 16-column steps of the first row)
#define LGPU_DPX_PRIV_CLASSES(X)                                                                                   \
    X(8, 4) X(8, 5) X(8, 6) X(8, 7) X(8, 8) X(8, 9) X(8, 10) X(8, 12) X(8, 14) X(8, 16) X(8, 20) X(8, 24)          \
    X(8, 28) X(8, 32) X(16, 20) X(16, 24) X(16, 28) X(16, 32) X(32, 20) X(32, 24) X(32, 28) X(32, 32)

// Pass 2 of protein searches: one warp per alignment (survivors of the e-value filter are about one per query, so
// there is no profile to share), 64 strips of K columns.  Tried: two alignments per warp (T = 16, K = 2 .. 32: 31-step
// ramp, twice the cell pairs per step) -- no faster on the 300-aa benchmark (trace stage 7.05 vs 7.03 ms) and slower on
// the real length distribution (2.69 vs 2.26 ms): two 10.8 KB profiles per warp cost more than the shorter ramp saves.
#define LGPU_DPX_TRACE32_CLASSES(X)                                                                                \
    X(32, 1) X(32, 2) X(32, 3) X(32, 4) X(32, 5) X(32, 6) X(32, 8) X(32, 10) X(32, 12) X(32, 16) X(32, 20)         \
    X(32, 24) X(32, 28) X(32, 32)

#define LGPU_DPX_CLASS_ENTRY(T, K) {T, K},
#define LGPU_DPX_CLASS_COUNT(T, K) +1
constexpr int kNumDpxClasses     = 0 LGPU_DPX_CLASSES(LGPU_DPX_CLASS_COUNT);
constexpr int kNumDpxPrivClasses = 0 LGPU_DPX_PRIV_CLASSES(LGPU_DPX_CLASS_COUNT);
constexpr int kNumDpxTr32Classes = 0 LGPU_DPX_TRACE32_CLASSES(LGPU_DPX_CLASS_COUNT);
constexpr int kMaxDpxClasses     = kNumDpxClasses; // the longest table
static_assert(kNumDpxClasses < 63 && kNumDpxPrivClasses <= kMaxDpxClasses && kNumDpxTr32Classes <= kMaxDpxClasses,
              "class id must fit the sort key / the per-class arrays");

// class tables: 0 = shared-profile score classes, 1 = private-profile classes, 2 = pass-2 classes of protein searches
enum { kDpxTabShared = 0, kDpxTabPriv = 1, kDpxTabTrace32 = 2 };
__host__ __device__ inline int dpxNumClasses(int tab)
{
    return tab == kDpxTabShared ? kNumDpxClasses : tab == kDpxTabPriv ? kNumDpxPrivClasses : kNumDpxTr32Classes;
}
__host__ __device__ inline DpxClass dpxClass(int tab, int cls)
{
    constexpr DpxClass t0[] = {LGPU_DPX_CLASSES(LGPU_DPX_CLASS_ENTRY)};
    constexpr DpxClass t1[] = {LGPU_DPX_PRIV_CLASSES(LGPU_DPX_CLASS_ENTRY)};
    constexpr DpxClass t2[] = {LGPU_DPX_TRACE32_CLASSES(LGPU_DPX_CLASS_ENTRY)};
    return tab == kDpxTabShared ? t0[cls] : tab == kDpxTabPriv ? t1[cls] : t2[cls];
}
// smallest class of the table that covers nq columns; dpxNumClasses(tab) = none (scalar wavefront kernel)
__host__ __device__ inline int dpxClassOf(int tab, unsigned int nq)
{
    int const n = dpxNumClasses(tab);
    for (int c = 0; c < n; ++c)
    {
        DpxClass const k = dpxClass(tab, c);
        if (nq <= static_cast<unsigned int>(2 * k.T * k.K))
            return c;
    }
    return n;
}

constexpr unsigned int kDpxMaxWindow = 8192; // longer windows go to the scalar kernel

// alignments per job (= groups per warp) of each class; the scalar class has one alignment per job
__host__ __device__ inline unsigned int dpxGroupsOf(int tab, int cls)
{
    return cls < dpxNumClasses(tab) ? static_cast<unsigned int>(32 / dpxClass(tab, cls).T) : 1u;
}

// Sort keys of the alignments of one pass (see kDpxSegShift) + per-class counts / longest window / total cells.
// tab selects the class table; withPlanes (pass 2) also writes the words of every alignment's residue plane.
__global__ void classifyKernel(lgpu_match const * tasks, unsigned int n, unsigned int bsMode, int tab, unsigned int maxColsInt16,
                               unsigned long long * keys,
                               unsigned int * idx, unsigned int * classCount, unsigned int * classMaxNt, unsigned int * maxNq,
                               unsigned long long * cells, unsigned long long * planeWords)
{
    unsigned int const t = blockIdx.x * blockDim.x + threadIdx.x;
    unsigned long long myCells = 0;
    int                c  = -1;
    unsigned int       nq = 0, nt = 0;
    if (t < n)
    {
        nq = tasks[t].qry_end - tasks[t].qry_start;
        nt = tasks[t].subj_end - tasks[t].subj_start;
        c  = dpxClassOf(tab, nq);
        // the packed kernels compute in int16 lanes: only alignments whose best possible score (columns x largest
        // matrix entry) fits may run there; the scalar kernel computes in 32 bits and reports an overflow
        if (nt > kDpxMaxWindow || nq > maxColsInt16)
            c = dpxNumClasses(tab);
        unsigned int const ntc = nt < (1u << kDpxSegShift) ? nt : (1u << kDpxSegShift) - 1u;
        if (tab == kDpxTabShared)
        {
            unsigned long long const seg = (static_cast<unsigned long long>(tasks[t].qry_id) << 1) | (tasks[t].subj_id & bsMode);
            keys[t] = (static_cast<unsigned long long>(c) << kDpxClassShift) | (seg << kDpxSegShift) | ntc;
        }
        else
            keys[t] = (static_cast<unsigned long long>(c) << kDpxClassShift) | ((1u << kDpxSegShift) - 1u - ntc);
        idx[t]  = t;
        myCells = static_cast<unsigned long long>(nq) * nt;
        if (planeWords)
            planeWords[t] = c < dpxNumClasses(tab) ? dpxPlaneWords(dpxClass(tab, c).T, dpxClass(tab, c).K, nt) : 0ull;
    }
    // one atomic per warp instead of one per alignment when the whole warp is of one class (the usual
    // case: alignments arrive sorted by query); otherwise every lane reports for itself
    {
        unsigned int const validMask = __ballot_sync(0xffffffffu, c >= 0);
        unsigned int const peers     = __match_any_sync(0xffffffffu, c);
        bool const         uniform   = __all_sync(0xffffffffu, c < 0 || peers == validMask);
        if (uniform)
        {
            unsigned int mt = nt, mq = nq;
#pragma unroll
            for (int off = 16; off > 0; off >>= 1)
            {
                mt = max(mt, __shfl_xor_sync(0xffffffffu, mt, off));
                mq = max(mq, __shfl_xor_sync(0xffffffffu, mq, off));
            }
            if (validMask && (threadIdx.x & 31u) == static_cast<unsigned int>(__ffs(validMask) - 1))
            {
                atomicAdd(&classCount[c], static_cast<unsigned int>(__popc(validMask)));
                atomicMax(&classMaxNt[c], mt);
                atomicMax(maxNq, mq);
            }
        }
        else if (c >= 0)
        {
            atomicAdd(&classCount[c], 1u);
            atomicMax(&classMaxNt[c], nt);
            atomicMax(maxNq, nq);
        }
    }
    // block-level reduction of the cell count
    __shared__ unsigned long long sh[8];
#pragma unroll
    for (int off = 16; off > 0; off >>= 1)
        myCells += __shfl_down_sync(0xffffffffu, myCells, off);
    if ((threadIdx.x & 31) == 0)
        sh[threadIdx.x >> 5] = myCells;
    __syncthreads();
    if (threadIdx.x == 0)
    {
        unsigned long long s = 0;
        for (unsigned int w = 0; w < blockDim.x / 32; ++w)
            s += sh[w];
        if (s)
            atomicAdd(cells, s);
    }
}

// dst[t] = src[order[t]]  (plane sizes in sorted order, scanned into plane offsets)
__global__ void gatherU64Kernel(unsigned long long const * src, unsigned int const * order, unsigned int n, unsigned long long * dst)
{
    unsigned int const t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t < n)
        dst[t] = src[order[t]];
}

// segStart[t] = t if sorted slot t opens a new (class, query) segment else 0  (max-scanned afterwards)
__global__ void segFlagKernel(unsigned long long const * keys, unsigned int const * order, lgpu_match const * tasks,
                              unsigned int n, unsigned int * segStart)
{
    unsigned int const t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n)
        return;
    bool head = (t == 0) || (keys[t] >> kDpxSegShift) != (keys[t - 1] >> kDpxSegShift);
    if (!head)
    {
        // the search path always aligns whole queries; the stage API may pass partial query ranges
        lgpu_match const a = tasks[order[t]], b = tasks[order[t - 1]];
        head               = a.qry_start != b.qry_start || a.qry_end != b.qry_end;
    }
    segStart[t] = head ? t : 0u;
}

// head[t] = 1 iff slot t is the first alignment of a job; counts the jobs per class (shared-profile table)
__global__ void jobHeadKernel(unsigned long long const * keys, unsigned int const * segStart, unsigned int n, unsigned int * head,
                              unsigned int * classJobs)
{
    unsigned int const t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n)
        return;
    int const          cls = static_cast<int>(keys[t] >> kDpxClassShift);
    unsigned int const h   = ((t - segStart[t]) % dpxGroupsOf(kDpxTabShared, cls)) == 0 ? 1u : 0u;
    head[t]                = h;
    if (h)
    {
        // aggregate the lanes of one class into a single atomic
        unsigned int const peers = __match_any_sync(__activemask(), cls);
        if ((threadIdx.x & 31u) == static_cast<unsigned int>(__ffs(peers) - 1))
            atomicAdd(&classJobs[cls], static_cast<unsigned int>(__popc(peers)));
    }
}

__global__ void jobEmitKernel(unsigned int const * head, unsigned int const * posIncl, unsigned int n, unsigned int * jobs)
{
    unsigned int const t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t < n && head[t])
        jobs[posIncl[t] - 1] = t;
}

} // namespace lgpu
