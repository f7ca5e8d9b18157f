// DP pass 1 (score only) on Blackwell's packed-int16 DPX instructions.
//
// Same recurrence as the reference's SIMD pass (_performAlignment<withTrace=false>,
// src/search_algo.hpp:1246; SQ/align/dp_formula_affine.h:66-126), restated so that one cell update is
// four half-rate DPX instructions + one full-rate VIMNMX + half a VIMNMX3 on two int16 lanes, one PRMT that builds the
// operand and one plain 32-bit add that runs on the other (FMA/IMAD) pipe:
//
//   E^ = E - go, F^ = F - go                 (the gap states are kept shifted; E^, F^ >= 0)
//   y   = Hdiag + sub'                       IADD (32 bit)          sub' = M[q][s] - go  (one profile byte, >= 0)
//   t   = max(y, E^)                         VIMNMX.S16x2 (full rate)
//   Ht  = max(t + go, 0)                     VIADDMNMX.S16x2.RELU   = max(Hdiag + sub, E, 0)
//   H   = max(F^ + go, Ht)                   VIADDMNMX.S16x2
//   F^' = max(F^ + ge, Ht)                   VIADDMNMX.S16x2        (= max(F^ + ge, H) because go <= ge;
//                                                                    the only op on the row's dependency chain)
//   E^' = max(E^ + ge, H)                    VIADDMNMX.S16x2
//   best = max(best, H, H of the next cell)  VIMNMX3.S16x2 every other cell
//
// The 32-bit add is exact on both halves because H >= 0 and sub' >= 0 (checked on the host: every matrix
// entry >= go), so the low half never carries.  Negative gap states never matter (H is clamped at 0 and
// E', F' are then dominated by H + go), so starting E^ = F^ = 0 instead of -inf changes nothing.
//
// Work decomposition: a group of T threads (8/16/32) owns one alignment.  The query is cut into 2T
// strips of K columns; thread p holds strip p in the low int16 half and strip p + T in the high half
// of every register, so both halves are always busy on different cells of the SAME alignment (no
// pairing of alignments, no length mismatch).  Strip v works on subject row j = s - v at step s: the
// 2T strips form an anti-diagonal wavefront, and the only communication per step is the rotation of
// (H, F^) of a strip's last column to the next strip (two shuffles).  Rows outside the window and
// columns past the query end use a "null" profile entry sub' = 0 (a substitution score of go): such a
// cell is max(Hdiag + go, E, F, 0), which can never exceed a value that already exists and leaves the
// all-zero state in front of the window untouched, so no masking is needed anywhere in the inner loop
// and all groups of a warp can simply run for the longest window among them.
//
// Shared memory per group: the query profile P[code][word][strip] (int8, row stride a multiple of 32
// words so the 4-byte loads of a warp are bank-conflict free) and the padded subject window.
#pragma once

#include <cstdint>
#include <cuda_runtime.h>

#include "../../include/lambda_b200.h"
#include "kernels_fm.cuh"

namespace lgpu
{

struct DpxParams
{
    DevIndex             ix;
    DevQueries           Q;
    lgpu_match const *         tasks;
    unsigned int const *       order;   // task indices sorted by (class, query, window length)
    unsigned long long const * keys;    // the sort keys belonging to `order`
    unsigned int               nSorted; // entries in order / keys
    unsigned int const *       jobs;    // first sorted slot of every job of this class
    unsigned int               nJobs;
    signed char const *  matrix; // 2 x (32 x 32)
    int                  go, ge;
    unsigned int         nCodes; // alphabet size + 1 (last row = null)
    unsigned int         winCap; // bytes reserved per group for the padded window
    unsigned int *       workCounter;
    int *                scores;
};

__device__ __forceinline__ unsigned int prmt(unsigned int a, unsigned int b, unsigned int sel)
{
    unsigned int d;
    asm("prmt.b32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(sel));
    return d;
}

// words per profile row: ((K+3)/4) * 2T rounded up to a multiple of 32, plus 8, so that rows of
// different residue codes start 8 banks apart (the groups of a warp read different rows at once)
__host__ __device__ constexpr int dpxRowWords(int T, int K)
{
    return (((K + 3) / 4) * 2 * T + 31) / 32 * 32 + ((T == 32) ? 0 : 8);
}

// sort key = class << 58 | (qryId << 1 | matrix selector) << 20 | min(nt, 2^20 - 1); the selector is the
// subject parity in bisulfite mode (reverse matrix for odd subjects), else 0: all alignments of a job
// share one profile
constexpr unsigned int kDpxSegShift = 20;

// Inner-loop formulation (compile-time experiment switch; see the header comment):
//   0  W = H + go stored, 6.5 ALU-pipe instructions per cell pair, profile null = -128
//   1  shifted gap states, VIMNMX3 on the row's dependency chain (3 ops deep), 5.5 per cell pair
//   2  shifted gap states, one op on the dependency chain, 6.0 per cell pair
#ifndef LGPU_DPX_FORM
#define LGPU_DPX_FORM 0
#endif
#if LGPU_DPX_FORM == 0
constexpr unsigned int kDpxNullWord = 0x80808080u;
constexpr int          kDpxNullVal  = -128;
#else
constexpr unsigned int kDpxNullWord = 0u; // null: sub' = 0
constexpr int          kDpxNullVal  = 0;
#endif

// A job = up to G = 32/T consecutive (sorted) alignments of the SAME query: the warp builds the query
// profile once in shared memory and its G groups run one alignment each against it.
template <int T, int K>
__global__ void __launch_bounds__(32) swScoreDpxKernel(DpxParams P)
{
    constexpr int KW   = (K + 3) / 4;  // profile words per strip
    constexpr int ROWW = dpxRowWords(T, K);
    constexpr int PAD  = 2 * T;        // null rows in front of the window

    extern __shared__ unsigned int smem[];
    unsigned int const lane = threadIdx.x;
    unsigned int const grp  = lane / T;
    unsigned int const gl   = lane % T;
    unsigned int const profWords = P.nCodes * ROWW;
    unsigned int *     prof = smem;
    // window buffers of the groups start 8 banks apart as well (winCap is a multiple of 128 bytes)
    unsigned char *    win  = reinterpret_cast<unsigned char *>(smem + profWords) + grp * (P.winCap + 32);
    unsigned int const nullCode = P.nCodes - 1;

    unsigned int const go2  = (static_cast<unsigned int>(P.go) & 0xffffu) * 0x10001u;
    unsigned int const ge2  = (static_cast<unsigned int>(P.ge) & 0xffffu) * 0x10001u;

    for (;;)
    {
        unsigned int job = 0;
        if (lane == 0)
            job = atomicAdd(P.workCounter, 1u);
        job = __shfl_sync(0xffffffffu, job, 0);
        if (job >= P.nJobs)
            break;
        unsigned int const       slot0 = P.jobs[job];
        unsigned long long const seg   = P.keys[slot0] >> kDpxSegShift;
        unsigned int const       slot  = slot0 + grp;
        // all alignments of a job share the query (frame) and the query range: take them from the first
        lgpu_match const         m0   = P.tasks[P.order[slot0]];
        bool                     valid = slot < P.nSorted && (P.keys[slot] >> kDpxSegShift) == seg;
        if (valid)
        {
            lgpu_match const m = P.tasks[P.order[slot]];
            valid              = m.qry_start == m0.qry_start && m.qry_end == m0.qry_end;
        }
        {
            // a job ends at the first slot that differs (same rule as segFlagKernel)
            unsigned int const eq = __ballot_sync(0xffffffffu, valid);
            for (unsigned int g2 = 0; g2 < grp; ++g2)
                valid = valid && ((eq >> (g2 * T)) & 1u);
        }
        unsigned int const       q    = m0.qry_id / P.Q.F;
        unsigned int const       f    = m0.qry_id % P.Q.F;
        unsigned long long const qb   = P.Q.offs[q];
        unsigned int const       qLen = static_cast<unsigned int>(P.Q.offs[q + 1] - qb);
        unsigned char const *    qs   = P.Q.trans + P.Q.F * qb + static_cast<unsigned long long>(f) * qLen + m0.qry_start;
        unsigned int const       nq   = m0.qry_end - m0.qry_start;
        signed char const *      M    = P.matrix + matrixOffset(P.ix, m0.subj_id);
        unsigned int             task = 0, nt = 0;
        unsigned char const *    ts   = nullptr;
        if (valid)
        {
            task               = P.order[slot];
            lgpu_match const m = P.tasks[task];
            ts                 = P.ix.seqs + sbjBase(P.ix, m.subj_id) + m.subj_start;
            nt                 = m.subj_end - m.subj_start;
        }
        // the warp runs for its longest window; extra steps are null rows for the shorter ones
        unsigned int ntMax = nt;
#pragma unroll
        for (int off = 16; off > 0; off >>= 1)
            ntMax = max(ntMax, __shfl_xor_sync(0xffffffffu, ntMax, off));
        unsigned int const nSteps = ntMax + 2 * T - 1;

        __syncwarp();
        // ---- query profile: P[c][w][v], byte r%4 of word w = r/4 of strip v <-> column i = v*K + r ----
        for (unsigned int idx = lane; idx < profWords; idx += 32)
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

#if LGPU_DPX_FORM == 0
        unsigned int const neg2 = 0xC000C000u; // -16384 in both halves
        unsigned int const init = go2, initE = neg2, border = go2, borderF = neg2;
#else
        unsigned int const init = 0u, initE = 0u, border = 0u, borderF = 0u;
#endif
        unsigned int E[K], H[K]; // form 0: H holds W = H + go
#pragma unroll
        for (int r = 0; r < K; ++r)
        {
            E[r] = initE;
            H[r] = init;
        }
        unsigned int best = init;
        unsigned int outH = init, outF = initE, diagIn = init;

        // profile words of the current step (software pipelined one step ahead)
        unsigned int wl[KW], wh[KW];
        {
            unsigned int const cLo = win[PAD - gl];           // row 0 - gl  (null for gl > 0)
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
            // (H, F^) of the left strip's last column for the row this strip works on now
            unsigned int inH = __shfl_sync(0xffffffffu, outH, (lane - 1) & (T - 1), T);
            unsigned int inF = __shfl_sync(0xffffffffu, outF, (lane - 1) & (T - 1), T);
            if (gl == 0)
            {
                // strip 0 sees the matrix border (H = 0, no horizontal gap); strip T continues strip T-1
                inH = prmt(border, inH, 0x5410);
                inF = prmt(borderF, inF, 0x5410);
            }
            unsigned int diag = diagIn; // H of the left strip at the previous row
            diagIn            = inH;
            unsigned int F    = inF;
#pragma unroll
            for (int r = 0; r < K; ++r)
            {
                // {lo byte, hi byte} of column r, each widened to 16 bits (values are 0 .. 127)
                unsigned int const b   = r & 3;
                unsigned int const sel = ((0xCu + b) << 12) | ((4u + b) << 8) | ((8u + b) << 4) | b;
                unsigned int const sub = prmt(wl[r >> 2], wh[r >> 2], sel);
#if LGPU_DPX_FORM == 0
                unsigned int const t   = __viaddmax_s16x2_relu(diag, sub, E[r]);
                unsigned int const u   = __vadd2(t, go2);
                unsigned int const h   = __viaddmax_s16x2(F, go2, u);
                F                      = __viaddmax_s16x2(F, ge2, u);
                E[r]                   = __viaddmax_s16x2(E[r], ge2, h);
#elif LGPU_DPX_FORM == 1
                unsigned int const y   = diag + sub; // both halves non-negative: no carry across the halves
                unsigned int const m   = __vimax3_s16x2(y, E[r], F);
                unsigned int const h   = __viaddmax_s16x2_relu(m, go2, 0u);
                F                      = __viaddmax_s16x2(F, ge2, h);
                E[r]                   = __viaddmax_s16x2(E[r], ge2, h);
#else
                unsigned int const y   = diag + sub; // both halves non-negative: no carry across the halves
                unsigned int const t   = __vmaxs2(y, E[r]);
                unsigned int const ht  = __viaddmax_s16x2_relu(t, go2, 0u); // max(Hdiag + sub, E, 0)
                unsigned int const h   = __viaddmax_s16x2(F, go2, ht);      // ... and F
                F                      = __viaddmax_s16x2(F, ge2, ht);      // = max(F^ + ge, H) because go <= ge
                E[r]                   = __viaddmax_s16x2(E[r], ge2, h);
#endif
                diag                   = H[r];
                H[r]                   = h;
                best                   = __vmaxs2(best, h);
            }
            outH = H[K - 1];
            outF = F;
#pragma unroll
            for (int k = 0; k < KW; ++k)
            {
                wl[k] = nl[k];
                wh[k] = nh[k];
            }
        }
        // reduce over the group: both halves, all T threads
        int b = max(static_cast<int>(static_cast<short>(best & 0xffffu)), static_cast<int>(best) >> 16);
#pragma unroll
        for (int off = T / 2; off > 0; off >>= 1)
            b = max(b, __shfl_xor_sync(0xffffffffu, b, off));
        if (valid && gl == 0)
            P.scores[task] = b - (LGPU_DPX_FORM == 0 ? P.go : 0);
    }
}

// ---------------------------------------------------------------------------------------------
// classification of tasks into (T, K) classes
// ---------------------------------------------------------------------------------------------

// (T, K) classes of the packed kernel, ascending by the columns they cover (2 * T * K).  Fine steps
// (16 columns) where protein lengths concentrate, so that a query wastes few padded columns.
#define LGPU_DPX_CLASSES(X)                                                                                        \
    X(8, 4) X(8, 6) X(8, 8) X(8, 9) X(8, 10) X(8, 11) X(8, 12) X(8, 13) X(8, 14) X(8, 15) X(8, 16) X(8, 17)       \
    X(8, 18) X(8, 19) X(8, 20) X(8, 21) X(8, 22) X(8, 23) X(8, 24) X(8, 26) X(8, 28) X(8, 30) X(8, 32)            \
    X(16, 20) X(16, 24) X(16, 28) X(16, 32) X(32, 20) X(32, 24) X(32, 28) X(32, 32)

struct DpxClass
{
    int T, K;
};
#define LGPU_DPX_CLASS_ENTRY(T, K) {T, K},
__host__ __device__ inline DpxClass dpxClass(int cls)
{
    constexpr DpxClass tab[] = {LGPU_DPX_CLASSES(LGPU_DPX_CLASS_ENTRY)};
    return tab[cls];
}
#define LGPU_DPX_CLASS_COUNT(T, K) +1
constexpr int kNumDpxClasses = 0 LGPU_DPX_CLASSES(LGPU_DPX_CLASS_COUNT);
static_assert(kNumDpxClasses < 63, "class id must fit the sort key");
constexpr unsigned int kDpxClassShift = 58; // sort key: class in the top 6 bits

__host__ __device__ inline int dpxClassOf(unsigned int nq)
{
    for (int c = 0; c < kNumDpxClasses; ++c)
    {
        DpxClass const k = dpxClass(c);
        if (nq <= static_cast<unsigned int>(2 * k.T * k.K))
            return c;
    }
    return kNumDpxClasses; // too long: scalar wavefront kernel
}

constexpr unsigned int kDpxMaxWindow = 8192; // longer windows go to the scalar kernel

// alignments per job (= groups per warp) of each class; the scalar class has one alignment per job
__host__ __device__ inline unsigned int dpxGroupsOf(int cls)
{
    return cls < kNumDpxClasses ? static_cast<unsigned int>(32 / dpxClass(cls).T) : 1u;
}

// key (see kDpxSegShift); also per-class counts / max window / total cells
__global__ void classifyKernel(lgpu_match const * tasks, unsigned int n, unsigned int bsMode, unsigned long long * keys, unsigned int * idx,
                               unsigned int * classCount, unsigned int * classMaxNt, unsigned int * maxNq,
                               unsigned long long * cells)
{
    unsigned int const t = blockIdx.x * blockDim.x + threadIdx.x;
    unsigned long long myCells = 0;
    int                c  = -1;
    unsigned int       nq = 0, nt = 0;
    if (t < n)
    {
        nq = tasks[t].qry_end - tasks[t].qry_start;
        nt = tasks[t].subj_end - tasks[t].subj_start;
        c  = dpxClassOf(nq);
        if (nt > kDpxMaxWindow)
            c = kNumDpxClasses;
        unsigned long long const seg = (static_cast<unsigned long long>(tasks[t].qry_id) << 1) | (tasks[t].subj_id & bsMode);
        keys[t] = (static_cast<unsigned long long>(c) << kDpxClassShift) | (seg << kDpxSegShift) |
                  (nt < (1u << kDpxSegShift) ? nt : (1u << kDpxSegShift) - 1u);
        idx[t]  = t;
        myCells = static_cast<unsigned long long>(nq) * nt;
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

// head[t] = 1 iff slot t is the first alignment of a job; counts the jobs per class
__global__ void jobHeadKernel(unsigned long long const * keys, unsigned int const * segStart, unsigned int n, unsigned int * head,
                              unsigned int * classJobs)
{
    unsigned int const t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n)
        return;
    int const          cls = static_cast<int>(keys[t] >> kDpxClassShift);
    unsigned int const h   = ((t - segStart[t]) % dpxGroupsOf(cls)) == 0 ? 1u : 0u;
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
