// Widen / merge and the extension DP kernels (sm_100a).
//
// Device restatement of
//   _widenMatch / _widenAndPreprocessMatches   reference src/search_algo.hpp:920-938,1137-1175
//   local affine DP (Smith-Waterman-Gotoh)      SQ/align/dp_formula_affine.h:66-126,
//                                               dp_formula.h:136-243 (trace bits, CompleteTrace ties)
//   first strict maximum, column-major          SQ/align/dp_scout_simd.h:216-229,565-578
//   traceback (GapsLeft, affine)                SQ/align/dp_traceback_impl.h:223-258,302-337,379-474
//   alignment statistics                        SQ/align/evaluate_alignment.h:215-300
//
// DP orientation follows the reference: query = columns i (outer), subject window = rows j (inner).
//   hgap(i,j) = max(hgap(i-1,j) + ge, S(i-1,j) + go)      "horizontal": gap in the subject row
//   vgap(i,j) = max(vgap(i,j-1) + ge, S(i,j-1) + go)      "vertical":   gap in the query row
//   S(i,j)    = max(S(i-1,j-1) + M[q_i][t_j], vgap, hgap); S <= 0 -> S = 0, trace = 0
// with go = gapOpen + gapExtend, ge = gapExtend (src/search_algo.hpp:226-227).
//
// Wavefront kernel: one warp per alignment.  Lane p owns K consecutive query columns and walks the
// subject rows; at step s it is on row j = s - p, so the 32 lanes form an anti-diagonal and the only
// communication is one shuffle per step (S and hgap of the lane's last column, packed in 32 bits).
// Queries longer than 32*K columns are processed in column blocks; the right edge of a block is kept
// in a per-warp scratch row in global memory (L2 resident).
#pragma once

#include <cstdint>
#include <cuda_runtime.h>

#include "../../include/lambda_b200.h"
#include "kernels_fm.cuh"

namespace lgpu
{

// ---------------------------------------------------------------------------------------------
// widen + sort keys; merge of overlapping windows
// ---------------------------------------------------------------------------------------------

// integer floor(sqrt(x)) identical to static_cast<int64_t>(std::sqrt(double(x))) for x < 2^52
__device__ __forceinline__ unsigned int isqrtFloor(unsigned int x)
{
    unsigned int r = static_cast<unsigned int>(sqrt(static_cast<double>(x)));
    while (static_cast<unsigned long long>(r) * r > x)
        --r;
    while (static_cast<unsigned long long>(r + 1) * (r + 1) <= x)
        ++r;
    return r;
}

// key1 = (qryId << 32) | subjId ; key2 = (subjStart << 32) | subjEnd   (qryStart/qryEnd are constant
// per qryId after widening, so the reference's 6-field lexicographic order reduces to these two keys)
__global__ void widenKernel(lgpu_match const * in, unsigned long long n, DevQueries Q, DevIndex ix, unsigned int windowBand,
                            unsigned long long * key1, unsigned long long * key2)
{
    unsigned long long const t = blockIdx.x * static_cast<unsigned long long>(blockDim.x) + threadIdx.x;
    if (t >= n)
        return;
    lgpu_match const         m    = in[t];
    unsigned int const       q    = m.qry_id / Q.F;
    // lengths in translated space (src/search_algo.hpp:925-926)
    unsigned long long const qLen = qryFrameLen(Q, static_cast<unsigned int>(Q.offs[q + 1] - Q.offs[q]), m.qry_id % Q.F);
    unsigned long long const sLen = sbjLength(ix, m.subj_id);
    unsigned long long const s0   = (m.subj_start < m.qry_start) ? 0ull : static_cast<unsigned long long>(m.subj_start) - m.qry_start;
    // _bandSize (src/search_misc.hpp:46-50) unless lgpu_params.window_band overrides it
    unsigned long long const band = windowBand ? static_cast<unsigned long long>(windowBand)
                                               : static_cast<unsigned long long>(isqrtFloor(static_cast<unsigned int>(qLen))) + 1ull;
    unsigned long long       e    = s0 + qLen + band;
    if (e > sLen)
        e = sLen;
    unsigned long long const b = (band < s0) ? s0 - band : 0ull;
    key1[t]                    = (static_cast<unsigned long long>(m.qry_id) << 32) | m.subj_id;
    key2[t]                    = (b << 32) | e;
}

__global__ void gatherKernel(unsigned long long const * src, unsigned int const * perm, unsigned long long n,
                             unsigned long long * dst)
{
    unsigned long long const t = blockIdx.x * static_cast<unsigned long long>(blockDim.x) + threadIdx.x;
    if (t < n)
        dst[t] = src[perm[t]];
}

__global__ void iotaKernel(unsigned int * p, unsigned long long n)
{
    unsigned long long const t = blockIdx.x * static_cast<unsigned long long>(blockDim.x) + threadIdx.x;
    if (t < n)
        p[t] = static_cast<unsigned int>(t);
}

__global__ void iotaFromKernel(unsigned int * p, unsigned int n, unsigned int first)
{
    unsigned int const t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t < n)
        p[t] = first + t;
}

// After sorting, windows of one (qry, subj) pair form chains: element t continues the chain of t-1
// iff same pair and end[t-1] >= start[t] (the reference's forward pass compares the untouched
// end of the left element with the untouched start of the right one).  The reference's forward +
// backward passes + unique leave exactly one window per chain: [start of first, end of last]
// (window ends are non-decreasing in sort order).  head[t] = 1 marks chain starts.
__global__ void chainHeadKernel(unsigned long long const * key1, unsigned long long const * key2, unsigned long long n,
                                unsigned int * head)
{
    unsigned long long const t = blockIdx.x * static_cast<unsigned long long>(blockDim.x) + threadIdx.x;
    if (t >= n)
        return;
    unsigned int h = 1;
    if (t > 0 && key1[t] == key1[t - 1])
    {
        unsigned long long const prevEnd = key2[t - 1] & 0xffffffffull;
        unsigned long long const start   = key2[t] >> 32;
        if (prevEnd >= start)
            h = 0;
    }
    head[t] = h;
}

// chainId = inclusive scan of head - 1.  The head writes qry/subj/start, the tail writes end.
__global__ void chainEmitKernel(unsigned long long const * key1, unsigned long long const * key2, unsigned int const * head,
                                unsigned int const * chainIdIncl, unsigned long long n, DevQueries Q, lgpu_match * out)
{
    unsigned long long const t = blockIdx.x * static_cast<unsigned long long>(blockDim.x) + threadIdx.x;
    if (t >= n)
        return;
    unsigned int const c = chainIdIncl[t] - 1;
    if (head[t])
    {
        unsigned int const qryId = static_cast<unsigned int>(key1[t] >> 32);
        unsigned int const q     = qryId / Q.F;
        out[c].qry_id            = qryId;
        out[c].subj_id           = static_cast<unsigned int>(key1[t]);
        out[c].qry_start         = 0;
        out[c].qry_end           = qryFrameLen(Q, static_cast<unsigned int>(Q.offs[q + 1] - Q.offs[q]), qryId % Q.F);
        out[c].subj_start        = static_cast<unsigned int>(key2[t] >> 32);
    }
    if (t + 1 == n || head[t + 1])
        out[c].subj_end = static_cast<unsigned int>(key2[t]);
}

// ---------------------------------------------------------------------------------------------
// extension DP
// ---------------------------------------------------------------------------------------------

enum : unsigned int
{
    T_DIAG  = 1,
    T_HORI  = 2,
    T_VERT  = 4,
    T_HOPEN = 8,
    T_VOPEN = 16,
    T_MAXH  = 32,
    T_MAXV  = 64
};

struct ExtParams
{
    DevIndex              ix;
    DevQueries            Q;
    lgpu_match const *    tasks;
    unsigned int const *  order;        // optional indirection: work item t is task order[t]
    unsigned int          nTasks;       // number of work items
    signed char const *   matrix; // 2 x (32 x 32)
    int                   go, ge;
    unsigned int *        workCounter;  // dynamic task scheduler
    int *                 scores;       // out: best score per task
    unsigned int *        bestPos;      // out (trace): 2 x u32 per task: end column i, end row j (1-based)
    unsigned char *       trace;        // trace bytes, task t at traceOff[t], row-major [j-1][i-1], row stride traceStride[t]
    unsigned long long const * traceOff;
    unsigned int *        boundary;     // per-warp scratch: maxRows packed (S | hgap << 16)
    unsigned int          maxRows;
    unsigned int *        overflowFlag; // set when a score leaves the reference's int16 range (src/search_algo.hpp:1047,1087)
};

constexpr int kNegInf = -16384; // INT16_MIN / 2, SQ/align/dp_cell.h:144-146

// "identical" column of computeAlignmentStats: equal residues (SQ/align/evaluate_alignment.h:270-277);
// bisulfite: score(q, s) == score(q, q) under the direction's matrix (src/evaluate_bisulfite_alignment.hpp:97)
__device__ __forceinline__ bool alignedIdentical(DevIndex const & ix, signed char const * M, unsigned int a, unsigned int b)
{
    return ix.bsMode ? (M[a * 32 + b] == M[a * 32 + a]) : (a == b);
}

__device__ __forceinline__ unsigned int packSH(int s, int h)
{
    return (static_cast<unsigned int>(s) & 0xffffu) | (static_cast<unsigned int>(h) << 16);
}
__device__ __forceinline__ int unpackLo(unsigned int v)
{
    return static_cast<int>(static_cast<short>(v & 0xffffu));
}
__device__ __forceinline__ int unpackHi(unsigned int v)
{
    return static_cast<int>(v) >> 16;
}

template <int K, bool TRACE>
__global__ void __launch_bounds__(128) swWavefrontKernel(ExtParams P)
{
    __shared__ signed char sMM[2048];
    for (int i = threadIdx.x; i < 2048; i += blockDim.x)
        sMM[i] = P.matrix[i];
    __syncthreads();

    unsigned int const lane   = threadIdx.x & 31u;
    unsigned int const warpId = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    unsigned int *     bnd    = P.boundary + static_cast<unsigned long long>(warpId) * P.maxRows;
    int const          go = P.go, ge = P.ge;
    constexpr int      kColBits = 9; // 32 * K <= 512 columns per block
    static_assert(32 * K <= (1 << kColBits), "column block too wide for the packed maximum key");

    for (;;)
    {
        unsigned int task = 0;
        if (lane == 0)
            task = atomicAdd(P.workCounter, 1u);
        task = __shfl_sync(0xffffffffu, task, 0);
        if (task >= P.nTasks)
            break;
        if (P.order)
            task = P.order[task];

        lgpu_match const         m    = P.tasks[task];
        unsigned int const       q    = m.qry_id / P.Q.F;
        unsigned int const       f    = m.qry_id % P.Q.F;
        unsigned long long const qb   = P.Q.offs[q];
        unsigned int const       qLen = static_cast<unsigned int>(P.Q.offs[q + 1] - qb);
        unsigned char const *    qs   = P.Q.trans + P.Q.F * qb + static_cast<unsigned long long>(f) * qLen + m.qry_start;
        unsigned int const       nq   = m.qry_end - m.qry_start;
        unsigned char const *    ts   = P.ix.seqs + sbjBase(P.ix, m.subj_id) + m.subj_start;
        unsigned int const       nt   = m.subj_end - m.subj_start;
        signed char const *      sM   = sMM + matrixOffset(P.ix, m.subj_id);

        // trace storage: every lane owns KS = roundup4(K) bytes per row and column block (word stores)
        constexpr unsigned int   KS     = (K + 3) / 4 * 4;
        unsigned int const       stride = (nq + 32 * K - 1) / (32 * K) * (32 * KS);
        unsigned char *          T      = TRACE ? P.trace + P.traceOff[task] : nullptr;

        // The running maximum is one integer: (score << 9) | (511 - column inside the block).  A strictly
        // greater key = higher score, or the same score in a smaller column; rows are visited in
        // increasing order, so the first row of that column wins -- the reference's "first strict
        // maximum in column-major order".  Keys of earlier column blocks get the best possible low bits.
        int          bestKey = 0;
        unsigned int bi = 0, bj = 0;

        for (unsigned int c0 = 0; c0 < nq; c0 += 32 * K)
        {
            unsigned int const colBase = c0 + lane * K;
            int                qoff[K]; // 32 * query residue (0 for columns past the query end, masked below)
            int                S[K], V[K];
#pragma unroll
            for (int r = 0; r < K; ++r)
            {
                unsigned int const i = colBase + r;
                qoff[r]              = (i < nq) ? 32 * static_cast<int>(qs[i]) : 0;
                S[r]                 = 0;       // row 0
                V[r]                 = kNegInf; // vertical gap above row 1
            }
            bestKey |= (1 << kColBits) - 1;
            bool const firstBlock = (c0 == 0);
            int        dLeft      = 0;           // S(i0-1, j-1)
            unsigned int outPrev  = packSH(0, kNegInf);
            unsigned int const nSteps = nt + 31;
            for (unsigned int s = 0; s < nSteps; ++s)
            {
                // left neighbour's (S, hgap) of the row this lane is about to compute
                unsigned int in = __shfl_up_sync(0xffffffffu, outPrev, 1);
                int const    j  = static_cast<int>(s) - static_cast<int>(lane); // 0-based row
                bool const   rowActive = (j >= 0) && (j < static_cast<int>(nt));
                if (lane == 0)
                    in = (firstBlock || !rowActive) ? packSH(0, kNegInf) : bnd[j];
                if (!rowActive)
                    continue; // lanes outside the window keep outPrev; their neighbours are inactive too
                int       sLeft = unpackLo(in);
                int       hLeft = unpackHi(in);
                int       diagS = dLeft;
                dLeft           = sLeft;
                int const tOff  = static_cast<int>(__ldg(ts + j));
                unsigned int traceWord[(K + 3) / 4];
#pragma unroll
                for (int w = 0; w < (K + 3) / 4; ++w)
                    traceWord[w] = 0;
#pragma unroll
                for (int r = 0; r < K; ++r)
                {
                    bool const colActive = colBase + r < nq;
                    int const  sub  = static_cast<int>(sM[qoff[r] + tOff]);
                    int const  diag = diagS + sub;
                    int const  a1 = hLeft + ge, b1 = sLeft + go; // horizontal: extend / open
                    int const  a2 = V[r] + ge, b2 = S[r] + go;   // vertical:   extend / open
                    int const  h = max(a1, b1);
                    int const  v = max(a2, b2);
                    int const  g = max(v, h);
                    int        cur = max(diag, g);
                    if (TRACE)
                    {
                        // CompleteTrace: ties set both bits (SQ/align/dp_formula.h:210-222)
                        unsigned int tv = (a1 >= b1 ? T_HORI : 0u) | (b1 >= a1 ? T_HOPEN : 0u) | (a2 >= b2 ? T_VERT : 0u) |
                                          (b2 >= a2 ? T_VOPEN : 0u);
                        unsigned int const t2 = (v >= h ? T_MAXV : 0u) | (h >= v ? T_MAXH : 0u);
                        tv |= (diag >= g ? T_DIAG : 0u) | (g >= diag ? t2 : 0u);
                        tv = (cur <= 0) ? 0u : tv;
                        traceWord[r / 4] |= tv << (8 * (r % 4));
                    }
                    cur   = max(cur, 0);
                    diagS = S[r]; // S(i, j-1) is the diagonal of column i+1
                    S[r]  = cur;
                    V[r]  = v;
                    sLeft = cur;
                    hLeft = h;
                    int const key = colActive ? ((cur << kColBits) | ((1 << kColBits) - 1 - static_cast<int>(lane * K + r))) : 0;
                    if (key > bestKey)
                    {
                        bestKey = key;
                        bi      = colBase + r + 1;
                        bj      = static_cast<unsigned int>(j) + 1;
                    }
                }
                outPrev = packSH(sLeft, hLeft);
                if (lane == 31 && c0 + 32 * K < nq)
                    bnd[j] = outPrev; // right edge of this column block, read by lane 0 of the next block
                if (TRACE)
                {
                    unsigned int * dst = reinterpret_cast<unsigned int *>(T + static_cast<unsigned long long>(j) * stride +
                                                                          (c0 / K) * KS + lane * KS);
#pragma unroll
                    for (int w = 0; w < (K + 3) / 4; ++w)
                        dst[w] = traceWord[w];
                }
            }
            __syncwarp();
        }

        // warp reduction: highest score; ties -> smallest column, then smallest row
        int best = bestKey >> kColBits;
#pragma unroll
        for (int off = 16; off > 0; off >>= 1)
        {
            int const          ob  = __shfl_down_sync(0xffffffffu, best, off);
            unsigned int const obi = __shfl_down_sync(0xffffffffu, bi, off);
            unsigned int const obj = __shfl_down_sync(0xffffffffu, bj, off);
            if (ob > best || (ob == best && ob > 0 && (obi < bi || (obi == bi && obj < bj))))
            {
                best = ob;
                bi   = obi;
                bj   = obj;
            }
        }
        if (lane == 0)
        {
            P.scores[task] = best;
            if (best > 32767 && P.overflowFlag)
                atomicOr(P.overflowFlag, 1u);
            if (TRACE)
            {
                P.bestPos[2 * task]     = bi;
                P.bestPos[2 * task + 1] = bj;
            }
        }
    }
}

// ---------------------------------------------------------------------------------------------
// traceback + statistics: one thread per alignment (a chain of dependent byte loads; the parallelism
// is across the ~10^5 surviving alignments)
// ---------------------------------------------------------------------------------------------

struct TracebackParams
{
    DevIndex                   ix;
    DevQueries                 Q;
    lgpu_match const *         tasks;
    unsigned int               nTasks;
    signed char const *        matrix; // 2 x (32 x 32)
    int const *                scores;
    unsigned int const *       bestPos;
    unsigned char const *      trace;
    unsigned long long const * traceOff;
    unsigned int               K;            // columns per lane used by the fill kernel (storage: roundup4(K) bytes)
    lgpu_hit *                 out;
    unsigned int const *       outIndex;     // record of task t goes to out[outIndex[t]] (nullptr: out[t])
    // second pass (lgpu_params.want_cigar): emit the runs of the path instead of the record
    int                        emit;
    unsigned int *             cigarOps;
    unsigned int const *       cigarOff; // per task
    unsigned int               cigarBase;
};

__global__ void __launch_bounds__(128) tracebackKernel(TracebackParams P)
{
    unsigned int const task = blockIdx.x * blockDim.x + threadIdx.x;
    if (task >= P.nTasks)
        return;
    lgpu_match const         m    = P.tasks[task];
    unsigned int const       q    = m.qry_id / P.Q.F;
    unsigned int const       f    = m.qry_id % P.Q.F;
    unsigned long long const qb   = P.Q.offs[q];
    unsigned int const       qLen = static_cast<unsigned int>(P.Q.offs[q + 1] - qb);
    unsigned char const *    qs   = P.Q.trans + P.Q.F * qb + static_cast<unsigned long long>(f) * qLen + m.qry_start;
    unsigned int const       nq   = m.qry_end - m.qry_start;
    unsigned int const       sId  = m.subj_id / P.ix.sbjFrames;
    unsigned char const *    ts   = P.ix.seqs + sbjBase(P.ix, m.subj_id) + m.subj_start;
    signed char const *      M    = P.matrix + matrixOffset(P.ix, m.subj_id);
    unsigned int const       KS     = (P.K + 3) / 4 * 4;
    unsigned int const       stride = (nq + 32 * P.K - 1) / (32 * P.K) * (32 * KS);
    unsigned char const *    T      = P.trace + P.traceOff[task];

    unsigned int       i = P.bestPos[2 * task], j = P.bestPos[2 * task + 1];
    unsigned int const bi = i, bj = j;
    unsigned int nMatch = 0, nMismatch = 0, nPositive = 0, nGapOpen = 0, nGapExt = 0, alnLen = 0;
    unsigned int nOps = 0;
    unsigned int * const ops = P.emit ? P.cigarOps + P.cigarOff[task] : nullptr;

    auto tr = [&](unsigned int ii, unsigned int jj) -> unsigned int {
        // column i-1 lives in strip (i-1) / K at byte (i-1) % K of that strip's KS-byte slot
        return (ii > 0 && jj > 0) ? T[static_cast<unsigned long long>(jj - 1) * stride + ((ii - 1) / P.K) * KS + (ii - 1) % P.K] : 0u;
    };

    if (P.scores[task] > 0)
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

    unsigned int const outSlot = P.outIndex ? P.outIndex[task] : task;
    if (P.emit)
    {
        P.out[outSlot].cigar_off = P.cigarBase + P.cigarOff[task];
        P.out[outSlot].cigar_len = nOps;
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
    h.score      = P.scores[task];
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
    P.out[outSlot] = h;
}

// out[t] = in[idx[t]]
__global__ void gatherTasksKernel(lgpu_match const * in, unsigned int const * idx, unsigned int n, lgpu_match * out)
{
    unsigned int const t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t < n)
        out[t] = in[idx[t]];
}

// upper bound of the runs of every alignment: gaps and aligned stretches alternate
__global__ void cigarCapKernel(lgpu_hit const * hits, unsigned int const * order, unsigned int n, unsigned int * cap)
{
    unsigned int const t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t < n)
        cap[t] = 2u * hits[order ? order[t] : t].n_gap_open + 1u;
}

// pass-1 filter: keep[t] = score passes both integer thresholds of its query
__global__ void filterKernel(lgpu_match const * tasks, int const * scores, unsigned int n, unsigned int F,
                             int const * minBit, int const * minEval, unsigned int * keep, unsigned long long * counters)
{
    unsigned int const t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n)
        return;
    unsigned int const q = tasks[t].qry_id / F;
    int const          s = scores[t];
    unsigned int       k = 1;
    if (s < minBit[q])
    {
        k = 0;
        atomicAdd(&counters[0], 1ull);
    }
    else if (s < minEval[q])
    {
        k = 0;
        atomicAdd(&counters[1], 1ull);
    }
    keep[t] = k;
}

__global__ void compactKernel(lgpu_match const * tasks, unsigned int const * keep, unsigned int const * posIncl, unsigned int n,
                              lgpu_match * out)
{
    unsigned int const t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t < n && keep[t])
        out[posIncl[t] - 1] = tasks[t];
}

} // namespace lgpu
