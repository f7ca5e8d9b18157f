// DP pass 2 with checkpoints instead of a stored trace matrix (EXPERIMENTAL: LAMBDA_B200_TRACE=ckpt).
// Parity-green, but measured slower than the stored planes of kernels_dpx_trace.cuh on the benchmark
// workloads (profiles/r1_trace_ckpt_vs_planes.jsonl: the fill gets cheaper, the per-thread tile
// recomputation of the traceback costs more than it saves), so it is not the default.
//
// Reference: _performAlignment<withTrace=true> (src/search_algo.hpp:1296), trace bits as defined in
// SQ/align/dp_formula.h:136-243 + dp_formula_affine.h:66-126 (CompleteTrace: ties set both bits),
// _doTraceback (SQ/align/dp_traceback_impl.h:223-474), computeAlignmentStats
// (SQ/align/evaluate_alignment.h:215-300).
//
// A stored trace costs 1-3 bytes per DP cell of HBM writes, which made pass 2 write-bound
// (kernels_dpx_trace.cuh: 40 GB per benchmark step).  The traceback only ever looks at the ~600 cells on
// the alignment path, so the fill kernel here stores just enough to RECOMPUTE any 8-row x K-column tile
// exactly:
//   row checkpoints   every 8th wavefront step: (W = H + go, E entering the row) of all columns
//   column checkpoints every step: (W, F entering the cell) of the last column of every strip
// = 0.5 + 4/K bytes per cell (0.7 B/cell for K = 20), laid out as one record per (8-step block, strip
// pair): [E of K columns | W of K columns | (W, F) of the last column for the 8 steps], staged through
// shared memory so that the warp writes whole records with coalesced 16-byte stores and the traceback
// finds everything a tile needs in ~7 sectors.  The fill loop is the score kernel's
// (kernels_dpx.cuh, same packed-int16 DPX recurrence, T in {8,16,32} lanes per alignment, 32/T
// alignments per warp, each with its own query profile) plus a per-column running maximum.
// The traceback kernel (one thread per alignment) walks the path tile by tile: it rebuilds the tile the
// path is in from the two checkpoints with the scalar recurrence (H, H - E, H - F per cell), derives
// SeqAn's trace byte of a cell from those exactly like kernels_dpx_trace.cuh does, and accumulates the
// alignment statistics.  The end cell follows the reference's rule (first strict maximum in column-major
// order): the fill kernel reports the smallest column holding the best score and the 8-step block in
// which that column's strip first reached it; the traceback scans that strip downwards from there.
#pragma once

#include <cstdint>
#include <cuda_runtime.h>

#include "../../include/lambda_b200.h"
#include "kernels_dpx.cuh"
#include "kernels_extend.cuh"

namespace lgpu
{

constexpr int kCkRows = 8; // rows per tile = wavefront steps between two row checkpoints

// (T, K) classes of the checkpointing fill kernel, ascending by columns (2 * T * K)
#define LGPU_CK_CLASSES(X)                                                                                            \
    X(8, 4) X(8, 8) X(8, 10) X(8, 12) X(16, 8) X(16, 10) X(16, 12) X(16, 16) X(16, 20) X(16, 24) X(16, 32) X(32, 20)  \
    X(32, 24) X(32, 32)
#define LGPU_CK_CLASS_ENTRY(T, K) {T, K},
#define LGPU_CK_CLASS_COUNT(T, K) +1
constexpr int kNumCkClasses = 0 LGPU_CK_CLASSES(LGPU_CK_CLASS_COUNT);
__host__ __device__ inline DpxClass ckClass(int cls)
{
    constexpr DpxClass tab[] = {LGPU_CK_CLASSES(LGPU_CK_CLASS_ENTRY)};
    return tab[cls];
}
__host__ __device__ inline int ckClassOf(unsigned int nq)
{
    for (int c = 0; c < kNumCkClasses; ++c)
    {
        DpxClass const k = ckClass(c);
        if (nq <= static_cast<unsigned int>(2 * k.T * k.K))
            return c;
    }
    return kNumCkClasses; // scalar wavefront kernel
}

// checkpoint words of one alignment with `nt` subject rows: records [block][lane][2K + 16], one block per
// 8 wavefront steps (the last one may be partial)
__host__ __device__ inline unsigned int ckSteps(int T, unsigned int nt) { return nt + 2u * static_cast<unsigned int>(T) - 1u; }
__host__ __device__ constexpr unsigned int ckRecWords(int K) { return 2u * static_cast<unsigned int>(K) + 2u * kCkRows; }
__host__ __device__ inline unsigned long long ckWords(int T, int K, unsigned int nt)
{
    unsigned long long const nBlk = (ckSteps(T, nt) + kCkRows - 1) / kCkRows;
    return nBlk * static_cast<unsigned int>(T) * ckRecWords(K);
}

struct CkTraceParams
{
    DevIndex                   ix;
    DevQueries                 Q;
    lgpu_match const *         tasks;
    unsigned int const *       order;  // tasks of this class (indices into `tasks`)
    unsigned int               nTasks; // entries in `order`
    signed char const *        matrix; // 2 x (32 x 32)
    int                        go, ge;
    unsigned int               nCodes; // alphabet size + 1 (last row = null)
    unsigned int               winCap; // bytes reserved per group for the padded window
    unsigned int *             workCounter;
    unsigned int *             ck;    // checkpoints
    unsigned long long const * ckOff; // word offset of every task's checkpoints (indexed by task)
    int *                      scores;   // out, indexed by task
    unsigned int *             bestCol;  // out: 1-based column of the end cell
    unsigned int *             firstBlk; // out: block in which the end cell's strip first reached the best score
};

template <int T, int K>
__global__ void __launch_bounds__(32) swTraceCkKernel(CkTraceParams P)
{
    constexpr int KW   = (K + 3) / 4;
    constexpr int ROWW = dpxRowWords(T, K);
    constexpr int PAD  = 2 * T;
    constexpr int G    = 32 / T;

    extern __shared__ unsigned int smem[];
    unsigned int const lane      = threadIdx.x;
    unsigned int const grp       = lane / T;
    unsigned int const gl        = lane % T;
    unsigned int const profWords = P.nCodes * ROWW;
    unsigned int *     prof      = smem + grp * profWords;
    unsigned char *    win       = reinterpret_cast<unsigned char *>(smem + G * profWords) + grp * (P.winCap + 32);
    constexpr unsigned int REC   = ckRecWords(K);
    static_assert(K % 2 == 0, "records are moved as 16-byte words");
    // staging area of the group's current block record, [lane][REC]
    unsigned int *     stage = reinterpret_cast<unsigned int *>(reinterpret_cast<unsigned char *>(smem + G * profWords) +
                                                            G * (P.winCap + 32)) + grp * (T * REC);
    unsigned int *     my    = stage + gl * REC;
    unsigned int const groupMask = (T == 32) ? 0xffffffffu : (((1u << T) - 1u) << (grp * T));
    unsigned int const nullCode  = P.nCodes - 1;

    unsigned int const go2  = (static_cast<unsigned int>(P.go) & 0xffffu) * 0x10001u;
    unsigned int const ge2  = (static_cast<unsigned int>(P.ge) & 0xffffu) * 0x10001u;
    unsigned int const neg2 = 0xE000E000u; // -8192: far below any real gap value, and H - E cannot overflow int16

    for (;;)
    {
        unsigned int job = 0;
        if (lane == 0)
            job = atomicAdd(P.workCounter, 1u);
        job = __shfl_sync(0xffffffffu, job, 0);
        if (job * G >= P.nTasks)
            break;
        unsigned int const slot  = job * G + grp;
        bool const         valid = slot < P.nTasks;
        unsigned int       task = 0, nq = 0, nt = 0;
        unsigned char const * qs = nullptr;
        unsigned char const * ts = nullptr;
        signed char const *   M  = P.matrix;
        if (valid)
        {
            task                          = P.order[slot];
            lgpu_match const         m    = P.tasks[task];
            unsigned int const       q    = m.qry_id / P.Q.F;
            unsigned int const       f    = m.qry_id % P.Q.F;
            unsigned long long const qb   = P.Q.offs[q];
            unsigned int const       qLen = static_cast<unsigned int>(P.Q.offs[q + 1] - qb);
            qs = P.Q.trans + P.Q.F * qb + static_cast<unsigned long long>(f) * qLen + m.qry_start;
            nq = m.qry_end - m.qry_start;
            ts = P.ix.seqs + sbjBase(P.ix, m.subj_id) + m.subj_start;
            nt = m.subj_end - m.subj_start;
            M  = P.matrix + matrixOffset(P.ix, m.subj_id);
        }
        unsigned int ntMax = nt;
#pragma unroll
        for (int off = 16; off > 0; off >>= 1)
            ntMax = max(ntMax, __shfl_xor_sync(0xffffffffu, ntMax, off));
        unsigned int const nSteps    = ntMax + 2 * T - 1;
        unsigned int const nStepsOwn = valid ? nt + 2 * T - 1 : 0u;
        unsigned int *     recBase   = P.ck + (valid ? P.ckOff[task] : 0ull);

        __syncwarp();
        // ---- query profile of this group: P[c][w][v], byte r%4 of word w = r/4 of strip v <-> column v*K + r ----
        for (unsigned int idx = gl; idx < profWords; idx += T)
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
        for (unsigned int idx = gl; idx < P.winCap; idx += T)
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
        unsigned int prevMax = go2, fbLo = 0, fbHi = 0; // strip maxima at the last checkpoint, block of their last rise

        unsigned int wl[KW], wh[KW];
        {
            unsigned int const cLo = win[PAD - gl];
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
            unsigned int inW = __shfl_sync(0xffffffffu, outW, (lane - 1) & (T - 1), T);
            unsigned int inF = __shfl_sync(0xffffffffu, outF, (lane - 1) & (T - 1), T);
            if (gl == 0)
            {
                inW = prmt(go2, inW, 0x5410);
                inF = prmt(neg2, inF, 0x5410);
            }
            bool const own    = s < nStepsOwn;
            bool const ckStep = ((s & (kCkRows - 1)) == kCkRows - 1) && own;
            bool const flushB = own && (ckStep || s == nStepsOwn - 1); // the last block may be partial
            if (flushB)
            {
                // E entering the checkpoint row of every column
#pragma unroll
                for (int r = 0; r < K; ++r)
                    my[r] = E[r];
            }
            unsigned int diag = diagIn;
            diagIn            = inW;
            unsigned int F    = inF;
            unsigned int fLast = inF; // F entering the strip's last column
#pragma unroll
            for (int r = 0; r < K; ++r)
            {
                unsigned int const b   = r & 3;
                unsigned int const sel = ((0xCu + b) << 12) | ((4u + b) << 8) | ((8u + b) << 4) | b;
                unsigned int const sub = prmt(wl[r >> 2], wh[r >> 2], sel);
                unsigned int const t   = __viaddmax_s16x2_relu(diag, sub, E[r]);
                unsigned int const u   = __vadd2(t, go2);
                unsigned int const w   = __viaddmax_s16x2(F, go2, u);
                if (r == K - 1)
                    fLast = F;
                F     = __viaddmax_s16x2(F, ge2, u);
                E[r]  = __viaddmax_s16x2(E[r], ge2, w);
                diag  = W[r];
                W[r]  = w;
                CB[r] = __vmaxs2(CB[r], w);
            }
            outW = W[K - 1];
            outF = F;
            if (own)
                *reinterpret_cast<uint2 *>(my + 2 * K + (s & (kCkRows - 1)) * 2) = make_uint2(outW, fLast);
            if (ckStep)
            {
                unsigned int mx = CB[0];
#pragma unroll
                for (int r = 1; r < K; ++r)
                    mx = __vmaxs2(mx, CB[r]);
                unsigned int const ch = mx ^ prevMax;
                if (ch & 0xffffu)
                    fbLo = s / kCkRows;
                if (ch >> 16)
                    fbHi = s / kCkRows;
                prevMax = mx;
            }
            if (flushB)
            {
#pragma unroll
                for (int r = 0; r < K; ++r)
                    my[K + r] = W[r];
                __syncwarp(groupMask);
                // the group's record block [lane][REC] leaves as one contiguous run of 16-byte words
                uint4 const * src = reinterpret_cast<uint4 const *>(stage);
                uint4 *       dst = reinterpret_cast<uint4 *>(recBase + static_cast<unsigned long long>(s / kCkRows) * (T * REC));
                for (unsigned int idx = gl; idx < T * REC / 4; idx += T)
                    dst[idx] = src[idx];
                __syncwarp(groupMask);
            }
#pragma unroll
            for (int k = 0; k < KW; ++k)
            {
                wl[k] = nl[k];
                wh[k] = nh[k];
            }
        }
        {
            // the last, partial block
            unsigned int mx = CB[0];
#pragma unroll
            for (int r = 1; r < K; ++r)
                mx = __vmaxs2(mx, CB[r]);
            unsigned int const ch = mx ^ prevMax;
            if (ch & 0xffffu)
                fbLo = nStepsOwn / kCkRows;
            if (ch >> 16)
                fbHi = nStepsOwn / kCkRows;
        }
        // best score and the smallest column that holds it (padded columns never reach the maximum)
        int          best = P.go;
        unsigned int bcol = 0xffffffffu;
#pragma unroll
        for (int r = 0; r < K; ++r)
        {
            int const          lo = static_cast<int>(static_cast<short>(CB[r] & 0xffffu));
            int const          hi = static_cast<int>(CB[r]) >> 16;
            unsigned int const cl = gl * K + r, ch = (gl + T) * K + r; // 0-based columns
            if (lo > best || (lo == best && cl < bcol)) { best = lo; bcol = cl; }
            if (hi > best || (hi == best && ch < bcol)) { best = hi; bcol = ch; }
        }
#pragma unroll
        for (int off = T / 2; off > 0; off >>= 1)
        {
            int const          ob = __shfl_xor_sync(0xffffffffu, best, off);
            unsigned int const oc = __shfl_xor_sync(0xffffffffu, bcol, off);
            if (ob > best || (ob == best && oc < bcol)) { best = ob; bcol = oc; }
        }
        {
            unsigned int const strip = (bcol == 0xffffffffu) ? 0u : bcol / K;
            unsigned int const src   = (lane & ~(T - 1)) | (strip % T);
            unsigned int const lo    = __shfl_sync(0xffffffffu, fbLo, src);
            unsigned int const hi    = __shfl_sync(0xffffffffu, fbHi, src);
            if (valid && gl == 0)
            {
                P.scores[task]   = best - P.go;
                P.bestCol[task]  = bcol + 1;
                P.firstBlk[task] = (strip / T) ? hi : lo;
            }
        }
    }
}

// ---------------------------------------------------------------------------------------------
// traceback on recomputed tiles
// ---------------------------------------------------------------------------------------------

struct TracebackCkParams
{
    DevIndex                   ix;
    DevQueries                 Q;
    lgpu_match const *         tasks;
    unsigned int const *       order;
    unsigned int               nTasks;
    signed char const *        matrix; // 2 x (32 x 32)
    int                        go, ge;
    unsigned int               T, K;
    int const *                scores;
    unsigned int const *       bestCol;
    unsigned int const *       firstBlk;
    unsigned int const *       ck;
    unsigned long long const * ckOff;
    lgpu_hit *                 out; // indexed by task
};

constexpr int kCkTbThreads = 64; // alignments per block of the traceback kernel

// shared memory of one traceback block: H (int16) and dE | dF << 4 (uint8) of a (8 + 1) x (K + 1) tile per
// thread, thread-minor so that the threads of a warp never collide on a bank
__host__ __device__ constexpr unsigned int ckTbSmemBytes(int K)
{
    return static_cast<unsigned int>((kCkRows + 1) * (K + 1) * kCkTbThreads * 3);
}

template <int K>
__global__ void __launch_bounds__(kCkTbThreads) tracebackCkKernel(TracebackCkParams P)
{
    extern __shared__ unsigned char tbSmem[];
    short *         sH  = reinterpret_cast<short *>(tbSmem);
    unsigned char * sN  = tbSmem + (kCkRows + 1) * (K + 1) * kCkTbThreads * 2;
    unsigned int const tid = threadIdx.x;
    unsigned int const t   = blockIdx.x * blockDim.x + threadIdx.x;
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
    int const                nq   = static_cast<int>(m.qry_end - m.qry_start);
    int const                nt   = static_cast<int>(m.subj_end - m.subj_start);
    int const                T    = static_cast<int>(P.T);
    int const                go = P.go, ge = P.ge, D = P.ge - P.go;
    constexpr unsigned int   REC  = ckRecWords(K);
    unsigned int const *     recBase = P.ck + P.ckOff[task];

    // the current tile: strip tileV, block tileB; cells indexed [y + 1][x + 1], y = -1 .. 7, x = -1 .. K-1
#define LGPU_HH(y1, x1) sH[((y1) * (K + 1) + (x1)) * kCkTbThreads + tid]
#define LGPU_NN(y1, x1) sN[((y1) * (K + 1) + (x1)) * kCkTbThreads + tid]
    int tileV = -1, tileB = -1, tileJ0 = 0;

    auto half16 = [](unsigned int w, int half) -> int {
        return half ? (static_cast<int>(w) >> 16) : static_cast<int>(static_cast<short>(w & 0xffffu));
    };
    auto loadTile = [&](int v, int b) {
        int const gl = v % T, half = v / T;
        int const j0 = kCkRows * b - v; // 0-based subject row of y = 0
        int       Eout[K];
        // top halo: the row checkpoint of block b - 1 (or the matrix border)
        if (j0 - 1 >= 0)
        {
            unsigned int const * rec = recBase + (static_cast<unsigned long long>(b - 1) * T + gl) * REC;
#pragma unroll
            for (int x = 0; x < K; ++x)
            {
                int const ein = half16(__ldg(rec + x), half), w = half16(__ldg(rec + K + x), half);
                int const H   = w - go;
                LGPU_HH(0, x + 1) = static_cast<short>(H);
                LGPU_NN(0, x + 1) = static_cast<unsigned char>(min(H - ein, 15));
                Eout[x]           = max(ein + ge, w);
            }
        }
        else
        {
#pragma unroll
            for (int x = 0; x < K; ++x)
            {
                // border row: H = 0, a gap opened there enters row 0 with go
                LGPU_HH(0, x + 1) = 0;
                LGPU_NN(0, x + 1) = 15;
                Eout[x]           = go;
            }
        }
        // left halo (and the corner): the column checkpoints of strip v - 1 (or the matrix border)
        int       Fout[kCkRows];
        int const glL = (v - 1 + T) % T, halfL = (v - 1) / T;
#pragma unroll
        for (int y = -1; y < kCkRows; ++y)
        {
            int const jr = j0 + y;
            int       H = 0, dF = 15, fo = go;
            if (v > 0 && jr >= 0 && jr < nt)
            {
                int const            sp  = jr + v - 1; // step at which strip v - 1 worked on row jr
                unsigned int const * rec = recBase + (static_cast<unsigned long long>(sp / kCkRows) * T + glL) * REC;
                uint2 const          cw  = __ldg(reinterpret_cast<uint2 const *>(rec + 2 * K + (sp % kCkRows) * 2));
                int const            w = half16(cw.x, halfL), fin = half16(cw.y, halfL);
                H  = w - go;
                dF = min(H - fin, 15);
                fo = max(fin + ge, w);
            }
            LGPU_HH(y + 1, 0) = static_cast<short>(H);
            LGPU_NN(y + 1, 0) = static_cast<unsigned char>(dF << 4);
            if (y >= 0)
                Fout[y] = fo;
        }
#pragma unroll
        for (int y = 0; y < kCkRows; ++y)
        {
            int const jr = j0 + y;
            if (jr >= nt)
                break;
            if (jr < 0)
            {
                // above the matrix: border row (H = 0); E entering row 0 stays go
#pragma unroll
                for (int x = 0; x <= K; ++x)
                    LGPU_HH(y + 1, x) = 0;
                continue;
            }
            int                F  = Fout[y];
            unsigned int const sc = ts[jr];
            int                hd = LGPU_HH(y, 0);
#pragma unroll
            for (int x = 0; x < K; ++x)
            {
                int const ic = v * K + x;
                if (ic < nq)
                {
                    int const ein = Eout[x];
                    int const up  = LGPU_HH(y, x + 1); // H(ic, jr - 1): the diagonal of the next column
                    int       h   = hd + static_cast<int>(M[qs[ic] * 32 + sc]);
                    h             = max(max(h, ein), max(F, 0));
                    LGPU_HH(y + 1, x + 1) = static_cast<short>(h);
                    LGPU_NN(y + 1, x + 1) = static_cast<unsigned char>(min(h - ein, 15) | (min(h - F, 15) << 4));
                    Eout[x]               = max(ein + ge, h + go);
                    F                     = max(F + ge, h + go);
                    hd                    = up;
                }
            }
        }
        tileV  = v;
        tileB  = b;
        tileJ0 = j0;
    };
    // SeqAn's trace byte of cell (ii, jj) (1-based) inside the loaded tile; 0 on the matrix border
    auto tr = [&](int ii, int jj) -> unsigned int {
        if (ii == 0 || jj == 0)
            return 0u;
        int const x = (ii - 1) - tileV * K, y = (jj - 1) - tileJ0;
        int const H = LGPU_HH(y + 1, x + 1);
        if (H <= 0)
            return 0u;
        unsigned int const n   = LGPU_NN(y + 1, x + 1);
        int const          dE = static_cast<int>(n & 15u), dF = static_cast<int>(n >> 4);
        int const          dFl = (ii > 1) ? static_cast<int>(LGPU_NN(y + 1, x) >> 4) : 15;
        int const          dEu = (jj > 1) ? static_cast<int>(LGPU_NN(y, x + 1) & 15u) : 15;
        unsigned int       tv = (dFl <= D ? T_HORI : 0u) | (dFl >= D ? T_HOPEN : 0u) | (dEu <= D ? T_VERT : 0u) |
                          (dEu >= D ? T_VOPEN : 0u);
        if (min(dE, dF) > 0)
            tv |= T_DIAG;
        else
        {
            tv |= (dE <= dF ? T_MAXV : 0u) | (dF <= dE ? T_MAXH : 0u);
            int const hd = (ii > 1 && jj > 1) ? static_cast<int>(LGPU_HH(y, x)) : 0;
            if (hd + static_cast<int>(M[qs[ii - 1] * 32 + ts[jj - 1]]) == H)
                tv |= T_DIAG;
        }
        return tv;
    };

    int const score = P.scores[task];
    int       i = (score > 0) ? static_cast<int>(P.bestCol[task]) : 0, j = 0; // no positive cell: empty alignment at (0, 0)
    bool      scanning = score > 0;
    int       scanB    = static_cast<int>(P.firstBlk[task]);
    int       bi = i, bj = 0;
    unsigned int nMatch = 0, nMismatch = 0, nPositive = 0, nGapOpen = 0, nGapExt = 0, alnLen = 0;
    int          last = 0; // 0 diag, 1 horizontal, 2 vertical
    unsigned int run  = 0;
    int          mode = 0; // 0 decide, 1 inside a vertical run, 2 inside a horizontal run
    bool         first = true, done = score <= 0;

    auto flush = [&]() {
        if (run)
        {
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

    while (!done)
    {
        if (!scanning && mode == 0 && !(i > 0 && j > 0))
            break;
        // ---- one tile per outer iteration (keeps the threads of a warp in step) ----
        int const v = (i - 1) / K;
        if (scanning)
        {
            // first row of the best column that holds the best score
            loadTile(v, scanB);
            int const x = (i - 1) - v * K;
            for (int y = 0; y < kCkRows; ++y)
            {
                int const jr = tileJ0 + y;
                if (jr >= 0 && jr < nt && LGPU_HH(y + 1, x + 1) == score)
                {
                    j        = jr + 1;
                    bj       = j;
                    scanning = false;
                    break;
                }
            }
            if (scanning)
            {
                ++scanB;
                if (kCkRows * scanB - v >= nt) // cannot happen: the fill kernel saw the score in this column
                {
                    scanning = false;
                    done     = true;
                    i = j = 0;
                    bi = bj = 0;
                }
                continue;
            }
        }
        else
        {
            int const b = ((j - 1) + v) / kCkRows;
            if (v != tileV || b != tileB)
                loadTile(v, b);
        }
        // ---- walk while the path stays inside the tile ----
        for (;;)
        {
            if (mode == 0 && !(i > 0 && j > 0))
            {
                done = true;
                break;
            }
            if (i > 0 && j > 0 && ((i - 1) < tileV * K || (j - 1) < tileJ0))
                break; // left the tile
            unsigned int tv = tr(i, j);
            if (first)
            {
                first = false;
                if (tv & T_MAXV) { tv &= (T_VERT | T_VOPEN | T_MAXV); last = 2; }
                else if (tv & T_MAXH) { tv &= (T_HORI | T_HOPEN | T_MAXH); last = 1; }
                else last = 0;
            }
            if (mode == 0)
            {
                if (tv == 0)
                {
                    done = true;
                    break;
                }
                if (tv & T_DIAG)
                {
                    switchTo(0);
                    unsigned int const a = qs[i - 1], b2 = ts[j - 1];
                    if (alignedIdentical(P.ix, M, a, b2)) ++nMatch; else ++nMismatch;
                    if (M[a * 32 + b2] > 0) ++nPositive;
                    --i; --j; ++run;
                    continue;
                }
                else if ((tv & T_MAXV) && (tv & T_VERT))
                {
                    switchTo(2);
                    mode = 1;
                }
                else if ((tv & T_MAXV) && (tv & T_VOPEN))
                {
                    switchTo(2);
                    --j; ++run;
                    continue;
                }
                else if ((tv & T_MAXH) && (tv & T_HORI))
                {
                    switchTo(1);
                    mode = 2;
                }
                else if ((tv & T_MAXH) && (tv & T_HOPEN))
                {
                    switchTo(1);
                    --i; ++run;
                    continue;
                }
                else
                {
                    done = true;
                    break;
                }
            }
            if (mode == 1)
            {
                if ((!(tv & T_VOPEN) || (tv & T_VERT)) && j != 1)
                {
                    --j; ++run;
                }
                else
                {
                    --j; ++run;
                    mode = 0;
                }
            }
            else if (mode == 2)
            {
                if ((!(tv & T_HOPEN) || (tv & T_HORI)) && i != 1)
                {
                    --i; ++run;
                }
                else
                {
                    --i; ++run;
                    mode = 0;
                }
            }
        }
    }
    if (score > 0)
        flush();

    lgpu_hit h;
    h.q_id       = q;
    h.s_id       = sId;
    h.q_start    = m.qry_start + static_cast<unsigned int>(i);
    h.q_end      = m.qry_start + static_cast<unsigned int>(bi);
    h.s_start    = m.subj_start + static_cast<unsigned int>(j);
    h.s_end      = m.subj_start + static_cast<unsigned int>(bj);
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
#undef LGPU_HH
#undef LGPU_NN
}

} // namespace lgpu
