// Device side of what the reference does with the records of a batch after the alignments are computed:
//   identity cut-off + phase bookkeeping        src/search_algo.hpp:1308-1322, iterativeSearchPre/Post :1391-1460
//   writeRecords / _writeRecord                 src/search_algo.hpp:1335-1362 / :821-913
//     sort by (subject, coordinates, frames, bit score descending), unique on everything but the score,
//     stable sort by bit score descending, keep the best maxMatches
// The bit score is strictly increasing in the raw score ((lambda S - ln K) / ln 2, lambda > 0), so the integer score
// orders the records exactly like the reference's doubles; the doubles themselves (bit score, e-value) are filled in
// on the host from tables, with the host's libm, after the final records came back.
#pragma once

#include <cstdint>
#include <cuda_runtime.h>

#include "../../include/lambda_b200.h"

namespace lgpu
{

// After the trace pass of one phase: keep[t] = record passes the identity cut-off
// (src/search_algo.hpp:1310-1316: float identity = 100 * matches / alignment length, dropped if < idCutOff);
// kept records get their phase, and their query is marked as done for the iterative search.
__global__ void postTraceKernel(lgpu_hit * hits, unsigned int n, int idCutoff, unsigned char phase, unsigned int * keep,
                                unsigned int * qryHasHit, unsigned long long * nFailedIdentity)
{
    unsigned int const t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n)
        return;
    lgpu_hit const h        = hits[t];
    float const    identity = static_cast<float>(__ddiv_rn(__dmul_rn(100.0, static_cast<double>(static_cast<float>(h.n_match))),
                                                            static_cast<double>(static_cast<float>(h.aln_len))));
    bool const     ok       = !(identity < static_cast<float>(idCutoff));
    keep[t]                 = ok ? 1u : 0u;
    if (ok)
    {
        hits[t].phase      = phase;
        qryHasHit[h.q_id]  = 1u;
    }
    else
        atomicAdd(nFailedIdentity, 1ull);
}

// out[base + pos] = in[t] for kept records (posIncl = inclusive scan of keep)
__global__ void appendHitsKernel(lgpu_hit const * in, unsigned int const * keep, unsigned int const * posIncl, unsigned int n,
                                 lgpu_hit * out)
{
    unsigned int const t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t < n && keep[t])
        out[posIncl[t] - 1] = in[t];
}

// iterativeSearchPost: the queries without a surviving hit go on to phase 2
__global__ void notDoneKernel(unsigned int const * qryHasHit, unsigned int n, unsigned int * flag)
{
    unsigned int const q = blockIdx.x * blockDim.x + threadIdx.x;
    if (q < n)
        flag[q] = qryHasHit[q] ? 0u : 1u;
}

// active[pos] = q for flagged queries; info[0] = longest active query
__global__ void activeEmitKernel(unsigned int const * flag, unsigned int const * posIncl, unsigned long long const * offs,
                                 unsigned int n, unsigned int * active, unsigned int * info)
{
    unsigned int const q = blockIdx.x * blockDim.x + threadIdx.x;
    unsigned int       len = 0;
    if (q < n && flag[q])
    {
        active[posIncl[q] - 1] = q;
        len                    = static_cast<unsigned int>(offs[q + 1] - offs[q]);
    }
#pragma unroll
    for (int off = 16; off > 0; off >>= 1)
        len = max(len, __shfl_xor_sync(0xffffffffu, len, off));
    if ((threadIdx.x & 31u) == 0 && len)
        atomicMax(&info[0], len);
}

// ---- _writeRecord --------------------------------------------------------------------------------------------

// first order: query, then the reference's 7-tuple, then score descending (src/search_algo.hpp:844-853)
struct RecordLess
{
    lgpu_hit const * h;
    __device__ bool  operator()(unsigned int a, unsigned int b) const
    {
        lgpu_hit const & x = h[a];
        lgpu_hit const & y = h[b];
        if (x.q_id != y.q_id) return x.q_id < y.q_id;
        if (x.s_id != y.s_id) return x.s_id < y.s_id;
        if (x.q_start != y.q_start) return x.q_start < y.q_start;
        if (x.q_end != y.q_end) return x.q_end < y.q_end;
        if (x.s_start != y.s_start) return x.s_start < y.s_start;
        if (x.s_end != y.s_end) return x.s_end < y.s_end;
        if (x.q_frame != y.q_frame) return x.q_frame < y.q_frame;
        if (x.s_frame != y.s_frame) return x.s_frame < y.s_frame;
        return x.score > y.score;
    }
};

__device__ __forceinline__ bool sameAlignment(lgpu_hit const & x, lgpu_hit const & y)
{
    return x.q_id == y.q_id && x.s_id == y.s_id && x.q_start == y.q_start && x.q_end == y.q_end && x.s_start == y.s_start &&
           x.s_end == y.s_end && x.q_frame == y.q_frame && x.s_frame == y.s_frame;
}

// dup[t] = 1 iff sorted record t repeats its predecessor (std::unique keeps the first = best-scoring one)
__global__ void markDuplicatesKernel(lgpu_hit const * h, unsigned int const * idx, unsigned int n, unsigned char * dropped,
                                     unsigned long long * nDup)
{
    unsigned int const t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n)
        return;
    bool const dup   = t > 0 && sameAlignment(h[idx[t]], h[idx[t - 1]]);
    dropped[idx[t]]  = dup ? 1 : 0;
    if (dup)
        atomicAdd(nDup, 1ull);
}

// second order: surviving records first, by query, by score descending; stable, so equal scores keep the first order
// (std::stable_sort by bitScore, src/search_algo.hpp:866-871)
struct RankLess
{
    lgpu_hit const *      h;
    unsigned char const * dropped;
    __device__ bool       operator()(unsigned int a, unsigned int b) const
    {
        if (dropped[a] != dropped[b]) return dropped[a] < dropped[b];
        if (h[a].q_id != h[b].q_id) return h[a].q_id < h[b].q_id;
        return h[a].score > h[b].score;
    }
};

// segHead[t] = t if sorted record t is the first (surviving) record of its query, else 0 (max-scanned afterwards)
__global__ void queryHeadKernel(lgpu_hit const * h, unsigned int const * idx, unsigned char const * dropped, unsigned int n,
                                unsigned int * segHead)
{
    unsigned int const t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n)
        return;
    bool const head = t == 0 || dropped[idx[t]] || h[idx[t]].q_id != h[idx[t - 1]].q_id;
    segHead[t]      = head ? t : 0u;
}

// keep[t] = surviving record with rank < maxMatches inside its query; counts queries with a hit and the records cut
__global__ void rankCutKernel(unsigned int const * idx, unsigned char * dropped, unsigned int const * segStart, unsigned int n,
                              unsigned int maxMatches, unsigned int * keep, unsigned long long * counters /* [0] abundant, [1] queries */)
{
    unsigned int const t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n)
        return;
    unsigned int k = 0;
    if (!dropped[idx[t]])
    {
        unsigned int const rank = t - segStart[t];
        if (rank == 0)
            atomicAdd(&counters[1], 1ull);
        if (rank < maxMatches)
            k = 1;
        else
        {
            atomicAdd(&counters[0], 1ull);
            dropped[idx[t]] = 2; // cut, not a duplicate
        }
    }
    keep[t] = k;
}

__global__ void gatherFinalKernel(lgpu_hit const * h, unsigned int const * idx, unsigned int const * keep,
                                  unsigned int const * posIncl, unsigned int n, lgpu_hit * out)
{
    unsigned int const t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t < n && keep[t])
        out[posIncl[t] - 1] = h[idx[t]];
}

// pairs = distinct (query, subject) among the final records: walk the FIRST order (grouped by query, subject); a final
// record counts if no earlier record of its group is final
__global__ void countPairsKernel(lgpu_hit const * h, unsigned int const * idx1, unsigned char const * dropped, unsigned int n,
                                 unsigned long long * pairs)
{
    unsigned int const t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n || dropped[idx1[t]])
        return;
    lgpu_hit const & me = h[idx1[t]];
    for (unsigned int u = t; u > 0; --u)
    {
        lgpu_hit const & o = h[idx1[u - 1]];
        if (o.q_id != me.q_id || o.s_id != me.s_id)
            break;
        if (!dropped[idx1[u - 1]])
            return;
    }
    atomicAdd(pairs, 1ull);
}

// records of one context into a caller's device buffer, query ids rebased (multi-GPU gather, sub-batch merge)
__global__ void exportHitsKernel(lgpu_hit const * in, unsigned long long n, unsigned int qBase, unsigned int cigarBase, lgpu_hit * out)
{
    unsigned long long const t = blockIdx.x * static_cast<unsigned long long>(blockDim.x) + threadIdx.x;
    if (t >= n)
        return;
    lgpu_hit h = in[t];
    h.q_id += qBase;
    h.cigar_off += cigarBase;
    out[t] = h;
}

} // namespace lgpu
