// FM-index primitives and the seeding kernel (sm_100a).
//
// Device restatement of
//   rank            FMC occtable/InterleavedEPRV2.h:81-86,211-216
//   extendRight     FMC ReverseFMIndexCursor.h:30-34
//   rank_symbol     FMC occtable/InterleavedEPRV2.h:121-138,264-270
//   CSA / bitvector FMC CSA.h:104-113, BitvectorCompact.h:26-72
//   locate          FMC ReverseFMIndex.h:62-91, locate.h:28-35
//   search()        reference src/search_algo.hpp:607-762 (+ searchHalfExactImpl :538-604,
//                   seedLooksPromising :427-481)
//
// Work decomposition: seeding has a serial dependency inside one query (`hitsThisSeq` steers the
// adaptive seed elongation of every later seed, search_algo.hpp:695-703), so the unit of
// parallelism is the query: one thread walks one query's frames/seeds in the reference's order.
// The index (GBs) lives in HBM; every LF step is two dependent random block reads, so throughput
// comes from the >10^5 independent chains in flight, not from bandwidth.
#pragma once

#include <cstdint>
#include <cuda_runtime.h>

#include "../../include/lambda_b200.h"
#include "n_random.hpp"

namespace lgpu
{

struct CsaSuperDev
{
    unsigned long long entry;
    unsigned char      blocks[4];
    unsigned char      pad[4];
    unsigned long long bits[4];
};
static_assert(sizeof(CsaSuperDev) == 48, "csa superblock layout");

struct DevIndex
{
    unsigned char const *      occ;
    unsigned long long const * super;
    unsigned long long const * ssa;
    CsaSuperDev const *        csa;
    // Subject sequences the alignments run on (the reference's transSbjSeqs), translated-alphabet
    // ranks, 1 B/residue: the stored sequences themselves, or -- for TBLASTN/TBLASTX -- their six-frame
    // translations made once at index creation.  Frame-expanded subject id s lives at
    // seqs[seqDelims[s >> sbjShift] ...): bisulfite subjects 2k and 2k+1 are the same sequence
    // (views::duplicate, src/view_duplicate.hpp:50-53), translated frames have their own entries.
    unsigned char const *      seqs;
    unsigned long long const * seqDelims;
    unsigned long long const * origDelims; // delimiters of the original (untranslated) sequences
    unsigned int               sbjShift;   // 1: bisulfite, else 0
    unsigned int               sbjFrames;  // sbjNumFrames: 1, 2 (bisulfite) or 6 (translated subjects)
    unsigned int               bsMode;     // 1: odd subjects are scored with the reverse bisulfite matrix
    unsigned long long         nSeqs;
    unsigned long long         nRows;      // C[sigma] = length of the BWT
    unsigned long long         posMask;
    unsigned int               bitsForPos;
    unsigned int               blockBytes, planesOff, sigma, sigmaBits;
    unsigned int               singleSuper; // 1: only one super block -> folded into Cbase
    unsigned long long         Cbase[32];   // C[s] (+ superBlocks[0][s] if singleSuper)
    // Occurrence table re-packed at index creation (packOccKernel) so that a rank touches ONE aligned 32-byte sector
    // (dna4: 24 B of bit planes + 4 x u16 counts) or one 64-byte line (Li10: 32 B planes + 10 x u16; dna3bs): the
    // reference's 48- / 80-byte blocks at 8-byte alignment straddle two to four sectors for the same 28 / 36 bytes.
    // Same arithmetic (FMC occtable/InterleavedEPRV2.h:81-86,211-216): counts are kept relative to groups of 1024
    // blocks (u16), `mid` holds C[s] + super block + count in front of every group.  Symbol 0 (the sentinel) is only
    // ranked by locate steps that run into a sequence border and keeps using the original blocks.
    unsigned char const *      occP;    // nullptr: not packed
    unsigned long long const * mid;     // [group][sigma]
    unsigned int               pStride; // 32 or 64
    unsigned int               pCntOff; // 8 * sigmaBits
};

__device__ __forceinline__ unsigned long long ldg64(void const * p)
{
    return __ldg(reinterpret_cast<unsigned long long const *>(p));
}

// start of frame-expanded subject `subjId` inside ix.seqs, and its length
__device__ __forceinline__ unsigned long long sbjBase(DevIndex const & ix, unsigned int subjId)
{
    return __ldg(ix.seqDelims + (subjId >> ix.sbjShift));
}
__device__ __forceinline__ unsigned long long sbjLength(DevIndex const & ix, unsigned int subjId)
{
    unsigned int const k = subjId >> ix.sbjShift;
    return __ldg(ix.seqDelims + k + 1) - __ldg(ix.seqDelims + k);
}
// byte offset of the scoring matrix for this subject inside the 2 x (32 x 32) matrix block
// (bisulfite: scoringSchemeAlignBSRev for odd subjects, src/search_algo.hpp:464-466,1098)
__device__ __forceinline__ unsigned int matrixOffset(DevIndex const & ix, unsigned int subjId)
{
    return (subjId & ix.bsMode) << 10;
}

// canonical genetic code over dna5 ranks (uploaded from kDna5Translate at index creation)
__constant__ unsigned char cDna5Translate[125];

__device__ __forceinline__ unsigned long long fmSymbolMask(DevIndex const & ix, unsigned char const * b, unsigned int symb)
{
    unsigned long long mask = ~0ull;
#pragma unroll
    for (unsigned int p = 0; p < 5; ++p)
        if (p < ix.sigmaBits)
        {
            unsigned long long const plane = ldg64(b + ix.planesOff + 8 * p);
            mask &= ((symb >> p) & 1u) ? plane : ~plane;
        }
    return mask;
}

// bit mask of the positions of `symb` inside a packed block (planes at the start of the block)
__device__ __forceinline__ unsigned long long fmSymbolMaskPacked(DevIndex const & ix, unsigned char const * b, unsigned int symb)
{
    unsigned long long mask = ~0ull;
#pragma unroll
    for (unsigned int p = 0; p < 5; ++p)
        if (p < ix.sigmaBits)
        {
            unsigned long long const plane = ldg64(b + 8 * p);
            mask &= ((symb >> p) & 1u) ? plane : ~plane;
        }
    return mask;
}

__device__ __forceinline__ unsigned long long fmRankPacked(DevIndex const & ix, unsigned long long idx, unsigned int symb)
{
    unsigned char const *    b    = ix.occP + (idx >> 6) * ix.pStride;
    unsigned int const       cnt  = __ldg(reinterpret_cast<unsigned short const *>(b + ix.pCntOff) + (symb - 1u));
    unsigned long long const mask = fmSymbolMaskPacked(ix, b, symb);
    unsigned int const       bit  = static_cast<unsigned int>(idx) & 63u;
    unsigned int const       pc   = bit ? __popcll(mask << (64u - bit)) : 0u;
    return __ldg(ix.mid + (idx >> 16) * ix.sigma + symb) + cnt + pc;
}

__device__ __forceinline__ unsigned long long fmRank(DevIndex const & ix, unsigned long long idx, unsigned int symb)
{
    if (ix.occP && symb != 0)
        return fmRankPacked(ix, idx, symb);
    unsigned char const *    b    = ix.occ + (idx >> 6) * ix.blockBytes;
    unsigned int const       cnt  = __ldg(reinterpret_cast<unsigned int const *>(b) + symb);
    unsigned long long const mask = fmSymbolMask(ix, b, symb);
    unsigned int const       bit  = static_cast<unsigned int>(idx) & 63u;
    // std::bitset<64> << 64 yields 0 in the reference; a 64-bit shift by 64 would be undefined here
    unsigned int const pc = bit ? __popcll(mask << (64u - bit)) : 0u;
    unsigned long long r  = static_cast<unsigned long long>(cnt) + pc + ix.Cbase[symb];
    if (!ix.singleSuper)
        r += __ldg(ix.super + (idx >> 32) * ix.sigma + symb);
    return r;
}

// one LF step from BWT row idx: rank of the symbol found at idx
__device__ __forceinline__ unsigned long long fmRankSymbol(DevIndex const & ix, unsigned long long idx)
{
    if (ix.occP)
    {
        unsigned char const * pb   = ix.occP + (idx >> 6) * ix.pStride;
        unsigned int const    bit  = static_cast<unsigned int>(idx) & 63u;
        unsigned int          symb = 0;
        unsigned long long    mask = ~0ull;
#pragma unroll
        for (unsigned int p = 0; p < 5; ++p)
            if (p < ix.sigmaBits)
            {
                unsigned long long const plane = ldg64(pb + 8 * p);
                unsigned int const       v     = static_cast<unsigned int>(plane >> bit) & 1u;
                symb |= v << p;
                mask &= v ? plane : ~plane;
            }
        if (symb != 0)
        {
            unsigned int const cnt = __ldg(reinterpret_cast<unsigned short const *>(pb + ix.pCntOff) + (symb - 1u));
            unsigned int const pc  = bit ? __popcll(mask << (64u - bit)) : 0u;
            return __ldg(ix.mid + (idx >> 16) * ix.sigma + symb) + cnt + pc;
        }
    }
    unsigned char const * b    = ix.occ + (idx >> 6) * ix.blockBytes;
    unsigned int const    bit  = static_cast<unsigned int>(idx) & 63u;
    unsigned int          symb = 0;
    unsigned long long    mask = ~0ull;
#pragma unroll
    for (unsigned int p = 0; p < 5; ++p)
        if (p < ix.sigmaBits)
        {
            unsigned long long const plane = ldg64(b + ix.planesOff + 8 * p);
            unsigned int const       v     = static_cast<unsigned int>(plane >> bit) & 1u;
            symb |= v << p;
            mask &= v ? plane : ~plane;
        }
    unsigned int const cnt = __ldg(reinterpret_cast<unsigned int const *>(b) + symb);
    unsigned int const pc  = bit ? __popcll(mask << (64u - bit)) : 0u;
    unsigned long long r   = static_cast<unsigned long long>(cnt) + pc + ix.Cbase[symb];
    if (!ix.singleSuper)
        r += __ldg(ix.super + (idx >> 32) * ix.sigma + symb);
    return r;
}

__device__ __forceinline__ void fmLocate(DevIndex const & ix, unsigned long long row, unsigned long long & subj,
                                         unsigned long long & pos)
{
    unsigned long long steps = 0;
    for (;;)
    {
        unsigned long long const i  = row + 1; // bit i of the vector is stored at position i + 1
        CsaSuperDev const *      sb = ix.csa + (i >> 8);
        unsigned long long const w  = ldg64(&sb->bits[(i & 255u) >> 6]);
        if ((w >> (i & 63u)) & 1ull)
            break;
        row = fmRankSymbol(ix, row);
        ++steps;
    }
    CsaSuperDev const *      sb  = ix.csa + (row >> 8);
    unsigned int const       blk = static_cast<unsigned int>(row & 255u) >> 6;
    unsigned int const       bit = static_cast<unsigned int>(row) & 63u;
    unsigned long long const w   = ldg64(&sb->bits[blk]);
    unsigned long long const r   = ldg64(&sb->entry) + __ldg(&sb->blocks[blk]) + __popcll(w << (63u - bit));
    unsigned long long const v   = __ldg(ix.ssa + r);
    subj                         = v >> ix.bitsForPos;
    pos                          = (v & ix.posMask) - steps;
}

struct Cursor
{
    unsigned long long lb, len;
};

__device__ __forceinline__ Cursor fmExtendRight(DevIndex const & ix, Cursor c, unsigned int symb)
{
    unsigned long long const lb = fmRank(ix, c.lb, symb);
    Cursor                   n;
    n.lb  = lb;
    n.len = fmRank(ix, c.lb + c.len, symb) - lb;
    return n;
}

// packed block b = bit planes + u16 counts of symbols 1 .. sigma-1 relative to the first block of its group of 1024;
// mid[g][s] = C[s] + super block + count in front of group g  (C passed by value: Cbase may already contain the super block)
struct PackOccParams
{
    DevIndex             ix;
    unsigned long long   nBlocks;
    unsigned long long   C[32];
    unsigned char *      occP;
    unsigned long long * mid;
};

__global__ void packOccKernel(PackOccParams P)
{
    DevIndex const &         ix = P.ix;
    unsigned long long const b  = blockIdx.x * static_cast<unsigned long long>(blockDim.x) + threadIdx.x;
    if (b >= P.nBlocks)
        return;
    unsigned long long const b0   = b & ~1023ull;
    unsigned char const *    src  = ix.occ + b * ix.blockBytes;
    unsigned char const *    src0 = ix.occ + b0 * ix.blockBytes;
    unsigned char *          dst  = P.occP + b * ix.pStride;
    for (unsigned int p = 0; p < ix.sigmaBits; ++p)
        reinterpret_cast<unsigned long long *>(dst)[p] = ldg64(src + ix.planesOff + 8 * p);
    for (unsigned int s = 1; s < ix.sigma; ++s)
    {
        unsigned int const c  = __ldg(reinterpret_cast<unsigned int const *>(src) + s);
        unsigned int const c0 = __ldg(reinterpret_cast<unsigned int const *>(src0) + s);
        reinterpret_cast<unsigned short *>(dst + ix.pCntOff)[s - 1] = static_cast<unsigned short>(c - c0);
        if (b == b0)
            P.mid[(b >> 10) * ix.sigma + s] = P.C[s] + __ldg(ix.super + (b >> 26) * ix.sigma + s) + c0;
    }
    if (b == b0)
        P.mid[(b >> 10) * ix.sigma] = 0;
    for (unsigned int k = ix.pCntOff + 2 * (ix.sigma - 1); k < ix.pStride; ++k)
        dst[k] = 0;
}

// ---------------------------------------------------------------------------------------------
// known-answer test kernels
// ---------------------------------------------------------------------------------------------

__global__ void fmRankKernel(DevIndex ix, unsigned long long const * idx, unsigned char const * symb, unsigned long long n,
                             unsigned long long * out)
{
    unsigned long long const i = blockIdx.x * static_cast<unsigned long long>(blockDim.x) + threadIdx.x;
    if (i < n)
        out[i] = fmRank(ix, idx[i], symb[i]);
}

__global__ void fmLocateKernel(DevIndex ix, unsigned long long const * rows, unsigned long long n, unsigned long long * subj,
                               unsigned long long * pos)
{
    unsigned long long const i = blockIdx.x * static_cast<unsigned long long>(blockDim.x) + threadIdx.x;
    if (i < n)
        fmLocate(ix, rows[i], subj[i], pos[i]);
}

// Cross-checks of an uploaded index that are O(size) and therefore run here, not in the host parser (a corrupt or
// crafted file must not make the search kernels read out of bounds): flags |= 1 occ block counts do not continue the
// previous block, 2 a CSA super block ranks outside the sampled suffix array, 4 a sampled entry names a sequence or a
// position that does not exist.
__global__ void validateIndexKernel(DevIndex ix, unsigned long long nBlocks, unsigned long long nSsa, unsigned long long nCsaSb,
                                    unsigned long long nFrameSubjects, unsigned int * flags)
{
    unsigned long long const stride = static_cast<unsigned long long>(gridDim.x) * blockDim.x;
    unsigned int             bad    = 0;
    unsigned long long const fullBlocks = ix.nRows / 64; // blocks 0 .. fullBlocks-1 hold 64 BWT symbols each
    for (unsigned long long b = blockIdx.x * static_cast<unsigned long long>(blockDim.x) + threadIdx.x; b + 1 < nBlocks && b < fullBlocks;
         b += stride)
    {
        if (((b + 1) & ((1ull << 26) - 1)) == 0)
            continue; // the counts restart with every super block (2^32 symbols)
        unsigned char const * cur = ix.occ + b * ix.blockBytes;
        unsigned char const * nxt = cur + ix.blockBytes;
        for (unsigned int s = 0; s < ix.sigma; ++s)
        {
            unsigned int const c0 = __ldg(reinterpret_cast<unsigned int const *>(cur) + s);
            unsigned int const c1 = __ldg(reinterpret_cast<unsigned int const *>(nxt) + s);
            if (c1 != c0 + static_cast<unsigned int>(__popcll(fmSymbolMask(ix, cur, s))))
                bad |= 1u;
        }
    }
    for (unsigned long long k = blockIdx.x * static_cast<unsigned long long>(blockDim.x) + threadIdx.x; k < nCsaSb; k += stride)
    {
        CsaSuperDev const * sb = ix.csa + k;
        unsigned long long  r  = ldg64(&sb->entry);
        unsigned int        in = 0;
        for (unsigned int w = 0; w < 4; ++w)
        {
            // the writer fills the in-block counts only up to the word that holds the last row (bit i lives at i + 1)
            if (k * 256 + w * 64 <= ix.nRows && __ldg(&sb->blocks[w]) != in)
                bad |= 2u;
            in += static_cast<unsigned int>(__popcll(ldg64(&sb->bits[w])));
        }
        if (r + in > nSsa)
            bad |= 2u;
    }
    for (unsigned long long k = blockIdx.x * static_cast<unsigned long long>(blockDim.x) + threadIdx.x; k < nSsa; k += stride)
    {
        unsigned long long const v    = __ldg(ix.ssa + k);
        unsigned long long const subj = v >> ix.bitsForPos;
        if (subj >= nFrameSubjects || (v & ix.posMask) > sbjLength(ix, static_cast<unsigned int>(subj)) + 1) // positions 0 .. length + 1
            bad |= 4u;
    }
    if (bad)
        atomicOr(flags, bad);
}

// ---------------------------------------------------------------------------------------------
// query preparation: frames (reverse complement) and alphabet reduction
// ---------------------------------------------------------------------------------------------

// Frame-expanded layout: frame f of query q starts at F * offs[q] + f * len(q) (len = original length;
// translated frames are shorter and leave the rest of their slot unused).
//   frameMode 0  one frame (protein query)
//             1  forward, reverse complement                      (bio::views::add_reverse_complement)
//             2  six-frame translation                            (bio::views::translate_join)
//             3  bisulfite: fwd, fwd, rc, rc; even frames reduced C->T, odd frames G->A
//                (src/shared_definitions.hpp:257-281, src/view_reduce_to_bisulfite.hpp:51-52,133-135)
struct DevQueries
{
    unsigned char const *      orig;  // original ranks
    unsigned long long const * offs;  // n + 1
    unsigned char *            trans; // F * total
    unsigned char *            red;   // F * total
    unsigned int               n, F;
    unsigned int               frameMode;
    unsigned char              redTab[2][32]; // [frame & 1] (the two differ in bisulfite mode only)
    unsigned char              compTab[8];
};

// length of frame f of a query with `len` original residues (BIO ranges/views/translate_single.hpp:95-113)
__device__ __forceinline__ unsigned int qryFrameLen(DevQueries const & Q, unsigned int len, unsigned int f)
{
    if (Q.frameMode != 2)
        return len;
    unsigned int const o = f % 3;
    return (max(len, o) - o) / 3;
}

// _setFrames (src/search_algo.hpp:769-814)
__device__ __forceinline__ void setFrames(DevQueries const & Q, DevIndex const & ix, unsigned int qryId, unsigned int subjId,
                                          signed char & qFrame, signed char & sFrame)
{
    int qf = 0, sf = 0;
    if (Q.frameMode == 2)
    {
        qf = static_cast<int>(qryId % 3) + 1;
        if (qryId % 6 > 2)
            qf = -qf;
    }
    else if (Q.frameMode == 3)
    {
        qf = static_cast<int>(qryId % 2) + 1;
        if (qryId % 4 > 1)
            qf = -qf;
    }
    else if (Q.frameMode == 1)
        qf = (qryId % 2) ? -1 : 1;
    if (ix.sbjFrames == 6)
    {
        sf = static_cast<int>(subjId % 3) + 1;
        if (subjId % 6 > 2)
            sf = -sf;
    }
    else if (ix.bsMode)
        sf = static_cast<int>(subjId % 2) + 1;
    qFrame = static_cast<signed char>(qf);
    sFrame = static_cast<signed char>(sf);
}

// Reduced query symbol of a seed search / of an elongation step.  An 'N' of a nucleotide query is stored as
// kNMarker | (frame & 1) and resolved like the reference's views::dna_n_to_random does (n_random.hpp): inside a
// seed by the number of 'N's in front of it in that seed, in an elongation step always to the first random rank.
__device__ __forceinline__ unsigned int seedSym(unsigned char const * red, unsigned int seedBegin, unsigned int p, bool bisulfite)
{
    unsigned int const s = red[p];
    return s < kNMarker ? s : redSymbol(red, seedBegin, p, bisulfite);
}
__device__ __forceinline__ unsigned int elongSym(unsigned char const * red, unsigned int p, bool bisulfite)
{
    unsigned int const s = red[p];
    return s < kNMarker ? s : redSymbol(red, p, p, bisulfite);
}

__global__ void prepQueriesKernel(DevQueries Q)
{
    // one block per query keeps the index math trivial; residues are strided over the block
    for (unsigned int q = blockIdx.x; q < Q.n; q += gridDim.x)
    {
        unsigned long long const b   = Q.offs[q];
        unsigned int const       len = static_cast<unsigned int>(Q.offs[q + 1] - b);
        for (unsigned int f = 0; f < Q.F; ++f)
        {
            unsigned long long const o    = Q.F * b + static_cast<unsigned long long>(f) * len;
            unsigned int const       fLen = qryFrameLen(Q, len, f);
            bool const               rc   = Q.frameMode == 1 ? (f & 1u) : Q.frameMode == 2 ? (f >= 3) : Q.frameMode == 3 ? (f >= 2) : false;
            for (unsigned int k = threadIdx.x; k < fLen; k += blockDim.x)
            {
                unsigned char r;
                if (Q.frameMode == 2)
                {
                    unsigned int const p = 3 * k + f % 3;
                    unsigned int       n1, n2, n3;
                    if (!rc)
                    {
                        n1 = Q.orig[b + p];
                        n2 = Q.orig[b + p + 1];
                        n3 = Q.orig[b + p + 2];
                    }
                    else
                    {
                        n1 = Q.compTab[Q.orig[b + len - p - 1]];
                        n2 = Q.compTab[Q.orig[b + len - p - 2]];
                        n3 = Q.compTab[Q.orig[b + len - p - 3]];
                    }
                    r = cDna5Translate[(n1 * 5 + n2) * 5 + n3];
                }
                else
                    r = rc ? Q.compTab[Q.orig[b + len - 1 - k]] : Q.orig[b + k];
                Q.trans[o + k] = r;
                Q.red[o + k]   = Q.redTab[f & 1u][r];
            }
        }
    }
}

// six-frame translation of the stored subjects (TBLASTN / TBLASTX), once per index: frame-expanded
// subject 6 * s + f -> out[outDelims[6 * s + f] ...)
__global__ void translateSubjectsKernel(unsigned char const * seqs, unsigned long long const * delims, unsigned long long nSeqs,
                                        unsigned char const * compTab5, unsigned long long const * outDelims, unsigned char * out)
{
    for (unsigned long long sf = blockIdx.x; sf < nSeqs * 6; sf += gridDim.x)
    {
        unsigned long long const s = sf / 6;
        unsigned int const       f = static_cast<unsigned int>(sf % 6);
        unsigned long long const b = delims[s], len = delims[s + 1] - b;
        unsigned long long const o = outDelims[sf], fLen = outDelims[sf + 1] - o;
        for (unsigned long long k = threadIdx.x; k < fLen; k += blockDim.x)
        {
            unsigned long long const p = 3 * k + f % 3;
            unsigned int             n1, n2, n3;
            if (f < 3)
            {
                n1 = seqs[b + p];
                n2 = seqs[b + p + 1];
                n3 = seqs[b + p + 2];
            }
            else
            {
                n1 = compTab5[seqs[b + len - p - 1]];
                n2 = compTab5[seqs[b + len - p - 2]];
                n3 = compTab5[seqs[b + len - p - 3]];
            }
            out[o + k] = cDna5Translate[(n1 * 5 + n2) * 5 + n3];
        }
    }
}

// ---------------------------------------------------------------------------------------------
// seeding
// ---------------------------------------------------------------------------------------------

struct SeedParams
{
    DevIndex             ix;
    DevQueries           Q;
    unsigned int const * active; // list of query ids to seed
    unsigned int         nActive;
    unsigned int         seedLength, seedOffset, maxSeedDist, halfExact, adaptive;
    unsigned int         fullHamming; // max_seed_dist = 1 over the WHOLE seed (seed_half_exact off)
    unsigned int         maxMatches;
    int                  preScoring;
    double               preScoringThresh;
    unsigned int         unknownRank;
    signed char const *  matrix; // 2 x (32 x 32) int8, translated-alphabet ranks (second: bisulfite reverse)
    lgpu_match *         out;
    unsigned long long   cap;
    unsigned long long * counters; // [0] matches emitted, [1] hitsAfterSeeding, [2] hitsFailedPreExtendTest
    // depth-k prefix table: cursor of every k-mer over the reduced alphabet, in place of the first k LF steps of a seed's
    // exactly matched part (the same cursor the steps would produce); nullptr = none
    uint2 const *        prefixTab;
    unsigned int         prefixK;
    // A cursor with ONE occurrence is elongated by locating it once and comparing query and subject text directly (each
    // further LF step would just test the next subject residue); textElong = 0 keeps the rank steps
    unsigned int         textElong;
    unsigned char        sbjRed[2][32]; // subject residue -> reduced rank, per subject parity (bisulfite); 0xff = unknown
};

constexpr int kMaxHalf2 = 16; // longest supported second seed half (seed length <= 32)

__global__ void __launch_bounds__(128) seedKernel(SeedParams P)
{
    __shared__ signed char sM[2048];
    for (int i = threadIdx.x; i < 2048; i += blockDim.x)
        sM[i] = P.matrix[i];
    __syncthreads();

    unsigned int const t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= P.nActive)
        return;
    unsigned int const       q    = P.active[t];
    DevIndex const &         ix   = P.ix;
    unsigned int const       F    = P.Q.F;
    unsigned long long const qb   = P.Q.offs[q];
    unsigned int const       L    = P.seedLength;
    unsigned int const       redN = ix.sigma - 1;

    unsigned long long nAfter = 0, nFailed = 0;
    unsigned int const origLen = static_cast<unsigned int>(P.Q.offs[q + 1] - qb);
    if (qryFrameLen(P.Q, origLen, 0) >= L) // frame 0 is the longest frame
    {
        unsigned long long hitsThisSeq = 0;
        unsigned long long needlesSum  = 0;
        for (unsigned int f = 0; f < F; ++f)
            needlesSum += qryFrameLen(P.Q, origLen, f);
        unsigned long long needlesPos = 0;
        bool const         half       = (P.halfExact && P.maxSeedDist != 0) || P.fullHamming;
        unsigned int const h1         = P.fullHamming ? 0u : (half ? L / 2 : L);
        unsigned int const n2         = L - h1;

        for (unsigned int f = 0; f < F; ++f)
        {
            unsigned int const len = qryFrameLen(P.Q, origLen, f);
            if (len < L)
                continue; // too short a frame is skipped without advancing needlesPos (search_algo.hpp:637)
            unsigned int const    qryId = q * F + f;
            unsigned char const * trans = P.Q.trans + F * qb + static_cast<unsigned long long>(f) * origLen;
            unsigned char const * red   = P.Q.red + F * qb + static_cast<unsigned long long>(f) * origLen;

            for (unsigned int seedBegin = 0;; seedBegin += P.seedOffset)
            {
                while (seedBegin < len - L &&
                       (trans[seedBegin] == P.unknownRank || trans[seedBegin] == trans[seedBegin + 1]))
                    ++seedBegin;
                if (seedBegin > len - L)
                    break;

                // exact chain E[0..K]: E[0] = cursor after the exactly matched first half
                Cursor E[kMaxHalf2 + 1];
                int    K = -1;
                {
                    Cursor c;
                    c.lb  = 0;
                    c.len = ix.nRows;
                    bool ok = true;
                    for (unsigned int i = 0; i < h1; ++i)
                    {
                        c = fmExtendRight(ix, c, seedSym(red, seedBegin, seedBegin + i, ix.bsMode != 0) + 1u);
                        if (c.len == 0)
                        {
                            ok = false;
                            break;
                        }
                    }
                    if (ok)
                    {
                        E[0] = c;
                        K    = 0;
                        for (unsigned int i = 0; i < n2; ++i)
                        {
                            c = fmExtendRight(ix, c, seedSym(red, seedBegin, seedBegin + h1 + i, ix.bsMode != 0) + 1u);
                            if (c.len == 0)
                                break;
                            E[i + 1] = c;
                            K        = static_cast<int>(i) + 1;
                        }
                    }
                }
                if (K < 0)
                    continue;

                // Enumerate the half-exact cursors in the reference's BFS order (= leaf order of the
                // search tree): mismatches below the seed symbol per level going down, the exact
                // cursor, then mismatches above the seed symbol per level going back up.
                int const    kMax  = (K < static_cast<int>(n2) - 1) ? K : static_cast<int>(n2) - 1;
                int          stage = 0, lvl = 0;
                unsigned int r     = 0;
                for (;;)
                {
                    Cursor cursor;
                    bool   have = false;
                    while (!have)
                    {
                        if (stage == 0)
                        {
                            if (lvl > kMax)
                            {
                                stage = 1;
                                continue;
                            }
                            unsigned int const want = seedSym(red, seedBegin, seedBegin + h1 + lvl, ix.bsMode != 0);
                            if (r >= want)
                            {
                                ++lvl;
                                r = 0;
                                continue;
                            }
                        }
                        else if (stage == 1)
                        {
                            stage = 2;
                            lvl   = kMax;
                            r     = (lvl >= 0) ? seedSym(red, seedBegin, seedBegin + h1 + lvl, ix.bsMode != 0) + 1u : 0u;
                            if (K == static_cast<int>(n2))
                            {
                                cursor = E[n2];
                                have   = true;
                            }
                            continue;
                        }
                        else
                        {
                            if (lvl < 0)
                                break;
                            if (r >= redN)
                            {
                                --lvl;
                                r = (lvl >= 0) ? seedSym(red, seedBegin, seedBegin + h1 + lvl, ix.bsMode != 0) + 1u : 0u;
                                continue;
                            }
                        }
                        // mismatch r at level lvl, then the rest of the seed exactly
                        Cursor c = fmExtendRight(ix, E[lvl], r + 1u);
                        ++r;
                        for (unsigned int l = lvl + 1; c.len != 0 && l < n2; ++l)
                            c = fmExtendRight(ix, c, seedSym(red, seedBegin, seedBegin + h1 + l, ix.bsMode != 0) + 1u);
                        if (c.len != 0)
                        {
                            cursor = c;
                            have   = true;
                        }
                    }
                    if (!have)
                        break;

                    // ---- adaptive elongation (search_algo.hpp:679-726) ----
                    unsigned int seedLen = L;
                    if (P.adaptive)
                    {
                        unsigned long long desired = 1;
                        if (hitsThisSeq < P.maxMatches)
                        {
                            unsigned long long remaining = (needlesSum - needlesPos - seedBegin) / P.seedOffset;
                            if (remaining < 1)
                                remaining = 1;
                            desired = (P.maxMatches - hitsThisSeq) * 10ull / remaining;
                            if (desired == 0)
                                desired = 1;
                        }
                        unsigned long long oldCount = cursor.len;
                        while (seedBegin + seedLen < len)
                        {
                            Cursor const n = fmExtendRight(ix, cursor, elongSym(red, seedBegin + seedLen, ix.bsMode != 0) + 1u);
                            if (n.len < desired && n.len < oldCount)
                                break; // keep the previous cursor
                            cursor   = n;
                            oldCount = n.len;
                            ++seedLen;
                        }
                    }
                    // over-abundant seeds are dropped (search_algo.hpp:729)
                    if (cursor.len > 10ull * P.maxMatches)
                        continue;

                    // ---- locate + pre-scoring ----
                    for (unsigned long long row = cursor.lb; row < cursor.lb + cursor.len; ++row)
                    {
                        unsigned long long subj, pos;
                        fmLocate(ix, row, subj, pos);
                        pos -= seedLen; // the reversed-text cursor reports the seed's end
                        ++nAfter;

                        // seedLooksPromising
                        long long                qB     = seedBegin;
                        long long                sB     = static_cast<long long>(pos);
                        unsigned long long const actual = seedLen;
                        unsigned long long       eff    = static_cast<unsigned long long>(L * P.preScoring);
                        if (eff < actual)
                            eff = actual;
                        unsigned long long const sBase = sbjBase(ix, static_cast<unsigned int>(subj));
                        unsigned long long const sLen  = sbjLength(ix, static_cast<unsigned int>(subj));
                        signed char const *      M     = sM + matrixOffset(ix, static_cast<unsigned int>(subj));
                        if (eff > actual)
                        {
                            qB -= static_cast<long long>((eff - actual) / 2);
                            sB -= static_cast<long long>((eff - actual) / 2);
                            long long const mn = qB < sB ? qB : sB;
                            if (mn < 0)
                            {
                                qB -= mn;
                                sB -= mn;
                                eff += mn;
                            }
                            unsigned long long const qRem = static_cast<unsigned long long>(len) - qB;
                            unsigned long long const sRem = sLen - sB;
                            if (qRem < eff)
                                eff = qRem;
                            if (sRem < eff)
                                eff = sRem;
                        }
                        int const             thresh = static_cast<int>(P.preScoringThresh * static_cast<double>(eff));
                        unsigned char const * qs     = trans + qB;
                        unsigned char const * ss     = ix.seqs + sBase + sB;
                        int                   s = 0, mx = 0;
                        bool                  pass = false;
                        for (unsigned long long i = 0; i < eff; ++i)
                        {
                            s += M[qs[i] * 32 + __ldg(ss + i)];
                            if (s < 0)
                                s = 0;
                            else if (s > mx)
                                mx = s;
                            if (mx >= thresh)
                            {
                                pass = true;
                                break;
                            }
                        }
                        if (!pass)
                        {
                            ++nFailed;
                            continue;
                        }
                        ++hitsThisSeq;
                        unsigned long long const slot = atomicAdd(&P.counters[0], 1ull);
                        if (slot < P.cap)
                        {
                            lgpu_match m;
                            m.qry_id     = qryId;
                            m.subj_id    = static_cast<unsigned int>(subj);
                            m.qry_start  = seedBegin;
                            m.qry_end    = seedBegin + seedLen;
                            m.subj_start = static_cast<unsigned int>(pos);
                            m.subj_end   = static_cast<unsigned int>(pos) + seedLen;
                            P.out[slot]  = m;
                        }
                    }
                }
            }
            needlesPos += len;
        }
    }
    if (nAfter)
        atomicAdd(&P.counters[1], nAfter);
    if (nFailed)
        atomicAdd(&P.counters[2], nFailed);
}

// ---------------------------------------------------------------------------------------------
// warp-cooperative building blocks (shared by the warp-per-query and block-per-query kernels)
// ---------------------------------------------------------------------------------------------

// cursor after the first `n` symbols of the seed starting at seedBegin, through the prefix table where it applies
__device__ __forceinline__ Cursor seedPrefixCursor(SeedParams const & P, unsigned char const * red, unsigned int seedBegin,
                                                   unsigned int n)
{
    DevIndex const & ix = P.ix;
    Cursor           c;
    c.lb           = 0;
    c.len          = ix.nRows;
    unsigned int i = 0;
    if (P.prefixTab && n >= P.prefixK)
    {
        unsigned int       code = 0;
        unsigned int const A    = ix.sigma - 1;
        for (; i < P.prefixK; ++i)
            code = code * A + seedSym(red, seedBegin, seedBegin + i, ix.bsMode != 0);
        uint2 const e = __ldg(P.prefixTab + code);
        c.lb          = e.x;
        c.len         = e.y;
    }
    for (; i < n && c.len != 0; ++i)
        c = fmExtendRight(ix, c, seedSym(red, seedBegin, seedBegin + i, ix.bsMode != 0) + 1u);
    return c;
}

// tab[code] = cursor of the k-mer `code` (base sigma-1 digits, first symbol most significant); len 0 = does not occur
__global__ void prefixTableKernel(DevIndex ix, unsigned int k, unsigned int nEntries, uint2 * tab)
{
    unsigned int const t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= nEntries)
        return;
    unsigned int const A = ix.sigma - 1;
    unsigned int       div = 1;
    for (unsigned int i = 1; i < k; ++i)
        div *= A;
    Cursor c;
    c.lb              = 0;
    c.len             = ix.nRows;
    unsigned int rest = t;
    for (unsigned int i = 0; i < k && c.len != 0; ++i)
    {
        unsigned int const sym = rest / div;
        rest -= sym * div;
        div = div > 1 ? div / A : 1;
        c   = fmExtendRight(ix, c, sym + 1u);
    }
    tab[t] = make_uint2(static_cast<unsigned int>(c.lb), static_cast<unsigned int>(c.len));
}

// Exact chain of one seed: E[0] = cursor after the exactly matched first half, E[i+1] = E[i] extended by
// the (h1+i)-th seed symbol.  Returns the deepest non-empty level K (-1: the first half does not occur).
__device__ __forceinline__ int seedExactChain(SeedParams const & P, unsigned char const * red, unsigned int seedBegin,
                                              unsigned int h1, unsigned int n2, Cursor * E)
{
    DevIndex const & ix = P.ix;
    Cursor           c  = seedPrefixCursor(P, red, seedBegin, h1);
    if (c.len == 0)
        return -1;
    E[0]  = c;
    int K = 0;
    for (unsigned int i = 0; i < n2; ++i)
    {
        c = fmExtendRight(ix, c, seedSym(red, seedBegin, seedBegin + h1 + i, ix.bsMode != 0) + 1u);
        if (c.len == 0)
            break;
        E[i + 1] = c;
        K        = static_cast<int>(i) + 1;
    }
    return K;
}

// Leaf `idx` of the half-exact search tree in the reference's (BFS = leaf) order: mismatches below the
// seed symbol per level going down, the exact cursor, mismatches above the seed symbol going back up.
// levelOrder (max_seed_dist = 1 without seed_half_exact, FMC search/BacktrackingWithBuffers.h:40-83): the
// mismatches of level 0, 1, ... in symbol order are delegated as soon as their error budget is used up, the
// exact cursor is what remains in the buffer at the end.
__device__ __forceinline__ Cursor seedLeaf(DevIndex const & ix, Cursor const * E, unsigned char const * red,
                                           unsigned int seedBegin, unsigned int h1, unsigned int n2, int kMax, bool hasExact,
                                           unsigned int redN, unsigned int idx, bool levelOrder = false)
{
    Cursor mine;
    mine.lb  = 0;
    mine.len = 0;
    unsigned int i     = idx;
    int          lvl   = -1;
    unsigned int r     = 0;
    bool         exact = false;
    if (levelOrder)
    {
        unsigned int const per = redN - 1;
        if (i < static_cast<unsigned int>(kMax + 1) * per)
        {
            lvl                     = static_cast<int>(i / per);
            unsigned int const k    = i % per;
            unsigned int const want = seedSym(red, seedBegin, seedBegin + h1 + lvl, ix.bsMode != 0);
            r                       = k < want ? k : k + 1;
        }
        else if (hasExact)
            exact = true;
    }
    else
    for (int l = 0; l <= kMax; ++l)
    {
        unsigned int const want = seedSym(red, seedBegin, seedBegin + h1 + l, ix.bsMode != 0);
        if (i < want)
        {
            lvl = l;
            r   = i;
            break;
        }
        i -= want;
    }
    if (!levelOrder && lvl < 0 && hasExact)
    {
        if (i == 0)
            exact = true;
        else
            --i;
    }
    if (!levelOrder && lvl < 0 && !exact)
        for (int l = kMax; l >= 0; --l)
        {
            unsigned int const want = seedSym(red, seedBegin, seedBegin + h1 + l, ix.bsMode != 0);
            unsigned int const cnt  = redN - 1 - want;
            if (i < cnt)
            {
                lvl = l;
                r   = want + 1 + i;
                break;
            }
            i -= cnt;
        }
    if (exact)
        mine = E[n2];
    else if (lvl >= 0)
    {
        Cursor c = fmExtendRight(ix, E[lvl], r + 1u);
        for (unsigned int l = lvl + 1; c.len != 0 && l < n2; ++l)
            c = fmExtendRight(ix, c, seedSym(red, seedBegin, seedBegin + h1 + l, ix.bsMode != 0) + 1u);
        mine = c;
    }
    return mine;
}

// ---- seeds with up to TWO mismatches (max_seed_dist = 2) ---------------------------------------------------------
// number of strings of length m within Hamming distance <= b (b <= 2) of a given one over an alphabet of A symbols
__device__ __forceinline__ unsigned int hammingBall(unsigned int m, unsigned int b, unsigned int A)
{
    unsigned int n = 1;
    if (b >= 1)
        n += m * (A - 1);
    if (b >= 2)
        n += m * (m - 1) / 2 * (A - 1) * (A - 1);
    return n;
}

// How many leaves the enumeration below produces for one seed (empty ones included: a lane finds out by walking).
__device__ __forceinline__ unsigned int seedNumLeaves2(unsigned int n2, unsigned int A, bool levelOrder)
{
    (void) levelOrder; // both orders enumerate the whole ball
    return hammingBall(n2, 2, A);
}

// the mismatch (level, symbol) number `i` of the "down then back up" order over levels 0 .. nLevels-1 WITHOUT the
// exact string: symbols below the seed symbol per level going down, then symbols above it per level going back up
__device__ __forceinline__ void seedOneMismatch(unsigned char const * red, unsigned int seedBegin, unsigned int h1, bool bs,
                                                unsigned int nLevels, unsigned int A, unsigned int i, int & lvl, unsigned int & r)
{
    lvl = -1;
    r   = 0;
    for (unsigned int l = 0; l < nLevels; ++l)
    {
        unsigned int const want = seedSym(red, seedBegin, seedBegin + h1 + l, bs);
        if (i < want)
        {
            lvl = static_cast<int>(l);
            r   = i;
            return;
        }
        i -= want;
    }
    for (int l = static_cast<int>(nLevels) - 1; l >= 0; --l)
    {
        unsigned int const want = seedSym(red, seedBegin, seedBegin + h1 + l, bs);
        unsigned int const cnt  = A - 1 - want;
        if (i < cnt)
        {
            lvl = l;
            r   = want + 1 + i;
            return;
        }
        i -= cnt;
    }
}

// Leaf `idx` of the search tree of a seed part of n2 symbols with up to two mismatches, in the reference's production
// order; E[l] = cursor of the exactly matched first l symbols (levels 0 .. K non-empty).
//   levelOrder = false (seed_half_exact, src/search_algo.hpp:538-604): breadth-first with ordered children = the strings
//     of the Hamming ball in LEXICOGRAPHIC order; unranked position by position with the ball sizes.
//   levelOrder = true (FMC search/BacktrackingWithBuffers.h:40-83): a branch is delegated to the exact search as soon
//     as its error budget is used up, i.e. at the level of its SECOND mismatch: level by level, for every buffered
//     one-mismatch prefix (they sit in the buffer in lexicographic order) the mismatching symbols in order; what is
//     left in the buffer at the end -- the strings with at most one mismatch, lexicographically -- comes last.
__device__ __forceinline__ Cursor seedLeaf2(DevIndex const & ix, Cursor const * E, int K, unsigned char const * red,
                                            unsigned int seedBegin, unsigned int h1, unsigned int n2, unsigned int A,
                                            unsigned int idx, bool levelOrder)
{
    bool const bs = ix.bsMode != 0;
    Cursor     none;
    none.lb  = 0;
    none.len = 0;
    int          l1 = -1, l2 = -1; // mismatch levels (l1 < l2), -1 = none
    unsigned int r1 = 0, r2 = 0;
    if (!levelOrder)
    {
        unsigned int b = 2, i = idx;
        for (unsigned int pos = 0; pos < n2 && b > 0; ++pos)
        {
            unsigned int const m     = n2 - pos - 1;
            unsigned int const want  = seedSym(red, seedBegin, seedBegin + h1 + pos, bs);
            unsigned int const sub   = hammingBall(m, b - 1, A); // leaves below one mismatching symbol here
            unsigned int const below = want * sub;
            unsigned int       sym   = want;
            if (i < below)
            {
                sym = i / sub;
                i -= sym * sub;
            }
            else
            {
                i -= below;
                unsigned int const exact = hammingBall(m, b, A);
                if (i >= exact)
                {
                    i -= exact;
                    sym = want + 1 + i / sub;
                    i %= sub;
                }
            }
            if (sym != want)
            {
                if (l1 < 0) { l1 = static_cast<int>(pos); r1 = sym; }
                else        { l2 = static_cast<int>(pos); r2 = sym; }
                --b;
            }
        }
    }
    else
    {
        unsigned int const per2  = (A - 1) * (A - 1);
        unsigned int const nTwo  = n2 * (n2 - 1) / 2 * per2;
        if (idx < nTwo)
        {
            // second mismatch at level i: i * (A-1) one-mismatch prefixes, A-1 symbols each
            unsigned int i = 1, t = idx;
            while (t >= i * per2)
            {
                t -= i * per2;
                ++i;
            }
            unsigned int const p    = t / (A - 1), k = t % (A - 1);
            unsigned int const want = seedSym(red, seedBegin, seedBegin + h1 + i, bs);
            l2                      = static_cast<int>(i);
            r2                      = k < want ? k : k + 1;
            seedOneMismatch(red, seedBegin, h1, bs, i, A, p, l1, r1);
        }
        else
        {
            // at most one mismatch, lexicographically: one-mismatch strings below the seed, the seed, those above
            unsigned int       t     = idx - nTwo;
            unsigned int       below = 0;
            for (unsigned int l = 0; l < n2; ++l)
                below += seedSym(red, seedBegin, seedBegin + h1 + l, bs);
            if (t < below)
                seedOneMismatch(red, seedBegin, h1, bs, n2, A, t, l1, r1);
            else if (t > below)
                seedOneMismatch(red, seedBegin, h1, bs, n2, A, t - 1, l1, r1);
        }
    }
    // walk: exact prefix up to the first mismatch comes from the chain
    int const first = l1 >= 0 ? l1 : static_cast<int>(n2);
    if (first > K)
        return none; // the exactly matched prefix does not occur
    if (l1 < 0)
        return E[n2];
    Cursor c = fmExtendRight(ix, E[l1], r1 + 1u);
    for (unsigned int l = static_cast<unsigned int>(l1) + 1; c.len != 0 && l < n2; ++l)
    {
        unsigned int const sym = (static_cast<int>(l) == l2) ? r2 : seedSym(red, seedBegin, seedBegin + h1 + l, bs);
        c                      = fmExtendRight(ix, c, sym + 1u);
    }
    return c;
}

// seedLooksPromising (src/search_algo.hpp:427-481): ungapped running-maximum score on the seed diagonal
// over max(seedLength * preScoring, seedLen) residues centred on the seed, clipped to both sequences.
__device__ __forceinline__ bool seedPreScore(SeedParams const & P, signed char const * sM, unsigned char const * trans,
                                             unsigned int len, unsigned int seedBegin, unsigned int seedLen,
                                             unsigned long long subj, unsigned long long pos)
{
    DevIndex const &         ix     = P.ix;
    long long                qB     = seedBegin;
    long long                sB     = static_cast<long long>(pos);
    unsigned long long const actual = seedLen;
    unsigned long long       eff    = static_cast<unsigned long long>(P.seedLength * P.preScoring);
    if (eff < actual)
        eff = actual;
    unsigned long long const sBase = sbjBase(ix, static_cast<unsigned int>(subj));
    unsigned long long const sLen  = sbjLength(ix, static_cast<unsigned int>(subj));
    signed char const *      M     = sM + matrixOffset(ix, static_cast<unsigned int>(subj));
    if (eff > actual)
    {
        qB -= static_cast<long long>((eff - actual) / 2);
        sB -= static_cast<long long>((eff - actual) / 2);
        long long const mn = qB < sB ? qB : sB;
        if (mn < 0)
        {
            qB -= mn;
            sB -= mn;
            eff += mn;
        }
        unsigned long long const qRem = static_cast<unsigned long long>(len) - qB;
        unsigned long long const sRem = sLen - sB;
        if (qRem < eff)
            eff = qRem;
        if (sRem < eff)
            eff = sRem;
    }
    int const             thresh = static_cast<int>(P.preScoringThresh * static_cast<double>(eff));
    unsigned char const * qs     = trans + qB;
    unsigned char const * ss     = ix.seqs + sBase + sB;
    int                   sc = 0, mx = 0;
    for (unsigned long long i = 0; i < eff; ++i)
    {
        sc += M[qs[i] * 32 + __ldg(ss + i)];
        if (sc < 0)
            sc = 0;
        else if (sc > mx)
            mx = sc;
        if (mx >= thresh)
            return true;
    }
    return false;
}

// desiredOccs of the adaptive seed elongation (src/search_algo.hpp:683-701)
__device__ __forceinline__ unsigned long long seedDesiredOccs(SeedParams const & P, unsigned long long hitsThisSeq,
                                                              unsigned long long needlesSum, unsigned long long needlesPos,
                                                              unsigned int seedBegin)
{
    unsigned long long desired = 1;
    if (hitsThisSeq < P.maxMatches)
    {
        unsigned long long remaining = (needlesSum - needlesPos - seedBegin) / P.seedOffset;
        if (remaining < 1)
            remaining = 1;
        desired = (P.maxMatches - hitsThisSeq) * 10ull / remaining;
        if (desired == 0)
            desired = 1;
    }
    return desired;
}

// locate + pre-scoring of all occurrences of one (already elongated) cursor, 32 rows at a time
__device__ __forceinline__ void seedConsumeRows(SeedParams const & P, signed char const * sM, unsigned int lane,
                                                unsigned int qryId, unsigned char const * trans, unsigned int len,
                                                unsigned int seedBegin, Cursor cursor, unsigned int seedLen,
                                                unsigned long long & hitsThisSeq, unsigned long long & nAfter,
                                                unsigned long long & nFailed)
{
    DevIndex const &   ix     = P.ix;
    unsigned int const ltMask = (1u << lane) - 1u;
    for (unsigned long long row0 = 0; row0 < cursor.len; row0 += 32)
    {
        bool const         act  = row0 + lane < cursor.len;
        bool               pass = false;
        unsigned long long subj = 0, pos = 0;
        if (act)
        {
            fmLocate(ix, cursor.lb + row0 + lane, subj, pos);
            pos -= seedLen;
            ++nAfter;
            pass = seedPreScore(P, sM, trans, len, seedBegin, seedLen, subj, pos);
            if (!pass)
                ++nFailed;
        }
        unsigned int const passMask = __ballot_sync(0xffffffffu, pass);
        unsigned int const cnt      = __popc(passMask);
        if (cnt)
        {
            unsigned long long slot0 = 0;
            if (lane == 0)
                slot0 = atomicAdd(&P.counters[0], static_cast<unsigned long long>(cnt));
            slot0 = __shfl_sync(0xffffffffu, slot0, 0);
            if (pass)
            {
                unsigned long long const slot = slot0 + __popc(passMask & ltMask);
                if (slot < P.cap)
                {
                    lgpu_match m;
                    m.qry_id     = qryId;
                    m.subj_id    = static_cast<unsigned int>(subj);
                    m.qry_start  = seedBegin;
                    m.qry_end    = seedBegin + seedLen;
                    m.subj_start = static_cast<unsigned int>(pos);
                    m.subj_end   = static_cast<unsigned int>(pos) + seedLen;
                    P.out[slot]  = m;
                }
            }
            hitsThisSeq += cnt;
        }
    }
}

// Everything the reference does with one cursor of one seed (src/search_algo.hpp:674-757), executed by a
// full warp: adaptive elongation (uniform), abundance cut, then locate + pre-scoring 32 rows at a time.
__device__ __forceinline__ void seedConsumeCursor(SeedParams const & P, signed char const * sM, unsigned int lane,
                                                  unsigned int qryId, unsigned char const * trans, unsigned char const * red,
                                                  unsigned int len, unsigned int seedBegin, Cursor cursor,
                                                  unsigned long long & hitsThisSeq, unsigned long long needlesSum,
                                                  unsigned long long needlesPos, unsigned long long & nAfter,
                                                  unsigned long long & nFailed)
{
    DevIndex const & ix      = P.ix;
    unsigned int     seedLen = P.seedLength;
    if (P.adaptive)
    {
        unsigned long long const desired  = seedDesiredOccs(P, hitsThisSeq, needlesSum, needlesPos, seedBegin);
        unsigned long long       oldCount = cursor.len;
        while (seedBegin + seedLen < len)
        {
            Cursor const n = fmExtendRight(ix, cursor, elongSym(red, seedBegin + seedLen, ix.bsMode != 0) + 1u);
            if (n.len < desired && n.len < oldCount)
                break;
            cursor   = n;
            oldCount = n.len;
            ++seedLen;
        }
    }
    if (cursor.len > 10ull * P.maxMatches)
        return;
    seedConsumeRows(P, sM, lane, qryId, trans, len, seedBegin, cursor, seedLen, hitsThisSeq, nAfter, nFailed);
}

// ---------------------------------------------------------------------------------------------
// speculative consumption of up to 32 cursors at once
// ---------------------------------------------------------------------------------------------
// The reference consumes the cursors of a query strictly one after the other because the adaptive
// elongation of cursor i looks at hitsThisSeq, the number of hits all earlier cursors produced
// (src/search_algo.hpp:683-703).  That number only enters through `desired`, and `desired` only decides
// the elongation at steps where the occurrence count drops to a non-zero value.  So every lane elongates
// its own cursor with the hitsThisSeq known at the start of the round and remembers for which interval
// [lo, hi] of `desired` it would have taken exactly the same decisions; the occurrences of all lanes are
// located and pre-scored together (flattened, 32 rows at a time, results parked in shared memory); a warp
// scan then yields the true hitsThisSeq in front of every lane, and the longest prefix of lanes whose true
// `desired` lies inside their interval is committed.  The first lane of a round is never speculative, so
// every round commits at least one cursor; the others are retried.  Results are identical to the serial
// order; the dependent-load chain of a query shrinks by up to 32x.
constexpr int kSpecRowCap = 256; // rows (occurrences) one round can park; 10 * maxMatches = 250 by default

struct SpecScratch
{
    unsigned int  subj[kSpecRowCap];
    unsigned int  pos[kSpecRowCap];
    unsigned int  cnt[32];
    unsigned int  passBits[kSpecRowCap / 32];
    unsigned char owner[kSpecRowCap];
};

__device__ __forceinline__ unsigned int warpInclusiveScan(unsigned int v, unsigned int lane)
{
#pragma unroll
    for (int off = 1; off < 32; off <<= 1)
    {
        unsigned int const n = __shfl_up_sync(0xffffffffu, v, off);
        if (lane >= static_cast<unsigned int>(off))
            v += n;
    }
    return v;
}

// `mine` (len 0 = none) is this lane's cursor of the seed starting at `seedBegin` in frame `f`; lanes are
// in the reference's consumption order.  hitsThisSeq stays uniform across the warp.
__device__ __forceinline__ void seedConsumeChunkSpec(SeedParams const & P, signed char const * sM, SpecScratch & S,
                                                     unsigned int lane, unsigned int q, unsigned long long qb,
                                                     unsigned int origLen, unsigned long long needlesSum, Cursor mine,
                                                     unsigned int f, unsigned int seedBegin, unsigned long long needlesPosF,
                                                     unsigned long long & hitsThisSeq, unsigned long long & nAfter,
                                                     unsigned long long & nFailed)
{
    DevIndex const &   ix     = P.ix;
    unsigned int const F      = P.Q.F;
    unsigned int const L      = P.seedLength;
    unsigned int const ltMask = (1u << lane) - 1u;
    unsigned int       pending = __ballot_sync(0xffffffffu, mine.len != 0);
    if (!pending)
        return;
    unsigned int const    len = qryFrameLen(P.Q, origLen, f);
    unsigned char const * red = P.Q.red + F * qb + static_cast<unsigned long long>(f) * origLen;

    bool               haveEl = false;
    Cursor             el     = mine;
    unsigned int       elLen  = L;
    unsigned long long lo = 1, hi = ~0ull;
    // text mode: the cursor has ONE occurrence, already located: subject tSubj, seed start tStart
    bool               elText = false;
    unsigned int       tSubj = 0, tStart = 0;
    while (pending)
    {
        unsigned long long const hits0     = hitsThisSeq;
        bool const               isPending = (pending >> lane) & 1u;
        if (isPending)
        {
            unsigned long long const desired =
              P.adaptive ? seedDesiredOccs(P, hits0, needlesSum, needlesPosF, seedBegin) : 1ull;
            if (!haveEl || desired < lo || desired > hi)
            {
                el     = mine;
                elLen  = L;
                lo     = 1;
                hi     = ~0ull;
                elText = false;
                if (P.adaptive)
                {
                    unsigned long long oldCount = el.len;
                    bool               stopped  = false;
                    // LF steps while the cursor has several occurrences (all of them, if text mode is off)
                    while (seedBegin + elLen < len && (el.len > 1 || !P.textElong))
                    {
                        Cursor const n = fmExtendRight(ix, el, elongSym(red, seedBegin + elLen, ix.bsMode != 0) + 1u);
                        if (n.len < oldCount)
                        {
                            if (n.len < desired)
                            {
                                lo      = max(lo, n.len + 1); // same stop for every desired > n.len
                                stopped = true;
                                break;
                            }
                            hi = min(hi, n.len); // same continuation for every desired <= n.len
                        }
                        el       = n;
                        oldCount = n.len;
                        ++elLen;
                    }
                    if (!stopped && seedBegin + elLen < len)
                    {
                        // ONE occurrence left: every further LF step would only test whether the next subject residue
                        // equals the next query symbol (count 1 -> 1: go on; 1 -> 0 < desired: stop, `lo` stays >= 1).
                        // Locate it once and compare the texts instead.
                        unsigned long long subj, pos;
                        fmLocate(ix, el.lb, subj, pos); // pos = one past the seed's end in the subject
                        unsigned long long const sBase  = sbjBase(ix, static_cast<unsigned int>(subj));
                        unsigned long long const sLen   = sbjLength(ix, static_cast<unsigned int>(subj));
                        unsigned char const *    rt     = P.sbjRed[subj & ix.bsMode];
                        unsigned int const       len0   = elLen;
                        bool                     replay = false;
                        while (seedBegin + elLen < len)
                        {
                            unsigned long long const sp = pos + (elLen - len0);
                            if (sp >= sLen)
                                break; // the sentinel behind the subject matches no query symbol
                            unsigned int const rs = rt[__ldg(ix.seqs + sBase + sp)];
                            if (rs == 0xffu)
                            {
                                replay = true; // a residue whose index symbol this table does not know
                                break;
                            }
                            if (rs != elongSym(red, seedBegin + elLen, ix.bsMode != 0))
                                break;
                            ++elLen;
                        }
                        if (!replay)
                        {
                            elText = true;
                            tSubj  = static_cast<unsigned int>(subj);
                            tStart = static_cast<unsigned int>(pos - len0);
                        }
                        else
                        {
                            // rare: redo the matched stretch with LF steps and go on as the reference does
                            for (unsigned int t = len0; t < elLen; ++t)
                                el = fmExtendRight(ix, el, elongSym(red, seedBegin + t, ix.bsMode != 0) + 1u);
                            oldCount = el.len;
                            while (seedBegin + elLen < len)
                            {
                                Cursor const n = fmExtendRight(ix, el, elongSym(red, seedBegin + elLen, ix.bsMode != 0) + 1u);
                                if (n.len < oldCount)
                                {
                                    if (n.len < desired)
                                    {
                                        lo = max(lo, n.len + 1);
                                        break;
                                    }
                                    hi = min(hi, n.len);
                                }
                                el       = n;
                                oldCount = n.len;
                                ++elLen;
                            }
                        }
                    }
                }
                haveEl = true;
            }
        }
        __syncwarp();
        // over-abundant cursors are dropped (src/search_algo.hpp:729): no rows, but still validated in order
        unsigned int const rows = (isPending && el.len <= 10ull * P.maxMatches) ? static_cast<unsigned int>(el.len) : 0u;
        unsigned int const incl = warpInclusiveScan(rows, lane);
        int const          head = __ffs(pending) - 1;
        unsigned int const base0 = __shfl_sync(0xffffffffu, incl - rows, head);
        bool const         inChunk   = isPending && (incl - base0 <= static_cast<unsigned int>(kSpecRowCap));
        unsigned int const chunkMask = __ballot_sync(0xffffffffu, inChunk);
        if (!((chunkMask >> head) & 1u))
        {
            // the head alone has more rows than a round can park; it is not speculative -> emit directly
            Cursor hc;
            hc.lb                      = __shfl_sync(0xffffffffu, el.lb, head);
            hc.len                     = __shfl_sync(0xffffffffu, el.len, head);
            unsigned int const hLen    = __shfl_sync(0xffffffffu, elLen, head);
            unsigned int const hf      = __shfl_sync(0xffffffffu, f, head);
            unsigned int const hsb     = __shfl_sync(0xffffffffu, seedBegin, head);
            unsigned char const * htr  = P.Q.trans + F * qb + static_cast<unsigned long long>(hf) * origLen;
            seedConsumeRows(P, sM, lane, q * F + hf, htr, qryFrameLen(P.Q, origLen, hf), hsb, hc, hLen, hitsThisSeq, nAfter,
                            nFailed);
            pending &= ~(1u << head);
            continue;
        }
        int const          last = 31 - __clz(chunkMask);
        unsigned int const R    = __shfl_sync(0xffffffffu, incl, last) - base0;
        S.cnt[lane]             = 0;
        __syncwarp();
        for (unsigned int r0 = 0; r0 < R; r0 += 32)
        {
            unsigned int const r      = r0 + lane;
            bool const         act    = r < R;
            unsigned int const target = base0 + (act ? r : 0u);
            // owner = first lane whose inclusive row count exceeds the row number
            int j = 0;
#pragma unroll
            for (int step = 16; step > 0; step >>= 1)
            {
                unsigned int const v = __shfl_sync(0xffffffffu, incl, j + step - 1);
                if (v <= target)
                    j += step;
            }
            unsigned int const       oExcl = __shfl_sync(0xffffffffu, incl - rows, j);
            unsigned long long const oLb   = __shfl_sync(0xffffffffu, el.lb, j);
            unsigned int const       oLen  = __shfl_sync(0xffffffffu, elLen, j);
            unsigned int const       oSb   = __shfl_sync(0xffffffffu, seedBegin, j);
            unsigned int const       oF    = __shfl_sync(0xffffffffu, f, j);
            bool const               oText = __shfl_sync(0xffffffffu, elText ? 1 : 0, j) != 0;
            unsigned int const       oTS   = __shfl_sync(0xffffffffu, tSubj, j);
            unsigned int const       oTP   = __shfl_sync(0xffffffffu, tStart, j);
            bool                     pass  = false;
            if (act)
            {
                unsigned long long subj = oTS, pos = oTP;
                if (!oText)
                {
                    fmLocate(ix, oLb + (target - oExcl), subj, pos);
                    pos -= oLen;
                }
                unsigned char const * tr = P.Q.trans + F * qb + static_cast<unsigned long long>(oF) * origLen;
                pass       = seedPreScore(P, sM, tr, qryFrameLen(P.Q, origLen, oF), oSb, oLen, subj, pos);
                S.subj[r]  = static_cast<unsigned int>(subj);
                S.pos[r]   = static_cast<unsigned int>(pos);
                S.owner[r] = static_cast<unsigned char>(j);
                if (pass)
                    atomicAdd(&S.cnt[j], 1u);
            }
            unsigned int const pm = __ballot_sync(0xffffffffu, pass);
            if (lane == 0)
                S.passBits[r0 >> 5] = pm;
        }
        __syncwarp();
        // ---- validation: true hitsThisSeq in front of every lane ----
        unsigned int const myPass   = inChunk ? S.cnt[lane] : 0u;
        unsigned int const inclPass = warpInclusiveScan(myPass, lane);
        bool               ok       = true;
        if (inChunk && P.adaptive)
        {
            unsigned long long const d = seedDesiredOccs(P, hits0 + inclPass - myPass, needlesSum, needlesPosF, seedBegin);
            ok                         = d >= lo && d <= hi;
        }
        unsigned int const bad        = __ballot_sync(0xffffffffu, !ok) & chunkMask & ~(1u << head);
        unsigned int const commitMask = bad ? (chunkMask & ((1u << (__ffs(bad) - 1)) - 1u)) : chunkMask;
        int const          lastC      = 31 - __clz(commitMask);
        unsigned int const Rc         = __shfl_sync(0xffffffffu, incl, lastC) - base0;
        unsigned int const hitsAdd    = __shfl_sync(0xffffffffu, inclPass, lastC);
        if (hitsAdd)
        {
            unsigned long long slot0 = 0;
            if (lane == 0)
                slot0 = atomicAdd(&P.counters[0], static_cast<unsigned long long>(hitsAdd));
            slot0            = __shfl_sync(0xffffffffu, slot0, 0);
            unsigned int run = 0;
            for (unsigned int r0 = 0; r0 < Rc; r0 += 32)
            {
                unsigned int const r  = r0 + lane;
                unsigned int       pm = S.passBits[r0 >> 5];
                if (Rc - r0 < 32)
                    pm &= (1u << (Rc - r0)) - 1u;
                int const          j    = (r < Rc) ? static_cast<int>(S.owner[r]) : head;
                unsigned int const oLen = __shfl_sync(0xffffffffu, elLen, j);
                unsigned int const oSb  = __shfl_sync(0xffffffffu, seedBegin, j);
                unsigned int const oF   = __shfl_sync(0xffffffffu, f, j);
                if ((pm >> lane) & 1u)
                {
                    unsigned long long const slot = slot0 + run + __popc(pm & ltMask);
                    if (slot < P.cap)
                    {
                        lgpu_match m;
                        m.qry_id     = q * F + oF;
                        m.subj_id    = S.subj[r];
                        m.qry_start  = oSb;
                        m.qry_end    = oSb + oLen;
                        m.subj_start = S.pos[r];
                        m.subj_end   = S.pos[r] + oLen;
                        P.out[slot]  = m;
                    }
                }
                run += __popc(pm);
            }
        }
        if (lane == 0)
        {
            nAfter += Rc;
            nFailed += Rc - hitsAdd;
        }
        hitsThisSeq += hitsAdd;
        pending &= ~commitMask;
        __syncwarp();
    }
}

__device__ __forceinline__ void seedFlushCounters(SeedParams const & P, unsigned int lane, unsigned long long nAfter,
                                                  unsigned long long nFailed)
{
#pragma unroll
    for (int off = 16; off > 0; off >>= 1)
    {
        nAfter += __shfl_down_sync(0xffffffffu, nAfter, off);
        nFailed += __shfl_down_sync(0xffffffffu, nFailed, off);
    }
    if (lane == 0)
    {
        if (nAfter)
            atomicAdd(&P.counters[1], nAfter);
        if (nFailed)
            atomicAdd(&P.counters[2], nFailed);
    }
}

// next seed start at or after `seedBegin` (src/search_algo.hpp:652-656); false = no more seeds
__device__ __forceinline__ bool seedNextStart(unsigned char const * trans, unsigned int len, unsigned int L,
                                              unsigned int unknownRank, unsigned int & seedBegin)
{
    while (seedBegin < len - L && (trans[seedBegin] == unknownRank || trans[seedBegin] == trans[seedBegin + 1]))
        ++seedBegin;
    return seedBegin <= len - L;
}

// The same for a whole warp at once (all lanes call it with the same arguments): the "skip this start" test of 32
// positions is one coalesced load and a ballot, kept in `bits` for the word `word` of the frame (callers reset
// word = ~0u when the frame changes); finding the next start is then bit arithmetic instead of a loop of dependent
// byte loads that every lane repeats.
__device__ __forceinline__ bool seedNextStartWarp(unsigned char const * trans, unsigned int len, unsigned int L,
                                                  unsigned int unknownRank, unsigned int & seedBegin, unsigned int lane,
                                                  unsigned int & word, unsigned int & bits)
{
    unsigned int const last = len - L; // the last possible start is taken without the test
    for (;;)
    {
        if (seedBegin >= last)
            return seedBegin <= last;
        unsigned int const w = seedBegin >> 5;
        if (w != word)
        {
            unsigned int const p    = (w << 5) + lane;
            bool               skip = false;
            if (p < last)
            {
                unsigned int const a = trans[p];
                skip                 = a == unknownRank || a == trans[p + 1];
            }
            bits = __ballot_sync(0xffffffffu, skip);
            word = w;
        }
        unsigned int const off  = seedBegin & 31u;
        unsigned int const keep = ~(bits >> off);                 // 1 = a start that is not skipped (bits beyond the word: 1)
        unsigned int const k    = static_cast<unsigned int>(__ffs(keep)) - 1u; // keep != 0: the shift fills with zeros unless off = 0
        if (keep != 0 && k < 32u - off)
        {
            seedBegin += k;
            return seedBegin <= last;
        }
        seedBegin = (w + 1u) << 5;
    }
}

// ---------------------------------------------------------------------------------------------
// seeding, one WARP per query
// ---------------------------------------------------------------------------------------------
// Same semantics and emission rules as seedKernel, but the work inside one query is spread over the
// 32 lanes wherever the reference's order does not forbid it:
//   * the half-exact search tree of a seed (up to (sigma-1) * L/2 mismatch branches, each a chain of
//     dependent LF steps) is evaluated 32 branches at a time, then consumed in the reference's order;
//   * the occurrences of a cursor are located and pre-scored 32 rows at a time.
// Only the exact chain, the adaptive elongation and the running `hitsThisSeq` stay serial.
__global__ void __launch_bounds__(128) seedWarpKernel(SeedParams P)
{
    __shared__ signed char sM[2048];
    for (int i = threadIdx.x; i < 2048; i += blockDim.x)
        sM[i] = P.matrix[i];
    __syncthreads();

    unsigned int const lane = threadIdx.x & 31u;
    unsigned int const w    = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (w >= P.nActive)
        return;
    unsigned int const       q    = P.active[w];
    DevIndex const &         ix   = P.ix;
    unsigned int const       F    = P.Q.F;
    unsigned long long const qb      = P.Q.offs[q];
    unsigned int const       origLen = static_cast<unsigned int>(P.Q.offs[q + 1] - qb);
    unsigned int const       L       = P.seedLength;
    unsigned int const       redN    = ix.sigma - 1;

    unsigned long long nAfter = 0, nFailed = 0; // per lane
    if (qryFrameLen(P.Q, origLen, 0) >= L)
    {
        unsigned long long hitsThisSeq = 0;
        unsigned long long needlesSum  = 0;
        for (unsigned int f = 0; f < F; ++f)
            needlesSum += qryFrameLen(P.Q, origLen, f);
        unsigned long long needlesPos = 0;
        bool const         half       = (P.halfExact && P.maxSeedDist != 0) || P.fullHamming;
        unsigned int const h1         = P.fullHamming ? 0u : (half ? L / 2 : L);
        unsigned int const n2         = L - h1;

        for (unsigned int f = 0; f < F; ++f)
        {
            unsigned int const len = qryFrameLen(P.Q, origLen, f);
            if (len < L)
                continue;
            unsigned int const    qryId = q * F + f;
            unsigned char const * trans = P.Q.trans + F * qb + static_cast<unsigned long long>(f) * origLen;
            unsigned char const * red   = P.Q.red + F * qb + static_cast<unsigned long long>(f) * origLen;

            for (unsigned int seedBegin = 0;; seedBegin += P.seedOffset)
            {
                if (!seedNextStart(trans, len, L, P.unknownRank, seedBegin))
                    break;
                // exact chain, computed redundantly by all lanes (same addresses -> one transaction)
                Cursor    E[kMaxHalf2 + 1];
                int const K = seedExactChain(P, red, seedBegin, h1, n2, E);
                if (K < 0)
                    continue;
                int const          kMax     = (K < static_cast<int>(n2) - 1) ? K : static_cast<int>(n2) - 1;
                bool const         hasExact = (K == static_cast<int>(n2));
                unsigned int const nLeaves  = P.maxSeedDist >= 2 ? seedNumLeaves2(n2, redN, P.fullHamming != 0)
                                                                  : static_cast<unsigned int>(kMax + 1) * (redN - 1) + (hasExact ? 1u : 0u);
                for (unsigned int base = 0; base < nLeaves; base += 32)
                {
                    Cursor mine;
                    mine.lb  = 0;
                    mine.len = 0;
                    if (base + lane < nLeaves)
                        mine = P.maxSeedDist >= 2 ? seedLeaf2(ix, E, K, red, seedBegin, h1, n2, redN, base + lane, P.fullHamming != 0)
                                                      : seedLeaf(ix, E, red, seedBegin, h1, n2, kMax, hasExact, redN, base + lane, P.fullHamming != 0);
                    unsigned int live = __ballot_sync(0xffffffffu, mine.len != 0);
                    while (live)
                    {
                        int const src = __ffs(live) - 1;
                        live &= live - 1;
                        Cursor cursor;
                        cursor.lb  = __shfl_sync(0xffffffffu, mine.lb, src);
                        cursor.len = __shfl_sync(0xffffffffu, mine.len, src);
                        seedConsumeCursor(P, sM, lane, qryId, trans, red, len, seedBegin, cursor, hitsThisSeq, needlesSum,
                                          needlesPos, nAfter, nFailed);
                    }
                }
            }
            needlesPos += len;
        }
    }
    seedFlushCounters(P, lane, nAfter, nFailed);
}

// ---------------------------------------------------------------------------------------------
// seeding, one WARP per query, cursors consumed speculatively 32 at a time (the default)
// ---------------------------------------------------------------------------------------------
// Exact seeds (phase 1): the 32 lanes take 32 consecutive seeds of the query, each lane walks its own
// exact chain (all lanes issue their LF steps together: 32 independent block reads per instruction
// instead of 32 threads diverged over LF / locate / pre-scoring), then seedConsumeChunkSpec.
// Half-exact seeds (phase 2): per seed the exact chain is computed by all lanes, the leaves of the
// mismatch tree are evaluated 32 at a time and handed to seedConsumeChunkSpec in leaf order.
constexpr int kSpecWarps = 4;

// Resident blocks per SM: the kernel is bound by the latency of dependent random reads, so warps in flight count more
// than registers.  8 blocks (32 warps, 64 registers, ~0.5 KB of spills that stay in L1) measured 17.6 / 18.0 ms of
// seeding on searchn / searchbs against 20.1 / 20.1 ms with 6 blocks (80 registers) and 23.0 / 23.0 ms with 5 blocks
// (96 registers, no spills) -- profiles/r2_spec_occupancy_ab.json.
#ifndef LGPU_SPEC_MINBLOCKS
#define LGPU_SPEC_MINBLOCKS 8
#endif
__global__ void __launch_bounds__(32 * kSpecWarps, LGPU_SPEC_MINBLOCKS) seedSpecKernel(SeedParams P)
{
    __shared__ signed char sM[2048];
    __shared__ SpecScratch sS[kSpecWarps];
    for (int i = threadIdx.x; i < 2048; i += blockDim.x)
        sM[i] = P.matrix[i];
    __syncthreads();

    unsigned int const lane = threadIdx.x & 31u;
    unsigned int const w    = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (w >= P.nActive)
        return;
    SpecScratch &            S       = sS[threadIdx.x >> 5];
    unsigned int const       q       = P.active[w];
    DevIndex const &         ix      = P.ix;
    unsigned int const       F       = P.Q.F;
    unsigned long long const qb      = P.Q.offs[q];
    unsigned int const       origLen = static_cast<unsigned int>(P.Q.offs[q + 1] - qb);
    unsigned int const       L       = P.seedLength;
    unsigned int const       redN    = ix.sigma - 1;

    unsigned long long nAfter = 0, nFailed = 0; // per lane
    if (qryFrameLen(P.Q, origLen, 0) >= L)
    {
        unsigned long long hitsThisSeq = 0;
        unsigned long long needlesSum  = 0;
        for (unsigned int f = 0; f < F; ++f)
            needlesSum += qryFrameLen(P.Q, origLen, f);
        unsigned long long needlesPos = 0;
        bool const         half       = (P.halfExact && P.maxSeedDist != 0) || P.fullHamming;
        unsigned int const h1         = P.fullHamming ? 0u : (half ? L / 2 : L);
        unsigned int const n2         = L - h1;

        if (!half)
        {
            // Exact seeds: the 32 lanes take the next 32 seeds of the query in the reference's order -- across frame
            // borders, so that short reads (9 - 16 seeds per frame) still fill the warp.  Every lane runs the same
            // scan and keeps its own (frame, seed start, needlesPos of that frame).
            unsigned int       f = 0, seedBegin = 0;
            unsigned int       scanWord = ~0u, scanBits = 0; // cached "skip" bits of 32 start positions (seedNextStartWarp)
            unsigned long long npos = 0;
            bool               more = true;
            while (more)
            {
                unsigned int       mySeed = 0xffffffffu, myF = 0;
                unsigned long long myNpos = 0;
                for (unsigned int t = 0; t < 32 && more; ++t)
                {
                    for (;;)
                    {
                        if (f >= F)
                        {
                            more = false;
                            break;
                        }
                        unsigned int const len = qryFrameLen(P.Q, origLen, f);
                        if (len < L)
                        {
                            ++f; // too short a frame is skipped without advancing needlesPos (search_algo.hpp:637)
                            seedBegin = 0;
                            scanWord  = ~0u;
                            continue;
                        }
                        unsigned char const * trans = P.Q.trans + F * qb + static_cast<unsigned long long>(f) * origLen;
                        if (seedNextStartWarp(trans, len, L, P.unknownRank, seedBegin, lane, scanWord, scanBits))
                            break;
                        npos += len;
                        ++f;
                        seedBegin = 0;
                        scanWord  = ~0u;
                    }
                    if (!more)
                        break;
                    if (lane == t)
                    {
                        mySeed = seedBegin;
                        myF    = f;
                        myNpos = npos;
                    }
                    seedBegin += P.seedOffset;
                }
                Cursor c;
                c.lb  = 0;
                c.len = 0;
                if (mySeed != 0xffffffffu)
                    c = seedPrefixCursor(P, P.Q.red + F * qb + static_cast<unsigned long long>(myF) * origLen, mySeed, L);
                else
                    mySeed = 0;
                __syncwarp();
                seedConsumeChunkSpec(P, sM, S, lane, q, qb, origLen, needlesSum, c, myF, mySeed, myNpos, hitsThisSeq, nAfter,
                                     nFailed);
            }
        }
        else
        for (unsigned int f = 0; f < F; ++f)
        {
            unsigned int const len = qryFrameLen(P.Q, origLen, f);
            if (len < L)
                continue;
            unsigned char const * trans = P.Q.trans + F * qb + static_cast<unsigned long long>(f) * origLen;
            unsigned char const * red   = P.Q.red + F * qb + static_cast<unsigned long long>(f) * origLen;
            {
                for (unsigned int seedBegin = 0;; seedBegin += P.seedOffset)
                {
                    if (!seedNextStart(trans, len, L, P.unknownRank, seedBegin))
                        break;
                    Cursor    E[kMaxHalf2 + 1];
                    int const K = seedExactChain(P, red, seedBegin, h1, n2, E);
                    if (K < 0)
                        continue;
                    int const          kMax     = (K < static_cast<int>(n2) - 1) ? K : static_cast<int>(n2) - 1;
                    bool const         hasExact = (K == static_cast<int>(n2));
                    unsigned int const nLeaves  = P.maxSeedDist >= 2 ? seedNumLeaves2(n2, redN, P.fullHamming != 0)
                                                                  : static_cast<unsigned int>(kMax + 1) * (redN - 1) + (hasExact ? 1u : 0u);
                    for (unsigned int base = 0; base < nLeaves; base += 32)
                    {
                        Cursor mine;
                        mine.lb  = 0;
                        mine.len = 0;
                        if (base + lane < nLeaves)
                            mine = P.maxSeedDist >= 2 ? seedLeaf2(ix, E, K, red, seedBegin, h1, n2, redN, base + lane, P.fullHamming != 0)
                                                      : seedLeaf(ix, E, red, seedBegin, h1, n2, kMax, hasExact, redN, base + lane, P.fullHamming != 0);
                        __syncwarp();
                        seedConsumeChunkSpec(P, sM, S, lane, q, qb, origLen, needlesSum, mine, f, seedBegin, needlesPos,
                                             hitsThisSeq, nAfter, nFailed);
                    }
                }
            }
            needlesPos += len;
        }
    }
    seedFlushCounters(P, lane, nAfter, nFailed);
}

// ---------------------------------------------------------------------------------------------
// seeding, one BLOCK per query (few queries: latency matters, not throughput)
// ---------------------------------------------------------------------------------------------
// The cursors of different seeds do not depend on each other -- only their consumption does (through
// `hitsThisSeq`).  Phase A: the warps of the block search the seeds of the query in parallel and park
// the non-empty cursors of every seed, in the reference's order, in a scratch list.  Phase B: warp 0
// consumes the lists seed by seed exactly like the warp kernel.  With W warps the dependent-load chain
// of a query shrinks from (#seeds x tree depth) to (#seeds / W x tree depth) + consumption.
constexpr int kSeedBlockWarps = 16;
constexpr int kSeedBlockMaxSeeds = 2048;

struct SeedScratch
{
    Cursor *       cursors;  // [nActive][maxSeeds][maxLeaves]
    unsigned int * counts;   // [nActive][maxSeeds]
    unsigned int   maxSeeds; // seeds per query the scratch was sized for
    unsigned int   maxLeaves;
};

__global__ void __launch_bounds__(32 * kSeedBlockWarps) seedBlockKernel(SeedParams P, SeedScratch S)
{
    __shared__ signed char  sM[2048];
    __shared__ unsigned int sSeed[kSeedBlockMaxSeeds]; // frame << 24 | seedBegin
    __shared__ unsigned int sNumSeeds;
    __shared__ SpecScratch  sSpec;
    __shared__ unsigned long long sNeedlesPos[8], sNeedlesSum;
    for (int i = threadIdx.x; i < 2048; i += blockDim.x)
        sM[i] = P.matrix[i];

    unsigned int const lane = threadIdx.x & 31u;
    unsigned int const warp = threadIdx.x >> 5;
    unsigned int const b    = blockIdx.x;
    unsigned int const       q    = P.active[b];
    DevIndex const &         ix   = P.ix;
    unsigned int const       F    = P.Q.F;
    unsigned long long const qb      = P.Q.offs[q];
    unsigned int const       origLen = static_cast<unsigned int>(P.Q.offs[q + 1] - qb);
    unsigned int const       L       = P.seedLength;
    unsigned int const       redN    = ix.sigma - 1;
    bool const               half    = (P.halfExact && P.maxSeedDist != 0) || P.fullHamming;
    unsigned int const       h1      = P.fullHamming ? 0u : (half ? L / 2 : L);
    unsigned int const       n2      = L - h1;

    if (threadIdx.x == 0)
    {
        unsigned int       n   = 0;
        unsigned long long pos = 0, sum = 0;
        for (unsigned int f = 0; f < F; ++f)
            sum += qryFrameLen(P.Q, origLen, f);
        sNeedlesSum = sum;
        if (qryFrameLen(P.Q, origLen, 0) >= L)
            for (unsigned int f = 0; f < F; ++f)
            {
                unsigned int const len = qryFrameLen(P.Q, origLen, f);
                sNeedlesPos[f]         = pos;
                if (len < L)
                    continue;
                pos += len;
                unsigned char const * trans = P.Q.trans + F * qb + static_cast<unsigned long long>(f) * origLen;
                for (unsigned int seedBegin = 0;; seedBegin += P.seedOffset)
                {
                    if (!seedNextStart(trans, len, L, P.unknownRank, seedBegin))
                        break;
                    if (n < kSeedBlockMaxSeeds)
                        sSeed[n] = (f << 24) | seedBegin;
                    ++n;
                }
            }
        sNumSeeds = n; // the host only routes queries here whose seed count fits (checked again below)
    }
    __syncthreads();
    unsigned int const nSeeds = min(sNumSeeds, min(static_cast<unsigned int>(kSeedBlockMaxSeeds), S.maxSeeds));
    Cursor *           myCur  = S.cursors + static_cast<unsigned long long>(b) * S.maxSeeds * S.maxLeaves;
    unsigned int *     myCnt  = S.counts + static_cast<unsigned long long>(b) * S.maxSeeds;

    // ---- phase A: all warps, seeds round-robin ----
    for (unsigned int k = warp; k < nSeeds; k += kSeedBlockWarps)
    {
        unsigned int const    f         = sSeed[k] >> 24;
        unsigned int const    seedBegin = sSeed[k] & 0xffffffu;
        unsigned char const * red       = P.Q.red + F * qb + static_cast<unsigned long long>(f) * origLen;
        Cursor                E[kMaxHalf2 + 1];
        int const             K   = seedExactChain(P, red, seedBegin, h1, n2, E);
        unsigned int          out = 0;
        if (K >= 0)
        {
            int const          kMax     = (K < static_cast<int>(n2) - 1) ? K : static_cast<int>(n2) - 1;
            bool const         hasExact = (K == static_cast<int>(n2));
            unsigned int const nLeaves  = P.maxSeedDist >= 2 ? seedNumLeaves2(n2, redN, P.fullHamming != 0)
                                                                  : static_cast<unsigned int>(kMax + 1) * (redN - 1) + (hasExact ? 1u : 0u);
            for (unsigned int base = 0; base < nLeaves; base += 32)
            {
                Cursor mine;
                mine.lb  = 0;
                mine.len = 0;
                if (base + lane < nLeaves)
                    mine = P.maxSeedDist >= 2 ? seedLeaf2(ix, E, K, red, seedBegin, h1, n2, redN, base + lane, P.fullHamming != 0)
                                                      : seedLeaf(ix, E, red, seedBegin, h1, n2, kMax, hasExact, redN, base + lane, P.fullHamming != 0);
                unsigned int const live = __ballot_sync(0xffffffffu, mine.len != 0);
                if (mine.len != 0)
                    myCur[static_cast<unsigned long long>(k) * S.maxLeaves + out + __popc(live & ((1u << lane) - 1u))] = mine;
                out += __popc(live);
            }
        }
        if (lane == 0)
            myCnt[k] = out;
    }
    __syncthreads();

    // ---- phase B: warp 0 consumes the parked cursors in order, 32 at a time (speculatively) ----
    if (warp != 0)
        return;
    unsigned long long nAfter = 0, nFailed = 0;
    unsigned long long hitsThisSeq = 0;
    unsigned long long const needlesSum = sNeedlesSum;
    unsigned int k = 0, ci = 0;
    for (;;)
    {
        Cursor mine;
        mine.lb  = 0;
        mine.len = 0;
        unsigned int myF = 0, mySb = 0, got = 0;
        while (got < 32 && k < nSeeds)
        {
            unsigned int const n    = myCnt[k];
            unsigned int const take = min(n - ci, 32u - got);
            if (lane >= got && lane < got + take)
            {
                mine = myCur[static_cast<unsigned long long>(k) * S.maxLeaves + ci + (lane - got)];
                myF  = sSeed[k] >> 24;
                mySb = sSeed[k] & 0xffffffu;
            }
            got += take;
            ci += take;
            if (ci == n)
            {
                ++k;
                ci = 0;
            }
        }
        if (!got)
            break;
        __syncwarp();
        seedConsumeChunkSpec(P, sM, sSpec, lane, q, qb, origLen, needlesSum, mine, myF, mySb, sNeedlesPos[myF], hitsThisSeq,
                             nAfter, nFailed);
    }
    seedFlushCounters(P, lane, nAfter, nFailed);
}

} // namespace lgpu
