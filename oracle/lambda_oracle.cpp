// lambda_oracle -- scalar CPU restatement of the reference's seed-and-extend hot path.
//
// TEST INFRASTRUCTURE ONLY.  Nothing under lambda_b200/ includes, links or executes this file;
// only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs do.
//
// Parity status: PINNED.  The reference ships no offline golden files (they are downloaded by
// test/data/datasources.cmake), so the pin is the reference binary itself: oracle/Makefile builds
// the unmodified lambda3 from /root/reference into oracle/_ref/, tests/golden/make_golden.py runs
// it on seeded synthetic inputs and commits inputs + .m8 outputs + funnel counters, and
// tests/test_oracle_vs_reference.py requires this restatement to reproduce them exactly.
//
// What is restated (plain loops, no SeqAn/BioC++/FMC types), with the reference location:
//   fm_rank / extendRight          FMC occtable/InterleavedEPRV2.h:81-86,211-216; ReverseFMIndexCursor.h:30-34
//   fm_locate                      FMC ReverseFMIndex.h:62-91, CSA.h:104-113, BitvectorCompact.h:26-72,
//                                  occtable/InterleavedEPRV2.h:121-138,264-270, locate.h:28-35
//   seed search (exact, half-exact) src/search_algo.hpp:484-494,538-604; FMC search/BacktrackingWithBuffers.h:74-83
//   search(): seed loop, adaptive elongation, abundance cut, locate, pre-scoring
//                                  src/search_algo.hpp:607-762, seedLooksPromising :427-481
//   widen / sort / merge / unique  src/search_algo.hpp:920-938,1137-1175; _bandSize src/search_misc.hpp:46-50
//   local affine DP + trace bits   SQ/align/dp_formula_affine.h:66-126, dp_formula.h:136-243,
//                                  dp_scout_simd.h:216-229,565-578 (first strict maximum, column-major)
//   traceback                      SQ/align/dp_traceback_impl.h:223-258,302-337,379-474
//   alignment statistics           SQ/align/evaluate_alignment.h:215-300
//   filters / phases               src/search_algo.hpp:1252-1322,1391-1460
// Statistics (bit score, e-value, length adjustment), per-query finalisation and m8 formatting are
// host code shared with the product (lambda_b200/csrc/host_params.hpp, host_finalize.hpp): they are
// not device work and are pinned end-to-end by the same golden files.
#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <fstream>
#include <memory>
#include <string>
#include <tuple>
#include <vector>

#ifdef _OPENMP
#include <omp.h>
#endif

#include "../lambda_b200/csrc/host_finalize.hpp"
#include "../lambda_b200/csrc/host_params.hpp"
#include "../lambda_b200/csrc/n_random.hpp"
#include "../lambda_b200/csrc/lba_index.hpp"

namespace orc
{

using lgpu::DomainInfo;
using lgpu::Scoring;

struct Index
{
    std::unique_ptr<lgpu::LbaFile> file;
    lgpu_index_desc                d{};
    uint64_t                       dbTotalLength = 0;
    // the reference's transSbjSeqs (src/shared_definitions.hpp:246-255): the stored sequences, each one
    // twice for bisulfite indexes (views::duplicate), or their six-frame translations (translate_join)
    uint32_t              sbjFrames = 1;
    std::vector<uint8_t>  tSeqs;   // translated subjects only
    std::vector<uint64_t> tDelims; // 6 * n_seqs + 1

    uint8_t const * sbjSeq(uint64_t subjId) const
    {
        return sbjFrames == 6 ? tSeqs.data() + tDelims[subjId] : d.seqs + d.seq_delims[subjId / sbjFrames];
    }
    uint64_t sbjLen(uint64_t subjId) const
    {
        if (sbjFrames == 6)
            return tDelims[subjId + 1] - tDelims[subjId];
        return d.seq_delims[subjId / sbjFrames + 1] - d.seq_delims[subjId / sbjFrames];
    }
};

// one translated frame of a dna5 sequence (BIO ranges/views/translate_single.hpp:95-150, canonical code)
static void translateFrame(uint8_t const * nt, uint64_t len, uint32_t frame, std::vector<uint8_t> & out)
{
    uint64_t const n = lgpu::translatedFrameLength(len, frame);
    uint64_t const o = frame % 3;
    for (uint64_t k = 0; k < n; ++k)
    {
        uint8_t a, b, c;
        if (frame < 3)
        {
            a = nt[3 * k + o];
            b = nt[3 * k + o + 1];
            c = nt[3 * k + o + 2];
        }
        else
        {
            a = lgpu::kDna5Complement[nt[len - 3 * k - o - 1]];
            b = lgpu::kDna5Complement[nt[len - 3 * k - o - 2]];
            c = lgpu::kDna5Complement[nt[len - 3 * k - o - 3]];
        }
        out.push_back(lgpu::kDna5Translate[(a * 5 + b) * 5 + c]);
    }
}

// ---------------------------------------------------------------------------------------------
// FM index primitives
// ---------------------------------------------------------------------------------------------

static inline uint64_t popShift(uint64_t bits, uint64_t idx) // popcount(bits << (64 - idx)), idx in [0,64)
{
    return idx == 0 ? 0 : static_cast<uint64_t>(__builtin_popcountll(bits << (64 - idx)));
}

static inline uint8_t const * blockPtr(lgpu_index_desc const & d, uint64_t blk)
{
    return static_cast<uint8_t const *>(d.occ_blocks) + blk * d.block_bytes;
}

static uint64_t fmRank(lgpu_index_desc const & d, uint64_t idx, uint32_t symb)
{
    uint8_t const * b = blockPtr(d, idx >> 6);
    uint32_t        cnt;
    std::memcpy(&cnt, b + 4 * symb, 4);
    uint64_t mask = ~0ull;
    for (uint32_t p = 0; p < d.sigma_bits; ++p)
    {
        uint64_t plane;
        std::memcpy(&plane, b + d.planes_offset + 8 * p, 8);
        mask &= ((symb >> p) & 1) ? plane : ~plane;
    }
    return cnt + popShift(mask, idx & 63) + d.super_blocks[(idx >> 32) * d.sigma + symb] + d.C[symb];
}

// symbol at BWT position idx and its rank there (one LF step)
static uint64_t fmRankSymbol(lgpu_index_desc const & d, uint64_t idx)
{
    uint8_t const * b    = blockPtr(d, idx >> 6);
    uint64_t const  bit  = idx & 63;
    uint32_t        symb = 0;
    uint64_t        mask = ~0ull;
    for (uint32_t p = 0; p < d.sigma_bits; ++p)
    {
        uint64_t plane;
        std::memcpy(&plane, b + d.planes_offset + 8 * p, 8);
        uint64_t const v = (plane >> bit) & 1;
        symb |= static_cast<uint32_t>(v) << p;
        mask &= v ? plane : ~plane;
    }
    uint32_t cnt;
    std::memcpy(&cnt, b + 4 * symb, 4);
    return cnt + popShift(mask, bit) + d.super_blocks[(idx >> 32) * d.sigma + symb] + d.C[symb];
}

struct CsaSuper
{
    uint64_t entry;
    uint8_t  blocks[4];
    uint8_t  pad[4];
    uint64_t bits[4];
};
static_assert(sizeof(CsaSuper) == 48, "csa superblock layout");

static inline bool csaSampled(lgpu_index_desc const & d, uint64_t row)
{
    uint64_t const   i  = row + 1; // bit i of the vector is stored at position i + 1
    CsaSuper const * sb = static_cast<CsaSuper const *>(d.csa_bv) + (i >> 8);
    return (sb->bits[(i & 255) >> 6] >> (i & 63)) & 1;
}

static inline uint64_t csaRank(lgpu_index_desc const & d, uint64_t row)
{
    CsaSuper const * sb  = static_cast<CsaSuper const *>(d.csa_bv) + (row >> 8);
    uint64_t const   blk = (row & 255) >> 6;
    uint64_t const   bit = row & 63;
    return sb->entry + sb->blocks[blk] + static_cast<uint64_t>(__builtin_popcountll(sb->bits[blk] << (63 - bit)));
}

static void fmLocate(lgpu_index_desc const & d, uint64_t row, uint64_t & subj, uint64_t & pos)
{
    uint64_t steps = 0;
    while (!csaSampled(d, row))
    {
        row = fmRankSymbol(d, row);
        ++steps;
    }
    uint64_t const v = d.ssa[csaRank(d, row)];
    subj             = v >> d.bits_for_position;
    pos              = (v & ((1ull << d.bits_for_position) - 1)) - steps;
}

struct Cursor
{
    uint64_t lb = 0, len = 0, depth = 0;
};

static inline Cursor extendRight(lgpu_index_desc const & d, Cursor const & c, uint32_t symb)
{
    uint64_t const lb = fmRank(d, c.lb, symb);
    return {lb, fmRank(d, c.lb + c.len, symb) - lb, c.depth + 1};
}

// ---------------------------------------------------------------------------------------------
// queries: frames, translation, reduction
// ---------------------------------------------------------------------------------------------

struct Queries
{
    uint32_t              nFrames = 1;
    uint64_t              n       = 0; // original queries
    std::vector<uint8_t>  trans;       // all frames, translated-alphabet ranks
    std::vector<uint8_t>  red;         // same layout, reduced-alphabet ranks
    std::vector<uint64_t> offs;        // n * nFrames + 1
    std::vector<uint32_t> origLen;     // n
};

static uint8_t const * reductionTable(uint32_t redAlph)
{
    static uint8_t const identity27[27] = {0,  1,  2,  3,  4,  5,  6,  7,  8,  9,  10, 11, 12, 13,
                                           14, 15, 16, 17, 18, 19, 20, 21, 22, 23, 24, 25, 26};
    static uint8_t const dna5to4[5]     = {0, 1, 2, 0, 3}; // N: replaced by the marker in makeQueries (n_random.hpp)
    switch (redAlph)
    {
        case LGPU_ALPH_LI10: return lgpu::kAa27ToLi10;
        case LGPU_ALPH_MURPHY10: return lgpu::kAa27ToMurphy10;
        case LGPU_ALPH_AMINO_ACID: return identity27;
        case LGPU_ALPH_DNA4: return dna5to4;
        case LGPU_ALPH_DNA3BS: return dna5to4; // then mapped per frame direction, see makeQueries
        default: return nullptr;
    }
}

static Queries makeQueries(DomainInfo const & di, uint32_t redAlph, uint8_t const * res, uint64_t const * offs, uint64_t n)
{
    Queries         q;
    uint8_t const * red = reductionTable(redAlph);
    q.nFrames           = di.qryNumFrames;
    q.n                 = n;
    q.offs.push_back(0);
    std::vector<uint8_t> frame;
    for (uint64_t i = 0; i < n; ++i)
    {
        uint64_t const len = offs[i + 1] - offs[i];
        q.origLen.push_back(static_cast<uint32_t>(len));
        for (uint32_t f = 0; f < q.nFrames; ++f)
        {
            // protein: the query itself; nucleotide: (fwd, rc); translated: six frames; bisulfite:
            // (fwd, fwd, rc, rc), each pair reduced with the forward (C->T) and the reverse (G->A)
            // bisulfite alphabet alternately (src/shared_definitions.hpp:257-281,
            // src/view_reduce_to_bisulfite.hpp:51-52,133-135)
            bool const           bs          = redAlph == LGPU_ALPH_DNA3BS;
            static uint8_t const bsTab[2][4] = {{0, 1, 2, 1}, {3, 4, 3, 5}};
            frame.clear();
            if (di.qFrameMode == 2)
                translateFrame(res + offs[i], len, f, frame);
            else
            {
                bool const revComp = di.qFrameMode == 3 ? (f >= 2) : (di.qFrameMode == 1 && f == 1);
                for (uint64_t k = 0; k < len; ++k)
                    frame.push_back(!revComp ? res[offs[i] + k] : lgpu::kDna5Complement[res[offs[i] + len - 1 - k]]);
            }
            bool const nucl = redAlph == LGPU_ALPH_DNA4 || bs; // views::dna_n_to_random sits in these reduced views only
            for (uint8_t r : frame)
            {
                q.trans.push_back(r);
                if (nucl && r == 3) // dna5 'N': randomised per read of the view (n_random.hpp)
                    q.red.push_back(static_cast<uint8_t>(lgpu::kNMarker | (f & 1u)));
                else
                    q.red.push_back(bs ? bsTab[f % 2][red[r]] : red[r]);
            }
            q.offs.push_back(q.trans.size());
        }
    }
    return q;
}

// ---------------------------------------------------------------------------------------------
// seeding
// ---------------------------------------------------------------------------------------------

struct Ctx
{
    Index const * idx = nullptr;
    lgpu_params   p{};
    Scoring       sc;
    DomainInfo    di;
};

// scoring matrix for a (frame-expanded) subject: bisulfite subjects with odd id were converted in the
// reverse direction and use the reverse matrix (src/search_algo.hpp:464-466,1098,1249)
static inline int8_t const * matrixFor(Ctx const & c, uint32_t subjId)
{
    return (c.p.domain == LGPU_DOMAIN_BISULFITE && (subjId & 1u)) ? c.sc.matrixRev : c.sc.matrix;
}

static bool seedLooksPromising(Ctx const & c, Queries const & q, lgpu_search_opts const & so, lgpu_match const & m)
{
    int64_t                 qBegin  = m.qry_start;
    int64_t                 sBegin  = m.subj_start;
    uint64_t const          actual  = m.qry_end - m.qry_start;
    uint64_t                effLen  = std::max<uint64_t>(static_cast<uint64_t>(so.seed_length * c.p.pre_scoring), actual);
    uint64_t const          qLen    = q.offs[m.qry_id + 1] - q.offs[m.qry_id];
    uint64_t const          sLen    = c.idx->sbjLen(m.subj_id);
    int8_t const *          M       = matrixFor(c, m.subj_id);
    if (effLen > actual)
    {
        qBegin -= (effLen - actual) / 2;
        sBegin -= (effLen - actual) / 2;
        int64_t const mn = std::min(qBegin, sBegin);
        if (mn < 0)
        {
            qBegin -= mn;
            sBegin -= mn;
            effLen += mn;
        }
        effLen = std::min({static_cast<uint64_t>(qLen - qBegin), static_cast<uint64_t>(sLen - sBegin), effLen});
    }
    uint8_t const * qs     = q.trans.data() + q.offs[m.qry_id] + qBegin;
    uint8_t const * ss     = c.idx->sbjSeq(m.subj_id) + sBegin;
    int             s      = 0;
    int             maxS   = 0;
    int const       thresh = static_cast<int>(c.p.pre_scoring_thresh * effLen);
    for (uint64_t i = 0; i < effLen; ++i)
    {
        s += M[qs[i] * 32 + ss[i]];
        if (s < 0)
            s = 0;
        else if (s > maxS)
            maxS = s;
        if (maxS >= thresh)
            return true;
    }
    return false;
}

// all cursors for one seed, in the reference's production order
static void seedCursors(Ctx const & c, lgpu_search_opts const & so, uint8_t const * redSeedRaw, std::vector<Cursor> & out)
{
    lgpu_index_desc const & d = c.idx->d;
    out.clear();
    Cursor const root{0, d.C[d.sigma], 0};
    // One instance of views::dna_n_to_random serves this whole seed search: every READ of an 'N' takes the next
    // random rank (n_random.hpp).  `redSeed` below restates exactly the reads the reference makes, in its order.
    bool const   bs     = d.red_alph == LGPU_ALPH_DNA3BS;
    unsigned int nReads = 0;
    struct SeedReads
    {
        uint8_t const * raw;
        bool            bs;
        unsigned int *  nReads;
        uint8_t operator[](uint32_t i) const
        {
            uint8_t const s = raw[i];
            if (s < lgpu::kNMarker)
                return s;
            return static_cast<uint8_t>(lgpu::nReducedRank(lgpu::nRandomRank((*nReads)++), bs, s & 1u));
        }
    } const redSeed{redSeedRaw, bs, &nReads};
    if (!(c.p.seed_half_exact && so.max_seed_dist != 0))
    {
        // search_impl -> search_backtracking_with_buffers (src/search_algo.hpp:484-494,
        // FMC search/BacktrackingWithBuffers.h:40-83): Hamming distance over the whole seed
        auto noErrors = [&](Cursor cur, uint32_t i)
        {
            for (; i < so.seed_length; ++i)
            {
                cur = extendRight(d, cur, redSeed[i] + 1u);
                if (cur.len == 0)
                    return;
            }
            out.push_back(cur);
        };
        if (so.max_seed_dist == 0)
        {
            noErrors(root, 0);
            return;
        }
        uint32_t const                           sigmaAll = d.sigma; // symbols 1 .. sigma-1 (0 = sentinel)
        std::vector<std::pair<Cursor, uint32_t>> b1, b2;
        b1.emplace_back(root, 0u);
        for (uint32_t i = 0; i < so.seed_length; ++i)
        {
            uint32_t const r = redSeed[i] + 1u;
            for (auto const & [cu, err] : b1)
                for (uint32_t s = 1; s < sigmaAll; ++s)
                {
                    Cursor const n = extendRight(d, cu, s);
                    if (n.len == 0)
                        continue;
                    uint32_t const e2 = err + (r != s ? 1u : 0u);
                    if (e2 < so.max_seed_dist)
                        b2.emplace_back(n, e2);
                    else
                        noErrors(n, i + 1);
                }
            b1.clear();
            std::swap(b1, b2);
        }
        for (auto const & e : b1)
            out.push_back(e.first);
        return;
    }
    uint32_t const half1 = so.seed_length / 2;
    uint32_t const half2 = so.seed_length - half1;
    uint32_t const sigma = d.sigma - 1; // size of the reduced alphabet
    std::vector<std::pair<Cursor, uint32_t>> buf, buf2;
    Cursor                                   cur = root;
    for (uint32_t i = 0; i < half1; ++i)
    {
        cur = extendRight(d, cur, redSeed[i] + 1u);
        if (cur.len == 0)
            return;
    }
    buf.emplace_back(cur, 0u);
    for (uint32_t i = 0; i < half2; ++i)
    {
        uint8_t const want = redSeed[half1 + i];
        for (auto const & [cu, err] : buf)
        {
            if (err < so.max_seed_dist)
            {
                for (uint32_t r = 0; r < sigma; ++r)
                {
                    Cursor const n = extendRight(d, cu, r + 1u);
                    if (n.len != 0)
                        buf2.emplace_back(n, err + (r != want));
                }
            }
            else
            {
                Cursor const n = extendRight(d, cu, want + 1u);
                if (n.len != 0)
                    buf2.emplace_back(n, err);
            }
        }
        buf.clear();
        std::swap(buf, buf2);
    }
    for (auto const & e : buf)
        out.push_back(e.first);
}

// search() for the queries flagged active; appends pre-score-passing matches
static void seedQueries(Ctx const & c, Queries const & q, lgpu_search_opts const & so, std::vector<uint8_t> const & active,
                        std::vector<lgpu_match> & matches, lgpu_stats & st)
{
    lgpu_index_desc const & d  = c.idx->d;
    uint32_t const          F  = q.nFrames;
    uint64_t const          hf = 10; // heuristicFactor
    std::vector<Cursor>     cursors;
    uint64_t hitsThisSeq = 0, needlesSum = 0, needlesPos = 0;
    for (uint64_t i = 0; i < q.n * F; ++i)
    {
        if (!active[i / F])
            continue;
        uint64_t const  len   = q.offs[i + 1] - q.offs[i];
        uint8_t const * red   = q.red.data() + q.offs[i];
        uint8_t const * trans = q.trans.data() + q.offs[i];
        if (len < so.seed_length)
            continue;
        if (i % F == 0)
        {
            hitsThisSeq = 0;
            needlesSum  = 0;
            needlesPos  = 0;
            for (uint32_t j = 0; j < F; ++j)
                needlesSum += q.offs[i + j + 1] - q.offs[i + j];
        }
        for (uint64_t seedBegin = 0;; seedBegin += so.seed_offset)
        {
            while (seedBegin < len - so.seed_length &&
                   (trans[seedBegin] == c.di.unknownRank || trans[seedBegin] == trans[seedBegin + 1]))
                ++seedBegin;
            if (seedBegin > len - so.seed_length)
                break;
            seedCursors(c, so, red + seedBegin, cursors);
            for (Cursor cursor : cursors)
            {
                uint64_t seedLength = so.seed_length;
                if (c.p.adaptive_seeding)
                {
                    uint64_t desiredOccs =
                      hitsThisSeq >= c.p.max_matches
                        ? 1
                        : (c.p.max_matches - hitsThisSeq) * hf /
                            std::max<uint64_t>((needlesSum - needlesPos - seedBegin) / so.seed_offset, 1);
                    if (desiredOccs == 0)
                        desiredOccs = 1;
                    Cursor   oldCursor = cursor;
                    uint64_t oldCount  = cursor.len;
                    while (seedBegin + seedLength < len)
                    {
                        // a fresh view instance per read: an 'N' here is always the first random rank
                        cursor = extendRight(d, cursor, lgpu::redSymbol(red, seedBegin + seedLength, seedBegin + seedLength,
                                                                        d.red_alph == LGPU_ALPH_DNA3BS) + 1u);
                        uint64_t const newCount = cursor.len;
                        if (newCount < desiredOccs && newCount < oldCount)
                        {
                            cursor = oldCursor;
                            break;
                        }
                        ++seedLength;
                        oldCount  = newCount;
                        oldCursor = cursor;
                    }
                }
                if (cursor.len > hf * c.p.max_matches)
                    continue;
                for (uint64_t row = cursor.lb; row < cursor.lb + cursor.len; ++row)
                {
                    uint64_t subj, pos;
                    fmLocate(d, row, subj, pos);
                    pos -= cursor.depth;
                    lgpu_match m{static_cast<uint32_t>(i),
                                 static_cast<uint32_t>(subj),
                                 static_cast<uint32_t>(seedBegin),
                                 static_cast<uint32_t>(seedBegin + seedLength),
                                 static_cast<uint32_t>(pos),
                                 static_cast<uint32_t>(pos + seedLength)};
                    ++st.hits_after_seeding;
                    if (!seedLooksPromising(c, q, so, m))
                        ++st.hits_failed_pre_extend;
                    else
                    {
                        matches.push_back(m);
                        ++hitsThisSeq;
                    }
                }
            }
        }
        needlesPos += len;
    }
}

// ---------------------------------------------------------------------------------------------
// widen / merge
// ---------------------------------------------------------------------------------------------

static inline bool matchLess(lgpu_match const & a, lgpu_match const & b)
{
    return std::tie(a.qry_id, a.subj_id, a.qry_start, a.qry_end, a.subj_start, a.subj_end) <
           std::tie(b.qry_id, b.subj_id, b.qry_start, b.qry_end, b.subj_start, b.subj_end);
}
static inline bool matchEq(lgpu_match const & a, lgpu_match const & b)
{
    return std::tie(a.qry_id, a.subj_id, a.qry_start, a.qry_end, a.subj_start, a.subj_end) ==
           std::tie(b.qry_id, b.subj_id, b.qry_start, b.qry_end, b.subj_start, b.subj_end);
}

static void widenAndMerge(Ctx const & c, Queries const & q, std::vector<lgpu_match> & ms, lgpu_stats & st)
{
    size_t const            before = ms.size();
    for (lgpu_match & m : ms)
    {
        uint64_t const qLen = q.offs[m.qry_id + 1] - q.offs[m.qry_id];
        uint64_t const sLen = c.idx->sbjLen(m.subj_id);
        uint64_t       s0   = (m.subj_start < m.qry_start) ? 0 : m.subj_start - m.qry_start;
        m.qry_start         = 0;
        m.qry_end           = static_cast<uint32_t>(qLen);
        uint64_t const band = c.p.window_band ? static_cast<uint64_t>(c.p.window_band) // lgpu_params.window_band (band sweep)
                                              : static_cast<uint64_t>(static_cast<int64_t>(std::sqrt(static_cast<double>(qLen))) + 1);
        m.subj_end          = static_cast<uint32_t>(std::min<uint64_t>(s0 + qLen + band, sLen));
        m.subj_start        = static_cast<uint32_t>((band < s0) ? s0 - band : 0);
    }
    std::sort(ms.begin(), ms.end(), matchLess);
    if (ms.size() > 1)
    {
        for (size_t i = 0; i + 1 < ms.size(); ++i)
        {
            lgpu_match & l = ms[i];
            lgpu_match & r = ms[i + 1];
            if (l.qry_id == r.qry_id && l.subj_id == r.subj_id && l.subj_end >= r.subj_start)
            {
                l.subj_end   = r.subj_end;
                r.subj_start = l.subj_start;
            }
        }
        for (size_t i = ms.size() - 1; i > 0; --i)
        {
            lgpu_match & r = ms[i];
            lgpu_match & l = ms[i - 1];
            if (r.qry_id == l.qry_id && r.subj_id == l.subj_id && r.subj_start < l.subj_end)
                l = r;
        }
        ms.erase(std::unique(ms.begin(), ms.end(), matchEq), ms.end());
        st.hits_duplicate += before - ms.size();
    }
}

// ---------------------------------------------------------------------------------------------
// extension DP
// ---------------------------------------------------------------------------------------------

enum : uint8_t { T_DIAG = 1, T_HORI = 2, T_VERT = 4, T_HOPEN = 8, T_VOPEN = 16, T_MAXH = 32, T_MAXV = 64 };

struct DpResult
{
    int      score = 0;
    uint32_t bi = 0, bj = 0; // end of the alignment (exclusive), query / window
    uint32_t ai = 0, aj = 0; // begin
    uint32_t nMatch = 0, nMismatch = 0, nGapOpen = 0, nGapExt = 0, nPositive = 0, alnLen = 0;
    std::vector<uint32_t> ops; // run-length operations of the path, traceback order (end first): run << 2 | LGPU_CIGAR_*
};

// query = columns (outer loop), subject window = rows (inner loop)
// `M` = 32-strided score matrix [query residue][subject residue]; `bsStats`: bisulfite rule for
// "identical" columns (src/evaluate_bisulfite_alignment.hpp:97)
static DpResult alignLocal(Scoring const & sc, int8_t const * M, bool bsStats, uint8_t const * qs, uint32_t nq,
                           uint8_t const * ts, uint32_t nt, bool withTrace, std::vector<uint8_t> & T, std::vector<int> & buf)
{
    constexpr int NEG = -16384; // INT16_MIN / 2 (SQ/align/dp_cell.h:144-146)
    int const     go = sc.gapOpenSeqan, ge = sc.gapExtend;
    DpResult      r;
    buf.assign(4 * (static_cast<size_t>(nt) + 1), 0);
    int * Sprev = buf.data();
    int * Scur  = Sprev + nt + 1;
    int * Hprev = Scur + nt + 1;
    int * Hcur  = Hprev + nt + 1;
    for (uint32_t j = 0; j <= nt; ++j)
    {
        Sprev[j] = 0;
        Hprev[j] = NEG;
    }
    size_t const stride = static_cast<size_t>(nt) + 1;
    if (withTrace)
        T.assign((static_cast<size_t>(nq) + 1) * stride, 0);
    for (uint32_t i = 1; i <= nq; ++i)
    {
        int v = NEG, sUp = 0;
        Scur[0] = 0;
        Hcur[0] = NEG;
        int8_t const * row = M + qs[i - 1] * 32;
        for (uint32_t j = 1; j <= nt; ++j)
        {
            int     diag = Sprev[j - 1] + row[ts[j - 1]];
            int     a = Hprev[j] + ge, b = Sprev[j] + go;
            int     h;
            uint8_t tv;
            if (a == b) { h = a; tv = T_HORI | T_HOPEN; }
            else if (a < b) { h = b; tv = T_HOPEN; }
            else { h = a; tv = T_HORI; }
            a = v + ge;
            b = sUp + go;
            if (a == b) { v = a; tv |= T_VERT | T_VOPEN; }
            else if (a < b) { v = b; tv |= T_VOPEN; }
            else { v = a; tv |= T_VERT; }
            int     g;
            uint8_t t2;
            if (v == h) { g = v; t2 = T_MAXV | T_MAXH; }
            else if (v < h) { g = h; t2 = T_MAXH; }
            else { g = v; t2 = T_MAXV; }
            int cur;
            if (diag == g) { cur = diag; tv |= T_DIAG | t2; }
            else if (diag < g) { cur = g; tv |= t2; }
            else { cur = diag; tv |= T_DIAG; }
            if (cur <= 0) { cur = 0; tv = 0; }
            Scur[j] = cur;
            Hcur[j] = h;
            if (withTrace)
                T[i * stride + j] = tv;
            sUp = cur;
            if (cur > r.score)
            {
                r.score = cur;
                r.bi    = i;
                r.bj    = j;
            }
        }
        std::swap(Sprev, Scur);
        std::swap(Hprev, Hcur);
    }
    if (!withTrace || r.score <= 0)
        return r;

    // traceback (GapsLeft, affine); statistics accumulated on the fly
    enum { DIAG, HORI, VERT } last;
    uint32_t i = r.bi, j = r.bj;
    uint8_t  tv = T[i * stride + j];
    if (tv & T_MAXV) { tv &= (T_VERT | T_VOPEN | T_MAXV); last = VERT; }
    else if (tv & T_MAXH) { tv &= (T_HORI | T_HOPEN | T_MAXH); last = HORI; }
    else last = DIAG;
    uint32_t run = 0;
    auto flush = [&](int kind, uint32_t n) {
        if (n == 0)
            return;
        // the gapped rows (alignRow0 / alignRow1 of the reference's BlastMatch) as runs: a diagonal run aligns residues,
        // a horizontal run consumes query residues only (gap in the subject row), a vertical run subject residues only
        r.ops.push_back((n << 2) | static_cast<uint32_t>(kind == DIAG ? LGPU_CIGAR_M : kind == HORI ? LGPU_CIGAR_I : LGPU_CIGAR_D));
        r.alnLen += n;
        if (kind != DIAG)
        {
            r.nGapOpen += 1;
            r.nGapExt += n - 1;
        }
    };
    auto switchTo = [&](decltype(last) k) {
        if (last != k)
        {
            flush(last, run);
            last = k;
            run  = 0;
        }
    };
    auto diagStats = [&](uint32_t ii, uint32_t jj) { // residues consumed by a diagonal step ending at (ii, jj)
        uint8_t const a = qs[ii - 1], b = ts[jj - 1];
        bool const isMatch = bsStats ? (M[a * 32 + b] == M[a * 32 + a]) : (a == b);
        if (isMatch) ++r.nMatch; else ++r.nMismatch;
        if (M[a * 32 + b] > 0) ++r.nPositive;
    };
    while (i > 0 && j > 0 && tv != 0)
    {
        if (tv & T_DIAG)
        {
            switchTo(DIAG);
            diagStats(i, j);
            --i; --j; tv = T[i * stride + j]; ++run;
        }
        else if ((tv & T_MAXV) && (tv & T_VERT))
        {
            switchTo(VERT);
            while ((!(tv & T_VOPEN) || (tv & T_VERT)) && j != 1)
            {
                --j; tv = T[i * stride + j]; ++run;
            }
            --j; tv = T[i * stride + j]; ++run;
        }
        else if ((tv & T_MAXV) && (tv & T_VOPEN))
        {
            switchTo(VERT);
            --j; tv = T[i * stride + j]; ++run;
        }
        else if ((tv & T_MAXH) && (tv & T_HORI))
        {
            switchTo(HORI);
            while ((!(tv & T_HOPEN) || (tv & T_HORI)) && i != 1)
            {
                --i; tv = T[i * stride + j]; ++run;
            }
            --i; tv = T[i * stride + j]; ++run;
        }
        else if ((tv & T_MAXH) && (tv & T_HOPEN))
        {
            switchTo(HORI);
            --i; tv = T[i * stride + j]; ++run;
        }
        else
            break; // unreachable for a consistent trace matrix
    }
    flush(last, run);
    r.ai = i;
    r.aj = j;
    return r;
}

static inline void windowOf(Ctx const & c, Queries const & q, lgpu_match const & m, uint8_t const *& qs, uint32_t & nq,
                            uint8_t const *& ts, uint32_t & nt)
{
    qs                        = q.trans.data() + q.offs[m.qry_id] + m.qry_start;
    nq                        = m.qry_end - m.qry_start;
    ts                        = c.idx->sbjSeq(m.subj_id) + m.subj_start;
    nt                        = m.subj_end - m.subj_start;
}

static lgpu_hit makeHit(Ctx const & c, Queries const & q, lgpu_match const & m, DpResult const & r, uint8_t phase)
{
    lgpu_index_desc const & d = c.idx->d;
    lgpu_hit                h{};
    h.q_id    = m.qry_id / q.nFrames;
    h.s_id    = m.subj_id / c.di.sbjNumFrames;
    h.q_len   = q.origLen[h.q_id];
    h.s_len   = static_cast<uint32_t>(d.seq_delims[h.s_id + 1] - d.seq_delims[h.s_id]);
    h.q_start = m.qry_start + r.ai;
    h.q_end   = m.qry_start + r.bi;
    h.s_start = m.subj_start + r.aj;
    h.s_end   = m.subj_start + r.bj;
    h.score   = r.score;
    h.n_match = r.nMatch; h.n_mismatch = r.nMismatch; h.n_gap_open = r.nGapOpen; h.n_gap_ext = r.nGapExt;
    h.n_positive = r.nPositive; h.aln_len = r.alnLen;
    // _setFrames (src/search_algo.hpp:769-814)
    h.q_frame = 0;
    h.s_frame = 0;
    if (c.di.qIsTranslated)
    {
        h.q_frame = static_cast<int8_t>((m.qry_id % 3) + 1);
        if (m.qry_id % 6 > 2)
            h.q_frame = static_cast<int8_t>(-h.q_frame);
    }
    else if (c.p.domain == LGPU_DOMAIN_BISULFITE)
    {
        h.q_frame = static_cast<int8_t>((m.qry_id % 2) + 1);
        if (m.qry_id % 4 > 1)
            h.q_frame = static_cast<int8_t>(-h.q_frame);
    }
    else if (c.p.domain == LGPU_DOMAIN_NUCLEOTIDE)
        h.q_frame = (m.qry_id % 2) ? -1 : 1;
    if (c.di.sIsTranslated)
    {
        h.s_frame = static_cast<int8_t>((m.subj_id % 3) + 1);
        if (m.subj_id % 6 > 2)
            h.s_frame = static_cast<int8_t>(-h.s_frame);
    }
    else if (c.p.domain == LGPU_DOMAIN_BISULFITE)
        h.s_frame = static_cast<int8_t>((m.subj_id % 2) + 1);
    h.phase   = phase;
    return h;
}

// iterateMatchesFullSimd for one phase; appends to `hits`
static void extendMatches(Ctx const & c, Queries const & q, std::vector<lgpu_match> & ms, lgpu::EValueComputer & ev,
                          uint8_t phase, std::vector<lgpu_hit> & hits, lgpu_stats & st, std::vector<uint32_t> * cigar = nullptr)
{
    widenAndMerge(c, q, ms, st);
    std::vector<uint8_t> T;
    std::vector<int>     buf;
    for (lgpu_match const & m : ms)
    {
        uint8_t const *qs, *ts;
        uint32_t       nq, nt;
        windowOf(c, q, m, qs, nq, ts, nt);
        ++st.n_extensions_score;
        st.cells_score += static_cast<uint64_t>(nq) * nt;
        int8_t const * M    = matrixFor(c, m.subj_id);
        bool const     bs   = c.p.domain == LGPU_DOMAIN_BISULFITE;
        DpResult const r1   = alignLocal(c.sc, M, bs, qs, nq, ts, nt, false, T, buf);
        uint32_t const qLen = q.origLen[m.qry_id / q.nFrames];
        double bits = 0, evalue = 0;
        if (c.p.min_bit_score >= 0)
        {
            bits = lgpu::bitScore(c.sc.ka, r1.score);
            if (bits < c.p.min_bit_score) { ++st.hits_failed_bitscore; continue; }
        }
        if (c.p.max_evalue >= 0)
        {
            evalue = ev.evalue(r1.score, qLen);
            if (evalue > c.p.max_evalue) { ++st.hits_failed_evalue; continue; }
        }
        ++st.n_extensions_trace;
        st.cells_trace += static_cast<uint64_t>(nq) * nt;
        DpResult const r2 = alignLocal(c.sc, M, bs, qs, nq, ts, nt, true, T, buf);
        lgpu_hit       h  = makeHit(c, q, m, r2, phase);
        float const identity = static_cast<float>(100.0 * static_cast<float>(h.n_match) / static_cast<float>(h.aln_len));
        if (identity < c.p.id_cutoff) { ++st.hits_failed_identity; continue; }
        // alignStats.alignmentScore is recomputed from the rows by computeAlignmentStats; it equals the DP score
        h.bit_score = lgpu::bitScore(c.sc.ka, h.score);
        h.evalue    = (c.p.max_evalue >= 0) ? evalue : ev.evalue(h.score, qLen);
        if (cigar && c.p.want_cigar) // lgpu_params.want_cigar: lgpu_hit.cigar_off / cigar_len index the call's run buffer
        {
            h.cigar_off = static_cast<uint32_t>(cigar->size());
            h.cigar_len = static_cast<uint32_t>(r2.ops.size());
            cigar->insert(cigar->end(), r2.ops.begin(), r2.ops.end());
        }
        hits.push_back(h);
    }
}

static int searchAll(Ctx const & c, uint8_t const * res, uint64_t const * offs, uint64_t n, std::vector<lgpu_hit> & hits,
                     lgpu_stats & st, std::vector<uint32_t> * cigar = nullptr)
{
    Queries              q = makeQueries(c.di, c.idx->d.red_alph, res, offs, n);
    lgpu::EValueComputer ev(c.sc.ka, c.idx->dbTotalLength, c.di.qIsTranslated);
    std::vector<uint8_t>    active(n, 1);
    std::vector<lgpu_match> ms;
    hits.clear();
    if (c.p.iterative_search)
    {
        seedQueries(c, q, c.p.opts0, active, ms, st);
        extendMatches(c, q, ms, ev, 1, hits, st, cigar);
        for (lgpu_hit const & h : hits)
            active[h.q_id] = 0;
        ms.clear();
        if (std::find(active.begin(), active.end(), 1) != active.end())
        {
            seedQueries(c, q, c.p.opts, active, ms, st);
            extendMatches(c, q, ms, ev, 2, hits, st, cigar);
        }
    }
    else
    {
        seedQueries(c, q, c.p.opts, active, ms, st);
        extendMatches(c, q, ms, ev, 2, hits, st, cigar);
    }
    if (c.p.finalize)
        lgpu::finalizeRecords(hits, c.p.max_matches, st);
    return 0;
}

} // namespace orc

// -------------------------------------------------------------------------------------------------
// C interface for the tests (ctypes)
// -------------------------------------------------------------------------------------------------

extern "C"
{

struct orc_handle
{
    orc::Index               idx;
    std::vector<lgpu_match>  matches;
    std::vector<lgpu_hit>    hits;
    std::vector<uint32_t>    cigar; // run buffer of the last orc_search call (lgpu_params.want_cigar)
    std::string              err;
};

orc_handle * orc_open(char const * path)
{
    auto * h = new orc_handle;
    try
    {
        h->idx.file.reset(new lgpu::LbaFile(path));
        h->idx.d = h->idx.file->desc;
        // dbTotalLength = sum of reduced subject lengths (src/search_algo.hpp:317-318)
        lgpu_index_desc const & d = h->idx.d;
        h->idx.sbjFrames          = d.red_alph == LGPU_ALPH_DNA3BS ? 2 : 1;
        h->idx.dbTotalLength      = d.n_residues * h->idx.sbjFrames;
        if (d.trans_alph == LGPU_ALPH_AMINO_ACID && d.orig_alph == LGPU_ALPH_DNA5)
        {
            h->idx.sbjFrames = 6;
            h->idx.tDelims.push_back(0);
            for (uint64_t sq = 0; sq < d.n_seqs; ++sq)
                for (uint32_t f = 0; f < 6; ++f)
                {
                    orc::translateFrame(d.seqs + d.seq_delims[sq], d.seq_delims[sq + 1] - d.seq_delims[sq], f, h->idx.tSeqs);
                    h->idx.tDelims.push_back(h->idx.tSeqs.size());
                }
            h->idx.dbTotalLength = h->idx.tSeqs.size();
        }
    }
    catch (std::exception const & e)
    {
        std::fprintf(stderr, "orc_open: %s\n", e.what());
        delete h;
        return nullptr;
    }
    return h;
}

void orc_close(orc_handle * h) { delete h; }

// run buffer of the last orc_search() call made with lgpu_params.want_cigar (see lgpu_hit.cigar_off / cigar_len)
uint64_t orc_last_cigar(orc_handle * h, uint32_t const ** ops)
{
    *ops = h->cigar.data();
    return h->cigar.size();
}

lgpu_index_desc const * orc_desc(orc_handle const * h) { return &h->idx.d; }

static int makeCtx(orc::Ctx & c, orc_handle const * h, lgpu_params const * p)
{
    c.idx = &h->idx;
    c.p   = *p;
    c.di  = lgpu::domainInfo(p->domain, h->idx.d.orig_alph, p->query_alph);
    return lgpu::makeScoring(c.sc, *p);
}

void orc_rank(orc_handle const * h, uint64_t const * idx, uint8_t const * symb, uint64_t n, uint64_t * out)
{
    for (uint64_t i = 0; i < n; ++i)
        out[i] = orc::fmRank(h->idx.d, idx[i], symb[i]);
}

void orc_locate(orc_handle const * h, uint64_t const * rows, uint64_t n, uint64_t * subj, uint64_t * pos)
{
    for (uint64_t i = 0; i < n; ++i)
        orc::fmLocate(h->idx.d, rows[i], subj[i], pos[i]);
}

int orc_seed(orc_handle * h, lgpu_params const * p, uint8_t const * res, uint64_t const * offs, uint64_t n, int phase,
             lgpu_match const ** out, uint64_t * nOut, lgpu_stats * st)
{
    orc::Ctx c;
    if (int rc = makeCtx(c, h, p)) return rc;
    orc::Queries         q = orc::makeQueries(c.di, h->idx.d.red_alph, res, offs, n);
    std::vector<uint8_t> active(n, 1);
    h->matches.clear();
    orc::seedQueries(c, q, phase == 1 ? p->opts0 : p->opts, active, h->matches, *st);
    *out  = h->matches.data();
    *nOut = h->matches.size();
    return 0;
}

int orc_merge(orc_handle * h, lgpu_params const * p, uint8_t const * res, uint64_t const * offs, uint64_t n,
              lgpu_match const * in, uint64_t nIn, lgpu_match const ** out, uint64_t * nOut, lgpu_stats * st)
{
    orc::Ctx c;
    if (int rc = makeCtx(c, h, p)) return rc;
    orc::Queries q = orc::makeQueries(c.di, h->idx.d.red_alph, res, offs, n);
    h->matches.assign(in, in + nIn);
    orc::widenAndMerge(c, q, h->matches, *st);
    *out  = h->matches.data();
    *nOut = h->matches.size();
    return 0;
}

int orc_extend(orc_handle * h, lgpu_params const * p, uint8_t const * res, uint64_t const * offs, uint64_t n,
               lgpu_match const * win, uint64_t nWin, int withTrace, int32_t * scores, lgpu_hit * hitsOut)
{
    orc::Ctx c;
    if (int rc = makeCtx(c, h, p)) return rc;
    orc::Queries         q = orc::makeQueries(c.di, h->idx.d.red_alph, res, offs, n);
    std::vector<uint8_t> T;
    std::vector<int>     buf;
    for (uint64_t i = 0; i < nWin; ++i)
    {
        uint8_t const *qs, *ts;
        uint32_t       nq, nt;
        orc::windowOf(c, q, win[i], qs, nq, ts, nt);
        orc::DpResult r = orc::alignLocal(c.sc, orc::matrixFor(c, win[i].subj_id), p->domain == LGPU_DOMAIN_BISULFITE, qs, nq,
                                          ts, nt, withTrace != 0, T, buf);
        if (scores) scores[i] = r.score;
        if (hitsOut && withTrace) hitsOut[i] = orc::makeHit(c, q, win[i], r, 0);
    }
    return 0;
}

// full search of one batch; threads > 1 shards the queries (results are per-query independent)
int orc_search(orc_handle * h, lgpu_params const * p, uint8_t const * res, uint64_t const * offs, uint64_t n, int threads,
               lgpu_hit const ** out, uint64_t * nOut, lgpu_stats * st)
{
    orc::Ctx c;
    if (int rc = makeCtx(c, h, p)) return rc;
    if (threads < 1) threads = 1;
    std::vector<std::vector<lgpu_hit>> parts(threads);
    std::vector<std::vector<uint32_t>> cigars(threads);
    std::vector<lgpu_stats>            stats(threads);
    std::memset(stats.data(), 0, sizeof(lgpu_stats) * threads);
#pragma omp parallel for num_threads(threads) schedule(static, 1)
    for (int t = 0; t < threads; ++t)
    {
        uint64_t const b = n * t / threads, e = n * (t + 1) / threads;
        if (b == e) continue;
        std::vector<uint64_t> o(offs + b, offs + e + 1);
        uint64_t const        base = o[0];
        for (auto & x : o) x -= base;
        orc::searchAll(c, res + base, o.data(), e - b, parts[t], stats[t], &cigars[t]);
        for (auto & hit : parts[t]) hit.q_id += static_cast<uint32_t>(b);
    }
    h->hits.clear();
    h->cigar.clear();
    for (int t = 0; t < threads; ++t)
    {
        for (auto & hit : parts[t]) hit.cigar_off += static_cast<uint32_t>(h->cigar.size());
        h->cigar.insert(h->cigar.end(), cigars[t].begin(), cigars[t].end());
        h->hits.insert(h->hits.end(), parts[t].begin(), parts[t].end());
        uint64_t const * s = reinterpret_cast<uint64_t const *>(&stats[t]);
        uint64_t *       d = reinterpret_cast<uint64_t *>(st);
        for (int k = 0; k < 16; ++k) d[k] += s[k];
    }
    *out  = h->hits.data();
    *nOut = h->hits.size();
    return 0;
}

} // extern "C"

extern "C" int orc_format_m8(uint32_t domain, lgpu_hit const * h, char const * qId, char const * sId, char * buf,
                             size_t cap)
{
    return lgpu::formatM8(domain, *h, qId, std::strlen(qId), sId, std::strlen(sId), buf, cap);
}

extern "C" int orc_params_default(lgpu_params * p, uint32_t domain, char const * profile)
{
    return lgpu::paramsDefault(*p, domain, profile);
}
