// Host-side option handling, scoring tables and BLAST statistics.
//
// Mirrors the parts of the reference the hot path reads:
//   per-domain defaults + profiles  src/search_options.hpp:263,290-337,631-682
//   prepareScoring                  src/search_algo.hpp:166-234
//   Karlin-Altschul values          SQ/blast/blast_statistics.h:92-361, _selectSet :560-624
//   _lengthAdjustment               SQ/blast/blast_statistics.h:903-982
//   computeBitScore / _computeEValue  :1027-1030 / :1081-1087
//   computeEValueThreadSafe         src/search_misc.hpp:57-80
// All floating point of the path lives here, on the host, in double like the reference; the device
// only ever sees integer score thresholds derived from these functions.
#pragma once

#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <string>
#include <unordered_map>
#include <vector>

#include "../../include/lambda_b200.h"

namespace lgpu
{

#include "tables_generated.inc"

inline int paramsDefault(lgpu_params & p, uint32_t domain, char const * profileC)
{
    std::string const profile = profileC ? profileC : "none";
    std::memset(&p, 0, sizeof(p));
    p.domain             = domain;
    p.seed_half_exact    = 1;
    p.adaptive_seeding   = 1;
    p.iterative_search   = 1;
    p.max_matches        = 25;
    p.pre_scoring        = 2;
    p.pre_scoring_thresh = 2.0;
    p.scoring_method     = 62;
    p.match              = 2;
    p.mismatch           = -3;
    p.min_bit_score      = -1;
    p.max_evalue         = 1e-2;
    p.id_cutoff          = 0;
    p.finalize           = 1;
    switch (domain)
    {
        case LGPU_DOMAIN_PROTEIN:
            p.gap_open   = -11;
            p.gap_extend = -1;
            p.opts0      = {10, 0, 5};
            p.opts       = {11, 1, 3};
            break;
        case LGPU_DOMAIN_NUCLEOTIDE:
            p.gap_open           = -5;
            p.gap_extend         = -2;
            p.opts0              = {14, 0, 9};
            p.opts               = {14, 1, 7};
            p.pre_scoring_thresh = 1.4;
            break;
        case LGPU_DOMAIN_BISULFITE:
            p.gap_open           = -5;
            p.gap_extend         = -2;
            p.max_evalue         = 1e-9;
            p.opts0              = {17, 0, 10};
            p.opts               = {17, 1, 10};
            p.pre_scoring_thresh = 1.5;
            break;
        default: return LGPU_ERR_ARG;
    }

    if (profile == "none" || profile.empty())
    {
    }
    else if (profile == "fast")
    {
        if (domain != LGPU_DOMAIN_PROTEIN)
        {
            p.iterative_search   = 0;
            p.opts.max_seed_dist = 0;
            if (domain == LGPU_DOMAIN_NUCLEOTIDE)
                p.opts.seed_offset = 9;
        }
        else
        {
            p.opts0.seed_length  = 12;
            p.opts0.seed_offset  = 8;
            p.opts.seed_length   = 10;
            p.opts.seed_offset   = 5;
            p.opts.max_seed_dist = 0;
        }
    }
    else if (profile == "sensitive" || profile.rfind("pairs", 0) == 0)
    {
        if (profile != "sensitive" && profile != "pairs-default" && profile != "pairs-sensitive")
            return LGPU_ERR_ARG;
        switch (domain)
        {
            case LGPU_DOMAIN_PROTEIN:
                p.opts0.seed_length  = 9;
                p.opts0.seed_offset  = 4;
                p.opts.seed_length   = 8;
                p.opts.seed_offset   = 3;
                p.pre_scoring        = 3;
                p.pre_scoring_thresh = 1.9;
                break;
            case LGPU_DOMAIN_NUCLEOTIDE:
                p.opts0.seed_offset = 3;
                p.opts.seed_offset  = 3;
                break;
            case LGPU_DOMAIN_BISULFITE:
                p.opts0.seed_length = 16;
                p.opts0.seed_offset = 8;
                p.opts.seed_length  = 15;
                p.opts.seed_offset  = 10;
                break;
        }
        if (profile.rfind("pairs", 0) == 0)
            p.iterative_search = 0;
        if (profile == "pairs-sensitive")
            --p.opts.seed_length;
    }
    else
    {
        return LGPU_ERR_ARG;
    }
    return LGPU_OK;
}

// ---------------------------------------------------------------------------------------------
// scoring
// ---------------------------------------------------------------------------------------------

struct KarlinAltschul
{
    double lambda = 0, K = 0, H = 0, alpha = 0, beta = 0;
    bool   valid = false;
};

// Scoring as the device consumes it: a dense int8 matrix over the *translated* alphabet in BioC++
// rank order (27x27 amino acids, 5x5 dna5), SeqAn-convention gap scores, and the KA parameters.
struct Scoring
{
    int            alphSize = 0;      // 27 or 5
    int8_t         matrix[32 * 32]{}; // [query residue * 32 + subject residue]
    int8_t         matrixRev[32 * 32]{}; // bisulfite: matrix for reverse-converted (odd) subjects; else = matrix
    int            gapOpenSeqan = 0;  // cost of the first gap character = gapOpen + gapExtend
    int            gapExtend    = 0;
    KarlinAltschul ka;
};

inline KarlinAltschul selectKA(lgpu_params const & p)
{
    KarlinAltschul ka;
    if (p.domain == LGPU_DOMAIN_PROTEIN)
    {
        double const(*tab)[8] = nullptr;
        int n                 = 0;
        switch (p.scoring_method)
        {
            case 45: tab = kKaBlosum45; n = sizeof(kKaBlosum45) / sizeof(kKaBlosum45[0]); break;
            case 62: tab = kKaBlosum62; n = sizeof(kKaBlosum62) / sizeof(kKaBlosum62[0]); break;
            case 80: tab = kKaBlosum80; n = sizeof(kKaBlosum80) / sizeof(kKaBlosum80[0]); break;
            default: return ka;
        }
        for (int i = 0; i < n; ++i)
            if (tab[i][0] == -p.gap_open && tab[i][1] == -p.gap_extend)
            {
                ka = {tab[i][3], tab[i][4], tab[i][5], tab[i][6], tab[i][7], true};
                break;
            }
    }
    else
    {
        int const n = sizeof(kKaNucl) / sizeof(kKaNucl[0]);
        for (int i = 0; i < n; ++i)
            if (kKaNucl[i][0] == p.match && kKaNucl[i][1] == -p.mismatch && kKaNucl[i][2] == -p.gap_open &&
                kKaNucl[i][3] == -p.gap_extend)
            {
                ka = {kKaNucl[i][4], kKaNucl[i][5], kKaNucl[i][6], kKaNucl[i][7], kKaNucl[i][8], true};
                break;
            }
    }
    return ka;
}

inline int makeScoring(Scoring & s, lgpu_params const & p)
{
    s = Scoring{};
    if (p.domain == LGPU_DOMAIN_PROTEIN)
    {
        int8_t const(*m)[27] = nullptr;
        switch (p.scoring_method)
        {
            case 45: m = kBlosum45; break;
            case 62: m = kBlosum62; break;
            case 80: m = kBlosum80; break;
            default: return LGPU_ERR_ARG;
        }
        s.alphSize = 27;
        for (int a = 0; a < 27; ++a)
            for (int b = 0; b < 27; ++b)
                s.matrix[a * 32 + b] = m[a][b];
    }
    else if (p.domain == LGPU_DOMAIN_NUCLEOTIDE)
    {
        // plain nucleotide scoring compares ranks: N vs N is a match
        // (SQ/score/score_simd_wrapper.h:149-154; scalar Score<int,Simple>)
        s.alphSize = 5;
        for (int a = 0; a < 5; ++a)
            for (int b = 0; b < 5; ++b)
                s.matrix[a * 32 + b] = static_cast<int8_t>(a == b ? p.match : p.mismatch);
    }
    else if (p.domain == LGPU_DOMAIN_BISULFITE)
    {
        // setScoreBisulfiteMatrix (src/bisulfite_scoring.hpp:67-93) is written over SeqAn's Dna5 order
        // (A,C,G,T,N); our ranks are BioC++ dna5 (A,C,G,N,T; src/seqan2_to_biocpp.hpp:360-364).
        // forward: a read T over a reference C is a match; reverse: a read A over a reference G; N never matches
        s.alphSize              = 5;
        int const toSeqan[5]    = {0, 1, 2, 4, 3};
        for (int a = 0; a < 5; ++a)
            for (int b = 0; b < 5; ++b)
            {
                int const i = toSeqan[a], j = toSeqan[b];
                bool const fwd = ((i == j) || (i == 3 && j == 1)) && i != 4;
                bool const rev = ((i == j) || (i == 0 && j == 2)) && i != 4;
                s.matrix[a * 32 + b]    = static_cast<int8_t>(fwd ? p.match : p.mismatch);
                s.matrixRev[a * 32 + b] = static_cast<int8_t>(rev ? p.match : p.mismatch);
            }
    }
    else
    {
        return LGPU_ERR_ARG;
    }
    if (p.domain != LGPU_DOMAIN_BISULFITE)
        std::memcpy(s.matrixRev, s.matrix, sizeof(s.matrix));
    s.gapOpenSeqan = p.gap_open + p.gap_extend; // src/search_algo.hpp:226
    s.gapExtend    = p.gap_extend;
    s.ka           = selectKA(p);
    if (!s.ka.valid)
        return LGPU_ERR_ARG; // "Could not compute Karlin-Altschul-Values for Scoring Scheme."
    return LGPU_OK;
}

// ---------------------------------------------------------------------------------------------
// statistics
// ---------------------------------------------------------------------------------------------

// NCBI's length-adjustment iteration as SeqAn ships it, including the early return that makes the
// final refinement unreachable (blast_statistics.h:962).
inline uint64_t lengthAdjustment(uint64_t dbLength, uint64_t queryLength, KarlinAltschul const & ka)
{
    double const logK          = std::log(ka.K);
    double const alphaByLambda = ka.alpha / ka.lambda;
    double const n             = static_cast<double>(dbLength);
    double const m             = static_cast<double>(queryLength);
    double       val = 0, valMin = 0, valMax;

    double const mb = m + n;
    double const c  = n * m - std::max(m, n) / ka.K;
    if (c < 0)
        return 0;
    valMax = 2 * c / (mb + std::sqrt(mb * mb - 4 * c));

    for (int i = 1; i <= 20; ++i)
    {
        double const totalLen = (m - val) * (n - val);
        double const valNew   = alphaByLambda * (logK + std::log(totalLen)) + ka.beta;
        if (valNew >= val)
        {
            valMin = val;
            if (valNew - valMin <= 1.0)
                break; // converged
            if (valMin == valMax)
                break;
        }
        else
        {
            valMax = val;
        }
        if (valMin <= valNew && valNew <= valMax)
            val = valNew;
        else
            val = (i == 1) ? valMax : (valMin + valMax) / 2;
    }
    return static_cast<uint64_t>(valMin);
}

inline double bitScore(KarlinAltschul const & ka, double rawScore)
{
    return (ka.lambda * rawScore - std::log(ka.K)) / std::log(2);
}

// computeEValueThreadSafe with its per-length cache; `queryLen` is the ORIGINAL query length
// (bm.qLength), divided by 3 for translated queries.  Per query length we cache, besides the length
// adjustment, the prefix K * (m - adj) * (n - adj) in the reference's evaluation order, so an e-value is
// one exp() and one multiplication with bit-identical results.
class EValueComputer
{
public:
    EValueComputer(KarlinAltschul const & ka, uint64_t dbTotalLength, bool qIsTranslated) :
      ka_(ka), dbLen_(dbTotalLength), div_(qIsTranslated ? 3 : 1), lnK_(std::log(ka.K)), ln2_(std::log(2))
    {}

    uint64_t adjustment(uint64_t queryLen) { return entry(queryLen).adj; }

    double evalue(int32_t rawScore, uint64_t queryLen)
    {
        return entry(queryLen).prefix * std::exp(-ka_.lambda * static_cast<double>(rawScore));
    }

    double bits(int32_t rawScore) const { return (ka_.lambda * static_cast<double>(rawScore) - lnK_) / ln2_; }

    // the same values through per-score tables (a record costs two loads instead of an exp and a division)
    double bitsCached(int32_t rawScore)
    {
        if (rawScore < 0 || rawScore >= kTable)
            return bits(rawScore);
        fillTables();
        return bitsTab_[rawScore];
    }
    double evalueCached(int32_t rawScore, uint64_t queryLen)
    {
        if (rawScore < 0 || rawScore >= kTable)
            return evalue(rawScore, queryLen);
        fillTables();
        return entry(queryLen).prefix * expTab_[rawScore];
    }

    KarlinAltschul const & ka() const { return ka_; }
    uint64_t               dbLen() const { return dbLen_; }

private:
    static constexpr int32_t kTable = 1 << 15;
    std::vector<double>      bitsTab_, expTab_;
    void                     fillTables()
    {
        if (!bitsTab_.empty())
            return;
        bitsTab_.resize(kTable);
        expTab_.resize(kTable);
        for (int32_t s = 0; s < kTable; ++s)
        {
            bitsTab_[s] = bits(s);
            expTab_[s]  = std::exp(-ka_.lambda * static_cast<double>(s));
        }
    }
    struct Entry
    {
        uint64_t adj;
        double   prefix; // (K * m) * n with m = ql - adj, n = dbLen - adj (unsigned subtraction, then double)
    };

    Entry const & entry(uint64_t queryLen)
    {
        if (queryLen == lastLen_)
            return last_;
        uint64_t const ql = queryLen / div_;
        auto           it = cache_.find(ql);
        if (it == cache_.end())
        {
            Entry e;
            e.adj          = lengthAdjustment(dbLen_, ql, ka_);
            double const m = static_cast<double>(ql - e.adj);
            double const n = static_cast<double>(dbLen_ - e.adj);
            e.prefix       = ka_.K * m * n;
            it             = cache_.emplace(ql, e).first;
        }
        lastLen_ = queryLen;
        last_    = it->second;
        return last_;
    }

    KarlinAltschul                      ka_;
    uint64_t                            dbLen_;
    uint64_t                            div_;
    double                              lnK_, ln2_;
    std::unordered_map<uint64_t, Entry> cache_;
    uint64_t                            lastLen_ = ~0ull;
    Entry                               last_{};
};

// Integer thresholds that reproduce the reference's two pass-1 filters exactly
// (src/search_algo.hpp:1256-1278): a hit is dropped by the bit-score test iff score < minBit, else
// by the e-value test iff score < minEval.  Both tests are monotone in the raw score for a fixed
// query length, so each reduces to one integer compare on the device.
struct ScoreThresholds
{
    int32_t minBit;  // smallest S with bitScore(S) >= minBitScore   (INT32_MIN if test disabled)
    int32_t minEval; // smallest S with evalue(S)  <= maxEValue      (INT32_MIN if test disabled)
};

inline ScoreThresholds scoreThresholds(lgpu_params const & p, EValueComputer & ev, uint64_t queryLen)
{
    constexpr int32_t kMaxScore = 1 << 20;
    ScoreThresholds   t{INT32_MIN, INT32_MIN};
    if (p.min_bit_score >= 0)
    {
        int32_t lo = 0, hi = kMaxScore; // first S in [lo,hi] with bits >= min
        while (lo < hi)
        {
            int32_t const mid = lo + (hi - lo) / 2;
            if (bitScore(ev.ka(), mid) < p.min_bit_score)
                lo = mid + 1;
            else
                hi = mid;
        }
        t.minBit = lo;
    }
    if (p.max_evalue >= 0)
    {
        int32_t lo = 0, hi = kMaxScore;
        while (lo < hi)
        {
            int32_t const mid = lo + (hi - lo) / 2;
            if (ev.evalue(mid, queryLen) > p.max_evalue)
                lo = mid + 1;
            else
                hi = mid;
        }
        t.minEval = lo;
    }
    return t;
}

// ---------------------------------------------------------------------------------------------
// domain-derived constants
// ---------------------------------------------------------------------------------------------

// Which BLAST program a (domain, index, query alphabet) combination is, and what follows from it
// (GlobalDataHolder::blastProgram / qryNumFrames / sbjNumFrames, src/search_datastructures.hpp:326-385).
struct DomainInfo
{
    uint32_t qryNumFrames  = 1;
    uint32_t sbjNumFrames  = 1;
    bool     qIsTranslated = false; // BLASTX / TBLASTX
    bool     sIsTranslated = false; // TBLASTN / TBLASTX
    uint8_t  unknownRank   = 23;    // 'X' in aa27, 'N' (3) in dna5  (src/search_algo.hpp:652-656)
    uint32_t qFrameMode    = 0;     // DevQueries::frameMode
};

// `sbjOrigAlph` = alphabet of the sequences stored in the index (LGPU_ALPH_*), `qryAlph` = alphabet of the
// query batch (0 = the domain's default: amino acids for protein searches, dna5 otherwise)
inline DomainInfo domainInfo(uint32_t domain, uint32_t sbjOrigAlph = 0, uint32_t qryAlph = 0)
{
    DomainInfo d;
    switch (domain)
    {
        case LGPU_DOMAIN_PROTEIN:
            d.qIsTranslated = qryAlph == LGPU_ALPH_DNA5;
            d.sIsTranslated = sbjOrigAlph == LGPU_ALPH_DNA5;
            d.qryNumFrames  = d.qIsTranslated ? 6 : 1;
            d.sbjNumFrames  = d.sIsTranslated ? 6 : 1;
            d.unknownRank   = 23;
            d.qFrameMode    = d.qIsTranslated ? 2 : 0;
            break;
        case LGPU_DOMAIN_NUCLEOTIDE: d = {2, 1, false, false, 3, 1}; break;
        case LGPU_DOMAIN_BISULFITE: d = {4, 2, false, false, 3, 3}; break;
    }
    return d;
}

// translated-frame length (BIO ranges/views/translate_single.hpp:95-113)
inline uint64_t translatedFrameLength(uint64_t len, uint32_t frame)
{
    uint64_t const o = frame % 3;
    return (std::max(len, o) - o) / 3;
}

} // namespace lgpu
