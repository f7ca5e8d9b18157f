// Host-side per-query finalisation and BLAST-tabular formatting.
//
//   writeRecords / _writeRecord   reference src/search_algo.hpp:1335-1362 / :821-913
//   m8 "std" columns              SQ/blast/blast_tabular_out.h:248-400
//   coordinate un-translation     SQ/blast/blast_base.h:337-372
// These stay on the host in the reference's design as well.
#pragma once

#include <algorithm>
#include <cstdio>
#include <cstring>
#include <tuple>
#include <unordered_set>
#include <vector>

#include "../../include/lambda_b200.h"

namespace lgpu
{

// Sort / unique / rank / truncate the hits of every query like _writeRecord does, in place.
// `hits` must be grouped by q_id (any order inside a group).  Returns the new size.
inline size_t finalizeRecords(std::vector<lgpu_hit> & hits, uint32_t maxMatches, lgpu_stats & st)
{
    // group by query, keeping the arrival order inside a group irrelevant: the comparators below
    // define a total order up to fully identical records
    auto const byQuery = [](lgpu_hit const & a, lgpu_hit const & b) { return a.q_id < b.q_id; };
    if (!std::is_sorted(hits.begin(), hits.end(), byQuery)) // the device emits hits grouped by query already
        std::stable_sort(hits.begin(), hits.end(), byQuery);
    size_t out = 0;
    size_t i   = 0;
    std::vector<lgpu_hit>        rec;
    std::unordered_set<uint32_t> uniq;
    while (i < hits.size())
    {
        size_t j = i;
        while (j < hits.size() && hits[j].q_id == hits[i].q_id)
            ++j;
        ++st.qrys_with_hit;
        if (j - i == 1) // the common case: nothing to sort, deduplicate or truncate
        {
            ++st.hits_final;
            ++st.pairs;
            hits[out++] = hits[i];
            i           = j;
            continue;
        }
        rec.assign(hits.begin() + i, hits.begin() + j);
        size_t const before = rec.size();
        // bitScore compared inverted so that larger scores come first among equal coordinates
        std::stable_sort(rec.begin(), rec.end(), [](lgpu_hit const & m1, lgpu_hit const & m2) {
            return std::tie(m1.s_id, m1.q_start, m1.q_end, m1.s_start, m1.s_end, m1.q_frame, m1.s_frame, m2.bit_score) <
                   std::tie(m2.s_id, m2.q_start, m2.q_end, m2.s_start, m2.s_end, m2.q_frame, m2.s_frame, m1.bit_score);
        });
        rec.erase(std::unique(rec.begin(), rec.end(),
                              [](lgpu_hit const & m1, lgpu_hit const & m2) {
                                  return std::tie(m1.s_id, m1.q_start, m1.q_end, m1.s_start, m1.s_end, m1.q_frame,
                                                  m1.s_frame) ==
                                         std::tie(m2.s_id, m2.q_start, m2.q_end, m2.s_start, m2.s_end, m2.q_frame,
                                                  m2.s_frame);
                              }),
                  rec.end());
        st.hits_duplicate2 += before - rec.size();
        std::stable_sort(rec.begin(), rec.end(),
                         [](lgpu_hit const & m1, lgpu_hit const & m2) { return m1.bit_score > m2.bit_score; });
        if (rec.size() > maxMatches)
        {
            st.hits_abundant += rec.size() - maxMatches;
            rec.resize(maxMatches);
        }
        st.hits_final += rec.size();
        uniq.clear();
        for (auto const & h : rec)
            uniq.insert(h.s_id);
        st.pairs += uniq.size();
        for (auto const & h : rec)
            hits[out++] = h;
        i = j;
    }
    hits.resize(out);
    return out;
}

inline char const * evalueFormat(double e)
{
    if (e < 1.0e-180) return "%3.1lf";
    if (e < 1.0e-99) return "%2.0le";
    if (e < 0.0009) return "%3.0le";
    if (e < 0.1) return "%4.3lf";
    if (e < 1.0) return "%3.2lf";
    if (e < 10.0) return "%2.1lf";
    return "%5.0lf";
}

inline char const * bitScoreFormat(double b)
{
    if (b > 9999) return "%4.3le";
    if (b > 99.9) return "%4.0lf";
    return "%4.1lf";
}

inline size_t idPrefixLen(char const * id, size_t len)
{
    void const * sp = std::memchr(id, ' ', len);
    return sp ? static_cast<size_t>(static_cast<char const *>(sp) - id) : len;
}

// One m8 line: qseqid sseqid pident length mismatch gapopen qstart qend sstart send evalue bitscore
inline int formatM8(uint32_t domain, lgpu_hit const & h, char const * qId, size_t qIdLen, char const * sId,
                    size_t sIdLen, char * buf, size_t cap)
{
    // _untranslateQPositions / _untranslateSPositions (SQ/blast/blast_base.h:337-420): codon positions of
    // translated frames back to nucleotides, reverse-strand coordinates flipped, starts 1-based.  Which
    // program ran is visible from the hit: only translated queries / subjects of a protein search carry
    // a non-zero frame, nucleotide and bisulfite queries have a strand (and BLASTN subjects none).
    uint64_t qs = h.q_start, qe = h.q_end, ss = h.s_start, se = h.s_end;
    auto const untranslate = [](uint64_t & b, uint64_t & e, int frame, uint64_t len, bool hasFrames) {
        if (hasFrames)
        {
            uint64_t const shift = static_cast<uint64_t>(frame < 0 ? -frame : frame) - 1;
            b                    = b * 3 + shift;
            e                    = e * 3 + shift;
        }
        if (frame > 0)
            ++b;
        else
        {
            b = len - b;
            e = len - e + 1;
        }
    };
    bool const qHasRevComp = domain != LGPU_DOMAIN_PROTEIN || h.q_frame != 0;
    if (qHasRevComp)
        untranslate(qs, qe, h.q_frame, h.q_len, domain == LGPU_DOMAIN_PROTEIN);
    else
        ++qs;
    if (domain == LGPU_DOMAIN_PROTEIN && h.s_frame != 0)
        untranslate(ss, se, h.s_frame, h.s_len, true);
    else
        ++ss; // BLASTN / BLASTP / BLASTX subjects have neither frames nor reverse complement
    float const identity = static_cast<float>(100.0 * static_cast<float>(h.n_match) / static_cast<float>(h.aln_len));
    char        ev[64], bs[64];
    std::snprintf(ev, sizeof(ev), evalueFormat(h.evalue), h.evalue);
    std::snprintf(bs, sizeof(bs), bitScoreFormat(h.bit_score), h.bit_score);
    int const n = std::snprintf(buf, cap, "%.*s\t%.*s\t%.2f\t%u\t%u\t%u\t%llu\t%llu\t%llu\t%llu\t%s\t%s\n",
                                static_cast<int>(idPrefixLen(qId, qIdLen)), qId,
                                static_cast<int>(idPrefixLen(sId, sIdLen)), sId, identity, h.aln_len, h.n_mismatch,
                                h.n_gap_open, static_cast<unsigned long long>(qs), static_cast<unsigned long long>(qe),
                                static_cast<unsigned long long>(ss), static_cast<unsigned long long>(se), ev, bs);
    return n;
}

} // namespace lgpu
