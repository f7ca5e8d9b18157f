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

// ---- BLAST tabular columns (--output-columns; SQ/blast/blast_tabular.h:403-454 optionLabels, :471-... columnLabels,
// :591-642 implemented) -- the index of a column is its position in the reference's BlastMatchField::Enum
constexpr int kNumTabColumns = 47;
inline char const * const * tabColumnOptionLabels()
{
    static char const * const l[kNumTabColumns] = {
      "std", "qseqid", "qgi", "qacc", "qaccver", "qlen", "sseqid", "sallseqid", "sgi", "sallgi", "sacc", "saccver",
      "sallacc", "slen", "qstart", "qend", "sstart", "send", "qseq", "sseq", "evalue", "bitscore", "score", "length",
      "pident", "nident", "mismatch", "positive", "gapopen", "gaps", "ppos", "frames", "qframe", "sframe", "btop",
      "staxids", "sscinames", "scomnames", "sblastnames", "sskingdoms", "stitle", "salltitles", "sstrand", "qcovs",
      "qcovhsp", "lcaid", "lcataxid"};
    return l;
}
inline char const * const * tabColumnLabels()
{
    static char const * const l[kNumTabColumns] = {
      "query id, subject id, % identity, alignment length, mismatches, gap opens, q. start, q. end, s. start, s. end, "
      "evalue, bit score",
      "query id", "query gi", "query acc.", "query acc.ver", "query length", "subject id", "subject ids", "subject gi",
      "subject gis", "subject acc.", "subject acc.ver", "subject accs.", "subject length", "q. start", "q. end",
      "s. start", "s. end", "query seq", "subject seq", "evalue", "bit score", "score", "alignment length",
      "% identity", "identical", "mismatches", "positives", "gap opens", "gaps", "% positives", "query/sbjct frames",
      "query frame", "sbjct frame", "BTOP", "subject tax ids", "subject sci names", "subject com names",
      "subject blast names", "subject super kingdoms", "subject title", "subject titles", "subject strand",
      "% subject coverage", "% hsp coverage", "lca id", "lca tax id"};
    return l;
}
// BlastMatchField::implemented (SQ/blast/blast_tabular.h:591-642): the columns whose value the writer prints; the
// others print "n/i"
inline bool tabColumnImplemented(uint32_t c)
{
    static bool const impl[kNumTabColumns] = {true,  true,  false, true,  false, true,  true,  false, false, false, true,  false,
                                              true,  true,  true,  true,  true,  true,  false, false, true,  true,  true,  true,
                                              true,  true,  true,  true,  true,  true,  true,  true,  true,  true,  false, true,
                                              false, false, false, false, false, false, false, false, false, true,  true};
    return c < static_cast<uint32_t>(kNumTabColumns) && impl[c];
}
enum TabColumn : uint32_t
{
    TAB_STD = 0, TAB_Q_SEQ_ID = 1, TAB_Q_ACC = 3, TAB_Q_LEN = 5, TAB_S_SEQ_ID = 6, TAB_S_ACC = 10, TAB_S_ALLACC = 12,
    TAB_S_LEN = 13, TAB_Q_START = 14, TAB_Q_END = 15, TAB_S_START = 16, TAB_S_END = 17, TAB_E_VALUE = 20,
    TAB_BIT_SCORE = 21, TAB_SCORE = 22, TAB_LENGTH = 23, TAB_P_IDENT = 24, TAB_N_IDENT = 25, TAB_MISMATCH = 26,
    TAB_POSITIVE = 27, TAB_GAP_OPEN = 28, TAB_GAPS = 29, TAB_P_POS = 30, TAB_FRAMES = 31, TAB_Q_FRAME = 32,
    TAB_S_FRAME = 33, TAB_S_TAX_IDS = 35, TAB_LCA_ID = 45, TAB_LCA_TAX_ID = 46
};

// One tabular line with the given columns (SQ/blast/blast_tabular_out.h:248-400 _writeField, :514-560 _writeMatch):
// "std" expands to qseqid sseqid pident length mismatch gapopen qstart qend sstart send evalue bitscore; columns
// the reference does not implement print "n/i"; accessions are never filled by the reference's search ("n/a").
// Taxonomy columns (staxids, lcaid, lcataxid) need the index's taxonomy, which this library does not load:
// the caller rejects them (returns -1 here).
inline int formatTabular(uint32_t domain, lgpu_hit const & h, char const * qId, size_t qIdLen, char const * sId,
                         size_t sIdLen, uint32_t const * cols, size_t nCols, char * buf, size_t cap)
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
    // frames as the tabular writer prints them: 0 where the program has a single frame (BLASTN subjects,
    // untranslated proteins); bisulfite runs as BLASTN there
    int const qFrame = h.q_frame;
    int const sFrame = domain == LGPU_DOMAIN_PROTEIN ? h.s_frame : 0;

    size_t pos = 0;
    bool   ok  = true;
    auto   put = [&](char const * fmt, auto... args) {
        if (pos >= cap)
        {
            ok = false;
            return;
        }
        int const n = std::snprintf(buf + pos, cap - pos, fmt, args...);
        if (n < 0 || static_cast<size_t>(n) >= cap - pos)
            ok = false;
        else
            pos += static_cast<size_t>(n);
    };
    bool first = true;
    auto field = [&](uint32_t c) -> bool {
        if (!first)
            put("%s", "\t");
        first = false;
        switch (c)
        {
            case TAB_Q_SEQ_ID: put("%.*s", static_cast<int>(idPrefixLen(qId, qIdLen)), qId); break;
            case TAB_S_SEQ_ID: put("%.*s", static_cast<int>(idPrefixLen(sId, sIdLen)), sId); break;
            case TAB_Q_ACC:
            case TAB_S_ACC:
            case TAB_S_ALLACC: put("%s", "n/a"); break;
            case TAB_Q_LEN: put("%u", h.q_len); break;
            case TAB_S_LEN: put("%u", h.s_len); break;
            case TAB_Q_START: put("%llu", static_cast<unsigned long long>(qs)); break;
            case TAB_Q_END: put("%llu", static_cast<unsigned long long>(qe)); break;
            case TAB_S_START: put("%llu", static_cast<unsigned long long>(ss)); break;
            case TAB_S_END: put("%llu", static_cast<unsigned long long>(se)); break;
            case TAB_E_VALUE: put(evalueFormat(h.evalue), h.evalue); break;
            case TAB_BIT_SCORE: put(bitScoreFormat(h.bit_score), h.bit_score); break;
            case TAB_SCORE: put("%d", h.score); break;
            case TAB_LENGTH: put("%u", h.aln_len); break;
            case TAB_P_IDENT:
            {
                float const identity = static_cast<float>(100.0 * static_cast<float>(h.n_match) / static_cast<float>(h.aln_len));
                put("%.2f", identity);
                break;
            }
            case TAB_N_IDENT: put("%u", h.n_match); break;
            case TAB_MISMATCH: put("%u", h.n_mismatch); break;
            case TAB_POSITIVE: put("%u", h.n_positive); break;
            case TAB_GAP_OPEN: put("%u", h.n_gap_open); break;
            case TAB_GAPS: put("%u", h.n_gap_open + h.n_gap_ext); break;
            case TAB_P_POS:
            {
                float const sim = static_cast<float>(100.0 * static_cast<float>(h.n_positive) / static_cast<float>(h.aln_len));
                put("%.2f", static_cast<double>(sim));
                break;
            }
            case TAB_FRAMES: put("%i/%i", qFrame, sFrame); break;
            case TAB_Q_FRAME: put("%i", qFrame); break;
            case TAB_S_FRAME: put("%i", sFrame); break;
            case TAB_S_TAX_IDS:
            case TAB_LCA_ID:
            case TAB_LCA_TAX_ID: return false;
            default: put("%s", "n/i"); break; // not implemented by the reference either
        }
        return true;
    };
    static uint32_t const stdCols[12] = {TAB_Q_SEQ_ID, TAB_S_SEQ_ID, TAB_P_IDENT, TAB_LENGTH, TAB_MISMATCH, TAB_GAP_OPEN,
                                         TAB_Q_START, TAB_Q_END, TAB_S_START, TAB_S_END, TAB_E_VALUE, TAB_BIT_SCORE};
    for (size_t i = 0; i < nCols; ++i)
    {
        if (cols[i] >= static_cast<uint32_t>(kNumTabColumns))
            return -1;
        if (cols[i] == TAB_STD)
        {
            for (uint32_t c : stdCols)
                field(c);
        }
        else if (!field(cols[i]))
            return -1;
    }
    put("%s", "\n");
    return ok ? static_cast<int>(pos) : -2;
}

// One m8 line: qseqid sseqid pident length mismatch gapopen qstart qend sstart send evalue bitscore
inline int formatM8(uint32_t domain, lgpu_hit const & h, char const * qId, size_t qIdLen, char const * sId,
                    size_t sIdLen, char * buf, size_t cap)
{
    uint32_t const std0 = TAB_STD;
    return formatTabular(domain, h, qId, qIdLen, sId, sIdLen, &std0, 1, buf, cap);
}

} // namespace lgpu
