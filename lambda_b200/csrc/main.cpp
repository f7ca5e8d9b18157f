// lambda3_b200 -- host program with the lambda3 searchp / searchn / searchbs command line.
//
// Replaces the loop body of the reference's realMain() (src/search.cpp:345-477) by calls through the C
// ABI (include/lambda_b200.h): load the reference's own .lba index, read the query FASTA, hand large
// query blocks to the device engine, write BLAST tabular (.m8).  Flags, defaults and profiles follow
// src/search_options.hpp; only the options that influence the hot path are accepted (output formats
// other than .m8, taxonomy/LCA and lazy query loading stay with the reference).
//
// Record order: the reference with -t 1 processes batches of `records_per_batch` queries
// (src/search_algo.hpp:354-356) and writes, per batch, the phase-1 successes in query order followed by
// that batch's phase-2 hits (src/search.cpp:449-457).  We reproduce exactly that order, so the file is
// byte-identical to `lambda3 search* -t 1`, not merely equal as a multiset.
#include <algorithm>
#include <atomic>
#include <cctype>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <string>
#include <thread>
#include <vector>

#include <cmath>
#include <zlib.h>

#include "../../include/lambda_b200.h"
#include "query_reader.hpp"
#include "tables_generated.inc" // alphabets and the genetic code, for the report-style outputs

namespace
{

// SamBamExtraTags (src/search_output.hpp:29-76), in enum order
enum SamTag { ST_AS, ST_OC, ST_NM, ST_IH, ST_ar, ST_ae, ST_ai, ST_ap, ST_qf, ST_qs, ST_sf, ST_st, ST_ls, ST_lt };
char const * const kSamTagKeys[14]  = {"AS", "OC", "NM", "IH", "ar", "ae", "ai", "ap", "qf", "qs", "sf", "st", "ls", "lt"};
char const * const kSamTagDescr[14] = {"bit score",
                                       "query protein cigar (* for BLASTN)",
                                       "edit distance (in protein space unless BLASTN)",
                                       "number of matches this query has",
                                       "raw score",
                                       "expect value",
                                       "% identity (in protein space unless BLASTN) ",
                                       "% positive (in protein space unless BLASTN)",
                                       "query frame",
                                       "query protein sequence (* for BLASTN)",
                                       "subject frame",
                                       "subject taxonomy IDs (* if n/a)",
                                       "lowest common ancestor scientific name",
                                       "lowest common ancestor taxonomy ID"};

struct Options
{
    uint32_t    domain = LGPU_DOMAIN_PROTEIN;
    std::string query, index, output = "output.m8", profile = "none", inputAlphabet = "auto";
    int         verbosity = 1;
    int         threads   = 1; // only used for the reference's records_per_batch formula
    int         gpus      = 1;
    bool        gzOutput  = false; // output path ends in .gz: the text is gzip-compressed
    bool        comments  = false; // .m9: BLAST tabular with comment lines
    // SAM / BAM dialect (src/search_options.hpp:276-370,765-823): tags in the order of SamBamExtraTags::Enum
    bool        samTags[14] = {true, false, true, false, false, true, true, false, true, false, false, false, false, false};
    bool        samWithRefHeader = false; // --sam-with-refheader: @SQ lines in .sam (always there in .bam)
    int         samBamSeq        = 1;     // --sam-bam-seq never|uniq|always = 0|1|2
    bool        samHardClip      = true;  // --sam-bam-clip hard|soft
    bool        hasSTaxIds = false; // a taxonomy column / tag was requested (src/search_options.hpp:744-750,806-814)
    bool        computeLCA = false; // ... one that needs the lowest common ancestor of a record
    std::string replayHits;            // --replay-hits FILE: format records computed elsewhere (test hook, no search)
    std::string outputColumns = "std"; // --output-columns (.m8 / .m9)
    std::vector<uint32_t> columns;     // ... resolved to BlastMatchField indices
    bool        report    = false; // .m0: BLAST pairwise report
    bool        bam       = false; // .bam: the same records, binary + BGZF
    bool        sam       = false; // .sam (default tags AS NM ae ai qf, --sam-bam-seq uniq, --sam-bam-clip hard)
    std::string commandLine;
    bool        versionToOutput = true;
    uint64_t    blockSize = 100000;
    lgpu_params params{};
};

[[noreturn]] void die(std::string const & msg)
{
    std::fprintf(stderr, "ERROR: %s\n", msg.c_str());
    std::exit(255); // the reference returns -1
}

bool endsWith(std::string const & s, char const * suf)
{
    size_t const n = std::strlen(suf);
    return s.size() >= n && s.compare(s.size() - n, n, suf) == 0;
}

void usage()
{
    std::puts("lambda3_b200 searchp|searchn|searchbs -q QUERY.(fasta|fastq)[.gz] -i INDEX.lba [-o output.m0|.m8|.m9|.sam|.bam] [OPTIONS]\n"
              "  -a, --input-alphabet   auto|dna5|aminoacid (searchp; dna queries are translated: BLASTX/TBLASTX)\n"
              "  -p, --profile          none|fast|sensitive|pairs-default|pairs-sensitive\n"
              "      --output-columns   'std' or space-separated NCBI column specifiers (.m8 / .m9; 'help' lists them)\n"
              "      --sam-bam-tags     'AS NM ae ai qf' (default) or any of AS OC NM IH ar ae ai ap qf qs sf ('help')\n"
              "      --sam-bam-seq      always|uniq|never   --sam-bam-clip hard|soft   --sam-with-refheader 0|1\n"
              "  -e, --e-value          maximum e-value (default 0.01; -1 = off)\n"
              "      --bit-score        minimum bit score (default -1 = off)\n"
              "      --percent-identity minimum identity in percent (default 0)\n"
              "  -n, --num-matches      maximum matches per query (default 25)\n"
              "      --seed-length / --seed-offset / --seed-delta           phase-2 seeds\n"
              "      --seed-length0 / --seed-offset0 / --seed-delta0        phase-1 seeds\n"
              "      --adaptive-seeding 0|1   --seed-half-exact 0|1   --search0 0|1 (phase 1 with the ...0 seeds)\n"
              "      --pre-scoring N    --pre-scoring-threshold X\n"
              "  -s, --scoring-scheme   45|62|80 (searchp)   --score-gap  --score-gap-open\n"
              "      --score-match / --score-mismatch (searchn)\n"
              "  -t, --threads          accepted for compatibility (affects record order only)\n"
              "      --gpus N           shard the queries over N GPUs (default 1)\n"
              "      --window-band N    subject residues added on either side of a seed's window (0 = lambda3's rule\n"
              "                         floor(sqrt(query length)) + 1; other values leave parity with lambda3)\n"
              "      --block-size N     queries per device batch (default 100000)\n"
              "  -v, --verbosity        0|1|2\n");
}

void parse(int argc, char ** argv, Options & o)
{
    if (argc < 2)
    {
        usage();
        std::exit(255);
    }
    std::string const cmd = argv[1];
    if (cmd == "searchp")
        o.domain = LGPU_DOMAIN_PROTEIN;
    else if (cmd == "searchn")
        o.domain = LGPU_DOMAIN_NUCLEOTIDE;
    else if (cmd == "searchbs")
        o.domain = LGPU_DOMAIN_BISULFITE;
    else if (cmd == "-h" || cmd == "--help")
    {
        usage();
        std::exit(0);
    }
    else
        die("unknown sub-command '" + cmd + "' (mkindex* stays with the reference's lambda3)");

    // first pass: profile (defaults depend on it), then explicit overrides in order
    for (int i = 2; i + 1 < argc; ++i)
        if (!std::strcmp(argv[i], "-p") || !std::strcmp(argv[i], "--profile"))
            o.profile = argv[i + 1];
    if (lgpu_params_default(&o.params, o.domain, o.profile.c_str()) != LGPU_OK)
        die("invalid profile '" + o.profile + "'");

    auto need = [&](int & i) -> char const * {
        if (i + 1 >= argc)
            die(std::string("missing value for ") + argv[i]);
        return argv[++i];
    };
    // boolean options take 0|1|true|false like the reference's parser
    auto needBool = [&](int & i) -> bool {
        std::string const opt = argv[i], v = need(i);
        if (v == "1" || v == "true")
            return true;
        if (v == "0" || v == "false")
            return false;
        die("Value parse failed for " + opt + ": Argument " + v + " could not be parsed as type bool.");
    };
    for (int i = 2; i < argc; ++i)
    {
        std::string const a = argv[i];
        if (a == "-q" || a == "--query") o.query = need(i);
        else if (a == "-i" || a == "--index") o.index = need(i);
        else if (a == "-o" || a == "--output") o.output = need(i);
        else if (a == "-p" || a == "--profile") need(i);
        else if (a == "-a" || a == "--input-alphabet") o.inputAlphabet = need(i);
        else if (a == "--lazy-query") (void)needBool(i); // accepted: the query file is always read completely, same results
        else if (a == "--output-columns") o.outputColumns = need(i);
        else if (a == "--replay-hits")
        {
            // test hook, not a search mode: only honoured when the test harness asks for it explicitly
            o.replayHits = need(i);
            char const * e = std::getenv("LAMBDA_B200_TEST_HOOKS");
            if (!e || std::strcmp(e, "1"))
                die("--replay-hits formats records computed elsewhere and performs no search; it is a hook for the output "
                    "tests (set LAMBDA_B200_TEST_HOOKS=1)");
        }
        else if (a == "--sam-with-refheader") o.samWithRefHeader = needBool(i);
        else if (a == "--sam-bam-seq")
        {
            std::string const v = need(i);
            if (v != "always" && v != "uniq" && v != "never")
                die("Value " + v + " is not one of [always,uniq,never].");
            o.samBamSeq = v == "never" ? 0 : v == "uniq" ? 1 : 2;
        }
        else if (a == "--sam-bam-clip")
        {
            std::string const v = need(i);
            if (v != "hard" && v != "soft")
                die("Value " + v + " is not one of [hard,soft].");
            o.samHardClip = v == "hard";
        }
        else if (a == "--sam-bam-tags")
        {
            std::string const v = need(i);
            if (v == "help")
            {
                std::puts("Please specify the tags in this format -oc 'tag1 tag2', i.e. space-separated and enclosed in quotes. "
                          "The order of tags is not preserved.\nThe following specifiers are supported:");
                for (int t = 0; t < 14; ++t)
                    std::printf("\t%s\t%s\n", kSamTagKeys[t], kSamTagDescr[t]);
                std::exit(0);
            }
            for (bool & b : o.samTags)
                b = false;
            std::string tok;
            auto        flush = [&]() {
                if (tok.empty())
                    return;
                int found = -1;
                for (int t = 0; t < 14; ++t)
                    if (tok == kSamTagKeys[t])
                        found = t;
                if (found < 0)
                    die("Unknown column specifier \"" + tok + "\". Please see \"--sam-bam-tags help\" for valid options.");
                if (found >= ST_st)
                {
                    o.hasSTaxIds = true;
                    if (found != ST_st)
                        o.computeLCA = true;
                }
                o.samTags[found] = true;
                tok.clear();
            };
            for (char ch : v)
                if (std::isspace(static_cast<unsigned char>(ch)))
                    flush();
                else
                    tok.push_back(ch);
            flush();
        }
        else if (a == "-e" || a == "--e-value") o.params.max_evalue = std::atof(need(i));
        else if (a == "--bit-score") o.params.min_bit_score = std::atoi(need(i));
        else if (a == "--percent-identity") o.params.id_cutoff = std::atoi(need(i));
        else if (a == "-n" || a == "--num-matches") o.params.max_matches = static_cast<uint32_t>(std::atoi(need(i)));
        else if (a == "--seed-length") o.params.opts.seed_length = static_cast<uint32_t>(std::atoi(need(i)));
        else if (a == "--seed-offset") o.params.opts.seed_offset = static_cast<uint32_t>(std::atoi(need(i)));
        else if (a == "--seed-delta") o.params.opts.max_seed_dist = static_cast<uint32_t>(std::atoi(need(i)));
        else if (a == "--seed-length0") o.params.opts0.seed_length = static_cast<uint32_t>(std::atoi(need(i)));
        else if (a == "--seed-offset0") o.params.opts0.seed_offset = static_cast<uint32_t>(std::atoi(need(i)));
        else if (a == "--seed-delta0") o.params.opts0.max_seed_dist = static_cast<uint32_t>(std::atoi(need(i)));
        else if (a == "--adaptive-seeding") o.params.adaptive_seeding = needBool(i);
        else if (a == "--seed-half-exact") o.params.seed_half_exact = needBool(i);
        // the reference spells it --search0 (src/search_options.hpp:452-456); --iterative-search is kept as an alias
        else if (a == "--search0" || a == "--iterative-search") o.params.iterative_search = needBool(i);
        else if (a == "--pre-scoring")
        {
            o.params.pre_scoring = std::atoi(need(i));
            if (o.params.pre_scoring < 1 || o.params.pre_scoring > 10)
                die("Validation failed for option --pre-scoring: Value " + std::to_string(o.params.pre_scoring) +
                    " is not in range [1,10].");
        }
        else if (a == "--pre-scoring-threshold") o.params.pre_scoring_thresh = std::atof(need(i));
        else if (a == "-s" || a == "--scoring-scheme") o.params.scoring_method = std::atoi(need(i));
        else if (a == "--score-gap") o.params.gap_extend = std::atoi(need(i));
        else if (a == "--score-gap-open") o.params.gap_open = std::atoi(need(i));
        else if (a == "--score-match") o.params.match = std::atoi(need(i));
        else if (a == "--score-mismatch") o.params.mismatch = std::atoi(need(i));
        else if (a == "-t" || a == "--threads") o.threads = std::max(1, std::atoi(need(i)));
        else if (a == "--window-band") o.params.window_band = static_cast<uint32_t>(std::max(0, std::atoi(need(i))));
        else if (a == "--gpus") o.gpus = std::max(1, std::atoi(need(i)));
        else if (a == "--block-size") o.blockSize = std::max<uint64_t>(1, std::strtoull(need(i), nullptr, 10));
        else if (a == "-v" || a == "--verbosity") o.verbosity = std::atoi(need(i));
        else if (a == "--version-to-outputfile") o.versionToOutput = needBool(i); // .m9 comment lines
        else if (a == "-h" || a == "--help") { usage(); std::exit(0); }
        else die("unknown option '" + a + "'");
    }
    if (o.query.empty() || o.index.empty())
        die("-q and -i are required");
    if (o.inputAlphabet != "auto" && o.inputAlphabet != "dna5" && o.inputAlphabet != "aminoacid")
        die("Invalid argument to --input-alphabet");
    if (o.domain != LGPU_DOMAIN_PROTEIN && o.inputAlphabet != "auto")
        die("--input-alphabet is a searchp option");
    // the format follows the extension, looked at without a trailing ".gz" (src/search_options.hpp:210-214,684-709)
    std::string fmtPath = o.output;
    o.gzOutput          = endsWith(fmtPath, ".gz");
    if (o.gzOutput)
        fmtPath.resize(fmtPath.size() - 3);
    o.comments = endsWith(fmtPath, ".m9");
    o.bam      = endsWith(fmtPath, ".bam");
    o.sam      = endsWith(fmtPath, ".sam") || o.bam;
    o.report   = endsWith(fmtPath, ".m0");
    if (!endsWith(fmtPath, ".m8") && !o.comments && !o.sam && !o.report)
        die("supported output formats: .m0, .m8, .m9, .sam, .bam (optionally followed by .gz)");
    if (o.bam && o.gzOutput)
        die(".bam is BGZF-compressed already; .bam.gz is not supported");
    o.params.want_cigar = (o.sam || o.report) ? 1u : 0u;
    // --output-columns (src/search_options.hpp:710-760): space-separated NCBI specifiers, "std" = the default twelve
    if (o.outputColumns == "help")
    {
        std::puts("Please specify the columns in this format -oc 'column1 column2', i.e. space-separated and enclosed in "
                  "single quotes.\nThe specifiers are the same as in NCBI Blast, currently the following are supported:");
        for (uint32_t c = 0; lgpu_tabular_column_name(c); ++c)
            if (lgpu_tabular_column_implemented(c)) // like the reference: only the columns with a value
                std::printf("\t%s%s%s\n", lgpu_tabular_column_name(c), std::strlen(lgpu_tabular_column_name(c)) >= 8 ? "\t" : "\t\t",
                            lgpu_tabular_column_label(c));
        std::exit(0);
    }
    {
        std::string tok;
        auto        flush = [&]() {
            if (tok.empty())
                return;
            int const c = lgpu_tabular_column(tok.c_str());
            if (c < 0)
                die("Unknown column specifier \"" + tok + "\". Please see -oc help for valid options.");
            if (!lgpu_tabular_column_supported(static_cast<uint32_t>(c))) // staxids / lcaid / lcataxid: formatted here
            {
                o.hasSTaxIds = true;
                if (tok != "staxids")
                    o.computeLCA = true;
            }
            o.columns.push_back(static_cast<uint32_t>(c));
            tok.clear();
        };
        for (char ch : o.outputColumns)
            if (std::isspace(static_cast<unsigned char>(ch)))
                flush();
            else
                tok.push_back(ch);
        flush();
    }
    for (int i = 0; i < argc; ++i)
        o.commandLine += (i ? " " : "") + std::string(argv[i]);
    if (std::ifstream(o.output).good())
        die("the output file already exists: " + o.output); // sharg's create_new validator
}

struct Fasta
{
    std::vector<std::string> ids;
    std::vector<uint8_t>     residues;
    std::vector<uint64_t>    offsets{0};
};

// detectSeqFileAlphabet (src/shared_misc.hpp:83-110): the first sequence decides
uint32_t detectAlphabet(std::string const & path)
{
    qio::SeqReader reader(path);
    qio::Record    rec;
    if (!reader.next(rec))
        die("Your query file contains no sequences.");
    std::string const & seq = rec.seq;
    auto const allIn = [&](char const * set) {
        for (char c : seq)
            if (!std::strchr(set, c))
                return false;
        return true;
    };
    if (allIn("ACGTUNacgtun"))
        return LGPU_ALPH_DNA5;
    if (allIn("ACGTUNRYKMSWBDHVacgtunrykmswbdhv"))
    {
        std::fprintf(stderr, "\nWARNING: You query file was detected as non-standard DNA, but it could be AminoAcid, too.\n"
                             "To explicitly read as AminoAcid, add '--query-alphabet aminoacid'.\n"
                             "To ignore and disable this warning, add '--query-alphabet dna5'.\n");
        return LGPU_ALPH_DNA5;
    }
    if (allIn("ABCDEFGHIJKLMNOPQRSTUVWXYZ*abcdefghijklmnopqrstuvwxyz"))
        return LGPU_ALPH_AMINO_ACID;
    die("Your query file contains illegal characters in the first sequence.");
}

// FASTA / FASTQ, plain or gzip-compressed (query_reader.hpp); residues become BioC++ ranks
Fasta readQueries(std::string const & path, bool aminoAcid)
{
    // char -> rank like BioC++ (aa27: unknown -> X; dna5: unknown -> N, U -> T)
    uint8_t tab[256];
    if (aminoAcid)
    {
        std::memset(tab, 23, sizeof(tab));
        char const * alph = "ABCDEFGHIJKLMNOPQRSTUVWXYZ*";
        for (int r = 0; alph[r]; ++r)
        {
            tab[static_cast<uint8_t>(alph[r])] = static_cast<uint8_t>(r);
            if (alph[r] >= 'A' && alph[r] <= 'Z')
                tab[static_cast<uint8_t>(alph[r] - 'A' + 'a')] = static_cast<uint8_t>(r);
        }
    }
    else
    {
        std::memset(tab, 3, sizeof(tab));
        char const * alph = "ACGNT";
        for (int r = 0; alph[r]; ++r)
        {
            tab[static_cast<uint8_t>(alph[r])]             = static_cast<uint8_t>(r);
            tab[static_cast<uint8_t>(alph[r] - 'A' + 'a')] = static_cast<uint8_t>(r);
        }
        tab[static_cast<uint8_t>('U')] = tab[static_cast<uint8_t>('u')] = 4;
    }
    Fasta          f;
    qio::SeqReader reader(path);
    qio::Record    rec;
    while (reader.next(rec))
    {
        f.ids.push_back(rec.id);
        for (char c : rec.seq)
            f.residues.push_back(tab[static_cast<uint8_t>(c)]);
        f.offsets.push_back(f.residues.size());
    }
    return f;
}

double now()
{
    return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

struct ShardResult
{
    std::vector<lgpu_hit> hits; // q_id already global
    std::vector<uint32_t> cigar; // run-length ops of the hits (SAM output)
    lgpu_stats            stats{};
    std::string           error;
    double                tUpload = 0, tSearch = 0, tTeardown = 0; // seconds
};

// `first`: the index on GPU 0 (uploaded from the host while the queries were read); the other GPUs copy it device to device
void runShard(Options const & o, lgpu_index * first, int device, Fasta const & f, uint64_t qBegin, uint64_t qEnd, ShardResult & out)
{
    double const t0 = now();
    lgpu_index * ix = first;
    if (device != 0 && lgpu_index_clone(&ix, first, device) != LGPU_OK)
    {
        out.error = lgpu_last_error(nullptr);
        return;
    }
    out.tUpload = now() - t0;
    double const t1 = now();
    lgpu_ctx * ctx = nullptr;
    if (lgpu_ctx_create(&ctx, ix, &o.params) != LGPU_OK)
    {
        out.error = lgpu_last_error(nullptr);
        lgpu_index_destroy(ix);
        return;
    }
    std::vector<uint64_t> offs;
    for (uint64_t b = qBegin; b < qEnd; b += o.blockSize)
    {
        uint64_t const e = std::min(qEnd, b + o.blockSize);
        offs.assign(f.offsets.begin() + b, f.offsets.begin() + e + 1);
        uint64_t const base = offs[0];
        for (auto & x : offs)
            x -= base;
        lgpu_query_batch qb{f.residues.data() + base, offs.data(), e - b, 0};
        lgpu_hits        hits{};
        if (lgpu_search_batch(ctx, &qb, &hits, &out.stats) != LGPU_OK)
        {
            out.error = lgpu_last_error(ctx);
            break;
        }
        size_t const   old       = out.hits.size();
        uint32_t const cigarBase = static_cast<uint32_t>(out.cigar.size());
        out.hits.insert(out.hits.end(), hits.hits, hits.hits + hits.n);
        if (hits.cigar_ops)
            out.cigar.insert(out.cigar.end(), hits.cigar_ops, hits.cigar_ops + hits.n_cigar_ops);
        for (size_t i = old; i < out.hits.size(); ++i)
        {
            out.hits[i].q_id += static_cast<uint32_t>(b);
            out.hits[i].cigar_off += cigarBase;
        }
    }
    out.tSearch = now() - t1;
    double const t2 = now();
    lgpu_ctx_destroy(ctx);
    if (device != 0)
        lgpu_index_destroy(ix);
    out.tTeardown = now() - t2;
}

} // namespace

// BGZF (SAM/BAM specification 4.1): independent gzip members of <= 64 KiB with the block size in an extra field
void bgzfBlock(FILE * fo, unsigned char const * p, size_t n)
{
    unsigned char out[65536 + 64];
    z_stream      zs{};
    if (deflateInit2(&zs, Z_DEFAULT_COMPRESSION, Z_DEFLATED, -15, 8, Z_DEFAULT_STRATEGY) != Z_OK)
        die("zlib: deflateInit2 failed");
    zs.next_in   = const_cast<unsigned char *>(p);
    zs.avail_in  = static_cast<uInt>(n);
    zs.next_out  = out + 18;
    zs.avail_out = sizeof(out) - 18 - 8;
    if (deflate(&zs, Z_FINISH) != Z_STREAM_END)
        die("zlib: deflate failed");
    size_t const clen = zs.total_out;
    deflateEnd(&zs);
    unsigned char const hdr[18] = {0x1f, 0x8b, 8, 4, 0, 0, 0, 0, 0, 0xff, 6, 0, 'B', 'C', 2, 0, 0, 0};
    std::memcpy(out, hdr, 18);
    size_t const total = 18 + clen + 8;
    out[16] = static_cast<unsigned char>((total - 1) & 0xff);
    out[17] = static_cast<unsigned char>((total - 1) >> 8);
    uint32_t const crc = static_cast<uint32_t>(crc32(crc32(0L, Z_NULL, 0), p, static_cast<uInt>(n)));
    uint32_t const isz = static_cast<uint32_t>(n);
    std::memcpy(out + 18 + clen, &crc, 4);
    std::memcpy(out + 18 + clen + 4, &isz, 4);
    std::fwrite(out, 1, total, fo);
}

// Writes the full 0xff00-byte blocks of `raw` (everything, if `final`) and keeps the rest: the BAM stream goes out
// block by block instead of being held in memory until the end.
void bgzfDrain(FILE * fo, std::string & raw, bool final)
{
    size_t off = 0;
    while (raw.size() - off >= 0xff00 || (final && off < raw.size()))
    {
        size_t const n = std::min<size_t>(0xff00, raw.size() - off);
        bgzfBlock(fo, reinterpret_cast<unsigned char const *>(raw.data()) + off, n);
        off += n;
    }
    raw.erase(0, off);
    if (final)
        bgzfBlock(fo, nullptr, 0); // end-of-file marker block
}

// stdio stream that compresses into a .gz file while it is written (no copy of the output in memory)
static ssize_t gzCookieWrite(void * c, char const * buf, size_t n)
{
    int const w = gzwrite(static_cast<gzFile>(c), buf, static_cast<unsigned>(std::min<size_t>(n, 1u << 30)));
    return w <= 0 ? 0 : w; // 0 = error for stdio
}
static int gzCookieClose(void * c) { return gzclose(static_cast<gzFile>(c)) == Z_OK ? 0 : -1; }


// `lambda3_b200 dumpq -q FILE [-a auto|dna5|aminoacid]`: parse the query file exactly like a search would and print
// "alphabet\t<name>" followed by "<id>\t<residues>" per record (no GPU involved; used by the CPU tests)
static int dumpQueries(int argc, char ** argv)
{
    std::string path, alph = "auto";
    for (int i = 2; i + 1 < argc; i += 2)
    {
        if (!std::strcmp(argv[i], "-q") || !std::strcmp(argv[i], "--query"))
            path = argv[i + 1];
        else if (!std::strcmp(argv[i], "-a") || !std::strcmp(argv[i], "--input-alphabet"))
            alph = argv[i + 1];
        else
            die(std::string("unknown option '") + argv[i] + "'");
    }
    if (path.empty())
        die("-q is required");
    uint32_t const a = alph == "dna5" ? static_cast<uint32_t>(LGPU_ALPH_DNA5)
                       : alph == "aminoacid" ? static_cast<uint32_t>(LGPU_ALPH_AMINO_ACID) : detectAlphabet(path);
    bool const  aa = a == LGPU_ALPH_AMINO_ACID;
    Fasta const f  = readQueries(path, aa);
    char const * letters = aa ? "ABCDEFGHIJKLMNOPQRSTUVWXYZ*" : "ACGNT";
    std::printf("alphabet\t%s\n", aa ? "aminoacid" : "dna5");
    for (size_t i = 0; i < f.ids.size(); ++i)
    {
        std::string seq;
        for (uint64_t k = f.offsets[i]; k < f.offsets[i + 1]; ++k)
            seq.push_back(letters[f.residues[k]]);
        std::printf("%s\t%s\n", f.ids[i].c_str(), seq.c_str());
    }
    return 0;
}

static int run(int argc, char ** argv);

int main(int argc, char ** argv)
{
    try
    {
        if (argc >= 2 && !std::strcmp(argv[1], "dumpq"))
            return dumpQueries(argc, argv);
        return run(argc, argv);
    }
    catch (std::exception const & e) // query parser errors (query_reader.hpp); same exit path as the reference's searchMain
    {
        die(e.what());
    }
}

static int run(int argc, char ** argv)
{
    Options o;
    parse(argc, argv, o);
    double const t0 = now();
    if (o.verbosity >= 1)
        std::printf("LAMBDA (B200 engine) - the Local Aligner for Massive Biological DatA\n\n");

    // Driver initialisation costs time per VISIBLE device (seconds on an 8-GPU box without persistence mode): unless the
    // user chose the devices, show the process only the GPUs it is going to use.
    if (!std::getenv("CUDA_VISIBLE_DEVICES") && o.replayHits.empty())
    {
        std::string vis;
        for (int g = 0; g < o.gpus; ++g)
            vis += (g ? "," : "") + std::to_string(g);
        setenv("CUDA_VISIBLE_DEVICES", vis.c_str(), 1);
    }
    lgpu_lba * lba = nullptr;
    if (lgpu_lba_open(&lba, o.index.c_str()) != LGPU_OK)
        die(lgpu_last_error(nullptr));
    lgpu_index_desc const * desc = lgpu_lba_desc(lba);
    lgpu_taxonomy const *   tax  = lgpu_lba_taxonomy(lba);
    // src/search_algo.hpp:303-314
    if (o.hasSTaxIds && (!tax->s_tax_delims || tax->n_s_tax_ids == 0))
        die("You requested printing of taxonomic IDs and/or taxonomic binning, but the index does not contain taxonomic "
            "information. Recreate it and provide --acc-tax-map .");
    if (o.computeLCA && (tax->n_taxa == 0 || !tax->taxon_name_delims))
        die("You requested taxonomic binning, but the index does not contain a taxonomic tree. Recreate it and provide "
            "--tax-dump-dir .");
    double const            t1   = now();
    // CUDA context creation and the index upload to GPU 0 run while this thread reads the queries
    lgpu_index *             ix0 = nullptr;
    std::string              ix0Error;
    double                   tUpload0 = 0;
    std::vector<std::thread> early;
    if (o.replayHits.empty())
    {
        for (int g = 1; g < o.gpus; ++g)
            early.emplace_back([g] { lgpu_device_warmup(g); });
        early.emplace_back([&] {
            double const tu = now();
            if (lgpu_index_create(&ix0, desc, 0) != LGPU_OK)
                ix0Error = lgpu_last_error(nullptr);
            tUpload0 = now() - tu;
        });
    }
    struct EarlyJoin
    {
        std::vector<std::thread> & t;
        ~EarlyJoin()
        {
            for (auto & x : t)
                if (x.joinable())
                    x.join();
        }
    } earlyJoin{early};
    // query alphabet: fixed for searchn / searchbs, given or auto-detected for searchp (src/search.cpp:209-216)
    uint32_t qryAlph = LGPU_ALPH_DNA5;
    if (o.domain == LGPU_DOMAIN_PROTEIN)
        qryAlph = o.inputAlphabet == "dna5"        ? static_cast<uint32_t>(LGPU_ALPH_DNA5)
                  : o.inputAlphabet == "aminoacid" ? static_cast<uint32_t>(LGPU_ALPH_AMINO_ACID)
                                                   : detectAlphabet(o.query);
    o.params.query_alph = qryAlph;
    Fasta const             f    = readQueries(o.query, qryAlph == LGPU_ALPH_AMINO_ACID);
    uint64_t const          nQ   = f.ids.size();
    double const            t2   = now();
    if (o.verbosity >= 2)
        std::printf("Index mapped in %.3fs (%llu subjects), %llu queries read in %.3fs\n", t1 - t0,
                    static_cast<unsigned long long>(desc->n_seqs), static_cast<unsigned long long>(nQ), t2 - t1);

    int const                nShards = o.replayHits.empty()
                                         ? static_cast<int>(std::min<uint64_t>(static_cast<uint64_t>(o.gpus), std::max<uint64_t>(nQ, 1)))
                                         : 1;
    std::vector<ShardResult> res(nShards);
    if (!o.replayHits.empty())
    {
        // Test hook: NO search happens here.  The file holds finished records (what lgpu_search_batch returned, or
        // what the test infrastructure computed) and this program only formats them, so that the five output
        // writers can be checked against the reference's files on a machine without a GPU.
        // Layout: "LGPUHITS", u64 n_hits, u64 n_ops, lgpu_stats, n_hits x lgpu_hit, n_ops x u32.
        FILE * fi = std::fopen(o.replayHits.c_str(), "rb");
        if (!fi)
            die("cannot open " + o.replayHits);
        char     magic[8];
        uint64_t nHits = 0, nOps = 0;
        bool     ok    = std::fread(magic, 1, 8, fi) == 8 && !std::memcmp(magic, "LGPUHITS", 8) && std::fread(&nHits, 8, 1, fi) == 1 &&
                  std::fread(&nOps, 8, 1, fi) == 1 && std::fread(&res[0].stats, sizeof(lgpu_stats), 1, fi) == 1;
        if (ok)
        {
            res[0].hits.resize(nHits);
            res[0].cigar.resize(nOps);
            ok = (nHits == 0 || std::fread(res[0].hits.data(), sizeof(lgpu_hit), nHits, fi) == nHits) &&
                 (nOps == 0 || std::fread(res[0].cigar.data(), 4, nOps, fi) == nOps);
        }
        std::fclose(fi);
        if (!ok)
            die("malformed hit file " + o.replayHits);
        for (auto const & h : res[0].hits)
            if (h.q_id >= nQ || h.s_id >= desc->n_seqs || static_cast<uint64_t>(h.cigar_off) + h.cigar_len > nOps)
                die("hit file " + o.replayHits + " does not belong to these queries / this index");
    }
    for (auto & t : early)
        t.join();
    early.clear();
    if (!ix0Error.empty())
        die(ix0Error);
    std::vector<std::thread> th;
    if (o.replayHits.empty())
        for (int g = 0; g < nShards; ++g)
            th.emplace_back(runShard, std::cref(o), ix0, g, std::cref(f), nQ * g / nShards, nQ * (g + 1) / nShards, std::ref(res[g]));
    for (auto & t : th)
        t.join();
    if (!res.empty() && o.replayHits.empty())
        res[0].tUpload += tUpload0;
    lgpu_stats total{};
    for (auto & r : res)
    {
        if (!r.error.empty())
            die(r.error);
        uint64_t const * s = reinterpret_cast<uint64_t const *>(&r.stats);
        uint64_t *       d = reinterpret_cast<uint64_t *>(&total);
        for (int k = 0; k < 16; ++k)
            d[k] += s[k];
        total.ms_total += r.stats.ms_total;
    }
    double const t3 = now();

    // --- write records in the reference's -t 1 order ---
    std::vector<std::vector<lgpu_hit const *>> perQuery(nQ);
    for (auto const & r : res)
        for (auto const & h : r.hits)
            perQuery[h.q_id].push_back(&h);
    // a hit's run-length ops live in the cigar vector of the shard that produced it
    auto cigarOf = [&](lgpu_hit const * h) -> uint32_t const * {
        for (auto const & r : res)
            if (!r.hits.empty() && h >= r.hits.data() && h < r.hits.data() + r.hits.size())
                return r.cigar.data() + h->cigar_off;
        return nullptr;
    };
    uint64_t const rpb = std::max<uint64_t>(std::min<uint64_t>(nQ / (static_cast<uint64_t>(o.threads) * 10), 10), 1);
    // .gz outputs are compressed while they are written
    FILE * fo = nullptr;
    if (o.gzOutput)
    {
        gzFile gz = gzopen(o.output.c_str(), "wb");
        if (gz)
        {
            cookie_io_functions_t io{};
            io.write = gzCookieWrite;
            io.close = gzCookieClose;
            fo       = fopencookie(gz, "w", io);
            if (!fo)
                gzclose(gz);
        }
    }
    else
        fo = std::fopen(o.output.c_str(), "wb");
    if (!fo)
        die("cannot create output file " + o.output);
    // a failed write (disk full, I/O error) must not leave a truncated file behind a zero exit code
    auto checkOutput = [&](bool closing) {
        bool bad = std::ferror(fo) != 0;
        if (closing)
            bad = (std::fclose(fo) != 0) || bad;
        if (bad)
        {
            if (!closing)
                std::fclose(fo);
            std::remove(o.output.c_str());
            die("error while writing " + o.output + " (disk full?); the partial file was removed");
        }
    };
    std::vector<char> line(1 << 16);
    auto              subjectId = [&](uint32_t s) {
        return std::string(desc->ids + desc->id_delims[s], desc->ids + desc->id_delims[s + 1]);
    };
    // the reference chunks the queries per thread first (src/search.cpp:384-385), then into batches
    // .m9: comment lines in front of every record = query with at least one match
    // (SQ/blast/blast_tabular_out.h:149-205, version string src/search_output.hpp:309-317)
    bool const  qTrans  = qryAlph == LGPU_ALPH_DNA5 && desc->trans_alph == LGPU_ALPH_AMINO_ACID;
    bool const  sTrans  = desc->orig_alph == LGPU_ALPH_DNA5 && desc->trans_alph == LGPU_ALPH_AMINO_ACID;
    char const * program = desc->trans_alph != LGPU_ALPH_AMINO_ACID ? "BLASTN"
                           : qTrans ? (sTrans ? "TBLASTX" : "BLASTX")
                                    : (sTrans ? "TBLASTN" : "BLASTP");
    std::string const versionLine =
      o.versionToOutput ? std::string(program) + " 2.2.26+ [created by LAMBDA-3.0.0, see http://seqan.de/lambda and please "
                                                 "cite correctly in your academic work]"
                        : std::string(program) + " 2.2.26+ [I/O Module of SeqAn-2.4.1, http://www.seqan.de]";
    // ---- taxonomy of a record (_writeRecord, src/search_algo.hpp:884-909; computeLCA, src/search_misc.hpp:86-112) ----
    uint32_t    recLcaTaxId = 0; // of the record that is being written
    std::string recLcaId;
    auto        taxIdsOf = [&](uint32_t sId) {
        return std::pair<uint32_t const *, uint32_t const *>(tax->s_tax_ids + tax->s_tax_delims[sId],
                                                             tax->s_tax_ids + tax->s_tax_delims[sId + 1]);
    };
    auto lca2 = [&](uint32_t n1, uint32_t n2) -> uint32_t {
        if (n1 == n2)
            return n1;
        for (unsigned i = tax->taxon_heights[n1]; i > tax->taxon_heights[n2]; --i) // bring both to the same height
            n1 = tax->taxon_parents[n1];
        for (unsigned i = tax->taxon_heights[n2]; i > tax->taxon_heights[n1]; --i)
            n2 = tax->taxon_parents[n2];
        while (n1 != 0 && n2 != 0)
        {
            if (n1 == n2)
                return n1;
            n1 = tax->taxon_parents[n1];
            n2 = tax->taxon_parents[n2];
        }
        die("LCA-computation error: One of the paths didn't lead to root.");
    };
    auto computeRecordLca = [&](uint64_t q, int phase) {
        recLcaTaxId = 0;
        recLcaId.clear();
        if (!o.computeLCA)
            return;
        for (lgpu_hit const * h : perQuery[q])
            if (h->phase == phase)
            {
                auto const [b, e] = taxIdsOf(h->s_id);
                if (b != e && tax->taxon_parents[*b] != 0)
                {
                    recLcaTaxId = *b;
                    break;
                }
            }
        if (recLcaTaxId != 0)
            for (lgpu_hit const * h : perQuery[q])
                if (h->phase == phase)
                {
                    auto const [b, e] = taxIdsOf(h->s_id);
                    for (uint32_t const * t = b; t != e; ++t)
                        if (tax->taxon_parents[*t] != 0) // unassigned subjects are ignored
                            recLcaTaxId = lca2(*t, recLcaTaxId);
                }
        recLcaId.assign(tax->taxon_names + tax->taxon_name_delims[recLcaTaxId],
                        tax->taxon_names + tax->taxon_name_delims[recLcaTaxId + 1]);
    };
    // one tabular line: the library formats the ordinary columns, the taxonomy columns are filled in here
    // (SQ/blast/blast_tabular_out.h:440-492)
    // Lines without taxonomy columns depend on nothing but their record: they are formatted up front by several threads
    // (the formatter is as long as the search on the benchmark workload), the writer below only copies them out.
    bool plainColumns = !o.sam && !o.report;
    for (uint32_t c : o.columns)
        plainColumns = plainColumns && lgpu_tabular_column_supported(c);
    std::vector<std::vector<std::string>> preLines(res.size());
    if (plainColumns)
    {
        unsigned int const nFmt = std::max(1u, std::min(8u, std::thread::hardware_concurrency()));
        for (size_t g = 0; g < res.size(); ++g)
        {
            preLines[g].resize(res[g].hits.size());
            std::vector<std::thread> ft;
            std::atomic<bool>        failed{false};
            for (unsigned int t = 0; t < nFmt; ++t)
                ft.emplace_back([&, g, t] {
                    std::vector<char> buf(1 << 16);
                    size_t const      n = res[g].hits.size(), b = n * t / nFmt, e = n * (t + 1) / nFmt;
                    for (size_t i = b; i < e; ++i)
                    {
                        lgpu_hit const &  h   = res[g].hits[i];
                        std::string const sId = subjectId(h.s_id);
                        int const len = lgpu_format_tabular(&o.params, &h, f.ids[h.q_id].c_str(), sId.c_str(), o.columns.data(),
                                                            o.columns.size(), buf.data(), buf.size());
                        if (len <= 0)
                        {
                            failed = true;
                            return;
                        }
                        preLines[g][i].assign(buf.data(), static_cast<size_t>(len));
                    }
                });
            for (auto & t : ft)
                t.join();
            if (failed)
                die("cannot format a tabular line");
        }
    }
    std::string tabLine;
    auto        tabularLine = [&](uint64_t q, lgpu_hit const * h) -> std::string const & {
        if (plainColumns)
            for (size_t g = 0; g < res.size(); ++g)
                if (!res[g].hits.empty() && h >= res[g].hits.data() && h < res[g].hits.data() + res[g].hits.size())
                    return preLines[g][static_cast<size_t>(h - res[g].hits.data())];
        tabLine.clear();
        std::string const sId = subjectId(h->s_id);
        auto isTax = [](uint32_t c) { return !lgpu_tabular_column_supported(c); };
        for (size_t i = 0; i < o.columns.size();)
        {
            if (i)
                tabLine += '\t';
            if (!isTax(o.columns[i]))
            {
                size_t j = i;
                while (j < o.columns.size() && !isTax(o.columns[j]))
                    ++j;
                int const n = lgpu_format_tabular(&o.params, h, f.ids[q].c_str(), sId.c_str(), o.columns.data() + i, j - i,
                                                  line.data(), line.size());
                if (n <= 0)
                    die("cannot format a tabular line");
                tabLine.append(line.data(), static_cast<size_t>(n) - 1); // without the newline
                i = j;
                continue;
            }
            char const * name = lgpu_tabular_column_name(o.columns[i]);
            if (!std::strcmp(name, "staxids"))
            {
                auto const            be = taxIdsOf(h->s_id);
                std::vector<uint32_t> ids(be.first, be.second);
                std::sort(ids.begin(), ids.end()); // "they have to be sorted numerically"
                if (ids.empty())
                    tabLine += "n/a";
                for (size_t k = 0; k < ids.size(); ++k)
                    tabLine += (k ? ";" : "") + std::to_string(ids[k]);
            }
            else if (!std::strcmp(name, "lcaid"))
            {
                if (recLcaId.empty())
                    tabLine += "n/a";
                for (char c : recLcaId)
                    tabLine += (c == ' ' || c == '\t') ? '_' : c;
            }
            else
                tabLine += std::to_string(recLcaTaxId);
            ++i;
        }
        tabLine += '\n';
        return tabLine;
    };
    uint64_t nRecords = 0;
    auto     recordHeader = [&](uint64_t q, size_t nHits) {
        ++nRecords;
        if (!o.comments)
            return;
        std::fprintf(fo, "# %s\n# Query: %s\n# Database: %s\n", versionLine.c_str(), f.ids[q].c_str(), o.index.c_str());
        if (nHits)
        {
            // _writeFieldLabels (SQ/blast/blast_tabular_out.h): the columns' labels joined by ", "
            std::fputs("# Fields: ", fo);
            for (size_t c = 0; c < o.columns.size(); ++c)
                std::fprintf(fo, "%s%s", c ? ", " : "", lgpu_tabular_column_label(o.columns[c]));
            std::fputc('\n', fo);
        }
        std::fprintf(fo, "# %zu hits found\n", nHits);
    };
    // residues of the frames of a stored sequence (used by the SAM tags and the pairwise report)
    auto       frameLen  = [](uint64_t len, unsigned f) { uint64_t const o = f % 3; return (std::max(len, o) - o) / 3; };
    // residue k (as a character) of frame `frame` of a stored sequence: strands for nucleotide searches,
    // six-frame translation (canonical code) for translated ones, the sequence itself otherwise
    auto frameChar = [&](uint8_t const * seq, uint64_t len, bool translated, bool nucleotide, int frame, uint64_t k) -> char {
        if (translated)
        {
            uint64_t const pos = 3 * k + static_cast<uint64_t>(std::abs(frame)) - 1;
            unsigned n1, n2, n3;
            if (frame > 0)
            {
                n1 = seq[pos]; n2 = seq[pos + 1]; n3 = seq[pos + 2];
            }
            else
            {
                n1 = kDna5Complement[seq[len - pos - 1]]; n2 = kDna5Complement[seq[len - pos - 2]];
                n3 = kDna5Complement[seq[len - pos - 3]];
            }
            return kAa27RankToChar[kDna5Translate[(n1 * 5 + n2) * 5 + n3]];
        }
        if (nucleotide)
            return kDna5RankToChar[frame < 0 ? kDna5Complement[seq[len - 1 - k]] : seq[k]];
        return kAa27RankToChar[seq[k]];
    };
    // ---- SAM (src/search_output.hpp:346-458 header, :482-716 records; defaults of src/search_options.hpp:339-370) ----
    bool const isBlastN = desc->trans_alph != LGPU_ALPH_AMINO_ACID;
    std::string bamRaw; // uncompressed BAM stream
    auto        put32 = [&](uint32_t v) { bamRaw.append(reinterpret_cast<char const *>(&v), 4); };
    auto        refName = [&](uint64_t sq) {
        std::string name = subjectId(static_cast<uint32_t>(sq));
        return name.substr(0, name.find(' '));
    };
    if (o.sam)
    {
        std::string text = "@HD\tVN:1.4\tGO:query\n";
        if (o.versionToOutput)
            text += "@PG\tID:lambda\tPN:lambda\tVN:3.0.0\tCL:" + o.commandLine + "\n";
        text += "@CO\tLambda is a high performance BLAST compatible local aligner, please see http://seqan.de/lambda for "
                "more information.\n"
                "@CO\tSAM/BAM dialect documentation is available here: https://github.com/seqan/lambda/wiki/Output-Formats\n"
                "@CO\tIf you use any results found by Lambda, please cite Hauswedell et al. (2014) doi: "
                "10.1093/bioinformatics/btu439\n"
                "@CO\tOptional tags as follow";
        for (int t = 0; t < 14; ++t)
            if (o.samTags[t])
                text += std::string("\t") + kSamTagKeys[t] + ":" + kSamTagDescr[t];
        text += "\n";
        // .sam with --sam-with-refheader: the default writeHeader() appends the @SQ lines after the header records
        if (!o.bam && o.samWithRefHeader)
            for (uint64_t sq = 0; sq < desc->n_seqs; ++sq)
                text += "@SQ\tSN:" + refName(sq) + "\tLN:" + std::to_string(desc->seq_delims[sq + 1] - desc->seq_delims[sq]) + "\n";
        if (o.bam)
        {
            bamRaw = "BAM\1";
            put32(static_cast<uint32_t>(text.size()));
            bamRaw += text;
            put32(static_cast<uint32_t>(desc->n_seqs));
            for (uint64_t sq = 0; sq < desc->n_seqs; ++sq)
            {
                std::string const name = refName(sq);
                put32(static_cast<uint32_t>(name.size() + 1));
                bamRaw.append(name.c_str(), name.size() + 1);
                put32(static_cast<uint32_t>(desc->seq_delims[sq + 1] - desc->seq_delims[sq]));
            }
        }
        else
            std::fputs(text.c_str(), fo);
    }
    using CigarVec = std::vector<std::pair<char, unsigned>>;
    auto cigarText = [](CigarVec const & el) {
        std::string t;
        for (auto const & e : el)
            t += std::to_string(e.second) + e.first;
        return t;
    };
    std::string seqStr;
    auto        samRecord = [&](uint64_t q, lgpu_hit const * h, lgpu_hit const * prev, size_t nHits) {
        uint64_t const qLen = f.offsets[q + 1] - f.offsets[q]; // record.qLength: original query length
        // POS (src/search_output.hpp:499-510; the qLength in the reverse-frame branch is the reference's)
        int32_t beginPos = static_cast<int32_t>(h->s_start);
        if (sTrans)
        {
            beginPos = static_cast<int32_t>(h->s_start * 3 + static_cast<uint32_t>(std::abs(h->s_frame)) - 1);
            if (h->s_frame < 0)
                beginPos = static_cast<int32_t>(qLen) - beginPos;
        }
        unsigned flag = prev ? 256u : 0u;
        if (h->q_frame < 0)
            flag |= 16u;
        // CIGARs (blastMatchOneCigar :116-196, blastMatchTwoCigar :199-298).  Clips caused by the frame are always
        // hard clips, those of the local alignment hard or soft (--sam-bam-clip).  The arithmetic is the
        // reference's, unsigned wrap-around for untranslated protein queries (frame 0) included.
        unsigned const leftFrameClip = static_cast<unsigned>(std::abs(h->q_frame)) - 1u;
        uint64_t const srcLen        = qTrans ? (std::max<uint64_t>(qLen, leftFrameClip) - leftFrameClip) / 3 : qLen; // frame length
        auto           oneCigar      = [&](unsigned transFac, bool translated) {
            unsigned const rightFrameClip = translated ? static_cast<unsigned>((static_cast<unsigned>(qLen) - leftFrameClip) % 3) : 0u;
            unsigned const leftClip       = h->q_start * transFac;
            unsigned const rightClip      = static_cast<unsigned>(srcLen - h->q_end) * transFac;
            CigarVec       el;
            if (o.samHardClip)
            {
                if (leftFrameClip + leftClip > 0)
                    el.push_back({'H', leftFrameClip + leftClip});
            }
            else
            {
                if (leftFrameClip > 0)
                    el.push_back({'H', leftFrameClip});
                if (leftClip > 0)
                    el.push_back({'S', leftClip});
            }
            uint32_t const * ops = cigarOf(h);
            for (uint32_t k = h->cigar_len; k-- > 0;) // stored END first
            {
                uint32_t const kind = ops[k] & 3u, run = ops[k] >> 2;
                el.push_back({kind == LGPU_CIGAR_M ? 'M' : kind == LGPU_CIGAR_I ? 'I' : 'D', run * transFac});
            }
            if (o.samHardClip)
            {
                if (rightFrameClip + rightClip > 0)
                    el.push_back({'H', rightFrameClip + rightClip});
            }
            else
            {
                if (rightClip > 0)
                    el.push_back({'S', rightClip});
                if (rightFrameClip > 0)
                    el.push_back({'H', rightFrameClip});
            }
            return el;
        };
        CigarVec cigarEl, protEl; // CIGAR column; OC tag (protein space)
        if (o.samTags[ST_OC])
        {
            if (isBlastN)
                cigarEl = oneCigar(1, false);
            else if (qTrans)
            {
                cigarEl = oneCigar(3, true);
                // the protein CIGAR of blastMatchTwoCigar: local clips only, never reversed
                unsigned const leftClip = h->q_start, rightClip = static_cast<unsigned>(srcLen - h->q_end);
                if (leftClip > 0)
                    protEl.push_back({o.samHardClip ? 'H' : 'S', leftClip});
                uint32_t const * ops = cigarOf(h);
                for (uint32_t k = h->cigar_len; k-- > 0;)
                {
                    uint32_t const kind = ops[k] & 3u, run = ops[k] >> 2;
                    protEl.push_back({kind == LGPU_CIGAR_M ? 'M' : kind == LGPU_CIGAR_I ? 'I' : 'D', run});
                }
                if (rightClip > 0)
                    protEl.push_back({o.samHardClip ? 'H' : 'S', rightClip});
            }
            else
            {
                protEl = oneCigar(1, false); // BLASTP / TBLASTN: the one CIGAR is the protein CIGAR
                if (h->q_frame < 0)
                    std::reverse(protEl.begin(), protEl.end());
            }
        }
        else if (isBlastN || qTrans)
            cigarEl = oneCigar(qTrans ? 3 : 1, qTrans);
        if (h->q_frame < 0)
            std::reverse(cigarEl.begin(), cigarEl.end());
        std::string const cigarStr = cigarEl.empty() ? "*" : cigarText(cigarEl);
        // SEQ (--sam-bam-seq; uniq: only when frame or aligned query range differ from the previous match)
        bool const writeSeq = o.samBamSeq > 1 ||
                              (o.samBamSeq == 1 && (!prev || prev->q_frame != h->q_frame || prev->q_start != h->q_start ||
                                                    prev->q_end != h->q_end));
        seqStr.clear();
        if (writeSeq && (isBlastN || qTrans))
        {
            static char const dna5[] = "ACGNT";
            uint8_t const *   src    = f.residues.data() + f.offsets[q];
            // hard clipping: the aligned part; soft clipping: everything the frame covers
            uint64_t b = o.samHardClip ? h->q_start : 0, e = o.samHardClip ? h->q_end : srcLen; // in the frame's sequence
            if (qTrans)
            {
                uint64_t const shift = static_cast<uint64_t>(std::abs(h->q_frame)) - 1;
                b                    = 3 * b + shift;
                e                    = 3 * e + shift;
            }
            static uint8_t const comp[5] = {4, 2, 1, 3, 0}; // dna5 ranks A C G N T
            if (h->q_frame >= 0)
                for (uint64_t i = b; i < e; ++i)
                    seqStr += dna5[src[i]];
            else // the frame is the reverse complement of the original query (_untranslateSequence, :84-109)
                for (uint64_t i = b; i < e; ++i)
                    seqStr += dna5[comp[src[qLen - 1 - i]]];
        }
        if (seqStr.empty())
            seqStr = "*";
        std::string const qName = f.ids[q].substr(0, f.ids[q].find_first_of(" \t\v\f\r\n"));
        std::string const sName = refName(h->s_id);
        float const identity   = static_cast<float>(100.0 * static_cast<float>(h->n_match) / static_cast<float>(h->aln_len));
        float const similarity = static_cast<float>(100.0 * static_cast<float>(h->n_positive) / static_cast<float>(h->aln_len));
        // qs: the protein sequence the alignment was computed on (* for BLASTN or when SEQ is left out), :672-692
        std::string qsStr = "*";
        if (o.samTags[ST_qs] && !isBlastN && writeSeq)
        {
            uint64_t const b = o.samHardClip ? h->q_start : 0, e = o.samHardClip ? h->q_end : srcLen;
            qsStr.clear();
            for (uint64_t k = b; k < e; ++k)
                qsStr += frameChar(f.residues.data() + f.offsets[q], qLen, qTrans, false, h->q_frame, k);
        }
        std::string const ocStr = protEl.empty() ? "*" : cigarText(protEl);
        // st: the subject's tax ids in stored order, * if it has none (:641-662)
        std::string stStr;
        if (o.samTags[ST_st])
        {
            auto const be = taxIdsOf(h->s_id);
            for (uint32_t const * t = be.first; t != be.second; ++t)
                stStr += (t != be.first ? ";" : "") + std::to_string(*t);
            if (stStr.empty())
                stStr = "*";
        }
        // typed optional fields in the order the reference appends them (:597-719)
        float const    evF = static_cast<float>(h->evalue);
        uint16_t const as  = static_cast<uint16_t>(h->bit_score);
        uint8_t const  ar  = static_cast<uint8_t>(h->score);
        uint8_t const  ai  = static_cast<uint8_t>(identity);
        uint16_t const ap  = static_cast<uint16_t>(similarity);
        uint32_t const nm  = h->aln_len - h->n_match;
        uint32_t const ih  = static_cast<uint32_t>(nHits);
        if (o.bam)
        {
            // SAM/BAM specification 4.2; bin from the reference span of the CIGAR (one base without CIGAR)
            bool const     haveSeq = seqStr != "*";
            uint32_t const lSeq    = haveSeq ? static_cast<uint32_t>(seqStr.size()) : 0u;
            int64_t        refLen  = 0;
            for (auto const & e : cigarEl)
                if (e.first == 'M' || e.first == 'D')
                    refLen += e.second;
            int64_t const beg = beginPos, end = beginPos + std::max<int64_t>(refLen, 1) - 1;
            uint32_t      bin = 0;
            if (beg >> 14 == end >> 14) bin = static_cast<uint32_t>(((1 << 15) - 1) / 7 + (beg >> 14));
            else if (beg >> 17 == end >> 17) bin = static_cast<uint32_t>(((1 << 12) - 1) / 7 + (beg >> 17));
            else if (beg >> 20 == end >> 20) bin = static_cast<uint32_t>(((1 << 9) - 1) / 7 + (beg >> 20));
            else if (beg >> 23 == end >> 23) bin = static_cast<uint32_t>(((1 << 6) - 1) / 7 + (beg >> 23));
            else if (beg >> 26 == end >> 26) bin = static_cast<uint32_t>(((1 << 3) - 1) / 7 + (beg >> 26));
            std::string rec;
            auto        r32 = [&](uint32_t v) { rec.append(reinterpret_cast<char const *>(&v), 4); };
            r32(h->s_id);
            r32(static_cast<uint32_t>(beginPos));
            r32((bin << 16) | (255u << 8) | static_cast<uint32_t>(qName.size() + 1));
            r32((flag << 16) | static_cast<uint32_t>(cigarEl.size()));
            r32(lSeq);
            r32(0xffffffffu);
            r32(0xffffffffu);
            r32(0);
            rec.append(qName.c_str(), qName.size() + 1);
            for (auto const & e : cigarEl)
                r32((e.second << 4) | (e.first == 'M' ? 0u : e.first == 'I' ? 1u : e.first == 'D' ? 2u : e.first == 'S' ? 4u : 5u));
            auto code = [](char c) -> unsigned { return c == 'A' ? 1u : c == 'C' ? 2u : c == 'G' ? 4u : c == 'T' ? 8u : 15u; };
            for (uint32_t i = 0; i < lSeq; i += 2)
                rec += static_cast<char>((code(seqStr[i]) << 4) | (i + 1 < lSeq ? code(seqStr[i + 1]) : 0u));
            rec.append(lSeq, static_cast<char>(0xff));
            auto tagRaw = [&](int t, char type, void const * v, size_t n) {
                rec += kSamTagKeys[t];
                rec += type;
                rec.append(static_cast<char const *>(v), n);
            };
            if (o.samTags[ST_ae]) tagRaw(ST_ae, 'f', &evF, 4);
            if (o.samTags[ST_AS]) tagRaw(ST_AS, 'S', &as, 2);
            if (o.samTags[ST_ar]) tagRaw(ST_ar, 'C', &ar, 1);
            if (o.samTags[ST_ai]) tagRaw(ST_ai, 'C', &ai, 1);
            if (o.samTags[ST_ap]) tagRaw(ST_ap, 'S', &ap, 2);
            if (o.samTags[ST_qf]) tagRaw(ST_qf, 'c', &h->q_frame, 1);
            if (o.samTags[ST_sf]) tagRaw(ST_sf, 'c', &h->s_frame, 1);
            if (o.samTags[ST_st]) tagRaw(ST_st, 'Z', stStr.c_str(), stStr.size() + 1);
            if (o.samTags[ST_ls]) tagRaw(ST_ls, 'Z', recLcaId.c_str(), recLcaId.size() + 1);
            if (o.samTags[ST_lt]) tagRaw(ST_lt, 'I', &recLcaTaxId, 4);
            if (o.samTags[ST_qs]) tagRaw(ST_qs, 'Z', qsStr.c_str(), qsStr.size() + 1);
            if (o.samTags[ST_OC]) tagRaw(ST_OC, 'Z', ocStr.c_str(), ocStr.size() + 1);
            if (o.samTags[ST_NM]) tagRaw(ST_NM, 'I', &nm, 4);
            if (o.samTags[ST_IH]) tagRaw(ST_IH, 'I', &ih, 4);
            put32(static_cast<uint32_t>(rec.size()));
            bamRaw += rec;
            bgzfDrain(fo, bamRaw, false);
            return;
        }
        std::string line = qName + "\t" + std::to_string(flag) + "\t" + sName + "\t" + std::to_string(beginPos + 1) + "\t255\t" +
                           cigarStr + "\t*\t0\t0\t" + seqStr + "\t*";
        char ev[64];
        std::snprintf(ev, sizeof(ev), "%g", static_cast<double>(evF));
        auto tagInt = [&](int t, long long v) { line += std::string("\t") + kSamTagKeys[t] + ":i:" + std::to_string(v); };
        if (o.samTags[ST_ae]) line += std::string("\tae:f:") + ev;
        if (o.samTags[ST_AS]) tagInt(ST_AS, as);
        if (o.samTags[ST_ar]) tagInt(ST_ar, ar);
        if (o.samTags[ST_ai]) tagInt(ST_ai, ai);
        if (o.samTags[ST_ap]) tagInt(ST_ap, ap);
        if (o.samTags[ST_qf]) tagInt(ST_qf, h->q_frame);
        if (o.samTags[ST_sf]) tagInt(ST_sf, h->s_frame);
        if (o.samTags[ST_st]) line += "\tst:Z:" + stStr;
        if (o.samTags[ST_ls]) line += "\tls:Z:" + recLcaId;
        if (o.samTags[ST_lt]) tagInt(ST_lt, recLcaTaxId);
        if (o.samTags[ST_qs]) line += "\tqs:Z:" + qsStr;
        if (o.samTags[ST_OC]) line += "\tOC:Z:" + ocStr;
        if (o.samTags[ST_NM]) tagInt(ST_NM, nm);
        if (o.samTags[ST_IH]) tagInt(ST_IH, ih);
        line += "\n";
        std::fwrite(line.data(), 1, line.size(), fo);
    };
    // ---- .m0: BLAST pairwise report (SQ/blast/blast_report_out.h:296-868) ----
    bool const bsIndex   = desc->red_alph == LGPU_ALPH_DNA3BS;
    double  kaLambda = 0, kaK = 0, kaH = 0;
    int8_t  scoreMat[32 * 32];
    uint64_t dbLetters = desc->n_residues, dbSeqs = desc->n_seqs;
    if (o.report)
    {
        if (lgpu_ka_params(&o.params, &kaLambda, &kaK, &kaH) != LGPU_OK || lgpu_score_matrix(&o.params, scoreMat) != LGPU_OK)
            die("no statistics for this scoring scheme");
        if (sTrans)
        {
            dbSeqs *= 6;
            dbLetters = 0;
            for (uint64_t sq = 0; sq < desc->n_seqs; ++sq)
                for (unsigned fr = 0; fr < 6; ++fr)
                    dbLetters += frameLen(desc->seq_delims[sq + 1] - desc->seq_delims[sq], fr);
        }
        else if (bsIndex)
        {
            dbSeqs *= 2;
            dbLetters *= 2;
        }
        std::fprintf(fo, "%s\n\n\nReference: Altschul, Stephen F., Thomas L. Madden, Alejandro A. Schaffer,\nJinghui Zhang, Zheng "
                         "Zhang, Webb Miller, and David J. Lipman (1997),\n\"Gapped BLAST and PSI-BLAST: a new generation of "
                         "protein database search\nprograms\",  Nucleic Acids Res. 25:3389-3402.\n\n\n\nReference for SeqAn: "
                         "Doering, A., D. Weese, T. Rausch, K. Reinert (2008): SeqAn --\nAn efficient, generic C++ library for "
                         "sequence analysis. BMC Bioinformatics,\n9(1), 11. BioMed Central Ltd. doi:10.1186/1471-2105-9-11\n"
                         "\n\n\nDatabase: %s\n           %llu sequences; %llu total letters\n\n",
                     versionLine.c_str(), o.index.c_str(), static_cast<unsigned long long>(dbSeqs),
                     static_cast<unsigned long long>(dbLetters));
    }
    auto untranslatePos = [](uint64_t & b, uint64_t & e, int frame, uint64_t len, bool hasFrames, bool hasRevComp) {
        // _untranslateQPositions / _untranslateSPositions (SQ/blast/blast_base.h:337-420), as in the m8 writer
        if (!hasRevComp && !hasFrames)
        {
            ++b;
            return;
        }
        if (hasFrames)
        {
            uint64_t const shift = static_cast<uint64_t>(frame < 0 ? -frame : frame) - 1;
            b = b * 3 + shift;
            e = e * 3 + shift;
        }
        if (frame > 0)
            ++b;
        else
        {
            b = len - b;
            e = len - e + 1;
        }
    };
    std::string row0, row1;
    auto        reportRecord = [&](uint64_t q, std::vector<lgpu_hit const *> const & ms) {
        uint64_t const  qLen = f.offsets[q + 1] - f.offsets[q];
        uint8_t const * qSeq = f.residues.data() + f.offsets[q];
        std::fprintf(fo, "\nQuery= %s\n\nLength=%llu\n", f.ids[q].c_str(), static_cast<unsigned long long>(qLen));
        std::fputs("                                                                   Score     E\n"
                   "Sequences producing significant alignments:                       (Bits)  Value\n\n", fo);
        for (lgpu_hit const * h : ms)
        {
            std::string const sId = subjectId(h->s_id);
            if (sId.size() <= 66)
                std::fprintf(fo, "%s%*s", sId.c_str(), static_cast<int>(66 - sId.size()), "");
            else
                std::fprintf(fo, "%s...", sId.substr(0, 63).c_str());
            std::fprintf(fo, " %4li  %.1g\n", static_cast<long>(h->bit_score), h->evalue);
        }
        std::fputs("\nALIGNMENTS\n", fo);
        for (lgpu_hit const * h : ms)
        {
            std::string const sId = subjectId(h->s_id);
            std::fputs("> ", fo);
            for (size_t beg = 0, end = 0; end < sId.size();)
            {
                end += beg == 0 ? 64 : 60;
                end = std::min(end, sId.size());
                std::fprintf(fo, "%s\n", sId.substr(beg, end - beg).c_str());
                beg = end;
            }
            std::fprintf(fo, "Length=%u\n\n", h->s_len);
            float const identity   = static_cast<float>(100.0 * static_cast<float>(h->n_match) / static_cast<float>(h->aln_len));
            float const similarity = static_cast<float>(100.0 * static_cast<float>(h->n_positive) / static_cast<float>(h->aln_len));
            unsigned const nGaps   = h->n_gap_open + h->n_gap_ext;
            std::fprintf(fo, " Score =  %.1f bits (%u), Expect =  %.1g\n Identities = %u/%u (%d%%)", h->bit_score,
                         static_cast<unsigned>(h->score), h->evalue, h->n_match, h->aln_len,
                         static_cast<int>(std::lround(identity)));
            if (!isBlastN)
                std::fprintf(fo, ", Positives = %u/%u (%d%%)", h->n_positive, h->aln_len, static_cast<int>(std::lround(similarity)));
            std::fprintf(fo, ", Gaps = %u/%u (%d%%)", nGaps, h->aln_len,
                         static_cast<int>(std::lround(static_cast<double>(nGaps) * 100 / h->aln_len)));
            if (isBlastN)
                std::fprintf(fo, "\n Strand=%s/%s\n\n", h->q_frame == 1 ? "Plus" : "Minus", h->s_frame == 1 ? "Plus" : "Minus");
            else
            {
                if (qTrans || sTrans)
                    std::fputs("\n Frame = ", fo);
                if (qTrans && sTrans)
                    std::fprintf(fo, "%+d/%+d", h->q_frame, h->s_frame);
                else if (qTrans)
                    std::fprintf(fo, "%+d", h->q_frame);
                else if (sTrans)
                    std::fprintf(fo, "%+d", h->s_frame);
                std::fputs("\n\n", fo);
            }
            // the gapped rows
            uint64_t const  sOff = desc->seq_delims[h->s_id];
            uint64_t const  sLen = desc->seq_delims[h->s_id + 1] - sOff;
            uint8_t const * sSeq = desc->seqs + sOff;
            row0.clear();
            row1.clear();
            uint64_t         qi = h->q_start, si = h->s_start;
            uint32_t const * ops = cigarOf(h);
            for (uint32_t k = h->cigar_len; k-- > 0;)
            {
                uint32_t const kind = ops[k] & 3u, run = ops[k] >> 2;
                for (uint32_t r = 0; r < run; ++r)
                {
                    row0 += kind == LGPU_CIGAR_D ? '-' : frameChar(qSeq, qLen, qTrans, isBlastN, h->q_frame, qi++);
                    row1 += kind == LGPU_CIGAR_I ? '-' : frameChar(sSeq, sLen, sTrans, isBlastN, sTrans ? h->s_frame : 1, si++);
                }
            }
            uint64_t effQS = h->q_start, effQE = h->q_end, effSS = h->s_start, effSE = h->s_end;
            untranslatePos(effQS, effQE, h->q_frame, qLen, qTrans, isBlastN || qTrans);
            untranslatePos(effSS, effSE, h->s_frame, h->s_len, sTrans, sTrans);
            int const qStepOne = h->q_frame < 0 ? -1 : 1, sStepOne = h->s_frame < 0 ? -1 : 1;
            int const qStep = qTrans ? 3 * qStepOne : qStepOne, sStep = sTrans ? 3 * sStepOne : sStepOne;
            uint64_t const maxPos = std::max(std::max(effQS, effQE), std::max(effSS, effSE));
            int const      width  = maxPos == 0 ? 1 : static_cast<int>(std::floor(std::log10(static_cast<double>(maxPos))) + 1);
            long long qPos = 0, sPos = 0;
            for (uint32_t aPos = 0; aPos < h->aln_len;)
            {
                uint32_t const end = std::min<uint32_t>(aPos + 60, h->aln_len);
                std::fprintf(fo, "Query  %-*d  ", width, static_cast<int>(qPos + static_cast<long long>(effQS)));
                for (uint32_t i = aPos; i < end; ++i)
                    if (row0[i] != '-')
                        qPos += qStep;
                std::fwrite(row0.data() + aPos, 1, end - aPos, fo);
                std::fprintf(fo, "  %-*d", width, static_cast<int>(qPos + static_cast<long long>(effQS) - qStepOne));
                std::fprintf(fo, "\n         %*s", width, "");
                for (uint32_t i = aPos; i < end; ++i)
                {
                    char const a = row0[i], b = row1[i];
                    char       c = ' ';
                    if (isBlastN)
                        c = a == b ? '|' : ' ';
                    else if (a == b)
                        c = a;
                    else if (a != '-' && b != '-' && scoreMat[kAa27CharToRank[static_cast<uint8_t>(a)] * 32 +
                                                              kAa27CharToRank[static_cast<uint8_t>(b)]] > 0)
                        c = '+';
                    std::fputc(c, fo);
                }
                std::fprintf(fo, "\nSbjct  %-*d  ", width, static_cast<int>(sPos + static_cast<long long>(effSS)));
                for (uint32_t i = aPos; i < end; ++i)
                    if (row1[i] != '-')
                        sPos += sStep;
                std::fwrite(row1.data() + aPos, 1, end - aPos, fo);
                std::fprintf(fo, "  %-*d\n\n", width, static_cast<int>(sPos + static_cast<long long>(effSS) - sStepOne));
                aPos = end;
            }
            std::fputc('\n', fo);
        }
        std::fprintf(fo, "\nLambda     K      H\n   %-4.3f   %-5.4f   %-5.4f\n\nGapped\nLambda     K      H\n   %-4.3f   %-5.4f   "
                         "%-5.4f\n\nEffective search space used: %llu\n\n",
                     kaLambda, kaK, kaH, kaLambda, kaK, kaH, static_cast<unsigned long long>(qLen * dbLetters));
    };
    // the reference chunks the queries per thread first (src/search.cpp:384-385), then into batches
    for (int t = 0; t < o.threads; ++t)
    {
        uint64_t const cb = nQ * t / o.threads, ce = nQ * (t + 1) / o.threads;
        for (uint64_t b = cb; b < ce; b += rpb)
        {
            uint64_t const e = std::min(ce, b + rpb);
            for (int phase = 1; phase <= 2; ++phase)
                for (uint64_t q = b; q < e; ++q)
                {
                    size_t nHits = 0;
                    for (lgpu_hit const * h : perQuery[q])
                        nHits += h->phase == phase;
                    if (nHits)
                    {
                        computeRecordLca(q, phase);
                        recordHeader(q, nHits);
                    }
                    if (o.report)
                    {
                        std::vector<lgpu_hit const *> ms;
                        for (lgpu_hit const * h : perQuery[q])
                            if (h->phase == phase)
                                ms.push_back(h);
                        if (!ms.empty())
                            reportRecord(q, ms);
                        continue;
                    }
                    if (o.sam)
                    {
                        lgpu_hit const * prev = nullptr;
                        for (lgpu_hit const * h : perQuery[q])
                            if (h->phase == phase)
                            {
                                samRecord(q, h, prev, nHits);
                                prev = h;
                            }
                        continue;
                    }
                    for (lgpu_hit const * h : perQuery[q])
                        if (h->phase == phase)
                        {
                            std::string const & l = tabularLine(q, h);
                            std::fwrite(l.data(), 1, l.size(), fo);
                        }
                }
        }
    }
    if (o.comments)
        std::fprintf(fo, "# BLAST processed %llu queries\n", static_cast<unsigned long long>(nRecords));
    if (o.bam)
        bgzfDrain(fo, bamRaw, true);
    if (o.report)
    {
        std::fprintf(fo, "\n  Database: %s\n  Number of letters in database: %llu\n  Number of sequences in database:  %llu\n\n\n\n"
                         "Matrix: ",
                     o.index.c_str(), static_cast<unsigned long long>(dbLetters), static_cast<unsigned long long>(dbSeqs));
        if (isBlastN)
            std::fprintf(fo, " blastn matrix:%d %d", o.params.match, o.params.mismatch);
        else
            std::fprintf(fo, "BLOSUM%d", o.params.scoring_method);
        std::fprintf(fo, "\nGap Penalties: Existence: %d, Extension: %d\n\n", -o.params.gap_open, -o.params.gap_extend);
    }
    checkOutput(true);
    lgpu_lba_close(lba);
    double const t4 = now();

    if (o.verbosity >= 2)
    {
        std::printf("Runtime total: %.3fs (search %.3fs incl. index upload, output %.3fs)\n", t4 - t0, t3 - t2, t4 - t3);
        for (size_t g = 0; g < res.size(); ++g)
            std::printf("  GPU %zu: CUDA init + index upload %.3fs, search %.3fs, teardown %.3fs\n", g, res[g].tUpload,
                        res[g].tSearch, res[g].tTeardown);
        std::printf("\n");
        std::printf("   HITS                             Remaining\n");
        uint64_t rem = total.hits_after_seeding;
        std::printf("   after Seeding               %10llu\n", static_cast<unsigned long long>(rem));
        auto row = [&](char const * name, uint64_t v) {
            rem -= v;
            std::printf(" - %-24s %10llu = %10llu\n", name, static_cast<unsigned long long>(v),
                        static_cast<unsigned long long>(rem));
        };
        row("failed pre-extend test", total.hits_failed_pre_extend);
        row("failed e-value test", total.hits_failed_evalue);
        row("failed bitScore test", total.hits_failed_bitscore);
        row("failed %-identity test", total.hits_failed_identity);
        row("duplicates", total.hits_duplicate);
        row("late duplicates", total.hits_duplicate2);
        row("abundant", total.hits_abundant);
        std::printf("\nNumber of total hits:                           %llu\n", static_cast<unsigned long long>(total.hits_final));
        std::printf("Number of Query-Subject pairs:                  %llu\n", static_cast<unsigned long long>(total.pairs));
        std::printf("Number of Queries with at least one valid hit:  %llu\n", static_cast<unsigned long long>(total.qrys_with_hit));
        std::printf("GPU: %llu kernel launches, %.1f ms device time, %.2f Gcells pass 1, %.2f Gcells pass 2\n",
                    static_cast<unsigned long long>(total.kernel_launches), total.ms_total, total.cells_score / 1e9,
                    total.cells_trace / 1e9);
    }
    return 0;
}
