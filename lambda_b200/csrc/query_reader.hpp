// Query file reader of the command-line host: FASTA and FASTQ, plain or gzip/BGZF-compressed.
//
// Mirrors what the reference accepts through bio::io::seq::reader (src/search_algo.hpp:342-348):
//   * the format follows the file extension, looked at after a compression extension was stripped
//     (BIO-IO format/fasta.hpp:60-68: fasta fa fna ffn faa frn fas; format/fastq.hpp:55: fastq fq);
//   * compression is detected from the magic bytes, not from the name (stream/transparent_istream.hpp) --
//     zlib's gzread does the same: it inflates gzip members (BGZF is a series of them) and passes anything else through;
//   * FASTA (format/fasta_input_handler.hpp): a record starts with '>' or ';', the id is the rest of that line
//     (truncate_ids = false, seq/reader_options.hpp:103), the sequence is every following line up to the next id
//     line with white space and digits dropped; a record without sequence characters is an error;
//   * FASTQ (format/fastq_input_handler.hpp): four lines per record ('@' id, sequence, '+...', qualities of the
//     same length); an empty sequence is allowed.
// bzip2 / zstd inputs (optional in the reference's build) are rejected with a message instead of being misparsed.
#pragma once

#include <cstdio>
#include <cstring>
#include <stdexcept>
#include <string>
#include <zlib.h>

namespace qio
{

enum class Format
{
    Fasta,
    Fastq
};

inline bool endsWithNoCase(std::string const & s, char const * suffix)
{
    size_t const n = std::strlen(suffix);
    if (s.size() < n)
        return false;
    for (size_t i = 0; i < n; ++i)
        if (std::tolower(static_cast<unsigned char>(s[s.size() - n + i])) != suffix[i])
            return false;
    return true;
}

// the extension decides (after one compression extension); anything else is the reference's
// "unhandled extension" error
inline Format formatOf(std::string path)
{
    for (char const * z : {".gz", ".bgzf", ".bz2", ".zst"})
        if (endsWithNoCase(path, z))
        {
            path.resize(path.size() - std::strlen(z));
            break;
        }
    for (char const * e : {".fasta", ".fa", ".fna", ".ffn", ".faa", ".frn", ".fas"})
        if (endsWithNoCase(path, e))
            return Format::Fasta;
    for (char const * e : {".fastq", ".fq"})
        if (endsWithNoCase(path, e))
            return Format::Fastq;
    throw std::runtime_error("The query file's extension is not handled (expected one of fasta fa fna ffn faa frn fas "
                             "fastq fq, optionally followed by .gz / .bgzf): " + path);
}

// buffered line reader over zlib (transparent for uncompressed files)
class LineReader
{
public:
    explicit LineReader(std::string const & path) : path_(path)
    {
        // refuse the two compressions zlib would pass through as "plain" bytes
        if (FILE * fp = std::fopen(path.c_str(), "rb"))
        {
            unsigned char m[4] = {0, 0, 0, 0};
            size_t const  n    = std::fread(m, 1, 4, fp);
            std::fclose(fp);
            if (n >= 3 && m[0] == 'B' && m[1] == 'Z' && m[2] == 'h')
                throw std::runtime_error("bzip2-compressed query files are not supported (use gzip): " + path);
            if (n == 4 && m[0] == 0x28 && m[1] == 0xB5 && m[2] == 0x2F && m[3] == 0xFD)
                throw std::runtime_error("zstd-compressed query files are not supported (use gzip): " + path);
        }
        f_ = gzopen(path.c_str(), "rb");
        if (!f_)
            throw std::runtime_error("Could not open file " + path + " for reading.");
        gzbuffer(f_, 1u << 20);
        buf_.resize(1u << 20);
    }
    LineReader(LineReader const &)             = delete;
    LineReader & operator=(LineReader const &) = delete;
    ~LineReader()
    {
        if (f_)
            gzclose(f_);
    }

    // next line without its terminator ("\n" or "\r\n"); false at end of file
    bool getline(std::string & line)
    {
        line.clear();
        bool any = false;
        for (;;)
        {
            if (pos_ == len_)
            {
                int const n = gzread(f_, buf_.data(), static_cast<unsigned int>(buf_.size()));
                if (n < 0)
                {
                    int          err = 0;
                    char const * msg = gzerror(f_, &err);
                    throw std::runtime_error("error while reading " + path_ + ": " + (msg ? msg : "zlib error"));
                }
                if (n == 0)
                {
                    checkCleanEnd();
                    break;
                }
                pos_ = 0;
                len_ = static_cast<size_t>(n);
            }
            any              = true;
            char const * b   = buf_.data() + pos_;
            char const * nl  = static_cast<char const *>(std::memchr(b, '\n', len_ - pos_));
            size_t const cnt = nl ? static_cast<size_t>(nl - b) : len_ - pos_;
            line.append(b, cnt);
            pos_ += cnt + (nl ? 1 : 0);
            if (nl)
                break;
        }
        if (!line.empty() && line.back() == '\r')
            line.pop_back();
        ++lineNo_;
        return any;
    }

    // first character of the next line without consuming it; -1 at end of file
    int peek()
    {
        if (pos_ == len_)
        {
            int const n = gzread(f_, buf_.data(), static_cast<unsigned int>(buf_.size()));
            if (n < 0)
                checkCleanEnd();
            if (n <= 0)
            {
                checkCleanEnd();
                return -1;
            }
            pos_ = 0;
            len_ = static_cast<size_t>(n);
        }
        return static_cast<unsigned char>(buf_[pos_]);
    }

    size_t lineNo() const { return lineNo_; }

private:
    // end of data: a compressed stream that stops in the middle (truncated download) is an error, not a short file
    void checkCleanEnd()
    {
        int          err = Z_OK;
        char const * msg = gzerror(f_, &err);
        if (err != Z_OK && err != Z_STREAM_END)
            throw std::runtime_error("error while reading " + path_ + ": " + (msg && *msg ? msg : "truncated or corrupt compressed stream"));
    }

    std::string path_;
    gzFile      f_ = nullptr;
    std::string buf_;
    size_t      pos_ = 0, len_ = 0, lineNo_ = 0;
};

struct Record
{
    std::string id, seq;
};

class SeqReader
{
public:
    explicit SeqReader(std::string const & path) : path_(path), format_(formatOf(path)), in_(path) {}

    Format format() const { return format_; }

    // next record; false at the (clean) end of the file.  Malformed input throws like the reference's parser.
    bool next(Record & r)
    {
        r.id.clear();
        r.seq.clear();
        // blank lines between records / at the end of the file are skipped (more lenient than the reference)
        for (int c = in_.peek(); c == '\n' || c == '\r'; c = in_.peek())
            in_.getline(line_);
        if (in_.peek() < 0)
            return false;
        return format_ == Format::Fasta ? nextFasta(r) : nextFastq(r);
    }

private:
    [[noreturn]] void error(std::string const & what)
    {
        throw std::runtime_error("[query file " + path_ + ", line " + std::to_string(in_.lineNo()) + "] " + what);
    }

    static bool isIdStart(int c) { return c == '>' || c == ';'; }

    bool nextFasta(Record & r)
    {
        if (!in_.getline(line_))
            return false;
        if (line_.empty())
            error("Expected to be on begin of record but is on empty line.");
        if (!isIdStart(static_cast<unsigned char>(line_[0])))
            error("Record does not begin with '>' or ';'.");
        r.id.assign(line_, 1, std::string::npos);
        for (int c = in_.peek(); c >= 0 && !isIdStart(c); c = in_.peek())
        {
            in_.getline(line_);
            for (char ch : line_)
                if (!std::isspace(static_cast<unsigned char>(ch)) && !std::isdigit(static_cast<unsigned char>(ch)))
                    r.seq.push_back(ch);
        }
        if (r.seq.empty())
            error("No sequence or no valid sequence characters.");
        return true;
    }

    bool nextFastq(Record & r)
    {
        if (!in_.getline(line_))
            return false;
        if (line_.empty())
            error("Expected to be on begin of record but line is empty.");
        if (line_[0] != '@')
            error("ID-line does not begin with '@'.");
        r.id.assign(line_, 1, std::string::npos);
        if (!in_.getline(r.seq))
            error("Reached end of file while trying to read SEQ.");
        if (!in_.getline(line_))
            error("Reached end of file while trying to read third FastQ record line.");
        if (line_.empty() || line_[0] != '+')
            error("Third FastQ record line does not begin with '+'.");
        if (!in_.getline(line_))
            error("Reached end of file while trying to read QUALITIES.");
        if (line_.size() != r.seq.size())
            error("Size mismatch between sequence (" + std::to_string(r.seq.size()) + ") and qualities (" +
                  std::to_string(line_.size()) + ").");
        return true;
    }

    std::string path_;
    Format      format_;
    LineReader  in_;
    std::string line_;
};

} // namespace qio
