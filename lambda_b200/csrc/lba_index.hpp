// Host-side reader for the reference's `.lba` index files.
//
// The file is a cereal BinaryOutputArchive of index_file<> (reference
// src/shared_definitions.hpp:330-379): little-endian, no padding between fields, vectors as
// `u64 count` + raw payload.  The field order below was validated against real indexes to exact
// EOF.  We do not link cereal: the big blobs (occ blocks, sampled SA,
// sequences) are raw in the file, so the loader just maps the file and records pointers -- loading
// a multi-GB index costs one mmap instead of a streamed parse.
#pragma once

#include <cstdint>
#include <cstring>
#include <stdexcept>
#include <mutex>
#include <string>
#include <vector>

#include <fcntl.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>

#include "../../include/lambda_b200.h"

namespace lgpu
{

struct LbaError : std::runtime_error
{
    using std::runtime_error::runtime_error;
};

// sigma (= reduced alphabet size + 1) per AlphabetEnum, FMC occtable parameters derived from it
inline uint32_t alphabetSize(uint32_t alph)
{
    switch (alph)
    {
        case LGPU_ALPH_DNA3BS: return 6;   // semialphabet_any<6>
        case LGPU_ALPH_DNA4: return 4;
        case LGPU_ALPH_DNA5: return 5;
        case LGPU_ALPH_AMINO_ACID: return 27;
        case LGPU_ALPH_MURPHY10: return 10;
        case LGPU_ALPH_LI10: return 10;
        default: throw LbaError("unknown alphabet id in index header");
    }
}

inline uint32_t bitLength(uint32_t v) // required_bits(): number of bits to represent v
{
    uint32_t b = 0;
    while (v) { ++b; v >>= 1; }
    return b;
}

// Registry of the mapped index files, so that a pointer into a mapping can be turned back into
// (file descriptor, file offset): the uploader then reads the file with pread() straight into pinned
// memory instead of touching (page-faulting) 4 KiB pages of the mapping one by one.
struct LbaMapping
{
    uint8_t const * base;
    size_t          size;
    int             fd;
};
inline std::mutex & lbaRegistryMutex()
{
    static std::mutex m;
    return m;
}
inline std::vector<LbaMapping> & lbaRegistry()
{
    static std::vector<LbaMapping> r;
    return r;
}
// true and (fd, offset) if [p, p + n) lies inside a mapped index file
inline bool lbaLocate(void const * p, size_t n, int & fd, size_t & offset)
{
    std::lock_guard<std::mutex> g(lbaRegistryMutex());
    uint8_t const *             q = static_cast<uint8_t const *>(p);
    for (LbaMapping const & m : lbaRegistry())
        if (q >= m.base && q + n <= m.base + m.size)
        {
            fd     = m.fd;
            offset = static_cast<size_t>(q - m.base);
            return true;
        }
    return false;
}

class LbaFile
{
public:
    explicit LbaFile(std::string const & path)
    {
        if (path.size() >= 3 && path.compare(path.size() - 3, 3, ".gz") == 0)
            throw LbaError("compressed indexes (.gz) are not supported; decompress first");
        if (path.size() >= 4 && path.compare(path.size() - 4, 4, ".lta") == 0)
            throw LbaError("JSON indexes (.lta) are not supported; use the binary .lba format");
        fd_ = ::open(path.c_str(), O_RDONLY);
        if (fd_ < 0)
            throw LbaError("cannot open index file " + path);
        struct stat st;
        if (fstat(fd_, &st) != 0)
        {
            release();
            throw LbaError("cannot stat index file " + path);
        }
        size_ = static_cast<size_t>(st.st_size);
        base_ = static_cast<uint8_t const *>(mmap(nullptr, size_, PROT_READ, MAP_PRIVATE, fd_, 0));
        if (base_ == MAP_FAILED)
        {
            base_ = nullptr;
            release();
            throw LbaError("cannot mmap index file " + path);
        }
        madvise(const_cast<uint8_t *>(base_), size_, MADV_SEQUENTIAL);
        try
        {
            parse();
        }
        catch (...)
        {
            release(); // the destructor does not run for a throwing constructor
            throw;
        }
        std::lock_guard<std::mutex> g(lbaRegistryMutex());
        lbaRegistry().push_back({base_, size_, fd_});
    }

    ~LbaFile()
    {
        {
            std::lock_guard<std::mutex> g(lbaRegistryMutex());
            auto &                      r = lbaRegistry();
            for (size_t i = 0; i < r.size(); ++i)
                if (r[i].base == base_)
                {
                    r.erase(r.begin() + static_cast<std::ptrdiff_t>(i));
                    break;
                }
        }
        release();
    }

    LbaFile(LbaFile const &)             = delete;
    LbaFile & operator=(LbaFile const &) = delete;

    lgpu_index_desc desc{};
    uint64_t        generation  = 0;
    uint8_t         geneticCode = 0;
    // taxonomy (shared_definitions.hpp:352-356): not used by the search, only by the per-record LCA and the taxonomy
    // columns / tags of the output (src/search_algo.hpp:886-908)
    lgpu_taxonomy tax{};

private:
    void release()
    {
        if (base_)
            munmap(const_cast<uint8_t *>(base_), size_);
        if (fd_ >= 0)
            ::close(fd_);
        base_ = nullptr;
        fd_   = -1;
    }

    int             fd_   = -1;
    uint8_t const * base_ = nullptr;
    size_t          size_ = 0;
    size_t          pos_  = 0;

    template <typename T>
    T get()
    {
        if (pos_ + sizeof(T) > size_)
            throw LbaError("index file truncated");
        T v;
        std::memcpy(&v, base_ + pos_, sizeof(T));
        pos_ += sizeof(T);
        return v;
    }

    uint8_t const * blob(uint64_t bytes)
    {
        if (bytes > size_ - pos_)
            throw LbaError("index file truncated");
        uint8_t const * p = base_ + pos_;
        pos_ += bytes;
        return p;
    }

    // n * elem bytes; the product cannot wrap (a count taken from the file is checked against the bytes that are left)
    uint8_t const * blobN(uint64_t n, uint64_t elem)
    {
        if (elem && n > (size_ - pos_) / elem)
            throw LbaError("index file truncated");
        return blob(n * elem);
    }

    // `u64 n` + n * elem bytes
    uint8_t const * vec(uint64_t elem, uint64_t & n)
    {
        n = get<uint64_t>();
        return blobN(n, elem);
    }

    // delimiters of a concatenated container: non-decreasing, the last one = size of the payload
    static void checkDelims(uint64_t const * d, uint64_t n, uint64_t payload, char const * what)
    {
        if (n == 0 || d[0] != 0 || d[n - 1] != payload)
            throw LbaError(std::string("index file corrupt: ") + what + " delimiters do not match their payload");
        for (uint64_t i = 1; i < n; ++i)
            if (d[i] < d[i - 1])
                throw LbaError(std::string("index file corrupt: ") + what + " delimiters are not sorted");
    }

    void parse()
    {
        // index_file_options (shared_definitions.hpp:316-343)
        generation      = get<uint64_t>();
        if (generation != 0)
            throw LbaError("unsupported index generation " + std::to_string(generation) +
                           " (this build reads generation 0)");
        desc.index_type = get<uint8_t>();
        desc.orig_alph  = get<uint8_t>();
        desc.trans_alph = get<uint8_t>();
        desc.red_alph   = get<uint8_t>();
        geneticCode     = get<uint8_t>();
        if (desc.index_type != LGPU_INDEX_FM)
            throw LbaError("bidirectional FM indexes are out of scope (LAMBDA_WITH_BIFM is off by default)");

        uint64_t n;
        // ids: concatenated_sequences<std::string> = data string + delimiters
        desc.ids       = reinterpret_cast<char const *>(vec(1, n));
        uint64_t const nIdBytes = n;
        desc.id_delims = reinterpret_cast<uint64_t const *>(vec(8, n));
        uint64_t const nIdDelims = n;
        // seqs: custom save (shared_definitions.hpp:290-307) = u64 n + n bytes; then delimiters
        desc.seqs       = vec(1, n);
        desc.n_residues = n;
        desc.seq_delims = reinterpret_cast<uint64_t const *>(vec(8, n));
        if (n == 0 || n != nIdDelims)
            throw LbaError("index file corrupt: id/sequence delimiter count mismatch");
        desc.n_seqs = n - 1;
        if (desc.seq_delims[desc.n_seqs] != desc.n_residues)
            throw LbaError("index file corrupt: last sequence delimiter != residue count");
        checkDelims(desc.id_delims, nIdDelims, nIdBytes, "id");
        checkDelims(desc.seq_delims, n, desc.n_residues, "sequence");
        // sTaxIds, taxonParentIDs, taxonHeights, taxonNames
        tax.s_tax_ids    = reinterpret_cast<uint32_t const *>(vec(4, tax.n_s_tax_ids));
        tax.s_tax_delims = reinterpret_cast<uint64_t const *>(vec(8, n));
        if (n != desc.n_seqs + 1)
        {
            if (tax.n_s_tax_ids != 0)
                throw LbaError("index file corrupt: taxonomy ids without one delimiter per subject");
            tax.s_tax_delims = nullptr; // index built without --acc-tax-map (the empty container holds one delimiter)
        }
        else
            checkDelims(tax.s_tax_delims, n, tax.n_s_tax_ids, "taxonomy id");
        tax.taxon_parents = reinterpret_cast<uint32_t const *>(vec(4, tax.n_taxa));
        tax.taxon_heights = vec(1, n);
        if (n != tax.n_taxa)
            throw LbaError("index file corrupt: taxonomy height / parent count mismatch");
        tax.taxon_names = reinterpret_cast<char const *>(vec(1, n));
        uint64_t const nNameBytes = n;
        tax.taxon_name_delims     = reinterpret_cast<uint64_t const *>(vec(8, n));
        if (tax.n_taxa == 0 || n != tax.n_taxa + 1)
            tax.taxon_name_delims = nullptr; // index built without --tax-dump-dir
        else
            checkDelims(tax.taxon_name_delims, n, nNameBytes, "taxon name");
        // every tax id / parent must index the taxon arrays (the host walks them for the lowest common ancestor)
        if (tax.n_taxa)
        {
            for (uint64_t i = 0; i < tax.n_s_tax_ids; ++i)
                if (tax.s_tax_ids[i] >= tax.n_taxa)
                    throw LbaError("index file corrupt: subject tax id outside the taxonomy");
            for (uint64_t i = 0; i < tax.n_taxa; ++i)
                if (tax.taxon_parents[i] >= tax.n_taxa)
                    throw LbaError("index file corrupt: taxon parent outside the taxonomy");
        }

        // index.occ (InterleavedEPRV2.h:272-306)
        uint32_t const redSize = alphabetSize(desc.red_alph);
        desc.sigma             = redSize + 1;
        desc.sigma_bits        = bitLength(desc.sigma - 1);
        desc.planes_offset     = (4 * desc.sigma + 7) / 8 * 8;
        desc.block_bytes       = desc.planes_offset + 8 * desc.sigma_bits;
        int32_t version        = get<int32_t>();
        if (version != 1)
            throw LbaError("occ table was not written with cereal's binary fast path (version != 1)");
        desc.n_blocks     = get<uint64_t>();
        desc.occ_blocks   = blobN(desc.n_blocks, desc.block_bytes);
        desc.n_super      = get<uint64_t>();
        desc.super_blocks = reinterpret_cast<uint64_t const *>(blobN(desc.n_super, static_cast<uint64_t>(desc.sigma) * 8));
        desc.C            = reinterpret_cast<uint64_t const *>(blob((desc.sigma + 1) * 8));

        // index.csa (CSA.h:115-118, BitvectorCompact.h:112-143)
        desc.ssa = reinterpret_cast<uint64_t const *>(vec(8, desc.n_ssa));
        version  = get<int32_t>();
        if (version != 1)
            throw LbaError("csa bit vector was not written with cereal's binary fast path");
        desc.n_csa_sb          = get<uint64_t>();
        desc.csa_bv            = blobN(desc.n_csa_sb, 48);
        desc.sampling_rate     = get<uint64_t>();
        desc.bits_for_position = get<uint64_t>();
        uint64_t const mask    = get<uint64_t>();
        if (desc.bits_for_position == 0 || desc.bits_for_position >= 64 ||
            mask != ((1ull << desc.bits_for_position) - 1))
            throw LbaError("index file corrupt: bad CSA position mask");
        if (pos_ != size_)
            throw LbaError("index file has " + std::to_string(size_ - pos_) + " trailing bytes");
        // size() of the index = C.back() must be addressable in the block array, the super-block rows and the CSA's
        // bit vector (bit i of the vector lives at position i + 1)
        uint64_t const nRows = desc.C[desc.sigma];
        if (nRows / 64 >= desc.n_blocks)
            throw LbaError("index file corrupt: occ blocks do not cover the BWT");
        if (desc.n_super == 0 || (nRows >> 32) >= desc.n_super)
            throw LbaError("index file corrupt: occ super blocks do not cover the BWT");
        for (uint32_t s2 = 0; s2 < desc.sigma; ++s2)
            if (desc.C[s2] > desc.C[s2 + 1])
                throw LbaError("index file corrupt: C array is not sorted");
        if (nRows / 256 >= desc.n_csa_sb)
            throw LbaError("index file corrupt: CSA bit vector does not cover the BWT");
        // sequences of the index text: two conversions per subject (bisulfite), six frames (translated subjects)
        uint64_t const perSubject = desc.red_alph == LGPU_ALPH_DNA3BS ? 2
                                    : (desc.trans_alph == LGPU_ALPH_AMINO_ACID && desc.orig_alph == LGPU_ALPH_DNA5) ? 6 : 1;
        if (desc.n_seqs > (1ull << (64 - desc.bits_for_position)) / perSubject)
            throw LbaError("index file corrupt: sequence ids do not fit the CSA entries");
        // The O(n) part of the cross-checks -- every sampled suffix-array entry names an existing (sequence, position),
        // every CSA super block ranks inside the sampled array -- runs on the device after the upload
        // (validateIndexKernel, engine.cu), where it costs a fraction of a millisecond.
    }
};

} // namespace lgpu
