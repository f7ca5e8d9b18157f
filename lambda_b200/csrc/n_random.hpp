// 'N' in nucleotide queries during seeding.
//
// The reference searches nucleotide queries through views::dna_n_to_random (src/view_dna_n_to_random.hpp:34-63,
// plugged into the reduced query view at src/shared_definitions.hpp:277-280): every time an 'N' is READ through
// one instance of that view it becomes the next output of a std::mt19937{0xDEADBEEF} owned by the instance,
// modulo 4, as a dna4 rank (A C G T).  search() (src/search_algo.hpp:607-762) creates a new instance with
// every `lH.redQrySeqs[i]` expression, so
//   * inside one seed search (`redQrySeqs[i] | slice(seedBegin, seedBegin + seedLength)`, :659-666) the k-th
//     read of an 'N' yields kNRandom(k).  The exact and the half-exact searches read every seed position once,
//     left to right (FMC search/BacktrackingWithBuffers.h:74-83, src/search_algo.hpp:565-577), so the 'N' at
//     position p gets kNRandom(number of 'N's in [seedBegin, p));
//   * every elongation step reads ONE symbol from a fresh instance (`redQrySeqs[i][seedBegin + seedLength]`,
//     :708): an 'N' there is always kNRandom(0) = C.
// Pre-scoring and the DP use the unreduced query, where 'N' stays 'N'.
// The table below is the first 256 outputs of std::mt19937{0xDEADBEEF} % 4 (generated with libstdc++; mt19937 is
// fully specified by the C++ standard), packed two bits each.  Shared by the CUDA kernels and the CPU oracle.
#pragma once

#include <cstdint>

#if defined(__CUDACC__)
#define LGPU_HD __host__ __device__
#else
#define LGPU_HD
#endif

namespace lgpu
{

constexpr unsigned int kNMarker = 0x80u; // reduced-query byte of an 'N': kNMarker | (frame & 1)

// dna4 rank (A=0 C=1 G=2 T=3) of the k-th 'N' read through one view instance
LGPU_HD inline unsigned int nRandomRank(unsigned int k)
{
    // 256 outputs, 32 per word (numpy: MT19937()._legacy_seeding(0xDEADBEEF).random_raw(256) % 4 gives the same)
    unsigned long long const t[8] = {0xa3736c5835666461ull, 0xf83739b5e56c0330ull, 0x8a0d496c4d639a57ull, 0x6a968fce7c214217ull,
                                     0x2c7b0190e28dee09ull, 0x7d47edbd2d3f94deull, 0xee9d805c62b94c97ull, 0x4e7974c9cd6db936ull};
    k &= 255u; // a seed search reads far fewer symbols than that
    return static_cast<unsigned int>((t[k >> 5] >> (2u * (k & 31u))) & 3ull);
}

// reduced-alphabet rank of a randomised 'N': dna4 itself, or the bisulfite reduction of the frame's direction
// (src/view_reduce_to_bisulfite.hpp:51-52: forward {A, C/T, G, C/T} -> {0,1,2,1}, reverse {A/G, C, A/G, T} -> {3,4,3,5})
LGPU_HD inline unsigned int nReducedRank(unsigned int dna4Rank, bool bisulfite, unsigned int frame)
{
    if (!bisulfite)
        return dna4Rank;
    unsigned int const fwd = 0x1210u, rev = 0x5343u; // one nibble per dna4 rank
    return (((frame & 1u) ? rev : fwd) >> (4u * dna4Rank)) & 15u;
}

// Reduced symbol at absolute position p of a frame whose reduced bytes are `red` (an 'N' is stored as
// kNMarker | (frame & 1)).  seedBegin: start of the seed slice the read belongs to; elongation reads pass
// seedBegin = p (no 'N' in front of it inside its one-symbol view).
LGPU_HD inline unsigned int redSymbol(unsigned char const * red, unsigned long long seedBegin, unsigned long long p, bool bisulfite)
{
    unsigned int const s = red[p];
    if (s < kNMarker)
        return s;
    unsigned int k = 0;
    for (unsigned long long q = seedBegin; q < p; ++q)
        k += red[q] >> 7;
    return nReducedRank(nRandomRank(k), bisulfite, s & 1u);
}

} // namespace lgpu
