// alphabet.hpp — d_dna4 / d_dna5 rank alphabets with delimiter, char<->rank, reverse complement.
//
// Stands in for ivs::d_dna4 / ivs::d_dna5, ivs::convert_char_to_rank, ivs::verify_rank and
// ivs::reverse_complement_rank as used at /root/reference/src/sahara/search.cpp:115-124 and
// /root/reference/src/sahara/index.cpp:53-72 (IVSigma 0.4.1 is not vendored; semantics per SURVEY.md §9.1):
// rank 0 = '$' (delimiter), A=1, C=2, G=3, T=4 (U and lower case fold), N=5 only in d_dna5;
// every other character converts to 255 (invalid).
#pragma once
#include <array>
#include <cstdint>
#include <optional>
#include <span>
#include <string_view>
#include <vector>

namespace sahara {

constexpr uint8_t kInvalidRank = 255;

template <bool WithN>
struct DnaAlphabet {
    static constexpr size_t size() { return WithN ? 6 : 5; }
    static constexpr std::array<uint8_t, 256> makeTable() {
        std::array<uint8_t, 256> t{};
        for (auto& v : t) v = kInvalidRank;
        t['$'] = 0;
        t['A'] = t['a'] = 1;
        t['C'] = t['c'] = 2;
        t['G'] = t['g'] = 3;
        t['T'] = t['t'] = t['U'] = t['u'] = 4;
        if (WithN) t['N'] = t['n'] = 5;
        return t;
    }
    static constexpr std::array<uint8_t, 256> table = makeTable();
    static constexpr uint8_t char_to_rank(char c) { return table[static_cast<uint8_t>(c)]; }
    static constexpr char rank_to_char(uint8_t r) { return r < size() ? "$ACGTN"[r] : '?'; }
    // complement on ranks: A<->T, C<->G, '$' and N unchanged
    static constexpr uint8_t complement_rank(uint8_t r) { return (r >= 1 && r <= 4) ? static_cast<uint8_t>(5 - r) : r; }
};
using d_dna4 = DnaAlphabet<false>;
using d_dna5 = DnaAlphabet<true>;

template <typename Alphabet>
std::vector<uint8_t> convert_char_to_rank(std::string_view seq) {
    std::vector<uint8_t> r(seq.size());
    for (size_t i = 0; i < seq.size(); ++i) r[i] = Alphabet::char_to_rank(seq[i]);
    return r;
}

inline bool verify_rank(uint8_t v) { return v != kInvalidRank; }

// position of the first invalid rank, if any
inline std::optional<size_t> verify_rank(std::span<uint8_t const> ranks) {
    for (size_t i = 0; i < ranks.size(); ++i)
        if (ranks[i] == kInvalidRank) return i;
    return std::nullopt;
}

template <typename Alphabet>
std::vector<uint8_t> reverse_complement_rank(std::span<uint8_t const> ranks) {
    std::vector<uint8_t> r(ranks.size());
    for (size_t i = 0; i < ranks.size(); ++i) r[i] = Alphabet::complement_rank(ranks[ranks.size() - 1 - i]);
    return r;
}

}  // namespace sahara
