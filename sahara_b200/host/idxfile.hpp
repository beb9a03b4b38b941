// idxfile.hpp — reader / writer of the `X.idx` file of `sahara index`.
//
// Replaces the cereal calls `archive(Sigma); archive(index)` of /root/reference/src/sahara/index.cpp:96-100
// (writer) and /root/reference/src/sahara/search.cpp:162-169, 278-283 (reader) for
// fmc::BiFMIndex<Sigma, fmc::string::InterleavedBitvector16>.
//
// cereal's binary archive rules: little-endian raw values, a u64 element count in front
// of every std::vector, std::array of arithmetic types raw.  The FIELD ORDER inside BiFMIndex is a
// reconstruction (fmindex-collection 1.1.0 is not available here, see DESIGN.md "index file layout"):
//   u64 sigma
//   bwt     : u64 nBlocks, nBlocks x { u16 cnt[sigma]; u64 bits[sigma] }, u64 nSuper, nSuper x u64[sigma], u64 rows
//   bwtRev  : same
//   C       : u64[sigma + 1]
//   csa     : u64 nSamples, u64 ssa[nSamples], u64 nWords, u64 markBits[nWords], u64 rows,
//             u64 samplingRate, u64 bitsForPosition
// The loader checks every size relation it can and refuses files it does not understand; the deep
// consistency checks (bitplanes disjoint, counters, histograms) run on the GPU in sb200_index_upload.
#pragma once
#include <cstdint>
#include <cstdio>
#include <stdexcept>
#include <string>
#include <vector>

#include "../../include/sahara_b200.h"

namespace sahara {

struct IndexImage {
    uint64_t sigma{}, n_rows{}, n_blocks{};
    std::vector<uint8_t> bwt_blocks, bwtrev_blocks;
    std::vector<uint64_t> bwt_super, bwtrev_super;
    std::vector<uint64_t> C, ssa, mark_bits;
    uint64_t sampling_rate{}, bits_for_position{};

    sb200_index_view view() const {
        sb200_index_view v{};
        v.sigma = sigma;
        v.n_rows = n_rows;
        v.n_blocks = n_blocks;
        v.bwt_blocks = bwt_blocks.data();
        v.bwt_super = bwt_super.data();
        v.bwtrev_blocks = bwtrev_blocks.data();
        v.bwtrev_super = bwtrev_super.data();
        v.C = C.data();
        v.ssa = ssa.data();
        v.n_ssa = ssa.size();
        v.mark_bits = mark_bits.data();
        v.sampling_rate = sampling_rate;
        v.bits_for_position = bits_for_position;
        return v;
    }
};

namespace detail {
struct File {
    FILE* f{};
    File(std::string const& path, char const* mode) : f(fopen(path.c_str(), mode)) {}
    ~File() { if (f) fclose(f); }
    File(File const&) = delete;
};
inline void readRaw(FILE* f, void* p, size_t n) {
    if (n && fread(p, 1, n, f) != n) throw std::runtime_error("index layout not understood: unexpected end of file");
}
inline uint64_t readU64(FILE* f) {
    uint64_t v;
    readRaw(f, &v, 8);
    return v;
}
inline void writeRaw(FILE* f, void const* p, size_t n) {
    if (n && fwrite(p, 1, n, f) != n) throw std::runtime_error("writing the index failed");
}
inline void writeU64(FILE* f, uint64_t v) { writeRaw(f, &v, 8); }
}  // namespace detail

// first 8 bytes of the file, as /root/reference/src/sahara/search.cpp:278-283
inline uint64_t peekSigma(std::string const& path) {
    detail::File file(path, "rb");
    if (!file.f) throw std::runtime_error("no valid index path at " + path);
    return detail::readU64(file.f);
}

inline IndexImage loadIndexFile(std::string const& path) {
    using namespace detail;
    File file(path, "rb");
    if (!file.f) throw std::runtime_error("no valid index path at " + path);
    FILE* f = file.f;
    IndexImage im;
    im.sigma = readU64(f);
    if (im.sigma != 5 && im.sigma != 6) throw std::runtime_error("unknown index with " + std::to_string(im.sigma) + " letters");
    auto readOcc = [&](std::vector<uint8_t>& blocks, std::vector<uint64_t>& super, uint64_t& rows, uint64_t& nBlocks) {
        nBlocks = readU64(f);
        if (nBlocks == 0 || nBlocks > (uint64_t{1} << 40)) throw std::runtime_error("index layout not understood: block count");
        blocks.resize(nBlocks * 10 * im.sigma);
        readRaw(f, blocks.data(), blocks.size());
        uint64_t nSuper = readU64(f);
        if (nSuper != (nBlocks + 1023) / 1024) throw std::runtime_error("index layout not understood: superblock count");
        super.resize(nSuper * im.sigma);
        readRaw(f, super.data(), super.size() * 8);
        rows = readU64(f);
        if (rows / 64 + 1 != nBlocks) throw std::runtime_error("index layout not understood: row count vs block count");
    };
    uint64_t rowsRev = 0, blocksRev = 0;
    readOcc(im.bwt_blocks, im.bwt_super, im.n_rows, im.n_blocks);
    readOcc(im.bwtrev_blocks, im.bwtrev_super, rowsRev, blocksRev);
    if (rowsRev != im.n_rows) throw std::runtime_error("index layout not understood: bwt and bwtRev differ in size");
    im.C.resize(im.sigma + 1);
    readRaw(f, im.C.data(), im.C.size() * 8);
    uint64_t nSamples = readU64(f);
    if (nSamples > im.n_rows) throw std::runtime_error("index layout not understood: sample count");
    im.ssa.resize(nSamples);
    readRaw(f, im.ssa.data(), nSamples * 8);
    uint64_t nWords = readU64(f);
    if (nWords != im.n_rows / 64 + 1) throw std::runtime_error("index layout not understood: marker words");
    im.mark_bits.resize(nWords);
    readRaw(f, im.mark_bits.data(), nWords * 8);
    if (readU64(f) != im.n_rows) throw std::runtime_error("index layout not understood: marker length");
    im.sampling_rate = readU64(f);
    im.bits_for_position = readU64(f);
    uint8_t extra;
    if (fread(&extra, 1, 1, f) != 0) throw std::runtime_error("index layout not understood: trailing bytes");
    return im;
}

inline void saveIndexFile(std::string const& path, sb200_index_view const& v) {
    using namespace detail;
    File file(path, "wb");
    if (!file.f) throw std::runtime_error("cannot open " + path + " for writing");
    FILE* f = file.f;
    writeU64(f, v.sigma);
    uint64_t nSuper = (v.n_blocks + 1023) / 1024;
    auto writeOcc = [&](void const* blocks, uint64_t const* super) {
        writeU64(f, v.n_blocks);
        writeRaw(f, blocks, v.n_blocks * 10 * v.sigma);
        writeU64(f, nSuper);
        writeRaw(f, super, nSuper * v.sigma * 8);
        writeU64(f, v.n_rows);
    };
    writeOcc(v.bwt_blocks, v.bwt_super);
    writeOcc(v.bwtrev_blocks, v.bwtrev_super);
    writeRaw(f, v.C, (v.sigma + 1) * 8);
    writeU64(f, v.n_ssa);
    writeRaw(f, v.ssa, v.n_ssa * 8);
    uint64_t nWords = v.n_rows / 64 + 1;
    writeU64(f, nWords);
    writeRaw(f, v.mark_bits, nWords * 8);
    writeU64(f, v.n_rows);
    writeU64(f, v.sampling_rate);
    writeU64(f, v.bits_for_position);
}

}  // namespace sahara
