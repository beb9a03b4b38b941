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
#include <algorithm>
#include <cstdint>
#include <cstdio>
#include <memory>
#include <stdexcept>
#include <string>
#include <vector>

#include "../../include/sahara_b200.h"

namespace sahara {

// storage the file is read into: uninitialised (a std::vector would write 8 GB of zeros first for a human-sized index)
template <typename T>
struct RawBuf {
    std::unique_ptr<T[]> p;
    size_t n{0};
    void resize(size_t count) {
        p.reset(count ? new T[count] : nullptr);
        n = count;
    }
    T* data() { return p.get(); }
    T const* data() const { return p.get(); }
    size_t size() const { return n; }
    T* begin() { return p.get(); }
};

struct IndexImage {
    uint64_t sigma{}, n_rows{}, n_blocks{};
    RawBuf<uint8_t> bwt_blocks, bwtrev_blocks;
    RawBuf<uint64_t> bwt_super, bwtrev_super;
    RawBuf<uint64_t> C, ssa, mark_bits;
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

// ---- the layout assumptions, one constant each (none of them has been checked against a file written by
// upstream sahara — tools/pin_against_sahara.sh does that where a real binary exists) ---------------------------------
namespace layout {
constexpr uint64_t kRowsPerBlock = 64;          // InterleavedBitvector16: one block per 64 rows
constexpr uint64_t kBlocksPerSuperblock = 1024; // u16 block counters: a superblock every 65536 rows
constexpr uint64_t kBlockAlign = 64;            // upstream declares the block alignas(64): the padded stride the loader also accepts
// bytes of one block as written: u16 counters + u64 one-hot bit words per symbol, packed
constexpr uint64_t blockBytesPacked(uint64_t sigma) { return (2 + 8) * sigma; }
constexpr uint64_t blockBytesPadded(uint64_t sigma) { return (blockBytesPacked(sigma) + kBlockAlign - 1) / kBlockAlign * kBlockAlign; }
constexpr uint64_t blocksForRows(uint64_t rows) { return rows / kRowsPerBlock + 1; }
constexpr uint64_t superblocksForBlocks(uint64_t blocks) { return (blocks + kBlocksPerSuperblock - 1) / kBlocksPerSuperblock; }
constexpr uint64_t markerWordsForRows(uint64_t rows) { return rows / 64 + 1; }
}  // namespace layout

namespace detail {
[[noreturn]] inline void layoutError(std::string const& what) {
    throw std::runtime_error("index layout not understood: " + what +
                             " (the field order of X.idx is reconstructed, see sahara_b200/host/idxfile.hpp; an index written by "
                             "this repo's `sahara index` always loads)");
}
struct File {
    FILE* f{};
    File(std::string const& path, char const* mode) : f(fopen(path.c_str(), mode)) {}
    ~File() { if (f) fclose(f); }
    File(File const&) = delete;
};
inline void readRaw(FILE* f, void* p, size_t n) {
    if (n && fread(p, 1, n, f) != n) layoutError("unexpected end of file");
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
    auto readOcc = [&](RawBuf<uint8_t>& blocks, RawBuf<uint64_t>& super, uint64_t& rows, uint64_t& nBlocks) {
        nBlocks = readU64(f);
        if (nBlocks == 0 || nBlocks > (uint64_t{1} << 40)) layoutError("block count " + std::to_string(nBlocks) + " where a u64 vector size was expected");
        const uint64_t packed = layout::blockBytesPacked(im.sigma), padded = layout::blockBytesPadded(im.sigma);
        const uint64_t wantSuper = layout::superblocksForBlocks(nBlocks);
        // which stride were the blocks written with?  The u64 behind them must be the superblock count.
        const long at = ftell(f);
        uint64_t stride = 0;
        for (uint64_t cand : {packed, padded}) {
            uint64_t v = 0;
            if (fseek(f, at + static_cast<long>(nBlocks * cand), SEEK_SET) == 0 && fread(&v, 8, 1, f) == 1 && v == wantSuper) {
                stride = cand;
                break;
            }
        }
        if (stride == 0)
            layoutError("no superblock count of " + std::to_string(wantSuper) + " (one per " + std::to_string(layout::kBlocksPerSuperblock) + " blocks) behind " +
                        std::to_string(nBlocks) + " blocks of " + std::to_string(packed) + " (packed) or " + std::to_string(padded) + " (64-byte aligned) bytes");
        fseek(f, at, SEEK_SET);
        blocks.resize(nBlocks * packed);
        if (stride == packed) readRaw(f, blocks.data(), blocks.size());
        else {  // padded blocks: keep the leading counters + bit words
            std::vector<uint8_t> buf(stride);
            for (uint64_t b = 0; b < nBlocks; ++b) {
                readRaw(f, buf.data(), stride);
                std::copy(buf.begin(), buf.begin() + static_cast<long>(packed), blocks.data() + b * packed);
            }
        }
        uint64_t nSuper = readU64(f);
        super.resize(nSuper * im.sigma);
        readRaw(f, super.data(), super.size() * 8);
        rows = readU64(f);
        if (layout::blocksForRows(rows) != nBlocks)
            layoutError("row count " + std::to_string(rows) + " does not give " + std::to_string(nBlocks) + " blocks of " + std::to_string(layout::kRowsPerBlock) + " rows");
    };
    uint64_t rowsRev = 0, blocksRev = 0;
    readOcc(im.bwt_blocks, im.bwt_super, im.n_rows, im.n_blocks);
    readOcc(im.bwtrev_blocks, im.bwtrev_super, rowsRev, blocksRev);
    if (rowsRev != im.n_rows) layoutError("bwt and bwtRev differ in size (expected: bwt, bwtRev, C, csa in this order)");
    im.C.resize(im.sigma + 1);
    readRaw(f, im.C.data(), im.C.size() * 8);
    uint64_t nSamples = readU64(f);
    if (nSamples > im.n_rows) layoutError("sample count " + std::to_string(nSamples) + " exceeds the row count (expected the csa's sample vector behind C[sigma + 1])");
    im.ssa.resize(nSamples);
    readRaw(f, im.ssa.data(), nSamples * 8);
    uint64_t nWords = readU64(f);
    if (nWords != layout::markerWordsForRows(im.n_rows)) layoutError("marker bitvector of " + std::to_string(nWords) + " words, expected rows / 64 + 1 behind the samples");
    im.mark_bits.resize(nWords);
    readRaw(f, im.mark_bits.data(), nWords * 8);
    if (readU64(f) != im.n_rows) layoutError("marker bitvector length differs from the row count");
    im.sampling_rate = readU64(f);
    im.bits_for_position = readU64(f);
    uint8_t extra;
    if (fread(&extra, 1, 1, f) != 0) layoutError("trailing bytes behind samplingRate / bitsForPosition");
    return im;
}

inline void saveIndexFile(std::string const& path, sb200_index_view const& v) {
    using namespace detail;
    File file(path, "wb");
    if (!file.f) throw std::runtime_error("cannot open " + path + " for writing");
    FILE* f = file.f;
    writeU64(f, v.sigma);
    uint64_t nSuper = layout::superblocksForBlocks(v.n_blocks);
    auto writeOcc = [&](void const* blocks, uint64_t const* super) {
        writeU64(f, v.n_blocks);
        writeRaw(f, blocks, v.n_blocks * layout::blockBytesPacked(v.sigma));
        writeU64(f, nSuper);
        writeRaw(f, super, nSuper * v.sigma * 8);
        writeU64(f, v.n_rows);
    };
    writeOcc(v.bwt_blocks, v.bwt_super);
    writeOcc(v.bwtrev_blocks, v.bwtrev_super);
    writeRaw(f, v.C, (v.sigma + 1) * 8);
    writeU64(f, v.n_ssa);
    writeRaw(f, v.ssa, v.n_ssa * 8);
    uint64_t nWords = layout::markerWordsForRows(v.n_rows);
    writeU64(f, nWords);
    writeRaw(f, v.mark_bits, nWords * 8);
    writeU64(f, v.n_rows);
    writeU64(f, v.sampling_rate);
    writeU64(f, v.bits_for_position);
}

}  // namespace sahara
