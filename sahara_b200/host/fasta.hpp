// fasta.hpp — minimal FASTA reader/writer.
//
// Stands in for ivio::fasta::reader as used at /root/reference/src/sahara/search.cpp:115 and
// /root/reference/src/sahara/index.cpp:53 (IVio 1.2.1 is not vendored): a record is a '>' header line
// (id = the line without '>') followed by sequence lines that are concatenated without line breaks.
#pragma once
#include <cstring>
#include <algorithm>
#include <array>
#include <cstdint>
#include <cstdio>
#include <fstream>
#include <functional>
#include <stdexcept>
#include <string>
#include <thread>
#include <vector>

namespace sahara::fasta {

struct Record {
    std::string id;
    std::string seq;
};

// calls cb(record) for every record; throws on I/O errors or data before the first header
inline void read(std::string const& path, std::function<void(Record&)> const& cb) {
    std::ifstream in(path, std::ios::binary);
    if (!in) throw std::runtime_error("cannot open fasta file " + path);
    std::vector<char> buf(1 << 22);
    in.rdbuf()->pubsetbuf(buf.data(), static_cast<std::streamsize>(buf.size()));
    Record rec;
    bool have = false;
    std::string line;
    while (std::getline(in, line)) {
        if (!line.empty() && line.back() == '\r') line.pop_back();
        if (!line.empty() && line[0] == '>') {
            if (have) cb(rec);
            rec.id.assign(line, 1, std::string::npos);
            rec.seq.clear();
            have = true;
        } else if (!line.empty()) {
            if (!have) throw std::runtime_error("fasta file " + path + " does not start with a '>' header");
            rec.seq += line;
        }
    }
    if (have) cb(rec);
}

inline std::vector<Record> readAll(std::string const& path) {
    std::vector<Record> r;
    read(path, [&](Record& rec) { r.push_back(rec); });
    return r;
}

// ---- read sets: all records converted to ranks, in parallel ---------------------------------------------
// The query file of `sahara search` (search.cpp:115-124) holds millions of short records; going through
// std::getline and one std::string per record costs more than the GPU search.  The file is cut at record
// boundaries into one piece per thread; every piece is converted straight into rank bytes (table[c], 255 =
// invalid) and the pieces are concatenated.  The first problem in file order is reported, as a sequential reader
// would: an invalid character (ivs::verify_rank) before a record of a different length.
struct ReadSet {
    std::vector<uint8_t> ranks;  // count * len rank bytes, record after record
    size_t count{0}, len{0};
    // first problem in file order (record = 0-based index; kind 0 = none, 1 = invalid character, 2 = length differs)
    int problem{0};
    size_t record{0}, pos{0}, length{0};
    char ch{0};
    std::string id;
};

inline ReadSet readRanksParallel(std::string const& path, std::array<uint8_t, 256> const& table, unsigned threads) {
    std::ifstream in(path, std::ios::binary | std::ios::ate);
    if (!in) throw std::runtime_error("cannot open fasta file " + path);
    const size_t size = static_cast<size_t>(in.tellg());
    std::vector<char> data(size);
    in.seekg(0);
    if (size && !in.read(data.data(), static_cast<std::streamsize>(size))) throw std::runtime_error("cannot read fasta file " + path);
    // data before the first header (blank lines are skipped, as by the sequential reader)
    size_t first = 0;
    while (first < size && (data[first] == '\n' || data[first] == '\r')) ++first;
    if (first < size && data[first] != '>') throw std::runtime_error("fasta file " + path + " does not start with a '>' header");
    if (threads == 0) threads = 1;
    // piece k starts at the first header at or behind k * size / threads
    std::vector<size_t> cut(threads + 1, size);
    cut[0] = first;
    for (unsigned k = 1; k < threads; ++k) {
        size_t p = std::max(cut[k - 1], size / threads * k);
        while (p < size && !(data[p] == '>' && (p == 0 || data[p - 1] == '\n'))) ++p;
        cut[k] = p;
    }
    struct Piece {
        std::vector<uint8_t> ranks;
        size_t count{0}, len{0};
        bool haveLen{false};
        int problem{0};
        size_t record{0}, pos{0}, length{0};
        char ch{0};
        std::string id;
    };
    std::vector<Piece> pieces(threads);
    auto work = [&](unsigned k) {
        Piece& pc = pieces[k];
        size_t p = cut[k];
        const size_t end = cut[k + 1];
        pc.ranks.reserve(end - p);
        while (p < end) {
            // header line
            size_t e = p;
            while (e < size && data[e] != '\n') ++e;
            size_t idEnd = e;
            if (idEnd > p && data[idEnd - 1] == '\r') --idEnd;
            const size_t idBegin = p + 1;
            p = e < size ? e + 1 : size;
            // sequence lines up to the next header
            const size_t start = pc.ranks.size();
            bool bad = false;
            while (p < size && data[p] != '>') {
                size_t le = p;
                while (le < size && data[le] != '\n') ++le;
                size_t ce = le;
                if (ce > p && data[ce - 1] == '\r') --ce;
                for (size_t i = p; i < ce; ++i) {
                    const uint8_t r = table[static_cast<uint8_t>(data[i])];
                    if (r == 255 && !bad && !pc.problem) {
                        bad = true;
                        pc.problem = 1;
                        pc.record = pc.count;
                        pc.pos = pc.ranks.size() - start;
                        pc.ch = data[i];
                        pc.id.assign(data.data() + idBegin, idEnd > idBegin ? idEnd - idBegin : 0);
                    }
                    pc.ranks.push_back(r);
                }
                p = le < size ? le + 1 : size;
            }
            const size_t len = pc.ranks.size() - start;
            if (!pc.haveLen) {
                pc.haveLen = true;
                pc.len = len;
            } else if (len != pc.len && !pc.problem) {
                pc.problem = 2;
                pc.record = pc.count;
                pc.length = len;
                pc.id.assign(data.data() + idBegin, idEnd > idBegin ? idEnd - idBegin : 0);
            }
            ++pc.count;
        }
    };
    {
        std::vector<std::thread> pool;
        for (unsigned k = 1; k < threads; ++k) pool.emplace_back(work, k);
        work(0);
        for (auto& t : pool) t.join();
    }
    ReadSet rs;
    size_t total = 0;
    for (auto const& pc : pieces) total += pc.ranks.size();
    rs.ranks.reserve(total);
    bool haveLen = false;
    for (auto& pc : pieces) {
        if (pc.count == 0) continue;
        if (!haveLen) {
            haveLen = true;
            rs.len = pc.len;
        }
        // a piece whose records all have one length that differs from the file's first record: its first record is the problem
        if (!rs.problem) {
            if (pc.len != rs.len && (pc.problem != 1 || pc.record != 0)) {
                rs.problem = 2;
                rs.record = rs.count;
                rs.length = pc.len;
                // (the id of that record is recovered below by the caller-independent rule: first record of the piece)
                size_t q = cut[&pc - pieces.data()];
                size_t e = q;
                while (e < size && data[e] != '\n') ++e;
                if (e > q && data[e - 1] == '\r') --e;
                rs.id.assign(data.data() + q + 1, e > q + 1 ? e - q - 1 : 0);
            } else if (pc.problem) {
                rs.problem = pc.problem;
                rs.record = rs.count + pc.record;
                rs.pos = pc.pos;
                rs.length = pc.length;
                rs.ch = pc.ch;
                rs.id = pc.id;
            }
        }
        rs.ranks.insert(rs.ranks.end(), pc.ranks.begin(), pc.ranks.end());
        rs.count += pc.count;
        std::vector<uint8_t>().swap(pc.ranks);
    }
    return rs;
}

struct Writer {
    FILE* f;
    size_t width;
    explicit Writer(std::string const& path, size_t lineWidth = 80) : f(fopen(path.c_str(), "w")), width(lineWidth) {
        if (!f) throw std::runtime_error("cannot open " + path + " for writing");
    }
    ~Writer() { if (f) fclose(f); }
    Writer(Writer const&) = delete;
    void write(std::string const& id, std::string const& seq) {
        fprintf(f, ">%s\n", id.c_str());
        size_t w = width ? width : seq.size();
        for (size_t i = 0; i < seq.size(); i += w) {
            fwrite(seq.data() + i, 1, std::min(w, seq.size() - i), f);
            fputc('\n', f);
        }
    }
};

// ranks (one byte per base) -> 4 bits per base, 8 bases per little-endian word, (len + 7) / 8 words per read; unused
// nibbles of the last word of a read are 0xF.  The format sb200_submit_reads takes as SB200_READS_PACKED4: half the
// bytes on the way to the GPU.  `threads` host threads, each a contiguous range of reads.
// ranks -> 2 bits per base (A, C, G, T = ranks 1 .. 4 -> 0 .. 3), 16 bases per little-endian 32-bit word, (len + 15) / 16 words
// per read (SB200_READS_PACKED2).  Returns the index of the first read that holds another symbol (it cannot be packed),
// or n_reads when all reads were packed.
inline uint64_t packReads2(const uint8_t* ranks, uint64_t n_reads, uint32_t len, unsigned threads, uint32_t* out) {
    const uint32_t W = (len + 15) / 16;
    std::vector<uint64_t> firstBad(std::max(1u, threads), n_reads);
    auto work = [&](unsigned t, uint64_t r0, uint64_t r1) {
        for (uint64_t r = r0; r < r1; ++r) {
            const uint8_t* src = ranks + r * len;
            uint32_t* dst = out + r * W;
            bool bad = false;
            for (uint32_t w = 0; w < W; ++w) {
                uint32_t v = 0;
                const uint32_t n = std::min<uint32_t>(16, len - w * 16);
                for (uint32_t j = 0; j < n; ++j) {
                    const uint32_t c = src[w * 16 + j];
                    bad = bad || c < 1 || c > 4;
                    v |= ((c - 1u) & 3u) << (2 * j);
                }
                dst[w] = v;
            }
            if (bad && firstBad[t] == n_reads) firstBad[t] = r;
        }
    };
    if (threads <= 1 || n_reads < 4096) {
        work(0, 0, n_reads);
    } else {
        std::vector<std::thread> pool;
        const uint64_t per = (n_reads + threads - 1) / threads;
        for (unsigned t = 0; t < threads; ++t) pool.emplace_back(work, t, std::min<uint64_t>(n_reads, per * t), std::min<uint64_t>(n_reads, per * (t + 1)));
        for (auto& th : pool) th.join();
    }
    return *std::min_element(firstBad.begin(), firstBad.end());
}

inline void packReads4(const uint8_t* ranks, uint64_t n_reads, uint32_t len, unsigned threads, uint32_t* out) {
    const uint32_t W = (len + 7) / 8;
    auto work = [&](uint64_t r0, uint64_t r1) {
        for (uint64_t r = r0; r < r1; ++r) {
            const uint8_t* src = ranks + r * len;
            uint32_t* dst = out + r * W;
            uint32_t i = 0;
            for (uint32_t w = 0; w + 1 < W || (w < W && len % 8 == 0); ++w, i += 8) {
                uint64_t x;
                std::memcpy(&x, src + i, 8);
                // low nibbles of 8 bytes -> 32 bits
                x &= 0x0f0f0f0f0f0f0f0full;
                x = (x | (x >> 4)) & 0x00ff00ff00ff00ffull;
                x = (x | (x >> 8)) & 0x0000ffff0000ffffull;
                x = (x | (x >> 16)) & 0x00000000ffffffffull;
                dst[w] = static_cast<uint32_t>(x);
            }
            if (len % 8) {
                uint32_t v = 0xffffffffu;
                for (uint32_t j = 0; i + j < len; ++j) v = (v & ~(0xfu << (4 * j))) | (uint32_t(src[i + j] & 0xf) << (4 * j));
                dst[W - 1] = v;
            }
        }
    };
    if (threads <= 1 || n_reads < 4096) {
        work(0, n_reads);
        return;
    }
    std::vector<std::thread> pool;
    const uint64_t per = (n_reads + threads - 1) / threads;
    for (unsigned t = 0; t < threads; ++t) pool.emplace_back(work, std::min<uint64_t>(n_reads, per * t), std::min<uint64_t>(n_reads, per * (t + 1)));
    for (auto& th : pool) th.join();
}

}  // namespace sahara::fasta
