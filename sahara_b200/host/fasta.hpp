// fasta.hpp — minimal FASTA reader/writer.
//
// Stands in for ivio::fasta::reader as used at /root/reference/src/sahara/search.cpp:115 and
// /root/reference/src/sahara/index.cpp:53 (IVio 1.2.1 is not vendored): a record is a '>' header line
// (id = the line without '>') followed by sequence lines that are concatenated without line breaks.
#pragma once
#include <cstdio>
#include <fstream>
#include <functional>
#include <stdexcept>
#include <string>
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

}  // namespace sahara::fasta
