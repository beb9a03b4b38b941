// host_capi.cpp — C entry points of the host-side helpers declared in include/sahara_host.h.
#include "../../include/sahara_host.h"

#include <cstdlib>
#include <cstring>
#include <string>

#include "alphabet.hpp"
#include "fasta.hpp"
#include "idxfile.hpp"
#include "scheme.hpp"

namespace {
thread_local std::string g_err;

template <typename F>
int guard(F&& f) {
    try {
        f();
        return 0;
    } catch (std::exception const& e) {
        g_err = e.what();
        return 1;
    } catch (...) {
        g_err = "unknown error";
        return 1;
    }
}

using sahara::scheme::Scheme;
using sahara::scheme::Search;

void exportTables(Scheme const& ss, uint32_t* n_searches, uint32_t* n_entries, uint16_t** pi, uint8_t** l, uint8_t** u) {
    size_t S = ss.size(), E = ss.empty() ? 0 : ss[0].pi.size();
    *n_searches = static_cast<uint32_t>(S);
    *n_entries = static_cast<uint32_t>(E);
    *pi = static_cast<uint16_t*>(std::malloc(std::max<size_t>(1, S * E) * 2));
    *l = static_cast<uint8_t*>(std::malloc(std::max<size_t>(1, S * E)));
    *u = static_cast<uint8_t*>(std::malloc(std::max<size_t>(1, S * E)));
    for (size_t j = 0; j < S; ++j) {
        if (ss[j].pi.size() != E) throw std::runtime_error("searches of a scheme differ in length");
        for (size_t i = 0; i < E; ++i) {
            if (ss[j].pi[i] > 0xffff || ss[j].l[i] > 255 || ss[j].u[i] > 255) throw std::runtime_error("scheme entry out of range");
            (*pi)[j * E + i] = static_cast<uint16_t>(ss[j].pi[i]);
            (*l)[j * E + i] = static_cast<uint8_t>(ss[j].l[i]);
            (*u)[j * E + i] = static_cast<uint8_t>(ss[j].u[i]);
        }
    }
}

Scheme importTables(uint32_t S, uint32_t E, const uint16_t* pi, const uint8_t* l, const uint8_t* u) {
    Scheme ss(S);
    for (uint32_t j = 0; j < S; ++j)
        for (uint32_t i = 0; i < E; ++i) {
            ss[j].pi.push_back(pi[size_t(j) * E + i]);
            ss[j].l.push_back(l[size_t(j) * E + i]);
            ss[j].u.push_back(u[size_t(j) * E + i]);
        }
    return ss;
}

Scheme finish(Scheme ss, uint32_t len, int limit) {
    if (len) ss = sahara::scheme::expand(ss, len);
    if (limit) ss = sahara::scheme::limitToHamming(ss);
    return ss;
}
}  // namespace

extern "C" {

const char* sbh_last_error(void) { return g_err.c_str(); }

const char* sbh_scheme_names(void) {
    static std::string names = [] {
        std::string s;
        for (auto const& [k, v] : sahara::scheme::generator::all()) s += (s.empty() ? "" : ",") + k;
        return s;
    }();
    return names.c_str();
}

int sbh_scheme_generate(const char* name, int minK, int maxK, uint32_t len, int limit, uint32_t* n_searches, uint32_t* n_entries,
                        uint16_t** pi, uint8_t** l, uint8_t** u) {
    return guard([&] {
        auto ss = sahara::scheme::generator::generate(name, minK, maxK);
        exportTables(finish(ss, len, limit), n_searches, n_entries, pi, l, u);
    });
}

int sbh_scheme_generate_dynamic(const char* name, int minK, int maxK, uint32_t len, int edit, uint64_t sigma, uint64_t ref_len,
                                uint32_t* n_searches, uint32_t* n_entries, uint16_t** pi, uint8_t** l, uint8_t** u, uint32_t* partition,
                                uint32_t partition_cap, uint32_t* n_parts) {
    return guard([&] {
        auto ss = sahara::scheme::generator::generate(name, minK, maxK);
        auto counts = edit ? sahara::scheme::optimizeByWNCTopDown<true>(ss, len, sigma, ref_len, 1)
                           : sahara::scheme::optimizeByWNCTopDown<false>(ss, len, sigma, ref_len, 1);
        if (n_parts) *n_parts = static_cast<uint32_t>(counts.size());
        for (size_t i = 0; partition && i < counts.size() && i < partition_cap; ++i) partition[i] = static_cast<uint32_t>(counts[i]);
        auto ex = sahara::scheme::expand(ss, counts);
        if (!edit) ex = sahara::scheme::limitToHamming(ex);
        exportTables(ex, n_searches, n_entries, pi, l, u);
    });
}

int sbh_scheme_from_columba(const char* text, uint32_t len, int limit, uint32_t* n_searches, uint32_t* n_entries, uint16_t** pi,
                            uint8_t** l, uint8_t** u) {
    return guard([&] {
        auto ss = sahara::scheme::fromColumba(text);
        exportTables(finish(ss, len, limit), n_searches, n_entries, pi, l, u);
    });
}

int sbh_scheme_check(uint32_t S, uint32_t P, const uint16_t* pi, const uint8_t* l, const uint8_t* u, int minK, int maxK, int* valid,
                     int* complete, int* non_redundant) {
    return guard([&] {
        auto ss = importTables(S, P, pi, l, u);
        *valid = sahara::scheme::isValid(ss);
        *complete = *valid && sahara::scheme::isComplete(ss, minK, maxK);
        *non_redundant = *valid && sahara::scheme::isNonRedundant(ss, minK, maxK);
    });
}

int sbh_scheme_node_count(uint32_t S, uint32_t len, const uint16_t* pi, const uint8_t* l, const uint8_t* u, int edit, uint64_t sigma,
                          uint64_t ref_len, double* nc, double* wnc) {
    return guard([&] {
        auto ss = importTables(S, len, pi, l, u);
        if (edit) {
            *nc = static_cast<double>(sahara::scheme::nodeCount<true>(ss, sigma));
            *wnc = static_cast<double>(sahara::scheme::weightedNodeCount<true>(ss, sigma, ref_len));
        } else {
            *nc = static_cast<double>(sahara::scheme::nodeCount<false>(ss, sigma));
            *wnc = static_cast<double>(sahara::scheme::weightedNodeCount<false>(ss, sigma, ref_len));
        }
    });
}

int sbh_idx_peek_sigma(const char* path, uint64_t* sigma) {
    return guard([&] { *sigma = sahara::peekSigma(path); });
}

int sbh_idx_load(const char* path, sb200_index_view* out, void** handle) {
    return guard([&] {
        auto* im = new sahara::IndexImage(sahara::loadIndexFile(path));
        *out = im->view();
        *handle = im;
    });
}

void sbh_idx_free(void* handle) { delete static_cast<sahara::IndexImage*>(handle); }

int sbh_idx_save(const char* path, const sb200_index_view* view) {
    return guard([&] { sahara::saveIndexFile(path, *view); });
}

int sbh_fasta_load_ranks(const char* path, uint64_t sigma, int with_revcomp, uint8_t** ranks, uint64_t** lens, uint64_t* n_seqs) {
    return guard([&] {
        if (sigma != 5 && sigma != 6) throw std::runtime_error("unknown index with " + std::to_string(sigma) + " letters");
        std::vector<uint8_t> all;
        std::vector<uint64_t> ls;
        size_t count = 0;
        sahara::fasta::read(path, [&](sahara::fasta::Record& rec) {
            ++count;
            auto r = sigma == 5 ? sahara::convert_char_to_rank<sahara::d_dna4>(rec.seq) : sahara::convert_char_to_rank<sahara::d_dna5>(rec.seq);
            if (auto pos = sahara::verify_rank(r); pos) {
                char buf[64];
                snprintf(buf, sizeof buf, "%x", static_cast<unsigned>(static_cast<uint8_t>(rec.seq[*pos])));
                throw std::runtime_error("query '" + rec.id + "' (" + std::to_string(count) + ") has invalid character at position " +
                                         std::to_string(*pos) + " '" + rec.seq[*pos] + "'(" + buf + ")");
            }
            all.insert(all.end(), r.begin(), r.end());
            ls.push_back(r.size());
            if (with_revcomp) {
                auto rc = sahara::reverse_complement_rank<sahara::d_dna5>(r);
                all.insert(all.end(), rc.begin(), rc.end());
                ls.push_back(rc.size());
            }
        });
        *ranks = static_cast<uint8_t*>(std::malloc(std::max<size_t>(1, all.size())));
        *lens = static_cast<uint64_t*>(std::malloc(std::max<size_t>(1, ls.size()) * 8));
        std::memcpy(*ranks, all.data(), all.size());
        std::memcpy(*lens, ls.data(), ls.size() * 8);
        *n_seqs = ls.size();
    });
}

int sbh_revcomp_ranks(const uint8_t* in, uint64_t n, uint8_t* out) {
    return guard([&] {
        for (uint64_t i = 0; i < n; ++i) out[i] = sahara::d_dna5::complement_rank(in[n - 1 - i]);
    });
}

void sbh_free(void* p) { std::free(p); }

int sbh_set_expand_rule(uint32_t rule) {
    return guard([&] {
        if (rule > 1) throw std::runtime_error("unknown expand rule");
        sahara::scheme::expandLowerRule() = rule;
    });
}

int sbh_pack_reads2(const uint8_t* ranks, uint64_t n_reads, uint32_t len, uint32_t threads, uint32_t* out) {
    return guard([&] {
        const uint64_t bad = sahara::fasta::packReads2(ranks, n_reads, len, threads, out);
        if (bad != n_reads)
            throw std::runtime_error("read " + std::to_string(bad) + " holds a symbol other than A, C, G, T: 2 bits per base cannot carry it, "
                                     "send this batch as SB200_READS_PACKED4 or SB200_READS_RANKS");
    });
}

int sbh_pack_reads4(const uint8_t* ranks, uint64_t n_reads, uint32_t len, uint32_t threads, uint32_t* out) {
    return guard([&] { sahara::fasta::packReads4(ranks, n_reads, len, threads, out); });
}

int sbh_decode_records(const uint32_t* hit_end, const uint8_t* records, uint64_t n_queries, uint64_t n_hits, uint32_t record_bytes,
                       uint32_t bits_for_position, int delta_coded, uint64_t first_query, uint64_t* out) {
    return guard([&] {
        if (record_bytes == 0 || record_bytes > 8 || bits_for_position == 0 || bits_for_position > 56) throw std::runtime_error("records: bad format");
        const uint64_t posMask = (uint64_t{1} << bits_for_position) - 1;
        uint64_t h = 0;
        auto put = [&](uint64_t q, uint64_t v) {
            if (h >= n_hits) throw std::runtime_error("records decode to more hits than announced");
            uint64_t* o = out + 4 * h++;
            o[0] = first_query + q;
            o[1] = (v >> 4) >> bits_for_position;
            o[2] = (v >> 4) & posMask;
            o[3] = v & 15;
        };
        for (uint64_t q = 0; q < n_queries; ++q) {
            const uint64_t b = q ? hit_end[q - 1] : 0, e = hit_end[q];
            if (!delta_coded) {
                for (uint64_t i = b; i < e; ++i) {
                    uint64_t v = 0;
                    std::memcpy(&v, records + i * record_bytes, record_bytes);
                    put(q, v);
                }
                continue;
            }
            uint64_t at = b, v = 0;
            while (at < e) {
                if (at == b) {
                    if (at + record_bytes > e) throw std::runtime_error("records: truncated first record of a query");
                    std::memcpy(&v, records + at, record_bytes);
                    at += record_bytes;
                } else {
                    uint64_t d = 0;
                    unsigned shift = 0;
                    while (true) {
                        if (at >= e || shift > 56) throw std::runtime_error("records: truncated difference");
                        const uint8_t byte = records[at++];
                        d |= uint64_t(byte & 0x7f) << shift;
                        shift += 7;
                        if (!(byte & 0x80)) break;
                    }
                    v += d;
                }
                put(q, v);
            }
        }
        if (h != n_hits) throw std::runtime_error("records decode to fewer hits than announced");
    });
}

int sbh_fasta_load_reads(const char* path, uint64_t sigma, uint32_t threads, uint8_t** ranks, uint64_t* n_reads, uint64_t* len) {
    return guard([&] {
        if (sigma != 5 && sigma != 6) throw std::runtime_error("unknown index with " + std::to_string(sigma) + " letters");
        auto rs = sahara::fasta::readRanksParallel(path, sigma == 5 ? sahara::d_dna4::table : sahara::d_dna5::table, threads);
        if (rs.problem == 1) {
            char buf[64];
            snprintf(buf, sizeof buf, "%x", static_cast<unsigned>(static_cast<uint8_t>(rs.ch)));
            throw std::runtime_error("query '" + rs.id + "' (" + std::to_string(rs.record + 1) + ") has invalid character at position " +
                                     std::to_string(rs.pos) + " '" + rs.ch + "'(" + buf + ")");
        }
        if (rs.problem == 2)
            throw std::runtime_error("query '" + rs.id + "' has length " + std::to_string(rs.length) +
                                     ", the search scheme is expanded for the length of the first query (" + std::to_string(rs.len) + ")");
        *ranks = static_cast<uint8_t*>(std::malloc(std::max<size_t>(1, rs.ranks.size())));
        std::memcpy(*ranks, rs.ranks.data(), rs.ranks.size());
        *n_reads = rs.count;
        *len = rs.len;
    });
}

}  // extern "C"
