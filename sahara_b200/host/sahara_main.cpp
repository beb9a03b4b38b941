// sahara_main.cpp — `sahara index` and `sahara search` on the B200 path.
//
// Drop-in for the two sub-commands of the reference that surround the hot path
//   /root/reference/src/sahara/index.cpp:20-121   (sahara index <fasta> [--ignore_unknown] [--dna4])
//   /root/reference/src/sahara/search.cpp:23-291  (sahara search -q Q -i X.idx [-o OUT] [-g GEN] [-e K]
//                                                  [--no-reverse] [-m all] [-d ham|lev] [--limit_queries N])
// Same flags, same index file, same "queryId seqId pos" output lines, same statistics block.  The
// library calls the reference makes into fmindex-collection are replaced by the C ABI of
// include/sahara_b200.h.  Queries are sharded over the visible GPUs (--gpus N,
// index replicated); hit lists are concatenated on the host in query order.
//
// -m besthits follows fmc::search_ng21::search_best (search.cpp:233-240): per query the strata of exactly
// 0, 1, ..., k errors are searched in turn (schemes generator(j, j)) and the first stratum with a hit ends the
// query; as in the reference this mode always uses edit distance.
// --max_hits n follows search_ng24::search_n / search_ng21::search_best_n (search.cpp:228,231,240): a query ends after n
// suffix-array rows, taken in the order of the reference's recursion (sb200_set_max_hits -> fm_ordered_kernel); with
// -m besthits the limit applies inside the first stratum that has a hit.  --dynamic_generator follows a
// reconstruction of optimizeByWNCTopDown (host/scheme.hpp).
#include <algorithm>
#include <chrono>
#include <cinttypes>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <filesystem>
#include <fstream>
#include <iterator>
#include <stdexcept>
#include <string>
#include <memory>
#include <thread>
#include <vector>

#include "../../include/sahara_b200.h"
#include "alphabet.hpp"
#include "fasta.hpp"
#include "idxfile.hpp"
#include "scheme.hpp"

namespace {

struct StopWatch {  // /root/reference/src/sahara/utils/StopWatch.h:8-29
    std::chrono::steady_clock::time_point start{std::chrono::steady_clock::now()};
    double reset() {
        auto now = std::chrono::steady_clock::now();
        double d = std::chrono::duration<double>(now - start).count();
        start = now;
        return d;
    }
};

[[noreturn]] void fail(std::string const& msg) { throw std::runtime_error(msg); }

void check(int rc) {
    if (rc != 0) fail(sb200_last_error());
}

struct Args {
    std::vector<std::string> pos;
    std::vector<std::pair<std::string, std::string>> opts;
    bool has(std::string const& k) const {
        for (auto const& [a, b] : opts)
            if (a == k) return true;
        return false;
    }
    std::string get(std::string const& k, std::string const& def = "") const {
        for (auto const& [a, b] : opts)
            if (a == k) return b;
        return def;
    }
};

// flags that take a value
bool takesValue(std::string const& f) {
    static char const* v[] = {"-q", "--query", "-i", "--index", "-o", "--output", "-g", "--generator", "-e", "--errors", "-m",
                              "--search_mode", "-d", "--distance-metric", "--max_hits", "--limit_queries", "--gpus", "--scheme-file",
                              "--batch", "--device-sa-rate", "--qgram"};
    for (auto s : v)
        if (f == s) return true;
    return false;
}

std::string canonical(std::string const& f) {
    if (f == "-q") return "--query";
    if (f == "-i") return "--index";
    if (f == "-o") return "--output";
    if (f == "-g") return "--generator";
    if (f == "-e") return "--errors";
    if (f == "-m") return "--search_mode";
    if (f == "-d") return "--distance-metric";
    return f;
}

Args parse(int argc, char** argv, int first) {
    Args a;
    for (int i = first; i < argc; ++i) {
        std::string s = argv[i];
        if (s.size() > 1 && s[0] == '-') {
            if (takesValue(s)) {
                if (i + 1 >= argc) fail("option " + s + " needs a value");
                a.opts.emplace_back(canonical(s), argv[++i]);
            } else {
                a.opts.emplace_back(canonical(s), "1");
            }
        } else {
            a.pos.push_back(s);
        }
    }
    return a;
}

// numeric option value; a malformed one names the option instead of surfacing a bare "stoul"
long long numOpt(Args const& a, std::string const& key, std::string const& def) {
    std::string v = a.get(key, def);
    try {
        size_t used = 0;
        long long x = std::stoll(v, &used);
        if (used != v.size()) throw std::invalid_argument(v);
        return x;
    } catch (std::exception const&) {
        fail("option " + key + " expects a number, got \"" + v + "\"");
    }
    return 0;
}
size_t sizeOpt(Args const& a, std::string const& key, std::string const& def) {
    long long x = numOpt(a, key, def);
    if (x < 0) fail("option " + key + " must not be negative");
    return static_cast<size_t>(x);
}

void printTiming(std::vector<std::pair<std::string, double>> const& timing, double* total) {
    printf("stats:\n");
    *total = 0;
    for (auto const& [key, t] : timing) {
        printf("  %-20s %10.2fs\n", (key + " time:").c_str(), t);
        *total += t;
    }
    printf("  total time:          %10.2fs\n", *total);
}

int gpuCount() {
    int n = 0;
    check(sb200_device_count(&n));
    if (n == 0) fail("sahara_b200 needs a CUDA device and has no CPU fallback");
    return n;
}

// ---- sahara index ---------------------------------------------------------------------------------
template <typename Alphabet>
void createIndex(Args const& a) {
    if (a.pos.empty()) fail("usage: sahara index <fasta> [--ignore_unknown] [--dna4]");
    std::string path = a.pos[0];
    bool ignoreUnknown = a.has("--ignore_unknown"), dna4 = a.has("--dna4");
    constexpr size_t Sigma = Alphabet::size();
    printf("constructing an index for %s\n", path.c_str());
    std::vector<std::pair<std::string, double>> timing;
    StopWatch sw;

    std::vector<uint8_t> all;
    std::vector<uint64_t> lens;
    size_t totalSize = 0, count = 0;
    sahara::fasta::read(path, [&](sahara::fasta::Record& rec) {
        ++count;
        totalSize += rec.seq.size();
        auto r = sahara::convert_char_to_rank<Alphabet>(rec.seq);
        if (ignoreUnknown) {
            for (auto& v : r) {
                if (sahara::verify_rank(v)) continue;
                v = dna4 ? static_cast<uint8_t>(Alphabet::char_to_rank('A') + rand() % 4) : Alphabet::char_to_rank('N');
            }
        }
        if (auto pos = sahara::verify_rank(r); pos) {
            char buf[256];
            snprintf(buf, sizeof buf, "ref '%s' (%zu) has invalid character '%c' (0x%02x) at position %zu", rec.id.c_str(), count,
                     rec.seq[*pos], static_cast<unsigned>(static_cast<uint8_t>(rec.seq[*pos])), *pos);
            fail(buf);
        }
        for (auto v : r)
            if (v == 0) fail("ref '" + rec.id + "' contains the delimiter character '$'");
        all.insert(all.end(), r.begin(), r.end());
        lens.push_back(r.size());
    });
    if (lens.empty()) fail("reference file " + path + " was empty - abort\n");
    printf("config:\n  file: %s\n  sigma: %zu\n  references: %zu\n  totalSize: %zu\n", path.c_str(), Sigma, lens.size(), totalSize);
    timing.emplace_back("ld queries", sw.reset());

    gpuCount();
    sb200_ctx* ctx = nullptr;
    check(sb200_create(0, &ctx));
    check(sb200_index_build(ctx, all.data(), lens.data(), lens.size(), static_cast<uint32_t>(Sigma), /*samplingRate*/ 16));
    timing.emplace_back("index creation", sw.reset());

    std::string indexPath = path + (dna4 ? ".dna4.idx" : ".idx");
    sb200_index_view view{};
    check(sb200_index_download(ctx, &view));
    sahara::saveIndexFile(indexPath, view);
    sb200_index_view_free(&view);
    sb200_destroy(ctx);
    timing.emplace_back("saving to disk", sw.reset());
    double total;
    printTiming(timing, &total);
}

// ---- sahara search --------------------------------------------------------------------------------
template <typename Alphabet>
void runSearch(Args const& a) {
    namespace ss = sahara::scheme;
    constexpr size_t Sigma = Alphabet::size();
    std::string queryPath = a.get("--query"), indexPath = a.get("--index"), outPath = a.get("--output", "sahara-output.txt");
    std::string generator = a.get("--generator", "h2-k2"), mode = a.get("--search_mode", "all"), metric = a.get("--distance-metric", "lev");
    size_t k = sizeOpt(a, "--errors", "0");
    bool noReverse = a.has("--no-reverse");
    long maxHits = static_cast<long>(numOpt(a, "--max_hits", "0"));
    size_t limitQueries = sizeOpt(a, "--limit_queries", "0");
    if (mode != "all" && mode != "besthits") fail("unknown search mode \"" + mode + "\"");
    if (metric != "ham" && metric != "lev") fail("unknown distance metric \"" + metric + "\"");
    const bool bestHits = mode == "besthits";
    if (maxHits < 0) fail("--max_hits must not be negative");
    const bool dynGenerator = a.has("--dynamic_generator");
    bool edit = metric == "lev" || bestHits;  // search_best has no Hamming variant (search.cpp:239)

    std::vector<std::pair<std::string, double>> timing;
    StopWatch sw;

    // load fasta file (search.cpp:115-124): the reads as ranks; queries[2i] = read i, queries[2i+1] = its reverse
    // complement — the second strand is made on the device (sb200_search_reads), or here for -m besthits
    const unsigned hostThreads = std::max(1u, std::min(16u, std::thread::hardware_concurrency()));
    const size_t per = noReverse ? 1 : 2;  // queries per read
    sahara::fasta::ReadSet rs = sahara::fasta::readRanksParallel(queryPath, Alphabet::table, hostThreads);
    if (rs.problem == 1) {
        char buf[512];
        snprintf(buf, sizeof buf, "query '%s' (%zu) has invalid character at position %zu '%c'(%x)", rs.id.c_str(), rs.record * per + 1, rs.pos,
                 rs.ch, static_cast<unsigned>(static_cast<uint8_t>(rs.ch)));
        fail(buf);
    }
    if (rs.problem == 2)
        fail("query '" + rs.id + "' has length " + std::to_string(rs.length) + ", the search scheme is expanded for the length of the first query (" +
             std::to_string(rs.len) + ")");
    std::vector<uint8_t>& reads = rs.ranks;
    size_t qlen = rs.len;
    size_t nQueries = rs.count * per;
    if (limitQueries) nQueries = std::min(limitQueries, nQueries);
    const size_t nReads = (nQueries + per - 1) / per;  // (an odd limit keeps the forward strand of the last read only)
    if (nQueries == 0) fail("query file " + queryPath + " was empty - abort\n");
    timing.emplace_back("ld queries", sw.reset());

    printf("config:\n  query:               %s\n  index:               %s\n  generator:           %s\n  dynamic expansion:   %s\n"
           "  allowed errors:      %zu\n  reverse complements: %s\n  search mode:         %s\n  max hits:            %ld\n"
           "  output path:         %s\n",
           queryPath.c_str(), indexPath.c_str(), generator.c_str(), dynGenerator ? "true" : "false", k, noReverse ? "false" : "true", mode.c_str(), maxHits,
           outPath.c_str());
    {
        size_t fwd = nQueries / (noReverse ? 1 : 2);
        printf("fwd queries: %zu\nbwd queries: %zu\n", fwd, nQueries - fwd);
    }
    if (!std::filesystem::exists(indexPath)) fail("no valid index path at " + indexPath);

    int nGpus = gpuCount();
    if (a.has("--gpus")) nGpus = std::min(nGpus, std::max(1, static_cast<int>(numOpt(a, "--gpus", "1"))));
    nGpus = static_cast<int>(std::min<size_t>(nGpus, (nQueries + 1) / 2));
    auto image = sahara::loadIndexFile(indexPath);
    auto view = image.view();
    std::vector<sb200_ctx*> ctxs(nGpus, nullptr);
    // GPU 0 gets the index from the file and derives its tables; the other GPUs copy everything from a GPU that already
    // has it — round r: GPUs 2^r .. 2^(r+1)-1 from GPU g - 2^r, all copies of a round at once (NVLink, sb200_index_clone)
    {
        check(sb200_create(0, &ctxs[0]));
        check(sb200_index_upload(ctxs[0], &view));
        if (a.has("--device-sa-rate")) check(sb200_index_densify(ctxs[0], static_cast<uint32_t>(sizeOpt(a, "--device-sa-rate", "16"))));
        // in-text verification (17 more bytes per row on the device) and the q-gram jump table are on by default
        if (!a.has("--no-text")) check(sb200_index_enable_text(ctxs[0], 1));
        unsigned q = 0;
        for (uint64_t n = image.n_rows; n >= 4 && q < 15; n /= 4) ++q;  // floor(log4(rows))
        q = std::min(15u, q);  // the depth at which cursors become unique; 4^15 x 16 B = 17 GB at most (3.1 Gbp genomes)
        if (a.has("--qgram")) q = static_cast<unsigned>(sizeOpt(a, "--qgram", "0"));
        check(sb200_index_build_qgram(ctxs[0], q));
    }
    for (int have = 1; have < nGpus; have *= 2) {
        std::vector<std::thread> setup;
        std::vector<std::string> setupErrors(nGpus);
        for (int g = have; g < std::min(2 * have, nGpus); ++g)
            setup.emplace_back([&, g, have] {
                try {
                    check(sb200_create(g, &ctxs[g]));
                    check(sb200_index_clone(ctxs[g], ctxs[g - have]));
                } catch (std::exception const& e) {
                    setupErrors[g] = e.what();
                }
            });
        for (auto& t : setup) t.join();
        for (auto const& e : setupErrors)
            if (!e.empty()) fail(e);
    }
    for (int g = 0; g < nGpus; ++g) check(sb200_set_max_hits(ctxs[g], static_cast<uint64_t>(maxHits)));  // search_n / search_best_n (search.cpp:228,231,240)
    timing.emplace_back("ld index", sw.reset());

    // search scheme (search.cpp:174-212, 226); besthits: one scheme per stratum of exactly j errors (search.cpp:234-237)
    struct Tables {
        uint32_t nSearches;
        std::vector<uint16_t> pi;
        std::vector<uint8_t> lo, up;
    };
    auto loadSearchScheme = [&](int minK, int maxK) {
        ss::Scheme scheme;
        if (a.has("--scheme-file")) {
            std::ifstream in(a.get("--scheme-file"));
            if (!in) fail("cannot open scheme file " + a.get("--scheme-file"));
            std::string text((std::istreambuf_iterator<char>(in)), std::istreambuf_iterator<char>());
            scheme = ss::fromColumba(text);
            if (!ss::isComplete(scheme, minK, maxK))
                fail("the scheme in " + a.get("--scheme-file") + " is not complete for " + std::to_string(minK) + ".." + std::to_string(maxK) + " errors");
        } else {
            scheme = ss::generator::generate(generator, minK, maxK);
        }
        if (!dynGenerator) {
            scheme = ss::expand(scheme, qlen);
        } else {  // search.cpp:193-195 / 203-205: part sizes by weighted node count (Sigma, index.size(), steps = 1)
            auto partition = edit ? ss::optimizeByWNCTopDown<true>(scheme, qlen, Sigma, image.n_rows, 1)
                                  : ss::optimizeByWNCTopDown<false>(scheme, qlen, Sigma, image.n_rows, 1);
            std::string txt;
            for (size_t i = 0; i < partition.size(); ++i) txt += (i ? ", " : "") + std::to_string(partition[i]);
            printf("partition: [%s]\n", txt.c_str());
            scheme = ss::expand(scheme, partition);
        }
        if (edit) {
            printf("node count: %.0Lf\n", ss::nodeCount<true>(scheme, Sigma));
            printf("weighted node count: %.2Lf\n", ss::weightedNodeCount<true>(scheme, Sigma, image.n_rows));
        } else {
            printf("node count: %.0Lf\n", ss::nodeCount<false>(scheme, Sigma));
            printf("weighted node count: %.2Lf\n", ss::weightedNodeCount<false>(scheme, Sigma, image.n_rows));
            scheme = ss::limitToHamming(scheme);
        }
        Tables t;
        t.nSearches = static_cast<uint32_t>(scheme.size());
        for (auto const& s : scheme)
            for (size_t i = 0; i < s.pi.size(); ++i) {
                t.pi.push_back(static_cast<uint16_t>(s.pi[i]));
                t.lo.push_back(static_cast<uint8_t>(s.l[i]));
                t.up.push_back(static_cast<uint8_t>(s.u[i]));
            }
        return t;
    };
    std::vector<Tables> schemes;
    if (!bestHits) schemes.push_back(loadSearchScheme(0, static_cast<int>(k)));
    else
        for (size_t j = 0; j <= k; ++j) schemes.push_back(loadSearchScheme(static_cast<int>(j), static_cast<int>(j)));
    auto setScheme = [&](sb200_ctx* c, Tables const& t) {
        check(sb200_set_scheme(c, t.nSearches, static_cast<uint32_t>(qlen), t.pi.data(), t.lo.data(), t.up.data(), edit));
    };
    timing.emplace_back("searchScheme", sw.reset());

    // search + locate: contiguous shards of reads per GPU, batches inside a shard.  Hits stay in the buffers the
    // library hands out (16-byte records, query ids local to the call) until they are written.
    struct HitBlock {
        sb200_hit32* hits;
        uint64_t n;
        uint64_t firstQuery;  // added to the query ids of the block
        bool owned;           // allocated here (besthits), not by the library
    };
    size_t batch = sizeOpt(a, "--batch", "2000000");
    batch += batch & 1;  // keep both strands of a read together
    const size_t batchReads = std::max<size_t>(1, batch / per);
    std::vector<std::vector<HitBlock>> results(nGpus);
    // the plain search formats the hit lines of a batch while the next batches run: text per GPU, in batch order
    struct TextBuf {  // (uninitialised storage: a std::vector would zero 33 bytes per hit first)
        std::unique_ptr<char[]> p;
        size_t n{0};
    };
    std::vector<std::vector<TextBuf>> text(nGpus);
    std::vector<size_t> textHits(nGpus, 0);
    const unsigned fmtThreads = std::max(1u, hostThreads / static_cast<unsigned>(nGpus));
    std::vector<std::string> errors(nGpus);
    std::vector<double> msSearch(nGpus, 0), msLocate(nGpus, 0);
    const size_t readsPerGpu = (nReads + nGpus - 1) / nGpus;
    // -m besthits works on single queries: both strands materialised on the host
    std::vector<uint8_t> queries;
    if (bestHits) {
        queries.resize(nQueries * qlen);
        for (size_t q = 0; q < nQueries; ++q) {
            const uint8_t* r = reads.data() + (q / per) * qlen;
            uint8_t* dst = queries.data() + q * qlen;
            if (per == 2 && (q & 1)) {
                for (size_t i = 0; i < qlen; ++i) dst[i] = Alphabet::complement_rank(r[qlen - 1 - i]);
            } else {
                std::memcpy(dst, r, qlen);
            }
        }
    }
    std::vector<std::thread> threads;
    for (int g = 0; g < nGpus; ++g) {
        threads.emplace_back([&, g] {
            const size_t r0 = std::min(nReads, readsPerGpu * g), r1 = std::min(nReads, readsPerGpu * (g + 1));
            try {
                auto account = [&] {
                    sb200_counters ct{};
                    sb200_get_counters(ctxs[g], &ct);
                    msSearch[g] += ct.ms_search;
                    msLocate[g] += ct.ms_locate + ct.ms_sort;
                };
                if (!bestHits && maxHits > 0) {
                    // search_n stops queries at a row limit and walks them again in recursion order: the synchronous call
                    setScheme(ctxs[g], schemes[0]);
                    for (size_t b = r0; b < r1; b += batchReads) {
                        const size_t n = std::min(batchReads, r1 - b);
                        sb200_hit32* hits = nullptr;
                        uint64_t nHits = 0;
                        check(sb200_search_reads(ctxs[g], reads.data() + b * qlen, n, static_cast<uint32_t>(qlen), noReverse ? 0 : 1, &hits,
                                                 &nHits));
                        results[g].push_back(HitBlock{hits, nHits, b * per, false});
                        account();
                    }
                } else if (!bestHits) {
                    // sb200_submit_reads / sb200_wait_batch: up to two batches in flight, so that the reads of the next batch
                    // travel to the GPU and the hits of the previous one come back (5-byte CSR records) while a batch computes;
                    // the hit lines "{queryId} {seqId} {pos}\n" (search.cpp:257-259) of a finished batch are formatted right
                    // here, by this GPU's share of the host threads, while the GPU works on the following batches
                    setScheme(ctxs[g], schemes[0]);
                    struct InFlight { uint64_t ticket, firstQuery; };
                    std::vector<InFlight> flight;
                    auto putNum = [](char*& out, uint64_t v, char sep) {
                        char tmp[24];
                        int n = 0;
                        do { tmp[n++] = static_cast<char>('0' + v % 10); v /= 10; } while (v);
                        while (n) *out++ = tmp[--n];
                        *out++ = sep;
                    };
                    auto collect = [&](InFlight const& fl) {
                        sb200_batch_result res{};
                        check(sb200_wait_batch(ctxs[g], fl.ticket, 1, &res));
                        msSearch[g] += res.ms_search;
                        msLocate[g] += res.ms_locate + res.ms_sort;
                        // slices of queries with about the same amount of records (hits, or bytes when delta coded), one per
                        // formatting thread
                        const uint64_t units = res.n_queries ? res.hit_end[res.n_queries - 1] : 0;
                        std::vector<uint64_t> cut(fmtThreads + 1, res.n_queries);
                        cut[0] = 0;
                        for (unsigned t = 1; t < fmtThreads; ++t)
                            cut[t] = static_cast<uint64_t>(std::lower_bound(res.hit_end, res.hit_end + res.n_queries,
                                                                             static_cast<uint32_t>(units * t / fmtThreads)) - res.hit_end);
                        const size_t firstBuf = text[g].size();
                        text[g].resize(firstBuf + fmtThreads);
                        std::vector<size_t> kept(fmtThreads, 0);
                        auto format = [&](unsigned t) {
                            const uint64_t qa = cut[t], qb = std::max(cut[t], cut[t + 1]);
                            const uint64_t u0 = qa ? res.hit_end[qa - 1] : 0, u1 = qb ? res.hit_end[qb - 1] : 0;
                            auto& buf = text[g][firstBuf + t];
                            // 3 numbers of at most 10 digits + separators per hit; a delta-coded hit takes at least one byte
                            buf.p.reset(new char[(u1 - u0) * 33 + 1]);
                            char* out = buf.p.get();
                            const uint64_t posMask = (uint64_t{1} << res.bits_for_position) - 1;
                            auto line = [&](uint64_t query, uint64_t v) {
                                v >>= 4;  // (the low 4 bits are the errors)
                                putNum(out, query, ' ');
                                putNum(out, v >> res.bits_for_position, ' ');
                                putNum(out, v & posMask, '\n');
                                ++kept[t];
                            };
                            uint64_t at = u0;
                            for (uint64_t q = qa; q < qb; ++q) {
                                const uint64_t end = res.hit_end[q], query = fl.firstQuery + q;
                                const bool skip = query >= nQueries;  // (the reverse strand of the last read when --limit_queries is odd)
                                if (!res.delta_coded) {
                                    for (; at < end; ++at) {
                                        uint64_t v = 0;
                                        std::memcpy(&v, res.records + at * res.record_bytes, res.record_bytes);
                                        if (!skip) line(query, v);
                                    }
                                    continue;
                                }
                                uint64_t v = 0;
                                for (bool first = true; at < end; first = false) {
                                    if (first) {
                                        std::memcpy(&v, res.records + at, res.record_bytes);
                                        at += res.record_bytes;
                                    } else {  // difference to the previous record: 7 bits per byte, low bits first
                                        uint64_t d = 0;
                                        for (unsigned shift = 0;; shift += 7) {
                                            const uint8_t byte = res.records[at++];
                                            d |= uint64_t(byte & 0x7f) << shift;
                                            if (!(byte & 0x80)) break;
                                        }
                                        v += d;
                                    }
                                    if (!skip) line(query, v);
                                }
                            }
                            buf.n = static_cast<size_t>(out - buf.p.get());
                        };
                        std::vector<std::thread> pool;
                        for (unsigned t = 1; t < fmtThreads; ++t) pool.emplace_back(format, t);
                        format(0);
                        for (auto& th : pool) th.join();
                        for (size_t kk : kept) textHits[g] += kk;
                        check(sb200_release_batch(ctxs[g], fl.ticket));
                    };
                    size_t done = 0;
                    for (size_t b = r0; b < r1; b += batchReads) {
                        const size_t n = std::min(batchReads, r1 - b);
                        uint64_t ticket = 0;
                        check(sb200_submit_reads(ctxs[g], reads.data() + b * qlen, n, static_cast<uint32_t>(qlen), SB200_READS_RANKS, noReverse ? 0 : 1,
                                                 &ticket));
                        flight.push_back(InFlight{ticket, b * per});
                        if (flight.size() - done >= 2) collect(flight[done++]);
                    }
                    while (done < flight.size()) collect(flight[done++]);
                } else {
                    // strata of exactly j errors; queries that found a hit leave the pool
                    const size_t q0 = std::min(nQueries, r0 * per), q1 = std::min(nQueries, r1 * per);
                    std::vector<sb200_hit> found_hits;
                    std::vector<uint64_t> active(q1 - q0);
                    for (size_t i = 0; i < active.size(); ++i) active[i] = q0 + i;
                    std::vector<uint8_t> dense;
                    for (size_t j = 0; j < schemes.size() && !active.empty(); ++j) {
                        setScheme(ctxs[g], schemes[j]);
                        std::vector<uint8_t> found(active.size(), 0);
                        for (size_t b = 0; b < active.size(); b += batch) {
                            size_t n = std::min(batch, active.size() - b);
                            dense.resize(n * qlen);
                            for (size_t i = 0; i < n; ++i) std::memcpy(dense.data() + i * qlen, queries.data() + active[b + i] * qlen, qlen);
                            sb200_hit* hits = nullptr;
                            uint64_t nHits = 0;
                            check(sb200_search(ctxs[g], dense.data(), n, static_cast<uint32_t>(qlen), &hits, &nHits));
                            size_t old = found_hits.size();
                            found_hits.resize(old + nHits);
                            for (uint64_t i = 0; i < nHits; ++i) {
                                found[b + hits[i].query_id] = 1;
                                found_hits[old + i] = hits[i];
                                found_hits[old + i].query_id = active[b + hits[i].query_id];
                            }
                            sb200_free(hits);
                            account();
                        }
                        std::vector<uint64_t> rest;
                        for (size_t i = 0; i < active.size(); ++i)
                            if (!found[i]) rest.push_back(active[i]);
                        active.swap(rest);
                    }
                    std::sort(found_hits.begin(), found_hits.end(), [](sb200_hit const& x, sb200_hit const& y) {
                        if (x.query_id != y.query_id) return x.query_id < y.query_id;
                        if (x.seq_id != y.seq_id) return x.seq_id < y.seq_id;
                        if (x.pos != y.pos) return x.pos < y.pos;
                        return x.errors < y.errors;
                    });
                    auto* block = new sb200_hit32[std::max<size_t>(1, found_hits.size())];
                    for (size_t i = 0; i < found_hits.size(); ++i)
                        block[i] = sb200_hit32{static_cast<uint32_t>(found_hits[i].query_id), static_cast<uint32_t>(found_hits[i].seq_id),
                                               static_cast<uint32_t>(found_hits[i].pos), static_cast<uint32_t>(found_hits[i].errors)};
                    results[g].push_back(HitBlock{block, found_hits.size(), 0, true});
                }
            } catch (std::exception const& ex) {
                errors[g] = ex.what();
            }
        });
    }
    for (auto& t : threads) t.join();
    for (auto const& e : errors)
        if (!e.empty()) fail(e);
    double total = sw.reset();
    double sMax = 0, lMax = 0;
    for (int g = 0; g < nGpus; ++g) {
        sMax = std::max(sMax, msSearch[g] * 1e-3);
        lMax = std::max(lMax, msLocate[g] * 1e-3);
    }
    // split the wall time of the fused phase in the proportion of the device timers
    double frac = (sMax + lMax) > 0 ? sMax / (sMax + lMax) : 1.0;
    timing.emplace_back("search", total * frac);
    timing.emplace_back("locate", total * (1 - frac));

    size_t nHitsTotal = 0;
    {
        FILE* ofs = fopen(outPath.c_str(), "w");
        if (!ofs) fail("cannot open " + outPath + " for writing");
        // "{queryId} {seqId} {pos}\n" (search.cpp:257-259), formatted by hand: every host thread formats its slice of a
        // block of hits into its own buffer, the buffers are written in order
        constexpr size_t kSlice = 1 << 18;  // hits per thread and round
        std::vector<std::vector<char>> bufs(hostThreads);
        std::vector<size_t> kept(hostThreads, 0);
        auto format = [&](HitBlock const& hb, size_t from, size_t to, std::vector<char>& buf, size_t& nKept) {
            buf.resize((to - from) * 33 + 1);  // 3 numbers of at most 10 digits + separators
            char* out = buf.data();
            nKept = 0;
            auto putNum = [&](uint64_t v, char sep) {
                char tmp[24];
                int n = 0;
                do { tmp[n++] = static_cast<char>('0' + v % 10); v /= 10; } while (v);
                while (n) *out++ = tmp[--n];
                *out++ = sep;
            };
            for (size_t i = from; i < to; ++i) {
                const uint64_t q = hb.firstQuery + hb.hits[i].query_id;
                if (q >= nQueries) continue;  // (the reverse strand of the last read when --limit_queries is odd)
                putNum(q, ' ');
                putNum(hb.hits[i].seq_id, ' ');
                putNum(hb.hits[i].pos, '\n');
                ++nKept;
            }
            buf.resize(static_cast<size_t>(out - buf.data()));
        };
        for (int g = 0; g < nGpus; ++g) {
            for (auto const& buf : text[g]) fwrite(buf.p.get(), 1, buf.n, ofs);
            nHitsTotal += textHits[g];
        }
        for (int g = 0; g < nGpus; ++g)
            for (auto const& hb : results[g]) {
                for (size_t base = 0; base < hb.n; base += kSlice * hostThreads) {
                    std::vector<std::thread> pool;
                    unsigned used = 0;
                    for (unsigned t = 0; t < hostThreads; ++t) {
                        const size_t from = base + t * kSlice, to = std::min<size_t>(hb.n, from + kSlice);
                        if (from >= to) break;
                        ++used;
                        if (t == 0) continue;  // slice 0 is formatted by this thread
                        pool.emplace_back(format, std::cref(hb), from, to, std::ref(bufs[t]), std::ref(kept[t]));
                    }
                    if (used) format(hb, base, std::min<size_t>(hb.n, base + kSlice), bufs[0], kept[0]);
                    for (auto& th : pool) th.join();
                    for (unsigned t = 0; t < used; ++t) {
                        fwrite(bufs[t].data(), 1, bufs[t].size(), ofs);
                        nHitsTotal += kept[t];
                    }
                }
                if (hb.owned) delete[] hb.hits;
                else sb200_free(hb.hits);
            }
        fclose(ofs);
    }
    timing.emplace_back("result", sw.reset());
    double totalTime;
    printTiming(timing, &totalTime);
    printf("  queries per second:  %10.0fq/s\n", nQueries / totalTime);
    printf("  number of hits:      %10zu\n", nHitsTotal);
    printf("  gpus:                %10d\n", nGpus);
    for (auto* c : ctxs) sb200_destroy(c);
}

void usage() {
    printf("sahara (B200 path)\n"
           "  sahara index <fasta> [--ignore_unknown] [--dna4]\n"
           "  sahara search -q <fasta> -i <index> [-o <out>] [-g <generator>] [-e <errors>] [--no-reverse]\n"
           "                [-m all] [-d ham|lev] [--limit_queries <n>] [--gpus <n>] [--scheme-file <columba.txt>]\n"
           "                [--batch <queries per call>] [--device-sa-rate <16|8|4|2|1>] [--qgram <q>] [--no-text]\n");
}

}  // namespace

int main(int argc, char** argv) {
    try {
        if (argc < 2 || std::string(argv[1]) == "--help" || std::string(argv[1]) == "-h") {
            usage();
            return argc < 2 ? 1 : 0;
        }
        std::string cmd = argv[1];
        Args a = parse(argc, argv, 2);
        if (cmd == "index") {
            if (a.has("--dna4")) createIndex<sahara::d_dna4>(a);
            else createIndex<sahara::d_dna5>(a);
        } else if (cmd == "search") {
            if (!a.has("--query") || !a.has("--index")) fail("sahara search needs --query and --index");
            uint64_t sigma = sahara::peekSigma(a.get("--index"));  // search.cpp:278-290
            if (sigma == 5) runSearch<sahara::d_dna4>(a);
            else if (sigma == 6) runSearch<sahara::d_dna5>(a);
            else fail("unknown index with " + std::to_string(sigma) + " letters");
        } else {
            fail("unknown command \"" + cmd + "\"");
        }
    } catch (std::exception const& e) {
        fprintf(stderr, "%s\n", e.what());  // clice prints the message and exits 1 (main.cpp:13)
        return 1;
    }
    return 0;
}
