// sahara_oracle.cpp — CPU ORACLE.  TEST INFRASTRUCTURE ONLY.
//
// This file restates, on the CPU, the algorithm that `sahara search` runs through the third-party
// library fmindex-collection (pinned 1.1.0 in /root/reference/cpm.dependencies:20-24).  It is the
// checker for the CUDA path in sahara_b200/csrc and the CPU baseline that bench.py times beside it.
// Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load it.
// The product (sahara_b200/) never includes, links or calls anything in this directory.
//
// PARITY UNPINNED: fmindex-collection is not vendored in /root/reference and cannot be fetched (no
// network), the reference ships no tests, fixtures or golden vectors.  What is
// restated here follows (1) the reference's own call sites in src/sahara/search.cpp and
// src/sahara/index.cpp, and (2) the published/recalled semantics of the library documented in
// SURVEY.md §9.  The Hamming hit set is mathematically defined and is cross-checked against a brute
// force scan (tests/); the edit-distance hit set is checked for soundness/completeness against DP.
//
// Reference call sites restated (paths relative to /root/reference/):
//   index type  fmc::BiFMIndex<Sigma, fmc::string::InterleavedBitvector16>   src/sahara/search.cpp:162, src/sahara/index.cpp:87
//   on-disk     archive(Sigma); archive(index)                               src/sahara/index.cpp:96-100, src/sahara/search.cpp:164-168
//   search      fmc::search_ng24::search<Edit>(index, queries, scheme, cb)    src/sahara/search.cpp:227-231
//   locate      for (auto [sae, offset] : fmc::LocateLinear{index, cursor})   src/sahara/search.cpp:245-249
//
// Build: see oracle/Makefile  (g++ -O3 -std=c++20 -fopenmp -shared -fPIC)
#include <algorithm>
#include <array>
#include <atomic>
#include <bit>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <memory>
#include <stdexcept>
#include <string>
#include <vector>
#if defined(_OPENMP)
#include <omp.h>
#include <parallel/algorithm>
#endif

// the ONE table of reconstructed rules of the recursion, shared with the CUDA kernels (a header of data, no code of the
// product): which operation may follow which, what may be reported, the order of the children
#include "../include/sahara_policy.h"

namespace {

sb200_policy g_policy = SB200_POLICY_DEFAULT;  // orc_set_policy()


using u8 = uint8_t;
using u16 = uint16_t;
using u32 = uint32_t;
using u64 = uint64_t;

thread_local std::string g_error;

// ---------------------------------------------------------------------------------------------
// Occurrence table in the InterleavedBitvector16 layout:
// one block per 64 BWT rows = Sigma u16 in-superblock counters + Sigma one-hot u64 bitplanes,
// 64-byte aligned in memory; one superblock row (Sigma u64) every 65536 rows.
// ---------------------------------------------------------------------------------------------
template <int S>
struct alignas(64) OccBlock {
    u16 cnt[S];
    u64 bits[S];
};
static_assert(sizeof(OccBlock<5>) == 64 && sizeof(OccBlock<6>) == 64);

template <int S>
struct OccTable {
    std::vector<OccBlock<S>> blocks;
    std::vector<std::array<u64, S>> superBlocks;
    u64 n{};  // number of rows

    void build(u8 const* bwt, u64 len) {
        n = len;
        u64 nBlocks = len / 64 + 1;  // rank(len) must be answerable
        blocks.assign(nBlocks, OccBlock<S>{});
        superBlocks.assign((nBlocks + 1023) / 1024, std::array<u64, S>{});
        std::array<u64, S> total{};
        std::array<u64, S> inSuper{};
        for (u64 b = 0; b < nBlocks; ++b) {
            if (b % 1024 == 0) {
                superBlocks[b / 1024] = total;
                inSuper.fill(0);
            }
            auto& blk = blocks[b];
            for (int s = 0; s < S; ++s) blk.cnt[s] = static_cast<u16>(inSuper[s]);
            u64 end = std::min<u64>(len, (b + 1) * 64);
            for (u64 r = b * 64; r < end; ++r) {
                u8 c = bwt[r];
                blk.bits[c] |= u64{1} << (r & 63);
                ++inSuper[c];
                ++total[c];
            }
        }
    }
    inline u64 rank(u64 i, int c) const {
        auto const& blk = blocks[i >> 6];
        u64 mask = (u64{1} << (i & 63)) - 1;
        return superBlocks[i >> 16][c] + blk.cnt[c] + std::popcount(blk.bits[c] & mask);
    }
    inline void all_ranks(u64 i, u64* out) const {
        auto const& blk = blocks[i >> 6];
        auto const& sb = superBlocks[i >> 16];
        u64 mask = (u64{1} << (i & 63)) - 1;
        for (int c = 0; c < S; ++c) out[c] = sb[c] + blk.cnt[c] + std::popcount(blk.bits[c] & mask);
    }
    inline int symbol(u64 i) const {
        auto const& blk = blocks[i >> 6];
        u64 bit = u64{1} << (i & 63);
        for (int c = 0; c < S; ++c)
            if (blk.bits[c] & bit) return c;
        return 0;
    }
};

// Bitvector with rank support marking the sampled suffix-array rows.
struct MarkVector {
    std::vector<u64> bits;
    std::vector<u64> before;  // number of set bits before word w
    u64 n{};
    void finalize() {
        before.assign(bits.size() + 1, 0);
        for (size_t w = 0; w < bits.size(); ++w) before[w + 1] = before[w] + std::popcount(bits[w]);
    }
    inline bool test(u64 i) const { return (bits[i >> 6] >> (i & 63)) & 1; }
    inline u64 rank(u64 i) const { return before[i >> 6] + std::popcount(bits[i >> 6] & ((u64{1} << (i & 63)) - 1)); }
    u64 ones() const { return before.back(); }
};

struct IndexBase {
    int sigma{};
    virtual ~IndexBase() = default;
};

template <int S>
struct Index : IndexBase {
    OccTable<S> bwt, bwtRev;
    std::array<u64, S + 1> C{};
    std::vector<u64> ssa;  // (seqId << bitsForPosition) | seqPos for every marked row, in row order
    MarkVector marks;
    u64 samplingRate{16};
    u64 bitsForPosition{};
    u64 size() const { return bwt.n; }
};

// ---------------------------------------------------------------------------------------------
// Index construction.  Text = every sequence followed by one delimiter (rank 0).
// Suffix order = plain lexicographic order with "end of text" smaller than every symbol.
// ---------------------------------------------------------------------------------------------
static std::vector<u64> suffix_array(std::vector<u8> const& text) {
    u64 n = text.size();
    constexpr int K = 21;  // symbols per 64-bit key, 3 bits each, code = rank+1, 0 = past the end
    std::vector<u64> key(n + K + 1, 0);
    {
        u64 k = 0;
        for (u64 i = n; i-- > 0;) {
            k = (k >> 3) | (u64(text[i] + 1) << (3 * (K - 1)));
            key[i] = k;
        }
    }
    std::vector<u64> sa(n);
    for (u64 i = 0; i < n; ++i) sa[i] = i;
    auto less = [&](u64 a, u64 b) {
        while (true) {
            u64 ka = a < n ? key[a] : 0, kb = b < n ? key[b] : 0;
            if (ka != kb) return ka < kb;
            if (a >= n || b >= n) return a > b;  // unreachable for distinct suffixes; keeps strict weak order
            a += K;
            b += K;
        }
    };
#if defined(_OPENMP)
    __gnu_parallel::sort(sa.begin(), sa.end(), less);
#else
    std::sort(sa.begin(), sa.end(), less);
#endif
    return sa;
}

template <int S>
static std::unique_ptr<Index<S>> build_index(u8 const* seqs, u64 const* lens, u64 nSeqs, u64 samplingRate) {
    auto idx = std::make_unique<Index<S>>();
    idx->sigma = S;
    idx->samplingRate = samplingRate;
    u64 n = 0, maxLen = 0;
    for (u64 i = 0; i < nSeqs; ++i) {
        n += lens[i] + 1;
        maxLen = std::max(maxLen, lens[i] + 1);
    }
    idx->bitsForPosition = std::max<u64>(1, std::bit_width(maxLen));
    std::vector<u64> seqStart(nSeqs + 1, 0);
    for (u64 i = 0; i < nSeqs; ++i) seqStart[i + 1] = seqStart[i] + lens[i] + 1;

    for (int pass = 0; pass < 2; ++pass) {
        std::vector<u8> text(n);
        u64 in = 0, out = 0;
        for (u64 i = 0; i < nSeqs; ++i) {
            for (u64 j = 0; j < lens[i]; ++j) {
                u8 c = pass == 0 ? seqs[in + j] : seqs[in + lens[i] - 1 - j];
                if (c == 0 || c >= S) throw std::runtime_error("sequence contains a rank outside 1..Sigma-1");
                text[out + j] = c;
            }
            text[out + lens[i]] = 0;
            in += lens[i];
            out += lens[i] + 1;
        }
        auto sa = suffix_array(text);
        std::vector<u8> bwt(n);
        for (u64 r = 0; r < n; ++r) bwt[r] = text[(sa[r] + n - 1) % n];
        if (pass == 0) {
            idx->bwt.build(bwt.data(), n);
            idx->C.fill(0);
            for (u64 i = 0; i < n; ++i) idx->C[text[i] + 1]++;
            for (int c = 0; c < S; ++c) idx->C[c + 1] += idx->C[c];
            idx->marks.n = n;
            idx->marks.bits.assign(n / 64 + 1, 0);
            for (u64 r = 0; r < n; ++r) {
                u64 p = sa[r];
                u64 sid = std::upper_bound(seqStart.begin(), seqStart.end(), p) - seqStart.begin() - 1;
                u64 pos = p - seqStart[sid];
                if (pos % samplingRate == 0) {
                    idx->marks.bits[r >> 6] |= u64{1} << (r & 63);
                    idx->ssa.push_back((sid << idx->bitsForPosition) | pos);
                }
            }
            idx->marks.finalize();
        } else {
            idx->bwtRev.build(bwt.data(), n);
        }
    }
    return idx;
}

// ---------------------------------------------------------------------------------------------
// On-disk format: cereal portable-binary rules — little-endian raw values, u64 size
// tag before every std::vector, std::array<arith,N> raw without tag.  Field order is the
// reconstruction documented in DESIGN.md ("index file layout"); the reader verifies every invariant
// it can and refuses anything it does not understand.
// ---------------------------------------------------------------------------------------------
struct Writer {
    FILE* f;
    void raw(void const* p, size_t n) {
        if (n && fwrite(p, 1, n, f) != n) throw std::runtime_error("write failed");
    }
    void u(u64 v) { raw(&v, 8); }
};
struct Reader {
    FILE* f;
    void raw(void* p, size_t n) {
        if (n && fread(p, 1, n, f) != n) throw std::runtime_error("index layout not understood: unexpected end of file");
    }
    u64 u() {
        u64 v;
        raw(&v, 8);
        return v;
    }
};

template <int S>
static void save_occ(Writer& w, OccTable<S> const& t) {
    w.u(t.blocks.size());
    for (auto const& b : t.blocks) {
        w.raw(b.cnt, sizeof(u16) * S);
        w.raw(b.bits, sizeof(u64) * S);
    }
    w.u(t.superBlocks.size());
    for (auto const& sb : t.superBlocks) w.raw(sb.data(), sizeof(u64) * S);
    w.u(t.n);
}
template <int S>
static void load_occ(Reader& r, OccTable<S>& t) {
    u64 nb = r.u();
    if (nb == 0 || nb > (u64{1} << 40)) throw std::runtime_error("index layout not understood: block count");
    t.blocks.resize(nb);
    for (auto& b : t.blocks) {
        r.raw(b.cnt, sizeof(u16) * S);
        r.raw(b.bits, sizeof(u64) * S);
    }
    u64 ns = r.u();
    if (ns != (nb + 1023) / 1024) throw std::runtime_error("index layout not understood: superblock count");
    t.superBlocks.resize(ns);
    for (auto& sb : t.superBlocks) r.raw(sb.data(), sizeof(u64) * S);
    t.n = r.u();
    if (t.n / 64 + 1 != nb) throw std::runtime_error("index layout not understood: row count vs block count");
}

template <int S>
static void save_index(Index<S> const& idx, char const* path) {
    FILE* f = fopen(path, "wb");
    if (!f) throw std::runtime_error(std::string("cannot open ") + path);
    Writer w{f};
    try {
        w.u(S);
        save_occ(w, idx.bwt);
        save_occ(w, idx.bwtRev);
        w.raw(idx.C.data(), 8 * (S + 1));
        w.u(idx.ssa.size());
        w.raw(idx.ssa.data(), 8 * idx.ssa.size());
        w.u(idx.marks.bits.size());
        w.raw(idx.marks.bits.data(), 8 * idx.marks.bits.size());
        w.u(idx.marks.n);
        w.u(idx.samplingRate);
        w.u(idx.bitsForPosition);
    } catch (...) {
        fclose(f);
        throw;
    }
    fclose(f);
}

template <int S>
static std::unique_ptr<Index<S>> load_index(FILE* f) {
    Reader r{f};
    auto idx = std::make_unique<Index<S>>();
    idx->sigma = S;
    load_occ(r, idx->bwt);
    load_occ(r, idx->bwtRev);
    r.raw(idx->C.data(), 8 * (S + 1));
    u64 ns = r.u();
    if (ns > idx->bwt.n) throw std::runtime_error("index layout not understood: sample count");
    idx->ssa.resize(ns);
    r.raw(idx->ssa.data(), 8 * ns);
    u64 nw = r.u();
    if (nw != idx->bwt.n / 64 + 1) throw std::runtime_error("index layout not understood: marker words");
    idx->marks.bits.resize(nw);
    r.raw(idx->marks.bits.data(), 8 * nw);
    idx->marks.n = r.u();
    idx->samplingRate = r.u();
    idx->bitsForPosition = r.u();
    idx->marks.finalize();
    u8 extra;
    if (fread(&extra, 1, 1, f) != 0) throw std::runtime_error("index layout not understood: trailing bytes");
    // self checks
    u64 n = idx->bwt.n;
    if (idx->bwtRev.n != n || idx->marks.n != n) throw std::runtime_error("index layout not understood: sizes differ");
    if (idx->C[0] != 0 || idx->C[S] != n) throw std::runtime_error("index layout not understood: C array");
    u64 ra[S], rb[S];
    idx->bwt.all_ranks(n, ra);
    idx->bwtRev.all_ranks(n, rb);
    for (int c = 0; c < S; ++c) {
        if (idx->C[c] > idx->C[c + 1]) throw std::runtime_error("index layout not understood: C not monotone");
        if (ra[c] != idx->C[c + 1] - idx->C[c] || rb[c] != ra[c])
            throw std::runtime_error("index layout not understood: symbol histogram");
    }
    if (idx->marks.ones() != ns) throw std::runtime_error("index layout not understood: marked rows != samples");
    return idx;
}


// In-memory image of the index in the on-disk packing (same fields as the file; mirror of the product's
// sb200_index_view, declared here so that the oracle stays self-contained).
struct OrcIndexView {
    u64 sigma, n_rows, n_blocks;
    void const* bwt_blocks;
    u64 const* bwt_super;
    void const* bwtrev_blocks;
    u64 const* bwtrev_super;
    u64 const* C;
    u64 const* ssa;
    u64 n_ssa;
    u64 const* mark_bits;
    u64 sampling_rate, bits_for_position;
};

template <int S>
static std::unique_ptr<Index<S>> index_from_view(OrcIndexView const& v) {
    auto idx = std::make_unique<Index<S>>();
    idx->sigma = S;
    auto fill = [&](OccTable<S>& t, void const* blocks, u64 const* super) {
        t.n = v.n_rows;
        t.blocks.resize(v.n_blocks);
        u8 const* p = static_cast<u8 const*>(blocks);
#if defined(_OPENMP)
#pragma omp parallel for schedule(static)
#endif
        for (u64 b = 0; b < v.n_blocks; ++b) {
            std::memcpy(t.blocks[b].cnt, p + b * 10 * S, 2 * S);
            std::memcpy(t.blocks[b].bits, p + b * 10 * S + 2 * S, 8 * S);
        }
        u64 ns = (v.n_blocks + 1023) / 1024;
        t.superBlocks.resize(ns);
        for (u64 i = 0; i < ns; ++i) std::memcpy(t.superBlocks[i].data(), super + i * S, 8 * S);
    };
    fill(idx->bwt, v.bwt_blocks, v.bwt_super);
    fill(idx->bwtRev, v.bwtrev_blocks, v.bwtrev_super);
    std::memcpy(idx->C.data(), v.C, 8 * (S + 1));
    idx->ssa.assign(v.ssa, v.ssa + v.n_ssa);
    idx->marks.n = v.n_rows;
    idx->marks.bits.assign(v.mark_bits, v.mark_bits + (v.n_rows / 64 + 1));
    idx->marks.finalize();
    idx->samplingRate = v.sampling_rate;
    idx->bitsForPosition = v.bits_for_position;
    if (idx->marks.ones() != v.n_ssa) throw std::runtime_error("index layout not understood: marked rows != samples");
    return idx;
}

// ---------------------------------------------------------------------------------------------
// Bidirectional cursor: (lb, lbRev, len); extendLeft uses bwt, extendRight bwtRev.
// ---------------------------------------------------------------------------------------------
struct Cursor {
    u64 lb, lbRev, len;
};

struct Counters {
    u64 nodes{};      // cursor extensions performed (each = 2 row probes)
    u64 rankOps{};    // row probes
    u64 cursors{};    // cursors reported
    u64 lfSteps{};    // LF steps during locate
    u64 hits{};       // located positions
};

struct Step {
    u16 pi;
    u8 l, u;
    bool right;
};

struct RawCursor {
    u64 qid, lb, len, e;
};
struct Hit {
    u64 qid, seqId, pos, e;
};

template <int S, bool Right>
static inline void extend_all(Index<S> const& ix, Cursor const& cur, Cursor* kids, Counters& ct) {
    auto const& occ = Right ? ix.bwtRev : ix.bwt;
    u64 lo = Right ? cur.lbRev : cur.lb;
    u64 r1[S], r2[S];
    occ.all_ranks(lo, r1);
    occ.all_ranks(lo + cur.len, r2);
    ct.nodes += 1;
    ct.rankOps += 2;
    u64 other = Right ? cur.lb : cur.lbRev;
    for (int c = 0; c < S; ++c) {
        u64 cnt = r2[c] - r1[c];
        u64 own = ix.C[c] + r1[c];
        if constexpr (Right) kids[c] = Cursor{other, own, cnt};
        else kids[c] = Cursor{own, other, cnt};
        other += cnt;
    }
}

template <int S, bool Right>
static inline Cursor extend_one(Index<S> const& ix, Cursor const& cur, int c, Counters& ct) {
    auto const& occ = Right ? ix.bwtRev : ix.bwt;
    u64 lo = Right ? cur.lbRev : cur.lb;
    u64 r1[S], r2[S];
    occ.all_ranks(lo, r1);
    occ.all_ranks(lo + cur.len, r2);
    ct.nodes += 1;
    ct.rankOps += 2;
    u64 smaller = 0;
    for (int s = 0; s < c; ++s) smaller += r2[s] - r1[s];
    u64 own = ix.C[c] + r1[c], cnt = r2[c] - r1[c];
    if constexpr (Right) return Cursor{cur.lb + smaller, own, cnt};
    else return Cursor{own, cur.lbRev + smaller, cnt};
}

// The backtracking search.  LInfo/RInfo = last operation at the left/right end.
template <int S, bool Edit>
struct Searcher {
    Index<S> const& ix;
    std::vector<Step> const& steps;
    u8 const* query;
    u64 qid;
    std::vector<RawCursor>& out;
    Counters& ct;
    // search_n (src/sahara/search.cpp:228,231): at most maxHits suffix-array rows are delivered per query, in the
    // order of the recursion; the cursor that crosses the limit is cut to its first rows and the query ends
    // (0 = unlimited).  *taken = rows delivered for this query so far, shared by the searches of the scheme.
    u64 maxHits{0};
    u64* taken{nullptr};

    bool stopped() const { return maxHits != 0 && *taken >= maxHits; }

    static constexpr unsigned info_code(char c) { return c == 'M' ? SB200_INFO_M : c == 'S' ? SB200_INFO_S : c == 'I' ? SB200_INFO_I : SB200_INFO_D; }

    template <char LInfo, char RInfo>
    void next(Cursor const& cur, int e, size_t i) {
        if (cur.len == 0) return;
        if (stopped()) return;
        if (i == steps.size()) {
            // end filter (policy): by default not behind a substitution or a deletion at either end
            if (!Edit || sb200_pol_end(&g_policy, info_code(LInfo), info_code(RInfo))) {
                u64 len = cur.len;
                if (maxHits != 0) {
                    if (*taken + len > maxHits) len = maxHits - *taken;
                    *taken += len;
                }
                out.push_back(RawCursor{qid, cur.lb, len, static_cast<u64>(e)});
                ct.cursors += 1;
            }
            return;
        }
        if (steps[i].right) dir<LInfo, RInfo, true>(cur, e, i);
        else dir<LInfo, RInfo, false>(cur, e, i);
    }

    template <char LInfo, char RInfo, bool Right>
    void dir(Cursor const& cur, int e, size_t i) {
        constexpr char T = Right ? RInfo : LInfo;
        // which operation may follow T at the end that is extended (policy)
        const bool DelOK = Edit && sb200_pol_del(&g_policy, info_code(T));
        const bool InsOK = Edit && sb200_pol_ins(&g_policy, info_code(T));
        const bool subFirst = (g_policy.child_order & SB200_CHILD_SUB_BEFORE_DEL) != 0;
        const bool insEarly = (g_policy.child_order & SB200_CHILD_INS_BEFORE_SYMBOLS) != 0;
        constexpr char ML = Right ? LInfo : 'M', MR = Right ? 'M' : RInfo;
        constexpr char SL = Right ? LInfo : 'S', SR = Right ? 'S' : RInfo;
        constexpr char DL = Right ? LInfo : 'D', DR = Right ? 'D' : RInfo;
        constexpr char IL = Right ? LInfo : 'I', IR = Right ? 'I' : RInfo;
        auto const& st = steps[i];
        int c = query[st.pi];
        bool matchOK = st.l <= e && e <= st.u;
        bool mismatchOK = st.l <= e + 1 && e + 1 <= st.u;
        if (mismatchOK) {
            Cursor kids[S];
            extend_all<S, Right>(ix, cur, kids, ct);
            // order of the children (policy): match; then per symbol deletion and substitution; the insertion last
            if (matchOK) next<ML, MR>(kids[c], e, i + 1);
            if (InsOK && insEarly) next<IL, IR>(cur, e + 1, i + 1);
            for (int s = 1; s < S; ++s) {
                if (s == c) continue;
                if (subFirst) next<SL, SR>(kids[s], e + 1, i + 1);
                if (DelOK) next<DL, DR>(kids[s], e + 1, i);
                if (!subFirst) next<SL, SR>(kids[s], e + 1, i + 1);
            }
            if (InsOK && !insEarly) next<IL, IR>(cur, e + 1, i + 1);
        } else if (matchOK) {
            next<ML, MR>(extend_one<S, Right>(ix, cur, c, ct), e, i + 1);
        }
    }
};

struct Scheme {
    std::vector<std::vector<Step>> searches;
};

static Scheme make_scheme(u64 nSearches, u64 m, u16 const* pi, u8 const* l, u8 const* u) {
    Scheme s;
    for (u64 j = 0; j < nSearches; ++j) {
        std::vector<Step> st(m);
        for (u64 i = 0; i < m; ++i) {
            bool right;
            if (i == 0) right = m < 2 ? true : pi[j * m] < pi[j * m + 1];
            else right = pi[j * m + i - 1] < pi[j * m + i];
            st[i] = Step{pi[j * m + i], l[j * m + i], u[j * m + i], right};
        }
        s.searches.push_back(std::move(st));
    }
    return s;
}

template <int S>
static void run_search(Index<S> const& ix, u8 const* queries, u64 nq, u64 m, Scheme const& sch, bool edit, int threads,
                       std::vector<RawCursor>& out, Counters& total, u64 maxHits = 0) {
    if (threads < 1) threads = 1;
    std::vector<std::vector<RawCursor>> outs(threads);
    std::vector<Counters> cts(threads);
    auto body = [&](int t, u64 q0, u64 q1) {
        for (u64 q = q0; q < q1; ++q) {
            u64 taken = 0;
            for (auto const& st : sch.searches) {
                Cursor root{0, 0, ix.size()};
                if (edit) {
                    Searcher<S, true> s{ix, st, queries + q * m, q, outs[t], cts[t], maxHits, &taken};
                    s.template next<'M', 'M'>(root, 0, 0);
                } else {
                    Searcher<S, false> s{ix, st, queries + q * m, q, outs[t], cts[t], maxHits, &taken};
                    s.template next<'M', 'M'>(root, 0, 0);
                }
            }
        }
    };
    if (threads == 1) {
        body(0, 0, nq);
    } else {
#if defined(_OPENMP)
        // contiguous chunks per thread so that concatenation keeps query order
#pragma omp parallel num_threads(threads)
        {
            int t = omp_get_thread_num();
            u64 per = (nq + threads - 1) / threads;
            u64 q0 = std::min<u64>(nq, per * t), q1 = std::min<u64>(nq, per * (t + 1));
            body(t, q0, q1);
        }
#else
        u64 per = (nq + threads - 1) / threads;
        for (int t = 0; t < threads; ++t) body(t, std::min<u64>(nq, per * t), std::min<u64>(nq, per * (t + 1)));
#endif
    }
    for (int t = 0; t < threads; ++t) {
        out.insert(out.end(), outs[t].begin(), outs[t].end());
        total.nodes += cts[t].nodes;
        total.rankOps += cts[t].rankOps;
        total.cursors += cts[t].cursors;
    }
}

// LocateLinear: per SA row LF-walk to a marked row; position = sample + steps.
template <int S>
static inline void locate_row(Index<S> const& ix, u64 row, u64& seqId, u64& pos, u64& steps) {
    steps = 0;
    while (!ix.marks.test(row)) {
        int c = ix.bwt.symbol(row);
        row = ix.C[c] + ix.bwt.rank(row, c);
        ++steps;
    }
    u64 v = ix.ssa[ix.marks.rank(row)];
    seqId = v >> ix.bitsForPosition;
    pos = (v & ((u64{1} << ix.bitsForPosition) - 1)) + steps;
}

template <int S>
static void run_locate(Index<S> const& ix, RawCursor const* cur, u64 n, int threads, std::vector<Hit>& out, Counters& ct) {
    std::vector<u64> off(n + 1, 0);
    for (u64 i = 0; i < n; ++i) off[i + 1] = off[i] + cur[i].len;
    out.resize(off[n]);
    u64 lf = 0;
    if (threads < 1) threads = 1;
#if defined(_OPENMP)
#pragma omp parallel for num_threads(threads) schedule(dynamic, 256) reduction(+ : lf)
#endif
    for (u64 i = 0; i < n; ++i) {
        for (u64 r = 0; r < cur[i].len; ++r) {
            u64 sid, pos, steps;
            locate_row(ix, cur[i].lb + r, sid, pos, steps);
            out[off[i] + r] = Hit{cur[i].qid, sid, pos, cur[i].e};
            lf += steps;
        }
    }
    ct.lfSteps += lf;
    ct.hits += off[n];
}

template <typename F>
static auto dispatch(IndexBase* b, F&& f) {
    if (b->sigma == 5) return f(*static_cast<Index<5>*>(b));
    if (b->sigma == 6) return f(*static_cast<Index<6>*>(b));
    throw std::runtime_error("unknown index with " + std::to_string(b->sigma) + " letters");
}

template <typename F>
static int guard(F&& f) {
    try {
        f();
        return 0;
    } catch (std::exception const& e) {
        g_error = e.what();
        return 1;
    } catch (...) {
        g_error = "unknown error";
        return 1;
    }
}

}  // namespace

// ---------------------------------------------------------------------------------------------
// C entry points (ctypes).  All return 0 on success; orc_last_error() gives the message otherwise.
// ---------------------------------------------------------------------------------------------
extern "C" {

char const* orc_last_error() { return g_error.c_str(); }

int orc_max_threads() {
#if defined(_OPENMP)
    return omp_get_max_threads();
#else
    return 1;
#endif
}

// seqs: concatenated ranks (no delimiters), lens[nSeqs]
int orc_index_build(u8 const* seqs, u64 const* lens, u64 nSeqs, int sigma, u64 samplingRate, void** out) {
    return guard([&] {
        if (nSeqs == 0) throw std::runtime_error("reference was empty - abort");
        if (sigma == 5) *out = static_cast<IndexBase*>(build_index<5>(seqs, lens, nSeqs, samplingRate).release());
        else if (sigma == 6) *out = static_cast<IndexBase*>(build_index<6>(seqs, lens, nSeqs, samplingRate).release());
        else throw std::runtime_error("unknown index with " + std::to_string(sigma) + " letters");
    });
}

int orc_index_from_view(void const* view, void** out) {
    return guard([&] {
        auto const& v = *static_cast<OrcIndexView const*>(view);
        if (v.sigma == 5) *out = static_cast<IndexBase*>(index_from_view<5>(v).release());
        else if (v.sigma == 6) *out = static_cast<IndexBase*>(index_from_view<6>(v).release());
        else throw std::runtime_error("unknown index with " + std::to_string(v.sigma) + " letters");
    });
}

void orc_index_free(void* h) { delete static_cast<IndexBase*>(h); }

int orc_index_save(void* h, char const* path) {
    return guard([&] { dispatch(static_cast<IndexBase*>(h), [&](auto& ix) { save_index(ix, path); return 0; }); });
}

int orc_index_load(char const* path, void** out) {
    return guard([&] {
        FILE* f = fopen(path, "rb");
        if (!f) throw std::runtime_error(std::string("no valid index path at ") + path);
        try {
            u64 sigma;
            if (fread(&sigma, 8, 1, f) != 1) throw std::runtime_error("index layout not understood: empty file");
            if (sigma == 5) *out = static_cast<IndexBase*>(load_index<5>(f).release());
            else if (sigma == 6) *out = static_cast<IndexBase*>(load_index<6>(f).release());
            else throw std::runtime_error("unknown index with " + std::to_string(sigma) + " letters");
        } catch (...) {
            fclose(f);
            throw;
        }
        fclose(f);
    });
}

// info[0]=sigma, [1]=rows, [2]=samples, [3]=samplingRate, [4]=bitsForPosition, [5..5+sigma]=C
int orc_index_info(void* h, u64* info) {
    return guard([&] {
        dispatch(static_cast<IndexBase*>(h), [&](auto& ix) {
            info[0] = ix.sigma;
            info[1] = ix.size();
            info[2] = ix.ssa.size();
            info[3] = ix.samplingRate;
            info[4] = ix.bitsForPosition;
            for (int c = 0; c <= ix.sigma; ++c) info[5 + c] = ix.C[c];
            return 0;
        });
    });
}

// which: 0 = bwt, 1 = bwtRev.  out[n][sigma]
int orc_all_ranks(void* h, int which, u64 const* pos, u64 n, u64* out) {
    return guard([&] {
        dispatch(static_cast<IndexBase*>(h), [&](auto& ix) {
            auto const& t = which ? ix.bwtRev : ix.bwt;
            for (u64 i = 0; i < n; ++i) {
                if (pos[i] > t.n) throw std::runtime_error("rank position out of range");
                t.all_ranks(pos[i], out + i * ix.sigma);
            }
            return 0;
        });
    });
}

int orc_bwt_symbols(void* h, int which, u8* out) {
    return guard([&] {
        dispatch(static_cast<IndexBase*>(h), [&](auto& ix) {
            auto const& t = which ? ix.bwtRev : ix.bwt;
            for (u64 i = 0; i < t.n; ++i) out[i] = static_cast<u8>(t.symbol(i));
            return 0;
        });
    });
}

// locate single rows: out[3*i] = seqId, pos, steps
int orc_locate_rows(void* h, u64 const* rows, u64 n, u64* out) {
    return guard([&] {
        dispatch(static_cast<IndexBase*>(h), [&](auto& ix) {
            for (u64 i = 0; i < n; ++i) locate_row(ix, rows[i], out[3 * i], out[3 * i + 1], out[3 * i + 2]);
            return 0;
        });
    });
}

// marker bits + sampled values (for comparing against the product's device structures)
int orc_index_samples(void* h, u64* marksOut /* rows/64+1 words */, u64* ssaOut) {
    return guard([&] {
        dispatch(static_cast<IndexBase*>(h), [&](auto& ix) {
            std::memcpy(marksOut, ix.marks.bits.data(), 8 * ix.marks.bits.size());
            std::memcpy(ssaOut, ix.ssa.data(), 8 * ix.ssa.size());
            return 0;
        });
    });
}

// search_ng24::search<Edit> over dense queries[nq][m]; scheme tables pi/l/u are [nSearches][m].
// counters[0..4] = nodes, rankOps, cursors, lfSteps, hits (accumulated).
static int search_entry(void* h, u8 const* queries, u64 nq, u64 m, u64 nSearches, u16 const* pi, u8 const* l, u8 const* u, int edit,
                        int threads, u64 maxHits, u64** cursorsOut, u64* nOut, u64* counters) {
    return guard([&] {
        if (nq == 0) throw std::runtime_error("query file was empty - abort");
        for (u64 j = 0; j < nSearches; ++j)
            for (u64 i = 0; i < m; ++i)
                if (pi[j * m + i] >= m) throw std::runtime_error("search scheme does not fit the query length");
        Scheme sch = make_scheme(nSearches, m, pi, l, u);
        std::vector<RawCursor> out;
        Counters ct;
        dispatch(static_cast<IndexBase*>(h), [&](auto& ix) {
            for (u64 i = 0; i < nq * m; ++i)
                if (queries[i] >= ix.sigma) throw std::runtime_error("query has invalid character");  // rank 0 ('$') is a valid character of the alphabet
            run_search(ix, queries, nq, m, sch, edit != 0, threads, out, ct, maxHits);
            return 0;
        });
        u64* buf = static_cast<u64*>(std::malloc(std::max<size_t>(1, out.size()) * sizeof(RawCursor)));
        std::memcpy(buf, out.data(), out.size() * sizeof(RawCursor));
        *cursorsOut = buf;
        *nOut = out.size();
        if (counters) {
            counters[0] += ct.nodes;
            counters[1] += ct.rankOps;
            counters[2] += ct.cursors;
        }
    });
}

int orc_search(void* h, u8 const* queries, u64 nq, u64 m, u64 nSearches, u16 const* pi, u8 const* l, u8 const* u, int edit,
               int threads, u64** cursorsOut, u64* nOut, u64* counters) {
    return search_entry(h, queries, nq, m, nSearches, pi, l, u, edit, threads, 0, cursorsOut, nOut, counters);
}

// search_ng24::search_n<Edit>(index, queries, scheme, maxHits, cb) (src/sahara/search.cpp:228,231): as orc_search, but a
// query ends once maxHits suffix-array rows were delivered (the last cursor is cut to its first rows).  The
// recursion order decides which hits those are.  [RECALL low: SURVEY.md 9.4]
int orc_search_n(void* h, u8 const* queries, u64 nq, u64 m, u64 nSearches, u16 const* pi, u8 const* l, u8 const* u, int edit,
                 int threads, u64 maxHits, u64** cursorsOut, u64* nOut, u64* counters) {
    return search_entry(h, queries, nq, m, nSearches, pi, l, u, edit, threads, maxHits, cursorsOut, nOut, counters);
}

// LocateLinear over cursors (qid, lb, len, e) -> hits (qid, seqId, pos, e), in cursor order then row order.
int orc_locate(void* h, u64 const* cursors, u64 n, int threads, u64** hitsOut, u64* nOut, u64* counters) {
    return guard([&] {
        std::vector<Hit> out;
        Counters ct;
        dispatch(static_cast<IndexBase*>(h), [&](auto& ix) {
            run_locate(ix, reinterpret_cast<RawCursor const*>(cursors), n, threads, out, ct);
            return 0;
        });
        u64* buf = static_cast<u64*>(std::malloc(std::max<size_t>(1, out.size()) * sizeof(Hit)));
        std::memcpy(buf, out.data(), out.size() * sizeof(Hit));
        *hitsOut = buf;
        *nOut = out.size();
        if (counters) {
            counters[3] += ct.lfSteps;
            counters[4] += ct.hits;
        }
    });
}

void orc_free(void* p) { std::free(p); }

// replaces the policy table in force (include/sahara_policy.h); nullptr restores the default
int orc_set_policy(sb200_policy const* p) {
    return guard([&] {
        sb200_policy def = SB200_POLICY_DEFAULT;
        if (p && !sb200_pol_valid(p)) throw std::runtime_error("invalid search policy");
        g_policy = p ? *p : def;
    });
}

// ---------------------------------------------------------------------------------------------
// Brute-force checkers (implementation-independent ground truth for tests).
// ---------------------------------------------------------------------------------------------

// All windows of one sequence with Hamming distance <= k.  Returns count; fills (pos, e) pairs up to cap.
u64 orc_bf_hamming(u8 const* seq, u64 len, u8 const* q, u64 m, u64 k, u64* out, u64 cap) {
    u64 n = 0;
    if (len < m) return 0;
    for (u64 p = 0; p + m <= len; ++p) {
        u64 e = 0;
        for (u64 i = 0; i < m && e <= k; ++i) e += seq[p + i] != q[i];
        if (e <= k) {
            if (n < cap) {
                out[2 * n] = p;
                out[2 * n + 1] = e;
            }
            ++n;
        }
    }
    return n;
}

// For every start position p: min over window lengths of the edit distance between q and seq[p, p+L),
// capped at k+1.  out[len+1].
void orc_bf_edit_starts(u8 const* seq, u64 len, u8 const* q, u64 m, u64 k, u8* out) {
    std::vector<u32> prev(m + 1), cur(m + 1);
    for (u64 p = 0; p <= len; ++p) {
        // D[j] after consuming t text chars = edit distance between q[0..j) and seq[p..p+t)
        for (u64 j = 0; j <= m; ++j) prev[j] = static_cast<u32>(j);
        u32 best = prev[m];
        u64 maxT = std::min<u64>(len - p, m + k);
        for (u64 t = 1; t <= maxT; ++t) {
            cur[0] = static_cast<u32>(t);
            u8 c = seq[p + t - 1];
            u32 rowMin = cur[0];
            for (u64 j = 1; j <= m; ++j) {
                u32 v = prev[j - 1] + (q[j - 1] != c);
                v = std::min(v, prev[j] + 1);
                v = std::min(v, cur[j - 1] + 1);
                cur[j] = v;
                rowMin = std::min(rowMin, v);
            }
            best = std::min(best, cur[m]);
            std::swap(prev, cur);
            if (rowMin > k) break;
        }
        out[p] = static_cast<u8>(std::min<u32>(best, static_cast<u32>(k + 1)));
    }
}

}  // extern "C"
