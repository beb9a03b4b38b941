// scheme.hpp — search schemes: generators, expansion to query length, validity / completeness checks.
//
// Host-side table generator of the hot path.  Replaces, behind the same names and
// argument meaning, what `sahara search` obtains from fmindex-collection at
//   /root/reference/src/sahara/search.cpp:174-212  (generator lookup by name, generator(minK,maxK,0,0),
//                                                    expand(scheme, queryLength))
//   /root/reference/src/sahara/search.cpp:226      (limitToHamming)
//   /root/reference/src/sahara/search_scheme.cpp:192 (the list of generator names)
//   /root/reference/src/sahara/search_scheme.cpp:252-276 (Columba text format "{pi} {L} {U}")
// A Search has the members pi, l, u (0-based parts) as used at /root/reference/src/sahara/tikz.h:14-26.
//
// Provenance of the tables (see DESIGN.md "search schemes"): fmindex-collection is not available here,
// so generators are implemented from their published definitions; every scheme returned by
// generate() is checked with isValid() and isComplete() and generation fails loudly otherwise.
#pragma once
#include <algorithm>
#include <cstddef>
#include <cstdint>
#include <functional>
#include <map>
#include <numeric>
#include <sstream>
#include <stdexcept>
#include <string>
#include <vector>

#include "../../include/sahara_policy.h"

namespace sahara::scheme {

struct Search {
    std::vector<size_t> pi, l, u;
    bool operator==(Search const&) const = default;
};
using Scheme = std::vector<Search>;

// ---- checks --------------------------------------------------------------------------------------

// pi is a permutation of 0..P-1 whose every prefix is a contiguous range; l,u non-decreasing, l<=u.
inline bool isValid(Search const& s) {
    size_t P = s.pi.size();
    if (P == 0 || s.l.size() != P || s.u.size() != P) return false;
    std::vector<bool> seen(P, false);
    size_t lo = s.pi[0], hi = s.pi[0];
    for (size_t i = 0; i < P; ++i) {
        size_t p = s.pi[i];
        if (p >= P || seen[p]) return false;
        seen[p] = true;
        if (i > 0) {
            if (p + 1 == lo) lo = p;
            else if (p == hi + 1) hi = p;
            else return false;
        }
        if (s.l[i] > s.u[i]) return false;
        if (i > 0 && (s.l[i] < s.l[i - 1] || s.u[i] < s.u[i - 1])) return false;
    }
    return true;
}
inline bool isValid(Scheme const& ss) {
    if (ss.empty()) return false;
    for (auto const& s : ss)
        if (!isValid(s) || s.pi.size() != ss[0].pi.size()) return false;
    return true;
}

// does the search accept the error distribution cfg (errors per part)?
inline bool covers(Search const& s, std::vector<size_t> const& cfg) {
    size_t a = 0;
    for (size_t i = 0; i < s.pi.size(); ++i) {
        a += cfg[s.pi[i]];
        if (a < s.l[i] || a > s.u[i]) return false;
    }
    return true;
}

// calls cb for every distribution of e errors (minK <= e <= maxK) over P parts
template <typename CB>
void forEachErrorConfig(size_t P, size_t minK, size_t maxK, CB&& cb) {
    std::vector<size_t> cfg(P, 0);
    std::function<void(size_t, size_t)> rec = [&](size_t part, size_t used) {
        if (part == P) {
            if (used >= minK) cb(cfg);
            return;
        }
        for (size_t e = 0; used + e <= maxK; ++e) {
            cfg[part] = e;
            rec(part + 1, used + e);
        }
        cfg[part] = 0;
    };
    rec(0, 0);
}

inline bool isComplete(Scheme const& ss, size_t minK, size_t maxK) {
    if (ss.empty()) return false;
    bool ok = true;
    forEachErrorConfig(ss[0].pi.size(), minK, maxK, [&](std::vector<size_t> const& cfg) {
        bool any = false;
        for (auto const& s : ss) any = any || covers(s, cfg);
        ok = ok && any;
    });
    return ok;
}

inline bool isNonRedundant(Scheme const& ss, size_t minK, size_t maxK) {
    if (ss.empty()) return false;
    bool ok = true;
    forEachErrorConfig(ss[0].pi.size(), minK, maxK, [&](std::vector<size_t> const& cfg) {
        size_t n = 0;
        for (auto const& s : ss) n += covers(s, cfg);
        ok = ok && n == 1;
    });
    return ok;
}

// ---- expansion to the query length (one entry per query character) ---------------------------------

// part sizes: len / P each, the first len % P parts one more
inline std::vector<size_t> expandCount(size_t parts, size_t len) {
    std::vector<size_t> c(parts, len / parts);
    for (size_t i = 0; i < len % parts; ++i) c[i] += 1;
    return c;
}

inline bool isExpandable(Search const& s, size_t len) { return s.pi.size() <= len; }

// Rule `expand_lower` of the policy table (include/sahara_policy.h): the lower bound of the characters of a part that
// are not its last one.  0 (the reconstruction in force): the previous part's lower bound — a part's own bound is only
// demanded once its last character is consumed; 1: the part's own bound at every character.
inline uint32_t& expandLowerRule() {
    static uint32_t rule = [] { sb200_policy p = SB200_POLICY_DEFAULT; return p.expand_lower; }();
    return rule;
}

inline Search expand(Search const& s, std::vector<size_t> const& counts) {
    size_t P = s.pi.size();
    std::vector<size_t> start(P + 1, 0);
    for (size_t i = 0; i < P; ++i) start[i + 1] = start[i] + counts[i];
    Search r;
    for (size_t i = 0; i < P; ++i) {
        size_t part = s.pi[i];
        bool right = (i == 0) ? (P < 2 || s.pi[0] < s.pi[1]) : (s.pi[i - 1] < s.pi[i]);
        for (size_t j = 0; j < counts[part]; ++j) {
            r.pi.push_back(right ? start[part] + j : start[part + 1] - 1 - j);
            // lower bound: previous part's value until the last character of the part (policy rule expand_lower)
            bool last = j + 1 == counts[part];
            r.l.push_back((last || expandLowerRule() == 1u) ? s.l[i] : (i > 0 ? s.l[i - 1] : 0));
            r.u.push_back(s.u[i]);
        }
    }
    return r;
}

inline Scheme expand(Scheme const& ss, std::vector<size_t> const& counts) {
    Scheme r;
    for (auto const& s : ss) r.push_back(expand(s, counts));
    return r;
}

inline Scheme expand(Scheme const& ss, size_t len) {
    if (ss.empty()) return {};
    size_t P = ss[0].pi.size();
    if (P > len) throw std::runtime_error("search scheme has more parts than the query has characters");
    return expand(ss, expandCount(P, len));
}

// after i+1 characters at most i+1 mismatches are possible — result neutral under Hamming distance
inline Search limitToHamming(Search s) {
    for (size_t i = 0; i < s.u.size(); ++i) s.u[i] = std::min(s.u[i], i + 1);
    return s;
}
inline Scheme limitToHamming(Scheme ss) {
    for (auto& s : ss) s = limitToHamming(s);
    return ss;
}

// ---- node counts (diagnostics printed by `sahara search`, search.cpp:197-198) ----------------------

// number of search-tree nodes of an expanded scheme over an alphabet of `sigma` symbols (incl. the
// delimiter, which is never a branch: sigma-1 usable symbols).  Edit adds insertion/deletion edges.
template <bool Edit>
long double nodeCount(Search const& s, size_t sigma) {
    size_t n = s.pi.size();
    size_t maxE = s.u.empty() ? 0 : s.u.back();
    long double alt = static_cast<long double>(sigma > 2 ? sigma - 2 : 0);  // mismatching symbols
    // cnt[e] = number of nodes at the current depth with e errors
    std::vector<long double> cnt(maxE + 2, 0.0L), nxt;
    cnt[0] = 1;
    long double total = 0;
    for (size_t i = 0; i < n; ++i) {
        nxt.assign(maxE + 2, 0.0L);
        if constexpr (Edit) {
            // deletions: stay on the step, consume a text character (counted as nodes of this step)
            for (size_t e = 0; e + 1 <= maxE; ++e) {
                if (s.l[i] <= e + 1 && e + 1 <= s.u[i]) {
                    long double d = cnt[e] * alt;
                    cnt[e + 1] += d;
                    total += d;
                }
            }
        }
        for (size_t e = 0; e <= maxE; ++e) {
            if (cnt[e] == 0) continue;
            if (s.l[i] <= e && e <= s.u[i]) nxt[e] += cnt[e];
            if (s.l[i] <= e + 1 && e + 1 <= s.u[i]) {
                nxt[e + 1] += cnt[e] * alt;
                if constexpr (Edit) nxt[e + 1] += cnt[e];  // insertion
            }
        }
        for (size_t e = 0; e <= maxE; ++e) total += nxt[e];
        cnt = nxt;
    }
    return total;
}
template <bool Edit>
long double nodeCount(Scheme const& ss, size_t sigma) {
    long double t = 0;
    for (auto const& s : ss) t += nodeCount<Edit>(s, sigma);
    return t;
}

// expected number of nodes that still have occurrences in a random text of length n:
// a node at depth d survives with probability min(1, n / (sigma-1)^d)
template <bool Edit>
long double weightedNodeCount(Search const& s, size_t sigma, size_t n) {
    size_t len = s.pi.size();
    size_t maxE = s.u.empty() ? 0 : s.u.back();
    long double alt = static_cast<long double>(sigma > 2 ? sigma - 2 : 0);
    long double base = static_cast<long double>(sigma > 1 ? sigma - 1 : 1);
    std::vector<long double> cnt(maxE + 2, 0.0L), nxt;
    cnt[0] = 1;
    long double total = 0;
    long double prob = static_cast<long double>(n);  // expected occurrences of a depth-d string
    for (size_t i = 0; i < len; ++i) {
        prob /= base;
        long double w = std::min<long double>(1.0L, prob);
        nxt.assign(maxE + 2, 0.0L);
        for (size_t e = 0; e <= maxE; ++e) {
            if (cnt[e] == 0) continue;
            if (s.l[i] <= e && e <= s.u[i]) nxt[e] += cnt[e];
            if (s.l[i] <= e + 1 && e + 1 <= s.u[i]) {
                nxt[e + 1] += cnt[e] * alt;
                if constexpr (Edit) nxt[e + 1] += cnt[e] * (1 + alt);  // insertion + deletions
            }
        }
        for (size_t e = 0; e <= maxE; ++e) total += nxt[e] * w;
        cnt = nxt;
    }
    return total;
}
template <bool Edit>
long double weightedNodeCount(Scheme const& ss, size_t sigma, size_t n) {
    long double t = 0;
    for (auto const& s : ss) t += weightedNodeCount<Edit>(s, sigma, n);
    return t;
}

// ---- generators ------------------------------------------------------------------------------------
// ---- dynamic expansion (`sahara search --dynamic_generator`, search.cpp:193-195,203-205) ------------
// optimizeByWNCTopDown / expandByWNCTopDown of fmindex-collection: part sizes chosen to minimise the weighted node
// count instead of len / P each.  Upstream source is not available here; this is a reconstruction
// of the published idea: start from the uniform partition ("top") and move `steps` characters from one part to
// another as long as the weighted node count over a text of N symbols goes down (first improvement, parts in
// order; every part keeps at least one character).  The partition is printed by the CLI like upstream's
// "partition: [..]" line; the hit SET of a complete scheme does not depend on it.
template <bool Edit>
inline std::vector<size_t> optimizeByWNCTopDown(Scheme const& ss, size_t len, size_t sigma, size_t N, size_t steps) {
    if (ss.empty()) return {};
    size_t const P = ss[0].pi.size();
    if (P > len) throw std::runtime_error("search scheme has more parts than the query has characters");
    if (steps == 0) steps = 1;
    auto counts = expandCount(P, len);
    auto best = weightedNodeCount<Edit>(expand(ss, counts), sigma, N);
    bool improved = true;
    while (improved) {
        improved = false;
        for (size_t i = 0; i < P; ++i)
            for (size_t j = 0; j < P; ++j) {
                if (i == j || counts[i] <= steps) continue;
                counts[i] -= steps;
                counts[j] += steps;
                auto v = weightedNodeCount<Edit>(expand(ss, counts), sigma, N);
                if (v < best) {
                    best = v;
                    improved = true;
                } else {
                    counts[i] += steps;
                    counts[j] -= steps;
                }
            }
    }
    return counts;
}

template <bool Edit>
inline Scheme expandByWNCTopDown(Scheme const& ss, size_t len, size_t sigma, size_t N, size_t steps) {
    return expand(ss, optimizeByWNCTopDown<Edit>(ss, len, sigma, N, steps));
}

namespace generator {

// order of parts for a search that starts at part `start`, walks right to the end, then left to 0
inline std::vector<size_t> rightThenLeft(size_t P, size_t start) {
    std::vector<size_t> pi;
    for (size_t j = start; j < P; ++j) pi.push_back(j);
    for (size_t j = start; j-- > 0;) pi.push_back(j);
    return pi;
}

inline void applyMinK(Scheme& ss, size_t minK) {
    for (auto& s : ss) s.l.back() = std::max(s.l.back(), minK);
}

// one part, plain backtracking with minK..K errors
inline Scheme backtracking(size_t minK, size_t K) { return {Search{{0}, {minK}, {K}}}; }

// pigeonhole: K+1 parts, one of them is free of errors; search i starts with part i exactly
inline Scheme pigeon_trivial(size_t minK, size_t K) {
    size_t P = K + 1;
    Scheme ss;
    for (size_t i = 0; i < P; ++i) {
        Search s{rightThenLeft(P, i), std::vector<size_t>(P, 0), std::vector<size_t>(P, K)};
        s.u[0] = 0;
        ss.push_back(s);
    }
    applyMinK(ss, minK);
    return ss;
}

// pigeonhole, non redundant: search i covers exactly the distributions whose first error-free part is i,
// hence every part left of i carries at least one error.
inline Scheme pigeon_opt(size_t minK, size_t K) {
    size_t P = K + 1;
    Scheme ss;
    for (size_t i = 0; i < P; ++i) {
        Search s{rightThenLeft(P, i), std::vector<size_t>(P, 0), std::vector<size_t>(P, 0)};
        size_t nRight = P - i;  // seed + parts to its right
        for (size_t j = 1; j < nRight; ++j) s.u[j] = K - i;
        for (size_t j = 1; j <= i; ++j) {  // j-th part to the left
            s.l[nRight - 1 + j] = j;
            s.u[nRight - 1 + j] = K - (i - j);
        }
        ss.push_back(s);
    }
    applyMinK(ss, minK);
    return ss;
}

// suffix filter (Kaerkkaeinen & Na): K+1 parts; there is a part i such that the j-th part after it holds
// at most j cumulated errors.
inline Scheme suffixFilter(size_t minK, size_t K) {
    size_t P = K + 1;
    Scheme ss;
    for (size_t i = 0; i < P; ++i) {
        Search s{rightThenLeft(P, i), std::vector<size_t>(P, 0), std::vector<size_t>(P, K)};
        for (size_t j = 0; j < P - i; ++j) s.u[j] = std::min(j, K);
        ss.push_back(s);
    }
    applyMinK(ss, minK);
    return ss;
}

// 01*0 seeds (Vroland et al.): K+2 parts; every distribution contains a part without error, followed by
// parts with exactly one error each, followed by a part without error.
inline Scheme zeroOnesZero_trivial(size_t minK, size_t K) {
    size_t P = K + 2;
    Scheme ss;
    for (size_t i = 0; i + 1 < P; ++i) {
        for (size_t j = i + 1; j < P; ++j) {
            Search s{rightThenLeft(P, i), std::vector<size_t>(P, 0), std::vector<size_t>(P, K)};
            for (size_t t = 0; t <= j - i; ++t) {
                size_t e = (t == j - i) ? t - 1 : t;  // cumulated errors after seed part i+t
                s.l[t] = e;
                s.u[t] = e;
            }
            for (size_t t = j - i + 1; t < P; ++t) s.l[t] = j - i - 1;
            ss.push_back(s);
        }
    }
    applyMinK(ss, minK);
    return ss;
}

// 01*0 with the seeds that share a start merged into one search (upper bound t after t parts)
inline Scheme zeroOnesZero_opt(size_t minK, size_t K) {
    size_t P = K + 2;
    Scheme ss;
    for (size_t i = 0; i + 1 < P; ++i) {
        Search s{rightThenLeft(P, i), std::vector<size_t>(P, 0), std::vector<size_t>(P, K)};
        for (size_t t = 0; t < P - i; ++t) s.u[t] = std::min(t, K);
        // the seed pattern needs a second error-free part: the last right part may not add the (t)-th error
        s.u[P - i - 1] = std::min<size_t>(P - i - 1 > 0 ? P - i - 2 : 0, K);
        ss.push_back(s);
    }
    applyMinK(ss, minK);
    return ss;
}

// Constructive scheme for any P >= K+1 ("zero-run seeded").  Every error distribution is assigned to one
// search: the one starting at the leftmost part of its longest run of error-free parts; the search order
// is right-then-left and the bounds are the envelope of the assigned distributions.  Used where no
// published table is available (reconstruction of the h2 family).
inline Scheme zeroRunSeeded(size_t P, size_t minK, size_t K) {
    if (P < K + 1) throw std::runtime_error("zeroRunSeeded needs at least K+1 parts");
    std::vector<Search> env(P);
    std::vector<bool> used(P, false);
    forEachErrorConfig(P, minK, K, [&](std::vector<size_t> const& cfg) {
        size_t best = 0, bestLen = 0;
        for (size_t i = 0; i < P;) {
            if (cfg[i] != 0) { ++i; continue; }
            size_t j = i;
            while (j < P && cfg[j] == 0) ++j;
            if (j - i > bestLen) { bestLen = j - i; best = i; }
            i = j;
        }
        auto pi = rightThenLeft(P, best);
        auto& s = env[best];
        if (!used[best]) {
            s.pi = pi;
            s.l.assign(P, K);
            s.u.assign(P, 0);
            used[best] = true;
        }
        size_t a = 0;
        for (size_t t = 0; t < P; ++t) {
            a += cfg[pi[t]];
            s.l[t] = std::min(s.l[t], a);
            s.u[t] = std::max(s.u[t], a);
        }
    });
    Scheme ss;
    for (size_t i = 0; i < P; ++i) {
        if (!used[i]) continue;
        auto s = env[i];
        for (size_t t = 1; t < P; ++t) {  // envelopes are monotone already; keep it explicit
            s.l[t] = std::max(s.l[t], s.l[t - 1]);
            s.u[t] = std::max(s.u[t], s.u[t - 1]);
        }
        ss.push_back(s);
    }
    return ss;
}

// published tables, written 1-based as in the papers
inline Search table(std::vector<size_t> pi1, std::vector<size_t> l, std::vector<size_t> u) {
    for (auto& p : pi1) p -= 1;
    return Search{std::move(pi1), std::move(l), std::move(u)};
}

// Optimum search schemes (Kianfar et al. 2018; the K=2 scheme with four parts is the one also shipped as
// SeqAn's optimum_search_scheme<0,2>).  minK > 0 or K > 2: constructive fallback.
inline Scheme optimum(size_t minK, size_t K) {
    if (minK == 0 && K == 0) return backtracking(0, 0);
    if (minK == 0 && K == 1) return {table({1, 2}, {0, 0}, {0, 1}), table({2, 1}, {0, 1}, {0, 1})};
    if (minK == 0 && K == 2)
        return {table({1, 2, 3, 4}, {0, 0, 1, 1}, {0, 0, 2, 2}), table({3, 2, 1, 4}, {0, 0, 0, 0}, {0, 1, 1, 2}),
                table({4, 3, 2, 1}, {0, 0, 0, 2}, {0, 1, 2, 2})};
    return zeroRunSeeded(K + 2, minK, K);
}

// Kianfar et al., schemes with K+1 parts
inline Scheme kianfar(size_t minK, size_t K) {
    if (minK == 0 && K == 0) return backtracking(0, 0);
    if (minK == 0 && K == 1) return {table({1, 2}, {0, 0}, {0, 1}), table({2, 1}, {0, 1}, {0, 1})};
    if (minK == 0 && K == 2)
        return {table({1, 2, 3}, {0, 0, 2}, {0, 1, 2}), table({3, 2, 1}, {0, 0, 0}, {0, 2, 2}),
                table({2, 3, 1}, {0, 1, 1}, {0, 1, 2})};
    return zeroRunSeeded(K + 1, minK, K);
}

// Kucherov et al. 2014 / Lam et al. 2009 style scheme with K+1 parts
inline Scheme kucherov_k1(size_t minK, size_t K) {
    if (minK == 0 && K == 0) return backtracking(0, 0);
    if (minK == 0 && K == 1) return {table({1, 2}, {0, 0}, {0, 1}), table({2, 1}, {0, 0}, {0, 1})};
    if (minK == 0 && K == 2)
        return {table({1, 2, 3}, {0, 0, 0}, {0, 2, 2}), table({3, 2, 1}, {0, 0, 0}, {0, 1, 2}),
                table({2, 3, 1}, {0, 0, 1}, {0, 1, 2})};
    return zeroRunSeeded(K + 1, minK, K);
}

// Lam et al. 2009 (2BWT): K+1 parts; the published cases are K = 1 (forward + backward search) and K = 2 (forward,
// backward and one bidirectional search that starts in the middle part).  Beyond K = 2 the paper gives no scheme:
// constructive fallback with K+1 parts.
inline Scheme lam(size_t minK, size_t K) {
    if (minK == 0 && K == 0) return backtracking(0, 0);
    if (minK == 0 && K == 1) return {table({1, 2}, {0, 0}, {0, 1}), table({2, 1}, {0, 0}, {0, 1})};
    if (minK == 0 && K == 2)
        return {table({1, 2, 3}, {0, 0, 0}, {0, 2, 2}), table({3, 2, 1}, {0, 0, 0}, {0, 1, 2}),
                table({2, 1, 3}, {0, 0, 1}, {0, 1, 2})};
    return zeroRunSeeded(K + 1, minK, K);
}

// PEX (Navarro & Baeza-Yates, hierarchical verification) written as a search scheme.  The pattern is cut into K+1
// leaves; an inner node with children of budgets kl and kr has the budget kl + kr + 1, so a node that matches within
// its budget has a child that matches within the child's.  One search per leaf: it starts with the leaf (0 errors) and
// climbs the tree, at every ancestor covering the sibling subtree (rightwards when the sibling is on the right, leftwards
// otherwise) under the ancestor's budget.  `lower`: a search that arrives from a right child only has to find what the
// left child's searches cannot, i.e. occurrences whose left sibling exceeds its budget (cumulated lower bound kl + 1).
struct PexNode {
    size_t lo, hi;     // leaves [lo, hi)
    size_t budget;
    int left{-1}, right{-1}, parent{-1};
};
inline std::vector<PexNode> pexTreeTopDown(size_t K) {
    std::vector<PexNode> t;
    std::function<int(size_t, size_t, size_t, int)> make = [&](size_t lo, size_t hi, size_t k, int parent) {
        int id = static_cast<int>(t.size());
        t.push_back(PexNode{lo, hi, k, -1, -1, parent});
        if (hi - lo > 1) {  // k + 1 leaves below a node of budget k
            size_t kl = k / 2;  // the left child gets ceil((k + 1) / 2) leaves
            size_t kr = k - 1 - kl;
            int l = make(lo, lo + kl + 1, kl, id);
            int r = make(lo + kl + 1, hi, kr, id);
            t[id].left = l;
            t[id].right = r;
        }
        return id;
    };
    make(0, K + 1, K, -1);
    return t;
}
inline std::vector<PexNode> pexTreeBottomUp(size_t K) {
    std::vector<PexNode> t;
    std::vector<int> level;
    for (size_t i = 0; i <= K; ++i) {
        level.push_back(static_cast<int>(t.size()));
        t.push_back(PexNode{i, i + 1, 0, -1, -1, -1});
    }
    while (level.size() > 1) {  // pair neighbours; an odd node at the end moves up unchanged
        std::vector<int> next;
        for (size_t i = 0; i + 1 < level.size(); i += 2) {
            int l = level[i], r = level[i + 1];
            int id = static_cast<int>(t.size());
            t.push_back(PexNode{t[l].lo, t[r].hi, t[l].budget + t[r].budget + 1, l, r, -1});
            t[l].parent = id;
            t[r].parent = id;
            next.push_back(id);
        }
        if (level.size() % 2) next.push_back(level.back());
        level = next;
    }
    return t;
}
inline Scheme pexScheme(std::vector<PexNode> const& t, size_t minK, size_t K, bool lower) {
    size_t P = K + 1;
    Scheme ss;
    for (size_t leaf = 0; leaf < P; ++leaf) {
        int node = -1;
        for (size_t i = 0; i < t.size(); ++i)
            if (t[i].left < 0 && t[i].lo == leaf) node = static_cast<int>(i);
        Search s{{leaf}, {0}, {0}};
        size_t lb = 0;
        while (t[node].parent >= 0) {
            int par = t[node].parent;
            bool fromRight = t[par].right == node;
            int sib = fromRight ? t[par].left : t[par].right;
            std::vector<size_t> parts;
            if (fromRight) for (size_t j = t[sib].hi; j-- > t[sib].lo;) parts.push_back(j);
            else for (size_t j = t[sib].lo; j < t[sib].hi; ++j) parts.push_back(j);
            for (size_t j = 0; j < parts.size(); ++j) {
                s.pi.push_back(parts[j]);
                s.u.push_back(t[par].budget);
                if (lower && fromRight && j + 1 == parts.size()) lb = std::max(lb, t[sib].budget + 1);
                s.l.push_back(lb);
            }
            node = par;
        }
        ss.push_back(s);
    }
    applyMinK(ss, minK);
    return ss;
}

inline Scheme h2(size_t P, size_t minK, size_t K) {
    // Reconstruction: with the part counts for which an optimum table is known use it, otherwise the
    // constructive zero-run seeded scheme (the upstream h2 tables are not available).
    if (minK == 0 && K == 1 && P == 2) return optimum(0, 1);
    if (minK == 0 && K == 2 && P == 4) return optimum(0, 2);
    if (minK == 0 && K == 2 && P == 3) return kianfar(0, 2);
    if (P == 1) return backtracking(minK, K);
    return zeroRunSeeded(P, minK, K);
}

struct Entry {
    std::string name;
    std::string description;
    std::function<Scheme(int, int, int, int)> generator;  // (minK, maxK, sigma (unused), refLen (unused))
};

inline std::map<std::string, Entry> const& all() {
    static std::map<std::string, Entry> const m = [] {
        std::map<std::string, Entry> r;
        auto add = [&](std::string name, std::string desc, std::function<Scheme(size_t, size_t)> f) {
            r[name] = Entry{name, desc, [f](int minK, int maxK, int, int) {
                                if (minK < 0 || maxK < minK) throw std::runtime_error("invalid error bounds");
                                return f(static_cast<size_t>(minK), static_cast<size_t>(maxK));
                            }};
        };
        add("backtracking", "single part, plain backtracking", backtracking);
        add("optimum", "optimum search schemes (Kianfar et al.), K+2 parts for K=2", optimum);
        add("01*0", "01*0 seeds, K+2 parts", zeroOnesZero_trivial);
        add("01*0_opt", "01*0 seeds merged per start part", zeroOnesZero_opt);
        add("pigeon", "pigeonhole, K+1 parts", pigeon_trivial);
        add("pigeon_opt", "pigeonhole without redundancy", pigeon_opt);
        add("suffix", "suffix filter, K+1 parts", suffixFilter);
        add("h2-k1", "heuristic scheme with K+1 parts (reconstructed)", [](size_t a, size_t b) { return h2(b + 1, a, b); });
        add("h2-k2", "heuristic scheme with K+2 parts (reconstructed)", [](size_t a, size_t b) { return h2(b + 2, a, b); });
        add("h2-k3", "heuristic scheme with K+3 parts (reconstructed)", [](size_t a, size_t b) { return h2(b + 3, a, b); });
        add("kianfar", "Kianfar et al. schemes with K+1 parts", kianfar);
        add("kucherov-k1", "Kucherov et al. schemes with K+1 parts", kucherov_k1);
        add("kucherov-k2", "Kucherov et al. schemes with K+2 parts (constructive)",
            [](size_t a, size_t b) { return zeroRunSeeded(b + 2, a, b); });
        add("lam", "Lam et al. 2009, K+1 parts (published for K <= 2, constructive beyond)", lam);
        // upstream ships tables computed by the Hato tool (Renders et al. 2024, ILP + greedy search); they are not
        // available offline: constructive scheme with K+2 parts; real tables come in through --scheme-file
        add("hato", "Hato-designed schemes (tables not available: constructive K+2 parts, reconstructed)",
            [](size_t a, size_t b) { return b == 0 ? backtracking(a, 0) : zeroRunSeeded(b + 2, a, b); });
        add("pex-td", "PEX tree built top down, one search per leaf (reconstructed)",
            [](size_t a, size_t b) { return pexScheme(pexTreeTopDown(b), a, b, false); });
        add("pex-td-l", "PEX tree built top down, with lower bounds (reconstructed)",
            [](size_t a, size_t b) { return pexScheme(pexTreeTopDown(b), a, b, true); });
        add("pex-bu", "PEX tree built bottom up, one search per leaf (reconstructed)",
            [](size_t a, size_t b) { return pexScheme(pexTreeBottomUp(b), a, b, false); });
        add("pex-bu-l", "PEX tree built bottom up, with lower bounds (reconstructed)",
            [](size_t a, size_t b) { return pexScheme(pexTreeBottomUp(b), a, b, true); });
        return r;
    }();
    return m;
}

// lookup + generate + verify.  Error text follows /root/reference/src/sahara/search.cpp:181.
inline Scheme generate(std::string const& name, int minK, int maxK) {
    auto const& m = all();
    auto it = m.find(name);
    if (it == m.end()) {
        std::string names;
        for (auto const& [k, v] : m) names += (names.empty() ? "" : ", ") + k;
        throw std::runtime_error("unknown search scheme generetaror \"" + name + "\", valid generators are: " + names);
    }
    Scheme ss = it->second.generator(minK, maxK, 0, 0);
    if (!isValid(ss)) throw std::runtime_error("generator " + name + " produced an invalid search scheme");
    if (!isComplete(ss, static_cast<size_t>(minK), static_cast<size_t>(maxK)))
        throw std::runtime_error("generator " + name + " produced an incomplete search scheme");
    return ss;
}

}  // namespace generator

// ---- Columba text format: one search per line "{pi} {L} {U}", comma separated, 0-based ---------------
inline std::string toColumba(Scheme const& ss) {
    std::ostringstream os;
    auto join = [&](std::vector<size_t> const& v) {
        os << '{';
        for (size_t i = 0; i < v.size(); ++i) os << (i ? "," : "") << v[i];
        os << '}';
    };
    for (auto const& s : ss) {
        join(s.pi);
        os << ' ';
        join(s.l);
        os << ' ';
        join(s.u);
        os << '\n';
    }
    return os.str();
}

inline Scheme fromColumba(std::string const& text) {
    Scheme ss;
    std::istringstream is(text);
    std::string line;
    while (std::getline(is, line)) {
        if (line.find('{') == std::string::npos) continue;
        std::vector<std::vector<size_t>> groups;
        size_t pos = 0;
        while ((pos = line.find('{', pos)) != std::string::npos) {
            size_t end = line.find('}', pos);
            if (end == std::string::npos) throw std::runtime_error("malformed search scheme line: " + line);
            std::vector<size_t> v;
            std::string body = line.substr(pos + 1, end - pos - 1);
            std::replace(body.begin(), body.end(), ',', ' ');
            std::istringstream bs(body);
            size_t x;
            while (bs >> x) v.push_back(x);
            groups.push_back(v);
            pos = end + 1;
        }
        if (groups.size() != 3) throw std::runtime_error("malformed search scheme line: " + line);
        ss.push_back(Search{groups[0], groups[1], groups[2]});
    }
    if (!isValid(ss)) throw std::runtime_error("search scheme file holds an invalid scheme");
    return ss;
}

}  // namespace sahara::scheme
