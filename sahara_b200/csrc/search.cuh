// search.cuh — kernel 2: search-scheme backtracking over the bidirectional FM-index.
//
// Replaces fmc::search_ng24::search<Edit>(index, queries, scheme, delegate) as called at
// /root/reference/src/sahara/search.cpp:227-231 (semantics: SURVEY.md §9.4).
//
// Three kernels share one state machine:
//
//   fm_roots_kernel   makes the root frame of every (query, search): the q-gram table entry of the leading error-free
//                     steps, or the whole suffix array.
//   fm_items_kernel   walks the search trees while a cursor still covers several suffix-array rows.  A lane owns one
//                     query at a time and explores each search depth first with an explicit stack.  One loop
//                     iteration = ONE probe of one BWT (the rows lb and lb+len; a single 32-byte request when they
//                     share a block) from which the child cursors of all symbols are derived; then every search
//                     state that lives on this cursor is expanded (fm_node).
//   text_pool_kernel  takes over as soon as a cursor holds a single row ("seed"): its occurrence T[a, b) is unique,
//                     so whether a child cursor is empty is decided by one text symbol (T[a-1] or T[b]) instead of
//                     two rank probes (in-text verification).  The reported cursor is lb = ISA[a], len = 1.
//   fm_ordered_kernel the same states in the ORDER of the reference recursion, for search_n (--max_hits).
//
// The rules that are reconstructions of upstream behaviour (which operation may follow which, what may be reported,
// the order of the children) are NOT written into these bodies: they come from the policy table
// include/sahara_policy.h (SearchParams::pol), which the CPU oracle consumes as well.
//
// States.  A frame is a cursor plus a search state (step, e, LInfo, RInfo) — LInfo/RInfo = last operation at the
// left/right end of the match, as in the reference recursion.  The deletion (step, e+1, D) and the substitution
// (step+1, e+1, S) of a mismatching symbol share the child cursor: one PAIR frame.  The insertion child keeps the
// cursor of its parent and is expanded in the same iteration (insertion chain).  Every state is expanded exactly
// like the corresponding call of the reference, so the reported multiset of (query, cursor, errors) is the
// reference's; only the order differs (the reference does not depend on it, search.cpp:218-220).  The number of
// expanded states equals the number of cursor extensions of the reference ("nodes").
//
// Children are pushed match first, so the error subtrees of a node finish before its match child resumes: a
// path keeps siblings only at the nodes it left through an error edge, which bounds the stack (capi.cu).
#pragma once
#include <cstdio>
#include "../../include/sahara_policy.h"
#include "layout.cuh"

namespace sb200 {

// packed step entry: pi | l << 16 | u << 20 | right << 24
__host__ __device__ inline uint32_t pack_step(uint32_t pi, uint32_t l, uint32_t u, bool right) {
    return pi | (l << 16) | (u << 20) | (static_cast<uint32_t>(right) << 24);
}

enum : uint32_t { INFO_M = 0, INFO_S = 1, INFO_I = 2, INFO_D = 3 };

// frame meta: step (10 bits) | e << 10 (4 bits) | LInfo << 14 | RInfo << 16 | PAIR << 18 | tlen << 20 (10 bits)
// tlen = number of text symbols the match covers so far (matches, substitutions, deletions)
__host__ __device__ inline uint32_t pack_meta(uint32_t step, uint32_t e, uint32_t L, uint32_t R) {
    return step | (e << 10) | (L << 14) | (R << 16);
}
constexpr uint32_t META_PAIR = 1u << 18;
constexpr uint32_t META_STREAK = 1u << 19;  // (text_pool_kernel) the state was reached by a match that followed a match
constexpr uint32_t META_TLEN_SHIFT = 20;
#if !defined(SB200_EMIT_CHUNK)
#define SB200_EMIT_CHUNK 16
#endif
constexpr uint32_t kEmitChunk = SB200_EMIT_CHUNK;   // output slots a thread reserves per atomic
constexpr uint32_t kInvalidQid = 0xffffffffu;
constexpr uint32_t kRunE = 5;         // error levels 0..4 in the run table
constexpr uint32_t kCursorTextPosFlag = 0x10u;  // == kCursorTextPos of locate.cuh
// rare paths are kept out of line: text_pool_kernel is as sensitive to the size of its code (instruction fetch) as to the
// instructions it executes
#if defined(SB200_HOST_EMU) || !defined(__CUDACC__)
#define SB200_COLD inline
#else
#define SB200_COLD __noinline__
#endif

// Match-only runs.  At (step, e) with u[step] == e and l[step] <= e only a match is possible (no error may be
// added) — in the text kernel such steps are compared symbol by symbol without touching the stack.  run(step, e)
// = number of consecutive such steps that extend the same end over consecutive query positions.
inline void build_runs(uint32_t n_searches, uint32_t len, const uint32_t* steps, uint8_t* runs) {
    for (uint32_t j = 0; j < n_searches; ++j)
        for (uint32_t e = 0; e < kRunE; ++e)
            for (uint32_t i = len; i-- > 0;) {
                const uint32_t st = steps[j * len + i];
                const uint32_t l = (st >> 16) & 0xfu, u = (st >> 20) & 0xfu, right = (st >> 24) & 1u, pi = st & 0xffffu;
                uint32_t r = 0;
                if (u == e && l <= e) {
                    r = 1;
                    if (i + 1 < len) {
                        const uint32_t sn = steps[j * len + i + 1];
                        const uint32_t pn = sn & 0xffffu;
                        bool consecutive = ((sn >> 24) & 1u) == right && (right ? pn == pi + 1 : pn + 1 == pi);
                        if (consecutive) r += runs[((j * len) + i + 1) * kRunE + e];
                    }
                    if (r > 255) r = 255;
                }
                runs[((j * len) + i) * kRunE + e] = static_cast<uint8_t>(r);
            }
}

// State flags.  What a state (step, e) of the in-text verification may do depends on the scheme tables only; the
// conditions are evaluated once per scheme instead of once per expanded state (text_states).  The table sits behind
// the run table: flags(step, e) = runs[state_flags_offset(n_steps) + step * kRunE + e].
enum : uint32_t {
    SF_MATCH = 1,     // l <= e <= u: the matching symbol may be taken
    SF_MISMATCH = 2,  // l <= e + 1 <= u: an error may be added
    SF_M_ALIVE = 4,   // not the last step and the match child survives the next lower bound
    SF_SUB_ALIVE = 8, // not the last step and a child with e + 1 errors survives the next lower bound
    SF_PAIR = 16,     // SF_SUB_ALIVE, the next step extends the same end and the policy lets no insertion follow a D or an S:
                      // deletion + substitution share a frame
    SF_RUN_M = 32,    // the match child (step + 1, e) starts a match-only run
    SF_RUN_D = 64,    // the deletion child (step, e + 1) starts a match-only run
    SF_RUN_S = 128    // the substitution child (step + 1, e + 1) starts a match-only run
};
__host__ __device__ inline uint32_t state_flags_offset(uint32_t n_steps) { return (n_steps * kRunE + 3u) & ~3u; }
__host__ __device__ inline uint32_t run_table_bytes(uint32_t n_steps) { return 3u * state_flags_offset(n_steps); }
// Path windows.  window(step, e) = w > 0: the states (step + j, e), j < w, of a search whose extended end carries M can be
// expanded together by text_path — they extend the same end over consecutive query positions, none of them or of the states
// of their insertion chains is the last step, and the state flags of every error level from e on are the same over the
// window, the chains and the children's first steps.  The table sits behind the state flags:
// window(step, e) = runs[2 * state_flags_offset(n_steps) + step * kRunE + e].
constexpr uint32_t kPathWindow = 8;   // steps per window: the symbols of one packed word
inline void build_path_windows(uint32_t n_searches, uint32_t len, const uint32_t* steps, uint8_t* runs) {
    const uint32_t off = state_flags_offset(n_searches * len);
    const uint8_t* flags = runs + off;
    uint8_t* win = runs + 2u * off;
    uint32_t kmax = 0;
    for (uint32_t i = 0; i < n_searches * len; ++i) kmax = ((steps[i] >> 20) & 0xfu) > kmax ? (steps[i] >> 20) & 0xfu : kmax;
    if (kmax >= kRunE) kmax = kRunE - 1;
    for (uint32_t j = 0; j < n_searches; ++j)
        for (uint32_t e = 0; e < kRunE; ++e) {
            uint32_t same = 0;  // steps behind i that look like step i (directions, positions, flags of the levels e ..)
            for (uint32_t i = len; i-- > 0;) {
                const uint32_t idx = (j * len + i) * kRunE;
                bool like_next = false;
                if (i + 1 < len) {
                    const uint32_t st = steps[j * len + i], sn = steps[j * len + i + 1];
                    const uint32_t right = (st >> 24) & 1u, pi = st & 0xffffu, pn = sn & 0xffffu;
                    like_next = ((sn >> 24) & 1u) == right && (right ? pn == pi + 1 : pn + 1 == pi);
                    for (uint32_t ee = e; ee < kRunE && like_next; ++ee) like_next = flags[idx + ee] == flags[idx + kRunE + ee];
                }
                same = like_next ? same + 1 : 0;
                const uint32_t f = flags[idx + e];
                const bool walk = e <= kmax && runs[idx + e] == 0 && (f & SF_MATCH) && (f & SF_M_ALIVE) && !(f & SF_RUN_M);
                // behind the window: the insertion chains (kmax - e steps) and the first step of their children
                const uint32_t margin = kmax - (e <= kmax ? e : kmax) + 1u;
                uint32_t w = walk && same > margin ? same - margin : 0u;
                win[idx + e] = static_cast<uint8_t>(w > kPathWindow ? kPathWindow : w);
            }
        }
}
inline void build_state_flags(uint32_t n_searches, uint32_t len, const uint32_t* steps, uint8_t* runs, const sb200_policy& pol) {
    const bool pairs = sb200_pol_pairs(&pol) != 0;  // (a PAIR frame expands no insertion child of either half)
    uint8_t* flags = runs + state_flags_offset(n_searches * len);
    for (uint32_t j = 0; j < n_searches; ++j)
        for (uint32_t i = 0; i < len; ++i)
            for (uint32_t e = 0; e < kRunE; ++e) {
                const uint32_t st = steps[j * len + i];
                const uint32_t l = (st >> 16) & 0xfu, u = (st >> 20) & 0xfu, right = (st >> 24) & 1u;
                const bool last = i + 1 == len;
                const uint32_t sn = last ? 0u : steps[j * len + i + 1];
                const uint32_t lnext = (sn >> 16) & 0xfu;
                auto run = [&](uint32_t ii, uint32_t ee) { return ii < len && ee < kRunE && runs[((j * len) + ii) * kRunE + ee] != 0; };
                uint32_t f = 0;
                if (l <= e && e <= u) f |= SF_MATCH;
                if (l <= e + 1 && e + 1 <= u) f |= SF_MISMATCH;
                if (!last && lnext <= e + 1) f |= SF_M_ALIVE;
                if (!last && lnext <= e + 2) f |= SF_SUB_ALIVE;
                if (pairs && !last && lnext <= e + 2 && ((sn >> 24) & 1u) == right) f |= SF_PAIR;
                if (run(i + 1, e)) f |= SF_RUN_M;
                if (run(i, e + 1)) f |= SF_RUN_D;
                if (run(i + 1, e + 1)) f |= SF_RUN_S;
                flags[((j * len) + i) * kRunE + e] = static_cast<uint8_t>(f);
            }
}

// counters (unsigned long long each)
enum : int {
    CT_NEXT_QUERY = 0,   // work distribution of fm_ordered_kernel
    CT_OUT_SLOTS = 1,    // cursor output slots reserved
    CT_NODES = 2,        // states expanded
    CT_OVERFLOW = 3,     // stack overflow flag
    CT_LF_STEPS = 4,     // (locate kernel)
    CT_MAX_SP = 5,       // deepest stack seen
    CT_CURSORS = 6,      // cursors reported
    CT_SEED_SLOTS = 7,   // seed slots reserved by fm_items_kernel
    CT_NEXT_SEED = 8,    // work distribution of text_pool_kernel
    CT_SEEDS = 9,        // seeds produced
    CT_BAD_QUERY = 10,   // 1 + offset of a query symbol outside the alphabet (0 = none)
    CT_NODES_TEXT = 11,  // states expanded by text_pool_kernel (subset of CT_NODES)
    CT_NEXT_ITEM = 12,   // work distribution of fm_items_kernel
    CT_TOTAL_ROWS = 13,  // (locate) rows beyond the first of every cursor (overflow guard of the u32 hit scan)
    CT_COUNT = 16
};

struct SearchParams {
    OccTable bwt, bwtRev;
    uint32_t C[8];
    uint32_t n_rows;
    const uint32_t* packed;  // [n_queries][packed_words(len)] queries, 8 symbols per word
    uint32_t n_queries, len, n_searches;
    const uint32_t* steps;   // [n_searches][len] packed
    const uint8_t* runs;     // [n_searches][len][kRunE]: length of the match-only run that starts at (step, e)
    uint4* out;              // (qid, lb, len, e)
    uint32_t out_cap;
    unsigned long long* counters;
    const uint4* qgram;      // optional q-gram jump table [4^q] (lb, lbRev, len, 0)
    uint32_t qgram_q;
    uint32_t debug_flags;    // 1: no pair frames, 2: no insertion chains in fm_node (diagnostics only)
    sb200_policy pol;        // the reconstructed rules of the recursion (include/sahara_policy.h)
    // in-text verification (nullptr = off): suffix array, its inverse, the text packed 8 symbols per word
    const uint32_t* sa32;
    const uint32_t* isa32;
    const uint32_t* text4;
    uint4* seeds;            // (qid, lb, search, meta) handed from fm_items_kernel to text_pool_kernel
    uint32_t seed_cap;
    // work items of fm_items_kernel, n_searches slots per query: root frame (lb, lbRev, len, meta) and
    // (qid, search | toText << 8 | live slots of the query << 16)
    uint4* items;
    uint2* item_tags;
    uint32_t textpos_out;    // 1: verified occurrences are reported as (qid, text position, 1, e | kCursorTextPos) for the locate step
    // ordered walk with a hit limit (fm_ordered_kernel): rows per query, the threads' stacks
    uint32_t max_hits;
    uint4* ostack;
    uint32_t ostack_frames;  // frames per thread
    const uint32_t* redo;    // optional: the queries to walk (n_queries = their number); nullptr = all queries
    uint32_t* qcount;        // optional: rows reported per query, counted as the cursors are written (what hit_count_kernel of
                             // locate.cuh would count in a pass of its own); rows beyond the first of a cursor also go to CT_TOTAL_ROWS
};

__host__ __device__ inline uint32_t packed_words(uint32_t len) { return (len + 7) / 8; }

#if !defined(SB200_HOST_EMU)
// the 8 bytes that start at byte `addr` of buffer q (total bytes; q 4-byte aligned): three aligned word loads instead
// of eight byte loads; bytes behind the buffer read as 0
__device__ __forceinline__ uint2 load8_aligned(const uint8_t* q, uint64_t addr, uint64_t total) {
    const uint64_t a0 = addr & ~uint64_t{3};
    const uint32_t sh = static_cast<uint32_t>(addr & 3u) * 8u;
    const uint32_t* p = reinterpret_cast<const uint32_t*>(q + a0);
    const uint64_t end = (total + 3u) & ~uint64_t{3};
    const uint32_t w0 = p[0];
    const uint32_t w1 = a0 + 4 < end ? p[1] : 0u;
    const uint32_t w2 = (sh != 0 && a0 + 8 < end) ? p[2] : 0u;
    return make_uint2(__funnelshift_r(w0, w1, sh), __funnelshift_r(w1, w2, sh));
}
__device__ __forceinline__ uint2 load8_bytes(const uint8_t* q, uint64_t addr, uint64_t total) {
    uint32_t lo = 0, hi = 0;
    for (uint32_t k = 0; k < 4; ++k) {
        if (addr + k < total) lo |= static_cast<uint32_t>(q[addr + k]) << (8 * k);
        if (addr + 4 + k < total) hi |= static_cast<uint32_t>(q[addr + 4 + k]) << (8 * k);
    }
    return make_uint2(lo, hi);
}
// low nibbles of 4 bytes -> 16 bits (byte k in nibble k)
__device__ __forceinline__ uint32_t nibbles4(uint32_t x) {
    x &= 0x0f0f0f0fu;
    x = (x | (x >> 4)) & 0x00ff00ffu;
    return (x | (x >> 8)) & 0xffffu;
}
// packs the n (<= 8) symbols in bytes (lo, hi) into one word (unused nibbles 0xF) and reports symbols >= sigma
// (offset of symbol 0 in the query array = first)
__device__ __forceinline__ uint32_t pack8(uint32_t lo, uint32_t hi, uint32_t n, uint32_t sigma, uint64_t first, unsigned long long* counters) {
    const uint32_t lim = sigma * 0x01010101u;
    uint32_t badLo = __vcmpgeu4(lo, lim), badHi = __vcmpgeu4(hi, lim);
    if (n < 4) badLo &= (1u << (8 * n)) - 1u;
    if (n < 8) badHi &= n > 4 ? (1u << (8 * (n - 4))) - 1u : 0u;
    if (badLo | badHi) {  // verify_rank (/root/reference/src/sahara/search.cpp:118-120): the offset of the last bad symbol
        const uint32_t k = badHi ? 4u + (31u - static_cast<uint32_t>(__clz(static_cast<int>(badHi)))) / 8u
                                 : (31u - static_cast<uint32_t>(__clz(static_cast<int>(badLo)))) / 8u;
        atomicMax(&counters[CT_BAD_QUERY], static_cast<unsigned long long>(first + k + 1));
    }
    uint32_t v = nibbles4(lo) | (nibbles4(hi) << 16);
    if (n < 8) v |= ~((1u << (4 * n)) - 1u);
    return v;
}

// packs the queries 8 symbols per word and verifies them (verify_rank of /root/reference/src/sahara/search.cpp:118-120)
__global__ void pack_queries_kernel(const uint8_t* q, uint64_t n_queries, uint32_t len, uint32_t sigma, uint32_t* out,
                                    unsigned long long* counters) {
    uint32_t W = packed_words(len);
    uint64_t i = blockIdx.x * static_cast<uint64_t>(blockDim.x) + threadIdx.x;
    if (i >= n_queries * W) return;
    uint64_t qi = i / W;
    uint32_t w = static_cast<uint32_t>(i % W);
    const uint64_t addr = qi * len + w * 8, total = n_queries * len;
    const bool aligned = (reinterpret_cast<uintptr_t>(q) & 3u) == 0;
    const uint2 b = aligned ? load8_aligned(q, addr, total) : load8_bytes(q, addr, total);
    const uint32_t n = len - w * 8 < 8u ? len - w * 8 : 8u;
    out[i] = pack8(b.x, b.y, n, sigma, addr, counters);
}

// the same from the reads alone: query 2r = read r, query 2r + 1 = its reverse complement (A<->T, C<->G on ranks 1..4,
// others unchanged; /root/reference/src/sahara/search.cpp:121-123) — packed directly, no byte copy of the second strand
__global__ void pack_reads_kernel(const uint8_t* reads, uint64_t n_queries, uint32_t len, uint32_t sigma, uint32_t* out,
                                  unsigned long long* counters) {
    uint32_t W = packed_words(len);
    uint64_t i = blockIdx.x * static_cast<uint64_t>(blockDim.x) + threadIdx.x;
    if (i >= n_queries * W) return;
    uint64_t qi = i / W;
    uint32_t w = static_cast<uint32_t>(i % W);
    const bool rc = qi & 1u;
    const uint32_t n = len - w * 8 < 8u ? len - w * 8 : 8u;
    const uint64_t total = ((n_queries + 1) >> 1) * len;
    // symbols w*8 .. w*8+n-1 of the query: of the read itself, or (reverse strand) the n bytes that end at len-1-w*8, reversed
    const uint64_t addr = (qi >> 1) * len + (rc ? len - w * 8 - n : w * 8);
    const bool aligned = (reinterpret_cast<uintptr_t>(reads) & 3u) == 0;
    uint2 b = aligned ? load8_aligned(reads, addr, total) : load8_bytes(reads, addr, total);
    if (rc) {
        // byte j of the reversed 8 bytes = byte 7-j; the n valid ones move down to byte 0
        unsigned long long x = (static_cast<unsigned long long>(__byte_perm(b.x, 0, 0x0123)) << 32) | __byte_perm(b.y, 0, 0x0123);
        x >>= 8 * (8 - n);
        b = make_uint2(static_cast<uint32_t>(x), static_cast<uint32_t>(x >> 32));
        // complement of ranks 1..4
        const uint32_t mLo = __vcmpgeu4(b.x, 0x01010101u) & __vcmpleu4(b.x, 0x04040404u);
        const uint32_t mHi = __vcmpgeu4(b.y, 0x01010101u) & __vcmpleu4(b.y, 0x04040404u);
        b.x = (b.x & ~mLo) | (__vsub4(0x05050505u, b.x) & mLo);
        b.y = (b.y & ~mHi) | (__vsub4(0x05050505u, b.y) & mHi);
    }
    out[i] = pack8(b.x, b.y, n, sigma, qi * len + w * 8, counters);
}

// the same from reads the host packed already (SB200_READS_PACKED4: 4 bits per base, 8 bases per little-endian word,
// packed_words(len) words per read — the layout of `out`): half the PCIe bytes of the rank bytes.  Query qi is read qi,
// or with with_reverse read qi / 2 and, for odd qi, its reverse complement.  Nibbles behind the read are forced to 0xF.
__global__ void pack_packed4_kernel(const uint32_t* reads, uint64_t n_queries, uint32_t len, uint32_t sigma, uint32_t with_reverse,
                                    uint32_t* out, unsigned long long* counters) {
    const uint32_t W = packed_words(len);
    const uint64_t i = blockIdx.x * static_cast<uint64_t>(blockDim.x) + threadIdx.x;
    if (i >= n_queries * W) return;
    const uint64_t qi = i / W;
    const uint32_t w = static_cast<uint32_t>(i % W);
    const bool rc = with_reverse && (qi & 1u);
    const uint32_t* src = reads + (with_reverse ? (qi >> 1) : qi) * W;
    const uint32_t n = len - w * 8 < 8u ? len - w * 8 : 8u;
    uint32_t v;
    if (!rc) {
        v = src[w];
    } else {
        // symbols w*8 .. w*8+n-1 of the reverse strand = symbols p_hi .. p_hi-n+1 of the read, complemented
        const uint32_t p_hi = len - 1 - w * 8, p_lo = p_hi + 1 - n;
        const uint32_t w0 = p_lo >> 3;
        const uint32_t lo = src[w0], hi = w0 + 1 < W ? src[w0 + 1] : 0u;
        uint32_t x = __funnelshift_r(lo, hi, (p_lo & 7u) * 4u);  // nibble j = symbol p_lo + j
        // reverse the 8 nibbles, then move the n valid ones (now at the top) down
        x = ((x & 0x0f0f0f0fu) << 4) | ((x >> 4) & 0x0f0f0f0fu);
        x = __byte_perm(x, 0, 0x0123);
        x >>= 4u * (8u - n);
        // complement of ranks 1..4 (nibble-wise through the two byte lanes)
        uint32_t ev = x & 0x0f0f0f0fu, od = (x >> 4) & 0x0f0f0f0fu;
        const uint32_t me = __vcmpgeu4(ev, 0x01010101u) & __vcmpleu4(ev, 0x04040404u);
        const uint32_t mo = __vcmpgeu4(od, 0x01010101u) & __vcmpleu4(od, 0x04040404u);
        ev = (ev & ~me) | (__vsub4(0x05050505u, ev) & me);
        od = (od & ~mo) | (__vsub4(0x05050505u, od) & mo);
        v = ev | (od << 4);
    }
    // verify_rank (/root/reference/src/sahara/search.cpp:118-120): the offset of the last symbol >= sigma
    const uint32_t ev = v & 0x0f0f0f0fu, od = (v >> 4) & 0x0f0f0f0fu;
    const uint32_t lim = sigma * 0x01010101u;
    uint32_t bad = 0;  // bit j: symbol j is outside the alphabet
    const uint32_t be = __vcmpgeu4(ev, lim), bo = __vcmpgeu4(od, lim);
    for (uint32_t j = 0; j < 4; ++j) {
        bad |= ((be >> (8 * j)) & 1u) << (2 * j);
        bad |= ((bo >> (8 * j)) & 1u) << (2 * j + 1);
    }
    if (n < 8) bad &= (1u << n) - 1u;
    if (bad) atomicMax(&counters[CT_BAD_QUERY], static_cast<unsigned long long>(qi * len + w * 8 + (31u - static_cast<uint32_t>(__clz(static_cast<int>(bad)))) + 1));
    if (n < 8) v |= ~((1u << (4 * n)) - 1u);
    out[i] = v;
}
#endif

#if !defined(SB200_HOST_EMU)
// the same from reads of 2 bits per base (A, C, G, T = 0 .. 3; 16 bases per little-endian word, (len + 15) / 16 words per
// read): half the bytes of the 4-bit form over PCIe.  Nothing to verify — every code is a symbol of the alphabet.
__global__ void pack_packed2_kernel(const uint32_t* reads, uint64_t n_queries, uint32_t len, uint32_t with_reverse, uint32_t* out) {
    const uint32_t W = packed_words(len), W2 = (len + 15u) / 16u;
    const uint64_t i = blockIdx.x * static_cast<uint64_t>(blockDim.x) + threadIdx.x;
    if (i >= n_queries * W) return;
    const uint64_t qi = i / W;
    const uint32_t w = static_cast<uint32_t>(i % W);
    const bool rc = with_reverse && (qi & 1u);
    const uint32_t* src = reads + (with_reverse ? (qi >> 1) : qi) * W2;
    const uint32_t n = len - w * 8 < 8u ? len - w * 8 : 8u;
    // the 8 codes that start at symbol p of the read (forward: w * 8; reverse strand: the n symbols that end at len - 1 - w * 8)
    const uint32_t p = rc ? len - w * 8 - n : w * 8;
    const uint32_t w0 = p >> 4;
    const uint32_t lo = src[w0], hi = w0 + 1 < W2 ? src[w0 + 1] : 0u;
    uint32_t x = __funnelshift_r(lo, hi, (p & 15u) * 2u) & 0xffffu;
    if (rc) x ^= 0xffffu;  // complement: A <-> T, C <-> G = 3 - code
    // 8 codes of 2 bits -> 8 nibbles
    x = (x | (x << 8)) & 0x00ff00ffu;
    x = (x | (x << 4)) & 0x0f0f0f0fu;
    x = (x | (x << 2)) & 0x33333333u;
    if (rc) {  // reverse the nibbles, then move the n valid ones (now at the top) down
        x = ((x & 0x0f0f0f0fu) << 4) | ((x >> 4) & 0x0f0f0f0fu);
        x = __byte_perm(x, 0, 0x0123);
        x >>= 4u * (8u - n);
    }
    uint32_t v = x + 0x11111111u;  // ranks 1 .. 4
    if (n < 8) v |= ~((1u << (4 * n)) - 1u);
    out[i] = v;
}
#endif

// chunked append to a global array: one atomic per CHUNK entries
template <uint32_t CHUNK>
struct ChunkWriterT {
    uint32_t pos{0}, end{0};
    __device__ __forceinline__ void put(uint4* buf, uint32_t cap, unsigned long long* counter, uint4 v) {
        if (pos == end) {
            pos = static_cast<uint32_t>(atomicAdd(counter, static_cast<unsigned long long>(CHUNK)));
            end = pos + CHUNK;
        }
        if (pos < cap) buf[pos] = v;
        ++pos;
    }
    // unused slots of the last chunk become empty entries
    __device__ __forceinline__ void finish(uint4* buf, uint32_t cap) {
#pragma unroll 1
        for (; pos < end; ++pos)
            if (pos < cap) buf[pos] = make_uint4(kInvalidQid, 0, 0, 0);
    }
};
using ChunkWriter = ChunkWriterT<kEmitChunk>;
#if !defined(SB200_SEED_CHUNK)
#define SB200_SEED_CHUNK 64
#endif
// seed slots a thread of the walk reserves per atomic.  All threads add to ONE counter, and same-address atomics are served one
// after the other: with 16 slots per atomic the reservations of fm_items_kernel took as long as its probes (45 % of its stall
// samples; 16 / 32 / 64 / 128 slots: 0.96 / 0.77 / 0.73 / 0.75 ms for the walk of the headline workload); the unused slots of a
// thread's last chunk are empty entries that text_pool_kernel skips
constexpr uint32_t kSeedChunk = SB200_SEED_CHUNK;

// ================================================================================================
// Shared pieces of the FM-index walk.
// ================================================================================================
#if defined(SB200_HOST_EMU)
// the emulated "warp" has one lane
static inline uint32_t warp_ballot(bool p) { return p ? 1u : 0u; }
static inline uint32_t warp_bcast0(uint32_t v) { return v; }
static inline uint32_t warp_lane() { return 0u; }
static inline uint32_t popc32(uint32_t v) { return static_cast<uint32_t>(__builtin_popcount(v)); }
static inline void warp_counter_add(unsigned long long* counter, uint32_t v) { *counter += v; }
static inline void warp_counter_max(unsigned long long* counter, uint32_t v) { if (v > *counter) *counter = v; }
static inline uint32_t ldg32(const uint32_t* p) { return *p; }
static inline uint4 ldg128(const uint4* p) { return *p; }
static inline uint2 ldg64(const uint2* p) { return *p; }
#else
__device__ __forceinline__ uint32_t warp_ballot(bool p) { return __ballot_sync(0xffffffffu, p); }
__device__ __forceinline__ uint32_t warp_bcast0(uint32_t v) { return __shfl_sync(0xffffffffu, v, 0); }
__device__ __forceinline__ uint32_t warp_lane() { return threadIdx.x & 31u; }
__device__ __forceinline__ uint32_t popc32(uint32_t v) { return static_cast<uint32_t>(__popc(v)); }
// the lanes of a converged warp add their values to one counter: reduced over the warp first, one atomic per warp (all
// warps of a kernel end at about the same time, and same-address atomics are served one after the other)
__device__ __forceinline__ void warp_counter_add(unsigned long long* counter, uint32_t v) {
    const uint32_t sum = __reduce_add_sync(0xffffffffu, v);
    if ((threadIdx.x & 31u) == 0 && sum != 0) atomicAdd(counter, static_cast<unsigned long long>(sum));
}
__device__ __forceinline__ void warp_counter_max(unsigned long long* counter, uint32_t v) {
    const uint32_t mx = __reduce_max_sync(0xffffffffu, v);
    if ((threadIdx.x & 31u) == 0 && mx != 0) atomicMax(counter, static_cast<unsigned long long>(mx));
}
__device__ __forceinline__ uint32_t ldg32(const uint32_t* p) { return __ldg(p); }
__device__ __forceinline__ uint4 ldg128(const uint4* p) { return __ldg(p); }
__device__ __forceinline__ uint2 ldg64(const uint2* p) { return __ldg(p); }
#endif

// Root frame of one search of one query.  tbl: the steps of the search, qsym(pos): query symbol.
// Returns 0 = the search cannot start (dead), 1 = `root` is the frame to expand, 2 = the q-gram covers the
// whole query: `root` is the final cursor (lb, -, len, -) with 0 errors.
template <typename QSym>
__device__ __forceinline__ int fm_root(const SearchParams& P, const uint32_t* tbl, QSym&& qsym, uint4& root) {
    const uint32_t qlen = P.len;
    const uint32_t st0 = tbl[0];
    if (((st0 >> 16) & 0xfu) > 1) return 0;  // neither a match nor a mismatch allowed at step 0
    root = make_uint4(0, 0, P.n_rows, 0);
    if (P.qgram_q) {  // q-gram jump: skip the leading steps that allow no error
        const uint32_t qq = P.qgram_q;
        bool ok = qq <= qlen;
        uint32_t code = 0;
        const bool right0 = (st0 >> 24) & 1u;
        for (uint32_t i = 0; ok && i < qq; ++i) {
            const uint32_t st = tbl[i];
            const uint32_t c = qsym(st & 0xffffu);
            ok = ((st >> 20) & 0xfu) == 0 && c >= 1 && c <= 4 && (((st >> 24) & 1u) == right0);
            // the table is keyed by the string in text order, first symbol most significant
            if (right0) code = (code << 2) | (c - 1);
            else code |= (c - 1) << (2 * i);
        }
        if (ok) {
            const uint4 g = P.qgram[code];
            if (g.z == 0) return 0;
            if (qq == qlen) {
                root = g;
                return 2;
            }
            if (((tbl[qq] >> 16) & 0xfu) > 1) return 0;
            root = make_uint4(g.x, g.y, g.z, pack_meta(qq, 0, INFO_M, INFO_M) | (qq << META_TLEN_SHIFT));
        }
    }
    return 1;
}

// One node of the walk: ONE probe of the occurrence table for the cursor of frame f, then every state that lives
// on this cursor is expanded.  push(lb, lbRev, len, meta) takes a child frame, emit(lb, len, e) a reported cursor.
template <int SIGMA, bool EDIT, typename QSym, typename Push, typename Emit>
__device__ __forceinline__ void fm_node(const SearchParams& P, const uint32_t* tbl, const uint4 f, uint32_t& nodes, QSym&& qsym, Push&& push,
                                        Emit&& emit) {
    const uint32_t qlen = P.len;
    const uint32_t lb = f.x, lbRev = f.y, len = f.z, meta = f.w;
    uint32_t step = meta & 0x3ffu, e = (meta >> 10) & 0xfu;
    uint32_t Linfo = (meta >> 14) & 3u, Rinfo = (meta >> 16) & 3u;
    const bool pair = (meta & META_PAIR) != 0;
    const uint32_t tlen = (meta >> META_TLEN_SHIFT) & 0x3ffu;
    const bool right = (tbl[step] >> 24) & 1u;
    uint32_t c_next = qsym(tbl[step] & 0xffffu);  // query symbol of the first state: requested before the probe
    // child cursors per symbol: (klb[s], klbRev[s], cnt[s]).  The probed side continues at C[s] + rank(lo, s),
    // the other side moves by the number of smaller symbols inside the interval.  (The selection by `right`
    // is done once here, outside the state loop.)
    uint32_t klb[SIGMA], klbRev[SIGMA], cnt[SIGMA];
    {
        const OccTable& tab = right ? P.bwtRev : P.bwt;
        const uint32_t lo = right ? lbRev : lb;
        const uint32_t hi = lo + len;
        OccBlk b1 = load_blk(tab.blk + (lo >> kBlkShift));
        OccSup s1 = load_sup(tab.sup + (lo >> kSupShift));
        OccBlk b2 = b1;
        OccSup s2 = s1;
        if ((lo >> kBlkShift) != (hi >> kBlkShift)) {
            b2 = load_blk(tab.blk + (hi >> kBlkShift));
            if ((lo >> kSupShift) != (hi >> kSupShift)) s2 = load_sup(tab.sup + (hi >> kSupShift));
        }
        uint32_t own[SIGMA];
        uint32_t sum1 = 0, sumc = 0;
#pragma unroll
        for (int s = 1; s < SIGMA; ++s) {
            uint32_t a = s1.c[s] + blk_ctr(b1, s) + blk_count(b1, lo & 63u, s);
            uint32_t b = s2.c[s] + blk_ctr(b2, s) + blk_count(b2, hi & 63u, s);
            own[s] = P.C[s] + a;
            cnt[s] = b - a;
            sum1 += a;
            sumc += b - a;
        }
        own[0] = lo - sum1;  // C[0] == 0
        cnt[0] = len - sumc;
        uint32_t other = right ? lb : lbRev;  // interval start on the side that is not probed
#pragma unroll
        for (int s = 0; s < SIGMA; ++s) {
            klb[s] = right ? other : own[s];
            klbRev[s] = right ? own[s] : other;
            other += cnt[s];
        }
    }
    const uint32_t tlenSame = tlen << META_TLEN_SHIFT, tlenNext = (tlen + 1) << META_TLEN_SHIFT;

    // ---- expand every state that lives on this cursor ------------------------------------------
    bool second = false;  // second half of a pair already taken
    bool first = true;
    while (true) {
        ++nodes;
#if defined(SB200_TRACE)
        if (P.debug_flags & 4u)
            printf("STATE lb=%u lbRev=%u len=%u step=%u e=%u L=%u R=%u pair=%d second=%d right=%d\n", lb, lbRev, len, step, e, Linfo, Rinfo,
                   (int)pair, (int)second, (int)right);
#endif
        const uint32_t st = tbl[step];
        const uint32_t l = (st >> 16) & 0xfu, u = (st >> 20) & 0xfu;
        const uint32_t c = first ? c_next : qsym(st & 0xffffu);
        first = false;
        const bool last = step + 1 == qlen;
        const uint32_t stn = last ? 0u : tbl[step + 1];
        const uint32_t lnext = (stn >> 16) & 0xfu;
        const bool rightNext = (stn >> 24) & 1u;
        const bool sameDirNext = !last && (rightNext == right);
        const bool matchOK = l <= e && e <= u;
        const bool mmOK = l <= e + 1 && e + 1 <= u;
        const uint32_t T = right ? Rinfo : Linfo;
        const uint32_t O = right ? Linfo : Rinfo;          // info of the other end
        // end filter of the policy: may a cursor be reported whose moving end carries M / I / S (the other end: O)?
        const bool otherEndOK = !EDIT || sb200_pol_end1(&P.pol, O);
        const bool endM = otherEndOK && (!EDIT || sb200_pol_end1(&P.pol, INFO_M));
        const bool endI = otherEndOK && (!EDIT || sb200_pol_end1(&P.pol, INFO_I));
        const bool endS = otherEndOK && (!EDIT || sb200_pol_end1(&P.pol, INFO_S));
        // metas of the possible children: the moving side gets the new info
        const uint32_t keepL = right ? Linfo : 0u, keepR = right ? 0u : Rinfo;
        const uint32_t sideShift = right ? 16u : 14u;
        const uint32_t metaBase = (keepL << 14) | (keepR << 16);
        const uint32_t mM = metaBase | (step + 1) | (e << 10) | (INFO_M << sideShift) | tlenNext;
        const uint32_t mD = metaBase | step | ((e + 1) << 10) | (INFO_D << sideShift) | tlenNext;
        const uint32_t mS = metaBase | (step + 1) | ((e + 1) << 10) | (INFO_S << sideShift) | tlenNext;
        const uint32_t mI = metaBase | (step + 1) | ((e + 1) << 10) | (INFO_I << sideShift) | tlenSame;
        // match
        {
            uint32_t mc = 0, nlb = 0, nlbRev = 0;
#pragma unroll
            for (int s = 0; s < SIGMA; ++s)
                if (static_cast<uint32_t>(s) == c) { mc = cnt[s]; nlb = klb[s]; nlbRev = klbRev[s]; }
            const bool alive = matchOK && mc != 0;
            if (alive && last) {
                if (endM) emit(nlb, mc, e);
            } else if (alive && lnext <= e + 1) {
                push(nlb, nlbRev, mc, mM);
            }
        }
        if (mmOK) {
            const bool delOK = EDIT && sb200_pol_del(&P.pol, T);
            const bool subAlive = !last && lnext <= e + 2;
            const bool asPair = delOK && subAlive && sameDirNext && sb200_pol_pairs(&P.pol) && !(P.debug_flags & 1u);
#pragma unroll
            for (int s = 1; s < SIGMA; ++s) {
                const bool live = static_cast<uint32_t>(s) != c && cnt[s] != 0;
                const uint32_t nlb = klb[s], nlbRev = klbRev[s];
                if (live && (asPair || delOK)) push(nlb, nlbRev, cnt[s], asPair ? (mD | META_PAIR) : mD);
                if (live && !asPair && subAlive) push(nlb, nlbRev, cnt[s], mS);
                if (live && last && endS) emit(nlb, cnt[s], e + 1);  // (edit distance: only when the policy reports an S at an end)
            }
        }
        // next state on the same cursor
        if (pair) {
            if (second) break;
            second = true;
            // second half of the pair: the substitution (step + 1, e, side = S)
            step += 1;
            if (right) Rinfo = INFO_S; else Linfo = INFO_S;
            continue;
        }
        const bool insOK = EDIT && mmOK && sb200_pol_ins(&P.pol, T);
        if (!insOK) break;
        if (last) {
            if (endI) emit(lb, len, e + 1);
            break;
        }
        if (lnext > e + 2) break;  // dead at the next step
        if (!sameDirNext || (P.debug_flags & 2u)) {  // direction changes: needs a probe of the other table
            push(lb, lbRev, len, mI);
            break;
        }
        step += 1;
        e += 1;
        if (right) Rinfo = INFO_I; else Linfo = INFO_I;
    }
}

// ================================================================================================
// Item-based walk (fm_roots_kernel + fm_items_kernel).
// In fm_thread a lane that runs out of frames fetches its next query (work atomic, staging, q-gram entries of up
// to n_searches dead searches in a row) while the other 31 lanes of the warp wait at the reconvergence point in
// front of the probe: 12.8 us per warp iteration instead of one memory latency (profiles/r01i_*).  Here
//   * fm_roots_kernel makes the root frames of every query in one fully parallel pass: the live ones of query q
//     sit in items[q * n_searches + 0 .. count) with tags (qid, search | toText << 8 | count << 16);
//   * fm_items_kernel walks them: the loop is warp synchronous; a lane still owns one query at a time (so its
//     seeds stay grouped by query, which the in-text verification needs for its cache hits), but the root it
//     continues with has ALREADY been loaded: the next slot of its query, or slot 0 of the next query of the range
//     the warp claims 32 queries at a time (the next claim is issued before the range runs out).  Query symbols
//     are read through L1 from the packed queries (requested before the probe).  No long-latency operation of the
//     refill is on the critical path of an iteration any more.
// ================================================================================================
constexpr uint32_t kItemClaim = 32;  // queries a warp claims per atomic (fewer when the batch is small: see item_claim)
constexpr uint32_t kItemToText = 0x100u;

// queries per claim: every warp should come back for work at least ~6 times, or the last claims decide the run time
__device__ __forceinline__ uint32_t item_claim(uint32_t n_queries) {
#if defined(SB200_HOST_EMU)
    const uint32_t warps = 1;
#else
    const uint32_t warps = gridDim.x * (blockDim.x >> 5);
#endif
    const uint32_t c = n_queries / (warps * 6u);
    return c < 4u ? 4u : (c > kItemClaim ? kItemClaim : c);
}

// root frames of all searches of query qid -> its slots of P.items / P.item_tags
__device__ __forceinline__ void fm_make_items(const SearchParams& P, const uint32_t* steps, uint32_t qid) {
    const uint32_t W = packed_words(P.len);
    const uint32_t* q = P.packed + static_cast<uint64_t>(qid) * W;
    auto qsym = [&](uint32_t pos) -> uint32_t { return (ldg32(q + (pos >> 3)) >> ((pos & 7u) * 4u)) & 0xfu; };
    bool delim = false;
    for (uint32_t w = 0; w < W; ++w) {
        const uint32_t v = ldg32(q + w);
        delim = delim || (((v - 0x11111111u) & ~v & 0x88888888u) != 0);  // some nibble is 0
    }
    const uint32_t flags = (P.sa32 != nullptr && !delim) ? kItemToText : 0u;  // a query with the delimiter stays on the FM path
    const uint64_t base = static_cast<uint64_t>(qid) * P.n_searches;
    uint32_t count = 0;
    for (uint32_t j = 0; j < P.n_searches; ++j) {
        uint4 root;
        if (fm_root(P, steps + j * P.len, qsym, root) != 1) continue;  // (the host never enables a table with q >= len here)
        P.items[base + count] = root;
        P.item_tags[base + count] = make_uint2(qid, j | flags);
        ++count;
    }
    // the number of live slots goes into every tag of the query (slot 0 is written even when it is 0)
    for (uint32_t j = 0; j < (count ? count : 1u); ++j) {
        const uint32_t y = j < count ? P.item_tags[base + j].y : 0u;
        P.item_tags[base + j] = make_uint2(qid, y | (count << 16));
    }
}

template <int SIGMA, bool EDIT, int STACK>
__device__ __forceinline__ void fm_items_thread(const SearchParams& P, const uint32_t* s_steps) {
    uint4 stack[STACK];
    int sp = 0;
    uint32_t nodes = 0, emitted = 0, seeded = 0;
    ChunkWriter outW;
    ChunkWriterT<kSeedChunk> seedW;
    bool overflow = false;
    int maxsp = 0;
    const uint32_t qlen = P.len;
    const uint32_t W = packed_words(qlen);
    const uint32_t n_queries = P.n_queries;
    const uint32_t lane = warp_lane();
    const uint32_t claim = item_claim(n_queries);

    // the item being walked
    uint32_t qid = 0, search = 0;
    bool toText = false;
    const uint32_t* tbl = s_steps;
    const uint32_t* qwords = P.packed;
    // the root loaded ahead: slot pj of query pq
    uint4 nroot = make_uint4(0, 0, 0, 0);
    uint2 ntag = make_uint2(0, 0);
    uint32_t pq = 0, pj = 0;
    bool have_next = false;
    // warp-uniform: the claimed range of queries, the claim made ahead (valid in lane 0), end of the batch seen
    uint32_t wnext = 0, wend = 0, cnext = 0;
    bool chave = false, exhausted = n_queries == 0;

    auto qsym = [&](uint32_t pos) -> uint32_t { return (ldg32(qwords + (pos >> 3)) >> ((pos & 7u) * 4u)) & 0xfu; };
    auto emit = [&](uint32_t lb, uint32_t len, uint32_t e) {
        outW.put(P.out, P.out_cap, &P.counters[CT_OUT_SLOTS], make_uint4(qid, lb, len, e));
        if (P.qcount != nullptr && len != 0) {
            atomicAdd(&P.qcount[qid], len);
            if (len > 1) atomicAdd(&P.counters[CT_TOTAL_ROWS], static_cast<unsigned long long>(len - 1));
        }
        ++emitted;
    };
    auto push = [&](uint32_t nlb, uint32_t nlbRev, uint32_t nlen, uint32_t m) {
        if (toText && nlen == 1) {
            seedW.put(P.seeds, P.seed_cap, &P.counters[CT_SEED_SLOTS], make_uint4(qid, nlb, search, m));
            ++seeded;
        } else {
            if (sp < STACK) stack[sp] = make_uint4(nlb, nlbRev, nlen, m);
            else overflow = true;
            ++sp;
        }
    };

    while (true) {
        // ---- a lane without frames continues with the root it loaded ahead ---------------------------
        if (sp == 0 && have_next) {
            const uint32_t count = ntag.y >> 16;
            if (pj < count) {
                qid = pq;
                search = ntag.y & 0xffu;
                toText = (ntag.y & kItemToText) != 0;
                tbl = s_steps + search * qlen;
                qwords = P.packed + static_cast<uint64_t>(qid) * W;
                push(nroot.x, nroot.y, nroot.z, nroot.w);
            }
            have_next = false;
            if (pj + 1 < count) {  // the next slot of the same query
                ++pj;
                const uint64_t at = static_cast<uint64_t>(pq) * P.n_searches + pj;
                nroot = ldg128(P.items + at);
                ntag = ldg64(P.item_tags + at);
                have_next = true;
            }
        }
        // ---- lanes without a root loaded ahead take the next queries of the warp's range -------------
        const uint32_t want = exhausted ? 0u : warp_ballot(!have_next);
        if (want != 0) {
            if (wnext == wend) {  // the range is used up: continue with the claim made ahead
                if (!chave && lane == 0) cnext = static_cast<uint32_t>(atomicAdd(&P.counters[CT_NEXT_ITEM], static_cast<unsigned long long>(claim)));
                wnext = warp_bcast0(cnext);
                chave = false;
                if (wnext >= n_queries) {
                    exhausted = true;
                    wnext = wend = 0;
                } else {
                    wend = wnext + claim < n_queries ? wnext + claim : n_queries;
                }
            }
            const uint32_t left = wend - wnext;
            const uint32_t n = popc32(want);
            const uint32_t take = n < left ? n : left;
            const uint32_t rank = popc32(want & ((1u << lane) - 1u));
            if (!have_next && rank < take) {
                pq = wnext + rank;
                pj = 0;
                const uint64_t at = static_cast<uint64_t>(pq) * P.n_searches;
                nroot = ldg128(P.items + at);
                ntag = ldg64(P.item_tags + at);
                have_next = true;
            }
            wnext += take;
            if (!chave && !exhausted && wend - wnext < claim / 2) {  // claim ahead: the result is needed much later
                if (lane == 0) cnext = static_cast<uint32_t>(atomicAdd(&P.counters[CT_NEXT_ITEM], static_cast<unsigned long long>(claim)));
                chave = true;
            }
        }
        if (warp_ballot(sp != 0 || have_next) == 0) break;
        // ---- one node ---------------------------------------------------------------------------
        if (sp != 0) {
            maxsp = sp > maxsp ? sp : maxsp;
            const uint4 f = stack[--sp];
            fm_node<SIGMA, EDIT>(P, tbl, f, nodes, qsym, push, emit);
        }
    }
    outW.finish(P.out, P.out_cap);
    seedW.finish(P.seeds, P.seed_cap);
    // (the warp is converged here: its loop ends on a vote)
    warp_counter_add(&P.counters[CT_NODES], nodes);
    if (overflow) atomicExch(&P.counters[CT_OVERFLOW], 1ull);
    warp_counter_max(&P.counters[CT_MAX_SP], maxsp);
    warp_counter_add(&P.counters[CT_CURSORS], emitted);
    warp_counter_add(&P.counters[CT_SEEDS], seeded);
}

__device__ __forceinline__ uint32_t nib_mask(uint32_t n) { return n >= 8u ? 0xffffffffu : ((1u << (4u * n)) - 1u); }
#if defined(SB200_HOST_EMU)
static inline uint32_t ctz32(uint32_t x) { return static_cast<uint32_t>(__builtin_ctz(x)); }
static inline uint32_t clz32(uint32_t x) { return static_cast<uint32_t>(__builtin_clz(x)); }
static inline uint32_t funnel_r(uint32_t lo, uint32_t hi, uint32_t sh) { return sh ? (lo >> sh) | (hi << (32u - sh)) : lo; }
#else
__device__ __forceinline__ uint32_t ctz32(uint32_t x) { return static_cast<uint32_t>(__ffs(static_cast<int>(x)) - 1); }
__device__ __forceinline__ uint32_t clz32(uint32_t x) { return static_cast<uint32_t>(__clz(static_cast<int>(x))); }
__device__ __forceinline__ uint32_t funnel_r(uint32_t lo, uint32_t hi, uint32_t sh) { return __funnelshift_r(lo, hi, sh); }
#endif
// the 8 text symbols that start at position pos (symbol i in nibble i); text4 is padded by two words
__device__ __forceinline__ uint32_t text8(const uint32_t* text4, uint32_t pos) {
    const uint32_t w = pos >> 3;
    return funnel_r(text4[w], text4[w + 1], (pos & 7u) * 4u);
}

// nibble j of the result = nibble 7 - j of x
__device__ __forceinline__ uint32_t rev8(uint32_t x) {
    x = ((x & 0x0f0f0f0fu) << 4) | ((x >> 4) & 0x0f0f0f0fu);
    x = ((x & 0x00ff00ffu) << 8) | ((x >> 8) & 0x00ff00ffu);
    return (x << 16) | (x >> 16);
}
// bit 4j of the result is set when nibble j of x is not zero
__device__ __forceinline__ uint32_t nz_nibbles(uint32_t x) {
    x |= x >> 1;
    x |= x >> 2;
    return x & 0x11111111u;
}

// ================================================================================================
// Ordered walk with a hit limit (fm_ordered_kernel): fmc::search_ng24::search_n<Edit>(index, queries, scheme, maxHits, cb)
// as called at /root/reference/src/sahara/search.cpp:228,231.  A query ends as soon as maxHits suffix-array rows were
// delivered, the cursor that crosses the limit is cut to its first rows — WHICH hits those are is decided by the order
// of the reference recursion (searches in scheme order; at a node the match child, then the symbols in
// ascending order with the deletion before the substitution, then the insertion).  The other kernels expand the same
// states in another order, so this path has its own walk: one thread per query, an explicit stack (global memory,
// frame i of a thread at stack[i * stride]) onto which the children of a node are pushed in REVERSE recursion order,
// so that they are popped in recursion order.  A frame is one call of the recursion with a non-empty cursor:
// (lb, lbRev, len, step | e | LInfo | RInfo); step == query length reports.  Every popped frame with step < length is
// one cursor extension of the reference ("node").  A cursor that holds one row continues as a text frame when the
// verification tables are loaded (one text symbol instead of two rank probes per node), and a search starts from the
// q-gram table like the other walks.  A query that reaches the limit ends early by definition.  That is also how the work is split (capi.cu, search_only): the
// throughput kernels search everything first; a query with at most maxHits rows is complete and identical to its
// search_n result, only the queries with MORE rows (`redo` list, found by the kernels below) are walked again in
// order, and their cursors from the first pass are dropped.
// ================================================================================================
template <int SIGMA>
__device__ __forceinline__ void probe_children(const SearchParams& P, bool right, uint32_t lb, uint32_t lbRev, uint32_t len, uint32_t* klb,
                                               uint32_t* klbRev, uint32_t* cnt) {
    const OccTable& tab = right ? P.bwtRev : P.bwt;
    const uint32_t lo = right ? lbRev : lb;
    const uint32_t hi = lo + len;
    OccBlk b1 = load_blk(tab.blk + (lo >> kBlkShift));
    OccSup s1 = load_sup(tab.sup + (lo >> kSupShift));
    OccBlk b2 = b1;
    OccSup s2 = s1;
    if ((lo >> kBlkShift) != (hi >> kBlkShift)) {
        b2 = load_blk(tab.blk + (hi >> kBlkShift));
        if ((lo >> kSupShift) != (hi >> kSupShift)) s2 = load_sup(tab.sup + (hi >> kSupShift));
    }
    uint32_t own[SIGMA];
    uint32_t sum1 = 0, sumc = 0;
#pragma unroll
    for (int s = 1; s < SIGMA; ++s) {
        const uint32_t a = s1.c[s] + blk_ctr(b1, s) + blk_count(b1, lo & 63u, s);
        const uint32_t b = s2.c[s] + blk_ctr(b2, s) + blk_count(b2, hi & 63u, s);
        own[s] = P.C[s] + a;
        cnt[s] = b - a;
        sum1 += a;
        sumc += b - a;
    }
    own[0] = lo - sum1;  // C[0] == 0
    cnt[0] = len - sumc;
    uint32_t other = right ? lb : lbRev;  // interval start on the side that is not probed
#pragma unroll
    for (int s = 0; s < SIGMA; ++s) {
        klb[s] = right ? other : own[s];
        klbRev[s] = right ? own[s] : other;
        other += cnt[s];
    }
}

// frames a thread's stack must hold: a node pushes at most 2 (SIGMA - 2) + 2 children and a path has at most
// length + kmax nodes (a deletion stays on its step)
__host__ __device__ inline uint32_t ordered_stack_frames(uint32_t len, uint32_t sigma) { return (len + 6u) * (2u * sigma - 2u) + 2u; }

// ordered walk: the frame is (a, -, 1, meta) — a cursor with ONE row whose occurrence is T[a, a + tlen): its children
// follow from one text symbol (in-text verification as in text_kernel) instead of two rank probes
constexpr uint32_t META_OTEXT = META_PAIR;

template <int SIGMA, bool EDIT>
__device__ __forceinline__ void fm_ordered_thread(const SearchParams& P, const uint32_t* s_steps, uint4* stack, uint32_t stride, uint32_t cap) {
    const uint32_t qlen = P.len;
    const uint32_t W = packed_words(qlen);
    const uint32_t max_hits = P.max_hits;
    uint32_t nodes = 0, emitted = 0, maxsp = 0;
    bool overflow = false;
    ChunkWriter outW;
    // ONE loop for "next query / next search / next frame" (as in fm_thread): with nested loops the lanes of a warp
    // would wait for the slowest search of the 32 before any of them starts its next one
    uint32_t qid = 0;
    const uint32_t* q = P.packed;
    bool toText = false;
    uint32_t taken = 0;               // rows delivered for the current query
    uint32_t j = P.n_searches;        // next search of the current query (forces the first claim)
    const uint32_t* tbl = s_steps;
    // the stack: frames 0 .. sp-1 in global memory, the newest frame in registers (`top`) — the child that is
    // visited next (mostly the match child) never makes the round trip through memory
    uint32_t sp = 0;
    uint4 top = make_uint4(0, 0, 0, 0);
    bool haveTop = false;
    auto qsym = [&](uint32_t pos) -> uint32_t { return (ldg32(q + (pos >> 3)) >> ((pos & 7u) * 4u)) & 0xfu; };
    auto push = [&](uint32_t nlb, uint32_t nlbRev, uint32_t nlen, uint32_t m) {
        if (toText && nlen == 1 && !(m & META_OTEXT)) {  // a unique cursor: continue at its text position SA[lb]
            nlb = ldg32(P.sa32 + nlb);
            m |= META_OTEXT;
        }
        if (haveTop) {
            if (sp < cap) stack[static_cast<uint64_t>(sp) * stride] = top;
            else overflow = true;
            ++sp;
        }
        top = make_uint4(nlb, nlbRev, nlen, m);
        haveTop = true;
    };
    constexpr unsigned long long kOrderedClaim = 4;
    unsigned long long wnext = 0, wend = 0;
    while (true) {
        if (taken >= max_hits) {  // the limit is reached: the query ends
            sp = 0;
            haveTop = false;
            j = P.n_searches;
        }
        if (!haveTop && sp == 0) {  // next search of the query, or the next query
            if (j == P.n_searches) {
                // (queries are claimed kOrderedClaim at a time: every thread of the launch adds to this one counter)
                if (wnext == wend) {
                    wnext = atomicAdd(&P.counters[CT_NEXT_QUERY], static_cast<unsigned long long>(kOrderedClaim));
                    wend = wnext + kOrderedClaim;
                }
                const unsigned long long w = wnext++;
                if (w >= P.n_queries) break;
                qid = P.redo ? ldg32(P.redo + w) : static_cast<uint32_t>(w);
                q = P.packed + static_cast<uint64_t>(qid) * W;
                // unique cursors are verified in the text, unless the query contains the delimiter (as in fm_kernel)
                toText = P.sa32 != nullptr;
                for (uint32_t i = 0; toText && i < W; ++i) {
                    const uint32_t v = ldg32(q + i);
                    toText = ((v - 0x11111111u) & ~v & 0x88888888u) == 0;  // no nibble is 0
                }
                taken = 0;
                j = 0;
            }
            tbl = s_steps + j * qlen;
            ++j;
            // the root: the q-gram entry when the leading steps allow no error (their only child is the match
            // child, so the jump keeps the order)
            uint4 root;
            if (fm_root(P, tbl, qsym, root) == 1) push(root.x, root.y, root.z, root.w);
            if (!haveTop) continue;  // the search cannot start
        }
        maxsp = sp + 1 > maxsp ? sp + 1 : maxsp;
        uint4 f = top;
        if (!haveTop) {
            --sp;
            if (sp >= cap) continue;  // (frame lost to an overflow: the call fails)
            f = stack[static_cast<uint64_t>(sp) * stride];
        }
        haveTop = false;
        const uint32_t lbRev = f.y, len = f.z, meta = f.w;
        uint32_t lb = f.x;  // (text frames: the text position a)
        uint32_t step = meta & 0x3ffu;
        const uint32_t e = (meta >> 10) & 0xfu;
        uint32_t Linfo = (meta >> 14) & 3u, Rinfo = (meta >> 16) & 3u;
        const bool inText = (meta & META_OTEXT) != 0;
        uint32_t tlen = (meta >> META_TLEN_SHIFT) & 0x3ffu;
        // match-only run of a text frame (steps with u == e: the match child is the only child, so walking them in
        // one go keeps the order): query and text compared 8 symbols per round on the packed words, as in text_kernel
        if (inText && step != qlen) {
            const uint8_t* runs = P.runs + static_cast<size_t>(j - 1) * qlen * kRunE;
            auto query8 = [&](uint32_t pos) -> uint32_t {  // the 8 query symbols from position pos; behind the query: 0xF
                const uint32_t w = pos >> 3;
                const uint32_t lo = ldg32(q + w);
                const uint32_t hi = w + 1 < W ? ldg32(q + w + 1) : 0xffffffffu;
                return funnel_r(lo, hi, (pos & 7u) * 4u);
            };
            bool dead = false;
            uint32_t R = runs[step * kRunE + e];
            while (R != 0) {
                const uint32_t st = tbl[step];
                const bool right = (st >> 24) & 1u;
                const uint32_t p0 = st & 0xffffu;
                uint32_t r = 0;
                if (right) {
                    while (r < R) {
                        const uint32_t n = R - r < 8u ? R - r : 8u;
                        const uint32_t x = (text8(P.text4, lb + tlen + r) ^ query8(p0 + r)) & nib_mask(n);
                        if (x != 0) { r += (ctz32(x) >> 2); break; }
                        r += n;
                    }
                } else {
                    while (r < R) {
                        if (lb < r + 1) break;  // the delimiter before position 0
                        const uint32_t endT = lb - 1 - r, endQ = p0 - r;  // compare the symbols ending here, downwards
                        uint32_t n = R - r < 8u ? R - r : 8u;
                        if (n > endT + 1) n = endT + 1;
                        const uint32_t x = (text8(P.text4, endT + 1 - n) ^ query8(endQ + 1 - n)) & nib_mask(n);
                        if (x != 0) { r += n - 1 - ((31u - clz32(x)) >> 2); break; }
                        r += n;
                    }
                }
                if (r < R) {  // a symbol differs: the state at step + r has no child
                    nodes += r + 1;
                    dead = true;
                    break;
                }
                nodes += R;
                step += R;
                tlen += R;
                if (right) Rinfo = INFO_M;
                else { Linfo = INFO_M; lb -= R; }
                if (step == qlen) break;  // reported below
                if (((tbl[step] >> 16) & 0xfu) > e + 1) { dead = true; break; }  // dead at the next step
                R = runs[step * kRunE + e];
            }
            if (dead) continue;
        }
        if (step == qlen) {  // the end of the query: report (edit distance: not behind a substitution or deletion at either end)
            if (!EDIT || sb200_pol_end(&P.pol, Linfo, Rinfo)) {
                const uint32_t n = len < max_hits - taken ? len : max_hits - taken;
                taken += n;
                uint4 cu = make_uint4(qid, lb, n, e);
                if (inText) cu = P.textpos_out ? make_uint4(qid, lb, 1, e | kCursorTextPosFlag) : make_uint4(qid, ldg32(P.isa32 + lb), 1, e);
                outW.put(P.out, P.out_cap, &P.counters[CT_OUT_SLOTS], cu);
                ++emitted;
            }
            continue;
        }
        const uint32_t st = tbl[step];
        const uint32_t l = (st >> 16) & 0xfu, u = (st >> 20) & 0xfu;
        const bool right = (st >> 24) & 1u;
        const bool matchOK = l <= e && e <= u;
        const bool mmOK = l <= e + 1 && e + 1 <= u;
        if (!matchOK && !mmOK) continue;
        const uint32_t c = qsym(st & 0xffffu);
        ++nodes;
        const uint32_t T = right ? Rinfo : Linfo;
        const uint32_t sideShift = right ? 16u : 14u;
        const uint32_t metaBase = ((right ? Linfo : 0u) << 14) | ((right ? 0u : Rinfo) << 16) | (meta & META_OTEXT);
        const uint32_t tlenSame = tlen << META_TLEN_SHIFT, tlenNext = (tlen + 1) << META_TLEN_SHIFT;
        const uint32_t mM = metaBase | (step + 1) | (e << 10) | (INFO_M << sideShift) | tlenNext;
        const uint32_t mD = metaBase | step | ((e + 1) << 10) | (INFO_D << sideShift) | tlenNext;
        const uint32_t mS = metaBase | (step + 1) | ((e + 1) << 10) | (INFO_S << sideShift) | tlenNext;
        const uint32_t mI = metaBase | (step + 1) | ((e + 1) << 10) | (INFO_I << sideShift) | tlenSame;
        const bool delOK = EDIT && sb200_pol_del(&P.pol, T);
        const bool insOK = EDIT && sb200_pol_ins(&P.pol, T);
        // order of the children (policy; children are pushed in REVERSE visiting order): the match child first, then by
        // default per symbol the deletion before the substitution, the insertion last
        const bool subFirst = (P.pol.child_order & SB200_CHILD_SUB_BEFORE_DEL) != 0;
        const bool insEarly = (P.pol.child_order & SB200_CHILD_INS_BEFORE_SYMBOLS) != 0;
        // the deletion / substitution children of one symbol (the same cursor)
        auto push_symbol = [&](uint32_t nlb, uint32_t nlbRev, uint32_t nlen, uint32_t flag) {
            if (subFirst) {
                if (delOK) push(nlb, nlbRev, nlen, mD | flag);
                push(nlb, nlbRev, nlen, mS | flag);
            } else {
                push(nlb, nlbRev, nlen, mS | flag);
                if (delOK) push(nlb, nlbRev, nlen, mD | flag);
            }
        };
        if (inText) {
            // the occurrence is T[a, a + tlen): the only non-empty child is the one of the text symbol next to it
            const uint32_t a = lb;
            uint32_t t = 0;  // (the delimiter before position 0)
            if (right) t = (ldg32(P.text4 + ((a + tlen) >> 3)) >> (((a + tlen) & 7u) * 4u)) & 0xfu;
            else if (a != 0) t = (ldg32(P.text4 + ((a - 1) >> 3)) >> (((a - 1) & 7u) * 4u)) & 0xfu;
            const uint32_t na = right ? a : a - 1;
            if (mmOK) {
                if (insOK && !insEarly) push(a, 0, 1, mI);
                if (t != c && t != 0) push_symbol(na, 0, 1, 0u);
                if (insOK && insEarly) push(a, 0, 1, mI);
            }
            if (matchOK && t == c) push(na, 0, 1, mM);  // popped first
            continue;
        }
        uint32_t klb[SIGMA], klbRev[SIGMA], cnt[SIGMA];
        probe_children<SIGMA>(P, right, lb, lbRev, len, klb, klbRev, cnt);
        if (mmOK) {
            if (insOK && !insEarly) push(lb, lbRev, len, mI);
#pragma unroll
            for (int s = SIGMA - 1; s >= 1; --s) {
                if (static_cast<uint32_t>(s) == c || cnt[s] == 0) continue;
                if (toText && cnt[s] == 1) push_symbol(ldg32(P.sa32 + klb[s]), 0, 1, META_OTEXT);  // (one load of the text position for both frames)
                else push_symbol(klb[s], klbRev[s], cnt[s], 0u);
            }
            if (insOK && insEarly) push(lb, lbRev, len, mI);
        }
        if (matchOK) {  // popped first
            uint32_t mc = 0, nlb = 0, nlbRev = 0;
#pragma unroll
            for (int s = 0; s < SIGMA; ++s)
                if (static_cast<uint32_t>(s) == c) { mc = cnt[s]; nlb = klb[s]; nlbRev = klbRev[s]; }
            if (mc != 0) push(nlb, nlbRev, mc, mM);
        }
    }
    outW.finish(P.out, P.out_cap);
    if (nodes) atomicAdd(&P.counters[CT_NODES], static_cast<unsigned long long>(nodes));
    if (overflow) atomicExch(&P.counters[CT_OVERFLOW], 1ull);
    atomicMax(&P.counters[CT_MAX_SP], static_cast<unsigned long long>(maxsp));
    if (emitted) atomicAdd(&P.counters[CT_CURSORS], static_cast<unsigned long long>(emitted));
}

// ================================================================================================
// text_pool_kernel: the same in-text verification, but the frames of a warp live in a pool in shared memory
// instead of 32 private stacks.  Every trip the warp pops up to 32 frames, each lane expands one of them (any
// lane, any seed: the frame names the seed slot whose staged query it needs) and pushes the children back.  Lanes
// therefore stay busy as long as the warp has frames at all — the private-stack version (text_thread) keeps 6.6 of
// 32 lanes active because the lanes drift into different phases — and a heavy seed is expanded by many lanes at
// once, which also shortens the drain at the end of the kernel.
//   * The pool is three stacks: RUN frames (next step is a match-only run: compare packed words), STATE frames (expand
//     search states, text_states) and PATH frames (a state at the start of a path window: up to 8 states along a matching
//     stretch expanded at once, text_path).  A trip pops from one of them (pool_pick), so all lanes of a trip execute
//     the same code.
//   * kPoolSlots seeds are in flight per warp; slots whose seed has no frame left (live[] counts them) are refilled
//     kRefillMin at a time, the queries of the new seeds staged by all lanes.
//   * A run frame compares at most kRunRounds x 8 symbols per pop, then is pushed back.
//   * The first kPoolCapS / kPoolCapR frames of a stack live in shared memory, the rest spills to global memory
//     (rare).  Pops of the state stack only narrow down when even the spill area is nearly full; with one frame per
//     trip the order is depth first and the stack cannot grow by more than the private-stack bound (STACK), which
//     is kept as head room.  Run frames never multiply (a pop pushes at most one frame); the path stack has no spill
//     area (a path frame that does not fit is an ordinary state frame).
// ================================================================================================
#if defined(SB200_POOL_CAP)
constexpr uint32_t kPoolCapS = SB200_POOL_CAP;     // (tests: tiny pools exercise the spill area and the narrow pops)
constexpr uint32_t kPoolCapR = SB200_POOL_CAP;
constexpr uint32_t kPoolCapP = SB200_POOL_CAP / 2;
constexpr uint32_t kSpillCap = SB200_SPILL_CAP;
#else
#if !defined(SB200_CAP_S)
#define SB200_CAP_S 40
#define SB200_CAP_R 44
#define SB200_CAP_P 34
#endif
constexpr uint32_t kPoolCapS = SB200_CAP_S;    // state frames per warp held in shared memory
constexpr uint32_t kPoolCapR = SB200_CAP_R;    // run frames per warp held in shared memory
constexpr uint32_t kPoolCapP = SB200_CAP_P;    // path frames per warp (shared memory only: one that does not fit is an ordinary state frame)
constexpr uint32_t kSpillCap = 2048;  // frames per warp and stack that spill to global memory behind them (rare)
#endif
#if defined(SB200_POOL_SLOTS)
constexpr uint32_t kPoolSlots = SB200_POOL_SLOTS;
constexpr uint32_t kPoolThreads = SB200_POOL_THREADS;
#else
constexpr uint32_t kPoolSlots = 52;   // seeds a warp works on at a time (52: 36 warps per SM still fit next to the tables of a 4-search scheme at 150 bp)
constexpr uint32_t kPoolThreads = 384;  // at most 12 warps per block: three blocks (36 warps) fit the shared memory of an SM at 150 bp
#endif
#if !defined(SB200_REFILL_MIN)
#define SB200_REFILL_MIN 16
#define SB200_REFILL_DRY 48
#endif
constexpr uint32_t kRefillMin = SB200_REFILL_MIN;  // free seed slots that trigger a refill ...
constexpr uint32_t kRefillDry = SB200_REFILL_DRY;  // ... or fewer frames than this in the pool
constexpr uint32_t kRunRounds = 3;    // rounds of 8 symbols per pop (measured with path frames: 1 -> 4.32 ms, 2 -> 4.10, 3 -> 4.04, 4 -> 4.05, 8 -> 4.13)

struct FrameStack {
    uint2* frames;   // [cap] (a, meta): frames 0 .. cap-1 of the stack, shared memory
    uint8_t* slots;  // [cap] seed slot of the frame
    uint4* spill;    // [kSpillCap] (a, meta, slot, -): frames cap .. of the stack, global memory
    uint32_t* top;   // number of frames
    uint32_t cap;
};
struct TextPool {
    FrameStack S, R;       // state frames, run frames
    FrameStack Pth;        // path frames: state frames at the start of a path window (text_path); no spill area
    uint32_t* live;        // [kPoolSlots] frames of the slot's seed that are still in the pool or being expanded
    uint32_t* ctx_qid;     // [kPoolSlots] query id of the seed in the slot
    uint32_t* ctx_search;  // [kPoolSlots] its search, as the index of the search's first step in the tables (search * len)
    uint32_t* query;       // [kPoolSlots][Wp] staged packed queries
    uint32_t Wp;           // odd stride (words) between the queries: spreads the slots over the banks
};
struct PoolLane {
    uint32_t nodes{0}, emitted{0};
    ChunkWriter outW;
    bool overflow{false};
};
__host__ __device__ inline uint32_t pool_query_stride(uint32_t len) { return packed_words(len) | 1u; }
// bytes of shared memory one warp's pool takes
__host__ __device__ inline uint32_t pool_bytes(uint32_t len) {
    const uint32_t frames = kPoolCapS + kPoolCapR + kPoolCapP;
    return frames * 8u + ((frames + 7u) & ~7u) + 16u + 3u * kPoolSlots * 4u + kPoolSlots * pool_query_stride(len) * 4u;
}
// carves one warp's pool out of `base` (8-byte aligned, pool_bytes(len) bytes); spill: 2 * kSpillCap entries
__host__ __device__ inline TextPool pool_carve(uint8_t* base, uint4* spill, uint32_t len) {
    TextPool pool;
    const uint32_t frames = kPoolCapS + kPoolCapR + kPoolCapP;
    uint2* fr = reinterpret_cast<uint2*>(base);
    uint8_t* sl = base + frames * 8u;
    uint32_t* words = reinterpret_cast<uint32_t*>(base + frames * 8u + ((frames + 7u) & ~7u));
    pool.S = FrameStack{fr, sl, spill, words, kPoolCapS};
    pool.R = FrameStack{fr + kPoolCapS, sl + kPoolCapS, spill + kSpillCap, words + 1, kPoolCapR};
    pool.Pth = FrameStack{fr + kPoolCapS + kPoolCapR, sl + kPoolCapS + kPoolCapR, nullptr, words + 2, kPoolCapP};
    pool.live = words + 4;
    pool.ctx_qid = words + 4 + kPoolSlots;
    pool.ctx_search = words + 4 + 2 * kPoolSlots;
    pool.query = words + 4 + 3 * kPoolSlots;
    pool.Wp = pool_query_stride(len);
    return pool;
}
// How many state frames the warp may pop when `top` are present and one frame pushes at most `maxpush` children
// (into either stack; `topR` run frames are present).
__host__ __device__ inline uint32_t pool_pop_width(uint32_t top, uint32_t topR, uint32_t maxpush, uint32_t lanes, uint32_t stack) {
    const uint32_t cap = kPoolCapS + kSpillCap - stack;
    uint32_t n = top < lanes ? top : lanes;
    if (top + lanes * (maxpush - 1u) <= cap && topR + lanes * maxpush <= kPoolCapR + kSpillCap) return n;  // (no division on the common path)
    const uint32_t room = top < cap ? (cap - top) / (maxpush - 1u) : 0u;
    if (n > room) n = room ? room : 1u;
    const uint32_t capR = kPoolCapR + kSpillCap;
    const uint32_t roomR = topR < capR ? (capR - topR) / maxpush : 0u;  // callers pop run frames first when this is 0
    if (n > roomR) n = roomR;
    return n;
}
#if defined(SB200_HOST_EMU)
static inline uint32_t top_fetch_add(uint32_t* p, uint32_t n) { const uint32_t v = *p; *p = v + n; return v; }
#else
__device__ __forceinline__ uint32_t top_fetch_add(uint32_t* p, uint32_t n) { return atomicAdd(p, n); }
#endif
__device__ __forceinline__ void stack_push(const FrameStack& st, uint32_t a, uint32_t meta, uint32_t slot, PoolLane& ls) {
#if defined(SB200_HOST_EMU)
    const uint32_t idx = (*st.top)++;
#else
    const uint32_t idx = atomicAdd(st.top, 1u);
#endif
    if (idx < st.cap) {
        st.frames[idx] = make_uint2(a, meta);
        st.slots[idx] = static_cast<uint8_t>(slot);
    } else if (idx - st.cap < kSpillCap) {
        st.spill[idx - st.cap] = make_uint4(a, meta, slot, 0);
    } else {
        ls.overflow = true;
    }
}
// push onto the run stack (run) or the state stack of the pool: pool_carve lays the run stack right behind the state
// stack (frames, slots, tops, spill), so the choice is an offset instead of a select over five pointers
__device__ __forceinline__ void pool_push_to(const TextPool& pool, bool run, uint32_t a, uint32_t meta, uint32_t slot, PoolLane& ls) {
#if defined(SB200_HOST_EMU)
    const uint32_t idx = pool.S.top[run ? 1 : 0]++;
#else
    const uint32_t idx = atomicAdd(pool.S.top + (run ? 1 : 0), 1u);
#endif
    const uint32_t cap = run ? kPoolCapR : kPoolCapS;
    if (idx < cap) {
        const uint32_t at = idx + (run ? kPoolCapS : 0u);
        pool.S.frames[at] = make_uint2(a, meta);
        pool.S.slots[at] = static_cast<uint8_t>(slot);
    } else if (idx - cap < kSpillCap) {
        pool.S.spill[idx - cap + (run ? kSpillCap : 0u)] = make_uint4(a, meta, slot, 0);
    } else {
        ls.overflow = true;
    }
}
// a PATH frame (the path stack is small: a frame that does not fit is pushed as an ordinary state frame, which is always right)
__device__ __forceinline__ void path_push(const TextPool& pool, uint32_t a, uint32_t meta, uint32_t slot, PoolLane& ls) {
#if defined(SB200_HOST_EMU)
    const uint32_t idx = (*pool.Pth.top)++;
#else
    const uint32_t idx = atomicAdd(pool.Pth.top, 1u);
#endif
    if (idx < kPoolCapP) {
        pool.Pth.frames[idx] = make_uint2(a, meta);
        pool.Pth.slots[idx] = static_cast<uint8_t>(slot);
        return;
    }
#if defined(SB200_HOST_EMU)
    (*pool.Pth.top)--;
#else
    atomicSub(pool.Pth.top, 1u);
#endif
    pool_push_to(pool, false, a, meta, slot, ls);
}
// frame idx of the stack -> (a, meta), slot
__device__ __forceinline__ uint2 stack_get(const FrameStack& st, uint32_t idx, uint32_t& slot) {
    if (idx < st.cap) {
        slot = st.slots[idx];
        return st.frames[idx];
    }
    const uint4 v = st.spill[idx - st.cap];
    slot = v.z;
    return make_uint2(v.x, v.y);
}
// a frame goes to the run stack when its next step starts a match-only run
__device__ __forceinline__ void pool_push(const TextPool& pool, const uint8_t* runs, uint32_t a, uint32_t meta, uint32_t slot, PoolLane& ls) {
    const uint32_t step = meta & 0x3ffu, e = (meta >> 10) & 0xfu;
    const bool run = (meta & META_PAIR) == 0 && runs[step * kRunE + e] != 0;
    stack_push(run ? pool.R : pool.S, a, meta, slot, ls);
}

// per-frame view of the seed context
struct SeedCtx {
    const uint32_t* q;
    const uint32_t* tbl;
    const uint8_t* runs;
    const uint8_t* flags;  // state flags of the search (build_state_flags)
    const uint8_t* win;    // path windows of the search (build_path_windows)
    uint32_t qid, W;
    __device__ __forceinline__ uint32_t qsym(uint32_t pos) const { return (q[pos >> 3] >> ((pos & 7u) * 4u)) & 0xfu; }
    // the 8 query symbols that start at position pos; positions behind the query read as 0xF
    __device__ __forceinline__ uint32_t query8(uint32_t pos) const {
        const uint32_t w = pos >> 3;
        const uint32_t lo = q[w];
        const uint32_t hi = w + 1 < W ? q[w + 1] : 0xffffffffu;
        return funnel_r(lo, hi, (pos & 7u) * 4u);
    }
    // the 8 query symbols at the positions pos, pos - 1, .., pos - 7 (nibble j = position pos - j); positions before the
    // query read as 0xF
    __device__ __forceinline__ uint32_t query8_down(uint32_t pos) const {
        if (pos >= 7u) return rev8(query8(pos - 7u));
        const uint32_t sh = (7u - pos) * 4u;
        return rev8((query8(0) << sh) | ((1u << sh) - 1u));
    }
};
__device__ __forceinline__ SeedCtx seed_ctx(const SearchParams& P, const uint32_t* s_steps, const uint8_t* s_runs, const TextPool& pool,
                                            uint32_t slot) {
    const uint32_t first_step = pool.ctx_search[slot];  // search * len: the first step of the seed's search in the tables
    const uint8_t* runs = s_runs + first_step * kRunE;
    const uint32_t off = state_flags_offset(P.n_searches * P.len);
    return SeedCtx{pool.query + slot * pool.Wp, s_steps + first_step, runs, runs + off, runs + 2u * off, pool.ctx_qid[slot], packed_words(P.len)};
}
__device__ __forceinline__ void pool_emit(const SearchParams& P, PoolLane& ls, uint32_t qid, uint32_t a, uint32_t e) {
    ls.outW.put(P.out, P.out_cap, &P.counters[CT_OUT_SLOTS], P.textpos_out ? make_uint4(qid, a, 1, e | kCursorTextPosFlag) : make_uint4(qid, P.isa32[a], 1, e));
    if (P.qcount != nullptr) atomicAdd(&P.qcount[qid], 1u);
    ++ls.emitted;
}

// where a frame whose extended end carries M goes: the run stack when its step starts a match-only run, the path stack
// when a path window starts there, else the state stack
__device__ __forceinline__ void route_push(const TextPool& pool, const SeedCtx& cx, uint32_t a, uint32_t meta, uint32_t slot, PoolLane& ls) {
    const uint32_t at = (meta & 0x3ffu) * kRunE + ((meta >> 10) & 0xfu);
    if (cx.runs[at] != 0) stack_push(pool.R, a, meta, slot, ls);
    else if (cx.win[at] >= 2u) path_push(pool, a, meta | META_STREAK, slot, ls);
    else stack_push(pool.S, a, meta, slot, ls);
}

// RUN frame: the match-only run(s) that start at its step, comparing packed words of query and text; returns the
// number of frames pushed (0 or 1)
template <bool EDIT>
__device__ __forceinline__ uint32_t text_run(const SearchParams& P, const uint32_t* s_steps, const uint8_t* s_runs, const TextPool& pool,
                                             const uint2 f, const uint32_t slot, PoolLane& ls, const uint32_t run_rounds) {
    const SeedCtx cx = seed_ctx(P, s_steps, s_runs, pool, slot);
    const uint32_t qlen = P.len;
    uint32_t a = f.x;
    const uint32_t meta = f.y;
    uint32_t step = meta & 0x3ffu;
    const uint32_t e = (meta >> 10) & 0xfu;
    uint32_t Linfo = (meta >> 14) & 3u, Rinfo = (meta >> 16) & 3u;
    uint32_t tlen = (meta >> META_TLEN_SHIFT) & 0x3ffu;
    uint32_t nodes = 0;
    uint32_t R = cx.runs[step * kRunE + e];
    uint32_t budget = run_rounds;
    while (R != 0) {
        const uint32_t st = cx.tbl[step];
        const bool right = (st >> 24) & 1u;
        const uint32_t p0 = st & 0xffffu;
        uint32_t r = 0;
        bool differs = false;
        // one code path for both directions: a window of n symbols that starts at tpos / qpos
        while (r < R && budget != 0) {
            uint32_t n = R - r < 8u ? R - r : 8u;
            if (!right) {
                if (a < r + 1) { differs = true; break; }  // the delimiter before position 0
                if (n > a - r) n = a - r;
            }
            --budget;
            const uint32_t tpos = right ? a + tlen + r : a - r - n;
            const uint32_t qpos = right ? p0 + r : p0 + 1 - r - n;
            const uint32_t x = (text8(P.text4, tpos) ^ cx.query8(qpos)) & nib_mask(n);
            if (x != 0) {
                r += right ? (ctz32(x) >> 2) : n - 1 - ((31u - clz32(x)) >> 2);
                differs = true;
                break;
            }
            r += n;
        }
        if (differs) {  // the state at step + r has no child
            ls.nodes += nodes + r + 1;
            return 0;
        }
        nodes += r;
        step += r;
        tlen += r;
        if (r != 0) {
            if (right) Rinfo = INFO_M;
            else { Linfo = INFO_M; a -= r; }
        }
        if (r < R) {  // out of rounds: the rest of the run waits for the next pop
            ls.nodes += nodes;
            stack_push(pool.R, a, pack_meta(step, e, Linfo, Rinfo) | (tlen << META_TLEN_SHIFT), slot, ls);
            return 1;
        }
        if (step == qlen) {  // the last step matched
            if (!EDIT || sb200_pol_end(&P.pol, Linfo, Rinfo)) pool_emit(P, ls, cx.qid, a, e);  // (the end that moved carries M)
            ls.nodes += nodes;
            return 0;
        }
        if (((cx.tbl[step] >> 16) & 0xfu) > e + 1) { ls.nodes += nodes; return 0; }  // dead at the next step
        R = cx.runs[step * kRunE + e];
    }
    ls.nodes += nodes;
    // states follow; behind a run the extended end carries M: a path window may start here
    const uint32_t m2 = pack_meta(step, e, Linfo, Rinfo) | (tlen << META_TLEN_SHIFT);
    const bool right = (cx.tbl[step] >> 24) & 1u;
    if (nodes != 0 && (right ? Rinfo : Linfo) == INFO_M && cx.win[step * kRunE + e] >= 2u) path_push(pool, a, m2 | META_STREAK, slot, ls);
    else stack_push(pool.S, a, m2, slot, ls);
    return 1;
}

// STATE frame: the states on its cursor (pair halves, insertion chain), expanded like fm_node does (the conditions
// that depend on the scheme tables only come from the state flags); returns the number of frames pushed
template <bool EDIT>
__device__ __forceinline__ uint32_t text_states(const SearchParams& P, const uint32_t* s_steps, const uint8_t* s_runs, const TextPool& pool,
                                                const uint2 f, const uint32_t slot, PoolLane& ls) {
    const SeedCtx cx = seed_ctx(P, s_steps, s_runs, pool, slot);
    const uint32_t qlen = P.len;
    const uint32_t* tbl = cx.tbl;
    uint32_t pushed = 0;
    auto push = [&](uint32_t a, uint32_t m, bool run) {
        pool_push_to(pool, run, a, m, slot, ls);
        ++pushed;
    };
    const uint32_t a = f.x;
    const uint32_t meta = f.y;
    uint32_t step = meta & 0x3ffu, e = (meta >> 10) & 0xfu;
    uint32_t Linfo = (meta >> 14) & 3u, Rinfo = (meta >> 16) & 3u;
    const bool pair = (meta & META_PAIR) != 0;
    const uint32_t tlen = (meta >> META_TLEN_SHIFT) & 0x3ffu;
    uint32_t nodes = 0;

    const uint32_t b = a + tlen;
    const bool rightFrame = (tbl[step] >> 24) & 1u;  // the end the frame's own error (D / S) sits on
    // the text symbols left and right of the occurrence (the delimiter before position 0)
    uint32_t tL = 0;
    if (a != 0) tL = (P.text4[(a - 1) >> 3] >> (((a - 1) & 7u) * 4u)) & 0xfu;
    const uint32_t tR = (P.text4[b >> 3] >> ((b & 7u) * 4u)) & 0xfu;
    const uint32_t tlenNext = (tlen + 1) << META_TLEN_SHIFT;

    bool second = false;
    bool first = !pair;  // the first state of a frame that is not a pair: where a matching path runs through
    while (true) {
        ++nodes;
        const uint32_t st = tbl[step];
        const uint32_t fl = cx.flags[step * kRunE + e];
        const bool right = (st >> 24) & 1u;
        const uint32_t c = cx.qsym(st & 0xffffu);
        const bool last = step + 1 == qlen;
        const bool mmOK = (fl & SF_MISMATCH) != 0;
        const uint32_t T = right ? Rinfo : Linfo;
        const uint32_t O = right ? Linfo : Rinfo;
        const bool otherEndOK = !EDIT || sb200_pol_end1(&P.pol, O);  // end filter of the policy, other end
        const uint32_t t = right ? tR : tL;    // the only symbol whose child cursor is not empty
        const uint32_t na = right ? a : a - 1;  // child occurrence T[na, na + tlen + 1)
        const uint32_t sideShift = right ? 16u : 14u;
        const uint32_t metaBase = ((right ? Linfo : 0u) << 14) | ((right ? 0u : Rinfo) << 16) | tlenNext;
        if (t != 0) {
            if (t == c) {
                if (fl & SF_MATCH) {
                    if (last) {
                        if (otherEndOK && (!EDIT || sb200_pol_end1(&P.pol, INFO_M))) pool_emit(P, ls, cx.qid, na, e);
                    } else if (fl & SF_M_ALIVE) {
                        const uint32_t mM = metaBase | (step + 1) | (e << 10) | (INFO_M << sideShift);
                        if (fl & SF_RUN_M) push(na, mM, true);
                        else if (first && T == INFO_M && cx.win[(step + 1) * kRunE + e] >= 2u &&
                                 ((tbl[step + 1] >> 24) & 1u) == static_cast<uint32_t>(right)) {
                            // the second match in a row at this end: most likely an alignment, and a path window starts at the child
                            path_push(pool, na, mM | META_STREAK, slot, ls);
                            ++pushed;
                        } else push(na, mM, false);
                    }
                }
            } else if (mmOK) {
                const bool delOK = EDIT && sb200_pol_del(&P.pol, T);
                const uint32_t mD = metaBase | step | ((e + 1) << 10) | (INFO_D << sideShift);
                const uint32_t mS = metaBase | (step + 1) | ((e + 1) << 10) | (INFO_S << sideShift);
                // a pair only when both halves extend the same end (as in fm_node) and the policy allows pairs (SF_PAIR)
                if (delOK && (fl & SF_PAIR)) push(na, mD | META_PAIR, false);
                else {
                    if (delOK) push(na, mD, (fl & SF_RUN_D) != 0);
                    if (fl & SF_SUB_ALIVE) push(na, mS, (fl & SF_RUN_S) != 0);
                }
                // a substitution at the last step is reported under Hamming distance; under edit distance only when the
                // policy reports an S at an end
                if (last && (!EDIT || (otherEndOK && sb200_pol_end1(&P.pol, INFO_S)))) pool_emit(P, ls, cx.qid, na, e + 1);
            }
        }
        // next state on the same cursor
        first = false;
        if (pair) {
            if (second) break;
            second = true;
            step += 1;  // the substitution half: (step + 1, e), S at the end both halves extend
            if (rightFrame) Rinfo = INFO_S; else Linfo = INFO_S;
            continue;
        }
        const bool insOK = EDIT && mmOK && sb200_pol_ins(&P.pol, T);
        if (!insOK) break;
        if (last) {
            if (otherEndOK && sb200_pol_end1(&P.pol, INFO_I)) pool_emit(P, ls, cx.qid, a, e + 1);
            break;
        }
        if (!(fl & SF_SUB_ALIVE)) break;  // dead at the next step
        step += 1;
        e += 1;
        if (right) Rinfo = INFO_I; else Linfo = INFO_I;
    }
    ls.nodes += nodes;
    return pushed;
}

// PATH frame: a state (step, e) whose extended end carries M at the start of a path window (build_path_windows).  The lane
// expands up to 8 consecutive states (step + j, e) along the matching symbols at once — what text_states does with them one
// frame at a time — from packed-word compares of the text window with the query on the diagonals 0 .. k - e:
//   diagonal 0   how far the path matches (m states);
//   diagonal i   the i-th state of the insertion chain of every path state j < m compares the text symbol t_j with the
//                query symbol of step + j + i: the match child, or the deletion / substitution children, per path state.
// The flags of the window are those of its first step (that is what makes it a window), so the only data-dependent part
// is which nibbles are equal.  A child that starts a match-only run is compared with its first symbol right here (the
// next nibble of its diagonal): a run that ends there costs its one node and no frame.  Children and node counts are
// exactly those of text_states + text_run on the same states.  wlim: states a lane may take this trip (stack room).
template <bool EDIT>
__device__ __forceinline__ uint32_t text_path(const SearchParams& P, const uint32_t* s_steps, const uint8_t* s_runs, const TextPool& pool,
                                              const uint2 f, const uint32_t slot, PoolLane& ls, const uint32_t wlim) {
    const SeedCtx cx = seed_ctx(P, s_steps, s_runs, pool, slot);
    const uint32_t a = f.x, meta = f.y & ~META_STREAK;
    const uint32_t step = meta & 0x3ffu, e = (meta >> 10) & 0xfu;
    const uint32_t Linfo = (meta >> 14) & 3u, Rinfo = (meta >> 16) & 3u;
    const uint32_t tlen = (meta >> META_TLEN_SHIFT) & 0x3ffu;
    const uint32_t st = cx.tbl[step];
    const bool right = (st >> 24) & 1u;
    const uint32_t p0 = st & 0xffffu;
    uint32_t w = cx.win[step * kRunE + e];
    if (w > wlim) w = wlim;
    // no window after all (pair frame, another operation at the end, too close to the text start): an ordinary state frame
    if (w == 0 || (meta & META_PAIR) || (right ? Rinfo : Linfo) != INFO_M || (!right && a < 8u)) {
        stack_push(pool.S, a, meta, slot, ls);
        return 1;
    }
    const uint32_t tw = right ? text8(P.text4, a + tlen) : rev8(text8(P.text4, a - 8u));  // nibble j = the symbol path state j consumes
    const uint32_t nz0 = nz_nibbles(tw ^ (right ? cx.query8(p0) : cx.query8_down(p0)));
    uint32_t m = nz0 ? ctz32(nz0) >> 2 : 8u;  // path states whose symbol matches
    if (m > w) m = w;
    if (m == 0) {
        stack_push(pool.S, a, meta, slot, ls);
        return 1;
    }
    uint32_t pushed = 0, nodes = m;
    const uint32_t sideShift = right ? 16u : 14u;
    const uint32_t keep = ((right ? Linfo : 0u) << 14) | ((right ? 0u : Rinfo) << 16);  // the other end
    if (EDIT) {
        const uint32_t valid = nib_mask(m) & 0x11111111u;
        const bool delOK = sb200_pol_del(&P.pol, INFO_I) != 0;  // (the chain states carry I at the extended end)
        uint32_t flp = cx.flags[step * kRunE + e], Tp = INFO_M, nzp = nz0;
        for (uint32_t i = 1; (flp & SF_MISMATCH) && sb200_pol_ins(&P.pol, Tp) && (flp & SF_SUB_ALIVE); ++i) {
            const uint32_t fl = cx.flags[(step + i) * kRunE + e + i];
            nodes += m;
            const uint32_t nzi = nz_nibbles(tw ^ (right ? cx.query8(p0 + i) : cx.query8_down(p0 - i)));
            const bool pairF = delOK && (fl & SF_PAIR);
            // the children of this chain level, one kind after the other (one copy of the push loop: code size):
            // match children of the path states whose symbol equals, deletion (or pair) and substitution children of the others
#pragma unroll 1
            for (uint32_t q = 0; q < 3u; ++q) {
                uint32_t mk = 0, look = nzi, info = INFO_M, rbit = SF_RUN_M, dstep = 1, de = 0, extra = 0;
                if (q == 0) {
                    if ((fl & SF_MATCH) && (fl & SF_M_ALIVE)) mk = ~nzi & valid;
                } else if (fl & SF_MISMATCH) {
                    if (q == 1) {
                        if (delOK) {
                            mk = nzi & valid;
                            info = INFO_D;
                            rbit = pairF ? 0u : static_cast<uint32_t>(SF_RUN_D);
                            dstep = 0;
                            de = 1;
                            look = nzp;  // a deletion consumes no query symbol: its first step compares on the previous diagonal
                            extra = pairF ? META_PAIR : 0u;
                        }
                    } else if (!pairF && (fl & SF_SUB_ALIVE)) {
                        mk = nzi & valid;
                        info = INFO_S;
                        rbit = SF_RUN_S;
                        de = 1;
                    }
                }
                if (mk == 0) continue;
                const bool run = (fl & rbit) != 0;
                if (run) {  // runs that end at their first symbol (the next nibble of their diagonal): one node each, no frame
                    const uint32_t dead = mk & (look >> 4);
                    nodes += popc32(dead);
                    mk &= ~dead;
                    if (mk == 0) continue;
                }
                const uint32_t n = popc32(mk);
                uint32_t idx = top_fetch_add(pool.S.top + (run ? 1 : 0), n);  // one reservation for all of them
                const uint32_t cap = run ? kPoolCapR : kPoolCapS, base = run ? kPoolCapS : 0u;
                const uint32_t m0 = keep | ((tlen + 1u) << META_TLEN_SHIFT) | (step + i + dstep) | ((e + i + de) << 10) | (info << sideShift) | extra;
                for (; mk; mk &= mk - 1u, ++idx) {
                    const uint32_t jj = ctz32(mk) >> 2;
                    const uint32_t ca = right ? a : a - jj - 1u;
                    const uint32_t cmeta = m0 + jj * ((1u << META_TLEN_SHIFT) + 1u);  // step + j, tlen + j
                    if (idx < cap) {
                        pool.S.frames[base + idx] = make_uint2(ca, cmeta);
                        pool.S.slots[base + idx] = static_cast<uint8_t>(slot);
                    } else if (idx - cap < kSpillCap) {
                        pool.S.spill[idx - cap + (run ? kSpillCap : 0u)] = make_uint4(ca, cmeta, slot, 0);
                    } else {
                        ls.overflow = true;
                    }
                }
                pushed += n;
            }
            flp = fl;
            Tp = INFO_I;
            nzp = nzi;
        }
    }
    // the state behind the last matching one: unseen when the window (or the trip's limit) ended, else a state whose
    // symbol differs — the ordinary expansion
    const uint32_t cm = keep | ((tlen + m) << META_TLEN_SHIFT) | (step + m) | (e << 10) | (static_cast<uint32_t>(INFO_M) << sideShift);
    if (m == w) route_push(pool, cx, right ? a : a - m, cm, slot, ls);
    else stack_push(pool.S, right ? a : a - m, cm, slot, ls);
    ls.nodes += nodes;
    return pushed + 1u;
}

// one seed into slot `slot`: stage its query (COPY; the kernel stages the queries of a refill with all lanes instead),
// remember its context, push its root frame
template <bool COPY>
__device__ __forceinline__ void pool_load_seed(const SearchParams& P, const uint8_t* s_runs, const TextPool& pool, uint32_t slot,
                                               const uint4 seed, PoolLane& ls) {
    if (COPY) {
        const uint32_t W = packed_words(P.len);
        const uint32_t* src = P.packed + static_cast<uint64_t>(seed.x) * W;
        uint32_t* dst = pool.query + slot * pool.Wp;
        for (uint32_t w = 0; w < W; ++w) dst[w] = src[w];
    }
    pool.ctx_qid[slot] = seed.x;
    pool.ctx_search[slot] = seed.z * P.len;
    pool.live[slot] = 1;
    // the root frame: run stack, path stack (the extended end carries M and a path window starts) or state stack
    const uint32_t step = seed.w & 0x3ffu, e = (seed.w >> 10) & 0xfu;
    const uint32_t at = (seed.z * P.len + step) * kRunE + e;
    const bool right = (P.steps[seed.z * P.len + step] >> 24) & 1u;
    const uint32_t a = P.sa32[seed.y];
    if (seed.w & META_PAIR) stack_push(pool.S, a, seed.w, slot, ls);
    else if (s_runs[at] != 0) stack_push(pool.R, a, seed.w, slot, ls);
    else if (((seed.w >> (right ? 16u : 14u)) & 3u) == INFO_M && s_runs[2u * state_flags_offset(P.n_searches * P.len) + at] >= 2u)
        path_push(pool, a, seed.w | META_STREAK, slot, ls);
    else stack_push(pool.S, a, seed.w, slot, ls);
}

// The next trip of a warp: which stack it pops, how many lanes, and for a path trip how many states a lane may take.
// Run frames first when there is a full trip of them or the run stack leaves no room for the children of a state trip;
// then full trips of path and state frames; else the fullest stack.  A path lane pushes up to
// 1 + w * (maxpush - 2) frames (w states, per state and chain level a match child or a deletion + substitution).
struct PoolTrip { uint32_t kind, n, w; };  // kind: 0 state, 1 run, 2 path
// (the stacks are nearly full: lanes and states per lane from the room that is left; rare, not inlined)
__host__ __device__ SB200_COLD PoolTrip pool_pick_tight(uint32_t topS, uint32_t topR, uint32_t topP, uint32_t maxpush, uint32_t stack) {
    const uint32_t capS = kPoolCapS + kSpillCap - stack, capR = kPoolCapR + kSpillCap;
    const uint32_t nR = topR < 32u ? topR : 32u;
    if (topR != 0 && (topS + topP == 0 || topR >= 32u || topR + 32u * maxpush > capR)) return PoolTrip{1u, nR, 0u};
    uint32_t nP = 0, w = 0;
    if (topP != 0) {
        const uint32_t roomS = topS < capS ? capS - topS : 0u, roomR = capR - topR;
        const uint32_t room = roomS < roomR ? roomS : roomR;
        const uint32_t per = maxpush > 2u ? maxpush - 2u : 1u;
        nP = topP < 32u ? topP : 32u;
        w = room / nP > 1u ? (room / nP - 1u) / per : 0u;
        if (w > kPathWindow) w = kPathWindow;
        if (w == 0) {  // one state per lane, fewer lanes
            nP = room / (1u + per) < nP ? room / (1u + per) : nP;
            w = nP ? 1u : 0u;
        }
    }
    if (w != 0 && (topP >= 32u || (topS < 32u && topP >= topS && topP >= topR))) return PoolTrip{2u, nP, w};
    if (topS != 0 && (topS >= topR || topR == 0)) {
        const uint32_t n = pool_pop_width(topS, topR, maxpush, 32u, stack);
        if (n != 0) return PoolTrip{0u, n, 0u};
    }
    if (topR != 0) return PoolTrip{1u, nR, 0u};
    if (topS != 0) return PoolTrip{0u, pool_pop_width(topS, topR, maxpush, 32u, stack), 0u};
    return PoolTrip{2u, nP, w};  // only path frames are left (w > 0: both other stacks are empty)
}
__host__ __device__ inline PoolTrip pool_pick(uint32_t topS, uint32_t topR, uint32_t topP, uint32_t maxpush, uint32_t stack) {
    const uint32_t capS = kPoolCapS + kSpillCap - stack, capR = kPoolCapR + kSpillCap;
    const uint32_t per = maxpush > 2u ? maxpush - 2u : 1u;
    const uint32_t wfull = per <= 6u ? kPathWindow : kPathWindow / 2u;  // (k = 4: half windows, so that a full trip always has room)
    const uint32_t worst = 32u * (1u + wfull * per);  // frames a trip can push (a state trip: 32 * maxpush, less)
    if (topS + worst > capS || topR + worst > capR) return pool_pick_tight(topS, topR, topP, maxpush, stack);
    const uint32_t nS = topS < 32u ? topS : 32u, nR = topR < 32u ? topR : 32u, nP = topP < 32u ? topP : 32u;
    if (topR != 0 && (topS + topP == 0 || topR >= 32u)) return PoolTrip{1u, nR, 0u};
    if (topP != 0 && (topP >= 32u || (topS < 32u && topP >= topS && topP >= topR))) return PoolTrip{2u, nP, wfull};
    if (topS != 0 && topS >= topR) return PoolTrip{0u, nS, 0u};
    if (topR != 0) return PoolTrip{1u, nR, 0u};
    return PoolTrip{2u, nP, wfull};
}

// the expanded frame is gone, `pushed` frames of the same seed were added
__device__ __forceinline__ void pool_retire(const TextPool& pool, uint32_t slot, uint32_t pushed) {
    if (pushed == 1) return;
#if defined(SB200_HOST_EMU)
    pool.live[slot] += pushed - 1u;
#else
    atomicAdd(&pool.live[slot], pushed - 1u);
#endif
}

__device__ __forceinline__ void pool_finish(const SearchParams& P, PoolLane& ls, uint32_t maxtop) {
    ls.outW.finish(P.out, P.out_cap);
    if (ls.overflow) atomicExch(&P.counters[CT_OVERFLOW], 1ull);
#if defined(SB200_HOST_EMU)
    // (the emulation calls this once per lane, one after the other)
    P.counters[CT_NODES] += ls.nodes;
    P.counters[CT_NODES_TEXT] += ls.nodes;
    P.counters[CT_CURSORS] += ls.emitted;
    if (maxtop > P.counters[CT_MAX_SP]) P.counters[CT_MAX_SP] = maxtop;
#else
    // (the warp is converged: its loop ends on warp-uniform conditions)
    const uint32_t nodes = __reduce_add_sync(0xffffffffu, ls.nodes);
    if ((threadIdx.x & 31u) == 0 && nodes != 0) {
        atomicAdd(&P.counters[CT_NODES], static_cast<unsigned long long>(nodes));
        atomicAdd(&P.counters[CT_NODES_TEXT], static_cast<unsigned long long>(nodes));
    }
    warp_counter_add(&P.counters[CT_CURSORS], ls.emitted);
    warp_counter_max(&P.counters[CT_MAX_SP], maxtop);
#endif
}

#if !defined(SB200_HOST_EMU)
// ---- search_n, selection of the queries that exceed the limit -----------------------------------------
// rows[q] += rows of every cursor of query q (cursors: the output slots of the first pass)
__global__ void __launch_bounds__(256) cursor_rows_kernel(const uint4* cursors, uint64_t n, unsigned long long* rows) {
    const uint64_t i = blockIdx.x * static_cast<uint64_t>(blockDim.x) + threadIdx.x;
    if (i >= n) return;
    const uint4 cu = cursors[i];
    if (cu.x != kInvalidQid && cu.z != 0) atomicAdd(&rows[cu.x], static_cast<unsigned long long>(cu.z));
}
// queries with more rows than the limit -> redo[0 .. tally[0]); tally[1] = cursors the ordered walk can report for them.
// (One reservation per warp: with a low limit nearly every query is on the list, and a million atomics on one counter
// are served one after the other.)  The order of the list does not matter.
__global__ void __launch_bounds__(256) redo_list_kernel(const unsigned long long* rows, uint32_t n_queries, uint32_t max_hits, uint32_t* redo,
                                                        unsigned long long* tally) {
    const uint32_t q = blockIdx.x * blockDim.x + threadIdx.x;
    const bool over = q < n_queries && rows[q] > max_hits;
    const uint32_t m = __ballot_sync(0xffffffffu, over);
    if (m == 0) return;
    const uint32_t lane = threadIdx.x & 31u, n = static_cast<uint32_t>(__popc(m));
    unsigned long long base = 0;
    if (lane == 0) {
        base = atomicAdd(&tally[0], static_cast<unsigned long long>(n));
        atomicAdd(&tally[1], static_cast<unsigned long long>(n) * max_hits);
    }
    base = __shfl_sync(0xffffffffu, base, 0);
    if (over) redo[base + __popc(m & ((1u << lane) - 1u))] = q;
}
// cursors of those queries become empty entries; tally[2] = how many were dropped
__global__ void __launch_bounds__(256) drop_cursors_kernel(uint4* cursors, uint64_t n, const unsigned long long* rows, uint32_t max_hits,
                                                           unsigned long long* tally) {
    const uint64_t i = blockIdx.x * static_cast<uint64_t>(blockDim.x) + threadIdx.x;
    bool drop = false;
    if (i < n) {
        const uint4 cu = cursors[i];
        drop = cu.x != kInvalidQid && rows[cu.x] > max_hits;
        if (drop) cursors[i] = make_uint4(kInvalidQid, 0, 0, 0);
    }
    const uint32_t m = __ballot_sync(0xffffffffu, drop);
    if (m != 0 && (threadIdx.x & 31u) == 0) atomicAdd(&tally[2], static_cast<unsigned long long>(__popc(m)));
}

// search_n: one thread per query, children visited in the order of the reference recursion (fm_ordered_thread)
template <int SIGMA, bool EDIT>
__global__ void __launch_bounds__(256, 4) fm_ordered_kernel(const SearchParams P) {
    extern __shared__ uint32_t s_steps[];
    const uint32_t n_steps = P.n_searches * P.len;
    for (uint32_t i = threadIdx.x; i < n_steps; i += blockDim.x) s_steps[i] = P.steps[i];
    __syncthreads();
    const uint32_t stride = gridDim.x * blockDim.x;
    fm_ordered_thread<SIGMA, EDIT>(P, s_steps, P.ostack + (blockIdx.x * blockDim.x + threadIdx.x), stride, P.ostack_frames);
}

// one thread per query: its live root frames become the work items of fm_items_kernel
__global__ void __launch_bounds__(256) fm_roots_kernel(const SearchParams P) {
    const uint64_t i = blockIdx.x * static_cast<uint64_t>(blockDim.x) + threadIdx.x;
    if (i < P.n_queries) fm_make_items(P, P.steps, static_cast<uint32_t>(i));
}

template <int SIGMA, bool EDIT, int STACK>
// Two blocks of 256 threads per SM: the walk saturates the random-access rate of the memory system with 512 threads
// per SM already, more threads only thrash L1 with their stacks (measured at 3.1 Gbp, fm phase of 2 M queries:
// 4 blocks / 64 registers 1.13 ms, 3 blocks 1.01, 2 blocks / 102 registers and no spills 0.94, 1 block 0.87; at
// 100 Mbp: 0.80 / - / 0.62 / 0.71).
#if !defined(SB200_FM_ITEMS_BLOCKS)
#define SB200_FM_ITEMS_BLOCKS 2
#endif
__global__ void __launch_bounds__(256, SB200_FM_ITEMS_BLOCKS) fm_items_kernel(const SearchParams P) {
    extern __shared__ uint32_t s_steps[];
    const uint32_t n_steps = P.n_searches * P.len;
    for (uint32_t i = threadIdx.x; i < n_steps; i += blockDim.x) s_steps[i] = P.steps[i];
    __syncthreads();
    fm_items_thread<SIGMA, EDIT, STACK>(P, s_steps);
}

template <bool EDIT, int STACK>
#if !defined(SB200_POOL_MINBLOCKS)
#define SB200_POOL_MINBLOCKS 3
#endif
__global__ void __launch_bounds__(kPoolThreads, SB200_POOL_MINBLOCKS) text_pool_kernel(const SearchParams P, const uint32_t maxpush, const uint32_t run_rounds, uint4* spill) {
    extern __shared__ uint32_t s_steps[];
    const uint32_t n_steps = P.n_searches * P.len;
    const uint32_t n_run_words = run_table_bytes(n_steps) / 4;  // run lengths + state flags
    uint32_t* s_runs = s_steps + n_steps;
    for (uint32_t i = threadIdx.x; i < n_steps; i += blockDim.x) s_steps[i] = P.steps[i];
    for (uint32_t i = threadIdx.x; i < n_run_words; i += blockDim.x) s_runs[i] = reinterpret_cast<const uint32_t*>(P.runs)[i];
    // this warp's pool (8-byte aligned region behind the tables)
    const uint32_t lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
    uint8_t* base = reinterpret_cast<uint8_t*>(s_steps) + (((n_steps + n_run_words) * 4u + 7u) & ~7u) + warp * pool_bytes(P.len);
    const TextPool pool = pool_carve(base, spill + (static_cast<size_t>(blockIdx.x) * (blockDim.x >> 5) + warp) * 2u * kSpillCap, P.len);
    if (lane == 0) { *pool.S.top = 0; *pool.R.top = 0; *pool.Pth.top = 0; }
    for (uint32_t i = lane; i < kPoolSlots; i += 32u) pool.live[i] = 0;
    __syncthreads();

    PoolLane ls;
    const unsigned long long slots = P.counters[CT_SEED_SLOTS];
    const uint32_t n_slots = static_cast<uint32_t>(slots < P.seed_cap ? slots : P.seed_cap);
    const uint8_t* runs8 = reinterpret_cast<const uint8_t*>(s_runs);
    bool exhausted = false;  // (warp uniform) the seed list has been handed out
    while (true) {
        uint32_t topS = *pool.S.top, topR = *pool.R.top, topP = *pool.Pth.top;
        if (topS < 32u && topR < 32u && topP < 32u && !exhausted) {  // no full trip: new seeds into the free slots
            constexpr uint32_t kRounds = (kPoolSlots + 31u) / 32u;
            uint32_t n_free = 0;
#pragma unroll
            for (uint32_t j = 0; j < kRounds; ++j)
                n_free += __popc(__ballot_sync(0xffffffffu, j * 32u + lane < kPoolSlots && pool.live[j * 32u + lane] == 0));
            // a refill costs the same instructions for one seed as for 32: wait for kRefillMin free slots unless the pool runs dry
            if (n_free >= kRefillMin || (n_free != 0 && topS + topR + topP < kRefillDry)) {
                uint32_t first = 0;
                if (lane == 0) first = static_cast<uint32_t>(atomicAdd(&P.counters[CT_NEXT_SEED], static_cast<unsigned long long>(n_free)));
                first = __shfl_sync(0xffffffffu, first, 0);
                exhausted = first + n_free >= n_slots;
                const uint32_t W = packed_words(P.len);
#pragma unroll 1
                for (uint32_t j = 0; j < kRounds; ++j) {  // (one copy of pool_load_seed: code size)
                    const bool fr = j * 32u + lane < kPoolSlots && pool.live[j * 32u + lane] == 0;
                    const uint32_t m = __ballot_sync(0xffffffffu, fr);
                    const uint32_t i = first + __popc(m & ((1u << lane) - 1u));
                    first += __popc(m);
                    uint4 seed = make_uint4(kInvalidQid, 0, 0, 0);
                    if (fr && i < n_slots) seed = P.seeds[i];
                    const bool has = seed.x != kInvalidQid;
                    // the queries of this round's seeds, staged by all lanes: word w of a query by lane w
                    for (uint32_t mm = __ballot_sync(0xffffffffu, has); mm; mm &= mm - 1u) {
                        const uint32_t src = static_cast<uint32_t>(__ffs(static_cast<int>(mm)) - 1);
                        const uint32_t qx = __shfl_sync(0xffffffffu, seed.x, src);
                        const uint32_t* from = P.packed + static_cast<uint64_t>(qx) * W;
                        uint32_t* to = pool.query + (j * 32u + src) * pool.Wp;
                        for (uint32_t w = lane; w < W; w += 32u) to[w] = from[w];
                    }
                    if (has) pool_load_seed<false>(P, runs8, pool, j * 32u + lane, seed, ls);
                }
                __syncwarp();
                topS = *pool.S.top;
                topR = *pool.R.top;
                topP = *pool.Pth.top;
            }
        }
        if (topS + topR + topP == 0) {
            if (exhausted) break;
            continue;  // only padding entries were fetched
        }
        uint2 f = make_uint2(0, 0);
        uint32_t slot = 0;
        const PoolTrip trip = pool_pick(topS, topR, topP, maxpush, STACK);
        if (trip.kind == 2u) {
            if (lane < trip.n) {
                slot = pool.Pth.slots[topP - 1 - lane];
                f = pool.Pth.frames[topP - 1 - lane];
            }
            __syncwarp();
            if (lane == 0) *pool.Pth.top = topP - trip.n;
            __syncwarp();
            if (lane < trip.n) pool_retire(pool, slot, text_path<EDIT>(P, s_steps, runs8, pool, f, slot, ls, trip.w));
        } else if (trip.kind == 1u) {
            if (lane < trip.n) f = stack_get(pool.R, topR - 1 - lane, slot);
            __syncwarp();
            if (lane == 0) *pool.R.top = topR - trip.n;
            __syncwarp();
            if (lane < trip.n) pool_retire(pool, slot, text_run<EDIT>(P, s_steps, runs8, pool, f, slot, ls, run_rounds));
        } else {
            if (lane < trip.n) f = stack_get(pool.S, topS - 1 - lane, slot);
            __syncwarp();
            if (lane == 0) *pool.S.top = topS - trip.n;
            __syncwarp();
            if (lane < trip.n) pool_retire(pool, slot, text_states<EDIT>(P, s_steps, runs8, pool, f, slot, ls));
        }
        __syncwarp();
    }
    pool_finish(P, ls, 0);  // (the deepest pool is not tracked: it would be five instructions in every trip for a diagnostic)
}
#endif

}  // namespace sb200
