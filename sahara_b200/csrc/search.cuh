// search.cuh — kernel 2: search-scheme backtracking over the bidirectional FM-index.
//
// Replaces fmc::search_ng24::search<Edit>(index, queries, scheme, delegate) as called at
// /root/reference/src/sahara/search.cpp:227-231 (semantics: SURVEY.md §9.4).  One CUDA thread owns one
// (query, search) pair at a time and walks its search tree depth first with an explicit stack; threads
// are persistent and pull the next pair from a global counter when their stack runs empty.
//
// Per loop iteration every lane performs exactly ONE cursor extension (= the two row probes lb and
// lb+len of one BWT), so warps stay converged at node granularity.  Children are generated in the
// order  match, (deletion, substitution) per symbol, insertion;  the last live child stays in registers
// and becomes the next node, the others are pushed.  Because the match child is generated first it is
// resumed only after all error children of the same node are finished, which bounds the stack by
// 9 * maxErrors frames (Sigma = 6): on the current path only nodes left through an error edge keep
// siblings on the stack, and a path has at most maxErrors error edges.
//
// The order in which cursors are reported differs from the reference's recursion order; the reported
// multiset is identical (the reference does not depend on order, src/sahara/search.cpp:218-220).
#pragma once
#include <cstdio>
#include "layout.cuh"

namespace sb200 {

// packed step entry: pi | l << 16 | u << 20 | right << 24
__host__ __device__ inline uint32_t pack_step(uint32_t pi, uint32_t l, uint32_t u, bool right) {
    return pi | (l << 16) | (u << 20) | (static_cast<uint32_t>(right) << 24);
}

enum : uint32_t { INFO_M = 0, INFO_S = 1, INFO_I = 2, INFO_D = 3 };

// node meta: step (10 bits) | e << 10 (4 bits) | LInfo << 14 | RInfo << 16
__device__ __forceinline__ uint32_t pack_meta(uint32_t step, uint32_t e, uint32_t L, uint32_t R) {
    return step | (e << 10) | (L << 14) | (R << 16);
}

struct SearchParams {
    OccTable bwt, bwtRev;
    uint32_t C[8];
    uint32_t n_rows;
    const uint8_t* queries;  // [n_queries][len] ranks
    const uint32_t* packed;  // [n_queries][packed_words(len)] 4-bit packed copies of the queries
    uint32_t n_queries, len, n_searches;
    const uint32_t* steps;   // [n_searches][len] packed
    uint4* out;              // (qid, lb, len, e)
    uint32_t out_cap;
    // counters: [0] next query, [1] output slots reserved, [2] nodes, [3] stack overflow flag, [5] max stack depth,
    //           [6] cursors reported
    unsigned long long* counters;
    // optional q-gram jump table (cursor after the first qgram_q characters of a search)
    const uint4* qgram;  // [4^q] (lb, lbRev, len, 0)
    uint32_t qgram_q;
    uint32_t debug_flags;  // 1: no pair frames, 2: no insertion chains (diagnostics only)
    // in-text verification (nullptr = off): suffix array, its inverse, and the text packed 8 symbols per word
    const uint32_t* sa32;
    const uint32_t* isa32;
    const uint32_t* text4;
};

// Frames and states.  A stack frame is a cursor plus one search state (step, e, LInfo, RInfo), or — flag
// PAIR — the two states a mismatching symbol produces on the same child cursor: the deletion
// (step, e, side = D) and the substitution (step + 1, e, side = S).  Both extend the same cursor, so one
// probe of the occurrence table serves both.  Likewise the insertion child of a state keeps the cursor
// of its parent: it is processed in the same iteration ("insertion chain") from the ranks already in
// registers.  Every state is expanded exactly like the corresponding call of the reference recursion,
// so the reported multiset is unchanged; only the number of memory probes shrinks.
constexpr uint32_t META_PAIR = 1u << 18;
// In-text verification.  Once a cursor holds a single row its occurrence T[a, b) is unique, and whether a
// child cursor is empty is decided by one text symbol (T[a-1] or T[b]) instead of two rank probes.  Frames
// flagged META_TEXT carry (a, b) in place of (lb, lbRev); the expansion of the states is unchanged, so the
// reported multiset is unchanged.  FM-mode metas carry the number of text symbols consumed so far (tlen,
// bits 20..29) so that b = a + tlen when the frame switches; a = SA[lb], and the reported cursor of a text
// frame is lb = ISA[a], len = 1.
constexpr uint32_t META_TEXT = 1u << 19;
constexpr uint32_t META_TLEN_SHIFT = 20;
constexpr uint32_t kEmitChunk = 16;       // output slots a thread reserves per atomic
constexpr uint32_t kQueryBatch = 2;       // queries a thread takes per atomic
constexpr uint32_t kInvalidQid = 0xffffffffu;

// queries packed 8 symbols per word (4 bits each) by pack_queries_kernel
__host__ __device__ inline uint32_t packed_words(uint32_t len) { return (len + 7) / 8; }

#if !defined(SB200_HOST_EMU)
__global__ void pack_queries_kernel(const uint8_t* q, uint64_t n_queries, uint32_t len, uint32_t* out) {
    uint32_t W = packed_words(len);
    uint64_t i = blockIdx.x * static_cast<uint64_t>(blockDim.x) + threadIdx.x;
    if (i >= n_queries * W) return;
    uint64_t qi = i / W;
    uint32_t w = static_cast<uint32_t>(i % W);
    const uint8_t* src = q + qi * len + w * 8;
    uint32_t v = 0xffffffffu;  // unused nibbles stay 0xF (never a valid symbol)
    for (uint32_t k = 0; k < 8 && w * 8 + k < len; ++k) v = (v & ~(0xfu << (4 * k))) | (static_cast<uint32_t>(src[k] & 0xfu) << (4 * k));
    out[i] = v;
}
#endif

// body of one (persistent) thread.
//   s_steps: the packed scheme table (shared memory on the device)
//   s_query: this thread's staging area for the packed query, element w at s_query[w * qstride]
template <int SIGMA, bool EDIT, int STACK>
__device__ __forceinline__ void search_thread(const SearchParams& P, const uint32_t* s_steps, uint32_t* s_query, uint32_t qstride) {
    uint4 stack[STACK];
    int sp = 0;
    const uint32_t* tbl = nullptr;
    uint32_t qid = 0, qid_end = 0, next_search = P.n_searches;  // forces the first fetch
    uint32_t nodes = 0, emitted = 0;
    uint32_t out_pos = 0, out_end = 0;
    bool overflow = false;
    bool textOK = false;  // in-text verification allowed for the current query
    int maxsp = 0;
    const uint32_t qlen = P.len;
    const uint32_t W = packed_words(qlen);

    auto qsym = [&](uint32_t pos) -> uint32_t { return (s_query[(pos >> 3) * qstride] >> ((pos & 7u) * 4u)) & 0xfu; };

    bool textFrame = false;  // the frame being expanded is in text mode
    auto emit = [&](uint32_t lb, uint32_t len, uint32_t e) {
        if (textFrame) lb = P.isa32[lb];  // (a, b) -> the row of the suffix starting at a; len is 1
        if (out_pos == out_end) {
            out_pos = static_cast<uint32_t>(atomicAdd(&P.counters[1], static_cast<unsigned long long>(kEmitChunk)));
            out_end = out_pos + kEmitChunk;
        }
#if defined(SB200_TRACE)
        if (P.debug_flags & 4u) printf("EMIT q=%u lb=%u len=%u e=%u\n", qid, lb, len, e);
#endif
        if (out_pos < P.out_cap) P.out[out_pos] = make_uint4(qid, lb, len, e);
        ++out_pos;
        ++emitted;
    };

    bool done = false;
    while (true) {
        // ---- make sure there is a frame: next search of the current query, next query, next batch ----
        while (sp == 0) {
            if (next_search == P.n_searches) {
                next_search = 0;
                ++qid;
                if (qid >= qid_end) {
                    unsigned long long w = atomicAdd(&P.counters[0], static_cast<unsigned long long>(kQueryBatch));
                    if (w >= P.n_queries) { done = true; break; }
                    qid = static_cast<uint32_t>(w);
                    qid_end = qid + kQueryBatch < P.n_queries ? qid + kQueryBatch : P.n_queries;
                }
                const uint32_t* src = P.packed + static_cast<uint64_t>(qid) * W;
                bool delim = false;
                for (uint32_t w = 0; w < W; ++w) {
                    uint32_t v = src[w];
                    s_query[w * qstride] = v;
                    delim = delim || (((v - 0x11111111u) & ~v & 0x88888888u) != 0);  // some nibble is 0
                }
                textOK = P.sa32 != nullptr && !delim;
            }
            tbl = s_steps + next_search * qlen;
            ++next_search;
            uint32_t st0 = tbl[0];
            if (((st0 >> 16) & 0xfu) > 1) continue;  // neither a match nor a mismatch allowed at step 0
            uint4 root = make_uint4(0, 0, P.n_rows, 0);
            if (P.qgram_q) {  // q-gram jump: skip the leading steps that allow no error
                uint32_t qq = P.qgram_q;
                bool ok = qq <= qlen;
                uint32_t code = 0;
                bool right0 = (st0 >> 24) & 1u;
                for (uint32_t i = 0; ok && i < qq; ++i) {
                    uint32_t st = tbl[i];
                    uint32_t c = qsym(st & 0xffffu);
                    ok = ((st >> 20) & 0xfu) == 0 && c >= 1 && c <= 4 && (((st >> 24) & 1u) == right0);
                    // the table is keyed by the string in text order, first symbol most significant
                    if (right0) code = (code << 2) | (c - 1);
                    else code |= (c - 1) << (2 * i);
                }
                if (ok) {
                    uint4 g = P.qgram[code];
                    if (g.z == 0) continue;
                    if (qq == qlen) {
                        textFrame = false;
                        emit(g.x, g.z, 0);
                        continue;
                    }
                    if (((tbl[qq] >> 16) & 0xfu) > 1) continue;
                    root = make_uint4(g.x, g.y, g.z, pack_meta(qq, 0, INFO_M, INFO_M) | (qq << META_TLEN_SHIFT));
                }
            }
            stack[0] = root;
            sp = 1;
        }
        if (done) break;
        maxsp = sp > maxsp ? sp : maxsp;
        uint32_t lb, lbRev, len, meta;
        {
            uint4 f = stack[--sp];
            lb = f.x; lbRev = f.y; len = f.z; meta = f.w;
        }

        // ---- one probe for this cursor: occurrence table (FM mode) or one text symbol (text mode) -------
        uint32_t step = meta & 0x3ffu, e = (meta >> 10) & 0xfu;
        uint32_t Linfo = (meta >> 14) & 3u, Rinfo = (meta >> 16) & 3u;
        const bool pair = (meta & META_PAIR) != 0;
        const bool right = (tbl[step] >> 24) & 1u;
        const uint32_t tlen = (meta >> META_TLEN_SHIFT) & 0x3ffu;
        textFrame = (meta & META_TEXT) != 0;
        if (!textFrame && textOK && len == 1) {  // unique occurrence: switch to the text
            uint32_t a = P.sa32[lb];
            lb = a;
            lbRev = a + tlen;
            textFrame = true;
        }
        // child cursors per symbol: (klb[s], klbRev[s], cnt[s])
        uint32_t klb[SIGMA], klbRev[SIGMA], cnt[SIGMA];
        if (textFrame) {
            // lb = a, lbRev = b.  Left extension reads T[a-1] (the delimiter before position 0), right T[b].
            const uint32_t pos = right ? lbRev : lb - 1;
            uint32_t t = 0;
            if (right || lb != 0) t = (P.text4[pos >> 3] >> ((pos & 7u) * 4u)) & 0xfu;
            const uint32_t na = right ? lb : lb - 1, nb = right ? lbRev + 1 : lbRev;
#pragma unroll
            for (int s = 0; s < SIGMA; ++s) {
                klb[s] = na;
                klbRev[s] = nb;
                cnt[s] = (s != 0 && static_cast<uint32_t>(s) == t) ? 1u : 0u;
            }
        } else {
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
            // The probed side continues at C[s] + rank(lo, s), the other side moves by the number of smaller
            // symbols inside the interval.
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
        // flags every child inherits: text mode, and the text length of children that consume a text symbol
        const uint32_t inheritSame = (textFrame ? META_TEXT : 0u) | (tlen << META_TLEN_SHIFT);
        const uint32_t inheritNext = (textFrame ? META_TEXT : 0u) | ((tlen + 1) << META_TLEN_SHIFT);

        // ---- expand every state that lives on this cursor ------------------------------------------
        bool second = false;  // second half of a pair already taken
        while (true) {
            ++nodes;
#if defined(SB200_TRACE)
            if (P.debug_flags & 4u)
                printf("STATE q=%u lb=%u lbRev=%u len=%u step=%u e=%u L=%u R=%u pair=%d second=%d right=%d cnt=%u,%u,%u,%u,%u,%u\n", qid, lb,
                       lbRev, len, step, e, Linfo, Rinfo, (int)pair, (int)second, (int)right, cnt[0], cnt[1], cnt[2], cnt[3], cnt[4],
                       SIGMA > 5 ? cnt[SIGMA - 1] : 0u);
#endif
            const uint32_t st = tbl[step];
            const uint32_t l = (st >> 16) & 0xfu, u = (st >> 20) & 0xfu;
            const uint32_t c = qsym(st & 0xffffu);
            const bool last = step + 1 == qlen;
            const uint32_t stn = last ? 0u : tbl[step + 1];
            const uint32_t lnext = (stn >> 16) & 0xfu;
            const bool rightNext = (stn >> 24) & 1u;
            const bool sameDirNext = !last && (rightNext == right);
            const bool matchOK = l <= e && e <= u;
            const bool mmOK = l <= e + 1 && e + 1 <= u;
            const uint32_t T = right ? Rinfo : Linfo;
            const uint32_t O = right ? Linfo : Rinfo;  // info of the other end
            const bool otherEndOK = !EDIT || (O & 1u) == 0;  // M or I
            // metas of the possible children: the moving side gets the new info
            const uint32_t keepL = right ? Linfo : 0u, keepR = right ? 0u : Rinfo;
            const uint32_t sideShift = right ? 16u : 14u;
            const uint32_t metaBase = (keepL << 14) | (keepR << 16);
            const uint32_t mM = metaBase | (step + 1) | (e << 10) | (INFO_M << sideShift) | inheritNext;
            const uint32_t mD = metaBase | step | ((e + 1) << 10) | (INFO_D << sideShift) | inheritNext;
            const uint32_t mS = metaBase | (step + 1) | ((e + 1) << 10) | (INFO_S << sideShift) | inheritNext;
            const uint32_t mI = metaBase | (step + 1) | ((e + 1) << 10) | (INFO_I << sideShift) | inheritSame;
            // match
            {
                uint32_t mc = 0, nlb = 0, nlbRev = 0;
#pragma unroll
                for (int s = 0; s < SIGMA; ++s)
                    if (static_cast<uint32_t>(s) == c) { mc = cnt[s]; nlb = klb[s]; nlbRev = klbRev[s]; }
                const bool alive = matchOK && mc != 0;
                if (alive && last) {
                    if (otherEndOK) emit(nlb, mc, e);
                } else if (alive && lnext <= e + 1) {
                    if (sp < STACK) stack[sp] = make_uint4(nlb, nlbRev, mc, mM);
                    else overflow = true;
                    ++sp;
                }
            }
            if (mmOK) {
                const bool delOK = EDIT && (T == INFO_M || T == INFO_D);
                const bool subAlive = !last && lnext <= e + 2;
                const bool asPair = delOK && subAlive && sameDirNext && !(P.debug_flags & 1u);
#pragma unroll
                for (int s = 1; s < SIGMA; ++s) {
                    const bool live = static_cast<uint32_t>(s) != c && cnt[s] != 0;
                    const uint32_t nlb = klb[s], nlbRev = klbRev[s];
                    if (live && (asPair || delOK)) {
                        if (sp < STACK) stack[sp] = make_uint4(nlb, nlbRev, cnt[s], asPair ? (mD | META_PAIR) : mD);
                        else overflow = true;
                        ++sp;
                    }
                    if (live && !asPair && subAlive) {
                        if (sp < STACK) stack[sp] = make_uint4(nlb, nlbRev, cnt[s], mS);
                        else overflow = true;
                        ++sp;
                        }
                    if (!EDIT && live && last) emit(nlb, cnt[s], e + 1);
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
            const bool insOK = EDIT && mmOK && (T == INFO_M || T == INFO_I);
            if (!insOK) break;
            if (last) {
                if (otherEndOK) emit(lb, len, e + 1);
                break;
            }
            if (lnext > e + 2) break;  // dead at the next step
            if (!sameDirNext || (P.debug_flags & 2u)) {  // direction changes: needs a probe of the other table
                if (sp < STACK) stack[sp] = make_uint4(lb, lbRev, len, mI);
                else overflow = true;
                ++sp;
                break;
            }
            step += 1;
            e += 1;
            if (right) Rinfo = INFO_I; else Linfo = INFO_I;
        }
    }
    // unused slots of the last reserved chunk become empty entries (len 0: they locate to nothing)
    for (; out_pos < out_end; ++out_pos)
        if (out_pos < P.out_cap) P.out[out_pos] = make_uint4(kInvalidQid, 0, 0, 0);
    if (nodes) atomicAdd(&P.counters[2], static_cast<unsigned long long>(nodes));
    if (overflow) atomicExch(&P.counters[3], 1ull);
    atomicMax(&P.counters[5], static_cast<unsigned long long>(maxsp));
    if (emitted) atomicAdd(&P.counters[6], static_cast<unsigned long long>(emitted));
}

#if !defined(SB200_HOST_EMU)
template <int SIGMA, bool EDIT, int STACK>
__global__ void __launch_bounds__(256, 4) search_kernel(const SearchParams P) {
    extern __shared__ uint32_t s_steps[];
    const uint32_t n_steps = P.n_searches * P.len;
    for (uint32_t i = threadIdx.x; i < n_steps; i += blockDim.x) s_steps[i] = P.steps[i];
    __syncthreads();
    // word w of this thread's query lives at s_query[w * blockDim.x]: every lane stays in its own bank
    search_thread<SIGMA, EDIT, STACK>(P, s_steps, s_steps + n_steps + threadIdx.x, blockDim.x);
}
#endif

}  // namespace sb200
