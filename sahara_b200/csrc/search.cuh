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
    uint32_t n_queries, len, n_searches;
    const uint32_t* steps;   // [n_searches][len] packed
    uint4* out;              // (qid, lb, len, e)
    uint32_t out_cap;
    // counters: [0] next work item, [1] cursors reported, [2] nodes, [3] stack overflow flag
    unsigned long long* counters;
    // optional q-gram jump table (cursor after the first qgram_q characters of a search)
    const uint4* qgram;  // [4^q] (lb, lbRev, len, 0)
    uint32_t qgram_q;
};

template <int SIGMA, bool EDIT, int STACK>
__global__ void __launch_bounds__(256) search_kernel(const SearchParams P) {
    extern __shared__ uint32_t s_steps[];
    for (uint32_t i = threadIdx.x; i < P.n_searches * P.len; i += blockDim.x) s_steps[i] = P.steps[i];
    __syncthreads();

    uint4 stack[STACK];
    int sp = 0;
    uint32_t lb = 0, lbRev = 0, len = 0, meta = 0;
    bool have = false;
    const uint8_t* q = nullptr;
    const uint32_t* tbl = nullptr;
    uint32_t qid = 0;
    uint32_t nodes = 0;
    bool overflow = false;
    const uint32_t total_items = P.n_queries * P.n_searches;
    const uint32_t qlen = P.len;

    while (true) {
        if (!have) {
            if (sp > 0) {
                uint4 f = stack[--sp];
                lb = f.x; lbRev = f.y; len = f.z; meta = f.w;
            } else {
                uint32_t w = static_cast<uint32_t>(atomicAdd(&P.counters[0], 1ull));
                if (w >= total_items) break;
                qid = w / P.n_searches;
                tbl = s_steps + (w % P.n_searches) * qlen;
                q = P.queries + static_cast<uint64_t>(qid) * qlen;
                lb = 0; lbRev = 0; len = P.n_rows; meta = 0;
                uint32_t st0 = tbl[0];
                if (((st0 >> 16) & 0xf) > 1) continue;  // neither a match nor a mismatch allowed at step 0
                // q-gram jump: skip the leading steps that allow no error
                if (P.qgram_q) {
                    uint32_t qq = P.qgram_q;
                    bool ok = qq <= qlen;
                    uint32_t code = 0;
                    bool right0 = (st0 >> 24) & 1;
                    for (uint32_t i = 0; ok && i < qq; ++i) {
                        uint32_t st = tbl[i];
                        ok = ((st >> 20) & 0xf) == 0;  // u == 0 (then l == 0 as well)
                        uint32_t c = q[st & 0xffff];
                        ok = ok && c >= 1 && c <= 4;
                        ok = ok && (((st >> 24) & 1) == right0);
                        // table is keyed by the string in text order
                        if (right0) code = (code << 2) | (c - 1);
                        else code |= (c - 1) << (2 * i);
                    }
                    if (ok) {
                        uint4 g = P.qgram[code];
                        if (g.z == 0) continue;
                        lb = g.x; lbRev = g.y; len = g.z;
                        uint32_t L = INFO_M, R = INFO_M;
                        if (qq == qlen) {
                            uint32_t idx = static_cast<uint32_t>(atomicAdd(&P.counters[1], 1ull));
                            if (idx < P.out_cap) P.out[idx] = make_uint4(qid, lb, len, 0);
                            continue;
                        }
                        uint32_t st = tbl[qq];
                        if (((st >> 16) & 0xf) > 1) continue;
                        meta = pack_meta(qq, 0, L, R);
                    }
                }
            }
        }
        have = false;

        // ---- one cursor extension ---------------------------------------------------------------
        const uint32_t step = meta & 0x3ffu, e = (meta >> 10) & 0xfu;
        const uint32_t Linfo = (meta >> 14) & 3u, Rinfo = (meta >> 16) & 3u;
        const uint32_t st = tbl[step];
        const uint32_t l = (st >> 16) & 0xfu, u = (st >> 20) & 0xfu;
        const bool right = (st >> 24) & 1u;
        const uint32_t c = q[st & 0xffffu];
        const bool matchOK = l <= e && e <= u;
        const bool mmOK = l <= e + 1 && e + 1 <= u;
        const uint32_t T = right ? Rinfo : Linfo;

        const OccTable& tab = right ? P.bwtRev : P.bwt;
        const uint32_t lo = right ? lbRev : lb;
        const uint32_t hi = lo + len;
        const bool sameBlk = (lo >> kBlkShift) == (hi >> kBlkShift);
        OccBlk b1 = load_blk(tab.blk + (lo >> kBlkShift));
        OccSup s1 = load_sup(tab.sup + (lo >> kSupShift));
        OccBlk b2 = b1;
        OccSup s2 = s1;
        if (!sameBlk) {
            b2 = load_blk(tab.blk + (hi >> kBlkShift));
            if ((lo >> kSupShift) != (hi >> kSupShift)) s2 = load_sup(tab.sup + (hi >> kSupShift));
        }
        ++nodes;

        uint32_t r1[SIGMA], cnt[SIGMA];
        {
            uint32_t sum1 = 0, sumc = 0;
#pragma unroll
            for (int s = 1; s < SIGMA; ++s) {
                uint32_t a = s1.c[s] + blk_ctr(b1, s) + blk_count(b1, lo & 63u, s);
                uint32_t b = s2.c[s] + blk_ctr(b2, s) + blk_count(b2, hi & 63u, s);
                r1[s] = a;
                cnt[s] = b - a;
                sum1 += a;
                sumc += b - a;
            }
            r1[0] = lo - sum1;
            cnt[0] = len - sumc;
        }

        // pending child kept in registers
        uint32_t plb = 0, plbRev = 0, plen = 0, pmeta = 0;
        bool phave = false;

        auto offer = [&](uint32_t nlb, uint32_t nlbRev, uint32_t nlen, uint32_t nstep, uint32_t ne, uint32_t nL,
                         uint32_t nR) {
            if (nlen == 0) return;
            if (nstep == qlen) {
                if (!EDIT || (((nL | nR) & 1u) == 0)) {  // both ends M or I
                    uint32_t idx = static_cast<uint32_t>(atomicAdd(&P.counters[1], 1ull));
                    if (idx < P.out_cap) P.out[idx] = make_uint4(qid, nlb, nlen, ne);
                }
                return;
            }
            uint32_t st2 = tbl[nstep];
            uint32_t l2 = (st2 >> 16) & 0xfu, u2 = (st2 >> 20) & 0xfu;
            if (!(ne <= u2 && l2 <= ne + 1)) return;  // neither match nor mismatch possible there
            if (phave) {
                if (sp < STACK) stack[sp++] = make_uint4(plb, plbRev, plen, pmeta);
                else overflow = true;
            }
            plb = nlb; plbRev = nlbRev; plen = nlen; pmeta = pack_meta(nstep, ne, nL, nR);
            phave = true;
        };

        // child cursor for symbol s
        uint32_t smaller = 0;  // occurrences of symbols < s inside the interval
        uint32_t mLb = 0, mLbRev = 0, mLen = 0;
        const bool delOK = EDIT && (T == INFO_M || T == INFO_D);
        const bool insOK = EDIT && (T == INFO_M || T == INFO_I);
        // match first
        {
            uint32_t sm = 0;
#pragma unroll
            for (int s = 0; s < SIGMA; ++s) {
                if (static_cast<uint32_t>(s) == c) {
                    uint32_t own = P.C[s] + r1[s];
                    mLb = right ? lb + sm : own;
                    mLbRev = right ? own : lbRev + sm;
                    mLen = cnt[s];
                }
                sm += cnt[s];
            }
        }
        if (matchOK) {
            offer(mLb, mLbRev, mLen, step + 1, e, right ? Linfo : INFO_M, right ? INFO_M : Rinfo);
        }
        if (mmOK) {
            smaller = cnt[0];
#pragma unroll
            for (int s = 1; s < SIGMA; ++s) {
                if (static_cast<uint32_t>(s) != c && cnt[s] != 0) {
                    uint32_t own = P.C[s] + r1[s];
                    uint32_t nlb = right ? lb + smaller : own;
                    uint32_t nlbRev = right ? own : lbRev + smaller;
                    if (delOK) offer(nlb, nlbRev, cnt[s], step, e + 1, right ? Linfo : INFO_D, right ? INFO_D : Rinfo);
                    offer(nlb, nlbRev, cnt[s], step + 1, e + 1, right ? Linfo : INFO_S, right ? INFO_S : Rinfo);
                }
                smaller += cnt[s];
            }
            if (insOK) offer(lb, lbRev, len, step + 1, e + 1, right ? Linfo : INFO_I, right ? INFO_I : Rinfo);
        }
        if (phave) {
            lb = plb; lbRev = plbRev; len = plen; meta = pmeta;
            have = true;
        }
    }
    if (nodes) atomicAdd(&P.counters[2], static_cast<unsigned long long>(nodes));
    if (overflow) atomicExch(&P.counters[3], 1ull);
}

}  // namespace sb200
