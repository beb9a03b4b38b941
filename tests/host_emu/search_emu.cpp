// search_emu.cpp — host emulation of the CUDA search kernel (test infrastructure).
// Compiles sahara_b200/csrc/search.cuh with g++ (SB200_HOST_EMU) and runs the per-thread body as a single
// thread over a device-layout index built on the host from BWT symbols.  Lets the `-m "not gpu"` tests
// compare the real kernel source with the oracle.
#define SB200_HOST_EMU 1
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <algorithm>
#include <cstring>
#include <vector>

#include "../../sahara_b200/csrc/search.cuh"

using namespace sb200;

namespace {
struct HostOcc {
    std::vector<OccBlk> blk;
    std::vector<OccSup> sup;
    void build(const uint8_t* bwt, uint64_t n) {
        uint64_t nb = n / 64 + 1, ns = nb / 64 + 1;
        blk.assign(nb + 1, OccBlk{0, 0, 0, 0});
        sup.assign(ns + 1, OccSup{});
        uint32_t abs[8] = {0}, rel[8] = {0};
        for (uint64_t b = 0; b < nb; ++b) {
            if (b % 64 == 0) {
                for (int c = 0; c < 8; ++c) { sup[b / 64].c[c] = abs[c]; rel[c] = 0; }
            }
            OccBlk& o = blk[b];
            for (int s = 1; s < 6; ++s) o.ctr |= uint64_t(rel[s] & 0xfff) << (12 * (s - 1));
            for (uint64_t r = b * 64; r < (b + 1) * 64 && r < n; ++r) {
                uint32_t c = bwt[r];
                o.p0 |= uint64_t(c & 1) << (r & 63);
                o.p1 |= uint64_t((c >> 1) & 1) << (r & 63);
                o.p2 |= uint64_t((c >> 2) & 1) << (r & 63);
                abs[c]++; rel[c]++;
            }
        }
    }
};

// trip statistics of the emulated warp (EMU_STATS=1 in the environment prints them): how full the trips of each kind are
struct TripStats {
    uint64_t trips[3]{}, lanes[3]{}, iters[3]{}, pushed[3]{};
    uint64_t max_lane_iters[3]{};  // sum over trips of the longest lane (what the warp waits for)
} g_stats;

// text_pool_kernel as one warp of 32 lanes in lockstep: the pops of a trip all read the pool before any lane
// expands (what the __syncwarp()s of the kernel guarantee), then the lanes expand one after the other.
template <bool EDIT>
void run_text_pool(const SearchParams& P, const uint32_t* steps, const uint8_t* runs, uint32_t kmax) {
    constexpr uint32_t STACK = 96, LANES = 32;
    std::vector<uint64_t> smem(pool_bytes(P.len) / 8 + 1, 0);
    std::vector<uint4> spill(2 * kSpillCap);
    const TextPool pool = pool_carve(reinterpret_cast<uint8_t*>(smem.data()), spill.data(), P.len);
    PoolLane lanes[LANES];
    const uint32_t maxpush = 2 * (kmax + 1);
    const unsigned long long slots_total = P.counters[CT_SEED_SLOTS];
    const uint32_t n_slots = static_cast<uint32_t>(slots_total < P.seed_cap ? slots_total : P.seed_cap);
    uint32_t maxtop = 0;
    bool exhausted = false;
    uint32_t &topS = *pool.S.top, &topR = *pool.R.top, &topP = *pool.Pth.top;
    while (true) {
        if (topS < LANES && topR < LANES && topP < LANES && !exhausted) {
            std::vector<uint32_t> free_slots;
            for (uint32_t sl = 0; sl < kPoolSlots; ++sl)
                if (pool.live[sl] == 0) free_slots.push_back(sl);
            const uint32_t first = free_slots.empty() ? 0u : static_cast<uint32_t>(atomicAdd(&P.counters[CT_NEXT_SEED], static_cast<unsigned long long>(free_slots.size())));
            if (!free_slots.empty()) exhausted = first + free_slots.size() >= n_slots;
            for (uint32_t r = 0; r < free_slots.size(); ++r) {
                const uint32_t i = first + r;
                if (i < n_slots && P.seeds[i].x != kInvalidQid)
                    pool_load_seed<true>(P, runs, pool, free_slots[r], P.seeds[i], lanes[free_slots[r] % LANES]);
            }
        }
        if (topS + topR + topP == 0) {
            if (exhausted) break;
            continue;
        }
        maxtop = std::max(maxtop, topS + topR + topP);
        uint2 f[LANES];
        uint32_t sl[LANES];
        const PoolTrip trip = pool_pick(topS, topR, topP, maxpush, STACK);
        if (trip.n == 0) {  // cannot happen: pool_pick always finds a stack to pop
            atomicExch(&P.counters[CT_OVERFLOW], 1ull);
            break;
        }
        const uint32_t n = trip.n;
        FrameStack const& st = trip.kind == 2u ? pool.Pth : trip.kind == 1u ? pool.R : pool.S;
        uint32_t& top = *st.top;
        for (uint32_t lane = 0; lane < n; ++lane) f[lane] = stack_get(st, top - 1 - lane, sl[lane]);
        top -= n;
        g_stats.trips[trip.kind]++;
        g_stats.lanes[trip.kind] += n;
        uint32_t longest = 0;
        for (uint32_t lane = 0; lane < n; ++lane) {
            const uint32_t n0 = lanes[lane].nodes;
            uint32_t pushed;
            if (trip.kind == 2u) pushed = text_path<EDIT>(P, steps, runs, pool, f[lane], sl[lane], lanes[lane], trip.w);
            else if (trip.kind == 1u) pushed = text_run<EDIT>(P, steps, runs, pool, f[lane], sl[lane], lanes[lane], kRunRounds);
            else pushed = text_states<EDIT>(P, steps, runs, pool, f[lane], sl[lane], lanes[lane]);
            pool_retire(pool, sl[lane], pushed);
            g_stats.iters[trip.kind] += lanes[lane].nodes - n0;
            g_stats.pushed[trip.kind] += pushed;
            longest = std::max(longest, lanes[lane].nodes - n0);
        }
        g_stats.max_lane_iters[trip.kind] += longest;
    }
    for (uint32_t sl2 = 0; sl2 < kPoolSlots; ++sl2)
        if (pool.live[sl2] != 0) atomicExch(&P.counters[CT_OVERFLOW], 1ull);  // a seed was lost: report as failure
    for (auto& ls : lanes) pool_finish(P, ls, maxtop);
    if (std::getenv("EMU_STATS")) {
        const char* names[3] = {"state", "run", "path"};
        for (int t = 0; t < 3; ++t)
            if (g_stats.trips[t])
                std::fprintf(stderr, "%-5s trips %9llu  lanes/trip %5.1f  nodes/trip %6.1f  longest-lane nodes/trip %5.2f  pushed/frame %4.2f\n", names[t],
                             (unsigned long long)g_stats.trips[t], double(g_stats.lanes[t]) / g_stats.trips[t], double(g_stats.iters[t]) / g_stats.trips[t],
                             double(g_stats.max_lane_iters[t]) / g_stats.trips[t], double(g_stats.pushed[t]) / g_stats.lanes[t]);
        g_stats = TripStats{};
    }
}
}  // namespace

uint32_t g_max_hits = 0;  // > 0: emu_search runs the ordered walk with a hit limit (fm_ordered_kernel)
uint32_t g_ordered_maxsp = 0;
sb200_policy g_policy = SB200_POLICY_DEFAULT;  // the table of reconstructed rules the kernels get (emu_set_policy)

extern "C" {

void emu_set_policy(const sb200_policy* p) {
    sb200_policy def = SB200_POLICY_DEFAULT;
    g_policy = p ? *p : def;
}

void emu_set_max_hits(uint32_t n) { g_max_hits = n; }
uint32_t emu_ordered_max_depth() { return g_ordered_maxsp; }

// bwt / bwtRev: symbols per row; C: sigma+1 entries; scheme tables as for sb200_set_scheme.
// out: library-allocated (qid, lb, len, e) u32 quadruples, release with emu_free.
int emu_search(const uint8_t* bwt, const uint8_t* bwtRev, uint64_t n_rows, int sigma, const uint64_t* C, const uint8_t* queries,
               uint64_t n_queries, uint32_t len, uint32_t n_searches, const uint16_t* pi, const uint8_t* l, const uint8_t* u, int edit,
               uint32_t debug_flags, const uint32_t* sa32, const uint8_t* text, uint32_t** out, uint64_t* n_out, uint64_t* nodes) {
    HostOcc a, b;
    a.build(bwt, n_rows);
    b.build(bwtRev, n_rows);
    std::vector<uint32_t> steps(size_t(n_searches) * len);
    uint32_t kmax = 0;
    for (uint32_t j = 0; j < n_searches; ++j)
        for (uint32_t i = 0; i < len; ++i) {
            size_t k = size_t(j) * len + i;
            bool right = (i == 0) ? (len < 2 || pi[k] < pi[k + 1]) : (pi[k - 1] < pi[k]);
            steps[k] = pack_step(pi[k], l[k], u[k], right);
            if (u[k] > kmax) kmax = u[k];
        }
    if (kmax > 4) return 2;
    std::vector<uint8_t> runs(run_table_bytes(static_cast<uint32_t>(steps.size())) + 4, 0);
    build_runs(n_searches, len, steps.data(), runs.data());
    build_state_flags(n_searches, len, steps.data(), runs.data(), g_policy);
    build_path_windows(n_searches, len, steps.data(), runs.data());
    // optional in-text verification tables
    std::vector<uint32_t> isa, text4;
    if (sa32 && text) {
        isa.assign(n_rows, 0);
        for (uint64_t r = 0; r < n_rows; ++r) isa[sa32[r]] = static_cast<uint32_t>(r);
        text4.assign(n_rows / 8 + 2, 0);
        for (uint64_t i = 0; i < n_rows; ++i) text4[i / 8] |= uint32_t(text[i] & 0xf) << (4 * (i % 8));
    }
    uint32_t W = packed_words(len);
    std::vector<uint32_t> packed(size_t(n_queries) * W, 0xffffffffu);
    for (uint64_t qi = 0; qi < n_queries; ++qi)
        for (uint32_t i = 0; i < len; ++i) {
            uint32_t& w = packed[qi * W + i / 8];
            w = (w & ~(0xfu << (4 * (i % 8)))) | (uint32_t(queries[qi * len + i] & 0xf) << (4 * (i % 8)));
        }
    // optional q-gram jump table (debug_flags bits 8..11 = q): cursor of every A/C/G/T string of length q, first
    // symbol most significant, made by right extensions over bwtRev (what qgram_level_kernel does on the device)
    const uint32_t qgram_q = (debug_flags >> 8) & 0xfu;
    std::vector<uint4> qgram;
    if (qgram_q) {
        qgram.assign(size_t(1) << (2 * qgram_q), uint4{0, 0, 0, 0});
        const OccTable rev{b.blk.data(), b.sup.data()};
        for (uint32_t code = 0; code < qgram.size(); ++code) {
            uint32_t lb = 0, lbRev = 0, ln = static_cast<uint32_t>(n_rows);
            for (uint32_t i = 0; i < qgram_q && ln; ++i) {
                const uint32_t c = ((code >> (2 * (qgram_q - 1 - i))) & 3u) + 1;
                uint32_t r1[8] = {0}, r2[8] = {0};
                all_ranks<6>(rev, lbRev, r1);
                all_ranks<6>(rev, lbRev + ln, r2);
                uint32_t smaller = 0;
                for (uint32_t t = 0; t < c; ++t) smaller += r2[t] - r1[t];
                lb += smaller;
                lbRev = static_cast<uint32_t>(C[c]) + r1[c];
                ln = r2[c] - r1[c];
            }
            qgram[code] = uint4{lb, lbRev, ln, 0};
        }
    }
    std::vector<uint4> items(size_t(n_queries) * n_searches + 1);
    std::vector<uint2> item_tags(size_t(n_queries) * n_searches + 1);
    uint64_t cap = 1 << 16, seed_cap = 1 << 16;
    std::vector<uint4> buf, seeds;
    unsigned long long counters[CT_COUNT];
    while (true) {
        buf.assign(cap + 1, uint4{0, 0, 0, 0});
        seeds.assign(seed_cap + 1, uint4{0, 0, 0, 0});
        std::memset(counters, 0, sizeof counters);
        SearchParams P{};
        P.bwt = OccTable{a.blk.data(), a.sup.data()};
        P.bwtRev = OccTable{b.blk.data(), b.sup.data()};
        for (int i = 0; i < 8; ++i) P.C[i] = static_cast<uint32_t>(i <= sigma ? C[i] : n_rows);
        P.n_rows = static_cast<uint32_t>(n_rows);
        P.packed = packed.data();
        P.seeds = seeds.data();
        P.seed_cap = static_cast<uint32_t>(seed_cap);
        P.n_queries = static_cast<uint32_t>(n_queries);
        P.len = len;
        P.n_searches = n_searches;
        P.steps = steps.data();
        P.runs = runs.data();
        P.out = buf.data();
        P.out_cap = static_cast<uint32_t>(cap);
        P.counters = counters;
        P.qgram = qgram_q ? qgram.data() : nullptr;
        P.qgram_q = qgram_q;
        P.debug_flags = debug_flags & 0xffu;
        P.pol = g_policy;
        P.sa32 = isa.empty() ? nullptr : sa32;
        P.isa32 = isa.empty() ? nullptr : isa.data();
        P.text4 = isa.empty() ? nullptr : text4.data();
        if (g_max_hits) {  // search_n: fm_ordered_kernel as a single thread
            P.max_hits = g_max_hits;
            if (P.qgram_q >= len) P.qgram = nullptr, P.qgram_q = 0;
            std::vector<uint4> ostack(ordered_stack_frames(len, static_cast<uint32_t>(sigma)));
            P.ostack = ostack.data();
            P.ostack_frames = static_cast<uint32_t>(ostack.size());
            if (sigma == 6) {
                if (edit) fm_ordered_thread<6, true>(P, steps.data(), ostack.data(), 1, P.ostack_frames);
                else fm_ordered_thread<6, false>(P, steps.data(), ostack.data(), 1, P.ostack_frames);
            } else if (sigma == 5) {
                if (edit) fm_ordered_thread<5, true>(P, steps.data(), ostack.data(), 1, P.ostack_frames);
                else fm_ordered_thread<5, false>(P, steps.data(), ostack.data(), 1, P.ostack_frames);
            } else return 3;
            g_ordered_maxsp = static_cast<uint32_t>(counters[CT_MAX_SP]);
            P.sa32 = nullptr;  // (no second kernel)
        } else {  // fm_roots_kernel as a host loop, then fm_items_kernel as a one-lane warp
            if (P.qgram_q >= len) P.qgram = nullptr, P.qgram_q = 0;
            P.items = items.data();
            P.item_tags = item_tags.data();
            for (uint64_t i = 0; i < n_queries; ++i) fm_make_items(P, steps.data(), static_cast<uint32_t>(i));
            if (sigma == 6) {
                if (edit) fm_items_thread<6, true, 96>(P, steps.data());
                else fm_items_thread<6, false, 96>(P, steps.data());
            } else if (sigma == 5) {
                if (edit) fm_items_thread<5, true, 96>(P, steps.data());
                else fm_items_thread<5, false, 96>(P, steps.data());
            } else return 3;
        }
        if (counters[CT_SEED_SLOTS] > seed_cap) {
            seed_cap = counters[CT_SEED_SLOTS];
            continue;
        }
        if (P.sa32) {  // second kernel (text_pool_kernel): one warp of 32 lanes in lockstep
            if (edit) run_text_pool<true>(P, steps.data(), runs.data(), kmax);
            else run_text_pool<false>(P, steps.data(), runs.data(), kmax);
        }
        if (counters[CT_OVERFLOW]) return 4;  // stack overflow
        if (counters[CT_OUT_SLOTS] <= cap) break;
        cap = counters[CT_OUT_SLOTS];
    }
    uint64_t slots = counters[CT_OUT_SLOTS], n = 0;
    *out = static_cast<uint32_t*>(std::malloc(std::max<uint64_t>(1, slots) * 16));
    for (uint64_t i = 0; i < slots; ++i)
        if (buf[i].x != kInvalidQid) std::memcpy(*out + 4 * n++, &buf[i], 16);
    if (n != counters[CT_CURSORS]) return 5;
    *n_out = n;
    if (nodes) *nodes = counters[CT_NODES];
    return 0;
}

void emu_free(void* p) { std::free(p); }
}
