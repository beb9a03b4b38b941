// capi.cu — implementation of the C ABI declared in include/sahara_b200.h.
//
// Host orchestration of the three kernels (rank / search / locate) for one GPU.  There is no CPU
// fallback: every entry point needs a CUDA device and fails loudly otherwise.
#include "../../include/sahara_b200.h"

#include <algorithm>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <stdexcept>
#include <string>
#include <vector>

#include <cub/cub.cuh>
#include <cuda_runtime.h>

#include "build.cuh"
#include "index.cuh"
#include "layout.cuh"
#include "locate.cuh"
#include "search.cuh"
#include "synth.cuh"

namespace {

using namespace sb200;

thread_local std::string g_err;

struct Error : std::runtime_error {
    using std::runtime_error::runtime_error;
};

#define CUDA_TRY(expr)                                                                                       \
    do {                                                                                                     \
        cudaError_t e__ = (expr);                                                                            \
        if (e__ != cudaSuccess)                                                                              \
            throw Error(std::string("CUDA error: ") + cudaGetErrorString(e__) + " at " + __FILE__ + ":" +   \
                        std::to_string(__LINE__) + " (" #expr ")");                                          \
    } while (0)

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

// device allocation that grows on demand; owning and move-only (released on scope exit, also when an error unwinds)
struct DevBuf {
    void* p{};
    size_t cap{};
    DevBuf() = default;
    DevBuf(const DevBuf&) = delete;
    DevBuf& operator=(const DevBuf&) = delete;
    DevBuf(DevBuf&& o) noexcept : p(o.p), cap(o.cap) { o.p = nullptr; o.cap = 0; }
    DevBuf& operator=(DevBuf&& o) noexcept {
        if (this != &o) {
            release();
            p = o.p;
            cap = o.cap;
            o.p = nullptr;
            o.cap = 0;
        }
        return *this;
    }
    ~DevBuf() { release(); }
    template <typename T = void>
    T* get() const { return static_cast<T*>(p); }
    void reserve(size_t bytes) {
        if (bytes <= cap) return;
        release();
        size_t want = bytes + bytes / 8 + 256;
        CUDA_TRY(cudaMalloc(&p, want));
        cap = want;
    }
    void release() {
        if (p) cudaFree(p);
        p = nullptr;
        cap = 0;
    }
};

// pinned host blocks handed out as results and recycled by sb200_free
struct PinnedPool {
    std::mutex mu;
    struct Blk { void* p; size_t cap; bool used; };
    std::vector<Blk> blocks;
    void* alloc(size_t bytes) {
        std::lock_guard<std::mutex> lk(mu);
        if (bytes == 0) bytes = 1;
        Blk* best = nullptr;
        for (auto& b : blocks)
            if (!b.used && b.cap >= bytes && (!best || b.cap < best->cap)) best = &b;
        if (best && best->cap <= 4 * bytes + (1 << 20)) {
            best->used = true;
            return best->p;
        }
        // drop unused blocks that are too small to be useful before growing
        for (size_t i = 0; i < blocks.size();) {
            if (!blocks[i].used && blocks[i].cap < bytes) {
                cudaFreeHost(blocks[i].p);
                blocks.erase(blocks.begin() + i);
            } else ++i;
        }
        size_t want = bytes + bytes / 4;
        void* p = nullptr;
        if (cudaHostAlloc(&p, want, cudaHostAllocDefault) != cudaSuccess) {
            cudaGetLastError();
            throw Error("cannot allocate " + std::to_string(want) + " bytes of pinned host memory");
        }
        blocks.push_back(Blk{p, want, true});
        return p;
    }
    bool free(void* p) {
        std::lock_guard<std::mutex> lk(mu);
        for (auto& b : blocks)
            if (b.p == p) {
                b.used = false;
                return true;
            }
        return false;
    }
};
PinnedPool g_pinned;

struct DeviceIndex {
    bool loaded{};
    uint32_t sigma{};
    uint64_t n_rows{}, n_blocks{}, n_sup{};
    DevBuf bwt_blk, bwt_sup, rev_blk, rev_sup;
    uint32_t C[8]{};
    uint64_t C64[8]{};
    DevBuf d_C;             // u32[8] copy for kernels that take a pointer
    DevBuf marks;           // MarkRec[n_mark]
    uint64_t n_mark{};
    bool full_sa{};         // device rate 1: ssa indexed by row, no marks
    DevBuf ssa;             // u64
    uint64_t n_ssa{};
    uint64_t sampling_rate{}, bits_for_position{}, device_rate{};
    uint32_t key_bits{};    // significant bits of an ssa value
    // reference image of the marks / ssa as uploaded or built (kept for download when densified)
    DevBuf ref_mark_words;  // u64[n_rows/64+1]
    DevBuf ref_ssa;         // u64[n_ref_ssa]   (only when densified; otherwise == ssa)
    uint64_t n_ref_ssa{};
    DevBuf qgram;
    uint32_t qgram_q{};
    DevBuf sa32, isa32, text4;  // in-text verification tables
    DevBuf seq_start;           // u64[n_seqs + 1]: start of every sequence in the delimited text (with the tables above)
    uint64_t n_seqs{};
    bool text_mode{};
    OccTable bwt() const { return OccTable{bwt_blk.get<OccBlk>(), bwt_sup.get<OccSup>()}; }
    OccTable rev() const { return OccTable{rev_blk.get<OccBlk>(), rev_sup.get<OccSup>()}; }
    uint64_t bytes() const {
        return bwt_blk.cap + bwt_sup.cap + rev_blk.cap + rev_sup.cap + marks.cap + ssa.cap + ref_mark_words.cap + ref_ssa.cap +
               qgram.cap + sa32.cap + isa32.cap + text4.cap + seq_start.cap;
    }
    void release() {
        for (DevBuf* b : {&bwt_blk, &bwt_sup, &rev_blk, &rev_sup, &d_C, &marks, &ssa, &ref_mark_words, &ref_ssa, &qgram, &sa32, &isa32, &text4, &seq_start})
            b->release();
        loaded = false;
        qgram_q = 0;
        text_mode = false;
    }
};

// knobs of the host orchestration (sb200_set_option); none of them changes results
struct Options {
    int bucket_sort{1};        // locate + sort through per-query buckets (0: global radix sort)
    int fused_sort{-1};        // global radix sort with the query id fused into the key (-1: when it saves a pass)
    int textpos{1};            // verified occurrences go to the locate step as text positions (0: as suffix-array rows)
    int ordered_only{0};       // search_n by the ordered walk over every query (0: plain search first, ordered walk for the queries above the limit)
    int debug{0};              // bits 0..7: SearchParams::debug_flags; any bit: progress lines on stderr
    uint64_t chunk{2000000};   // host-buffer calls: queries per middle chunk
    uint64_t edge_div{5};      // ... first and last chunk = 1 / edge_div of the batch
    int pool_blocks_per_sm{0}, pool_threads{0}, run_rounds{0};  // text_pool_kernel geometry (0 = default)
    int items_blocks_per_sm{0}, ordered_blocks_per_sm{0};       // fm_items_kernel / fm_ordered_kernel blocks per SM (0 = default)
    // how the batches of the pipelined calls share the GPU: 0 = one after the other on one stream; 1 = every slot on its own
    // stream, unconstrained (two batches then run in lockstep and their copies are exposed); 2 = own streams, but the in-text
    // verification kernel of a batch (a persistent kernel that fills every SM) starts only when its predecessor batch is
    // complete: packing and the walk over the occurrence tables of batch i+1 fill the tail of batch i, nothing else waits
    int overlap{2};
    int delta_records{1};      // CSR results of the asynchronous calls: records behind the first of a query as differences (varint)
};

enum : int { IN_QUERIES_RANKS = 0, IN_READS_RANKS = 1, IN_READS_PACKED4 = 2, IN_READS_PACKED2 = 3 };
// bytes of one read (or query) of the batch as the caller hands it over
inline size_t input_item_bytes(int in_fmt, uint32_t len) {
    return in_fmt == IN_READS_PACKED4 ? size_t((len + 7) / 8) * 4 : in_fmt == IN_READS_PACKED2 ? size_t((len + 15) / 16) * 4 : len;
}
enum : int { OUT_NONE = 0, OUT_HIT64 = 1, OUT_HIT32 = 2, OUT_CSR = 3 };
constexpr int kSlots = 3;

// one work slot: the buffers, stream and events of one batch in flight
struct Work {
    int id{};
    bool busy{};              // handed out by submit, not yet released
    uint64_t ticket{};
    cudaStream_t own_stream{}, stream{};
    bool pipelined{};         // the batch runs on the slot's own streams (asynchronous calls, chunks of the host-buffer calls)
    cudaEvent_t ev[12]{};     // timing: 0 start, 11 search end, 8..10 fm / text, 1..3 locate / sort
    cudaEvent_t ev_in{}, ev_done{}, ev_ready{}, ev_out{}, ev_fork{};
    DevBuf d_in, d_packed, d_items, d_item_tags, d_seeds, d_spill, d_cursors, d_counters, d_qpos, d_lc, d_tasks, d_bigsegs, d_keys[2], d_qids[2],
        d_offsets, d_tmp, d_scratch, d_rows, d_redo, d_ostack, d_out;
    uint64_t cursor_cap{}, seed_cap{}, hit_cap{};
    uint32_t task_cap{};
    unsigned long long* h_status{};      // pinned + mapped status block (publish_status_kernel)
    unsigned long long* h_status_dev{};
    // pinned result buffers of the asynchronous calls (owned by the slot, recycled)
    void* h_out{};
    size_t h_out_cap{};
    uint32_t* h_ends{};
    size_t h_ends_cap{};
    // the batch
    const uint8_t* d_src{};
    uint64_t n_queries{}, first_query{};
    uint32_t len{};
    int in_fmt{}, out_fmt{};
    bool with_reverse{}, do_locate{}, packed_ready{}, located{};
    uint32_t fused_shift{0};  // hit keys of a global sort carry the query id above this bit (0: separate array)
    int sorted_keys{0};       // d_keys[] buffer that holds the sorted hits
    uint64_t n_cursor_slots{}, n_real_cursors{}, n_hits{};
    DevBuf d_bsize, d_bpos, d_btmp;   // delta-coded records: bytes per query, their inclusive scan, scan scratch
    bool delta{};             // the records in d_out are delta coded
    bool counted{};           // the search kernels counted the rows per query into d_qpos (no hit_count_kernel pass)
    uint64_t n_rec_bytes{};   // their total size
    uint64_t h2d_bytes{}, d2h_bytes{};
    float ms_search{}, ms_locate{}, ms_sort{}, ms_fm{}, ms_text{};
    void release_buffers() {
        for (DevBuf* b : {&d_in, &d_packed, &d_items, &d_item_tags, &d_seeds, &d_spill, &d_cursors, &d_counters, &d_qpos, &d_lc, &d_tasks, &d_bigsegs,
                          &d_keys[0], &d_keys[1], &d_qids[0], &d_qids[1], &d_offsets, &d_tmp, &d_scratch, &d_rows, &d_redo, &d_ostack, &d_out, &d_bsize, &d_bpos, &d_btmp})
            b->release();
    }
};

}  // namespace

struct sb200_ctx {
    int device{};
    cudaStream_t own_stream{}, stream{};
    DeviceIndex idx;
    // scheme
    DevBuf d_steps, d_runs;
    uint32_t n_searches{}, qlen{}, kmax{};
    bool edit{}, have_scheme{};
    std::vector<uint32_t> h_steps;  // the packed steps as set (the state flags are rebuilt when the policy changes)
    sb200_policy policy = SB200_POLICY_DEFAULT;
    uint32_t max_hits{0};  // > 0: search_n (fm_ordered_kernel), at most this many rows per query
    Options opt;
    // work slots; synchronous calls use slot 0 on `stream`
    Work work[kSlots];
    Work* last{};          // slot that holds the result of the last synchronous search (sb200_fetch_hits, sb200_search_cursors)
    uint64_t next_ticket{1};
    Work* prev_slot{};        // slot of the batch submitted last (overlap mode 2 chains on its ev_done)
    uint64_t cap_cursor{}, cap_seed{}, cap_hit{};  // buffer capacities the work slots have learned (shared: a slot starts with them)
    DevBuf d_tmp, d_scratch, d_counters;  // index construction, rank benchmark
    uint64_t nodes_text{};
    float ms_fm{}, ms_text{};
    unsigned long long* h_counters{};  // pinned + mapped, CT_COUNT entries + 8 scratch words
    unsigned long long* h_counters_dev{};  // its device alias
    sb200_counters ct{};
    cudaEvent_t ev[12]{};
    int sms{};
    cudaStream_t s_in{}, s_out{};  // copy streams of the pipelined calls
};

namespace {

void use(sb200_ctx* c) {
    if (!c) throw Error("null context");
    CUDA_TRY(cudaSetDevice(c->device));
}

inline unsigned grid_for(uint64_t n, unsigned block = 256) { return static_cast<unsigned>((n + block - 1) / block); }

template <typename F>
auto with_sigma(uint32_t sigma, F&& f) {
    if (sigma == 5) return f(std::integral_constant<int, 5>{});
    if (sigma == 6) return f(std::integral_constant<int, 6>{});
    throw Error("unknown index with " + std::to_string(sigma) + " letters");
}

void launch_check(sb200_ctx* c) {
    c->ct.kernel_launches += 1;
    CUDA_TRY(cudaGetLastError());
}

// ---- marks: words -> MarkRec -----------------------------------------------------------------------
void build_mark_records(sb200_ctx* c, const uint64_t* d_words, uint64_t n_rows, DevBuf& out, uint64_t& n_rec, uint64_t& total) {
    uint64_t n_words = n_rows / 64 + 1;
    n_rec = n_rows / kRowsPerMark + 1;
    DevBuf popc, rank;
    popc.reserve(n_rec * 4);
    rank.reserve((n_rec + 1) * 4);
    mark_popc_kernel<<<grid_for(n_rec), 256, 0, c->stream>>>(d_words, n_words, n_rec, popc.get<uint32_t>());
    launch_check(c);
    size_t tmp_bytes = 0;
    CUDA_TRY(cub::DeviceScan::ExclusiveSum(nullptr, tmp_bytes, popc.get<uint32_t>(), rank.get<uint32_t>(), n_rec, c->stream));
    c->d_tmp.reserve(tmp_bytes);
    CUDA_TRY(cub::DeviceScan::ExclusiveSum(c->d_tmp.p, tmp_bytes, popc.get<uint32_t>(), rank.get<uint32_t>(), n_rec, c->stream));
    c->ct.kernel_launches += 2;
    out.reserve(n_rec * sizeof(MarkRec));
    mark_fill_kernel<<<grid_for(n_rec), 256, 0, c->stream>>>(d_words, n_words, n_rec, rank.get<uint32_t>(), out.get<MarkRec>());
    launch_check(c);
    uint32_t lastRank = 0, lastPopc = 0;
    CUDA_TRY(cudaMemcpyAsync(&lastRank, rank.get<uint32_t>() + (n_rec - 1), 4, cudaMemcpyDeviceToHost, c->stream));
    CUDA_TRY(cudaMemcpyAsync(&lastPopc, popc.get<uint32_t>() + (n_rec - 1), 4, cudaMemcpyDeviceToHost, c->stream));
    CUDA_TRY(cudaStreamSynchronize(c->stream));
    total = uint64_t(lastRank) + lastPopc;
    popc.release();
    rank.release();
}

uint32_t bits_of(uint64_t v) {
    uint32_t b = 0;
    while (v) { ++b; v >>= 1; }
    return b;
}

void finish_index(sb200_ctx* c) {
    auto& ix = c->idx;
    ix.d_C.reserve(sizeof(uint32_t) * 8);
    CUDA_TRY(cudaMemcpyAsync(ix.d_C.p, ix.C, sizeof(uint32_t) * 8, cudaMemcpyHostToDevice, c->stream));
    // largest sample value -> number of key bits for sorting hits
    uint64_t* d_max = nullptr;
    c->d_scratch.reserve(64);
    d_max = c->d_scratch.get<uint64_t>();
    size_t tmp_bytes = 0;
    CUDA_TRY(cub::DeviceReduce::Max(nullptr, tmp_bytes, ix.ssa.get<uint64_t>(), d_max, ix.n_ssa, c->stream));
    c->d_tmp.reserve(tmp_bytes);
    CUDA_TRY(cub::DeviceReduce::Max(c->d_tmp.p, tmp_bytes, ix.ssa.get<uint64_t>(), d_max, ix.n_ssa, c->stream));
    c->ct.kernel_launches += 1;
    uint64_t mx = 0;
    CUDA_TRY(cudaMemcpyAsync(&mx, d_max, 8, cudaMemcpyDeviceToHost, c->stream));
    CUDA_TRY(cudaStreamSynchronize(c->stream));
    // a located position can exceed its sample by at most sampling_rate-1 inside the same sequence
    ix.key_bits = std::max<uint32_t>(bits_of(mx + ix.sampling_rate), static_cast<uint32_t>(ix.bits_for_position));
    if (ix.key_bits + 4 > 64) throw Error("sampled suffix array values do not fit the 60-bit hit key");
    ix.loaded = true;
}

// verifies rank(n_rows, c) against C for both tables
void verify_histograms(sb200_ctx* c) {
    auto& ix = c->idx;
    DevBuf pos, out;
    pos.reserve(8);
    out.reserve(8 * 8);
    uint64_t n = ix.n_rows;
    CUDA_TRY(cudaMemcpyAsync(pos.p, &n, 8, cudaMemcpyHostToDevice, c->stream));
    for (int which = 0; which < 2; ++which) {
        with_sigma(ix.sigma, [&](auto S) {
            rank_probe_kernel<S()><<<1, 32, 0, c->stream>>>(which ? ix.rev() : ix.bwt(), pos.get<uint64_t>(), 1, out.get<uint64_t>());
            return 0;
        });
        launch_check(c);
        uint64_t r[8];
        CUDA_TRY(cudaMemcpyAsync(r, out.p, 8 * ix.sigma, cudaMemcpyDeviceToHost, c->stream));
        CUDA_TRY(cudaStreamSynchronize(c->stream));
        for (uint32_t s = 0; s < ix.sigma; ++s)
            if (r[s] != ix.C64[s + 1] - ix.C64[s])
                throw Error("index layout not understood: symbol histogram of " + std::string(which ? "bwtRev" : "bwt") +
                            " does not match C");
    }
    pos.release();
    out.release();
}

void upload_occ(sb200_ctx* c, uint32_t sigma, uint64_t n_rows, uint64_t n_blocks, const void* blocks, const uint64_t* super,
                DevBuf& blk, DevBuf& sup) {
    uint64_t stride = 10ull * sigma;
    uint64_t n_super = (n_blocks + 1023) / 1024;
    uint64_t n_sup = n_blocks / 64 + 1;
    DevBuf raw, dsuper, err;
    raw.reserve(n_blocks * stride + 64);
    dsuper.reserve(n_super * sigma * 8);
    err.reserve(4);
    CUDA_TRY(cudaMemcpyAsync(raw.p, blocks, n_blocks * stride, cudaMemcpyHostToDevice, c->stream));
    CUDA_TRY(cudaMemcpyAsync(dsuper.p, super, n_super * sigma * 8, cudaMemcpyHostToDevice, c->stream));
    CUDA_TRY(cudaMemsetAsync(err.p, 0, 4, c->stream));
    blk.reserve((n_blocks + 1) * sizeof(OccBlk));
    sup.reserve((n_sup + 1) * sizeof(OccSup));
    CUDA_TRY(cudaMemsetAsync(blk.p, 0, (n_blocks + 1) * sizeof(OccBlk), c->stream));
    CUDA_TRY(cudaMemsetAsync(sup.p, 0, (n_sup + 1) * sizeof(OccSup), c->stream));
    with_sigma(sigma, [&](auto S) {
        convert_ref_occ_kernel<S()><<<grid_for(n_blocks), 256, 0, c->stream>>>(raw.get<uint8_t>(), dsuper.get<uint64_t>(), n_blocks,
                                                                               n_rows, blk.get<OccBlk>(), sup.get<OccSup>(),
                                                                               err.get<unsigned int>());
        return 0;
    });
    launch_check(c);
    unsigned int e = 0;
    CUDA_TRY(cudaMemcpyAsync(&e, err.p, 4, cudaMemcpyDeviceToHost, c->stream));
    CUDA_TRY(cudaStreamSynchronize(c->stream));
    raw.release();
    dsuper.release();
    err.release();
    if (e & 1) throw Error("index layout not understood: bitplanes of a block overlap or do not cover its rows");
    if (e & 2) throw Error("index layout not understood: block / superblock counters are inconsistent with the bitplanes");
    if (e & 4) throw Error("index layout not understood: counter out of range");
}


// ---- index construction ------------------------------------------------------------------------------

template <typename T>
T read_back(sb200_ctx* c, const T* d) {
    T v;
    CUDA_TRY(cudaMemcpyAsync(&v, d, sizeof(T), cudaMemcpyDeviceToHost, c->stream));
    CUDA_TRY(cudaStreamSynchronize(c->stream));
    return v;
}

struct CountUnsorted {
    Unsorted pred;
    __device__ uint64_t operator()(uint64_t j) const { return pred(j) ? 1 : 0; }
};
struct UnsortedU32 {
    Unsorted pred;
    __device__ bool operator()(uint32_t j) const { return pred(j); }
};

// suffix array of d_text[0, n) (n < 2^32) into `sa` (u32[n])
void build_suffix_array(sb200_ctx* c, const uint8_t* d_text, uint64_t n, DevBuf& sa) {
    DevBuf keys0, keys1, vals0, flags, head, rnk;
    keys0.reserve(n * 8);
    keys1.reserve(n * 8);
    vals0.reserve(n * 4);
    sa.reserve(n * 4);
    make_keys_kernel<<<grid_for(n), 256, 0, c->stream>>>(d_text, n, keys0.get<uint64_t>(), vals0.get<uint32_t>());
    launch_check(c);
    size_t tb = 0;
    CUDA_TRY(cub::DeviceRadixSort::SortPairs(nullptr, tb, keys0.get<uint64_t>(), keys1.get<uint64_t>(), vals0.get<uint32_t>(),
                                             sa.get<uint32_t>(), n, 0, 3 * kKeySyms, c->stream));
    c->d_tmp.reserve(tb);
    CUDA_TRY(cub::DeviceRadixSort::SortPairs(c->d_tmp.p, tb, keys0.get<uint64_t>(), keys1.get<uint64_t>(), vals0.get<uint32_t>(),
                                             sa.get<uint32_t>(), n, 0, 3 * kKeySyms, c->stream));
    c->ct.kernel_launches += 9;
    CUDA_TRY(cudaStreamSynchronize(c->stream));
    keys0.release();
    vals0.release();
    flags.reserve(n + 1);
    group_flags_kernel<<<grid_for(n), 256, 0, c->stream>>>(keys1.get<uint64_t>(), n, flags.get<uint8_t>());
    launch_check(c);
    CUDA_TRY(cudaStreamSynchronize(c->stream));
    keys1.release();

    Unsorted pred{flags.get<uint8_t>(), n};
    DevBuf d_cnt;
    d_cnt.reserve(16);
    auto count_unsorted = [&]() -> uint64_t {
        auto it = cub::TransformInputIterator<uint64_t, CountUnsorted, cub::CountingInputIterator<uint64_t>>(
            cub::CountingInputIterator<uint64_t>(0), CountUnsorted{pred});
        size_t t = 0;
        CUDA_TRY(cub::DeviceReduce::Sum(nullptr, t, it, d_cnt.get<uint64_t>(), n, c->stream));
        c->d_tmp.reserve(t);
        CUDA_TRY(cub::DeviceReduce::Sum(c->d_tmp.p, t, it, d_cnt.get<uint64_t>(), n, c->stream));
        c->ct.kernel_launches += 1;
        return read_back(c, d_cnt.get<uint64_t>());
    };
    uint64_t m = count_unsorted();
    if (m > 0) {
        head.reserve(n * 4);
        rnk.reserve(n * 4);
        {
            auto it = cub::TransformInputIterator<uint32_t, HeadOf, cub::CountingInputIterator<uint64_t>>(
                cub::CountingInputIterator<uint64_t>(0), HeadOf{flags.get<uint8_t>()});
            size_t t = 0;
            CUDA_TRY(cub::DeviceScan::InclusiveScan(nullptr, t, it, head.get<uint32_t>(), MaxU32{}, n, c->stream));
            c->d_tmp.reserve(t);
            CUDA_TRY(cub::DeviceScan::InclusiveScan(c->d_tmp.p, t, it, head.get<uint32_t>(), MaxU32{}, n, c->stream));
            c->ct.kernel_launches += 1;
        }
        scatter_rank_kernel<<<grid_for(n), 256, 0, c->stream>>>(sa.get<uint32_t>(), head.get<uint32_t>(), n, rnk.get<uint32_t>());
        launch_check(c);
        uint64_t h = kKeySyms;
        for (int round = 0; m > 0; ++round) {
            if (round > 64) throw Error("internal error: suffix sorting did not converge");
            DevBuf U, k0, k1, v0, v1, nh0, nh1;
            U.reserve(m * 4);
            k0.reserve(m * 8); k1.reserve(m * 8);
            v0.reserve(m * 4); v1.reserve(m * 4);
            nh0.reserve(m * 4); nh1.reserve(m * 4);
            {
                size_t t = 0;
                cub::CountingInputIterator<uint32_t> cnt(0);
                CUDA_TRY(cub::DeviceSelect::If(nullptr, t, cnt, U.get<uint32_t>(), d_cnt.get<uint64_t>(), static_cast<int64_t>(n),
                                               UnsortedU32{pred}, c->stream));
                c->d_tmp.reserve(t);
                CUDA_TRY(cub::DeviceSelect::If(c->d_tmp.p, t, cnt, U.get<uint32_t>(), d_cnt.get<uint64_t>(), static_cast<int64_t>(n),
                                               UnsortedU32{pred}, c->stream));
                c->ct.kernel_launches += 1;
            }
            refine_keys_kernel<<<grid_for(m), 256, 0, c->stream>>>(U.get<uint32_t>(), m, sa.get<uint32_t>(), head.get<uint32_t>(),
                                                                 rnk.get<uint32_t>(), n, h, k0.get<uint64_t>(), v0.get<uint32_t>());
            launch_check(c);
            {
                size_t t = 0;
                CUDA_TRY(cub::DeviceRadixSort::SortPairs(nullptr, t, k0.get<uint64_t>(), k1.get<uint64_t>(), v0.get<uint32_t>(),
                                                         v1.get<uint32_t>(), m, 0, 64, c->stream));
                c->d_tmp.reserve(t);
                CUDA_TRY(cub::DeviceRadixSort::SortPairs(c->d_tmp.p, t, k0.get<uint64_t>(), k1.get<uint64_t>(), v0.get<uint32_t>(),
                                                         v1.get<uint32_t>(), m, 0, 64, c->stream));
                c->ct.kernel_launches += 9;
            }
            refine_flags_kernel<<<grid_for(m), 256, 0, c->stream>>>(U.get<uint32_t>(), m, k1.get<uint64_t>(), v1.get<uint32_t>(),
                                                                  sa.get<uint32_t>(), flags.get<uint8_t>(), nh0.get<uint32_t>());
            launch_check(c);
            {
                size_t t = 0;
                CUDA_TRY(cub::DeviceScan::InclusiveScan(nullptr, t, nh0.get<uint32_t>(), nh1.get<uint32_t>(), MaxU32{}, m, c->stream));
                c->d_tmp.reserve(t);
                CUDA_TRY(cub::DeviceScan::InclusiveScan(c->d_tmp.p, t, nh0.get<uint32_t>(), nh1.get<uint32_t>(), MaxU32{}, m, c->stream));
                c->ct.kernel_launches += 1;
            }
            refine_scatter_kernel<<<grid_for(m), 256, 0, c->stream>>>(U.get<uint32_t>(), m, v1.get<uint32_t>(), nh1.get<uint32_t>(),
                                                                    head.get<uint32_t>(), rnk.get<uint32_t>());
            launch_check(c);
            CUDA_TRY(cudaStreamSynchronize(c->stream));
            for (DevBuf* b : {&U, &k0, &k1, &v0, &v1, &nh0, &nh1}) b->release();
            h *= 2;
            m = count_unsorted();
        }
    }
    for (DevBuf* b : {&flags, &head, &rnk, &d_cnt}) b->release();
}

// BWT of the text given its suffix array, packed into the device occurrence table
void build_occ(sb200_ctx* c, const uint8_t* d_text, const uint32_t* d_sa, uint64_t n, DevBuf& blk, DevBuf& sup) {
    uint64_t n_blocks = n / 64 + 1;
    uint64_t n_sup = n_blocks / 64 + 1;
    DevBuf bwt, tot, scanned;
    bwt.reserve(n_sup * 4096);
    CUDA_TRY(cudaMemsetAsync(bwt.p, 0, n_sup * 4096, c->stream));
    bwt_kernel<<<grid_for(n), 256, 0, c->stream>>>(d_text, d_sa, n, bwt.get<uint8_t>());
    launch_check(c);
    blk.reserve((n_blocks + 1) * sizeof(OccBlk));
    sup.reserve((n_sup + 1) * sizeof(OccSup));
    CUDA_TRY(cudaMemsetAsync(blk.p, 0, (n_blocks + 1) * sizeof(OccBlk), c->stream));
    CUDA_TRY(cudaMemsetAsync(sup.p, 0, (n_sup + 1) * sizeof(OccSup), c->stream));
    tot.reserve(n_sup * sizeof(Cnt8));
    scanned.reserve(n_sup * sizeof(Cnt8));
    pack_occ_kernel<<<static_cast<unsigned>(n_sup), 64, 0, c->stream>>>(bwt.get<uint8_t>(), n, n_blocks, blk.get<OccBlk>(), tot.get<Cnt8>());
    launch_check(c);
    size_t t = 0;
    CUDA_TRY(cub::DeviceScan::ExclusiveScan(nullptr, t, tot.get<Cnt8>(), scanned.get<Cnt8>(), AddCnt8{}, Cnt8{}, n_sup, c->stream));
    c->d_tmp.reserve(t);
    CUDA_TRY(cub::DeviceScan::ExclusiveScan(c->d_tmp.p, t, tot.get<Cnt8>(), scanned.get<Cnt8>(), AddCnt8{}, Cnt8{}, n_sup, c->stream));
    c->ct.kernel_launches += 1;
    sup_write_kernel<<<grid_for(n_sup), 256, 0, c->stream>>>(scanned.get<Cnt8>(), n_sup, sup.get<OccSup>());
    launch_check(c);
    CUDA_TRY(cudaStreamSynchronize(c->stream));
    for (DevBuf* b : {&bwt, &tot, &scanned}) b->release();
}

void build_index_device(sb200_ctx* c, const uint8_t* d_src, const uint64_t* seq_lens, uint64_t n_seqs, uint32_t sigma, uint32_t rate) {
    if (sigma != 5 && sigma != 6) throw Error("unknown index with " + std::to_string(sigma) + " letters");
    if (n_seqs == 0) throw Error("reference file was empty - abort");
    if (n_seqs >= (1ull << 31)) throw Error("too many sequences");
    if (rate == 0) throw Error("sampling rate must be positive");
    std::vector<uint64_t> start(n_seqs + 1, 0);
    uint64_t maxLen = 0, srcLen = 0;
    for (uint64_t i = 0; i < n_seqs; ++i) {
        start[i + 1] = start[i] + seq_lens[i] + 1;
        maxLen = std::max(maxLen, seq_lens[i] + 1);
        srcLen += seq_lens[i];
    }
    uint64_t n = start[n_seqs];
    if (n >= (1ull << 32) - 8192) throw Error("text too long: the GPU index needs fewer than 2^32 rows");
    auto& ix = c->idx;
    ix.release();
    ix.sigma = sigma;
    ix.n_rows = n;
    ix.n_blocks = n / 64 + 1;
    ix.n_sup = ix.n_blocks / 64 + 1;
    ix.sampling_rate = rate;
    ix.device_rate = rate;
    ix.bits_for_position = std::max<uint32_t>(1, bits_of(maxLen));
    ix.full_sa = false;

    DevBuf d_start, err, text, sa;
    d_start.reserve((n_seqs + 1) * 8);
    CUDA_TRY(cudaMemcpyAsync(d_start.p, start.data(), (n_seqs + 1) * 8, cudaMemcpyHostToDevice, c->stream));
    err.reserve(4);
    CUDA_TRY(cudaMemsetAsync(err.p, 0, 4, c->stream));
    if (srcLen) {
        check_text_kernel<<<grid_for(srcLen), 256, 0, c->stream>>>(d_src, srcLen, sigma, err.get<unsigned int>());
        launch_check(c);
    }
    if (read_back(c, err.get<unsigned int>())) throw Error("reference has invalid character (rank outside 1..sigma-1)");
    SeqMap map{d_start.get<uint64_t>(), static_cast<uint32_t>(n_seqs)};
    text.reserve(n + 64);
    for (int pass = 0; pass < 2; ++pass) {
        make_text_kernel<<<grid_for(n), 256, 0, c->stream>>>(d_src, map, n, pass == 1, text.get<uint8_t>());
        launch_check(c);
        build_suffix_array(c, text.get<uint8_t>(), n, sa);
        if (pass == 0) {
            build_occ(c, text.get<uint8_t>(), sa.get<uint32_t>(), n, ix.bwt_blk, ix.bwt_sup);
            // sampled suffix array
            uint64_t n_words = n / 64 + 1;
            ix.ref_mark_words.reserve(n_words * 8);
            CUDA_TRY(cudaMemsetAsync(ix.ref_mark_words.p, 0, n_words * 8, c->stream));
            sample_marks_kernel<<<grid_for(n), 256, 0, c->stream>>>(sa.get<uint32_t>(), n, map, rate,
                                                                  reinterpret_cast<uint32_t*>(ix.ref_mark_words.p));
            launch_check(c);
            uint64_t total = 0;
            build_mark_records(c, ix.ref_mark_words.get<uint64_t>(), n, ix.marks, ix.n_mark, total);
            ix.ssa.reserve(std::max<uint64_t>(1, total) * 8);
            sample_values_kernel<<<grid_for(n), 256, 0, c->stream>>>(sa.get<uint32_t>(), n, map, rate,
                                                                   static_cast<uint32_t>(ix.bits_for_position), ix.marks.get<MarkRec>(),
                                                                   ix.ssa.get<uint64_t>());
            launch_check(c);
            ix.n_ssa = ix.n_ref_ssa = total;
        } else {
            build_occ(c, text.get<uint8_t>(), sa.get<uint32_t>(), n, ix.rev_blk, ix.rev_sup);
        }
    }
    CUDA_TRY(cudaStreamSynchronize(c->stream));
    for (DevBuf* b : {&text, &sa, &err}) b->release();
    // C from the symbol histogram = rank(n, s)
    {
        DevBuf pos, out;
        pos.reserve(8);
        out.reserve(64);
        CUDA_TRY(cudaMemcpyAsync(pos.p, &n, 8, cudaMemcpyHostToDevice, c->stream));
        with_sigma(sigma, [&](auto S) {
            rank_probe_kernel<S()><<<1, 32, 0, c->stream>>>(ix.bwt(), pos.get<uint64_t>(), 1, out.get<uint64_t>());
            return 0;
        });
        launch_check(c);
        uint64_t r[8] = {0};
        CUDA_TRY(cudaMemcpyAsync(r, out.p, 8 * sigma, cudaMemcpyDeviceToHost, c->stream));
        CUDA_TRY(cudaStreamSynchronize(c->stream));
        ix.C64[0] = 0;
        for (uint32_t s = 0; s < 8; ++s) {
            if (s < sigma) ix.C64[s + 1 > 7 ? 7 : s + 1] = ix.C64[s] + r[s];
        }
        for (uint32_t s = sigma + 1; s < 8; ++s) ix.C64[s] = n;
        for (int i = 0; i < 8; ++i) ix.C[i] = static_cast<uint32_t>(ix.C64[i]);
        if (ix.C64[sigma] != n) throw Error("internal error: symbol histogram does not add up");
        pos.release();
        out.release();
    }
    d_start.release();
    verify_histograms(c);
    finish_index(c);
}

// ---- search pipeline ---------------------------------------------------------------------------------
//
// A BATCH (contiguous range of queries) runs as one fixed sequence of launches on the stream of its work slot, with no
// host round trip in between: pack -> fm_roots -> fm_items -> text_pool -> hit_count -> scan -> locate_scatter ->
// locate_tasks -> segment_sort (+ big) -> output formatting -> publish.  Every size the next kernel needs (cursor slots,
// seeds, hits) stays in device memory; buffers are sized from capacities that grow when a batch did not fit (the kernels
// never write out of bounds, they raise flags; finish_batch() then starts the batch over with larger buffers).  The
// host reads the published status block once, when it waits for the batch.
//
// Slots make batches overlap: while slot A computes, the reads of the next batch are copied into slot B and the hits of
// the previous one leave slot C (copy streams s_in / s_out).  The public asynchronous pair sb200_submit_reads /
// sb200_wait_reads and the synchronous host-buffer calls (which cut their batch into chunks) share this code.

// worst-case stack depth of the pair/chain traversal (search.cuh): D(k) = 2, D(e) = F(e) - 1 + D(e + 1) with
// F(e) <= 9 (k - e) + 2 frames pushed per iteration
template <typename F>
void with_stack(uint32_t k, F&& f) {
    if (k == 0) f(std::integral_constant<int, 4>{});
    else if (k == 1) f(std::integral_constant<int, 16>{});
    else if (k == 2) f(std::integral_constant<int, 36>{});
    else if (k == 3) f(std::integral_constant<int, 64>{});
    else if (k == 4) f(std::integral_constant<int, 96>{});
    else throw Error("search schemes with more than 4 errors are not supported by the GPU kernel yet");
}

__global__ void mirror_words_kernel(const unsigned long long* src, unsigned long long* dst, int n) {
    if (threadIdx.x < n) dst[threadIdx.x] = src[threadIdx.x];
}
// n (<= CT_COUNT + 1) device words -> c->h_counters[at ...], then waits for the stream
void read_back_words(sb200_ctx* c, cudaStream_t s, const void* d_src, int n, int at = 0) {
    mirror_words_kernel<<<1, 32, 0, s>>>(static_cast<const unsigned long long*>(d_src), c->h_counters_dev + at, n);
    launch_check(c);
    CUDA_TRY(cudaStreamSynchronize(s));
}

// status block a batch publishes into mapped host memory (one kernel at its end)
enum : int { ST_COUNTERS = 0, ST_LC = CT_COUNT, ST_HITS = CT_COUNT + LC_COUNT, ST_BYTES = ST_HITS + 1, ST_COUNT = CT_COUNT + LC_COUNT + 2 };
__global__ void publish_status_kernel(const unsigned long long* counters, const unsigned int* lc, const uint32_t* total_hits,
                                      const uint32_t* total_bytes, unsigned long long* dst) {
    const int t = threadIdx.x;
    if (t < CT_COUNT) dst[ST_COUNTERS + t] = counters[t];
    else if (t < CT_COUNT + LC_COUNT) dst[t] = lc ? lc[t - CT_COUNT] : 0u;
    else if (t == ST_HITS) dst[t] = total_hits ? *total_hits : 0u;
    else if (t == ST_BYTES) dst[t] = total_bytes ? *total_bytes : 0u;
}

SearchParams search_params(sb200_ctx* c, Work& w) {
    auto& ix = c->idx;
    SearchParams P{};
    P.bwt = ix.bwt();
    P.bwtRev = ix.rev();
    for (int i = 0; i < 8; ++i) P.C[i] = ix.C[i];
    P.n_rows = static_cast<uint32_t>(ix.n_rows);
    P.packed = w.d_packed.get<uint32_t>();
    P.seeds = w.d_seeds.get<uint4>();
    P.seed_cap = static_cast<uint32_t>(std::min<uint64_t>(w.seed_cap, 0xfffffffeull));
    P.n_queries = static_cast<uint32_t>(w.n_queries);
    P.len = w.len;
    P.n_searches = c->n_searches;
    P.steps = c->d_steps.get<uint32_t>();
    P.runs = c->d_runs.get<uint8_t>();
    P.out = w.d_cursors.get<uint4>();
    P.out_cap = static_cast<uint32_t>(std::min<uint64_t>(w.cursor_cap, 0xfffffffeull));
    P.counters = w.d_counters.get<unsigned long long>();
    P.qgram = ix.qgram_q ? ix.qgram.get<uint4>() : nullptr;
    P.qgram_q = ix.qgram_q;
    if (P.qgram_q >= w.len) P.qgram = nullptr, P.qgram_q = 0;  // (a table that covers the whole query: plain walk)
    P.sa32 = ix.text_mode ? ix.sa32.get<uint32_t>() : nullptr;
    P.isa32 = ix.text_mode ? ix.isa32.get<uint32_t>() : nullptr;
    P.text4 = ix.text_mode ? ix.text4.get<uint32_t>() : nullptr;
    // cursors that go straight to the locate step carry the text position of a verified occurrence instead of its row
    P.textpos_out = (w.do_locate && ix.text_mode && c->opt.textpos) ? 1u : 0u;
    P.debug_flags = static_cast<uint32_t>(c->opt.debug) & 0xffu;
    P.pol = c->policy;
    return P;
}

// search_n: the ordered walk (fm_ordered_kernel), one thread per query with its stack in global memory
unsigned ordered_grid(sb200_ctx* c, uint32_t len, uint64_t n_queries) {
    const unsigned per_sm = c->opt.ordered_blocks_per_sm > 0 ? static_cast<unsigned>(c->opt.ordered_blocks_per_sm) : (len > 300 ? 1u : 4u);
    unsigned g = static_cast<unsigned>(c->sms) * per_sm;  // (the stacks: 16 B x frames per thread)
    return std::max(1u, std::min(g, grid_for(n_queries)));
}
void launch_ordered(sb200_ctx* c, Work& w, const SearchParams& P) {
    const size_t osmem = size_t(P.n_searches) * P.len * 4;
    if (osmem > 48 * 1024) throw Error("search scheme table does not fit shared memory (query too long)");
    const unsigned ogrid = ordered_grid(c, P.len, P.n_queries);
    SearchParams Q = P;
    Q.ostack_frames = ordered_stack_frames(P.len, c->idx.sigma);
    w.d_ostack.reserve(size_t(ogrid) * 256 * Q.ostack_frames * sizeof(uint4));
    Q.ostack = w.d_ostack.get<uint4>();
    with_sigma(c->idx.sigma, [&](auto S) {
        if (c->edit) fm_ordered_kernel<S(), true><<<ogrid, 256, osmem, w.stream>>>(Q);
        else fm_ordered_kernel<S(), false><<<ogrid, 256, osmem, w.stream>>>(Q);
        return 0;
    });
    launch_check(c);
}

// launch geometry of text_pool_kernel for this scheme: as many resident warps per SM as shared memory (227 KB, 1 KB
// reserved per block) and registers allow; measured on the headline workload: 36 warps 8.4 ms, 30 warps 9.2 ms, 24 warps
// 10.2 ms
struct PoolGeometry { unsigned threads, per_sm; size_t smem; };
PoolGeometry pool_geometry(sb200_ctx* c, uint32_t n_searches, uint32_t len) {
    const size_t tables = (size_t(n_searches) * len * 4 + run_table_bytes(n_searches * len) + 7) & ~size_t{7};
    PoolGeometry g{0, 1, 0};
    for (unsigned wps = kPoolThreads / 32; wps >= 4; --wps) {
        const size_t bytes = tables + wps * size_t(pool_bytes(len));
        const unsigned blocks = static_cast<unsigned>(std::min<size_t>((227 * 1024) / (bytes + 1024), 42 / wps));
        if (blocks * wps > g.per_sm * (g.threads / 32)) g = PoolGeometry{wps * 32, blocks, bytes};
    }
    if (c->opt.pool_threads > 0) {
        g.threads = std::min<unsigned>(kPoolThreads, std::max(32, c->opt.pool_threads)) & ~31u;
        g.smem = tables + (g.threads / 32) * size_t(pool_bytes(len));
        g.per_sm = static_cast<unsigned>(std::max<size_t>(1, std::min<size_t>(42 / (g.threads / 32), (227 * 1024) / (g.smem + 1024))));
    }
    if (c->opt.pool_blocks_per_sm > 0) g.per_sm = static_cast<unsigned>(c->opt.pool_blocks_per_sm);
    if (g.threads == 0 || g.smem > 226 * 1024) throw Error("search scheme table and frame pools do not fit shared memory (query too long)");
    return g;
}

// the search kernels of a batch (queries already packed in w.d_packed); nothing waits for the device
void enqueue_search_kernels(sb200_ctx* c, Work& w) {
    auto& ix = c->idx;
    const uint64_t n_queries = w.n_queries;
    w.d_cursors.reserve((w.cursor_cap + 1) * sizeof(uint4));
    if (ix.text_mode) w.d_seeds.reserve((w.seed_cap + 1) * sizeof(uint4));
    unsigned long long* ctr = w.d_counters.get<unsigned long long>();
    CUDA_TRY(cudaMemsetAsync(ctr, 0, CT_BAD_QUERY * sizeof(unsigned long long), w.stream));  // keeps CT_BAD_QUERY
    CUDA_TRY(cudaMemsetAsync(ctr + CT_NODES_TEXT, 0, (CT_COUNT - CT_NODES_TEXT) * sizeof(unsigned long long), w.stream));
    SearchParams P = search_params(c, w);
    // the rows per query are counted while the cursors are written when the bucketed locate follows right behind
    // (no row limit, no foreign cursors): saves its counting pass over the cursor list
    w.counted = w.do_locate && c->max_hits == 0 && c->opt.bucket_sort;
    if (w.counted) {
        w.d_qpos.reserve((n_queries + 1) * 4);
        CUDA_TRY(cudaMemsetAsync(w.d_qpos.p, 0, (n_queries + 1) * 4, w.stream));
        P.qcount = w.d_qpos.get<uint32_t>();
    }
    CUDA_TRY(cudaEventRecord(w.ev[8], w.stream));
    // search_n by the ordered walk alone (option ordered_only); the default is the plain search first, then the queries
    // above the limit again in recursion order (refine_max_hits)
    if (c->max_hits && c->opt.ordered_only) {
        P.max_hits = c->max_hits;
        launch_ordered(c, w, P);
        CUDA_TRY(cudaEventRecord(w.ev[9], w.stream));
        CUDA_TRY(cudaEventRecord(w.ev[10], w.stream));
        return;
    }
    // root frames of all (query, search) in one pass, then the warp-synchronous walk over the live ones
    const uint64_t total = n_queries * uint64_t(c->n_searches);
    w.d_items.reserve(total * sizeof(uint4));
    w.d_item_tags.reserve(total * sizeof(uint2));
    P.items = w.d_items.get<uint4>();
    P.item_tags = w.d_item_tags.get<uint2>();
    fm_roots_kernel<<<grid_for(n_queries), 256, 0, w.stream>>>(P);
    launch_check(c);
    const size_t ismem = size_t(P.n_searches) * P.len * 4;
    if (ismem > 48 * 1024) throw Error("search scheme table does not fit shared memory (query too long)");
    const unsigned iper = c->opt.items_blocks_per_sm > 0 ? static_cast<unsigned>(c->opt.items_blocks_per_sm) : SB200_FM_ITEMS_BLOCKS;
    const unsigned igrid = std::max(1u, std::min(static_cast<unsigned>(c->sms) * iper, grid_for(n_queries)));
    // (a policy without PAIR frames pushes the deletion and the substitution children separately: one level more)
    const uint32_t stack_k = std::min<uint32_t>(4, c->kmax + (sb200_pol_pairs(&c->policy) ? 0u : 1u));
    with_sigma(ix.sigma, [&](auto S) {
        with_stack(stack_k, [&](auto STACK) {
            if (c->edit) fm_items_kernel<S(), true, STACK()><<<igrid, 256, ismem, w.stream>>>(P);
            else fm_items_kernel<S(), false, STACK()><<<igrid, 256, ismem, w.stream>>>(P);
        });
        return 0;
    });
    launch_check(c);
    CUDA_TRY(cudaEventRecord(w.ev[9], w.stream));
    if (P.sa32) {  // in-text verification, one frame pool per warp (reads the seed count from device memory)
        const PoolGeometry g = pool_geometry(c, P.n_searches, P.len);
        const unsigned tgrid = static_cast<unsigned>(c->sms) * g.per_sm;
        w.d_spill.reserve(size_t(tgrid) * (g.threads / 32) * 2 * kSpillCap * sizeof(uint4));
        const uint32_t run_rounds = c->opt.run_rounds > 0 ? static_cast<uint32_t>(c->opt.run_rounds) : kRunRounds;
        // overlap mode 2: not before the batch submitted before this one is complete
        if (w.pipelined && c->opt.overlap == 2 && c->prev_slot && c->prev_slot != &w) CUDA_TRY(cudaStreamWaitEvent(w.stream, c->prev_slot->ev_done, 0));
        with_stack(stack_k, [&](auto STACK) {
            auto go = [&](auto kern) {
                CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(g.smem)));
                kern<<<tgrid, g.threads, g.smem, w.stream>>>(P, 2 * (c->kmax + 1), run_rounds, w.d_spill.get<uint4>());
            };
            if (c->edit) go(text_pool_kernel<true, STACK()>);
            else go(text_pool_kernel<false, STACK()>);
        });
        launch_check(c);
    }
    CUDA_TRY(cudaEventRecord(w.ev[10], w.stream));
}

LocateIndex locate_index(sb200_ctx* c) {
    auto& ix = c->idx;
    LocateIndex L{};
    L.bwt = ix.bwt();
    for (int i = 0; i < 8; ++i) L.C[i] = ix.C[i];
    L.marks = ix.full_sa ? nullptr : ix.marks.get<MarkRec>();
    L.ssa = ix.ssa.get<uint64_t>();
    L.seq_start = ix.text_mode ? ix.seq_start.get<uint64_t>() : nullptr;
    L.n_seqs = static_cast<uint32_t>(ix.n_seqs);
    L.bits = static_cast<uint32_t>(ix.bits_for_position);
    return L;
}

// Locate + sort through per-query buckets (locate.cuh) for the cursors the search kernels of this batch leave in
// w.d_cursors.  The cursor count is read on the device (CT_OUT_SLOTS); hits land in w.d_keys[0] (+ w.d_qids[0] when the
// output format needs the query of every hit), per-query ends in w.d_qpos.
void enqueue_locate_bucketed(sb200_ctx* c, Work& w) {
    auto& ix = c->idx;
    const uint64_t n_queries = w.n_queries;
    unsigned long long* ctr = w.d_counters.get<unsigned long long>();
    CUDA_TRY(cudaEventRecord(w.ev[1], w.stream));
    w.d_qpos.reserve((n_queries + 1) * 4);
    w.d_lc.reserve(LC_COUNT * 4);
    CUDA_TRY(cudaMemsetAsync(w.d_lc.p, 0, LC_COUNT * 4, w.stream));
    const uint32_t cur_cap = static_cast<uint32_t>(std::min<uint64_t>(w.cursor_cap, 0xfffffffeull));
    const unsigned wide = static_cast<unsigned>(c->sms) * 8;
    if (!w.counted) {  // (else the search kernels have counted the rows per query as they wrote the cursors)
        CUDA_TRY(cudaMemsetAsync(w.d_qpos.p, 0, (n_queries + 1) * 4, w.stream));
        hit_count_kernel<<<wide, 256, 0, w.stream>>>(w.d_cursors.get<uint4>(), cur_cap, ctr + CT_OUT_SLOTS, w.d_qpos.get<uint32_t>(), ctr + CT_TOTAL_ROWS);
        launch_check(c);
    }
    w.counted = false;  // (a later bucketed pass over this batch — search_n — counts for itself)
    size_t tmp_bytes = 0;
    CUDA_TRY(cub::DeviceScan::ExclusiveSum(nullptr, tmp_bytes, w.d_qpos.get<uint32_t>(), w.d_qpos.get<uint32_t>(), n_queries + 1, w.stream));
    w.d_tmp.reserve(tmp_bytes);
    CUDA_TRY(cub::DeviceScan::ExclusiveSum(w.d_tmp.p, tmp_bytes, w.d_qpos.get<uint32_t>(), w.d_qpos.get<uint32_t>(), n_queries + 1, w.stream));
    c->ct.kernel_launches += 2;
    const bool want_qids = w.out_fmt != OUT_CSR;
    w.d_keys[0].reserve(std::max<uint64_t>(1, w.hit_cap) * 8);
    if (want_qids) w.d_qids[0].reserve(std::max<uint64_t>(1, w.hit_cap) * 4);
    // the scan leaves the total behind the last query; the scatter turns qpos[q] into the END of segment q, so the total
    // is kept aside for the kernels that follow
    w.d_scratch.reserve(64);
    CUDA_TRY(cudaMemcpyAsync(w.d_scratch.p, w.d_qpos.get<uint32_t>() + n_queries, 4, cudaMemcpyDeviceToDevice, w.stream));
    BucketParams B{};
    B.index = locate_index(c);
    B.cursors = w.d_cursors.get<uint4>();
    B.n_cursors = cur_cap;
    B.n_cursors_dev = ctr + CT_OUT_SLOTS;
    B.n_queries = static_cast<uint32_t>(n_queries);
    B.qpos = w.d_qpos.get<uint32_t>();
    B.keys = w.d_keys[0].get<uint64_t>();
    B.qids = want_qids ? w.d_qids[0].get<uint32_t>() : nullptr;
    B.key_cap = static_cast<uint32_t>(std::min<uint64_t>(w.hit_cap, 0xfffffff0ull));
    B.task_cap = static_cast<uint32_t>(std::min<uint64_t>(w.hit_cap / kInlineRows + 1024, 0xfffffff0ull));
    B.big_cap = static_cast<uint32_t>(w.hit_cap / kWarpSeg + 16);
    w.d_tasks.reserve(uint64_t(B.task_cap) * sizeof(uint4));
    w.d_bigsegs.reserve(uint64_t(B.big_cap) * 4);
    B.tasks = w.d_tasks.get<uint4>();
    B.big_segs = w.d_bigsegs.get<uint32_t>();
    B.lc = w.d_lc.get<unsigned int>();
    B.counters = ctr;
    w.task_cap = B.task_cap;
    with_sigma(ix.sigma, [&](auto S) {
        locate_scatter_kernel<S()><<<wide, 256, 0, w.stream>>>(B);
        launch_check(c);
        locate_tasks_kernel<S()><<<static_cast<unsigned>(c->sms) * 4, 256, 0, w.stream>>>(B);
        return 0;
    });
    launch_check(c);
    CUDA_TRY(cudaEventRecord(w.ev[2], w.stream));
    segment_sort_kernel<<<grid_for((n_queries + 31) / 32 * 32), 256, 0, w.stream>>>(B);  // one lane per query
    launch_check(c);
    segment_sort_big_kernel<<<static_cast<unsigned>(c->sms) * 4, 256, 0, w.stream>>>(B);
    launch_check(c);
    CUDA_TRY(cudaEventRecord(w.ev[3], w.stream));
    w.fused_shift = 0;
    w.sorted_keys = 0;
}

size_t out_record_bytes(sb200_ctx* c, int fmt) {
    if (fmt == OUT_HIT64) return sizeof(sb200_hit);
    if (fmt == OUT_HIT32) return sizeof(sb200_hit32);
    if (fmt == OUT_CSR) return (c->idx.key_bits + 4 + 7) / 8;
    return 0;
}

// sorted hits of the batch (w.d_keys[w.sorted_keys], w.d_qids[0]) -> w.d_out in the wanted format; n = capacity (count read
// on the device) or, with n_dev == nullptr, the exact number
void enqueue_output(sb200_ctx* c, Work& w, uint64_t n, const uint32_t* n_dev) {
    if (w.out_fmt == OUT_NONE) return;
    auto& ix = c->idx;
    const size_t rec = out_record_bytes(c, w.out_fmt);
    w.d_out.reserve((std::max<uint64_t>(1, n) + 4) * rec + 64);
    const uint64_t* keys = w.d_keys[w.sorted_keys].get<uint64_t>();
    const unsigned grid = n_dev ? static_cast<unsigned>(c->sms) * 8 : std::max(1u, grid_for(n));
    if (w.out_fmt == OUT_HIT32)
        compact_hits_kernel<<<grid, 256, 0, w.stream>>>(keys, w.d_qids[0].get<uint32_t>(), n, n_dev, static_cast<uint32_t>(ix.bits_for_position),
                                                       static_cast<uint32_t>(w.first_query), w.fused_shift, w.d_out.get<uint4>());
    else if (w.out_fmt == OUT_HIT64)
        expand_hits_kernel<<<grid, 256, 0, w.stream>>>(keys, w.d_qids[0].get<uint32_t>(), n, n_dev, static_cast<uint32_t>(ix.bits_for_position),
                                                      w.first_query, w.fused_shift, w.d_out.get<uint64_t>());
    else if (!w.delta)
        pack_records_kernel<<<grid, 256, 0, w.stream>>>(keys, n, n_dev, w.fused_shift, static_cast<uint32_t>(rec), w.d_out.get<uint8_t>());
    else {
        // delta-coded records (locate.cuh): bytes per query, their scan, the bytes; the per-query ends are in w.d_qpos
        const uint32_t nq = static_cast<uint32_t>(w.n_queries);
        w.d_out.reserve((std::max<uint64_t>(1, n) + 4) * (rec + 1) + 64);  // (a difference takes at most rec + 1 bytes)
        w.d_bsize.reserve((size_t(nq) + 1) * 4);
        w.d_bpos.reserve((size_t(nq) + 1) * 4);
        delta_size_kernel<<<std::max(1u, grid_for(nq)), 256, 0, w.stream>>>(keys, static_cast<uint32_t>(std::min<uint64_t>(n, 0xfffffff0ull)), w.d_qpos.get<uint32_t>(), nq, w.fused_shift, static_cast<uint32_t>(rec),
                                                                             w.d_bsize.get<uint32_t>());
        launch_check(c);
        size_t tmp = 0;
        CUDA_TRY(cub::DeviceScan::InclusiveSum(nullptr, tmp, w.d_bsize.get<uint32_t>(), w.d_bpos.get<uint32_t>(), nq, w.stream));
        w.d_btmp.reserve(tmp);
        CUDA_TRY(cub::DeviceScan::InclusiveSum(w.d_btmp.p, tmp, w.d_bsize.get<uint32_t>(), w.d_bpos.get<uint32_t>(), nq, w.stream));
        c->ct.kernel_launches += 2;
        delta_write_kernel<<<std::max(1u, grid_for(nq)), 256, 0, w.stream>>>(keys, static_cast<uint32_t>(std::min<uint64_t>(n, 0xfffffff0ull)), w.d_qpos.get<uint32_t>(), w.d_bpos.get<uint32_t>(), nq, w.fused_shift,
                                                                              static_cast<uint32_t>(rec), w.d_out.get<uint8_t>());
    }
    launch_check(c);
}

void enqueue_publish(sb200_ctx* c, Work& w, bool located) {
    publish_status_kernel<<<1, 64, 0, w.stream>>>(w.d_counters.get<unsigned long long>(), located ? w.d_lc.get<unsigned int>() : nullptr,
                                                  located ? w.d_scratch.get<uint32_t>() : nullptr,
                                                  located && w.delta && w.n_queries ? w.d_bpos.get<uint32_t>() + (w.n_queries - 1) : nullptr, w.h_status_dev);
    launch_check(c);
    CUDA_TRY(cudaEventRecord(w.ev_done, w.stream));
}

// everything of a batch that can be queued without looking at device results: pack, search, locate, sort, output
void enqueue_compute(sb200_ctx* c, Work& w) {
    auto& ix = c->idx;
    const uint64_t n_queries = w.n_queries;
    const uint32_t len = w.len, W = packed_words(len);
    w.cursor_cap = std::max(w.cursor_cap, c->cap_cursor);  // (capacities other slots needed for this workload)
    w.seed_cap = std::max(w.seed_cap, c->cap_seed);
    w.hit_cap = std::max(w.hit_cap, c->cap_hit);
    if (w.cursor_cap < n_queries * 16) w.cursor_cap = std::max<uint64_t>(1 << 20, n_queries * 16);
    if (w.seed_cap < n_queries * 4) w.seed_cap = std::max<uint64_t>(1 << 20, n_queries * 4);
    if (w.hit_cap < n_queries * 12) w.hit_cap = std::max<uint64_t>(1 << 20, n_queries * 12);
    w.d_counters.reserve(CT_COUNT * sizeof(unsigned long long));
    w.d_packed.reserve(n_queries * W * 4);
    unsigned long long* ctr = w.d_counters.get<unsigned long long>();
    CUDA_TRY(cudaEventRecord(w.ev[0], w.stream));
    if (!w.packed_ready) {
        CUDA_TRY(cudaMemsetAsync(ctr, 0, CT_COUNT * sizeof(unsigned long long), w.stream));
        const uint8_t* in = w.d_src;
        const unsigned grid = grid_for(n_queries * W);
        if (w.in_fmt == IN_READS_PACKED2)
            pack_packed2_kernel<<<grid, 256, 0, w.stream>>>(reinterpret_cast<const uint32_t*>(in), n_queries, len, w.with_reverse ? 1u : 0u,
                                                           w.d_packed.get<uint32_t>());
        else if (w.in_fmt == IN_READS_PACKED4)
            pack_packed4_kernel<<<grid, 256, 0, w.stream>>>(reinterpret_cast<const uint32_t*>(in), n_queries, len, ix.sigma, w.with_reverse ? 1u : 0u,
                                                           w.d_packed.get<uint32_t>(), ctr);
        else if (w.in_fmt == IN_READS_RANKS && w.with_reverse)
            pack_reads_kernel<<<grid, 256, 0, w.stream>>>(in, n_queries, len, ix.sigma, w.d_packed.get<uint32_t>(), ctr);
        else
            pack_queries_kernel<<<grid, 256, 0, w.stream>>>(in, n_queries, len, ix.sigma, w.d_packed.get<uint32_t>(), ctr);
        launch_check(c);
        w.packed_ready = true;
    }
    enqueue_search_kernels(c, w);
    CUDA_TRY(cudaEventRecord(w.ev[11], w.stream));
    // search_n needs the host between search and locate (which queries exceed the limit); foreign cursor lists and the
    // option bucket_sort = 0 take the global radix sort: both are finished synchronously in finish_batch
    w.located = w.do_locate && c->max_hits == 0 && c->opt.bucket_sort;
    if (w.located) {
        enqueue_locate_bucketed(c, w);
        enqueue_output(c, w, w.hit_cap, w.d_scratch.get<uint32_t>());
    }
    enqueue_publish(c, w, w.located);
    if (w.pipelined) c->prev_slot = &w;
}

// search_n after the plain search: a query with at most max_hits rows is complete; the others (they end early by
// definition) are walked again in the reference's recursion order and their cursors of the first pass are dropped.
// n_slots: output slots of the first pass.  false = the cursor buffer is too small (cursor_cap was raised, the caller
// starts over).
bool refine_max_hits(sb200_ctx* c, Work& w, uint64_t n_slots) {
    SearchParams P = search_params(c, w);
    const uint32_t nq = P.n_queries;
    w.d_rows.reserve((size_t(nq) + 4) * sizeof(unsigned long long));
    w.d_redo.reserve(size_t(nq) * sizeof(uint32_t));
    unsigned long long* rows = w.d_rows.get<unsigned long long>();
    unsigned long long* tally = rows + nq;  // queries to redo, bound on their cursors, cursors dropped
    CUDA_TRY(cudaMemsetAsync(rows, 0, (size_t(nq) + 4) * sizeof(unsigned long long), w.stream));
    if (n_slots) cursor_rows_kernel<<<grid_for(n_slots), 256, 0, w.stream>>>(P.out, n_slots, rows);
    redo_list_kernel<<<grid_for(nq), 256, 0, w.stream>>>(rows, nq, c->max_hits, w.d_redo.get<uint32_t>(), tally);
    launch_check(c);
    read_back_words(c, w.stream, tally, 2, CT_COUNT);
    const uint64_t n_redo = c->h_counters[CT_COUNT], bound = c->h_counters[CT_COUNT + 1];
    if (n_redo == 0) return true;
    const uint64_t need = n_slots + bound + uint64_t(ordered_grid(c, P.len, n_redo)) * 256 * kEmitChunk;
    if (need >= 0xfffffffeull) throw Error("more than 2^32 cursors in one call; split the batch");
    if (need > w.cursor_cap) {
        w.cursor_cap = need + need / 8;
        return false;
    }
    drop_cursors_kernel<<<grid_for(n_slots), 256, 0, w.stream>>>(P.out, n_slots, rows, c->max_hits, tally);
    CUDA_TRY(cudaMemsetAsync(P.counters + CT_NEXT_QUERY, 0, sizeof(unsigned long long), w.stream));
    SearchParams Q = P;
    Q.max_hits = c->max_hits;
    Q.redo = w.d_redo.get<uint32_t>();
    Q.n_queries = static_cast<uint32_t>(n_redo);
    launch_ordered(c, w, Q);
    read_back_words(c, w.stream, tally + 2, 1, CT_COUNT);
    const uint64_t dropped = c->h_counters[CT_COUNT];
    read_back_words(c, w.stream, w.d_counters.p, CT_COUNT);
    if (c->h_counters[CT_OVERFLOW]) throw Error("internal error: search stack overflow");
    if (c->h_counters[CT_OUT_SLOTS] > w.cursor_cap) throw Error("internal error: cursor buffer of the ordered walk too small");
    for (int i = 0; i < CT_COUNT; ++i) w.h_status[ST_COUNTERS + i] = c->h_counters[i];
    w.h_status[ST_COUNTERS + CT_CURSORS] -= dropped;
    return true;
}

// kernel 3 over the n_cursors cursors in w.d_cursors (room for one extra slot) with a global sort by (qid, seq/pos, e):
// cursors with arbitrary query ids (sb200_locate), queries with thousands of hits, option bucket_sort = 0.  Synchronous.
// Sorted hits end in w.d_keys[w.sorted_keys] (fused keys) or w.d_keys[0] / w.d_qids[0].
void locate_radix(sb200_ctx* c, Work& w, uint64_t n_cursors, uint64_t n_queries_hint) {
    auto& ix = c->idx;
    CUDA_TRY(cudaEventRecord(w.ev[1], w.stream));
    w.d_counters.reserve(CT_COUNT * sizeof(unsigned long long));
    unsigned long long* ctr = w.d_counters.get<unsigned long long>();
    CUDA_TRY(cudaMemsetAsync(ctr + CT_LF_STEPS, 0, sizeof(unsigned long long), w.stream));
    uint64_t total_rows = 0;
    if (n_cursors > 0) {
        w.d_offsets.reserve((n_cursors + 1) * 8);
        auto lens = cub::TransformInputIterator<uint64_t, CursorLen, const uint4*>(w.d_cursors.get<uint4>(), CursorLen{});
        size_t tmp_bytes = 0;
        CUDA_TRY(cub::DeviceScan::ExclusiveSum(nullptr, tmp_bytes, lens, w.d_offsets.get<uint64_t>(), n_cursors + 1, w.stream));
        w.d_tmp.reserve(tmp_bytes);
        // n_cursors + 1 items are scanned: zero the slot behind the last cursor
        CUDA_TRY(cudaMemsetAsync(w.d_cursors.get<uint4>() + n_cursors, 0, sizeof(uint4), w.stream));
        CUDA_TRY(cub::DeviceScan::ExclusiveSum(w.d_tmp.p, tmp_bytes, lens, w.d_offsets.get<uint64_t>(), n_cursors + 1, w.stream));
        c->ct.kernel_launches += 1;
        read_back_words(c, w.stream, w.d_offsets.get<uint64_t>() + n_cursors, 1, CT_COUNT);
        total_rows = c->h_counters[CT_COUNT];
    }
    if (total_rows >= (1ull << 32)) throw Error("more than 2^32 hits in one call; split the batch");
    // one 64-bit key (query id above value and errors) when it fits: a single keys-only radix sort
    const int key_bits = static_cast<int>(ix.key_bits) + 4;
    const int qid_bits = std::max(1, static_cast<int>(bits_of(n_queries_hint ? n_queries_hint - 1 : 0xffffffffull)));
    // (measured at equal pass counts, 16 M hits: two pair sorts 1.19 ms, one 64-bit keys-only sort 1.38 ms — so only
    // when it saves a radix pass)
    const bool fewer_passes = (key_bits + qid_bits + 7) / 8 < (key_bits + 7) / 8 + (qid_bits + 7) / 8;
    const bool fused = key_bits + qid_bits <= 64 && (c->opt.fused_sort >= 0 ? c->opt.fused_sort != 0 : fewer_passes);
    w.fused_shift = fused ? static_cast<uint32_t>(key_bits) : 0u;
    w.sorted_keys = 0;
    for (int i = 0; i < 2; ++i) {
        w.d_keys[i].reserve(std::max<uint64_t>(1, total_rows) * 8);
        if (!fused) w.d_qids[i].reserve(std::max<uint64_t>(1, total_rows) * 4);
    }
    if (total_rows > 0) {
        LocateParams L{};
        L.index = locate_index(c);
        L.cursors = w.d_cursors.get<uint4>();
        L.offsets = w.d_offsets.get<uint64_t>();
        L.n_cursors = static_cast<uint32_t>(n_cursors);
        L.n_rows_total = total_rows;
        L.out_key = w.d_keys[0].get<uint64_t>();
        L.out_qid = fused ? nullptr : w.d_qids[0].get<uint32_t>();
        L.fused_shift = w.fused_shift;
        L.counters = ctr;
        with_sigma(ix.sigma, [&](auto S) {
            locate_kernel<S()><<<grid_for(total_rows), 256, 0, w.stream>>>(L);
            return 0;
        });
        launch_check(c);
    }
    CUDA_TRY(cudaEventRecord(w.ev[2], w.stream));
    if (total_rows > 1 && fused) {
        size_t t1 = 0;
        CUDA_TRY(cub::DeviceRadixSort::SortKeys(nullptr, t1, w.d_keys[0].get<uint64_t>(), w.d_keys[1].get<uint64_t>(), total_rows, 0,
                                                key_bits + qid_bits, w.stream));
        w.d_tmp.reserve(t1);
        CUDA_TRY(cub::DeviceRadixSort::SortKeys(w.d_tmp.p, t1, w.d_keys[0].get<uint64_t>(), w.d_keys[1].get<uint64_t>(), total_rows, 0,
                                                key_bits + qid_bits, w.stream));
        w.sorted_keys = 1;
        c->ct.kernel_launches += (key_bits + qid_bits + 7) / 8 + 2;
    } else if (total_rows > 1) {
        size_t t1 = 0, t2 = 0;
        CUDA_TRY(cub::DeviceRadixSort::SortPairs(nullptr, t1, w.d_keys[0].get<uint64_t>(), w.d_keys[1].get<uint64_t>(),
                                                 w.d_qids[0].get<uint32_t>(), w.d_qids[1].get<uint32_t>(), total_rows, 0, key_bits, w.stream));
        CUDA_TRY(cub::DeviceRadixSort::SortPairs(nullptr, t2, w.d_qids[1].get<uint32_t>(), w.d_qids[0].get<uint32_t>(),
                                                 w.d_keys[1].get<uint64_t>(), w.d_keys[0].get<uint64_t>(), total_rows, 0, qid_bits, w.stream));
        w.d_tmp.reserve(std::max(t1, t2));
        CUDA_TRY(cub::DeviceRadixSort::SortPairs(w.d_tmp.p, t1, w.d_keys[0].get<uint64_t>(), w.d_keys[1].get<uint64_t>(),
                                                 w.d_qids[0].get<uint32_t>(), w.d_qids[1].get<uint32_t>(), total_rows, 0, key_bits, w.stream));
        CUDA_TRY(cub::DeviceRadixSort::SortPairs(w.d_tmp.p, t2, w.d_qids[1].get<uint32_t>(), w.d_qids[0].get<uint32_t>(),
                                                 w.d_keys[1].get<uint64_t>(), w.d_keys[0].get<uint64_t>(), total_rows, 0, qid_bits, w.stream));
        c->ct.kernel_launches += 2 * ((key_bits + 7) / 8 + 1) + 2 * ((qid_bits + 7) / 8 + 1);
    }
    CUDA_TRY(cudaEventRecord(w.ev[3], w.stream));
    if (w.out_fmt == OUT_CSR) {  // per-query ends of the sorted list
        if (n_queries_hint == 0) throw Error("internal error: CSR output needs the number of queries");
        w.d_qpos.reserve((n_queries_hint + 1) * 4);
        csr_ends_kernel<<<std::max(1u, std::min(grid_for(total_rows + 1), static_cast<unsigned>(c->sms) * 8)), 256, 0, w.stream>>>(
            w.d_keys[w.sorted_keys].get<uint64_t>(), w.d_qids[0].get<uint32_t>(), total_rows, w.fused_shift, static_cast<uint32_t>(n_queries_hint),
            w.d_qpos.get<uint32_t>());
        launch_check(c);
    }
    enqueue_output(c, w, total_rows, nullptr);
    read_back_words(c, w.stream, ctr, CT_COUNT);
    w.h_status[ST_COUNTERS + CT_LF_STEPS] = c->h_counters[CT_LF_STEPS];
    w.n_hits = total_rows;
    if (w.delta && w.out_fmt == OUT_CSR && w.n_queries) {
        uint32_t bytes = 0;
        CUDA_TRY(cudaMemcpyAsync(&bytes, w.d_bpos.get<uint32_t>() + (w.n_queries - 1), 4, cudaMemcpyDeviceToHost, w.stream));
        CUDA_TRY(cudaStreamSynchronize(w.stream));
        w.n_rec_bytes = bytes;
    }
}

// waits for the batch, grows buffers and starts over when something did not fit, finishes the parts that need the host
// (search_n, global sort).  Afterwards: w.n_cursor_slots / w.n_real_cursors / w.n_hits are valid, hits sit in w.d_out.
void finish_batch(sb200_ctx* c, Work& w) {
    const uint32_t len = w.len;
    for (int attempt = 0;; ++attempt) {
        if (attempt > 8) throw Error("internal error: work buffers do not converge");
        CUDA_TRY(cudaEventSynchronize(w.ev_done));
        const unsigned long long* st = w.h_status;
        if (st[ST_COUNTERS + CT_BAD_QUERY]) {
            const uint64_t off = st[ST_COUNTERS + CT_BAD_QUERY] - 1 + w.first_query * len;
            throw Error("query has invalid character at offset " + std::to_string(off % len) + " of query " + std::to_string(off / len));
        }
        if (c->opt.debug)
            fprintf(stderr, "[sb200 debug] slot %d: max stack depth %llu overflow %llu seeds %llu cursors %llu hits %llu\n", w.id,
                    st[ST_COUNTERS + CT_MAX_SP], st[ST_COUNTERS + CT_OVERFLOW], st[ST_COUNTERS + CT_SEEDS], st[ST_COUNTERS + CT_OUT_SLOTS], st[ST_HITS]);
        if (st[ST_COUNTERS + CT_OVERFLOW]) throw Error("internal error: search stack overflow");
        const uint64_t n_slots = st[ST_COUNTERS + CT_OUT_SLOTS], n_seed_slots = st[ST_COUNTERS + CT_SEED_SLOTS];
        if (n_slots >= 0xfffffffeull || n_seed_slots >= 0xfffffffeull) throw Error("more than 2^32 cursors in one call; split the batch");
        bool fits = true;
        if (n_seed_slots > w.seed_cap) {  // the text kernel saw a truncated seed list
            // (the cursors counted so far come from the part of the list it saw: scale them, and the hits they will give,
            // so that one restart is enough instead of one per buffer)
            const uint64_t est = n_slots / std::max<uint64_t>(1, w.seed_cap) * n_seed_slots + n_slots % std::max<uint64_t>(1, w.seed_cap) * n_seed_slots / std::max<uint64_t>(1, w.seed_cap);
            w.seed_cap = n_seed_slots + n_seed_slots / 4;
            w.cursor_cap = std::max(w.cursor_cap, est + est / 4);
            w.hit_cap = std::max(w.hit_cap, est + est / 4);
            fits = false;
        }
        if (n_slots > w.cursor_cap) {
            w.cursor_cap = n_slots + n_slots / 4;
            w.hit_cap = std::max(w.hit_cap, w.cursor_cap);  // (every cursor gives at least one hit)
            fits = false;
        }
        bool radix = false;
        if (fits && w.located) {
            const uint64_t total = st[ST_HITS];
            // (the scan is u32: n_slots + rows beyond the first of every cursor bounds the true number of hits)
            if (n_slots + st[ST_COUNTERS + CT_TOTAL_ROWS] >= (1ull << 32)) throw Error("more than 2^32 hits in one call; split the batch");
            if (total > w.hit_cap || st[ST_LC + LC_KEY_OVERFLOW]) {
                w.hit_cap = total + total / 4 + 1024;
                fits = false;
            } else if (st[ST_LC + LC_TASKS] > w.task_cap) {
                w.hit_cap = std::max<uint64_t>(w.hit_cap * 2, uint64_t(st[ST_LC + LC_TASKS]) * kInlineRows);
                fits = false;
            } else {
                radix = st[ST_LC + LC_HUGE] != 0;  // a query with more hits than a block sorts
                w.n_hits = total;
                w.n_rec_bytes = st[ST_BYTES];
                // (the byte offsets of delta-coded records are u32 like the hit offsets: a record takes at most rec + 1 bytes)
                if (w.delta && total * (out_record_bytes(c, OUT_CSR) + 1) >= (1ull << 32))
                    throw Error("more than 2^32 bytes of hit records in one batch; split the batch");
            }
        }
        if (!fits) {
            c->ct.batch_restarts += 1;
            enqueue_compute(c, w);
            continue;
        }
        // what this slot learned about the workload holds for the other slots too
        c->cap_cursor = std::max(c->cap_cursor, w.cursor_cap);
        c->cap_seed = std::max(c->cap_seed, w.seed_cap);
        c->cap_hit = std::max(c->cap_hit, w.hit_cap);
        w.n_cursor_slots = n_slots;
        if (c->max_hits && !c->opt.ordered_only) {
            if (!refine_max_hits(c, w, n_slots)) {
                c->ct.batch_restarts += 1;
                enqueue_compute(c, w);
                continue;
            }
            w.n_cursor_slots = w.h_status[ST_COUNTERS + CT_OUT_SLOTS];
        }
        if (w.do_locate && (!w.located || radix)) {
            // (cursors of our own search carry query ids 0 .. n_queries - 1: per-query buckets unless switched off)
            if (!w.located && c->opt.bucket_sort) {  // search_n: the bucketed path, now that the cursor list is final
                unsigned long long* ctr = w.d_counters.get<unsigned long long>();
                CUDA_TRY(cudaMemcpyAsync(ctr + CT_OUT_SLOTS, &w.n_cursor_slots, 8, cudaMemcpyHostToDevice, w.stream));
                CUDA_TRY(cudaMemsetAsync(ctr + CT_LF_STEPS, 0, sizeof(unsigned long long), w.stream));
                CUDA_TRY(cudaMemsetAsync(ctr + CT_TOTAL_ROWS, 0, sizeof(unsigned long long), w.stream));
                const unsigned long long keep_cursors = w.h_status[ST_COUNTERS + CT_CURSORS], keep_nodes = w.h_status[ST_COUNTERS + CT_NODES],
                                         keep_text = w.h_status[ST_COUNTERS + CT_NODES_TEXT];
                bool again = true;
                while (again) {
                    enqueue_locate_bucketed(c, w);
                    enqueue_output(c, w, w.hit_cap, w.d_scratch.get<uint32_t>());
                    enqueue_publish(c, w, true);
                    CUDA_TRY(cudaEventSynchronize(w.ev_done));
                    const uint64_t total = w.h_status[ST_HITS];
                    again = false;
                    if (total > w.hit_cap || w.h_status[ST_LC + LC_KEY_OVERFLOW]) {
                        w.hit_cap = total + total / 4 + 1024;
                        again = true;
                    } else if (w.h_status[ST_LC + LC_TASKS] > w.task_cap) {
                        w.hit_cap = std::max<uint64_t>(w.hit_cap * 2, uint64_t(w.h_status[ST_LC + LC_TASKS]) * kInlineRows);
                        again = true;
                    } else {
                        radix = w.h_status[ST_LC + LC_HUGE] != 0;
                        w.n_hits = total;
                    }
                }
                w.h_status[ST_COUNTERS + CT_CURSORS] = keep_cursors;
                w.h_status[ST_COUNTERS + CT_NODES] = keep_nodes;
                w.h_status[ST_COUNTERS + CT_NODES_TEXT] = keep_text;
                w.located = true;
            } else {
                radix = true;
            }
            if (radix) locate_radix(c, w, w.n_cursor_slots, w.n_queries);
        }
        break;
    }
    const unsigned long long* st = w.h_status;
    w.n_real_cursors = st[ST_COUNTERS + CT_CURSORS];
    if (!w.do_locate) w.n_hits = 0;
    // accounting (everything of this batch is behind ev_done or a synchronous read-back)
    float ms = 0;
    CUDA_TRY(cudaEventElapsedTime(&ms, w.ev[0], w.ev[11]));
    w.ms_search = ms;
    CUDA_TRY(cudaEventElapsedTime(&w.ms_fm, w.ev[8], w.ev[9]));
    CUDA_TRY(cudaEventElapsedTime(&w.ms_text, w.ev[9], w.ev[10]));
    w.ms_locate = w.ms_sort = 0;
    if (w.do_locate) {
        CUDA_TRY(cudaEventElapsedTime(&w.ms_locate, w.ev[1], w.ev[2]));
        CUDA_TRY(cudaEventElapsedTime(&w.ms_sort, w.ev[2], w.ev[3]));
    }
    c->ct.nodes += st[ST_COUNTERS + CT_NODES];
    c->nodes_text += st[ST_COUNTERS + CT_NODES_TEXT];
    c->ct.rank_ops += 2 * (st[ST_COUNTERS + CT_NODES] - st[ST_COUNTERS + CT_NODES_TEXT]);
    c->ct.cursors += st[ST_COUNTERS + CT_CURSORS];
    if (w.do_locate) {
        c->ct.lf_steps += st[ST_COUNTERS + CT_LF_STEPS];
        c->ct.hits += w.n_hits;
    }
    c->ct.ms_search = w.ms_search;
    c->ct.ms_locate = w.ms_locate;
    c->ct.ms_sort = w.ms_sort;
    c->ms_fm = w.ms_fm;
    c->ms_text = w.ms_text;
}

// describes a batch in slot w; src = host pointer (copied on the input stream) or device pointer
void setup_batch(sb200_ctx* c, Work& w, const void* src, bool src_on_device, int in_fmt, bool with_reverse, uint64_t n_queries, uint32_t len,
                 bool do_locate, int out_fmt, uint64_t first_query, cudaStream_t stream) {
    auto& ix = c->idx;
    if (!ix.loaded) throw Error("no index loaded");
    if (!c->have_scheme) throw Error("no search scheme set");
    if (len != c->qlen) throw Error("query length " + std::to_string(len) + " does not match the expanded search scheme (" +
                                    std::to_string(c->qlen) + ")");
    if (!src || n_queries == 0) throw Error("query file was empty - abort");
    if (n_queries * uint64_t(c->n_searches) >= (1ull << 32)) throw Error("too many (query, search) pairs for one call; split the batch");
    if (c->n_searches > 255) throw Error("search schemes with more than 255 searches are not supported");
    w.pipelined = stream == w.own_stream;
    // overlap mode 0: the batches of all slots run one after the other on the stream of slot 0
    w.stream = (w.pipelined && c->opt.overlap == 0) ? c->work[0].own_stream : stream;
    w.n_queries = n_queries;
    w.len = len;
    w.in_fmt = in_fmt;
    w.with_reverse = with_reverse;
    w.do_locate = do_locate;
    w.out_fmt = do_locate ? out_fmt : OUT_NONE;
    w.delta = w.out_fmt == OUT_CSR && c->opt.delta_records != 0;
    w.n_rec_bytes = 0;
    w.first_query = first_query;
    w.packed_ready = false;
    w.located = false;
    w.n_hits = w.n_cursor_slots = w.n_real_cursors = 0;
    const uint64_t n_items = (with_reverse && in_fmt != IN_QUERIES_RANKS) ? n_queries / 2 : n_queries;
    const uint64_t bytes = n_items * input_item_bytes(in_fmt, len);
    if (src_on_device) {
        w.d_src = static_cast<const uint8_t*>(src);
    } else {
        w.d_in.reserve(bytes + 16);
        // the slot's previous batch is finished (slots are handed out only after finish_batch): the buffer is free
        CUDA_TRY(cudaMemcpyAsync(w.d_in.p, src, bytes, cudaMemcpyHostToDevice, c->s_in));
        CUDA_TRY(cudaEventRecord(w.ev_in, c->s_in));
        CUDA_TRY(cudaStreamWaitEvent(w.stream, w.ev_in, 0));
        w.d_src = w.d_in.get<uint8_t>();
        w.h2d_bytes = bytes;
    }
}

// copies the formatted hits of a finished batch (w.d_out) and, for CSR, the per-query ends to host memory on the output
// stream; returns without waiting (w.ev_out marks the end)
void enqueue_copy_out(sb200_ctx* c, Work& w, void* dst_records, uint32_t* dst_ends) {
    const size_t rec = out_record_bytes(c, w.out_fmt);
    // (finish_batch has waited for everything of this batch: no event needed; in overlap mode 0 the slot's stream also
    // carries the successors' kernels, which the copy must not wait for)
    const size_t rec_bytes_total = w.delta ? w.n_rec_bytes : w.n_hits * rec;
    if (w.n_hits) CUDA_TRY(cudaMemcpyAsync(dst_records, w.d_out.p, rec_bytes_total, cudaMemcpyDeviceToHost, c->s_out));
    // (delta-coded records: the ends are byte offsets)
    if (dst_ends) CUDA_TRY(cudaMemcpyAsync(dst_ends, w.delta ? w.d_bpos.p : w.d_qpos.p, w.n_queries * 4, cudaMemcpyDeviceToHost, c->s_out));
    CUDA_TRY(cudaEventRecord(w.ev_out, c->s_out));
    w.d2h_bytes = rec_bytes_total + (dst_ends ? w.n_queries * 4 : 0);
}

// the synchronous entry points run on slot 0 and the caller's stream
void run_pipeline(sb200_ctx* c, const uint8_t* d_queries, uint64_t n_queries, uint32_t len, bool do_locate, int in_fmt = IN_QUERIES_RANKS,
                  bool with_reverse = false, int out_fmt = OUT_NONE) {
    Work& w = c->work[0];
    if (w.busy) throw Error("a submitted batch is still in flight in slot 0: wait for it first");
    setup_batch(c, w, d_queries, true, in_fmt, with_reverse, n_queries, len, do_locate, out_fmt, 0, c->stream);
    enqueue_compute(c, w);
    finish_batch(c, w);
    c->last = &w;
}

void fetch_hits(sb200_ctx* c, sb200_hit** hits, uint64_t* n_hits) {
    Work* wp = c->last;
    if (!wp) throw Error("no search result to fetch");
    Work& w = *wp;
    const uint64_t n = w.n_hits;
    sb200_hit* out = static_cast<sb200_hit*>(g_pinned.alloc(std::max<uint64_t>(1, n) * sizeof(sb200_hit)));
    if (n) {
        // expand to the reference tuple on the device, then one pinned copy
        const int keep = w.out_fmt;
        w.out_fmt = OUT_HIT64;
        enqueue_output(c, w, n, nullptr);
        w.out_fmt = keep;
        CUDA_TRY(cudaEventRecord(c->ev[4], w.stream));
        CUDA_TRY(cudaMemcpyAsync(out, w.d_out.p, n * sizeof(sb200_hit), cudaMemcpyDeviceToHost, w.stream));
        CUDA_TRY(cudaEventRecord(c->ev[5], w.stream));
        CUDA_TRY(cudaStreamSynchronize(w.stream));
        CUDA_TRY(cudaEventElapsedTime(&c->ct.ms_d2h, c->ev[4], c->ev[5]));
    }
    *hits = out;
    *n_hits = n;
}

// search + locate from host buffers, pipelined over the work slots: the batch is cut into chunks of reads; the
// host->device copy of chunk i+1 and the device->host copy of the hits of chunk i-1 run on their own streams while chunk
// i computes, and the kernels of chunk i+1 are queued behind those of chunk i before the host looks at chunk i.
// Chunks are contiguous query ranges, so concatenating their sorted hit lists keeps the global order.
//   in_fmt / with_reverse : what the host buffer holds (sb200.h); the reverse complements are made on the device
//                           (queries[2i] = read i, queries[2i+1] = its reverse complement, search.cpp:121-123)
//   out_fmt               : OUT_HIT64 (the reference's 32-byte tuple) or OUT_HIT32 (16 bytes)
void search_host_pipelined(sb200_ctx* c, const uint8_t* src, uint64_t n_items, uint32_t len, int in_fmt, bool with_reverse, int out_fmt,
                           void** hits, uint64_t* n_hits) {
    auto& ix = c->idx;
    if (!ix.loaded) throw Error("no index loaded");
    if (!src || n_items == 0) throw Error("query file was empty - abort");
    for (auto& w : c->work)
        if (w.busy) throw Error("submitted batches are still in flight: wait for them first");
    const uint64_t per_item = (with_reverse && in_fmt != IN_QUERIES_RANKS) ? 2 : 1;  // queries per host item
    const uint64_t n_queries = n_items * per_item;
    const size_t hit_bytes = out_record_bytes(c, out_fmt);
    const size_t item_bytes = input_item_bytes(in_fmt, len);
    // Chunk boundaries (in queries).  Every chunk costs a kernel drain (the longest single seed), so few chunks: a short
    // first one (its copy-in cannot be hidden), a short last one (its copy-out cannot be hidden), and the rest in pieces
    // of at most `chunk` queries whose copies hide behind the neighbours' kernels.
    uint64_t chunk = std::max<uint64_t>(2, c->opt.chunk);
    const uint64_t edge_div = std::max<uint64_t>(2, c->opt.edge_div);
    chunk += chunk & 1;  // both strands of a read stay together
    std::vector<uint64_t> bounds{0};
    {
        uint64_t edge = std::min(chunk / 2, n_queries / edge_div) & ~uint64_t{1};
        if (edge >= 16384 && n_queries >= 4 * edge) {
            const uint64_t middle = n_queries - 2 * edge;
            const uint64_t pieces = (middle + chunk - 1) / chunk;
            bounds.push_back(edge);
            for (uint64_t i = 1; i < pieces; ++i) bounds.push_back(edge + ((middle * i / pieces) & ~uint64_t{1}));
            bounds.push_back(n_queries - edge);
        } else {
            while (bounds.back() + chunk < n_queries) bounds.push_back(bounds.back() + chunk);
        }
        bounds.push_back(n_queries);
    }
    const uint64_t n_chunks = bounds.size() - 1;
    // wait until earlier work on the caller's stream is done before the slot streams touch anything
    CUDA_TRY(cudaStreamSynchronize(c->stream));
    uint8_t* out = nullptr;
    uint64_t out_cap = 0, total = 0;
    float ms_search = 0, ms_locate = 0, ms_sort = 0;
    auto t0 = std::chrono::steady_clock::now();
    auto submit = [&](uint64_t k) {
        Work& w = c->work[k % kSlots];
        const uint64_t q0 = bounds[k], n = bounds[k + 1] - q0;
        setup_batch(c, w, src + (q0 / per_item) * item_bytes, false, in_fmt, with_reverse, n, len, true, out_fmt, q0, w.own_stream);
        enqueue_compute(c, w);
        w.busy = true;
    };
    try {
        const uint64_t depth = std::min<uint64_t>(kSlots - 1, n_chunks);  // one slot stays free for the copy-out of the oldest chunk
        for (uint64_t k = 0; k < depth; ++k) submit(k);
        for (uint64_t k = 0; k < n_chunks; ++k) {
            Work& w = c->work[k % kSlots];
            finish_batch(c, w);
            if (c->opt.debug)
                fprintf(stderr, "[sb200 debug] chunk %llu: %llu queries, device search %.3f (fm %.3f text %.3f) locate %.3f sort %.3f ms\n",
                        (unsigned long long)k, (unsigned long long)w.n_queries, w.ms_search, w.ms_fm, w.ms_text, w.ms_locate, w.ms_sort);
            ms_search += w.ms_search;
            ms_locate += w.ms_locate;
            ms_sort += w.ms_sort;
            const uint64_t nh = w.n_hits;
            // output buffer: sized from the first chunk, grown (rarely) when the estimate was too small
            if (total + nh > out_cap) {
                // first estimate: the hit density of the first chunk over the whole batch
                const uint64_t n = w.n_queries;
                uint64_t want = k == 0 ? nh * ((n_queries + n - 1) / n) + nh / 4 + 1024 : (total + nh) * 2;
                uint8_t* bigger = static_cast<uint8_t*>(g_pinned.alloc(std::max<uint64_t>(1, want) * hit_bytes));
                if (out) {
                    CUDA_TRY(cudaStreamSynchronize(c->s_out));  // copies into the old buffer must have landed
                    std::memcpy(bigger, out, total * hit_bytes);
                    g_pinned.free(out);
                }
                out = bigger;
                out_cap = want;
            }
            enqueue_copy_out(c, w, out + total * hit_bytes, nullptr);
            total += nh;
            // the slot of chunk k + depth is the one whose copy-out (chunk k + depth - kSlots) was queued one round ago
            if (k + depth < n_chunks) {
                Work& nw = c->work[(k + depth) % kSlots];
                if (nw.busy) {
                    CUDA_TRY(cudaEventSynchronize(nw.ev_out));
                    nw.busy = false;
                }
                submit(k + depth);
            }
        }
        CUDA_TRY(cudaStreamSynchronize(c->s_out));
        CUDA_TRY(cudaStreamSynchronize(c->s_in));
        for (auto& w : c->work) w.busy = false;
    } catch (...) {
        cudaDeviceSynchronize();
        for (auto& w : c->work) w.busy = false;
        if (out) g_pinned.free(out);
        throw;
    }
    if (!out) out = static_cast<uint8_t*>(g_pinned.alloc(hit_bytes));
    if (c->opt.debug)
        fprintf(stderr, "[sb200 debug] host-buffer search total host wall %.3f ms\n",
                std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count());
    c->ct.ms_search = ms_search;
    c->ct.ms_locate = ms_locate;
    c->ct.ms_sort = ms_sort;
    c->ct.ms_h2d = c->ct.ms_d2h = 0;  // overlapped with the kernels
    c->last = nullptr;
    *hits = out;
    *n_hits = total;
}

// scheme tables of the context -> device: packed steps, run lengths and the state flags (which depend on the policy)
void upload_scheme_tables(sb200_ctx* c) {
    const std::vector<uint32_t>& steps = c->h_steps;
    c->d_steps.reserve(steps.size() * 4);
    CUDA_TRY(cudaMemcpyAsync(c->d_steps.p, steps.data(), steps.size() * 4, cudaMemcpyHostToDevice, c->stream));
    std::vector<uint8_t> runs(run_table_bytes(static_cast<uint32_t>(steps.size())) + 4, 0);  // run lengths, state flags, path windows
    build_runs(c->n_searches, c->qlen, steps.data(), runs.data());
    build_state_flags(c->n_searches, c->qlen, steps.data(), runs.data(), c->policy);
    build_path_windows(c->n_searches, c->qlen, steps.data(), runs.data());
    c->d_runs.reserve(runs.size());
    CUDA_TRY(cudaMemcpyAsync(c->d_runs.p, runs.data(), runs.size(), cudaMemcpyHostToDevice, c->stream));
    CUDA_TRY(cudaStreamSynchronize(c->stream));
}

}  // namespace

// ======================================================================================================
extern "C" {

int sb200_abi_version(void) { return SB200_ABI_VERSION; }
const char* sb200_last_error(void) { return g_err.c_str(); }

int sb200_device_count(int* count) {
    return guard([&] {
        int n = 0;
        cudaError_t e = cudaGetDeviceCount(&n);
        if (e != cudaSuccess) {
            cudaGetLastError();
            throw Error(std::string("no usable CUDA device: ") + cudaGetErrorString(e));
        }
        *count = n;
    });
}

int sb200_create(int device, sb200_ctx** out) {
    return guard([&] {
        int n = 0;
        cudaError_t e = cudaGetDeviceCount(&n);
        if (e != cudaSuccess || n == 0) {
            cudaGetLastError();
            throw Error(std::string("sahara_b200 needs a CUDA device and has no CPU fallback: ") +
                        (e != cudaSuccess ? cudaGetErrorString(e) : "no device found"));
        }
        if (device < 0 || device >= n) throw Error("device ordinal out of range");
        CUDA_TRY(cudaSetDevice(device));
        auto c = new sb200_ctx();
        c->device = device;
        CUDA_TRY(cudaStreamCreateWithFlags(&c->own_stream, cudaStreamNonBlocking));
        c->stream = c->own_stream;
        for (auto& ev : c->ev) CUDA_TRY(cudaEventCreate(&ev));
        CUDA_TRY(cudaStreamCreateWithFlags(&c->s_in, cudaStreamNonBlocking));
        CUDA_TRY(cudaStreamCreateWithFlags(&c->s_out, cudaStreamNonBlocking));
        for (int i = 0; i < kSlots; ++i) {
            Work& w = c->work[i];
            w.id = i;
            CUDA_TRY(cudaStreamCreateWithFlags(&w.own_stream, cudaStreamNonBlocking));
            for (auto& ev : w.ev) CUDA_TRY(cudaEventCreate(&ev));
            for (cudaEvent_t* ev : {&w.ev_in, &w.ev_done, &w.ev_ready, &w.ev_out, &w.ev_fork})
                CUDA_TRY(cudaEventCreateWithFlags(ev, cudaEventDisableTiming));
            CUDA_TRY(cudaHostAlloc(reinterpret_cast<void**>(&w.h_status), ST_COUNT * sizeof(unsigned long long), cudaHostAllocMapped));
            CUDA_TRY(cudaHostGetDevicePointer(reinterpret_cast<void**>(&w.h_status_dev), w.h_status, 0));
        }
        // mapped: small read-backs are written by a kernel straight into host memory, so they never queue behind
        // a large hit transfer on the copy engine
        CUDA_TRY(cudaHostAlloc(reinterpret_cast<void**>(&c->h_counters), (CT_COUNT + 8) * sizeof(unsigned long long), cudaHostAllocMapped));
        CUDA_TRY(cudaHostGetDevicePointer(reinterpret_cast<void**>(&c->h_counters_dev), c->h_counters, 0));
        cudaDeviceProp prop;
        CUDA_TRY(cudaGetDeviceProperties(&prop, device));
        c->sms = prop.multiProcessorCount;
        *out = c;
    });
}

int sb200_destroy(sb200_ctx* c) {
    return guard([&] {
        if (!c) return;
        cudaSetDevice(c->device);
        cudaStreamSynchronize(c->stream);
        cudaDeviceSynchronize();
        c->idx.release();
        for (DevBuf* b : {&c->d_steps, &c->d_runs, &c->d_tmp, &c->d_scratch, &c->d_counters}) b->release();
        for (auto& ev : c->ev) cudaEventDestroy(ev);
        for (auto& w : c->work) {
            w.release_buffers();
            for (auto& ev : w.ev) cudaEventDestroy(ev);
            for (cudaEvent_t ev : {w.ev_in, w.ev_done, w.ev_ready, w.ev_out, w.ev_fork}) cudaEventDestroy(ev);
            cudaStreamDestroy(w.own_stream);
            cudaFreeHost(w.h_status);
            if (w.h_out) cudaFreeHost(w.h_out);
            if (w.h_ends) cudaFreeHost(w.h_ends);
        }
        cudaStreamDestroy(c->s_in);
        cudaStreamDestroy(c->s_out);
        cudaFreeHost(c->h_counters);
        cudaStreamDestroy(c->own_stream);
        delete c;
    });
}

int sb200_set_stream(sb200_ctx* c, void* s) {
    return guard([&] {
        use(c);
        c->stream = s ? static_cast<cudaStream_t>(s) : c->own_stream;
    });
}

int sb200_synchronize(sb200_ctx* c) {
    return guard([&] {
        use(c);
        CUDA_TRY(cudaStreamSynchronize(c->stream));
    });
}

int sb200_index_upload(sb200_ctx* c, const sb200_index_view* v) {
    return guard([&] {
        use(c);
        if (!v) throw Error("null index view");
        if (v->sigma != 5 && v->sigma != 6) throw Error("unknown index with " + std::to_string(v->sigma) + " letters");
        if (v->n_rows == 0 || v->n_rows >= (1ull << 32) - 8192) throw Error("index layout not understood: row count out of range (needs < 2^32 rows)");
        if (v->n_blocks != v->n_rows / 64 + 1) throw Error("index layout not understood: block count does not match the row count");
        if (v->sampling_rate == 0 || v->bits_for_position == 0 || v->bits_for_position > 56)
            throw Error("index layout not understood: sampling parameters");
        if (v->C[0] != 0 || v->C[v->sigma] != v->n_rows) throw Error("index layout not understood: C array");
        for (uint32_t s = 0; s < v->sigma; ++s)
            if (v->C[s] > v->C[s + 1]) throw Error("index layout not understood: C array is not monotone");
        auto& ix = c->idx;
        ix.release();
        ix.sigma = static_cast<uint32_t>(v->sigma);
        ix.n_rows = v->n_rows;
        ix.n_blocks = v->n_blocks;
        ix.n_sup = v->n_blocks / 64 + 1;
        for (int i = 0; i < 8; ++i) {
            ix.C64[i] = i <= static_cast<int>(v->sigma) ? v->C[i] : v->n_rows;
            ix.C[i] = static_cast<uint32_t>(ix.C64[i]);
        }
        upload_occ(c, ix.sigma, ix.n_rows, ix.n_blocks, v->bwt_blocks, v->bwt_super, ix.bwt_blk, ix.bwt_sup);
        upload_occ(c, ix.sigma, ix.n_rows, ix.n_blocks, v->bwtrev_blocks, v->bwtrev_super, ix.rev_blk, ix.rev_sup);
        verify_histograms(c);
        uint64_t n_words = v->n_rows / 64 + 1;
        ix.ref_mark_words.reserve(n_words * 8);
        CUDA_TRY(cudaMemcpyAsync(ix.ref_mark_words.p, v->mark_bits, n_words * 8, cudaMemcpyHostToDevice, c->stream));
        uint64_t total = 0;
        build_mark_records(c, ix.ref_mark_words.get<uint64_t>(), ix.n_rows, ix.marks, ix.n_mark, total);
        if (total != v->n_ssa) throw Error("index layout not understood: number of marked rows differs from the number of samples");
        ix.ssa.reserve(std::max<uint64_t>(1, v->n_ssa) * 8);
        CUDA_TRY(cudaMemcpyAsync(ix.ssa.p, v->ssa, v->n_ssa * 8, cudaMemcpyHostToDevice, c->stream));
        ix.n_ssa = v->n_ssa;
        ix.n_ref_ssa = v->n_ssa;
        ix.sampling_rate = v->sampling_rate;
        ix.device_rate = v->sampling_rate;
        ix.bits_for_position = v->bits_for_position;
        ix.full_sa = false;
        finish_index(c);
    });
}


int sb200_index_build_device(sb200_ctx* c, const uint8_t* d_seq_ranks, const uint64_t* seq_lens, uint64_t n_seqs, uint32_t sigma,
                             uint32_t sampling_rate) {
    return guard([&] {
        use(c);
        build_index_device(c, d_seq_ranks, seq_lens, n_seqs, sigma, sampling_rate);
    });
}

int sb200_index_build(sb200_ctx* c, const uint8_t* seq_ranks, const uint64_t* seq_lens, uint64_t n_seqs, uint32_t sigma,
                      uint32_t sampling_rate) {
    return guard([&] {
        use(c);
        uint64_t total = 0;
        for (uint64_t i = 0; i < n_seqs; ++i) total += seq_lens[i];
        DevBuf src;
        src.reserve(std::max<uint64_t>(1, total));
        CUDA_TRY(cudaMemcpyAsync(src.p, seq_ranks, total, cudaMemcpyHostToDevice, c->stream));
        try {
            build_index_device(c, src.get<uint8_t>(), seq_lens, n_seqs, sigma, sampling_rate);
        } catch (...) {
            src.release();
            throw;
        }
        src.release();
    });
}

int sb200_synth_genome_device(sb200_ctx* c, uint64_t n_bases, uint64_t seed, uint8_t* d_out) {
    return guard([&] {
        use(c);
        if (n_bases == 0) return;
        synth_genome_kernel<<<grid_for(n_bases), 256, 0, c->stream>>>(n_bases, seed, d_out);
        launch_check(c);
        CUDA_TRY(cudaStreamSynchronize(c->stream));
    });
}

int sb200_synth_reads_device(sb200_ctx* c, const uint8_t* d_genome, uint64_t n_bases, uint64_t n_reads, uint32_t len, uint32_t k,
                             int edit, uint64_t seed, uint64_t first_read, uint8_t* d_out) {
    return guard([&] {
        use(c);
        if (len == 0 || len > kMaxSynthLen - 16) throw Error("synthetic read length out of range");
        if (k > 8) throw Error("synthetic reads support at most 8 errors");
        if (n_bases < uint64_t(len) + k + 1) throw Error("genome shorter than a read");
        if (n_reads == 0) return;
        synth_reads_kernel<<<grid_for(n_reads, 128), 128, 0, c->stream>>>(d_genome, n_bases, n_reads, len, k, edit, seed, first_read, d_out);
        launch_check(c);
        CUDA_TRY(cudaStreamSynchronize(c->stream));
    });
}

int sb200_index_info_get(sb200_ctx* c, sb200_index_info* out) {
    return guard([&] {
        use(c);
        auto& ix = c->idx;
        if (!ix.loaded) throw Error("no index loaded");
        out->sigma = ix.sigma;
        out->n_rows = ix.n_rows;
        out->n_ssa = ix.n_ref_ssa;
        out->sampling_rate = ix.sampling_rate;
        out->bits_for_position = ix.bits_for_position;
        out->device_sampling_rate = ix.device_rate;
        out->device_bytes = ix.bytes();
        for (int i = 0; i < 8; ++i) out->C[i] = ix.C64[i];
    });
}

int sb200_index_download(sb200_ctx* c, sb200_index_view* out) {
    return guard([&] {
        use(c);
        auto& ix = c->idx;
        if (!ix.loaded) throw Error("no index loaded");
        std::memset(out, 0, sizeof(*out));
        uint64_t stride = 10ull * ix.sigma;
        uint64_t n_super = (ix.n_blocks + 1023) / 1024;
        DevBuf raw, dsuper;
        raw.reserve(ix.n_blocks * stride + 64);
        dsuper.reserve(n_super * ix.sigma * 8);
        auto pull = [&](OccTable t, void** blocks, uint64_t** super) {
            with_sigma(ix.sigma, [&](auto S) {
                export_ref_occ_kernel<S()><<<grid_for(ix.n_blocks), 256, 0, c->stream>>>(t.blk, t.sup, ix.n_blocks, ix.n_rows,
                                                                                       raw.get<uint8_t>(), dsuper.get<uint64_t>());
                return 0;
            });
            launch_check(c);
            *blocks = std::malloc(ix.n_blocks * stride);
            *super = static_cast<uint64_t*>(std::malloc(n_super * ix.sigma * 8));
            if (!*blocks || !*super) throw Error("out of host memory");
            CUDA_TRY(cudaMemcpyAsync(*blocks, raw.p, ix.n_blocks * stride, cudaMemcpyDeviceToHost, c->stream));
            CUDA_TRY(cudaMemcpyAsync(*super, dsuper.p, n_super * ix.sigma * 8, cudaMemcpyDeviceToHost, c->stream));
            CUDA_TRY(cudaStreamSynchronize(c->stream));
        };
        void* b1 = nullptr; void* b2 = nullptr;
        uint64_t* s1 = nullptr; uint64_t* s2 = nullptr;
        pull(ix.bwt(), &b1, &s1);
        out->bwt_blocks = b1; out->bwt_super = s1;
        pull(ix.rev(), &b2, &s2);
        out->bwtrev_blocks = b2; out->bwtrev_super = s2;
        raw.release();
        dsuper.release();
        out->sigma = ix.sigma;
        out->n_rows = ix.n_rows;
        out->n_blocks = ix.n_blocks;
        uint64_t* C = static_cast<uint64_t*>(std::malloc(8 * (ix.sigma + 1)));
        for (uint32_t i = 0; i <= ix.sigma; ++i) C[i] = ix.C64[i];
        out->C = C;
        uint64_t n_words = ix.n_rows / 64 + 1;
        uint64_t* words = static_cast<uint64_t*>(std::malloc(n_words * 8));
        CUDA_TRY(cudaMemcpyAsync(words, ix.ref_mark_words.p, n_words * 8, cudaMemcpyDeviceToHost, c->stream));
        out->mark_bits = words;
        const DevBuf& src = ix.ref_ssa.p ? ix.ref_ssa : ix.ssa;
        uint64_t* ssa = static_cast<uint64_t*>(std::malloc(std::max<uint64_t>(1, ix.n_ref_ssa) * 8));
        CUDA_TRY(cudaMemcpyAsync(ssa, src.p, ix.n_ref_ssa * 8, cudaMemcpyDeviceToHost, c->stream));
        CUDA_TRY(cudaStreamSynchronize(c->stream));
        out->ssa = ssa;
        out->n_ssa = ix.n_ref_ssa;
        out->sampling_rate = ix.sampling_rate;
        out->bits_for_position = ix.bits_for_position;
    });
}

void sb200_index_view_free(sb200_index_view* v) {
    if (!v) return;
    std::free(const_cast<void*>(v->bwt_blocks));
    std::free(const_cast<uint64_t*>(v->bwt_super));
    std::free(const_cast<void*>(v->bwtrev_blocks));
    std::free(const_cast<uint64_t*>(v->bwtrev_super));
    std::free(const_cast<uint64_t*>(v->C));
    std::free(const_cast<uint64_t*>(v->ssa));
    std::free(const_cast<uint64_t*>(v->mark_bits));
    std::memset(v, 0, sizeof(*v));
}

static void densify_index(sb200_ctx* c, uint32_t rate);

int sb200_index_densify(sb200_ctx* c, uint32_t rate) {
    return guard([&] {
        use(c);
        densify_index(c, rate);
    });
}

int sb200_index_enable_text(sb200_ctx* c, int enable) {
    return guard([&] {
        use(c);
        auto& ix = c->idx;
        if (!ix.loaded) throw Error("no index loaded");
        if (!enable) {
            for (DevBuf* b : {&ix.sa32, &ix.isa32, &ix.text4, &ix.seq_start}) b->release();
            ix.text_mode = false;
            return;
        }
        if (ix.text_mode) return;
        densify_index(c, 1);  // complete suffix array as (seqId, pos)
        uint64_t n = ix.n_rows;
        // sequence starts from the delimiter rows: rows [0, C[1]) are the suffixes that start with a delimiter,
        // their value is (seqId, length of that sequence)
        uint64_t n_seqs = ix.C64[1];
        if (n_seqs == 0) throw Error("index has no delimiter");
        std::vector<uint64_t> vals(n_seqs), start(n_seqs + 1, 0), lens(n_seqs, 0);
        CUDA_TRY(cudaMemcpyAsync(vals.data(), ix.ssa.p, n_seqs * 8, cudaMemcpyDeviceToHost, c->stream));
        CUDA_TRY(cudaStreamSynchronize(c->stream));
        uint64_t mask = (uint64_t{1} << ix.bits_for_position) - 1;
        for (uint64_t v : vals) {
            uint64_t sid = v >> ix.bits_for_position;
            if (sid >= n_seqs) throw Error("index layout not understood: sequence id of a delimiter row out of range");
            lens[sid] = v & mask;
        }
        for (uint64_t i = 0; i < n_seqs; ++i) start[i + 1] = start[i] + lens[i] + 1;
        if (start[n_seqs] != n) throw Error("index layout not understood: sequence lengths do not add up to the text length");
        DevBuf& d_start = ix.seq_start;
        d_start.reserve((n_seqs + 1) * 8);
        CUDA_TRY(cudaMemcpyAsync(d_start.p, start.data(), (n_seqs + 1) * 8, cudaMemcpyHostToDevice, c->stream));
        ix.n_seqs = n_seqs;
        ix.sa32.reserve(n * 4);
        ix.isa32.reserve(n * 4);
        uint64_t n_words = n / 8 + 2;
        ix.text4.reserve(n_words * 4);
        CUDA_TRY(cudaMemsetAsync(ix.text4.p, 0, n_words * 4, c->stream));
        sa32_kernel<<<grid_for(n), 256, 0, c->stream>>>(ix.ssa.get<uint64_t>(), d_start.get<uint64_t>(), n,
                                                         static_cast<uint32_t>(ix.bits_for_position), ix.sa32.get<uint32_t>(),
                                                         ix.isa32.get<uint32_t>());
        launch_check(c);
        with_sigma(ix.sigma, [&](auto S) {
            text4_kernel<S()><<<grid_for(n), 256, 0, c->stream>>>(ix.bwt(), ix.sa32.get<uint32_t>(), n, ix.text4.get<uint32_t>());
            return 0;
        });
        launch_check(c);
        CUDA_TRY(cudaStreamSynchronize(c->stream));
        ix.text_mode = true;
    });
}

static void densify_index(sb200_ctx* c, uint32_t rate) {
    {
        auto& ix = c->idx;
        if (!ix.loaded) throw Error("no index loaded");
        if (rate == 0 || (rate & (rate - 1)) || rate > ix.device_rate) throw Error("device sampling rate must be a power of two not above the current rate");
        if (rate == ix.device_rate) return;
        uint64_t n = ix.n_rows;
        uint64_t n_words = n / 64 + 1;
        DevBuf new_words, row_value;
        new_words.reserve(n_words * 8);
        row_value.reserve(n * 8);
        CUDA_TRY(cudaMemsetAsync(new_words.p, 0, n_words * 8, c->stream));
        DensifyParams D{};
        D.index.bwt = ix.bwt();
        for (int i = 0; i < 8; ++i) D.index.C[i] = ix.C[i];
        D.index.marks = ix.full_sa ? nullptr : ix.marks.get<MarkRec>();
        D.index.ssa = ix.ssa.get<uint64_t>();
        D.n_rows = static_cast<uint32_t>(n);
        D.new_rate = rate;
        D.pos_mask = (uint64_t{1} << ix.bits_for_position) - 1;
        D.new_mark_bits = new_words.get<uint64_t>();
        D.row_value = row_value.get<uint64_t>();
        with_sigma(ix.sigma, [&](auto S) {
            densify_kernel<S()><<<grid_for(n), 256, 0, c->stream>>>(D);
            return 0;
        });
        launch_check(c);
        if (!ix.ref_ssa.p) {  // keep the reference-rate samples for download
            ix.ref_ssa = std::move(ix.ssa);
        } else {
            ix.ssa.release();
        }
        if (rate == 1) {
            CUDA_TRY(cudaStreamSynchronize(c->stream));
            ix.ssa = std::move(row_value);
            ix.n_ssa = n;
            ix.full_sa = true;
            ix.marks.release();
        } else {
            uint64_t total = 0;
            build_mark_records(c, new_words.get<uint64_t>(), n, ix.marks, ix.n_mark, total);
            ix.ssa.reserve(std::max<uint64_t>(1, total) * 8);
            // compact row_value at marked rows, in row order
            DevBuf d_num;
            d_num.reserve(8);
            auto flags = cub::TransformInputIterator<bool, BitFlag, cub::CountingInputIterator<uint64_t>>(
                cub::CountingInputIterator<uint64_t>(0), BitFlag{new_words.get<uint64_t>()});
            size_t tmp_bytes = 0;
            CUDA_TRY(cub::DeviceSelect::Flagged(nullptr, tmp_bytes, row_value.get<uint64_t>(), flags, ix.ssa.get<uint64_t>(),
                                                d_num.get<uint64_t>(), static_cast<int64_t>(n), c->stream));
            c->d_tmp.reserve(tmp_bytes);
            CUDA_TRY(cub::DeviceSelect::Flagged(c->d_tmp.p, tmp_bytes, row_value.get<uint64_t>(), flags, ix.ssa.get<uint64_t>(),
                                                d_num.get<uint64_t>(), static_cast<int64_t>(n), c->stream));
            c->ct.kernel_launches += 1;
            uint64_t got = 0;
            CUDA_TRY(cudaMemcpyAsync(&got, d_num.p, 8, cudaMemcpyDeviceToHost, c->stream));
            CUDA_TRY(cudaStreamSynchronize(c->stream));
            d_num.release();
            if (got != total) throw Error("internal error: densify produced inconsistent sample counts");
            ix.n_ssa = total;
            ix.full_sa = false;
        }
        new_words.release();
        row_value.release();
        ix.device_rate = rate;
    }
}

int sb200_index_build_qgram(sb200_ctx* c, uint32_t q) {
    return guard([&] {
        use(c);
        auto& ix = c->idx;
        if (!ix.loaded) throw Error("no index loaded");
        if (q > 15) throw Error("q-gram length above 15 is not supported");
        ix.qgram.release();
        ix.qgram_q = 0;
        if (q == 0) return;
        DevBuf a, b;
        uint64_t n_final = 1ull << (2 * q);
        a.reserve(n_final * sizeof(uint4));
        b.reserve(std::max<uint64_t>(1, n_final / 4) * sizeof(uint4));
        // ping-pong so that the last level lands in `a`
        DevBuf* cur = (q % 2 == 0) ? &a : &b;
        DevBuf* nxt = (q % 2 == 0) ? &b : &a;
        uint4 root = make_uint4(0, 0, static_cast<uint32_t>(ix.n_rows), 0);
        CUDA_TRY(cudaMemcpyAsync(cur->p, &root, sizeof(uint4), cudaMemcpyHostToDevice, c->stream));
        for (uint32_t t = 1; t <= q; ++t) {
            uint32_t n_child = 1u << (2 * t);
            with_sigma(ix.sigma, [&](auto S) {
                qgram_level_kernel<S()><<<grid_for(n_child), 256, 0, c->stream>>>(ix.rev(), ix.d_C.get<uint32_t>(), cur->get<uint4>(),
                                                                                 nxt->get<uint4>(), n_child,
                                                                                 static_cast<uint32_t>(ix.n_rows));
                return 0;
            });
            launch_check(c);
            std::swap(cur, nxt);
        }
        CUDA_TRY(cudaStreamSynchronize(c->stream));
        if (cur != &a) throw Error("internal error: q-gram ping-pong");
        ix.qgram = std::move(a);
        b.release();
        ix.qgram_q = q;
    });
}

int sb200_index_clone(sb200_ctx* dst, sb200_ctx* src) {
    return guard([&] {
        if (!src || !dst || src == dst) throw Error("sb200_index_clone needs two different contexts");
        if (!src->idx.loaded) throw Error("no index loaded in the source context");
        use(dst);
        if (dst->device != src->device) {  // direct GPU-to-GPU copies where the devices are peers (else the runtime stages them)
            int can = 0;
            CUDA_TRY(cudaDeviceCanAccessPeer(&can, dst->device, src->device));
            if (can) {
                const cudaError_t e = cudaDeviceEnablePeerAccess(src->device, 0);
                if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) CUDA_TRY(e);
                cudaGetLastError();
            }
        }
        DeviceIndex& d = dst->idx;
        const DeviceIndex& s = src->idx;
        d.release();
        auto copy = [&](DevBuf& to, const DevBuf& from) {
            if (!from.p) return;
            CUDA_TRY(cudaMalloc(&to.p, from.cap));
            to.cap = from.cap;
            CUDA_TRY(cudaMemcpyPeerAsync(to.p, dst->device, from.p, src->device, from.cap, dst->stream));
        };
        copy(d.bwt_blk, s.bwt_blk); copy(d.bwt_sup, s.bwt_sup); copy(d.rev_blk, s.rev_blk); copy(d.rev_sup, s.rev_sup);
        copy(d.d_C, s.d_C); copy(d.marks, s.marks); copy(d.ssa, s.ssa); copy(d.ref_mark_words, s.ref_mark_words);
        copy(d.ref_ssa, s.ref_ssa); copy(d.qgram, s.qgram); copy(d.sa32, s.sa32); copy(d.isa32, s.isa32); copy(d.text4, s.text4);
        copy(d.seq_start, s.seq_start);
        d.sigma = s.sigma; d.n_rows = s.n_rows; d.n_blocks = s.n_blocks; d.n_sup = s.n_sup;
        for (int i = 0; i < 8; ++i) { d.C[i] = s.C[i]; d.C64[i] = s.C64[i]; }
        d.n_mark = s.n_mark; d.full_sa = s.full_sa; d.n_ssa = s.n_ssa; d.sampling_rate = s.sampling_rate;
        d.bits_for_position = s.bits_for_position; d.device_rate = s.device_rate; d.key_bits = s.key_bits; d.n_ref_ssa = s.n_ref_ssa;
        d.qgram_q = s.qgram_q; d.n_seqs = s.n_seqs; d.text_mode = s.text_mode;
        CUDA_TRY(cudaStreamSynchronize(dst->stream));
        d.loaded = true;
    });
}

int sb200_set_scheme(sb200_ctx* c, uint32_t n_searches, uint32_t len, const uint16_t* pi, const uint8_t* l, const uint8_t* u, int edit) {
    return guard([&] {
        use(c);
        if (n_searches == 0 || len == 0) throw Error("empty search scheme");
        if (len > 1000) throw Error("queries longer than 1000 characters are not supported");
        std::vector<uint32_t> steps(size_t(n_searches) * len);
        uint32_t kmax = 0;
        for (uint32_t j = 0; j < n_searches; ++j) {
            std::vector<bool> seen(len, false);
            for (uint32_t i = 0; i < len; ++i) {
                size_t k = size_t(j) * len + i;
                if (pi[k] >= len || seen[pi[k]]) throw Error("search scheme does not fit the query length (pi is not a permutation)");
                seen[pi[k]] = true;
                if (l[k] > u[k] || u[k] > 15) throw Error("search scheme has invalid error bounds");
                if (i > 0 && (l[k] < l[k - 1] || u[k] < u[k - 1])) throw Error("search scheme bounds must be non-decreasing");
                bool right = (i == 0) ? (len < 2 || pi[k] < pi[k + 1]) : (pi[k - 1] < pi[k]);
                steps[k] = pack_step(pi[k], l[k], u[k], right);
                kmax = std::max<uint32_t>(kmax, u[k]);
            }
        }
        if (kmax > 4) throw Error("search schemes with more than 4 errors are not supported by the GPU kernel yet");
        for (auto& w : c->work)
            if (w.busy) throw Error("submitted batches are still in flight: wait for them before changing the scheme");
        CUDA_TRY(cudaDeviceSynchronize());
        c->h_steps = std::move(steps);
        c->n_searches = n_searches;
        c->qlen = len;
        c->kmax = kmax;
        c->edit = edit != 0;
        upload_scheme_tables(c);
        c->have_scheme = true;
    });
}

int sb200_set_policy(sb200_ctx* c, const sb200_policy* policy) {
    return guard([&] {
        use(c);
        if (!policy || !sb200_pol_valid(policy)) throw Error("invalid search policy");
        for (auto& w : c->work)
            if (w.busy) throw Error("submitted batches are still in flight: wait for them before changing the policy");
        CUDA_TRY(cudaDeviceSynchronize());
        c->policy = *policy;
        if (c->have_scheme) upload_scheme_tables(c);  // (the state flags depend on it)
    });
}

int sb200_get_policy(sb200_ctx* c, sb200_policy* out) {
    return guard([&] {
        if (!c || !out) throw Error("null argument");
        *out = c->policy;
    });
}

int sb200_set_option(sb200_ctx* c, const char* name, int64_t value) {
    return guard([&] {
        if (!c || !name) throw Error("null argument");
        const std::string n(name);
        auto& o = c->opt;
        if (n == "bucket_sort") o.bucket_sort = value != 0;
        else if (n == "fused_sort") o.fused_sort = value < 0 ? -1 : (value != 0);
        else if (n == "textpos") o.textpos = value != 0;
        else if (n == "ordered_only") o.ordered_only = value != 0;
        else if (n == "debug") o.debug = static_cast<int>(value);
        else if (n == "chunk") o.chunk = value <= 0 ? Options{}.chunk : static_cast<uint64_t>(value);
        else if (n == "edge_div") o.edge_div = value <= 0 ? Options{}.edge_div : static_cast<uint64_t>(value);
        else if (n == "pool_blocks_per_sm") o.pool_blocks_per_sm = static_cast<int>(std::max<int64_t>(0, value));
        else if (n == "pool_threads") o.pool_threads = static_cast<int>(std::max<int64_t>(0, value));
        else if (n == "run_rounds") o.run_rounds = static_cast<int>(std::max<int64_t>(0, value));
        else if (n == "items_blocks_per_sm") o.items_blocks_per_sm = static_cast<int>(std::max<int64_t>(0, value));
        else if (n == "ordered_blocks_per_sm") o.ordered_blocks_per_sm = static_cast<int>(std::max<int64_t>(0, value));
        else if (n == "delta_records") o.delta_records = value != 0;
        else if (n == "overlap") {
            if (value < 0 || value > 2) throw Error("option overlap takes 0, 1 or 2");
            for (auto& w : c->work)
                if (w.busy) throw Error("submitted batches are still in flight");
            o.overlap = static_cast<int>(value);
            c->prev_slot = nullptr;
        }
        else throw Error("unknown option '" + n + "'");
    });
}

int sb200_set_max_hits(sb200_ctx* c, uint64_t max_hits) {
    return guard([&] {
        use(c);
        if (max_hits > 0xfffffffeull) throw Error("--max_hits above 2^32 - 2 is not supported");
        c->max_hits = static_cast<uint32_t>(max_hits);
    });
}

int sb200_search_device(sb200_ctx* c, const uint8_t* d_queries, uint64_t n_queries, uint32_t len, uint64_t* n_cursors, uint64_t* n_hits) {
    return guard([&] {
        use(c);
        run_pipeline(c, d_queries, n_queries, len, n_hits != nullptr);
        if (n_cursors) *n_cursors = c->last->n_real_cursors;
        if (n_hits) *n_hits = c->last->n_hits;
    });
}

int sb200_fetch_hits(sb200_ctx* c, sb200_hit** hits, uint64_t* n_hits) {
    return guard([&] {
        use(c);
        fetch_hits(c, hits, n_hits);
    });
}

int sb200_search(sb200_ctx* c, const uint8_t* queries, uint64_t n_queries, uint32_t len, sb200_hit** hits, uint64_t* n_hits) {
    return guard([&] {
        use(c);
        void* p = nullptr;
        search_host_pipelined(c, queries, n_queries, len, IN_QUERIES_RANKS, false, OUT_HIT64, &p, n_hits);
        *hits = static_cast<sb200_hit*>(p);
    });
}

int sb200_search_reads(sb200_ctx* c, const uint8_t* reads, uint64_t n_reads, uint32_t len, int with_reverse, sb200_hit32** hits,
                       uint64_t* n_hits) {
    return guard([&] {
        use(c);
        if (c->idx.loaded && c->idx.bits_for_position > 32) throw Error("sequences too long for 32-bit positions: use sb200_search");
        void* p = nullptr;
        search_host_pipelined(c, reads, n_reads, len, IN_READS_RANKS, with_reverse != 0, OUT_HIT32, &p, n_hits);
        *hits = static_cast<sb200_hit32*>(p);
    });
}

// ---- asynchronous batches ------------------------------------------------------------------------------

static Work& slot_of(sb200_ctx* c, uint64_t ticket) {
    for (auto& w : c->work)
        if (w.busy && w.ticket == ticket) return w;
    throw Error("unknown batch ticket");
}

static uint64_t submit_batch(sb200_ctx* c, const void* src, bool on_device, uint64_t n_items, uint32_t len, int in_fmt, bool with_reverse,
                             int out_fmt) {
    Work* free_slot = nullptr;
    for (auto& w : c->work)
        if (!w.busy) { free_slot = &w; break; }
    if (!free_slot) throw Error("too many batches in flight (" + std::to_string(kSlots) + "): wait for one and release it first");
    Work& w = *free_slot;
    const uint64_t n_queries = (with_reverse && in_fmt != IN_QUERIES_RANKS) ? 2 * n_items : n_items;
    // fork from the caller's stream: work queued there before this call happens before the batch
    CUDA_TRY(cudaEventRecord(w.ev_fork, c->stream));
    CUDA_TRY(cudaStreamWaitEvent(w.own_stream, w.ev_fork, 0));
    if (c->opt.overlap == 0) CUDA_TRY(cudaStreamWaitEvent(c->work[0].own_stream, w.ev_fork, 0));
    if (!on_device) CUDA_TRY(cudaStreamWaitEvent(c->s_in, w.ev_fork, 0));
    setup_batch(c, w, src, on_device, in_fmt, with_reverse, n_queries, len, true, out_fmt, 0, w.own_stream);
    enqueue_compute(c, w);
    w.busy = true;
    w.ticket = c->next_ticket++;
    return w.ticket;
}

int sb200_submit_reads(sb200_ctx* c, const void* reads, uint64_t n_reads, uint32_t len, int format, int with_reverse, uint64_t* ticket) {
    return guard([&] {
        use(c);
        if (!ticket) throw Error("null argument");
        if (format != SB200_READS_RANKS && format != SB200_READS_PACKED4 && format != SB200_READS_PACKED2) throw Error("unknown read format");
        if (c->max_hits) throw Error("sb200_submit_reads does not support --max_hits: use sb200_search_reads");
        *ticket = submit_batch(c, reads, false, n_reads, len, format == SB200_READS_PACKED4 ? IN_READS_PACKED4 : format == SB200_READS_PACKED2 ? IN_READS_PACKED2 : IN_READS_RANKS,
                               with_reverse != 0, OUT_CSR);
    });
}

int sb200_submit_device(sb200_ctx* c, const uint8_t* d_queries, uint64_t n_queries, uint32_t len, uint64_t* ticket) {
    return guard([&] {
        use(c);
        if (!ticket) throw Error("null argument");
        if (c->max_hits) throw Error("sb200_submit_device does not support --max_hits: use sb200_search_device");
        *ticket = submit_batch(c, d_queries, true, n_queries, len, IN_QUERIES_RANKS, false, OUT_CSR);
    });
}

int sb200_wait_batch(sb200_ctx* c, uint64_t ticket, int copy_to_host, sb200_batch_result* out) {
    return guard([&] {
        use(c);
        if (!out) throw Error("null argument");
        Work& w = slot_of(c, ticket);
        std::memset(out, 0, sizeof(*out));
        try {
            finish_batch(c, w);
        } catch (...) {
            w.busy = false;
            throw;
        }
        const size_t rec = out_record_bytes(c, OUT_CSR);
        out->n_queries = w.n_queries;
        out->n_hits = w.n_hits;
        out->n_cursors = w.n_real_cursors;
        out->record_bytes = static_cast<uint32_t>(rec);
        out->bits_for_position = static_cast<uint32_t>(c->idx.bits_for_position);
        out->delta_coded = w.delta ? 1u : 0u;
        out->n_record_bytes = w.delta ? w.n_rec_bytes : w.n_hits * rec;
        if (copy_to_host) {
            const size_t need = (w.delta ? w.n_rec_bytes : (w.n_hits + 4) * rec) + 64, need_ends = (w.n_queries + 1) * 4;
            if (need > w.h_out_cap) {
                if (w.h_out) cudaFreeHost(w.h_out);
                w.h_out = nullptr;
                w.h_out_cap = 0;
                CUDA_TRY(cudaHostAlloc(&w.h_out, need + need / 4, cudaHostAllocDefault));
                w.h_out_cap = need + need / 4;
            }
            if (need_ends > w.h_ends_cap) {
                if (w.h_ends) cudaFreeHost(w.h_ends);
                w.h_ends = nullptr;
                w.h_ends_cap = 0;
                CUDA_TRY(cudaHostAlloc(reinterpret_cast<void**>(&w.h_ends), need_ends + need_ends / 4, cudaHostAllocDefault));
                w.h_ends_cap = need_ends + need_ends / 4;
            }
            enqueue_copy_out(c, w, w.h_out, w.h_ends);
            CUDA_TRY(cudaEventSynchronize(w.ev_out));
            out->hit_end = w.h_ends;
            out->records = static_cast<const uint8_t*>(w.h_out);
            out->h2d_bytes = w.h2d_bytes;
            out->d2h_bytes = w.d2h_bytes;
            // join: work queued on the caller's stream after this call happens after the batch
            CUDA_TRY(cudaStreamWaitEvent(c->stream, w.ev_out, 0));
        } else {
            CUDA_TRY(cudaStreamWaitEvent(c->stream, w.ev_done, 0));
        }
        out->ms_search = w.ms_search;
        out->ms_locate = w.ms_locate;
        out->ms_sort = w.ms_sort;
    });
}

int sb200_release_batch(sb200_ctx* c, uint64_t ticket) {
    return guard([&] {
        if (!c) throw Error("null context");
        slot_of(c, ticket).busy = false;
    });
}

int sb200_search_cursors(sb200_ctx* c, const uint8_t* queries, uint64_t n_queries, uint32_t len, sb200_cursor** cursors, uint64_t* n_cursors) {
    return guard([&] {
        use(c);
        if (!c->idx.loaded) throw Error("no index loaded");
        if (!queries || n_queries == 0) throw Error("query file was empty - abort");
        Work& w = c->work[0];
        if (w.busy) throw Error("a submitted batch is still in flight in slot 0: wait for it first");
        setup_batch(c, w, queries, false, IN_QUERIES_RANKS, false, n_queries, len, false, OUT_NONE, 0, c->stream);
        enqueue_compute(c, w);
        finish_batch(c, w);
        c->last = &w;
        uint64_t n = w.n_cursor_slots;
        std::vector<uint4> tmp(n);
        CUDA_TRY(cudaMemcpyAsync(tmp.data(), w.d_cursors.p, n * sizeof(uint4), cudaMemcpyDeviceToHost, c->stream));
        CUDA_TRY(cudaStreamSynchronize(c->stream));
        tmp.erase(std::remove_if(tmp.begin(), tmp.end(), [](uint4 const& a) { return a.x == kInvalidQid; }), tmp.end());
        if (tmp.size() != w.n_real_cursors) throw Error("internal error: cursor count mismatch");
        n = tmp.size();
        std::sort(tmp.begin(), tmp.end(), [](uint4 const& a, uint4 const& b) {
            if (a.x != b.x) return a.x < b.x;
            if (a.y != b.y) return a.y < b.y;
            if (a.z != b.z) return a.z < b.z;
            return a.w < b.w;
        });
        sb200_cursor* out = static_cast<sb200_cursor*>(g_pinned.alloc(std::max<uint64_t>(1, n) * sizeof(sb200_cursor)));
        for (uint64_t i = 0; i < n; ++i) out[i] = sb200_cursor{tmp[i].x, tmp[i].y, tmp[i].z, tmp[i].w};
        *cursors = out;
        *n_cursors = n;
    });
}

int sb200_locate(sb200_ctx* c, const sb200_cursor* cursors, uint64_t n_cursors, sb200_hit** hits, uint64_t* n_hits) {
    return guard([&] {
        use(c);
        auto& ix = c->idx;
        if (!ix.loaded) throw Error("no index loaded");
        if (n_cursors >= 0xffffffffull) throw Error("too many cursors for one call");
        std::vector<uint4> tmp(n_cursors + 1);
        for (uint64_t i = 0; i < n_cursors; ++i) {
            auto const& k = cursors[i];
            // (written so that huge values cannot wrap around the check)
            if (k.len > ix.n_rows || k.lb > ix.n_rows - k.len || k.errors > 15 || k.query_id > 0xffffffffull) throw Error("cursor out of range");
            tmp[i] = make_uint4(static_cast<uint32_t>(k.query_id), static_cast<uint32_t>(k.lb), static_cast<uint32_t>(k.len),
                                static_cast<uint32_t>(k.errors));
        }
        tmp[n_cursors] = make_uint4(0, 0, 0, 0);
        Work& w = c->work[0];
        if (w.busy) throw Error("a submitted batch is still in flight in slot 0: wait for it first");
        w.stream = c->stream;
        w.d_cursors.reserve((n_cursors + 1) * sizeof(uint4));
        CUDA_TRY(cudaMemcpyAsync(w.d_cursors.p, tmp.data(), (n_cursors + 1) * sizeof(uint4), cudaMemcpyHostToDevice, c->stream));
        CUDA_TRY(cudaStreamSynchronize(c->stream));
        w.out_fmt = OUT_NONE;
        w.first_query = 0;
        w.n_queries = 0;
        w.do_locate = true;
        locate_radix(c, w, n_cursors, 0);
        CUDA_TRY(cudaStreamSynchronize(c->stream));
        CUDA_TRY(cudaEventElapsedTime(&c->ct.ms_locate, w.ev[1], w.ev[2]));
        CUDA_TRY(cudaEventElapsedTime(&c->ct.ms_sort, w.ev[2], w.ev[3]));
        c->ct.lf_steps += w.h_status[ST_COUNTERS + CT_LF_STEPS];
        c->ct.hits += w.n_hits;
        c->last = &w;
        fetch_hits(c, hits, n_hits);
    });
}

void sb200_free(void* p) {
    if (!p) return;
    if (!g_pinned.free(p)) std::free(p);
}

int sb200_host_alloc(uint64_t bytes, void** out) {
    return guard([&] { *out = g_pinned.alloc(bytes); });
}

int sb200_rank_probe(sb200_ctx* c, int which, const uint64_t* positions, uint64_t n, uint64_t* out) {
    return guard([&] {
        use(c);
        auto& ix = c->idx;
        if (!ix.loaded) throw Error("no index loaded");
        for (uint64_t i = 0; i < n; ++i)
            if (positions[i] > ix.n_rows) throw Error("rank position out of range");
        DevBuf dpos, dout;
        dpos.reserve(std::max<uint64_t>(1, n) * 8);
        dout.reserve(std::max<uint64_t>(1, n) * 8 * ix.sigma);
        CUDA_TRY(cudaMemcpyAsync(dpos.p, positions, n * 8, cudaMemcpyHostToDevice, c->stream));
        if (n) {
            with_sigma(ix.sigma, [&](auto S) {
                rank_probe_kernel<S()><<<grid_for(n), 256, 0, c->stream>>>(which ? ix.rev() : ix.bwt(), dpos.get<uint64_t>(), n,
                                                                         dout.get<uint64_t>());
                return 0;
            });
            launch_check(c);
        }
        CUDA_TRY(cudaMemcpyAsync(out, dout.p, n * 8 * ix.sigma, cudaMemcpyDeviceToHost, c->stream));
        CUDA_TRY(cudaStreamSynchronize(c->stream));
        dpos.release();
        dout.release();
    });
}

int sb200_rank_bench(sb200_ctx* c, int which, uint64_t n_chains, uint32_t iters, uint64_t seed, float* ms, uint64_t* checksum) {
    return guard([&] {
        use(c);
        auto& ix = c->idx;
        if (!ix.loaded) throw Error("no index loaded");
        c->d_counters.reserve(CT_COUNT * sizeof(unsigned long long));
        CUDA_TRY(cudaMemsetAsync(c->d_counters.p, 0, CT_COUNT * sizeof(unsigned long long), c->stream));
        CUDA_TRY(cudaEventRecord(c->ev[0], c->stream));
        with_sigma(ix.sigma, [&](auto S) {
            rank_bench_kernel<S()><<<grid_for(n_chains), 256, 0, c->stream>>>(which ? ix.rev() : ix.bwt(), static_cast<uint32_t>(ix.n_rows),
                                                                            n_chains, iters, static_cast<uint32_t>(seed),
                                                                            c->d_counters.get<unsigned long long>());
            return 0;
        });
        launch_check(c);
        CUDA_TRY(cudaEventRecord(c->ev[1], c->stream));
        CUDA_TRY(cudaMemcpyAsync(c->h_counters, c->d_counters.p, 8, cudaMemcpyDeviceToHost, c->stream));
        CUDA_TRY(cudaStreamSynchronize(c->stream));
        CUDA_TRY(cudaEventElapsedTime(ms, c->ev[0], c->ev[1]));
        if (checksum) *checksum = c->h_counters[0];
        c->ct.rank_ops += n_chains * iters;
    });
}

int sb200_get_counters(sb200_ctx* c, sb200_counters* out) {
    return guard([&] {
        if (!c) throw Error("null context");
        *out = c->ct;
        out->ms_fm = c->ms_fm;
        out->ms_text = c->ms_text;
        out->nodes_text = c->nodes_text;
    });
}
int sb200_reset_counters(sb200_ctx* c) {
    return guard([&] {
        if (!c) throw Error("null context");
        c->ct = sb200_counters{};
        c->nodes_text = 0;
    });
}

int sb200_device_alloc(sb200_ctx* c, uint64_t bytes, void** d_ptr) {
    return guard([&] {
        use(c);
        CUDA_TRY(cudaMalloc(d_ptr, std::max<uint64_t>(1, bytes)));
    });
}
int sb200_device_free(sb200_ctx* c, void* d_ptr) {
    return guard([&] {
        use(c);
        CUDA_TRY(cudaFree(d_ptr));
    });
}
int sb200_copy_to_host(sb200_ctx* c, void* dst, const void* d_src, uint64_t bytes) {
    return guard([&] {
        use(c);
        CUDA_TRY(cudaMemcpyAsync(dst, d_src, bytes, cudaMemcpyDeviceToHost, c->stream));
        CUDA_TRY(cudaStreamSynchronize(c->stream));
    });
}
int sb200_copy_to_device(sb200_ctx* c, void* d_dst, const void* src, uint64_t bytes) {
    return guard([&] {
        use(c);
        CUDA_TRY(cudaMemcpyAsync(d_dst, src, bytes, cudaMemcpyHostToDevice, c->stream));
        CUDA_TRY(cudaStreamSynchronize(c->stream));
    });
}

}  // extern "C"
