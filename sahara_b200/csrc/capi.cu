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

struct DevBuf {
    void* p{};
    size_t cap{};
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

}  // namespace

struct sb200_ctx {
    int device{};
    cudaStream_t own_stream{}, stream{};
    DeviceIndex idx;
    // scheme
    DevBuf d_steps, d_runs;
    uint32_t n_searches{}, qlen{}, kmax{};
    bool edit{}, have_scheme{};
    uint32_t max_hits{0};  // > 0: search_n (fm_ordered_kernel), at most this many rows per query
    // work buffers
    uint32_t fused_shift{0};  // hit keys of the last locate carry the query id above this bit (0: separate array)
    int sorted_keys{0};       // d_keys[] buffer that holds the sorted hits
    DevBuf d_qpos, d_tasks, d_bigsegs, d_lc, d_items, d_item_tags, d_ostack, d_rows, d_redo, d_seeds, d_spill, d_packed, d_queries, d_cursors, d_counters, d_offsets, d_keys[2], d_qids[2], d_tmp, d_scratch;
    uint64_t cursor_cap{}, seed_cap{};
    uint64_t last_cursors{}, last_real_cursors{}, last_hits{};
    uint64_t nodes_text{};
    float ms_fm{}, ms_text{};
    bool hits_in_second{};  // which of the double buffers holds the sorted hits
    unsigned long long* h_counters{};  // pinned + mapped, CT_COUNT entries + 8 scratch words
    unsigned long long* h_counters_dev{};  // its device alias
    sb200_counters ct{};
    cudaEvent_t ev[12]{};
    int sms{};
    // pipelined host-buffer search: copy streams, double buffers
    cudaStream_t s_in{}, s_out{};
    DevBuf d_qchunk[2], d_hitchunk[2];
    cudaEvent_t ev_in[2]{}, ev_free_q[2]{}, ev_expanded[2]{}, ev_out[2]{};
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

unsigned blocks_per_sm(const char* env, unsigned def) {
    if (const char* e = std::getenv(env)) return static_cast<unsigned>(std::max(1, std::atoi(e)));
    return def;
}

__global__ void mirror_words_kernel(const unsigned long long* src, unsigned long long* dst, int n) {
    if (threadIdx.x < n) dst[threadIdx.x] = src[threadIdx.x];
}
// n (<= CT_COUNT + 1) device words -> c->h_counters[at ...], then waits for the stream
void read_back_words(sb200_ctx* c, const void* d_src, int n, int at = 0) {
    mirror_words_kernel<<<1, 32, 0, c->stream>>>(static_cast<const unsigned long long*>(d_src), c->h_counters_dev + at, n);
    launch_check(c);
    CUDA_TRY(cudaStreamSynchronize(c->stream));
}

// search_n: the ordered walk (fm_ordered_kernel), one thread per query with its stack in global memory
unsigned ordered_grid(sb200_ctx* c, uint32_t len, uint64_t n_queries) {
    unsigned g = static_cast<unsigned>(c->sms) * blocks_per_sm("SB200_ORDERED_BLOCKS_PER_SM", len > 300 ? 1 : 4);  // (the stacks: 16 B x frames per thread)
    return std::max(1u, std::min(g, grid_for(n_queries)));
}
void launch_ordered(sb200_ctx* c, const SearchParams& P) {
    const size_t osmem = size_t(P.n_searches) * P.len * 4;
    if (osmem > 48 * 1024) throw Error("search scheme table does not fit shared memory (query too long)");
    const unsigned ogrid = ordered_grid(c, P.len, P.n_queries);
    SearchParams Q = P;
    Q.ostack_frames = ordered_stack_frames(P.len, c->idx.sigma);
    c->d_ostack.reserve(size_t(ogrid) * 256 * Q.ostack_frames * sizeof(uint4));
    Q.ostack = c->d_ostack.get<uint4>();
    with_sigma(c->idx.sigma, [&](auto S) {
        if (c->edit) fm_ordered_kernel<S(), true><<<ogrid, 256, osmem, c->stream>>>(Q);
        else fm_ordered_kernel<S(), false><<<ogrid, 256, osmem, c->stream>>>(Q);
        return 0;
    });
    launch_check(c);
}

// search_n after the plain search: a query with at most max_hits rows is complete; the others (they end early by
// definition) are walked again in the reference's recursion order and their cursors of the first pass are dropped.
// P: the parameters of the first pass, n_slots: its output slots.  false = the cursor buffer is too small (cursor_cap was
// raised, the caller starts over).
bool refine_max_hits(sb200_ctx* c, const SearchParams& P, uint64_t n_slots) {
    const uint32_t nq = P.n_queries;
    c->d_rows.reserve((size_t(nq) + 4) * sizeof(unsigned long long));
    c->d_redo.reserve(size_t(nq) * sizeof(uint32_t));
    unsigned long long* rows = c->d_rows.get<unsigned long long>();
    unsigned long long* tally = rows + nq;  // queries to redo, bound on their cursors, cursors dropped
    CUDA_TRY(cudaMemsetAsync(rows, 0, (size_t(nq) + 4) * sizeof(unsigned long long), c->stream));
    if (n_slots) cursor_rows_kernel<<<grid_for(n_slots), 256, 0, c->stream>>>(P.out, n_slots, rows);
    redo_list_kernel<<<grid_for(nq), 256, 0, c->stream>>>(rows, nq, c->max_hits, c->d_redo.get<uint32_t>(), tally);
    launch_check(c);
    read_back_words(c, tally, 2, CT_COUNT);
    const uint64_t n_redo = c->h_counters[CT_COUNT], bound = c->h_counters[CT_COUNT + 1];
    if (n_redo == 0) return true;
    const uint64_t need = n_slots + bound + uint64_t(ordered_grid(c, P.len, n_redo)) * 256 * kEmitChunk;
    if (need >= 0xfffffffeull) throw Error("more than 2^32 cursors in one call; split the batch");
    if (need > c->cursor_cap) {
        c->cursor_cap = need + need / 8;
        return false;
    }
    drop_cursors_kernel<<<grid_for(n_slots), 256, 0, c->stream>>>(P.out, n_slots, rows, c->max_hits, tally);
    CUDA_TRY(cudaMemsetAsync(P.counters + CT_NEXT_QUERY, 0, sizeof(unsigned long long), c->stream));
    SearchParams Q = P;
    Q.max_hits = c->max_hits;
    Q.redo = c->d_redo.get<uint32_t>();
    Q.n_queries = static_cast<uint32_t>(n_redo);
    Q.items = nullptr, Q.item_tags = nullptr;
    Q.qgram = c->idx.qgram_q && c->idx.qgram_q < P.len ? c->idx.qgram.get<uint4>() : nullptr;  // (not a table that covers the whole query)
    Q.qgram_q = Q.qgram ? c->idx.qgram_q : 0;
    launch_ordered(c, Q);
    read_back_words(c, tally + 2, 1, CT_COUNT);
    const uint64_t dropped = c->h_counters[CT_COUNT];
    read_back_words(c, c->d_counters.p, CT_COUNT);
    if (c->h_counters[CT_OVERFLOW]) throw Error("internal error: search stack overflow");
    if (c->h_counters[CT_OUT_SLOTS] > c->cursor_cap) throw Error("internal error: cursor buffer of the ordered walk too small");
    c->h_counters[CT_CURSORS] -= dropped;
    return true;
}

void launch_search(sb200_ctx* c, const SearchParams& P) {
    // shared memory: scheme table + one staged packed query per thread
    size_t smem = (size_t(P.n_searches) * P.len + (size_t(P.n_searches) * P.len * kRunE + 3) / 4 + size_t(packed_words(P.len)) * 256) * 4;
    if (smem > 100 * 1024) throw Error("search scheme table and staged queries do not fit shared memory (query too long)");
    unsigned grid = static_cast<unsigned>(c->sms) * blocks_per_sm("SB200_BLOCKS_PER_SM", 4);
    unsigned need = grid_for((uint64_t(P.n_queries) + kQueryBatch - 1) / kQueryBatch);
    if (need < grid) grid = std::max(1u, need);
    CUDA_TRY(cudaEventRecord(c->ev[8], c->stream));
    if (P.max_hits) {  // search_n by the ordered walk alone (SB200_ORDERED_ONLY=1)
        launch_ordered(c, P);
        CUDA_TRY(cudaEventRecord(c->ev[9], c->stream));
        CUDA_TRY(cudaEventRecord(c->ev[10], c->stream));
        return;
    }
    if (P.items) {  // root frames of all (query, search) in one pass, then the warp-synchronous walk over the live ones
        fm_roots_kernel<<<grid_for(P.n_queries), 256, 0, c->stream>>>(P);
        launch_check(c);
        const size_t ismem = size_t(P.n_searches) * P.len * 4;
        if (ismem > 48 * 1024) throw Error("search scheme table does not fit shared memory (query too long)");
        unsigned igrid = static_cast<unsigned>(c->sms) * blocks_per_sm("SB200_BLOCKS_PER_SM", SB200_FM_ITEMS_BLOCKS);
        igrid = std::max(1u, std::min(igrid, grid_for(P.n_queries)));
        with_sigma(c->idx.sigma, [&](auto S) {
            with_stack(c->kmax, [&](auto STACK) {
                if (c->edit) fm_items_kernel<S(), true, STACK()><<<igrid, 256, ismem, c->stream>>>(P);
                else fm_items_kernel<S(), false, STACK()><<<igrid, 256, ismem, c->stream>>>(P);
            });
            return 0;
        });
        launch_check(c);
    } else {
        with_sigma(c->idx.sigma, [&](auto S) {
            with_stack(c->kmax, [&](auto STACK) {
                auto go = [&](auto kern) {
                    if (smem > 48 * 1024) CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
                    kern<<<grid, 256, smem, c->stream>>>(P);
                };
                if (c->edit) go(fm_kernel<S(), true, STACK()>);
                else go(fm_kernel<S(), false, STACK()>);
            });
            return 0;
        });
        launch_check(c);
    }
    CUDA_TRY(cudaEventRecord(c->ev[9], c->stream));
    const bool use_pool = !(std::getenv("SB200_TEXT_POOL") && std::atoi(std::getenv("SB200_TEXT_POOL")) == 0);
    if (P.sa32 && use_pool) {  // in-text verification, one frame pool per warp
        with_stack(c->kmax, [&](auto STACK) {
            const size_t tables = (size_t(P.n_searches) * P.len * 4 + run_table_bytes(P.n_searches * P.len) + 7) & ~size_t{7};
            // warps per block: as many resident warps per SM as shared memory (227 KB, 1 KB reserved per block) and
            // registers (48 per thread: 42 warps) allow; measured on the headline workload: 36 warps 8.4 ms, 30 warps
            // 9.2 ms, 24 warps 10.2 ms
            unsigned threads = 0, per_sm = 1;
            size_t psmem = 0;
            for (unsigned w = kPoolThreads / 32; w >= 4; --w) {
                const size_t bytes = tables + w * size_t(pool_bytes(P.len));
                const unsigned blocks = static_cast<unsigned>(std::min<size_t>((227 * 1024) / (bytes + 1024), 42 / w));
                if (blocks * w > per_sm * (threads / 32)) {
                    threads = w * 32;
                    per_sm = blocks;
                    psmem = bytes;
                }
            }
            if (const char* e = std::getenv("SB200_POOL_THREADS")) {
                threads = std::min<unsigned>(kPoolThreads, std::max(32, std::atoi(e))) & ~31u;
                psmem = tables + (threads / 32) * size_t(pool_bytes(P.len));
                per_sm = static_cast<unsigned>(std::max<size_t>(1, std::min<size_t>(42 / (threads / 32), (227 * 1024) / (psmem + 1024))));
            }
            if (threads == 0 || psmem > 226 * 1024) throw Error("search scheme table and frame pools do not fit shared memory (query too long)");
            unsigned tgrid = static_cast<unsigned>(c->sms) * blocks_per_sm("SB200_POOL_BLOCKS_PER_SM", per_sm);
            auto go = [&](auto kern) {
                CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(psmem)));
                c->d_spill.reserve(size_t(tgrid) * (threads / 32) * 2 * kSpillCap * sizeof(uint4));
                kern<<<tgrid, threads, psmem, c->stream>>>(P, 2 * (c->kmax + 1), blocks_per_sm("SB200_RUN_ROUNDS", kRunRounds),
                                                           c->d_spill.get<uint4>());
            };
            if (c->edit) go(text_pool_kernel<true, STACK()>);
            else go(text_pool_kernel<false, STACK()>);
        });
        launch_check(c);
    } else if (P.sa32) {  // in-text verification of the seeds (reads the seed count from device memory: no host sync)
        unsigned tgrid = static_cast<unsigned>(c->sms) * blocks_per_sm("SB200_TEXT_BLOCKS_PER_SM", 6);
        with_stack(c->kmax, [&](auto STACK) {
            auto go = [&](auto kern) {
                if (smem > 48 * 1024) CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
                kern<<<tgrid, 256, smem, c->stream>>>(P);
            };
            if (c->edit) go(text_kernel<true, STACK()>);
            else go(text_kernel<false, STACK()>);
        });
        launch_check(c);
    }
    CUDA_TRY(cudaEventRecord(c->ev[10], c->stream));
}

// kernel 2 on device-resident queries; cursors stay in c->d_cursors
// d_queries: the queries (both strands), or with from_reads the n_queries / 2 reads (the reverse complements are made while packing)
void search_only(sb200_ctx* c, const uint8_t* d_queries, uint64_t n_queries, uint32_t len, bool for_locate = false, bool from_reads = false) {
    auto& ix = c->idx;
    if (!ix.loaded) throw Error("no index loaded");
    if (!c->have_scheme) throw Error("no search scheme set");
    if (len != c->qlen) throw Error("query length " + std::to_string(len) + " does not match the expanded search scheme (" +
                                    std::to_string(c->qlen) + ")");
    if (n_queries == 0) throw Error("query file was empty - abort");
    if (n_queries * uint64_t(c->n_searches) >= (1ull << 32)) throw Error("too many (query, search) pairs for one call; split the batch");

    if (c->cursor_cap < n_queries * 16) c->cursor_cap = std::max<uint64_t>(1 << 20, n_queries * 16);
    if (c->seed_cap < n_queries * 4) c->seed_cap = std::max<uint64_t>(1 << 20, n_queries * 4);
    c->d_counters.reserve(CT_COUNT * sizeof(unsigned long long));
    CUDA_TRY(cudaEventRecord(c->ev[0], c->stream));
    const uint32_t W = packed_words(len);
    c->d_packed.reserve(n_queries * W * 4);
    CUDA_TRY(cudaMemsetAsync(c->d_counters.p, 0, CT_COUNT * sizeof(unsigned long long), c->stream));
    if (from_reads)
        pack_reads_kernel<<<grid_for(n_queries * W), 256, 0, c->stream>>>(d_queries, n_queries, len, ix.sigma, c->d_packed.get<uint32_t>(),
                                                                          c->d_counters.get<unsigned long long>());
    else
        pack_queries_kernel<<<grid_for(n_queries * W), 256, 0, c->stream>>>(d_queries, n_queries, len, ix.sigma, c->d_packed.get<uint32_t>(),
                                                                            c->d_counters.get<unsigned long long>());
    launch_check(c);
    uint64_t n_cursors = 0;
    while (true) {
        c->d_cursors.reserve((c->cursor_cap + 1) * sizeof(uint4));
        if (ix.text_mode) c->d_seeds.reserve((c->seed_cap + 1) * sizeof(uint4));
        CUDA_TRY(cudaMemsetAsync(c->d_counters.p, 0, CT_BAD_QUERY * sizeof(unsigned long long), c->stream));  // keeps CT_BAD_QUERY
        CUDA_TRY(cudaMemsetAsync(c->d_counters.get<unsigned long long>() + CT_NODES_TEXT, 0, (CT_NEXT_ITEM + 1 - CT_NODES_TEXT) * sizeof(unsigned long long),
                                 c->stream));
        SearchParams P{};
        P.bwt = ix.bwt();
        P.bwtRev = ix.rev();
        for (int i = 0; i < 8; ++i) P.C[i] = ix.C[i];
        P.n_rows = static_cast<uint32_t>(ix.n_rows);
        P.packed = c->d_packed.get<uint32_t>();
        P.seeds = c->d_seeds.get<uint4>();
        P.seed_cap = static_cast<uint32_t>(std::min<uint64_t>(c->seed_cap, 0xfffffffeull));
        P.n_queries = static_cast<uint32_t>(n_queries);
        P.len = len;
        P.n_searches = c->n_searches;
        P.steps = c->d_steps.get<uint32_t>();
        P.runs = c->d_runs.get<uint8_t>();
        P.out = c->d_cursors.get<uint4>();
        P.out_cap = static_cast<uint32_t>(std::min<uint64_t>(c->cursor_cap, 0xfffffffeull));
        P.counters = c->d_counters.get<unsigned long long>();
        P.qgram = ix.qgram_q ? ix.qgram.get<uint4>() : nullptr;
        P.qgram_q = ix.qgram_q;
        // item-based walk (fm_roots_kernel + fm_items_kernel); SB200_FM_ITEMS=0 keeps the query-owning fm_kernel
        const bool use_items = !(std::getenv("SB200_FM_ITEMS") && std::atoi(std::getenv("SB200_FM_ITEMS")) == 0);
        if (use_items && c->n_searches <= 255) {  // (search index and slot count share a tag word)
            if (P.qgram_q >= len) P.qgram = nullptr, P.qgram_q = 0;  // (a table that covers the whole query: plain walk)
            const uint64_t total = n_queries * uint64_t(c->n_searches);
            c->d_items.reserve(total * sizeof(uint4));
            c->d_item_tags.reserve(total * sizeof(uint2));
            P.items = c->d_items.get<uint4>();
            P.item_tags = c->d_item_tags.get<uint2>();
        }
        P.sa32 = ix.text_mode ? ix.sa32.get<uint32_t>() : nullptr;
        P.isa32 = ix.text_mode ? ix.isa32.get<uint32_t>() : nullptr;
        P.text4 = ix.text_mode ? ix.text4.get<uint32_t>() : nullptr;
        // cursors that go straight to the locate step carry the text position of a verified occurrence instead of its row
        P.textpos_out = (for_locate && ix.text_mode && !(std::getenv("SB200_TEXTPOS") && std::atoi(std::getenv("SB200_TEXTPOS")) == 0)) ? 1u : 0u;
        if (const char* dbg = std::getenv("SB200_DEBUG")) P.debug_flags = static_cast<uint32_t>(std::atoi(dbg));
        // search_n: the plain search first, then the queries above the limit again in recursion order (refine_max_hits);
        // SB200_ORDERED_ONLY=1 walks every query in order instead
        const bool ordered_only = c->max_hits && std::getenv("SB200_ORDERED_ONLY") && std::atoi(std::getenv("SB200_ORDERED_ONLY")) != 0;
        if (ordered_only) {
            P.max_hits = c->max_hits;
            P.items = nullptr, P.item_tags = nullptr;
            if (P.qgram_q >= len) P.qgram = nullptr, P.qgram_q = 0;  // (a table that covers the whole query: plain walk)
        }
        launch_search(c, P);
        read_back_words(c, c->d_counters.p, CT_COUNT);
        if (std::getenv("SB200_DEBUG"))
            fprintf(stderr, "[sb200 debug] max stack depth %llu overflow %llu seeds %llu\n", c->h_counters[CT_MAX_SP], c->h_counters[CT_OVERFLOW],
                    c->h_counters[CT_SEEDS]);
        if (c->h_counters[CT_BAD_QUERY]) {
            uint64_t off = c->h_counters[CT_BAD_QUERY] - 1;
            throw Error("query has invalid character at offset " + std::to_string(off % len) + " of query " + std::to_string(off / len));
        }
        if (c->h_counters[CT_OVERFLOW]) throw Error("internal error: search stack overflow");
        n_cursors = c->h_counters[CT_OUT_SLOTS];
        uint64_t n_seed_slots = c->h_counters[CT_SEED_SLOTS];
        if (n_cursors >= 0xfffffffeull || n_seed_slots >= 0xfffffffeull) throw Error("more than 2^32 cursors in one call; split the batch");
        bool fits = true;
        if (n_seed_slots > c->seed_cap) {  // the text kernel saw a truncated seed list: rerun with a buffer that fits
            c->seed_cap = n_seed_slots + n_seed_slots / 4;
            fits = false;
        }
        if (n_cursors > c->cursor_cap) {
            c->cursor_cap = n_cursors + n_cursors / 4;
            fits = false;
        }
        if (fits && c->max_hits && !ordered_only) {
            if (!refine_max_hits(c, P, n_cursors)) continue;
            n_cursors = c->h_counters[CT_OUT_SLOTS];
        }
        if (fits) break;
    }
    CUDA_TRY(cudaEventRecord(c->ev[1], c->stream));
    CUDA_TRY(cudaEventSynchronize(c->ev[1]));
    CUDA_TRY(cudaEventElapsedTime(&c->ct.ms_search, c->ev[0], c->ev[1]));
    CUDA_TRY(cudaEventElapsedTime(&c->ms_fm, c->ev[8], c->ev[9]));
    CUDA_TRY(cudaEventElapsedTime(&c->ms_text, c->ev[9], c->ev[10]));
    c->ct.ms_locate = c->ct.ms_sort = 0;
    c->ct.nodes += c->h_counters[CT_NODES];
    c->nodes_text += c->h_counters[CT_NODES_TEXT];
    c->ct.rank_ops += 2 * (c->h_counters[CT_NODES] - c->h_counters[CT_NODES_TEXT]);
    c->ct.cursors += c->h_counters[CT_CURSORS];
    c->last_cursors = n_cursors;  // reserved output slots; unused ones are empty entries (qid 0xffffffff, len 0)
    c->last_real_cursors = c->h_counters[CT_CURSORS];
    c->last_hits = 0;
}

static_assert(kCursorTextPosFlag == kCursorTextPos, "search.cuh and locate.cuh agree on the flag");

__global__ void mirror_u32_kernel(const uint32_t* src, unsigned long long* dst, int n) {
    if (threadIdx.x < n) dst[threadIdx.x] = src[threadIdx.x];
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

// Locate + sort through per-query buckets (locate.cuh) for the cursors of a search over queries 0 .. n_queries-1.
// Returns false when some query has more hits than a block sorts (the caller then takes the global radix sort).
bool locate_bucketed(sb200_ctx* c, uint64_t n_cursors, uint64_t n_queries) {
    auto& ix = c->idx;
    CUDA_TRY(cudaEventRecord(c->ev[1], c->stream));
    c->d_counters.reserve(CT_COUNT * sizeof(unsigned long long));
    CUDA_TRY(cudaMemsetAsync(c->d_counters.get<unsigned long long>() + CT_LF_STEPS, 0, sizeof(unsigned long long), c->stream));
    c->d_qpos.reserve((n_queries + 1) * 4);
    c->d_lc.reserve(LC_COUNT * 4);
    CUDA_TRY(cudaMemsetAsync(c->d_qpos.p, 0, (n_queries + 1) * 4, c->stream));
    CUDA_TRY(cudaMemsetAsync(c->d_lc.p, 0, LC_COUNT * 4, c->stream));
    uint64_t total_rows = 0;
    if (n_cursors > 0) {
        CUDA_TRY(cudaMemsetAsync(c->d_counters.get<unsigned long long>() + CT_TOTAL_ROWS, 0, sizeof(unsigned long long), c->stream));
        hit_count_kernel<<<grid_for(n_cursors), 256, 0, c->stream>>>(c->d_cursors.get<uint4>(), static_cast<uint32_t>(n_cursors),
                                                                    c->d_qpos.get<uint32_t>(),
                                                                    c->d_counters.get<unsigned long long>() + CT_TOTAL_ROWS);
        launch_check(c);
        size_t tmp_bytes = 0;
        CUDA_TRY(cub::DeviceScan::ExclusiveSum(nullptr, tmp_bytes, c->d_qpos.get<uint32_t>(), c->d_qpos.get<uint32_t>(), n_queries + 1, c->stream));
        c->d_tmp.reserve(tmp_bytes);
        CUDA_TRY(cub::DeviceScan::ExclusiveSum(c->d_tmp.p, tmp_bytes, c->d_qpos.get<uint32_t>(), c->d_qpos.get<uint32_t>(), n_queries + 1, c->stream));
        c->ct.kernel_launches += 2;
        mirror_u32_kernel<<<1, 32, 0, c->stream>>>(c->d_qpos.get<uint32_t>() + n_queries, c->h_counters_dev + CT_COUNT, 1);
        launch_check(c);
        read_back_words(c, c->d_counters.get<unsigned long long>() + CT_TOTAL_ROWS, 1, CT_COUNT + 1);
        total_rows = c->h_counters[CT_COUNT];
        // (the scan is u32: n_cursors + rows beyond the first of every cursor bounds the true number of hits)
        if (n_cursors + c->h_counters[CT_COUNT + 1] >= (1ull << 32)) throw Error("more than 2^32 hits in one call; split the batch");
    }
    c->fused_shift = 0;
    c->sorted_keys = 0;
    c->d_keys[0].reserve(std::max<uint64_t>(1, total_rows) * 8);
    c->d_qids[0].reserve(std::max<uint64_t>(1, total_rows) * 4);
    bool ok = true;
    if (total_rows > 0) {
        BucketParams B{};
        B.index = locate_index(c);
        B.cursors = c->d_cursors.get<uint4>();
        B.n_cursors = static_cast<uint32_t>(n_cursors);
        B.n_queries = static_cast<uint32_t>(n_queries);
        B.qpos = c->d_qpos.get<uint32_t>();
        B.keys = c->d_keys[0].get<uint64_t>();
        B.qids = c->d_qids[0].get<uint32_t>();
        B.task_cap = static_cast<uint32_t>(std::min<uint64_t>(total_rows / kInlineRows + 1024, 0xfffffff0ull));
        B.big_cap = static_cast<uint32_t>(total_rows / kWarpSeg + 16);
        c->d_tasks.reserve(uint64_t(B.task_cap) * sizeof(uint4));
        c->d_bigsegs.reserve(uint64_t(B.big_cap) * 4);
        B.tasks = c->d_tasks.get<uint4>();
        B.big_segs = c->d_bigsegs.get<uint32_t>();
        B.lc = c->d_lc.get<unsigned int>();
        B.counters = c->d_counters.get<unsigned long long>();
        const unsigned wide = static_cast<unsigned>(c->sms) * 4;
        with_sigma(ix.sigma, [&](auto S) {
            locate_scatter_kernel<S()><<<grid_for(n_cursors), 256, 0, c->stream>>>(B);
            launch_check(c);
            locate_tasks_kernel<S()><<<wide, 256, 0, c->stream>>>(B);
            return 0;
        });
        launch_check(c);
        CUDA_TRY(cudaEventRecord(c->ev[2], c->stream));
        segment_sort_kernel<<<grid_for((n_queries + 31) / 32 * 32), 256, 0, c->stream>>>(B);  // one lane per query
        launch_check(c);
        segment_sort_big_kernel<<<wide, 256, 0, c->stream>>>(B);
        launch_check(c);
        CUDA_TRY(cudaEventRecord(c->ev[3], c->stream));
        mirror_u32_kernel<<<1, 32, 0, c->stream>>>(c->d_lc.get<uint32_t>(), c->h_counters_dev + CT_COUNT, LC_COUNT);
        launch_check(c);
        read_back_words(c, c->d_counters.p, CT_COUNT);
        ok = c->h_counters[CT_COUNT + LC_HUGE] == 0 && c->h_counters[CT_COUNT + LC_TASKS] <= B.task_cap;
    } else {
        CUDA_TRY(cudaEventRecord(c->ev[2], c->stream));
        CUDA_TRY(cudaEventRecord(c->ev[3], c->stream));
        read_back_words(c, c->d_counters.p, CT_COUNT);
    }
    if (!ok) return false;
    c->ct.lf_steps += c->h_counters[CT_LF_STEPS];
    c->ct.hits += total_rows;
    c->last_hits = total_rows;
    CUDA_TRY(cudaEventElapsedTime(&c->ct.ms_locate, c->ev[1], c->ev[2]));
    CUDA_TRY(cudaEventElapsedTime(&c->ct.ms_sort, c->ev[2], c->ev[3]));
    return true;
}

// kernel 3 over the n_cursors cursors in c->d_cursors (room for one extra slot), then sort by
// (qid, seq/pos, e).  Sorted hits end in d_keys[sorted_keys] (fused keys) or d_keys[0] / d_qids[0].
void locate_only(sb200_ctx* c, uint64_t n_cursors, uint64_t n_queries_hint) {
    auto& ix = c->idx;
    // cursors of our own search (query ids 0 .. n_queries_hint-1): per-query buckets; SB200_BUCKET_SORT=0 or a query with
    // thousands of hits: global radix sort
    const bool bucketed = n_queries_hint > 0 && n_queries_hint < 0xffffffffull && n_cursors < 0xffffffffull &&
                          !(std::getenv("SB200_BUCKET_SORT") && std::atoi(std::getenv("SB200_BUCKET_SORT")) == 0);
    if (bucketed && locate_bucketed(c, n_cursors, n_queries_hint)) return;
    CUDA_TRY(cudaEventRecord(c->ev[1], c->stream));
    c->d_counters.reserve(CT_COUNT * sizeof(unsigned long long));
    CUDA_TRY(cudaMemsetAsync(c->d_counters.get<unsigned long long>() + CT_LF_STEPS, 0, sizeof(unsigned long long), c->stream));
    uint64_t total_rows = 0;
    if (n_cursors > 0) {
        c->d_offsets.reserve((n_cursors + 1) * 8);
        auto lens = cub::TransformInputIterator<uint64_t, CursorLen, const uint4*>(c->d_cursors.get<uint4>(), CursorLen{});
        size_t tmp_bytes = 0;
        CUDA_TRY(cub::DeviceScan::ExclusiveSum(nullptr, tmp_bytes, lens, c->d_offsets.get<uint64_t>(), n_cursors + 1, c->stream));
        c->d_tmp.reserve(tmp_bytes);
        // n_cursors + 1 items are scanned: zero the slot behind the last cursor
        CUDA_TRY(cudaMemsetAsync(c->d_cursors.get<uint4>() + n_cursors, 0, sizeof(uint4), c->stream));
        CUDA_TRY(cub::DeviceScan::ExclusiveSum(c->d_tmp.p, tmp_bytes, lens, c->d_offsets.get<uint64_t>(), n_cursors + 1, c->stream));
        c->ct.kernel_launches += 1;
        read_back_words(c, c->d_offsets.get<uint64_t>() + n_cursors, 1, CT_COUNT);
        total_rows = c->h_counters[CT_COUNT];
    }
    if (total_rows >= (1ull << 32)) throw Error("more than 2^32 hits in one call; split the batch");
    // one 64-bit key (query id above value and errors) when it fits: a single keys-only radix sort
    const int key_bits = static_cast<int>(ix.key_bits) + 4;
    const int qid_bits = std::max(1, static_cast<int>(bits_of(n_queries_hint ? n_queries_hint - 1 : 0xffffffffull)));
    // (measured at equal pass counts, 16 M hits: two pair sorts 1.19 ms, one 64-bit keys-only sort 1.38 ms — so only
    // when it saves a radix pass)
    const bool fewer_passes = (key_bits + qid_bits + 7) / 8 < (key_bits + 7) / 8 + (qid_bits + 7) / 8;
    const char* force = std::getenv("SB200_FUSED_SORT");
    const bool fused = key_bits + qid_bits <= 64 && (force ? std::atoi(force) != 0 : fewer_passes);
    c->fused_shift = fused ? static_cast<uint32_t>(key_bits) : 0u;
    c->sorted_keys = 0;
    for (int i = 0; i < 2; ++i) {
        c->d_keys[i].reserve(std::max<uint64_t>(1, total_rows) * 8);
        if (!fused) c->d_qids[i].reserve(std::max<uint64_t>(1, total_rows) * 4);
    }
    if (total_rows > 0) {
        LocateParams L{};
        L.index = locate_index(c);
        L.cursors = c->d_cursors.get<uint4>();
        L.offsets = c->d_offsets.get<uint64_t>();
        L.n_cursors = static_cast<uint32_t>(n_cursors);
        L.n_rows_total = total_rows;
        L.out_key = c->d_keys[0].get<uint64_t>();
        L.out_qid = fused ? nullptr : c->d_qids[0].get<uint32_t>();
        L.fused_shift = c->fused_shift;
        L.counters = c->d_counters.get<unsigned long long>();
        with_sigma(ix.sigma, [&](auto S) {
            locate_kernel<S()><<<grid_for(total_rows), 256, 0, c->stream>>>(L);
            return 0;
        });
        launch_check(c);
    }
    CUDA_TRY(cudaEventRecord(c->ev[2], c->stream));
    if (total_rows > 1 && fused) {
        size_t t1 = 0;
        CUDA_TRY(cub::DeviceRadixSort::SortKeys(nullptr, t1, c->d_keys[0].get<uint64_t>(), c->d_keys[1].get<uint64_t>(), total_rows, 0,
                                                key_bits + qid_bits, c->stream));
        c->d_tmp.reserve(t1);
        CUDA_TRY(cub::DeviceRadixSort::SortKeys(c->d_tmp.p, t1, c->d_keys[0].get<uint64_t>(), c->d_keys[1].get<uint64_t>(), total_rows, 0,
                                                key_bits + qid_bits, c->stream));
        c->sorted_keys = 1;
        c->ct.kernel_launches += (key_bits + qid_bits + 7) / 8 + 2;
    } else if (total_rows > 1) {
        size_t t1 = 0, t2 = 0;
        CUDA_TRY(cub::DeviceRadixSort::SortPairs(nullptr, t1, c->d_keys[0].get<uint64_t>(), c->d_keys[1].get<uint64_t>(),
                                                 c->d_qids[0].get<uint32_t>(), c->d_qids[1].get<uint32_t>(), total_rows, 0, key_bits,
                                                 c->stream));
        CUDA_TRY(cub::DeviceRadixSort::SortPairs(nullptr, t2, c->d_qids[1].get<uint32_t>(), c->d_qids[0].get<uint32_t>(),
                                                 c->d_keys[1].get<uint64_t>(), c->d_keys[0].get<uint64_t>(), total_rows, 0, qid_bits,
                                                 c->stream));
        c->d_tmp.reserve(std::max(t1, t2));
        CUDA_TRY(cub::DeviceRadixSort::SortPairs(c->d_tmp.p, t1, c->d_keys[0].get<uint64_t>(), c->d_keys[1].get<uint64_t>(),
                                                 c->d_qids[0].get<uint32_t>(), c->d_qids[1].get<uint32_t>(), total_rows, 0, key_bits,
                                                 c->stream));
        CUDA_TRY(cub::DeviceRadixSort::SortPairs(c->d_tmp.p, t2, c->d_qids[1].get<uint32_t>(), c->d_qids[0].get<uint32_t>(),
                                                 c->d_keys[1].get<uint64_t>(), c->d_keys[0].get<uint64_t>(), total_rows, 0, qid_bits,
                                                 c->stream));
        c->ct.kernel_launches += 2 * ((key_bits + 7) / 8 + 1) + 2 * ((qid_bits + 7) / 8 + 1);
    }
    CUDA_TRY(cudaEventRecord(c->ev[3], c->stream));
    read_back_words(c, c->d_counters.p, CT_COUNT);
    c->ct.lf_steps += c->h_counters[CT_LF_STEPS];
    c->ct.hits += total_rows;
    c->last_hits = total_rows;
    CUDA_TRY(cudaEventElapsedTime(&c->ct.ms_locate, c->ev[1], c->ev[2]));
    CUDA_TRY(cudaEventElapsedTime(&c->ct.ms_sort, c->ev[2], c->ev[3]));
}

void run_pipeline(sb200_ctx* c, const uint8_t* d_queries, uint64_t n_queries, uint32_t len, bool do_locate, bool from_reads = false) {
    search_only(c, d_queries, n_queries, len, do_locate, from_reads);
    if (do_locate) locate_only(c, c->last_cursors, n_queries);
}

void fetch_hits(sb200_ctx* c, sb200_hit** hits, uint64_t* n_hits) {
    uint64_t n = c->last_hits;
    auto& ix = c->idx;
    sb200_hit* out = static_cast<sb200_hit*>(g_pinned.alloc(std::max<uint64_t>(1, n) * sizeof(sb200_hit)));
    if (n) {
        // expand to the reference tuple on the device, then one pinned copy
        c->d_scratch.reserve(n * sizeof(sb200_hit));
        expand_hits_kernel<<<grid_for(n), 256, 0, c->stream>>>(c->d_keys[c->sorted_keys].get<uint64_t>(), c->d_qids[0].get<uint32_t>(), n,
                                                               static_cast<uint32_t>(ix.bits_for_position), 0, c->fused_shift,
                                                               c->d_scratch.get<uint64_t>());
        launch_check(c);
        CUDA_TRY(cudaEventRecord(c->ev[4], c->stream));
        CUDA_TRY(cudaMemcpyAsync(out, c->d_scratch.p, n * sizeof(sb200_hit), cudaMemcpyDeviceToHost, c->stream));
        CUDA_TRY(cudaEventRecord(c->ev[5], c->stream));
        CUDA_TRY(cudaStreamSynchronize(c->stream));
        CUDA_TRY(cudaEventElapsedTime(&c->ct.ms_d2h, c->ev[4], c->ev[5]));
    }
    *hits = out;
    *n_hits = n;
}

const uint8_t* stage_queries(sb200_ctx* c, const uint8_t* queries, uint64_t n_queries, uint32_t len) {
    if (!c->idx.loaded) throw Error("no index loaded");
    if (!queries || n_queries == 0) throw Error("query file was empty - abort");
    c->d_queries.reserve(n_queries * len);
    CUDA_TRY(cudaEventRecord(c->ev[6], c->stream));
    CUDA_TRY(cudaMemcpyAsync(c->d_queries.p, queries, n_queries * len, cudaMemcpyHostToDevice, c->stream));
    CUDA_TRY(cudaEventRecord(c->ev[7], c->stream));
    return c->d_queries.get<uint8_t>();
}

// search + locate from host buffers, pipelined: the batch is cut into chunks of reads; the host->device copy of
// chunk i+1 and the device->host copy of the hits of chunk i-1 run on their own streams while chunk i computes.
// Chunks are contiguous query ranges, so concatenating their sorted hit lists keeps the global order.
//   make_rc : the host buffer holds only the reads; the reverse complements are made on the device
//             (queries[2i] = read i, queries[2i+1] = its reverse complement, search.cpp:121-123)
//   compact : hits are returned as sb200_hit32 (16 bytes) instead of the reference's 32-byte tuple
void search_host_pipelined(sb200_ctx* c, const uint8_t* src, uint64_t n_items, uint32_t len, bool make_rc, bool compact, void** hits,
                           uint64_t* n_hits) {
    auto& ix = c->idx;
    if (!ix.loaded) throw Error("no index loaded");
    if (!src || n_items == 0) throw Error("query file was empty - abort");
    const uint64_t per_item = make_rc ? 2 : 1;  // queries per host item
    const uint64_t n_queries = n_items * per_item;
    const size_t hit_bytes = compact ? sizeof(sb200_hit32) : sizeof(sb200_hit);
    // Chunk boundaries (in queries).  Every chunk costs ~0.9 ms of kernel drain (the longest single seed), so few
    // chunks: a short first one (its copy-in cannot be hidden), a short last one (its copy-out cannot be hidden),
    // and the rest in pieces of at most `chunk` queries whose copies hide behind the neighbours' kernels.
    uint64_t chunk = 2000000, edge_div = 5;
    if (const char* e = std::getenv("SB200_CHUNK")) chunk = std::max<uint64_t>(2, std::strtoull(e, nullptr, 10));
    if (const char* e = std::getenv("SB200_EDGE_DIV")) edge_div = std::max<uint64_t>(2, std::strtoull(e, nullptr, 10));
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
    uint64_t max_chunk = 0;
    for (uint64_t k = 0; k < n_chunks; ++k) max_chunk = std::max(max_chunk, bounds[k + 1] - bounds[k]);
    for (int i = 0; i < 2; ++i) c->d_qchunk[i].reserve(max_chunk * len);
    // wait until earlier work on the caller's stream is done before the copy streams touch the buffers
    CUDA_TRY(cudaStreamSynchronize(c->stream));
    auto copy_in = [&](uint64_t k) {
        const uint64_t q0 = bounds[k], n = bounds[k + 1] - q0;
        const int b = static_cast<int>(k & 1);
        if (k >= 2) CUDA_TRY(cudaStreamWaitEvent(c->s_in, c->ev_free_q[b], 0));  // chunk k-2 no longer reads this buffer
        CUDA_TRY(cudaMemcpyAsync(c->d_qchunk[b].p, src + (q0 / per_item) * len, (n / per_item) * len, cudaMemcpyHostToDevice, c->s_in));
        CUDA_TRY(cudaEventRecord(c->ev_in[b], c->s_in));
    };
    uint8_t* out = nullptr;
    uint64_t out_cap = 0, total = 0;
    auto t0 = std::chrono::steady_clock::now();
    copy_in(0);
    float ms_search = 0, ms_locate = 0, ms_sort = 0;
    try {
        for (uint64_t k = 0; k < n_chunks; ++k) {
            const uint64_t q0 = bounds[k], n = bounds[k + 1] - q0;
            const int b = static_cast<int>(k & 1);
            if (k + 1 < n_chunks) copy_in(k + 1);
            CUDA_TRY(cudaStreamWaitEvent(c->stream, c->ev_in[b], 0));
            const uint8_t* dq = c->d_qchunk[b].get<uint8_t>();
            auto tc0 = std::chrono::steady_clock::now();
            run_pipeline(c, dq, n, len, true, make_rc);  // search + locate + sort of this chunk (make_rc: dq holds the reads only)
            if (std::getenv("SB200_DEBUG"))
                fprintf(stderr,
                        "[sb200 debug] chunk %llu: %llu queries, host wall %.3f ms, device search %.3f (fm %.3f text %.3f) locate %.3f sort %.3f ms\n",
                        (unsigned long long)k, (unsigned long long)n,
                        std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - tc0).count(), c->ct.ms_search, c->ms_fm,
                        c->ms_text, c->ct.ms_locate, c->ct.ms_sort);
            CUDA_TRY(cudaEventRecord(c->ev_free_q[b], c->stream));
            ms_search += c->ct.ms_search;
            ms_locate += c->ct.ms_locate;
            ms_sort += c->ct.ms_sort;
            const uint64_t nh = c->last_hits;
            // output buffer: sized from the first chunk, grown (rarely) when the estimate was too small
            if (total + nh > out_cap) {
                // first estimate: the hit density of the first chunk over the whole batch
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
            if (nh) {
                if (k >= 2) CUDA_TRY(cudaStreamWaitEvent(c->stream, c->ev_out[b], 0));  // hits of chunk k-2 have left this buffer
                c->d_hitchunk[b].reserve(nh * hit_bytes);
                if (compact)
                    compact_hits_kernel<<<grid_for(nh), 256, 0, c->stream>>>(c->d_keys[c->sorted_keys].get<uint64_t>(),
                                                                             c->d_qids[0].get<uint32_t>(), nh,
                                                                             static_cast<uint32_t>(ix.bits_for_position),
                                                                             static_cast<uint32_t>(q0), c->fused_shift,
                                                                             c->d_hitchunk[b].get<uint4>());
                else
                    expand_hits_kernel<<<grid_for(nh), 256, 0, c->stream>>>(c->d_keys[c->sorted_keys].get<uint64_t>(),
                                                                            c->d_qids[0].get<uint32_t>(), nh,
                                                                            static_cast<uint32_t>(ix.bits_for_position), q0, c->fused_shift,
                                                                            c->d_hitchunk[b].get<uint64_t>());
                launch_check(c);
                CUDA_TRY(cudaEventRecord(c->ev_expanded[b], c->stream));
                CUDA_TRY(cudaStreamWaitEvent(c->s_out, c->ev_expanded[b], 0));
                CUDA_TRY(cudaMemcpyAsync(out + total * hit_bytes, c->d_hitchunk[b].p, nh * hit_bytes, cudaMemcpyDeviceToHost, c->s_out));
            }
            CUDA_TRY(cudaEventRecord(c->ev_out[b], c->s_out));
            total += nh;
        }
        CUDA_TRY(cudaStreamSynchronize(c->s_out));
        CUDA_TRY(cudaStreamSynchronize(c->s_in));
    } catch (...) {
        cudaStreamSynchronize(c->s_out);
        cudaStreamSynchronize(c->s_in);
        if (out) g_pinned.free(out);
        throw;
    }
    if (!out) out = static_cast<uint8_t*>(g_pinned.alloc(hit_bytes));
    if (std::getenv("SB200_DEBUG"))
        fprintf(stderr, "[sb200 debug] host-buffer search total host wall %.3f ms\n",
                std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count());
    c->ct.ms_search = ms_search;
    c->ct.ms_locate = ms_locate;
    c->ct.ms_sort = ms_sort;
    c->ct.ms_h2d = c->ct.ms_d2h = 0;  // overlapped with the kernels
    *hits = out;
    *n_hits = total;
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
        for (int i = 0; i < 2; ++i) {
            CUDA_TRY(cudaEventCreateWithFlags(&c->ev_in[i], cudaEventDisableTiming));
            CUDA_TRY(cudaEventCreateWithFlags(&c->ev_free_q[i], cudaEventDisableTiming));
            CUDA_TRY(cudaEventCreateWithFlags(&c->ev_expanded[i], cudaEventDisableTiming));
            CUDA_TRY(cudaEventCreateWithFlags(&c->ev_out[i], cudaEventDisableTiming));
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
        c->idx.release();
        for (DevBuf* b : {&c->d_steps, &c->d_runs, &c->d_qpos, &c->d_tasks, &c->d_bigsegs, &c->d_lc, &c->d_items, &c->d_item_tags, &c->d_seeds, &c->d_spill, &c->d_packed, &c->d_queries, &c->d_cursors, &c->d_counters, &c->d_offsets, &c->d_keys[0], &c->d_keys[1],
                          &c->d_qids[0], &c->d_qids[1], &c->d_tmp, &c->d_scratch})
            b->release();
        for (auto& ev : c->ev) cudaEventDestroy(ev);
        for (int i = 0; i < 2; ++i) {
            c->d_qchunk[i].release();
            c->d_hitchunk[i].release();
            cudaEventDestroy(c->ev_in[i]);
            cudaEventDestroy(c->ev_free_q[i]);
            cudaEventDestroy(c->ev_expanded[i]);
            cudaEventDestroy(c->ev_out[i]);
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
            ix.ref_ssa = ix.ssa;
            ix.ssa = DevBuf{};
        } else {
            ix.ssa.release();
        }
        if (rate == 1) {
            CUDA_TRY(cudaStreamSynchronize(c->stream));
            ix.ssa = row_value;
            row_value = DevBuf{};
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
        ix.qgram = a;
        a = DevBuf{};
        b.release();
        ix.qgram_q = q;
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
        c->d_steps.reserve(steps.size() * 4);
        CUDA_TRY(cudaMemcpyAsync(c->d_steps.p, steps.data(), steps.size() * 4, cudaMemcpyHostToDevice, c->stream));
        std::vector<uint8_t> runs(run_table_bytes(static_cast<uint32_t>(steps.size())) + 4, 0);  // run lengths, then state flags
        build_runs(n_searches, len, steps.data(), runs.data());
        build_state_flags(n_searches, len, steps.data(), runs.data());
        c->d_runs.reserve(runs.size());
        CUDA_TRY(cudaMemcpyAsync(c->d_runs.p, runs.data(), runs.size(), cudaMemcpyHostToDevice, c->stream));
        CUDA_TRY(cudaStreamSynchronize(c->stream));
        c->n_searches = n_searches;
        c->qlen = len;
        c->kmax = kmax;
        c->edit = edit != 0;
        c->have_scheme = true;
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
        if (n_cursors) *n_cursors = c->last_real_cursors;
        if (n_hits) *n_hits = c->last_hits;
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
        search_host_pipelined(c, queries, n_queries, len, false, false, &p, n_hits);
        *hits = static_cast<sb200_hit*>(p);
    });
}

int sb200_search_reads(sb200_ctx* c, const uint8_t* reads, uint64_t n_reads, uint32_t len, int with_reverse, sb200_hit32** hits,
                       uint64_t* n_hits) {
    return guard([&] {
        use(c);
        if (c->idx.loaded && c->idx.bits_for_position > 32) throw Error("sequences too long for 32-bit positions: use sb200_search");
        void* p = nullptr;
        search_host_pipelined(c, reads, n_reads, len, with_reverse != 0, true, &p, n_hits);
        *hits = static_cast<sb200_hit32*>(p);
    });
}

int sb200_search_cursors(sb200_ctx* c, const uint8_t* queries, uint64_t n_queries, uint32_t len, sb200_cursor** cursors, uint64_t* n_cursors) {
    return guard([&] {
        use(c);
        const uint8_t* dq = stage_queries(c, queries, n_queries, len);
        run_pipeline(c, dq, n_queries, len, false);
        uint64_t n = c->last_cursors;
        std::vector<uint4> tmp(n);
        CUDA_TRY(cudaMemcpyAsync(tmp.data(), c->d_cursors.p, n * sizeof(uint4), cudaMemcpyDeviceToHost, c->stream));
        CUDA_TRY(cudaStreamSynchronize(c->stream));
        tmp.erase(std::remove_if(tmp.begin(), tmp.end(), [](uint4 const& a) { return a.x == kInvalidQid; }), tmp.end());
        if (tmp.size() != c->last_real_cursors) throw Error("internal error: cursor count mismatch");
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
            if (k.lb + k.len > ix.n_rows || k.errors > 15 || k.query_id > 0xffffffffull) throw Error("cursor out of range");
            tmp[i] = make_uint4(static_cast<uint32_t>(k.query_id), static_cast<uint32_t>(k.lb), static_cast<uint32_t>(k.len),
                                static_cast<uint32_t>(k.errors));
        }
        tmp[n_cursors] = make_uint4(0, 0, 0, 0);
        c->d_cursors.reserve((n_cursors + 1) * sizeof(uint4));
        CUDA_TRY(cudaMemcpyAsync(c->d_cursors.p, tmp.data(), (n_cursors + 1) * sizeof(uint4), cudaMemcpyHostToDevice, c->stream));
        CUDA_TRY(cudaStreamSynchronize(c->stream));
        c->last_cursors = n_cursors;
        locate_only(c, n_cursors, 0);
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
