// build.cuh — index construction on the GPU.
//
// Replaces `fmc::BiFMIndex<Sigma, InterleavedBitvector16>{ref, samplingRate, threads}`
// (/root/reference/src/sahara/index.cpp:87; upstream: concatenate with delimiters, libsais suffix
// array, BWT of the text and of the reversed text, occurrence tables, sampled suffix array —
// SURVEY.md §9.2).  B200 design: the suffix array of a text of up to 2^32 symbols is built by one
// radix sort of 64-bit keys holding the first 21 symbols of every suffix (3 bits each), followed by
// prefix-doubling rounds that only touch the suffixes still tied (for random DNA ~0.1 % after the first
// pass).  Everything stays in HBM; a 3.1 Gbp text needs about 80 GB at the peak.
#pragma once
#include <cub/cub.cuh>
#include "layout.cuh"
#include "locate.cuh"

namespace sb200 {

struct SeqMap {
    const uint64_t* start;  // n_seqs + 1 text offsets (each sequence is followed by one delimiter)
    uint32_t n_seqs;
};

__device__ __forceinline__ uint32_t find_seq(const SeqMap& m, uint64_t p) {
    uint32_t lo = 0, hi = m.n_seqs;  // start[lo] <= p < start[hi]
    while (hi - lo > 1) {
        uint32_t mid = lo + ((hi - lo) >> 1);
        if (m.start[mid] <= p) lo = mid;
        else hi = mid;
    }
    return lo;
}

// text = seq_0 $ seq_1 $ ... ; reverse: every sequence reversed in place
__global__ void make_text_kernel(const uint8_t* src, SeqMap m, uint64_t n, bool reverse, uint8_t* text) {
    uint64_t i = blockIdx.x * static_cast<uint64_t>(blockDim.x) + threadIdx.x;
    if (i >= n) return;
    uint32_t s = find_seq(m, i);
    uint64_t off = i - m.start[s];
    uint64_t len = m.start[s + 1] - m.start[s] - 1;
    uint8_t c = 0;
    if (off < len) {
        uint64_t base = m.start[s] - s;  // s delimiters precede
        c = reverse ? src[base + len - 1 - off] : src[base + off];
    }
    text[i] = c;
}

// flags texts that contain a rank outside 1..sigma-1
__global__ void check_text_kernel(const uint8_t* src, uint64_t n, uint32_t sigma, unsigned int* err) {
    uint64_t i = blockIdx.x * static_cast<uint64_t>(blockDim.x) + threadIdx.x;
    if (i >= n) return;
    uint8_t c = src[i];
    if (c == 0 || c >= sigma) atomicOr(err, 1u);
}

constexpr int kKeySyms = 21;

__global__ void make_keys_kernel(const uint8_t* text, uint64_t n, uint64_t* keys, uint32_t* vals) {
    uint64_t i = blockIdx.x * static_cast<uint64_t>(blockDim.x) + threadIdx.x;
    if (i >= n) return;
    uint64_t k = 0;
#pragma unroll
    for (int t = 0; t < kKeySyms; ++t) {
        uint64_t p = i + t;
        uint64_t c = p < n ? uint64_t(text[p]) + 1 : 0;
        k = (k << 3) | c;
    }
    keys[i] = k;
    vals[i] = static_cast<uint32_t>(i);
}

// flags[j] = 1 when sorted key j starts a new group
__global__ void group_flags_kernel(const uint64_t* keys, uint64_t n, uint8_t* flags) {
    uint64_t j = blockIdx.x * static_cast<uint64_t>(blockDim.x) + threadIdx.x;
    if (j >= n) return;
    flags[j] = (j == 0 || keys[j] != keys[j - 1]) ? 1 : 0;
}

struct HeadOf {
    const uint8_t* flags;
    __device__ uint32_t operator()(uint64_t j) const { return flags[j] ? static_cast<uint32_t>(j) : 0u; }
};
struct MaxU32 {
    __device__ uint32_t operator()(uint32_t a, uint32_t b) const { return a > b ? a : b; }
};

__global__ void scatter_rank_kernel(const uint32_t* sa, const uint32_t* head, uint64_t n, uint32_t* rnk) {
    uint64_t j = blockIdx.x * static_cast<uint64_t>(blockDim.x) + threadIdx.x;
    if (j >= n) return;
    rnk[sa[j]] = head[j];
}

// positions that are not yet alone in their group
struct Unsorted {
    const uint8_t* flags;
    uint64_t n;
    __device__ bool operator()(uint64_t j) const {
        bool single = flags[j] && (j + 1 >= n || flags[j + 1]);
        return !single;
    }
};

__global__ void refine_keys_kernel(const uint32_t* U, uint64_t m, const uint32_t* sa, const uint32_t* head, const uint32_t* rnk,
                                   uint64_t n, uint64_t h, uint64_t* keys, uint32_t* vals) {
    uint64_t t = blockIdx.x * static_cast<uint64_t>(blockDim.x) + threadIdx.x;
    if (t >= m) return;
    uint32_t j = U[t];
    uint32_t s = sa[j];
    uint64_t p = uint64_t(s) + h;
    uint64_t r = p < n ? uint64_t(rnk[p]) + 1 : 0;
    keys[t] = (uint64_t(head[j]) << 32) | r;
    vals[t] = s;
}

__global__ void refine_flags_kernel(const uint32_t* U, uint64_t m, const uint64_t* keys, const uint32_t* vals, uint32_t* sa,
                                    uint8_t* flags, uint32_t* newhead_in) {
    uint64_t t = blockIdx.x * static_cast<uint64_t>(blockDim.x) + threadIdx.x;
    if (t >= m) return;
    uint32_t j = U[t];
    bool f = (t == 0 || keys[t] != keys[t - 1]);
    sa[j] = vals[t];
    flags[j] = f ? 1 : 0;
    newhead_in[t] = f ? j : 0u;
}

__global__ void refine_scatter_kernel(const uint32_t* U, uint64_t m, const uint32_t* vals, const uint32_t* newhead, uint32_t* head,
                                      uint32_t* rnk) {
    uint64_t t = blockIdx.x * static_cast<uint64_t>(blockDim.x) + threadIdx.x;
    if (t >= m) return;
    head[U[t]] = newhead[t];
    rnk[vals[t]] = newhead[t];
}

__global__ void bwt_kernel(const uint8_t* text, const uint32_t* sa, uint64_t n, uint8_t* bwt) {
    uint64_t j = blockIdx.x * static_cast<uint64_t>(blockDim.x) + threadIdx.x;
    if (j >= n) return;
    uint32_t p = sa[j];
    bwt[j] = text[p == 0 ? n - 1 : p - 1];
}

struct Cnt8 {
    uint32_t c[8];
};
struct AddCnt8 {
    __device__ Cnt8 operator()(const Cnt8& a, const Cnt8& b) const {
        Cnt8 r;
#pragma unroll
        for (int i = 0; i < 8; ++i) r.c[i] = a.c[i] + b.c[i];
        return r;
    }
};

// one thread block (64 threads) per superblock of 4096 rows; thread t packs block sb*64+t.
// bwt must be readable (zero padded) up to a multiple of 4096 rows.
__global__ void __launch_bounds__(64) pack_occ_kernel(const uint8_t* bwt, uint64_t n_rows, uint64_t n_blocks, OccBlk* blk, Cnt8* sup_tot) {
    uint64_t sb = blockIdx.x;
    uint64_t b = sb * 64 + threadIdx.x;
    uint64_t row0 = b * 64;
    uint64_t p0 = 0, p1 = 0, p2 = 0;
    const uint4* src = reinterpret_cast<const uint4*>(bwt + row0);
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        uint4 v = src[q];
        uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
        for (int k = 0; k < 4; ++k) {
#pragma unroll
            for (int by = 0; by < 4; ++by) {
                int r = q * 16 + k * 4 + by;
                uint32_t c = (w[k] >> (8 * by)) & 0xffu;
                if (row0 + r >= n_rows) c = 0;
                p0 |= uint64_t(c & 1u) << r;
                p1 |= uint64_t((c >> 1) & 1u) << r;
                p2 |= uint64_t((c >> 2) & 1u) << r;
            }
        }
    }
    uint32_t cnt[6];
    cnt[0] = 0;
    cnt[1] = __popcll(p0 & ~p1 & ~p2);
    cnt[2] = __popcll(p1 & ~p0);
    cnt[3] = __popcll(p0 & p1);
    cnt[4] = __popcll(p2 & ~p0);
    cnt[5] = __popcll(p2 & p0);
    // exclusive scan over the 64 threads of 16-bit packed counters (A,C,G,T) and a separate one for N
    typedef cub::BlockScan<uint64_t, 64> Scan64;
    typedef cub::BlockScan<uint32_t, 64> Scan32;
    __shared__ typename Scan64::TempStorage t64;
    __shared__ typename Scan32::TempStorage t32;
    uint64_t packed = uint64_t(cnt[1]) | (uint64_t(cnt[2]) << 16) | (uint64_t(cnt[3]) << 32) | (uint64_t(cnt[4]) << 48);
    uint64_t ex64;
    uint32_t ex32;
    Scan64(t64).ExclusiveSum(packed, ex64);
    Scan32(t32).ExclusiveSum(cnt[5], ex32);
    uint32_t ex[6];
    ex[1] = uint32_t(ex64) & 0xffffu;
    ex[2] = uint32_t(ex64 >> 16) & 0xffffu;
    ex[3] = uint32_t(ex64 >> 32) & 0xffffu;
    ex[4] = uint32_t(ex64 >> 48) & 0xffffu;
    ex[5] = ex32;
    if (b < n_blocks) {
        OccBlk o;
        o.p0 = p0; o.p1 = p1; o.p2 = p2;
        o.ctr = uint64_t(ex[1]) | (uint64_t(ex[2]) << 12) | (uint64_t(ex[3]) << 24) | (uint64_t(ex[4]) << 36) | (uint64_t(ex[5]) << 48);
        blk[b] = o;
    }
    if (threadIdx.x == 63) {
        Cnt8 tot;
        uint32_t sum = 0;
        for (int s = 1; s < 6; ++s) {
            tot.c[s] = ex[s] + cnt[s];
            sum += tot.c[s];
        }
        uint64_t first = sb * 4096;
        uint64_t valid = n_rows > first ? (n_rows - first < 4096 ? n_rows - first : 4096) : 0;
        tot.c[0] = static_cast<uint32_t>(valid) - sum;
        tot.c[6] = tot.c[7] = 0;
        sup_tot[sb] = tot;
    }
}

__global__ void sup_write_kernel(const Cnt8* scanned, uint64_t n_sup, OccSup* sup) {
    uint64_t i = blockIdx.x * static_cast<uint64_t>(blockDim.x) + threadIdx.x;
    if (i >= n_sup) return;
    OccSup s;
#pragma unroll
    for (int k = 0; k < 8; ++k) s.c[k] = scanned[i].c[k];
    sup[i] = s;
}

// marker words: row j is sampled when its in-sequence text position is a multiple of `rate`
__global__ void sample_marks_kernel(const uint32_t* sa, uint64_t n, SeqMap m, uint32_t rate, uint32_t* words32) {
    uint64_t j = blockIdx.x * static_cast<uint64_t>(blockDim.x) + threadIdx.x;
    bool marked = false;
    if (j < n) {
        uint64_t p = sa[j];
        uint32_t s = find_seq(m, p);
        marked = ((p - m.start[s]) % rate) == 0;
    }
    uint32_t ballot = __ballot_sync(0xffffffffu, marked);
    if ((threadIdx.x & 31) == 0 && (j >> 5) < 2 * (n / 64 + 1)) words32[j >> 5] = ballot;
}

// sampled values in row order: ssa[rank(j)] = (seqId << bits) | seqPos for every marked row j
__global__ void sample_values_kernel(const uint32_t* sa, uint64_t n, SeqMap m, uint32_t rate, uint32_t bits, const MarkRec* marks,
                                     uint64_t* ssa) {
    uint64_t j = blockIdx.x * static_cast<uint64_t>(blockDim.x) + threadIdx.x;
    if (j >= n) return;
    uint64_t p = sa[j];
    uint32_t s = find_seq(m, p);
    uint64_t pos = p - m.start[s];
    if (pos % rate) return;
    MarkRec r = marks[j / kRowsPerMark];
    uint32_t mo = static_cast<uint32_t>(j % kRowsPerMark);
    uint32_t w = mo >> 6, o = mo & 63u;
    uint32_t rk = r.rank + __popcll(r.bits[w] & ((uint64_t{1} << o) - 1));
    if (w > 0) rk += __popcll(r.bits[0]);
    if (w > 1) rk += __popcll(r.bits[1]);
    ssa[rk] = (uint64_t(s) << bits) | pos;
}

}  // namespace sb200
