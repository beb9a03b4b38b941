// locate.cuh — kernel 3: locate every suffix-array row of every reported cursor.
//
// Replaces the loop `for (auto [sae, offset] : fmc::LocateLinear{index, cursor})`
// (/root/reference/src/sahara/search.cpp:244-250; semantics SURVEY.md §8 a6, §9.2): for each row of
// [lb, lb+len) walk the LF mapping  row = C[c] + rank(row, c), c = bwt[row]  until a sampled row is
// reached; the hit position is sample + steps.
//
// One thread per (cursor, row) pair, found by binary search in the exclusive prefix sum of the cursor
// lengths, so repeats with thousands of rows are spread over many threads.  Per LF step a thread issues
// the marker record and the occ block of the row together (one round trip per step).
#pragma once
#include "layout.cuh"

namespace sb200 {

// marker bits of 192 consecutive rows + number of marked rows before the record: one 32-byte sector
struct __align__(32) MarkRec {
    uint64_t bits[3];
    uint32_t rank;
    uint32_t pad;
};
static_assert(sizeof(MarkRec) == 32, "one sector");
constexpr uint32_t kRowsPerMark = 192;

__device__ __forceinline__ MarkRec load_mark(const MarkRec* p) {
    MarkRec r;
    uint64_t w;
    asm volatile("ld.global.nc.v4.b64 {%0,%1,%2,%3}, [%4];" : "=l"(r.bits[0]), "=l"(r.bits[1]), "=l"(r.bits[2]), "=l"(w) : "l"(p));
    r.rank = static_cast<uint32_t>(w);
    r.pad = static_cast<uint32_t>(w >> 32);
    return r;
}

struct LocateIndex {
    OccTable bwt;
    uint32_t C[8];
    const MarkRec* marks;  // nullptr when the device suffix array is complete (rate 1)
    const uint64_t* ssa;   // (seqId << bits) | seqPos of the marked rows, in row order
};

// LF-walk from `row` to the next sampled row; returns the sample value + number of steps walked
template <int SIGMA>
__device__ __forceinline__ uint64_t locate_row(const LocateIndex& X, uint32_t row, uint32_t& steps) {
    if (X.marks == nullptr) return X.ssa[row];
    while (true) {
        uint32_t mr = row / kRowsPerMark, mo = row % kRowsPerMark;
        MarkRec m = load_mark(X.marks + mr);
        OccBlk b = load_blk(X.bwt.blk + (row >> kBlkShift));
        OccSup s = load_sup(X.bwt.sup + (row >> kSupShift));
        uint32_t w = mo >> 6, o = mo & 63u;
        uint64_t word = w == 0 ? m.bits[0] : (w == 1 ? m.bits[1] : m.bits[2]);
        if ((word >> o) & 1u) {
            uint32_t r = m.rank + __popcll(word & ((uint64_t{1} << o) - 1));
            if (w > 0) r += __popcll(m.bits[0]);
            if (w > 1) r += __popcll(m.bits[1]);
            return X.ssa[r] + steps;
        }
        uint32_t bo = row & 63u;
        int c = blk_symbol(b, bo);
        uint32_t rk;
        if (c == 0) {  // cannot happen for a well-formed index (sequence starts are sampled)
            uint32_t sum = 0;
#pragma unroll
            for (int t = 1; t < SIGMA; ++t) sum += s.c[t] + blk_ctr(b, t) + blk_count(b, bo, t);
            rk = row - sum;
        } else {
            rk = s.c[c] + blk_ctr(b, c) + blk_count(b, bo, c);
        }
        row = X.C[c] + rk;
        ++steps;
    }
}

struct LocateParams {
    LocateIndex index;
    const uint4* cursors;     // (qid, lb, len, e)
    const uint64_t* offsets;  // exclusive prefix sum of len, n_cursors + 1 entries
    uint32_t n_cursors;
    uint64_t n_rows_total;
    uint64_t* out_key;        // ((seqId << bits | pos) << 4) | e; with fused_shift != 0 the query id sits above it
    uint32_t* out_qid;        // unused when fused
    uint32_t fused_shift;     // 0, or the width of the (value, e) part: key = qid << fused_shift | (value << 4 | e)
    unsigned long long* counters;  // [4] LF steps
};

template <int SIGMA>
__global__ void __launch_bounds__(256) locate_kernel(const LocateParams P) {
    uint64_t j = blockIdx.x * static_cast<uint64_t>(blockDim.x) + threadIdx.x;
    uint32_t steps = 0;
    if (j < P.n_rows_total) {
        // largest ci with offsets[ci] <= j
        uint32_t lo = 0, hi = P.n_cursors;
        while (hi - lo > 1) {
            uint32_t mid = lo + ((hi - lo) >> 1);
            if (P.offsets[mid] <= j) lo = mid;
            else hi = mid;
        }
        uint4 cur = P.cursors[lo];
        uint32_t row = cur.y + static_cast<uint32_t>(j - P.offsets[lo]);
        uint64_t v = locate_row<SIGMA>(P.index, row, steps);
        if (P.fused_shift) {
            P.out_key[j] = (static_cast<uint64_t>(cur.x) << P.fused_shift) | (v << 4) | cur.w;
        } else {
            P.out_key[j] = (v << 4) | cur.w;
            P.out_qid[j] = cur.x;
        }
    }
    // warp-aggregated step counter
    for (int o = 16; o > 0; o >>= 1) steps += __shfl_xor_sync(0xffffffffu, steps, o);
    if ((threadIdx.x & 31) == 0 && steps) atomicAdd(&P.counters[4], static_cast<unsigned long long>(steps));
}

struct CursorLen {
    __host__ __device__ uint64_t operator()(const uint4& c) const { return c.z; }
};
struct BitFlag {
    const uint64_t* words;
    __host__ __device__ bool operator()(uint64_t i) const { return (words[i >> 6] >> (i & 63)) & 1u; }
};

// sorted hit i -> (query id, (value << 4) | e); fused keys carry the query id above bit fused_shift
__device__ __forceinline__ void hit_key(const uint64_t* keys, const uint32_t* qids, uint64_t i, uint32_t fused_shift, uint32_t& qid, uint64_t& k) {
    k = keys[i];
    if (fused_shift) {
        qid = static_cast<uint32_t>(k >> fused_shift);
        k &= (uint64_t{1} << fused_shift) - 1;
    } else {
        qid = qids[i];
    }
}

// sorted hits -> the reference's result tuple (queryId, seqId, pos, errors), 4 x u64
__global__ void expand_hits_kernel(const uint64_t* keys, const uint32_t* qids, uint64_t n, uint32_t bits, uint64_t first_query,
                                   uint32_t fused_shift, uint64_t* out) {
    uint64_t i = blockIdx.x * static_cast<uint64_t>(blockDim.x) + threadIdx.x;
    if (i >= n) return;
    uint32_t qid;
    uint64_t k;
    hit_key(keys, qids, i, fused_shift, qid, k);
    uint64_t v = k >> 4;
    ulonglong4 h;
    h.x = first_query + qid;
    h.y = v >> bits;
    h.z = v & ((uint64_t{1} << bits) - 1);
    h.w = k & 15u;
    reinterpret_cast<ulonglong4*>(out)[i] = h;
}

// sorted hits -> compact hits (query_id, seq_id, pos, errors) as 4 x u32
__global__ void compact_hits_kernel(const uint64_t* keys, const uint32_t* qids, uint64_t n, uint32_t bits, uint32_t first_query,
                                    uint32_t fused_shift, uint4* out) {
    uint64_t i = blockIdx.x * static_cast<uint64_t>(blockDim.x) + threadIdx.x;
    if (i >= n) return;
    uint32_t qid;
    uint64_t k;
    hit_key(keys, qids, i, fused_shift, qid, k);
    uint64_t v = k >> 4;
    out[i] = make_uint4(first_query + qid, static_cast<uint32_t>(v >> bits), static_cast<uint32_t>(v & ((uint64_t{1} << bits) - 1)),
                        static_cast<uint32_t>(k & 15u));
}

// queries[2i] = read i, queries[2i+1] = its reverse complement (A<->T, C<->G on ranks 1..4, others unchanged)
__global__ void revcomp_kernel(const uint8_t* reads, uint64_t n_reads, uint32_t len, uint8_t* queries) {
    uint64_t i = blockIdx.x * static_cast<uint64_t>(blockDim.x) + threadIdx.x;
    if (i >= n_reads * len) return;
    uint64_t r = i / len;
    uint32_t p = static_cast<uint32_t>(i % len);
    uint8_t c = reads[i];
    queries[(2 * r) * len + p] = c;
    queries[(2 * r + 1) * len + (len - 1 - p)] = (c >= 1 && c <= 4) ? static_cast<uint8_t>(5 - c) : c;
}

// ---- densify: re-sample the suffix array at a denser rate ------------------------------------------
// Every row locates itself through the existing samples; rows whose in-sequence position is a multiple
// of new_rate are marked and their value kept in row_value[row] (compacted afterwards in row order).
struct DensifyParams {
    LocateIndex index;
    uint32_t n_rows;
    uint32_t new_rate;
    uint64_t pos_mask;        // (1 << bits_for_position) - 1
    uint64_t* new_mark_bits;  // n_rows/64+1 words; one thread block owns whole words (blockDim % 64 == 0)
    uint64_t* row_value;      // [n_rows]
};

template <int SIGMA>
__global__ void __launch_bounds__(256) densify_kernel(const DensifyParams P) {
    uint32_t row = blockIdx.x * blockDim.x + threadIdx.x;
    bool marked = false;
    if (row < P.n_rows) {
        uint32_t steps = 0;
        uint64_t v = locate_row<SIGMA>(P.index, row, steps);
        marked = ((v & P.pos_mask) % P.new_rate) == 0;
        P.row_value[row] = v;
    }
    uint32_t ballot = __ballot_sync(0xffffffffu, marked);
    if ((threadIdx.x & 31) == 0 && (row >> 6) < (P.n_rows / 64 + 1)) {
        reinterpret_cast<uint32_t*>(P.new_mark_bits)[row >> 5] = ballot;
    }
}

}  // namespace sb200
