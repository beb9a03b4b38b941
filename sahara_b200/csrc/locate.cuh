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
    // for cursors that carry a text position instead of a row (in-text verification): start of every sequence in the
    // delimited text, n_seqs + 1 entries
    const uint64_t* seq_start;
    uint32_t n_seqs;
    uint32_t bits;         // bits_for_position
};

// cursor.w flag: cursor.y is the text position of the (single) occurrence, not a suffix-array row.  The in-text
// verification knows the position; going through ISA and SA again would cost two random reads per hit.
constexpr uint32_t kCursorTextPos = 0x10u;

// text position -> (seqId << bits) | seqPos
__device__ __forceinline__ uint64_t textpos_value(const LocateIndex& X, uint32_t a) {
    uint32_t lo = 0, hi = X.n_seqs;  // largest i with seq_start[i] <= a
    while (hi - lo > 1) {
        const uint32_t mid = lo + ((hi - lo) >> 1);
        if (X.seq_start[mid] <= a) lo = mid;
        else hi = mid;
    }
    return (static_cast<uint64_t>(lo) << X.bits) | (a - X.seq_start[lo]);
}

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
        uint64_t v = (cur.w & kCursorTextPos) ? textpos_value(P.index, cur.y) : locate_row<SIGMA>(P.index, row, steps);
        const uint32_t e = cur.w & 0xfu;
        if (P.fused_shift) {
            P.out_key[j] = (static_cast<uint64_t>(cur.x) << P.fused_shift) | (v << 4) | e;
        } else {
            P.out_key[j] = (v << 4) | e;
            P.out_qid[j] = cur.x;
        }
    }
    // warp-aggregated step counter
    for (int o = 16; o > 0; o >>= 1) steps += __shfl_xor_sync(0xffffffffu, steps, o);
    if ((threadIdx.x & 31) == 0 && steps) atomicAdd(&P.counters[4], static_cast<unsigned long long>(steps));
}

// ---- locate + sort by query buckets ------------------------------------------------------------------
// The hits of a search are wanted in the order (query, sequence, position, errors).  A query has few hits (8 on
// average in the headline workload), so instead of radix-sorting all hits by 57 bits (8 passes over 16 M pairs):
//   hit_count_kernel      hits per query (one pass over the cursors, red.add)
//   (exclusive scan)      first slot of every query
//   locate_scatter_kernel one thread per cursor: locates its rows and writes them into the slots of its query
//                         (cursors with many rows are cut into tasks for locate_tasks_kernel)
//   segment_sort_kernel   one warp per 32 queries: bitonic sort of each segment (<= 32 keys in registers over shuffles,
//                         <= 256 in shared memory); longer segments go to
//   segment_sort_big_kernel (one block per segment, bitonic sort of <= 2048 keys in shared memory); a query with
//                         more hits than that sets a flag and the host falls back to the global radix sort.
constexpr uint32_t kInlineRows = 8;      // rows a thread of locate_scatter_kernel locates itself
constexpr uint32_t kTaskRows = 1024;     // rows per task of locate_tasks_kernel
constexpr uint32_t kWarpSeg = 256;       // keys a warp of segment_sort_kernel sorts
constexpr uint32_t kRankSortMax = 20;    // ... by counting ranks instead of a sorting network
constexpr uint32_t kBigSeg = 2048;       // keys a block of segment_sort_big_kernel sorts
enum : int { LC_TASKS = 0, LC_BIG_SEGS = 1, LC_HUGE = 2, LC_KEY_OVERFLOW = 3, LC_COUNT = 4 };

// bytes of d as a variable-length integer, 7 bits per byte (delta-coded records, below)
__device__ __forceinline__ uint32_t varint_bytes(uint64_t d) {
    uint32_t n = 1;
    while (d >= 128u) { d >>= 7; ++n; }
    return n;
}

struct BucketParams {
    LocateIndex index;
    const uint4* cursors;  // (qid, lb or text position, len, e | flags)
    uint32_t n_cursors;    // number of cursor slots; with n_cursors_dev: the capacity of the cursor array
    const unsigned long long* n_cursors_dev;  // optional: the slot count as the search kernels left it in device memory (no host round trip)
    uint32_t n_queries;
    uint32_t* qpos;        // [n_queries + 1] hits per query -> first slot -> (after the scatter) end of the segment
    uint64_t* keys;        // (value << 4) | e
    uint32_t* qids;        // query of every hit (nullptr: not wanted — qpos already says it)
    uint32_t key_cap;      // capacity of keys / qids; a query whose segment does not fit sets lc[LC_KEY_OVERFLOW]
    uint4* tasks;          // (cursor, first row of the task, first slot, -)
    uint32_t task_cap;
    uint32_t* big_segs;    // queries with more than kWarpSeg hits
    uint32_t big_cap;
    unsigned int* lc;      // LC_* counters
    unsigned long long* counters;  // [4] LF steps
};

// hits per query.  The counts and their scan are u32; *extra accumulates the rows beyond the first of every cursor
// (64 bit, touched only by warps that see a cursor with several rows), so that n_cursors + *extra bounds the number of
// hits and the host can refuse a call whose scan might have wrapped.
__device__ __forceinline__ uint32_t bucket_cursor_count(uint32_t n_cursors, const unsigned long long* n_dev) {
    if (n_dev == nullptr) return n_cursors;
    const unsigned long long n = *n_dev;
    return n < n_cursors ? static_cast<uint32_t>(n) : n_cursors;
}

__global__ void __launch_bounds__(256) hit_count_kernel(const uint4* cursors, uint32_t n_cursors, const unsigned long long* n_dev, uint32_t* qcount,
                                                        unsigned long long* extra) {
    const uint32_t n = bucket_cursor_count(n_cursors, n_dev);
    const uint32_t stride = gridDim.x * blockDim.x;
    unsigned long long more = 0;
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        const uint4 c = cursors[i];
        if (c.z != 0) atomicAdd(&qcount[c.x], c.z);
        if (c.z > 1) more += c.z - 1;
    }
    if (__any_sync(0xffffffffu, more != 0)) {
        for (int o = 16; o > 0; o >>= 1) more += __shfl_xor_sync(0xffffffffu, more, o);
        if ((threadIdx.x & 31) == 0) atomicAdd(extra, more);
    }
}

template <int SIGMA>
__global__ void __launch_bounds__(256) locate_scatter_kernel(const BucketParams P) {
    const uint32_t n = bucket_cursor_count(P.n_cursors, P.n_cursors_dev);
    const uint32_t stride = gridDim.x * blockDim.x;
    uint32_t steps = 0;
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        const uint4 c = P.cursors[i];
        if (c.z == 0) continue;
        const uint32_t base = atomicAdd(&P.qpos[c.x], c.z);
        if (base + c.z > P.key_cap || base + c.z < base) {  // the hit buffer is too small: the host grows it and starts over
            P.lc[LC_KEY_OVERFLOW] = 1u;
            continue;
        }
        const uint32_t e = c.w & 0xfu;
        if (c.w & kCursorTextPos) {  // (always a single row)
            P.keys[base] = (textpos_value(P.index, c.y) << 4) | e;
            if (P.qids) P.qids[base] = c.x;
        } else if (c.z <= kInlineRows) {
            for (uint32_t r = 0; r < c.z; ++r) {
                uint32_t st = 0;
                P.keys[base + r] = (locate_row<SIGMA>(P.index, c.y + r, st) << 4) | e;
                if (P.qids) P.qids[base + r] = c.x;
                steps += st;
            }
        } else {
            for (uint32_t off = 0; off < c.z; off += kTaskRows) {
                const uint32_t t = atomicAdd(&P.lc[LC_TASKS], 1u);
                if (t < P.task_cap) P.tasks[t] = make_uint4(i, off, base + off, 0);
            }
        }
    }
    for (int o = 16; o > 0; o >>= 1) steps += __shfl_xor_sync(0xffffffffu, steps, o);
    if ((threadIdx.x & 31) == 0 && steps) atomicAdd(&P.counters[4], static_cast<unsigned long long>(steps));
}

// one warp per task: up to kTaskRows rows of one cursor
template <int SIGMA>
__global__ void __launch_bounds__(256) locate_tasks_kernel(const BucketParams P) {
    const uint32_t n_tasks = P.lc[LC_TASKS] < P.task_cap ? P.lc[LC_TASKS] : P.task_cap;
    const uint32_t lane = threadIdx.x & 31u;
    const uint32_t warps = gridDim.x * (blockDim.x >> 5);
    uint32_t steps = 0;
    for (uint32_t t = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); t < n_tasks; t += warps) {
        const uint4 task = P.tasks[t];
        const uint4 c = P.cursors[task.x];
        const uint32_t end = task.y + kTaskRows < c.z ? task.y + kTaskRows : c.z;
        for (uint32_t r = task.y + lane; r < end; r += 32u) {
            uint32_t st = 0;
            P.keys[task.z + (r - task.y)] = (locate_row<SIGMA>(P.index, c.y + r, st) << 4) | (c.w & 0xfu);
            if (P.qids) P.qids[task.z + (r - task.y)] = c.x;
            steps += st;
        }
    }
    for (int o = 16; o > 0; o >>= 1) steps += __shfl_xor_sync(0xffffffffu, steps, o);
    if (lane == 0 && steps) atomicAdd(&P.counters[4], static_cast<unsigned long long>(steps));
}

// one warp per 32 consecutive queries: the lanes read the segment bounds, then the warp sorts the segments with two
// or more keys one after the other — up to 32 keys in registers (bitonic network over shuffles), up to kWarpSeg keys
// in the warp's shared buffer; longer segments go to segment_sort_big_kernel
__global__ void __launch_bounds__(256) segment_sort_kernel(const BucketParams P) {
    __shared__ uint64_t sbuf[8][kWarpSeg];
    const uint32_t lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
    const uint32_t q = (blockIdx.x * (blockDim.x >> 5) + warp) * 32u + lane;
    uint32_t s_l = 0, n_l = 0;
    if (q < P.n_queries) {
        s_l = q ? P.qpos[q - 1] : 0u;
        n_l = P.qpos[q] - s_l;
        if (P.qpos[q] > P.key_cap || P.qpos[q] < s_l) n_l = 0;  // (overflowed hit buffer: the host starts over)
    }
    uint64_t* buf = sbuf[warp];
    uint32_t todo = __ballot_sync(0xffffffffu, n_l >= 2);
    while (todo != 0) {
        const int src = __ffs(static_cast<int>(todo)) - 1;
        todo &= todo - 1;
        const uint32_t s = __shfl_sync(0xffffffffu, s_l, src), n = __shfl_sync(0xffffffffu, n_l, src);
        uint64_t* k = P.keys + s;
        if (n <= kRankSortMax) {
            // few keys: every lane counts the keys that precede its own (n broadcasts), cheaper than the 15 stages
            // of the 32-key network
            const uint64_t v = lane < n ? k[lane] : ~uint64_t{0};
            uint32_t rank = 0;
            for (uint32_t j = 0; j < n; ++j) {
                const uint64_t o = __shfl_sync(0xffffffffu, v, static_cast<int>(j));
                rank += (o < v || (o == v && j < lane)) ? 1u : 0u;
            }
            if (lane < n) k[rank] = v;
        } else if (n <= 32u) {
            uint64_t v = lane < n ? k[lane] : ~uint64_t{0};
#pragma unroll
            for (uint32_t size = 2; size <= 32u; size <<= 1)
#pragma unroll
                for (uint32_t stride = size >> 1; stride > 0; stride >>= 1) {
                    const uint64_t o = __shfl_xor_sync(0xffffffffu, v, stride);
                    const bool up = (lane & size) == 0 || size == 32u;
                    const bool lower = (lane & stride) == 0;  // this lane keeps the smaller key of the pair when ascending
                    v = ((v < o) == (lower == up)) ? v : o;
                }
            if (lane < n) k[lane] = v;
        } else if (n <= kWarpSeg) {
            uint32_t m = 64;  // power of two >= n
            while (m < n) m <<= 1;
            for (uint32_t i = lane; i < m; i += 32u) buf[i] = i < n ? k[i] : ~uint64_t{0};
            __syncwarp();
            for (uint32_t size = 2; size <= m; size <<= 1)
                for (uint32_t stride = size >> 1; stride > 0; stride >>= 1) {
                    for (uint32_t i = lane; i < m / 2; i += 32u) {
                        const uint32_t lo = 2 * i - (i & (stride - 1));
                        const uint32_t hi = lo + stride;
                        const bool up = (lo & size) == 0;
                        const uint64_t x = buf[lo], y = buf[hi];
                        if ((x > y) == up) {
                            buf[lo] = y;
                            buf[hi] = x;
                        }
                    }
                    __syncwarp();
                }
            for (uint32_t i = lane; i < n; i += 32u) k[i] = buf[i];
            __syncwarp();
        } else if (lane == 0) {
            const uint32_t qq = (blockIdx.x * (blockDim.x >> 5) + warp) * 32u + static_cast<uint32_t>(src);
            if (n > kBigSeg) {
                P.lc[LC_HUGE] = 1u;
            } else {
                const uint32_t t = atomicAdd(&P.lc[LC_BIG_SEGS], 1u);
                if (t < P.big_cap) P.big_segs[t] = qq;
                else P.lc[LC_HUGE] = 1u;
            }
        }
    }
}

__global__ void __launch_bounds__(256) segment_sort_big_kernel(const BucketParams P) {
    __shared__ uint64_t sk[kBigSeg];
    const uint32_t n_big = P.lc[LC_BIG_SEGS] < P.big_cap ? P.lc[LC_BIG_SEGS] : P.big_cap;
    for (uint32_t b = blockIdx.x; b < n_big; b += gridDim.x) {
        const uint32_t q = P.big_segs[b];
        const uint32_t s = q ? P.qpos[q - 1] : 0u, n = P.qpos[q] - s;
        if (n > kBigSeg) continue;  // (cannot happen: segment_sort_kernel sends only segments that fit)
        uint32_t m = 64;  // power of two >= n
        while (m < n) m <<= 1;
        for (uint32_t i = threadIdx.x; i < m; i += blockDim.x) sk[i] = i < n ? P.keys[s + i] : ~uint64_t{0};
        __syncthreads();
        for (uint32_t size = 2; size <= m; size <<= 1)
            for (uint32_t stride = size >> 1; stride > 0; stride >>= 1) {
                for (uint32_t i = threadIdx.x; i < m / 2; i += blockDim.x) {
                    const uint32_t lo = 2 * i - (i & (stride - 1));  // element with the stride bit clear
                    const uint32_t hi = lo + stride;
                    const bool up = (lo & size) == 0;
                    const uint64_t x = sk[lo], y = sk[hi];
                    if ((x > y) == up) {
                        sk[lo] = y;
                        sk[hi] = x;
                    }
                }
                __syncthreads();
            }
        for (uint32_t i = threadIdx.x; i < n; i += blockDim.x) P.keys[s + i] = sk[i];
        __syncthreads();
    }
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

// number of hits: given by the host, or read from device memory (n_dev: the last entry of the per-query scan) and
// clamped to the capacity `n`
__device__ __forceinline__ uint64_t hit_total(uint64_t n, const uint32_t* n_dev) {
    if (n_dev == nullptr) return n;
    const uint64_t t = *n_dev;
    return t < n ? t : n;
}

// sorted hits -> the reference's result tuple (queryId, seqId, pos, errors), 4 x u64
__global__ void expand_hits_kernel(const uint64_t* keys, const uint32_t* qids, uint64_t n, const uint32_t* n_dev, uint32_t bits,
                                   uint64_t first_query, uint32_t fused_shift, uint64_t* out) {
    const uint64_t total = hit_total(n, n_dev), stride = static_cast<uint64_t>(gridDim.x) * blockDim.x;
    for (uint64_t i = blockIdx.x * static_cast<uint64_t>(blockDim.x) + threadIdx.x; i < total; i += stride) {
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
}

// sorted hits -> compact hits (query_id, seq_id, pos, errors) as 4 x u32
__global__ void compact_hits_kernel(const uint64_t* keys, const uint32_t* qids, uint64_t n, const uint32_t* n_dev, uint32_t bits,
                                    uint32_t first_query, uint32_t fused_shift, uint4* out) {
    const uint64_t total = hit_total(n, n_dev), stride = static_cast<uint64_t>(gridDim.x) * blockDim.x;
    for (uint64_t i = blockIdx.x * static_cast<uint64_t>(blockDim.x) + threadIdx.x; i < total; i += stride) {
        uint32_t qid;
        uint64_t k;
        hit_key(keys, qids, i, fused_shift, qid, k);
        uint64_t v = k >> 4;
        out[i] = make_uint4(first_query + qid, static_cast<uint32_t>(v >> bits), static_cast<uint32_t>(v & ((uint64_t{1} << bits) - 1)),
                            static_cast<uint32_t>(k & 15u));
    }
}

// sorted hits -> records of rec_bytes bytes, little endian: ((seqId << bits | pos) << 4) | errors.  The query of a record
// follows from the per-query ends (CSR), so a hit costs rec_bytes (5 for a human-sized genome) instead of 16 or 32 bytes
// on the way to the host.  One thread per 4 records: whole words are stored.
__global__ void pack_records_kernel(const uint64_t* keys, uint64_t n, const uint32_t* n_dev, uint32_t fused_shift, uint32_t rec_bytes,
                                    uint8_t* out) {
    const uint64_t total = hit_total(n, n_dev), stride = static_cast<uint64_t>(gridDim.x) * blockDim.x;
    const uint64_t mask = fused_shift ? (uint64_t{1} << fused_shift) - 1 : ~uint64_t{0};
    const uint64_t recmask = rec_bytes >= 8 ? ~uint64_t{0} : (uint64_t{1} << (8u * rec_bytes)) - 1;
    for (uint64_t g = blockIdx.x * static_cast<uint64_t>(blockDim.x) + threadIdx.x; g * 4 < total; g += stride) {
        // 4 records = 4 * rec_bytes bytes = rec_bytes aligned words
        uint32_t w[8] = {0, 0, 0, 0, 0, 0, 0, 0};
        for (uint32_t r = 0; r < 4; ++r) {
            const uint64_t i = g * 4 + r;
            const uint64_t k = i < total ? (keys[i] & mask & recmask) : 0;
            const uint32_t bit = r * rec_bytes * 8u;  // first bit of the record inside the group
            const uint32_t wi = bit >> 5, sh = bit & 31u;
            w[wi] |= static_cast<uint32_t>(k << sh);
            const uint64_t rest = sh ? (k >> (32u - sh)) : (k >> 16 >> 16);
            w[wi + 1] |= static_cast<uint32_t>(rest);
            if (wi + 2 < 8) w[wi + 2] |= static_cast<uint32_t>(rest >> 16 >> 16);
        }
        uint32_t* dst = reinterpret_cast<uint32_t*>(out) + g * rec_bytes;  // (out is 16-byte aligned, sized to whole groups)
        for (uint32_t j = 0; j < rec_bytes; ++j) dst[j] = w[j];
    }
}

// ---- delta-coded records: the hits of a query lie close together (the alignments of one locus), so behind the first record
// of a query (rec_bytes, as above) every record is the difference to its predecessor as a variable-length integer, 7 bits per
// byte, low bits first, the top bit of a byte says that another follows.  One thread per query; sizes first, then (after a
// scan of the sizes) the bytes.
// sizes and bytes: one thread per query.  (Measured and dropped: staging the keys of a warp's 32 queries in shared memory for
// coalesced loads / stores — 0.17 ms instead of 0.14 for the writes, the sizes unchanged; sizing inside segment_sort_kernel with
// all 32 segments of a warp sorted at once from a staged range, one query per lane (insertion sort) or one key per lane (rank
// counting) — the sort went from 0.37 to 0.52 / 0.47 ms on the headline workload, more than the separate pass costs.)
__device__ __forceinline__ uint64_t delta_key_mask(uint32_t fused_shift, uint32_t rec_bytes) {
    return (fused_shift ? (uint64_t{1} << fused_shift) - 1 : ~uint64_t{0}) & (rec_bytes >= 8 ? ~uint64_t{0} : (uint64_t{1} << (8u * rec_bytes)) - 1);
}
__global__ void __launch_bounds__(256) delta_size_kernel(const uint64_t* keys, uint32_t key_cap, const uint32_t* ends, uint32_t n_queries,
                                                         uint32_t fused_shift, uint32_t rec_bytes, uint32_t* sizes) {
    const uint32_t q = blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= n_queries) return;
    const uint64_t mask = delta_key_mask(fused_shift, rec_bytes);
    uint32_t b = q ? ends[q - 1] : 0u, e = ends[q];
    if (e > key_cap || e < b) b = e = 0;  // (overflowed hit buffer: the host starts the batch over)
    uint32_t sz = 0;
    uint64_t prev = 0;
    for (uint32_t i = b; i < e; ++i) {
        const uint64_t v = keys[i] & mask;
        sz += i == b ? rec_bytes : varint_bytes(v - prev);
        prev = v;
    }
    sizes[q] = sz;
}
__global__ void __launch_bounds__(256) delta_write_kernel(const uint64_t* keys, uint32_t key_cap, const uint32_t* ends, const uint32_t* byte_ends,
                                                          uint32_t n_queries, uint32_t fused_shift, uint32_t rec_bytes, uint8_t* out) {
    const uint32_t q = blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= n_queries) return;
    const uint64_t mask = delta_key_mask(fused_shift, rec_bytes);
    uint32_t b = q ? ends[q - 1] : 0u, e = ends[q];
    if (e > key_cap || e < b) return;  // (overflowed hit buffer: the host starts the batch over)
    uint8_t* dst = out + (q ? byte_ends[q - 1] : 0u);
    uint64_t prev = 0;
    for (uint32_t i = b; i < e; ++i) {
        const uint64_t v = keys[i] & mask;
        if (i == b) {
            for (uint32_t j = 0; j < rec_bytes; ++j) *dst++ = static_cast<uint8_t>(v >> (8u * j));
        } else {
            uint64_t d = v - prev;
            while (d >= 128u) { *dst++ = static_cast<uint8_t>(d) | 0x80u; d >>= 7; }
            *dst++ = static_cast<uint8_t>(d);
        }
        prev = v;
    }
}

// per-query ends of a hit list sorted by query (global radix-sort path): ends[q] = number of hits of queries <= q
__global__ void csr_ends_kernel(const uint64_t* keys, const uint32_t* qids, uint64_t n, uint32_t fused_shift, uint32_t n_queries, uint32_t* ends) {
    const uint64_t stride = static_cast<uint64_t>(gridDim.x) * blockDim.x;
    for (uint64_t i = blockIdx.x * static_cast<uint64_t>(blockDim.x) + threadIdx.x; i <= n; i += stride) {
        // hit i-1 belongs to query a, hit i to query b: every query in [a, b) ends at i
        uint32_t a = 0, b = n_queries;
        if (i > 0) a = fused_shift ? static_cast<uint32_t>(keys[i - 1] >> fused_shift) : qids[i - 1];
        if (i < n) b = fused_shift ? static_cast<uint32_t>(keys[i] >> fused_shift) : qids[i];
        for (uint32_t q = a; q < b && q < n_queries; ++q) ends[q] = static_cast<uint32_t>(i);
    }
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
