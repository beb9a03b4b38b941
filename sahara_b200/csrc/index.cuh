// index.cuh — conversion between the reference index image and the device layout, q-gram table.
//
// The reference image is what `archive(index)` stores for fmc::BiFMIndex<Sigma, InterleavedBitvector16>
// (/root/reference/src/sahara/index.cpp:96-100): per 64 rows Sigma u16 counters + Sigma one-hot u64
// bitplanes, unpadded (10*Sigma bytes); one row of Sigma u64 every 65536 rows.
#pragma once
#include "layout.cuh"
#include "locate.cuh"

namespace sb200 {

__device__ __forceinline__ uint64_t load_u64_unaligned2(const uint8_t* p) {  // p is 2-byte aligned
    const uint16_t* q = reinterpret_cast<const uint16_t*>(p);
    return uint64_t(q[0]) | (uint64_t(q[1]) << 16) | (uint64_t(q[2]) << 32) | (uint64_t(q[3]) << 48);
}
__device__ __forceinline__ void store_u64_unaligned2(uint8_t* p, uint64_t v) {
    uint16_t* q = reinterpret_cast<uint16_t*>(p);
    q[0] = uint16_t(v); q[1] = uint16_t(v >> 16); q[2] = uint16_t(v >> 32); q[3] = uint16_t(v >> 48);
}

// absolute counts of all symbols at the start of reference block b
template <int SIGMA>
__device__ __forceinline__ void ref_abs_counts(const uint8_t* raw, const uint64_t* super, uint64_t b, uint64_t* abs) {
    const uint16_t* cnt = reinterpret_cast<const uint16_t*>(raw + b * (10 * SIGMA));
#pragma unroll
    for (int s = 0; s < SIGMA; ++s) abs[s] = super[(b >> 10) * SIGMA + s] + cnt[s];
}

// error flags: 1 = bitplanes overlap / do not cover the rows, 2 = counters inconsistent, 4 = count overflow
template <int SIGMA>
__global__ void convert_ref_occ_kernel(const uint8_t* raw, const uint64_t* super, uint64_t n_blocks, uint64_t n_rows,
                                       OccBlk* blk, OccSup* sup, unsigned int* err) {
    uint64_t b = blockIdx.x * static_cast<uint64_t>(blockDim.x) + threadIdx.x;
    if (b >= n_blocks) return;
    const uint8_t* base = raw + b * (10 * SIGMA);
    uint64_t bits[SIGMA];
    uint64_t all = 0;
    uint32_t pc = 0;
#pragma unroll
    for (int s = 0; s < SIGMA; ++s) {
        bits[s] = load_u64_unaligned2(base + 2 * SIGMA + 8 * s);
        all |= bits[s];
        pc += __popcll(bits[s]);
    }
    uint64_t first = b * 64;
    uint64_t valid = n_rows > first ? n_rows - first : 0;
    uint64_t expect = valid >= 64 ? ~uint64_t{0} : ((uint64_t{1} << valid) - 1);
    if (all != expect || pc != __popcll(expect)) atomicOr(err, 1u);
    uint64_t abs[SIGMA], abs0[SIGMA];
    ref_abs_counts<SIGMA>(raw, super, b, abs);
    ref_abs_counts<SIGMA>(raw, super, b & ~uint64_t{63}, abs0);
    if (b + 1 < n_blocks) {
        uint64_t nxt[SIGMA];
        ref_abs_counts<SIGMA>(raw, super, b + 1, nxt);
#pragma unroll
        for (int s = 0; s < SIGMA; ++s)
            if (nxt[s] != abs[s] + __popcll(bits[s])) atomicOr(err, 2u);
    }
    if (b == 0) {
#pragma unroll
        for (int s = 0; s < SIGMA; ++s)
            if (abs[s] != 0) atomicOr(err, 2u);
    }
    uint64_t b5 = 0;
    if constexpr (SIGMA > 5) b5 = bits[5];
    OccBlk o;
    o.p0 = bits[1] | bits[3] | b5;
    o.p1 = bits[2] | bits[3];
    o.p2 = bits[4] | b5;
    uint64_t ctr = 0;
#pragma unroll
    for (int s = 1; s < SIGMA; ++s) {
        uint64_t rel = abs[s] - abs0[s];
        if (rel > 4095 || abs[s] > 0xffffffffull) atomicOr(err, 4u);
        ctr |= (rel & 0xfff) << (12 * (s - 1));
    }
    o.ctr = ctr;
    blk[b] = o;
    if ((b & 63) == 0) {
        OccSup su;
#pragma unroll
        for (int s = 0; s < 8; ++s) su.c[s] = s < SIGMA ? static_cast<uint32_t>(abs[s]) : 0u;
        sup[b >> 6] = su;
    }
}

// device layout -> reference image
template <int SIGMA>
__global__ void export_ref_occ_kernel(const OccBlk* blk, const OccSup* sup, uint64_t n_blocks, uint64_t n_rows, uint8_t* raw,
                                      uint64_t* super) {
    uint64_t b = blockIdx.x * static_cast<uint64_t>(blockDim.x) + threadIdx.x;
    if (b >= n_blocks) return;
    auto abs_at = [&](uint64_t bb, uint64_t* abs) {
        OccBlk k = blk[bb];
        OccSup s = sup[bb >> 6];
        uint64_t sum = 0;
#pragma unroll
        for (int c = 1; c < SIGMA; ++c) {
            abs[c] = uint64_t(s.c[c]) + blk_ctr(k, c);
            sum += abs[c];
        }
        abs[0] = bb * 64 - sum;
    };
    uint64_t abs[SIGMA], abs0[SIGMA];
    abs_at(b, abs);
    abs_at(b & ~uint64_t{1023}, abs0);
    OccBlk k = blk[b];
    uint64_t first = b * 64;
    uint64_t valid = n_rows > first ? n_rows - first : 0;
    uint64_t vm = valid >= 64 ? ~uint64_t{0} : ((uint64_t{1} << valid) - 1);
    uint64_t bits[6];
    bits[0] = ~(k.p0 | k.p1 | k.p2) & vm;
    bits[1] = k.p0 & ~k.p1 & ~k.p2;
    bits[2] = k.p1 & ~k.p0;
    bits[3] = k.p0 & k.p1;
    bits[4] = k.p2 & ~k.p0;
    bits[5] = k.p2 & k.p0;
    uint8_t* base = raw + b * (10 * SIGMA);
    uint16_t* cnt = reinterpret_cast<uint16_t*>(base);
#pragma unroll
    for (int s = 0; s < SIGMA; ++s) {
        cnt[s] = static_cast<uint16_t>(abs[s] - abs0[s]);
        store_u64_unaligned2(base + 2 * SIGMA + 8 * s, bits[s]);
    }
    if ((b & 1023) == 0) {
#pragma unroll
        for (int s = 0; s < SIGMA; ++s) super[(b >> 10) * SIGMA + s] = abs[s];
    }
}

// ---- marker records ----------------------------------------------------------------------------------
// mark words: bit r of word r/64 set <=> row r sampled.  192 rows = 3 words per record.
__global__ void mark_popc_kernel(const uint64_t* words, uint64_t n_words, uint64_t n_rec, uint32_t* popc) {
    uint64_t r = blockIdx.x * static_cast<uint64_t>(blockDim.x) + threadIdx.x;
    if (r >= n_rec) return;
    uint32_t c = 0;
    for (int k = 0; k < 3; ++k) {
        uint64_t w = r * 3 + k;
        if (w < n_words) c += __popcll(words[w]);
    }
    popc[r] = c;
}
__global__ void mark_fill_kernel(const uint64_t* words, uint64_t n_words, uint64_t n_rec, const uint32_t* rank_before,
                                 MarkRec* rec) {
    uint64_t r = blockIdx.x * static_cast<uint64_t>(blockDim.x) + threadIdx.x;
    if (r >= n_rec) return;
    MarkRec m;
    for (int k = 0; k < 3; ++k) {
        uint64_t w = r * 3 + k;
        m.bits[k] = w < n_words ? words[w] : 0;
    }
    m.rank = rank_before[r];
    m.pad = 0;
    rec[r] = m;
}
__global__ void mark_export_kernel(const MarkRec* rec, uint64_t n_words, uint64_t* words) {
    uint64_t w = blockIdx.x * static_cast<uint64_t>(blockDim.x) + threadIdx.x;
    if (w >= n_words) return;
    words[w] = rec[w / 3].bits[w % 3];
}

// ---- rank probe kernels (kernel 1) -----------------------------------------------------------------
template <int SIGMA>
__global__ void rank_probe_kernel(OccTable t, const uint64_t* pos, uint64_t n, uint64_t* out) {
    uint64_t i = blockIdx.x * static_cast<uint64_t>(blockDim.x) + threadIdx.x;
    if (i >= n) return;
    uint32_t r[SIGMA];
    all_ranks<SIGMA>(t, static_cast<uint32_t>(pos[i]), r);
#pragma unroll
    for (int c = 0; c < SIGMA; ++c) out[i * SIGMA + c] = r[c];
}

__device__ __forceinline__ uint32_t hash32(uint32_t x) {
    x ^= x >> 16; x *= 0x7feb352dU; x ^= x >> 15; x *= 0x846ca68bU; x ^= x >> 16;
    return x;
}

// dependent chains of all_ranks probes: the next row is derived from the ranks just computed, like a
// cursor extension depends on the previous one.  checksum = sum of all ranks (order independent).
template <int SIGMA>
__global__ void __launch_bounds__(256) rank_bench_kernel(OccTable t, uint32_t n_rows, uint64_t n_chains, uint32_t iters,
                                                         uint32_t seed, unsigned long long* checksum) {
    uint64_t i = blockIdx.x * static_cast<uint64_t>(blockDim.x) + threadIdx.x;
    unsigned long long acc = 0;
    if (i < n_chains) {
        uint32_t row = static_cast<uint32_t>((uint64_t(hash32(static_cast<uint32_t>(i) ^ seed)) * (uint64_t(n_rows) + 1)) >> 32);
        for (uint32_t it = 0; it < iters; ++it) {
            uint32_t r[SIGMA];
            all_ranks<SIGMA>(t, row, r);
            uint32_t mix = it;
#pragma unroll
            for (int c = 0; c < SIGMA; ++c) {
                acc += r[c];
                mix = mix * 0x9e3779b1u + r[c];
            }
            row = static_cast<uint32_t>((uint64_t(hash32(mix)) * (uint64_t(n_rows) + 1)) >> 32);
        }
    }
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if ((threadIdx.x & 31) == 0) atomicAdd(checksum, acc);
}

// ---- in-text verification tables -------------------------------------------------------------------
// sa32[row] = global text position of the suffix (start of its sequence + in-sequence position)
__global__ void sa32_kernel(const uint64_t* full_ssa, const uint64_t* seq_start, uint64_t n, uint32_t bits, uint32_t* sa32, uint32_t* isa32) {
    uint64_t r = blockIdx.x * static_cast<uint64_t>(blockDim.x) + threadIdx.x;
    if (r >= n) return;
    uint64_t v = full_ssa[r];
    uint32_t p = static_cast<uint32_t>(seq_start[v >> bits] + (v & ((uint64_t{1} << bits) - 1)));
    sa32[r] = p;
    isa32[p] = static_cast<uint32_t>(r);
}
// T[SA[r] - 1] = bwt[r] (cyclic): scatter the symbols into the zero-initialised packed text
template <int SIGMA>
__global__ void text4_kernel(OccTable bwt, const uint32_t* sa32, uint64_t n, uint32_t* text4) {
    uint64_t r = blockIdx.x * static_cast<uint64_t>(blockDim.x) + threadIdx.x;
    if (r >= n) return;
    uint32_t p = sa32[r];
    p = p == 0 ? static_cast<uint32_t>(n - 1) : p - 1;
    OccBlk b = bwt.blk[r >> kBlkShift];
    uint32_t c = static_cast<uint32_t>(blk_symbol(b, static_cast<uint32_t>(r & 63u)));
    if (c) atomicOr(&text4[p >> 3], c << ((p & 7u) * 4u));
}

// ---- q-gram jump table ------------------------------------------------------------------------------
// level t (strings of t symbols over A,C,G,T, first symbol most significant) from level t-1 by one
// extendRight: cursor(parent + c).
template <int SIGMA>
__global__ void qgram_level_kernel(OccTable bwtRev, const uint32_t* C, const uint4* parent, uint4* child, uint32_t n_child,
                                   uint32_t n_rows) {
    uint32_t code = blockIdx.x * blockDim.x + threadIdx.x;
    if (code >= n_child) return;
    uint4 p = parent[code >> 2];  // (lb, lbRev, len, 0)
    uint32_t c = (code & 3u) + 1;
    uint4 o = make_uint4(0, 0, 0, 0);
    if (p.z != 0) {
        uint32_t r1[SIGMA], r2[SIGMA];
        all_ranks<SIGMA>(bwtRev, p.y, r1);
        all_ranks<SIGMA>(bwtRev, p.y + p.z, r2);
        uint32_t smaller = 0;
        for (uint32_t s = 0; s < c; ++s) smaller += r2[s] - r1[s];
        o.x = p.x + smaller;
        o.y = C[c] + r1[c];
        o.z = r2[c] - r1[c];
    }
    child[code] = o;
}

}  // namespace sb200
