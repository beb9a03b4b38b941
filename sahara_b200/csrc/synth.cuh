// synth.cuh — deterministic synthetic genome / read generators.
//
// Counter based (stateless) so that the GPU kernels and the numpy mirror in sahara_b200/synth.py produce
// identical bytes for the same (seed, index).  The read model follows the behaviour of
// /root/reference/src/sahara/read_simulator.cpp:119-167,244-291: an edit transcript of the read length
// with substitutions / insertions replacing matches and deletions inserted, applied to a window of the
// genome; 10 % of the reads are uniform random.  Both strands of every read are emitted
// (/root/reference/src/sahara/search.cpp:121-123: [2i] = read, [2i+1] = reverse complement).
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace sb200 {

__host__ __device__ inline uint64_t splitmix(uint64_t x) {
    x += 0x9E3779B97F4A7C15ull;
    x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
    x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
    return x ^ (x >> 31);
}
__host__ __device__ inline uint64_t rnd(uint64_t seed, uint64_t counter) { return splitmix(seed * 0xD1342543DE82EF95ull + counter); }

__global__ void synth_genome_kernel(uint64_t n, uint64_t seed, uint8_t* out) {
    uint64_t i = blockIdx.x * static_cast<uint64_t>(blockDim.x) + threadIdx.x;
    if (i >= n) return;
    out[i] = static_cast<uint8_t>(1 + (rnd(seed, i) >> 62));
}

constexpr int kMaxSynthLen = 512;

// one thread per read; draws are numbered so that the numpy mirror can reproduce them:
//  0: random-read coin (x % 10 == 0)   1: number of errors (x % (k+1))   2..: error types (x % 3)
//  16..: transcript placement draws (consumed sequentially)   200: strand   201: window start
//  256 + i: substituted / inserted / random base for read position i
__global__ void synth_reads_kernel(const uint8_t* genome, uint64_t n_bases, uint64_t n_reads, uint32_t len, uint32_t k, int edit,
                                   uint64_t seed, uint64_t first_read, uint8_t* out) {
    uint64_t t = blockIdx.x * static_cast<uint64_t>(blockDim.x) + threadIdx.x;
    if (t >= n_reads) return;
    uint64_t r = first_read + t;
    uint64_t base = r * 1024;
    uint8_t* fwd = out + (2 * t) * uint64_t(len);
    uint8_t* rev = out + (2 * t + 1) * uint64_t(len);
    uint8_t read[kMaxSynthLen];
    if (rnd(seed, base + 0) % 10 == 0) {
        for (uint32_t i = 0; i < len; ++i) read[i] = static_cast<uint8_t>(1 + (rnd(seed, base + 256 + i) >> 62));
    } else {
        uint32_t ne = static_cast<uint32_t>(rnd(seed, base + 1) % (k + 1));
        char tr[kMaxSynthLen + 8];
        uint32_t tl = len;
        for (uint32_t i = 0; i < len; ++i) tr[i] = 'M';
        uint32_t draw = 16;
        for (uint32_t x = 0; x < ne; ++x) {
            uint32_t type = edit ? static_cast<uint32_t>(rnd(seed, base + 2 + x) % 3) : 0;
            if (type < 2) {  // substitution / insertion replace a match
                uint32_t pos;
                do { pos = static_cast<uint32_t>(rnd(seed, base + draw++) % tl); } while (tr[pos] != 'M');
                tr[pos] = type == 0 ? 'S' : 'I';
            } else {  // deletion is inserted
                uint32_t pos = static_cast<uint32_t>(rnd(seed, base + draw++) % (tl + 1));
                for (uint32_t i = tl; i > pos; --i) tr[i] = tr[i - 1];
                tr[pos] = 'D';
                ++tl;
            }
        }
        uint32_t ref_len = 0;
        for (uint32_t i = 0; i < tl; ++i) ref_len += tr[i] != 'I';
        uint64_t start = rnd(seed, base + 201) % (n_bases - ref_len + 1);
        uint64_t p = start;
        uint32_t o = 0;
        for (uint32_t i = 0; i < tl; ++i) {
            char c = tr[i];
            if (c == 'M') read[o++] = genome[p++];
            else if (c == 'S') {
                uint32_t g = genome[p++] - 1;
                read[o] = static_cast<uint8_t>(1 + (g + 1 + rnd(seed, base + 256 + o) % 3) % 4);
                ++o;
            } else if (c == 'I') {
                read[o] = static_cast<uint8_t>(1 + (rnd(seed, base + 256 + o) >> 62));
                ++o;
            } else ++p;
        }
        if (rnd(seed, base + 200) & 1) {  // sample from the reverse strand
            for (uint32_t i = 0; i < len / 2; ++i) {
                uint8_t a = read[i], b = read[len - 1 - i];
                read[i] = static_cast<uint8_t>(5 - b);
                read[len - 1 - i] = static_cast<uint8_t>(5 - a);
            }
            if (len & 1) read[len / 2] = static_cast<uint8_t>(5 - read[len / 2]);
        }
    }
    for (uint32_t i = 0; i < len; ++i) {
        fwd[i] = read[i];
        rev[len - 1 - i] = static_cast<uint8_t>(5 - read[i]);
    }
}

}  // namespace sb200
