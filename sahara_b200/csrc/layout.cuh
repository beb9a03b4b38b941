// layout.cuh — device layout of the occurrence tables and the rank primitives (kernel 1).
//
// Reference layout (fmc::string::InterleavedBitvector16, used at /root/reference/src/sahara/search.cpp:162):
// per 64 rows Sigma u16 counters + Sigma one-hot u64 bitplanes = 60 B padded to a 64-byte line, plus a
// superblock row every 65536 rows.
//
// B200 layout (measured motivation in DESIGN.md / profiles/r01_gather_microbench.txt): a random probe
// costs one L2-miss REQUEST whatever its size (<= 128 B), and every per-thread load instruction is its
// own request.  So one probed row must be ONE 32-byte sector fetched by ONE 256-bit load:
//
//   OccBlk (32 B, 64 rows): three binary bitplanes p0,p1,p2 of the symbol code (rank 0..5) and one word
//       holding five 12-bit counters = occurrences of ranks 1..5 between the superblock start and the
//       block start.  Rank 0 ('$') is derived: rows - sum(others).
//   OccSup (32 B, every 4096 rows): absolute u32 counts of ranks 0..5 at the superblock start.  The
//       table is n/4096*32 B (24 MB for a 3.1 Gbp text) and stays L2 resident.
//
// 4 bits per row instead of the reference's 8: a 3.1 Gbp index is 2 x 1.55 GB.
#pragma once
#include <cstdint>
#if defined(SB200_HOST_EMU)
// Host emulation of the device code (tests/host_emu): the kernel bodies are compiled with g++ and run as a
// single thread so that the traversal logic can be checked against the oracle without a GPU.
#include <cstring>
#define __device__
#define __host__
#define __forceinline__ inline
#define __align__(n) alignas(n)
struct uint4 { uint32_t x, y, z, w; };
struct uint2 { uint32_t x, y; };
static inline uint4 make_uint4(uint32_t x, uint32_t y, uint32_t z, uint32_t w) { return uint4{x, y, z, w}; }
static inline uint2 make_uint2(uint32_t x, uint32_t y) { return uint2{x, y}; }
static inline int __popcll(unsigned long long v) { return __builtin_popcountll(v); }
static inline unsigned long long atomicAdd(unsigned long long* p, unsigned long long v) { unsigned long long o = *p; *p += v; return o; }
static inline uint32_t atomicAdd(uint32_t* p, uint32_t v) { uint32_t o = *p; *p += v; return o; }
static inline unsigned long long atomicExch(unsigned long long* p, unsigned long long v) { unsigned long long o = *p; *p = v; return o; }
static inline unsigned long long atomicMax(unsigned long long* p, unsigned long long v) { unsigned long long o = *p; if (v > o) *p = v; return o; }
#else
#include <cuda_runtime.h>
#endif

namespace sb200 {

struct __align__(32) OccBlk {
    uint64_t p0, p1, p2, ctr;
};
struct __align__(32) OccSup {
    uint32_t c[8];
};
static_assert(sizeof(OccBlk) == 32 && sizeof(OccSup) == 32, "one sector each");

constexpr int kRowsPerBlk = 64;
constexpr int kBlkShift = 6;
constexpr int kSupShift = 12;  // 4096 rows
constexpr int kBlksPerSup = 64;

struct OccTable {
    const OccBlk* blk;
    const OccSup* sup;
};

struct U32x8 {
    uint32_t v[8];
};

// one 256-bit load = one memory request (LDG.E.ENL2.256 on sm_100a)
#if defined(SB200_HOST_EMU)
__device__ __forceinline__ OccBlk load_blk(const OccBlk* p) { return *p; }
__device__ __forceinline__ OccSup load_sup(const OccSup* p) { return *p; }
#elif defined(SB200_NO_LD256)
__device__ __forceinline__ OccBlk load_blk(const OccBlk* p) {
    const ulonglong2* q = reinterpret_cast<const ulonglong2*>(p);
    ulonglong2 a = __ldg(q), b = __ldg(q + 1);
    OccBlk r;
    r.p0 = a.x; r.p1 = a.y; r.p2 = b.x; r.ctr = b.y;
    return r;
}
__device__ __forceinline__ OccSup load_sup(const OccSup* p) {
    const uint4* q = reinterpret_cast<const uint4*>(p);
    uint4 a = __ldg(q), b = __ldg(q + 1);
    OccSup r;
    r.c[0] = a.x; r.c[1] = a.y; r.c[2] = a.z; r.c[3] = a.w; r.c[4] = b.x; r.c[5] = b.y; r.c[6] = b.z; r.c[7] = b.w;
    return r;
}
#else
__device__ __forceinline__ OccBlk load_blk(const OccBlk* p) {
    OccBlk r;
    asm volatile("ld.global.nc.v4.b64 {%0,%1,%2,%3}, [%4];" : "=l"(r.p0), "=l"(r.p1), "=l"(r.p2), "=l"(r.ctr) : "l"(p));
    return r;
}
__device__ __forceinline__ OccSup load_sup(const OccSup* p) {
    OccSup r;
    asm volatile("ld.global.nc.v8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(r.c[0]), "=r"(r.c[1]), "=r"(r.c[2]), "=r"(r.c[3]), "=r"(r.c[4]), "=r"(r.c[5]), "=r"(r.c[6]),
                   "=r"(r.c[7])
                 : "l"(p));
    return r;
}
#endif

// occurrences of rank s (1..5) among the first `o` rows of the block (o in 0..63)
__device__ __forceinline__ uint32_t blk_count(const OccBlk& b, uint32_t o, int s) {
    uint64_t m = (uint64_t{1} << o) - 1;
    uint64_t a = b.p0 & m, c1 = b.p1 & m, c2 = b.p2 & m;
    uint64_t sel;
    switch (s) {
        case 1: sel = a & ~c1 & ~c2; break;  // 001
        case 2: sel = c1 & ~a; break;        // 010
        case 3: sel = a & c1; break;         // 011
        case 4: sel = c2 & ~a; break;        // 100
        default: sel = c2 & a; break;        // 101
    }
    return __popcll(sel);
}

__device__ __forceinline__ uint32_t blk_ctr(const OccBlk& b, int s) {  // s in 1..5
    return static_cast<uint32_t>(b.ctr >> (12 * (s - 1))) & 0xfffu;
}

// symbol code at row offset o
__device__ __forceinline__ int blk_symbol(const OccBlk& b, uint32_t o) {
    return static_cast<int>(((b.p0 >> o) & 1) | (((b.p1 >> o) & 1) << 1) | (((b.p2 >> o) & 1) << 2));
}

// ranks of symbols 1..SIGMA-1 at row i into r[1..], r[0] = rank of the delimiter
template <int SIGMA>
__device__ __forceinline__ void all_ranks(const OccTable& t, uint32_t i, uint32_t* r) {
    OccBlk b = load_blk(t.blk + (i >> kBlkShift));
    OccSup s = load_sup(t.sup + (i >> kSupShift));
    uint32_t o = i & 63u;
    uint32_t sum = 0;
#pragma unroll
    for (int c = 1; c < SIGMA; ++c) {
        r[c] = s.c[c] + blk_ctr(b, c) + blk_count(b, o, c);
        sum += r[c];
    }
    r[0] = i - sum;
}

}  // namespace sb200
