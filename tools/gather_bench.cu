// gather_bench.cu — random dependent-gather microbenchmark for sm_100a.
//
// Purpose: pick the occurrence-table block granularity (32 / 64 / 128 bytes per probed BWT row) and
// the load method (thread-per-probe 128-bit loads, 256-bit loads, lane-group cooperative loads) for
// the rank kernel.  Each "chain" mimics one cursor: the next block index depends on the data just
// loaded, exactly like an FM-index extension depends on the previous rank.  ILP independent chains
// per thread mimic the two rows (lb, lb+len) probed per extension.
//
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo tools/gather_bench.cu -o build/gather_bench
// Run:   build/gather_bench [table_MiB ...]
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { \
  fprintf(stderr, "CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(1); } } while (0)

__device__ __forceinline__ uint32_t mix32(uint32_t x) {
  x ^= x >> 16; x *= 0x7feb352dU; x ^= x >> 15; x *= 0x846ca68bU; x ^= x >> 16; return x;
}

__global__ void fill_kernel(uint4* t, uint64_t n16) {
  uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
  uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
  for (; i < n16; i += stride) {
    uint32_t a = mix32((uint32_t)i * 4u + 1u), b = mix32((uint32_t)i * 4u + 2u);
    t[i] = make_uint4(a, b, mix32(a ^ 0x9e3779b9u), mix32(b + 0x85ebca6bu));
  }
}

__device__ __forceinline__ uint4 ldg128(const void* p) {
  uint4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
               : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
  return r;
}
struct U8x { uint32_t v[8]; };
__device__ __forceinline__ U8x ldg256(const void* p) {
  U8x r;
  asm volatile("ld.global.nc.L1::no_allocate.v8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=r"(r.v[0]), "=r"(r.v[1]), "=r"(r.v[2]), "=r"(r.v[3]),
                 "=r"(r.v[4]), "=r"(r.v[5]), "=r"(r.v[6]), "=r"(r.v[7]) : "l"(p));
  return r;
}

// MODE 0: one thread loads G bytes with 128-bit loads.  MODE 1: with 256-bit loads.
// MODE 2: a group of G/16 lanes loads one block cooperatively (16 B per lane) and shuffles the sum.
template <int G, int MODE, int ILP>
__global__ void __launch_bounds__(256) gather_kernel(const uint8_t* __restrict__ table, uint32_t nblk,
                                                     uint32_t iters, uint32_t* __restrict__ sink) {
  constexpr int LPG = (MODE == 2) ? (G / 16) : 1;  // lanes per group
  uint32_t tid = blockIdx.x * blockDim.x + threadIdx.x;
  uint32_t chain = tid / LPG;
  uint32_t sub = tid % LPG;
  uint32_t idx[ILP];
#pragma unroll
  for (int j = 0; j < ILP; ++j) idx[j] = (uint32_t)(((uint64_t)mix32(chain * ILP + j + 12345u) * nblk) >> 32);
  uint32_t acc = 0;
  for (uint32_t it = 0; it < iters; ++it) {
    uint32_t s[ILP];
    if constexpr (MODE == 0) {
      uint4 r[ILP][G / 16];
#pragma unroll
      for (int j = 0; j < ILP; ++j)
#pragma unroll
        for (int q = 0; q < G / 16; ++q) r[j][q] = ldg128(table + (uint64_t)idx[j] * G + q * 16);
#pragma unroll
      for (int j = 0; j < ILP; ++j) {
        s[j] = 0;
#pragma unroll
        for (int q = 0; q < G / 16; ++q) s[j] += __popc(r[j][q].x) + __popc(r[j][q].y) + r[j][q].z + r[j][q].w;
      }
    } else if constexpr (MODE == 1) {
      U8x r[ILP][G / 32];
#pragma unroll
      for (int j = 0; j < ILP; ++j)
#pragma unroll
        for (int q = 0; q < G / 32; ++q) r[j][q] = ldg256(table + (uint64_t)idx[j] * G + q * 32);
#pragma unroll
      for (int j = 0; j < ILP; ++j) {
        s[j] = 0;
#pragma unroll
        for (int q = 0; q < G / 32; ++q)
          s[j] += __popc(r[j][q].v[0]) + __popc(r[j][q].v[1]) + r[j][q].v[2] + r[j][q].v[3] +
                  __popc(r[j][q].v[4]) + __popc(r[j][q].v[5]) + r[j][q].v[6] + r[j][q].v[7];
      }
    } else {
      uint4 r[ILP];
#pragma unroll
      for (int j = 0; j < ILP; ++j) r[j] = ldg128(table + (uint64_t)idx[j] * G + sub * 16);
#pragma unroll
      for (int j = 0; j < ILP; ++j) {
        uint32_t v = __popc(r[j].x) + __popc(r[j].y) + r[j].z + r[j].w;
#pragma unroll
        for (int o = LPG / 2; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        s[j] = v;
      }
    }
#pragma unroll
    for (int j = 0; j < ILP; ++j) {
      acc += s[j];
      idx[j] = (uint32_t)(((uint64_t)mix32(s[j] + idx[j] * 0x9e3779b1u + it) * nblk) >> 32);
    }
  }
  if (acc == 0x12345678u) sink[0] = acc;  // keep the loads alive
}

template <int G, int MODE, int ILP>
static void run(const uint8_t* table, uint64_t bytes, uint32_t* sink, int sms, int blocks_per_sm, uint32_t iters) {
  uint32_t nblk = (uint32_t)(bytes / G);
  int maxb = 0;
  CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&maxb, gather_kernel<G, MODE, ILP>, 256, 0));
  if (blocks_per_sm > maxb) blocks_per_sm = maxb;
  int grid = sms * blocks_per_sm;
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  gather_kernel<G, MODE, ILP><<<grid, 256>>>(table, nblk, iters / 4 + 1, sink);  // warm-up
  CK(cudaDeviceSynchronize());
  float best = 1e30f;
  for (int rep = 0; rep < 3; ++rep) {
    CK(cudaEventRecord(e0));
    gather_kernel<G, MODE, ILP><<<grid, 256>>>(table, nblk, iters, sink);
    CK(cudaEventRecord(e1));
    CK(cudaEventSynchronize(e1));
    float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
    if (ms < best) best = ms;
  }
  constexpr int LPG = (MODE == 2) ? (G / 16) : 1;
  double probes = (double)grid * 256 / LPG * ILP * iters;
  double gps = probes / (best * 1e-3) / 1e9;
  printf("G=%3d mode=%d ilp=%d thr/SM=%4d table=%6.0fMiB  %8.2f Gprobe/s  %8.1f GB/s  (%.3f ms)\n", G, MODE, ILP,
         blocks_per_sm * 256, bytes / 1048576.0, gps, gps * G, best);
  fflush(stdout);
  CK(cudaEventDestroy(e0)); CK(cudaEventDestroy(e1));
}

int main(int argc, char** argv) {
  std::vector<uint64_t> sizes;
  for (int i = 1; i < argc; ++i) sizes.push_back(strtoull(argv[i], nullptr, 10));
  if (sizes.empty()) sizes = {64, 512, 3072};
  if (const char* g = getenv("L2_GRAN")) {
    CK(cudaDeviceSetLimit(cudaLimitMaxL2FetchGranularity, (size_t)atoi(g)));
  }
  size_t gran = 0;
  CK(cudaDeviceGetLimit(&gran, cudaLimitMaxL2FetchGranularity));
  printf("L2 fetch granularity limit: %zu\n", gran);
  cudaDeviceProp prop;
  CK(cudaGetDeviceProperties(&prop, 0));
  int sms = prop.multiProcessorCount;
  printf("device %s, %d SMs, L2 %d MiB\n", prop.name, sms, prop.l2CacheSize >> 20);
  uint64_t maxb = 0;
  for (auto s : sizes) if (s > maxb) maxb = s;
  maxb <<= 20;
  uint8_t* table; uint32_t* sink;
  CK(cudaMalloc(&table, maxb)); CK(cudaMalloc(&sink, 4));
  fill_kernel<<<sms * 8, 256>>>((uint4*)table, maxb / 16);
  CK(cudaDeviceSynchronize());
  const uint32_t iters = 200;
  for (auto s : sizes) {
    uint64_t bytes = s << 20;
    for (int bps : {8}) {
      run<32, 1, 1>(table, bytes, sink, sms, bps, iters);
      run<32, 1, 2>(table, bytes, sink, sms, bps, iters);
      run<64, 1, 1>(table, bytes, sink, sms, bps, iters);
      run<64, 2, 2>(table, bytes, sink, sms, bps, iters);
      run<128, 2, 2>(table, bytes, sink, sms, bps, iters);
    }
  }
  return 0;
}
