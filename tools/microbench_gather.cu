// Random-access HBM microbenchmark: defines the roofline the hash probe is held to.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o microbench_gather microbench_gather.cu
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>

__device__ __forceinline__ uint64_t fmix64(uint64_t k) {
  k ^= k >> 33; k *= 0xff51afd7ed558ccdULL; k ^= k >> 33; k *= 0xc4ceb9fe1a85ec53ULL; k ^= k >> 33; return k;
}
__device__ __forceinline__ void ld256(const uint32_t* p, uint32_t (&c)[8]) {
  asm volatile("ld.global.nc.L1::no_allocate.v8.u32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=r"(c[0]), "=r"(c[1]), "=r"(c[2]), "=r"(c[3]), "=r"(c[4]), "=r"(c[5]), "=r"(c[6]), "=r"(c[7]) : "l"(p));
}
__device__ __forceinline__ uint4 ld128(const uint32_t* p) {
  uint4 v; asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p)); return v;
}
__device__ __forceinline__ uint32_t ld32(const uint32_t* p) {
  uint32_t v; asm volatile("ld.global.nc.L1::no_allocate.u32 %0, [%1];" : "=r"(v) : "l"(p)); return v;
}

// MODE: bytes fetched per access (4, 16, 32, 64, 128); UNROLL independent accesses in flight per thread
template <int BYTES, int UNROLL>
__global__ void __launch_bounds__(256) k_gather(const uint32_t* __restrict__ t, uint64_t n_units, uint64_t n_reads, uint64_t seed, uint32_t* sink) {
  const uint64_t stride = (uint64_t)gridDim.x * blockDim.x * UNROLL;
  uint32_t acc = 0;
  for (uint64_t i = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) * UNROLL; i < n_reads; i += stride) {
    const uint32_t* p[UNROLL];
#pragma unroll
    for (int u = 0; u < UNROLL; u++) {
      uint64_t h = fmix64(i + u + seed);
      uint64_t unit = __umul64hi(h, n_units);
      p[u] = t + unit * (BYTES >= 32 ? BYTES / 4 : 8);
    }
#pragma unroll
    for (int u = 0; u < UNROLL; u++) {
      if (BYTES == 4) acc ^= ld32(p[u]);
      else if (BYTES == 16) { uint4 v = ld128(p[u]); acc ^= v.x ^ v.y ^ v.z ^ v.w; }
      else {
#pragma unroll
        for (int s = 0; s < BYTES / 32; s++) { uint32_t c[8]; ld256(p[u] + 8 * s, c); acc ^= c[0] ^ c[1] ^ c[2] ^ c[3] ^ c[4] ^ c[5] ^ c[6] ^ c[7]; }
      }
    }
  }
  if (acc == 0x9E3779B9u) sink[0] = acc;
}

template <int BYTES, int UNROLL>
double run(const uint32_t* t, uint64_t table_bytes, uint64_t n_reads, int blocks_per_sm, uint32_t* sink) {
  uint64_t unit_bytes = BYTES >= 32 ? BYTES : 32;
  uint64_t n_units = table_bytes / unit_bytes;
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  float best = 1e30f;
  for (int it = 0; it < 4; it++) {
    cudaEventRecord(e0);
    k_gather<BYTES, UNROLL><<<148 * blocks_per_sm, 256>>>(t, n_units, n_reads, 0x1234567ULL * (it + 1), sink);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    if (it > 0 && ms < best) best = ms;
  }
  return (double)n_reads / (best * 1e-3) / 1e9;  // G accesses / s
}

int main(int argc, char** argv) {
  size_t gran = 0; cudaDeviceGetLimit(&gran, cudaLimitMaxL2FetchGranularity);
  printf("default cudaLimitMaxL2FetchGranularity = %zu\n", gran);
  uint32_t* sink; cudaMalloc(&sink, 64);
  const uint64_t n_reads = 1ULL << 27;
  for (int g : {0, 32, 64, 128}) {
    if (g) { cudaError_t e = cudaDeviceSetLimit(cudaLimitMaxL2FetchGranularity, g); cudaDeviceGetLimit(&gran, cudaLimitMaxL2FetchGranularity); printf("== set granularity %d -> %s, now %zu\n", g, cudaGetErrorString(e), gran); }
    for (uint64_t gib : {1ULL, 8ULL, 32ULL}) {
      uint64_t bytes = gib << 30;
      uint32_t* t; if (cudaMalloc(&t, bytes) != cudaSuccess) { printf("alloc %llu GiB failed\n", (unsigned long long)gib); continue; }
      cudaMemset(t, 1, bytes);
      printf("table %2llu GiB:", (unsigned long long)gib);
      printf(" 4B u1 %.1f", run<4, 1>(t, bytes, n_reads, 8, sink));
      printf(" | 32B u1 %.1f u2 %.1f u4 %.1f", run<32, 1>(t, bytes, n_reads, 8, sink), run<32, 2>(t, bytes, n_reads, 8, sink), run<32, 4>(t, bytes, n_reads, 8, sink));
      printf(" | 32B u1 occ4 %.1f occ2 %.1f", run<32, 1>(t, bytes, n_reads, 4, sink), run<32, 1>(t, bytes, n_reads, 2, sink));
      printf(" | 64B u1 %.1f u2 %.1f", run<64, 1>(t, bytes, n_reads, 8, sink), run<64, 2>(t, bytes, n_reads, 8, sink));
      printf(" | 128B u1 %.1f  [G accesses/s]\n", run<128, 1>(t, bytes, n_reads, 8, sink));
      cudaFree(t);
    }
  }
  return 0;
}
