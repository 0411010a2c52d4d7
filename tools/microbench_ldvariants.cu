// Which load instruction gives the cheapest random 32-byte sector read on B200?
// Each variant is its own kernel so that `ncu` reports DRAM bytes / L2 sectors per variant.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o microbench_ldvariants microbench_ldvariants.cu
//   ./microbench_ldvariants [l2_fetch_granularity(0=default)] [table_gib]
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstdlib>

__device__ __forceinline__ uint64_t fmix64(uint64_t k) {
  k ^= k >> 33; k *= 0xff51afd7ed558ccdULL; k ^= k >> 33; k *= 0xc4ceb9fe1a85ec53ULL; k ^= k >> 33; return k;
}
#define LD8(name, op)                                                                        \
  __device__ __forceinline__ uint32_t name(const uint32_t* p) {                              \
    uint32_t c0, c1, c2, c3, c4, c5, c6, c7;                                                 \
    asm volatile(op " {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"                                      \
                 : "=r"(c0), "=r"(c1), "=r"(c2), "=r"(c3), "=r"(c4), "=r"(c5), "=r"(c6), "=r"(c7) : "l"(p)); \
    return c0 ^ c1 ^ c2 ^ c3 ^ c4 ^ c5 ^ c6 ^ c7;                                            \
  }
#define LD4(name, op)                                                                        \
  __device__ __forceinline__ uint32_t name(const uint32_t* p) {                              \
    uint32_t c0, c1, c2, c3;                                                                 \
    asm volatile(op " {%0,%1,%2,%3}, [%4];" : "=r"(c0), "=r"(c1), "=r"(c2), "=r"(c3) : "l"(p)); \
    return c0 ^ c1 ^ c2 ^ c3;                                                                \
  }
#define LD1(name, op)                                                                        \
  __device__ __forceinline__ uint32_t name(const uint32_t* p) {                              \
    uint32_t c0; asm volatile(op " %0, [%1];" : "=r"(c0) : "l"(p)); return c0; }

LD8(v8_nc_noalloc, "ld.global.nc.L1::no_allocate.v8.u32")
LD8(v8_plain, "ld.global.v8.u32")
LD8(v8_nc, "ld.global.nc.v8.u32")
LD8(v8_noalloc, "ld.global.L1::no_allocate.v8.u32")
LD8(v8_l2_64, "ld.global.L2::64B.v8.u32")
LD8(v8_evict_first, "ld.global.L1::evict_first.v8.u32")
LD4(v4_nc_noalloc, "ld.global.nc.L1::no_allocate.v4.u32")
LD4(v4_plain, "ld.global.v4.u32")
LD4(v4_cg, "ld.global.cg.v4.u32")
LD4(v4_cs, "ld.global.cs.v4.u32")
LD4(v4_cv, "ld.global.cv.v4.u32")
LD4(v4_lu, "ld.global.lu.v4.u32")
LD1(v1_plain, "ld.global.u32")
LD1(v1_cg, "ld.global.cg.u32")
LD1(v1_cv, "ld.global.cv.u32")

#define KERNEL(kname, expr)                                                                        \
  __global__ void __launch_bounds__(256) kname(const uint32_t* __restrict__ t, uint64_t n_sectors, \
                                               uint64_t n_reads, uint64_t seed, uint32_t* sink) {  \
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;                                      \
    uint32_t acc = 0;                                                                              \
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_reads; i += stride) { \
      const uint32_t* p = t + __umul64hi(fmix64(i + seed), n_sectors) * 8ULL;                      \
      acc ^= (expr);                                                                               \
    }                                                                                              \
    if (acc == 0x9E3779B9u) sink[0] = acc;                                                         \
  }

KERNEL(g_v8_nc_noalloc, v8_nc_noalloc(p))
KERNEL(g_v8_plain, v8_plain(p))
KERNEL(g_v8_nc, v8_nc(p))
KERNEL(g_v8_noalloc, v8_noalloc(p))
KERNEL(g_v8_l2_64, v8_l2_64(p))
KERNEL(g_v8_evict_first, v8_evict_first(p))
KERNEL(g_2xv4_nc_noalloc, v4_nc_noalloc(p) ^ v4_nc_noalloc(p + 4))
KERNEL(g_2xv4_plain, v4_plain(p) ^ v4_plain(p + 4))
KERNEL(g_2xv4_cg, v4_cg(p) ^ v4_cg(p + 4))
KERNEL(g_2xv4_cs, v4_cs(p) ^ v4_cs(p + 4))
KERNEL(g_2xv4_cv, v4_cv(p) ^ v4_cv(p + 4))
KERNEL(g_2xv4_lu, v4_lu(p) ^ v4_lu(p + 4))
KERNEL(g_1xv4_cg, v4_cg(p))
KERNEL(g_1xu32_plain, v1_plain(p))
KERNEL(g_1xu32_cg, v1_cg(p))
KERNEL(g_1xu32_cv, v1_cv(p))

typedef void (*kern_t)(const uint32_t*, uint64_t, uint64_t, uint64_t, uint32_t*);
struct V { const char* name; kern_t k; };

int main(int argc, char** argv) {
  int gran = argc > 1 ? atoi(argv[1]) : 0;
  uint64_t gib = argc > 2 ? strtoull(argv[2], 0, 10) : 8;
  if (gran) { cudaError_t e = cudaDeviceSetLimit(cudaLimitMaxL2FetchGranularity, gran); printf("set L2 fetch granularity %d: %s\n", gran, cudaGetErrorString(e)); }
  size_t g = 0; cudaDeviceGetLimit(&g, cudaLimitMaxL2FetchGranularity); printf("cudaLimitMaxL2FetchGranularity = %zu, table %llu GiB\n", g, (unsigned long long)gib);
  uint64_t bytes = gib << 30;
  uint32_t *t, *sink; cudaMalloc(&sink, 64);
  if (argc > 3) {  // size sweep of the v8.nc flavour: is the limit DRAM/TLB (drops past L2 / TLB reach) or the request path?
    cudaMalloc(&t, 16ULL << 30); cudaMemset(t, 1, 16ULL << 30);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (uint64_t mb : {16ULL, 32ULL, 64ULL, 96ULL, 128ULL, 192ULL, 256ULL, 384ULL, 512ULL, 1024ULL, 2048ULL, 4096ULL, 8192ULL, 16384ULL}) {
      float best = 1e30f;
      for (int it = 0; it < 4; it++) {
        cudaEventRecord(e0);
        g_v8_nc<<<148 * 8, 256>>>(t, (mb << 20) / 32, 1ULL << 27, 0x1234567ULL * (it + 1), sink);
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        if (it > 0 && ms < best) best = ms;
      }
      printf("window %6llu MiB: %6.2f G sectors/s\n", (unsigned long long)mb, (1ULL << 27) / (best * 1e-3) / 1e9);
    }
    return 0;
  }
  if (cudaMalloc(&t, bytes) != cudaSuccess) { printf("alloc failed\n"); return 1; }
  cudaMemset(t, 1, bytes);
  const uint64_t n_reads = 1ULL << 27, n_sectors = bytes / 32;
  V vs[] = {{"v8.nc.L1::no_allocate", g_v8_nc_noalloc}, {"v8 (plain)", g_v8_plain}, {"v8.nc", g_v8_nc},
            {"v8.L1::no_allocate", g_v8_noalloc}, {"v8.L2::64B", g_v8_l2_64}, {"v8.L1::evict_first", g_v8_evict_first},
            {"2xv4.nc.L1::no_allocate", g_2xv4_nc_noalloc}, {"2xv4 (plain)", g_2xv4_plain}, {"2xv4.cg", g_2xv4_cg},
            {"2xv4.cs", g_2xv4_cs}, {"2xv4.cv", g_2xv4_cv}, {"2xv4.lu", g_2xv4_lu}, {"1xv4.cg (16B)", g_1xv4_cg},
            {"1xu32 (plain)", g_1xu32_plain}, {"1xu32.cg", g_1xu32_cg}, {"1xu32.cv", g_1xu32_cv}};
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  for (auto& v : vs) {
    float best = 1e30f;
    for (int it = 0; it < 3; it++) {
      cudaEventRecord(e0);
      v.k<<<148 * 8, 256>>>(t, n_sectors, n_reads, 0x1234567ULL * (it + 1), sink);
      cudaEventRecord(e1); cudaEventSynchronize(e1);
      float ms; cudaEventElapsedTime(&ms, e0, e1);
      if (it > 0 && ms < best) best = ms;
    }
    cudaError_t e = cudaGetLastError();
    printf("%-26s %6.2f G sectors/s  %7.1f GB/s  %s\n", v.name, n_reads / (best * 1e-3) / 1e9, n_reads * 32.0 / (best * 1e-3) / 1e9, e == cudaSuccess ? "" : cudaGetErrorString(e));
  }
  return 0;
}
