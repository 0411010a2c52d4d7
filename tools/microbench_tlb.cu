// Is the 37 G requests/s random-access limit a TLB (pages touched) or an L2-miss (bytes touched) limit?
// Same number of distinct bytes, spread over few or many 2 MiB pages.
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
__device__ __forceinline__ uint64_t fmix64(uint64_t k) {
  k ^= k >> 33; k *= 0xff51afd7ed558ccdULL; k ^= k >> 33; k *= 0xc4ceb9fe1a85ec53ULL; k ^= k >> 33; return k;
}
__device__ __forceinline__ uint32_t ld8(const uint32_t* p) {
  uint32_t c0, c1, c2, c3, c4, c5, c6, c7;
  asm volatile("ld.global.nc.L1::no_allocate.v8.u32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=r"(c0), "=r"(c1), "=r"(c2), "=r"(c3), "=r"(c4), "=r"(c5), "=r"(c6), "=r"(c7) : "l"(p));
  return c0 ^ c1 ^ c2 ^ c3 ^ c4 ^ c5 ^ c6 ^ c7;
}
// n_pages pages of `page_stride` bytes; only the first `used` bytes of each page are touched
__global__ void __launch_bounds__(256) k(const uint32_t* __restrict__ t, uint64_t n_pages, uint64_t page_stride,
                                         uint64_t used, uint64_t n_items, uint64_t seed, uint32_t* sink) {
  const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
  uint32_t acc = 0;
  const uint64_t sec_per_page = used / 32;
  for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_items; i += stride) {
    const uint64_t h = fmix64(i + seed);
    const uint64_t s = __umul64hi(h, n_pages * sec_per_page);
    const uint64_t page = s / sec_per_page, in = s % sec_per_page;
    acc ^= ld8(t + (page * page_stride + in * 32) / 4);
  }
  if (acc == 0x9E3779B9u) sink[0] = acc;
}
int main() {
  const uint64_t bytes = 16ULL << 30; uint32_t *t, *sink; cudaMalloc(&t, bytes); cudaMalloc(&sink, 64); cudaMemset(t, 1, bytes);
  const uint64_t n = 1ULL << 26;
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  struct C { uint64_t pages, stride, used; } cs[] = {
      {128, 2ULL << 20, 2ULL << 20},    // 256 MiB dense: 128 pages
      {4096, 2ULL << 20, 64ULL << 10},  // 256 MiB spread over 4096 pages (8 GiB span)
      {8192, 2ULL << 20, 32ULL << 10},  // 256 MiB spread over 8192 pages (16 GiB span)
      {512, 2ULL << 20, 2ULL << 20},    // 1 GiB dense
      {8192, 2ULL << 20, 128ULL << 10}, // 1 GiB spread over 8192 pages
      {32, 2ULL << 20, 2ULL << 20},     // 64 MiB dense (L2 resident)
      {4096, 2ULL << 20, 16ULL << 10},  // 64 MiB spread over 4096 pages
      {8192, 2ULL << 20, 8ULL << 10},   // 64 MiB spread over 8192 pages
      {16, 512ULL << 20, 4ULL << 20},   // 64 MiB in 16 x 4 MiB islands 512 MiB apart
      {4096, 2ULL << 20, 2ULL << 20},   // 8 GiB dense
  };
  for (auto& c : cs) {
    float best = 1e30f;
    for (int it = 0; it < 3; it++) {
      cudaEventRecord(e0); k<<<148 * 8, 256>>>(t, c.pages, c.stride, c.used, n, 77 * (it + 1), sink); cudaEventRecord(e1); cudaEventSynchronize(e1);
      float ms; cudaEventElapsedTime(&ms, e0, e1); if (it > 0 && ms < best) best = ms;
    }
    printf("%5llu pages x %7llu KiB used (stride %4llu MiB, %6llu MiB touched): %7.2f G requests/s\n", (unsigned long long)c.pages,
           (unsigned long long)(c.used >> 10), (unsigned long long)(c.stride >> 20), (unsigned long long)((c.pages * c.used) >> 20), n / (best * 1e-3) / 1e9);
  }
  return 0;
}
