// Is the TLB that limits random access per SM, per group of SMs, or shared by the GPU?
// Each group of `share` SMs (by %smid) reads randomly inside its own window of `win_mb` MiB.
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
__device__ __forceinline__ uint64_t fmix64(uint64_t k) {
  k ^= k >> 33; k *= 0xff51afd7ed558ccdULL; k ^= k >> 33; k *= 0xc4ceb9fe1a85ec53ULL; k ^= k >> 33; return k;
}
template <int H> __device__ __forceinline__ uint32_t ld8(const uint32_t* p);
template <> __device__ __forceinline__ uint32_t ld8<1>(const uint32_t* p) {
  uint32_t c0, c1, c2, c3, c4, c5, c6, c7;
  asm volatile("ld.global.L2::64B.v8.u32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=r"(c0), "=r"(c1), "=r"(c2), "=r"(c3), "=r"(c4), "=r"(c5), "=r"(c6), "=r"(c7) : "l"(p));
  return c0 ^ c1 ^ c2 ^ c3 ^ c4 ^ c5 ^ c6 ^ c7;
}
template <> __device__ __forceinline__ uint32_t ld8<0>(const uint32_t* p) {
  uint32_t c0, c1, c2, c3, c4, c5, c6, c7;
  asm volatile("ld.global.nc.L1::no_allocate.v8.u32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=r"(c0), "=r"(c1), "=r"(c2), "=r"(c3), "=r"(c4), "=r"(c5), "=r"(c6), "=r"(c7) : "l"(p));
  return c0 ^ c1 ^ c2 ^ c3 ^ c4 ^ c5 ^ c6 ^ c7;
}
template <int H>
__global__ void __launch_bounds__(256) k(const uint32_t* __restrict__ t, uint64_t win_sectors, uint32_t share, uint32_t n_windows,
                                         uint64_t per_thread, uint64_t seed, uint32_t* sink, uint64_t page_used_sectors) {
  uint32_t smid; asm("mov.u32 %0, %%smid;" : "=r"(smid));
  const uint64_t w = (smid / share) % n_windows;
  const uint32_t* base = t + w * win_sectors * 8ULL;
  uint32_t acc = 0;
  const uint64_t tid = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  for (uint64_t i = 0; i < per_thread; i++)
  {
    uint64_t s = __umul64hi(fmix64(tid * per_thread + i + seed), win_sectors);
    if (page_used_sectors) { const uint64_t n_pages = win_sectors / 65536; s = __umul64hi(fmix64(tid * per_thread + i + seed), n_pages * page_used_sectors); s = (s / page_used_sectors) * 65536 + s % page_used_sectors; }
    acc ^= ld8<H>(base + s * 8ULL);
  }
  if (acc == 0x9E3779B9u) sink[0] = acc;
}
int main() {
  const uint64_t bytes = 20ULL << 30; uint32_t *t, *sink; cudaMalloc(&t, bytes); cudaMalloc(&sink, 64); cudaMemset(t, 1, bytes);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  const uint64_t per_thread = 256, threads = 148ULL * 8 * 256, n = per_thread * threads;
  struct C { uint32_t share; uint64_t win_mb; int hint; uint64_t used_kib; } cs[] = {
      {148, 8192, 0, 0}, {148, 8192, 1, 0}, {1, 64, 0, 0}, {1, 64, 1, 0}, {1, 32, 1, 0}, {2, 128, 1, 0}, {4, 128, 1, 0}, {8, 128, 1, 0}, {16, 256, 1, 0},
      {148, 8192, 0, 16}, {148, 8192, 1, 16}, {1, 64, 0, 256}, {148, 16384, 1, 0}};
  for (auto& c : cs) {
    uint32_t n_windows = (148 + c.share - 1) / c.share;
    if ((uint64_t)n_windows * (c.win_mb << 20) > bytes) { printf("skip\n"); continue; }
    float best = 1e30f;
    for (int it = 0; it < 3; it++) {
      cudaEventRecord(e0);
      if (c.hint) k<1><<<148 * 8, 256>>>(t, (c.win_mb << 20) / 32, c.share, n_windows, per_thread, 77 * (it + 1), sink, c.used_kib * 32);
      else k<0><<<148 * 8, 256>>>(t, (c.win_mb << 20) / 32, c.share, n_windows, per_thread, 77 * (it + 1), sink, c.used_kib * 32);
      cudaEventRecord(e1); cudaEventSynchronize(e1);
      float ms; cudaEventElapsedTime(&ms, e0, e1); if (it > 0 && ms < best) best = ms;
    }
    printf("%3u SMs share a %5llu MiB window (%3u windows) %s %s: %7.2f G requests/s\n", c.share, (unsigned long long)c.win_mb, n_windows,
           c.hint ? "L2::64B" : "default", c.used_kib ? "sparse(first KiB of each 2 MiB page used)" : "dense", n / (best * 1e-3) / 1e9);
  }
  return 0;
}
