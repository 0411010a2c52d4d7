// Does a VMM allocation (cuMemCreate + 512 MiB / 1 GiB aligned VA) get larger GPU pages than cudaMalloc,
// i.e. does the random-access rate rise above the 37 G requests/s TLB-miss limit?
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o microbench_vmm microbench_vmm.cu -lcuda
#include <cuda.h>
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
__global__ void __launch_bounds__(256) k(const uint32_t* __restrict__ t, uint64_t n_sectors, uint64_t n_items, uint64_t seed, uint32_t* sink) {
  const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
  uint32_t acc = 0;
  for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_items; i += stride)
    acc ^= ld8(t + __umul64hi(fmix64(i + seed), n_sectors) * 8ULL);
  if (acc == 0x9E3779B9u) sink[0] = acc;
}
static double rate(const uint32_t* t, uint64_t bytes, uint32_t* sink) {
  const uint64_t n = 1ULL << 26;
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  float best = 1e30f;
  for (int it = 0; it < 3; it++) {
    cudaEventRecord(e0); k<<<148 * 8, 256>>>(t, bytes / 32, n, 77 * (it + 1), sink); cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1); if (it > 0 && ms < best) best = ms;
  }
  cudaError_t e = cudaGetLastError(); if (e != cudaSuccess) printf("  kernel error: %s\n", cudaGetErrorString(e));
  return n / (best * 1e-3) / 1e9;
}
#define CK(x) do { CUresult r_ = (x); if (r_ != CUDA_SUCCESS) { const char* s_; cuGetErrorString(r_, &s_); printf("  %s -> %s\n", #x, s_); return 1; } } while (0)
int main() {
  cudaFree(0);
  uint32_t* sink; cudaMalloc(&sink, 64);
  const uint64_t bytes = 8ULL << 30;
  { uint32_t* t; cudaMalloc(&t, bytes); cudaMemset(t, 1, bytes); printf("cudaMalloc (ptr %p): %.2f G requests/s\n", (void*)t, rate(t, bytes, sink)); cudaFree(t); }
  CUmemAllocationProp prop = {}; prop.type = CU_MEM_ALLOCATION_TYPE_PINNED; prop.location.type = CU_MEM_LOCATION_TYPE_DEVICE; prop.location.id = 0;
  size_t gmin = 0, grec = 0;
  CK(cuMemGetAllocationGranularity(&gmin, &prop, CU_MEM_ALLOC_GRANULARITY_MINIMUM));
  CK(cuMemGetAllocationGranularity(&grec, &prop, CU_MEM_ALLOC_GRANULARITY_RECOMMENDED));
  printf("VMM granularity: minimum %zu KiB, recommended %zu KiB\n", gmin >> 10, grec >> 10);
  for (uint64_t align : {(uint64_t)(2ULL << 20), (uint64_t)(32ULL << 20), (uint64_t)(512ULL << 20), (uint64_t)(1ULL << 30)}) {
    for (uint64_t chunk : {(uint64_t)(8ULL << 30), (uint64_t)(512ULL << 20)}) {   // one physical handle, or 512 MiB handles
      CUdeviceptr va = 0; CK(cuMemAddressReserve(&va, bytes, align, 0, 0));
      const int nchunks = (int)(bytes / chunk);
      CUmemGenericAllocationHandle h[64];
      for (int i = 0; i < nchunks; i++) { CK(cuMemCreate(&h[i], chunk, &prop, 0)); CK(cuMemMap(va + (uint64_t)i * chunk, chunk, 0, h[i], 0)); }
      CUmemAccessDesc ad = {}; ad.location = prop.location; ad.flags = CU_MEM_ACCESS_FLAGS_PROT_READWRITE;
      CK(cuMemSetAccess(va, bytes, &ad, 1));
      cudaMemset((void*)va, 1, bytes);
      printf("VMM va align %4llu MiB, %2d handle(s) (ptr %p): %.2f G requests/s\n", (unsigned long long)(align >> 20), nchunks, (void*)va, rate((const uint32_t*)va, bytes, sink));
      CK(cuMemUnmap(va, bytes)); for (int i = 0; i < nchunks; i++) CK(cuMemRelease(h[i])); CK(cuMemAddressFree(va, bytes));
    }
  }
  return 0;
}
