/*
 * nh_math.h — integer building blocks shared by the sm_100a kernels and the
 * host-side unit checks (tests/test_host_math.py compiles this file with g++).
 * Everything here is plain 32/64-bit integer arithmetic written so that the
 * same source runs on the device and on the host.
 *
 * Upstream units restated (kraken2 @ Dockerfile:15,35-38 of the reference;
 * spec in SURVEY.md Appendix A):
 *   nh_fmix64      <- kv_store.h MurmurHash3()                    (A.4)
 *   nh_revcomp     <- mmscanner.cc reverse_complement()           (A.3)
 *   nh_pack8       <- mmscanner.cc lookup_table_ (A/C/G/T -> 0..3) (A.2)
 *   nh_fastmod     <- `hc % capacity_` in compact_hash.cc Get()   (A.4)
 */
#ifndef NH_MATH_H
#define NH_MATH_H

#include <stdint.h>

#ifdef __CUDACC__
#define NH_HD __host__ __device__ __forceinline__
#else
#define NH_HD static inline
#endif

NH_HD uint64_t nh_fmix64(uint64_t k) {
  k ^= k >> 33;
  k *= 0xff51afd7ed558ccdULL;
  k ^= k >> 33;
  k *= 0xc4ceb9fe1a85ec53ULL;
  k ^= k >> 33;
  return k;
}

NH_HD uint32_t nh_brev32(uint32_t x) {
#ifdef __CUDA_ARCH__
  return __brev(x);
#else
  x = ((x >> 1) & 0x55555555u) | ((x & 0x55555555u) << 1);
  x = ((x >> 2) & 0x33333333u) | ((x & 0x33333333u) << 2);
  x = ((x >> 4) & 0x0F0F0F0Fu) | ((x & 0x0F0F0F0Fu) << 4);
  x = ((x >> 8) & 0x00FF00FFu) | ((x & 0x00FF00FFu) << 8);
  return (x >> 16) | (x << 16);
#endif
}

/* Reverse complement of an l-mer held in the low 2l bits.  A full 64-bit bit
 * reversal followed by a swap inside every 2-bit group equals kraken2's
 * reversal of 2-bit groups; then complement and (revcom_version 1) shift the
 * 2l significant bits back down. */
NH_HD uint64_t nh_revcomp(uint64_t x, int l, int revcom_version) {
  uint32_t lo = (uint32_t)x, hi = (uint32_t)(x >> 32);
  uint32_t rlo = nh_brev32(hi), rhi = nh_brev32(lo);
  rlo = ((rlo >> 1) & 0x55555555u) | ((rlo & 0x55555555u) << 1);
  rhi = ((rhi >> 1) & 0x55555555u) | ((rhi & 0x55555555u) << 1);
  uint64_t r = ~(((uint64_t)rhi << 32) | rlo);
  uint64_t mask = (1ULL << (2 * l)) - 1;
  if (revcom_version == 0) return r & mask; /* pre-2.0.8 DBs */
  return (r >> (64 - 2 * l)) & mask;
}

/* Four ASCII bases (memory order = byte 0 first) -> 8 bits of 2-bit codes,
 * first base in the two most significant bits, plus a 4-bit ambiguity mask
 * (bit i = byte i is not one of ACGTacgt).  Ambiguous bases get an arbitrary
 * code; the caller masks them out through the ambiguity bitmap. */
NH_HD uint32_t nh_pack4(uint32_t v, uint32_t *amb4) {
  uint32_t c = ((v >> 1) ^ (v >> 2)) & 0x03030303u;
  uint32_t u = v & 0xDFDFDFDFu; /* fold lower case */
  uint32_t b0 = c & 0x01010101u;
  uint32_t b1 = (c >> 1) & 0x01010101u;
  uint32_t t = b0 & b1;
  /* expected upper-case letter for the code: A 0x41, C 0x43, G 0x47, T 0x54 */
  uint32_t expect = 0x40404040u | (t ^ 0x01010101u) | ((b0 ^ b1) << 1) | (b1 << 2) | (t << 4);
  uint32_t diff = u ^ expect;
  uint32_t nz = (diff | ((diff & 0x7F7F7F7Fu) + 0x7F7F7F7Fu)) & 0x80808080u;
  *amb4 = (((nz >> 7) * 0x01020408u) >> 24) & 0xFu;
  return (c * 0x40100401u) >> 24;
}

/* Exact a % d for an invariant 64-bit divisor (Granlund & Montgomery 1994,
 * unsigned case, N = 64): q = (t + ((a - t) >> sh1)) >> sh2 with
 * t = mulhi(m, a); valid for 1 <= d < 2^63. */
typedef struct {
  uint64_t d;
  uint64_t m;
  uint32_t sh1, sh2;
} nh_divisor;

static inline nh_divisor nh_make_divisor(uint64_t d) {
  nh_divisor r;
  r.d = d;
  uint32_t l = 0; /* l = ceil(log2 d), d < 2^63 */
  while (l < 63 && (1ULL << l) < d) l++;
  unsigned __int128 num = ((unsigned __int128)((1ULL << l) - d)) << 64;
  r.m = (uint64_t)(num / d) + 1;
  r.sh1 = l < 1 ? l : 1;
  r.sh2 = l > 1 ? l - 1 : 0;
  return r;
}

NH_HD uint64_t nh_mulhi64(uint64_t a, uint64_t b) {
#ifdef __CUDA_ARCH__
  return __umul64hi(a, b);
#else
  return (uint64_t)(((unsigned __int128)a * b) >> 64);
#endif
}

NH_HD uint64_t nh_fastmod(uint64_t a, uint64_t d, uint64_t m, uint32_t sh1, uint32_t sh2) {
  uint64_t t = nh_mulhi64(m, a);
  uint64_t q = (t + ((a - t) >> sh1)) >> sh2;
  return a - q * d;
}

/* l-mer starting at base index t of a stream packed 16 bases per 32-bit word,
 * first base of a word in its two most significant bits. */
NH_HD uint64_t nh_extract_lmer(const uint32_t *w, uint32_t t, int l) {
  uint32_t j = t >> 4, o2 = (t & 15u) * 2u;
  uint32_t w0 = w[j], w1 = w[j + 1], w2 = w[j + 2];
#ifdef __CUDA_ARCH__
  uint32_t h = __funnelshift_l(w1, w0, o2);
  uint32_t m = __funnelshift_l(w2, w1, o2);
#else
  uint32_t h = o2 ? (w0 << o2) | (w1 >> (32 - o2)) : w0;
  uint32_t m = o2 ? (w1 << o2) | (w2 >> (32 - o2)) : w1;
#endif
  return (((uint64_t)h << 32) | m) >> (64 - 2 * l);
}

#endif
