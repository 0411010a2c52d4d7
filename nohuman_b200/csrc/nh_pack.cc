/*
 * nh_pack.cc — host side of the packed transfer format (include/nohuman_gpu.h, nh_pack_reads):
 * ASCII bases -> 2-bit codes (4 bases per byte, first base in the top bits, the layout nh_pack4
 * produces in the kernel) + 1 validity bit per base, every sequence starting on a unit of 32 bases.
 * 0.4 bytes per base cross PCIe instead of 1.  What kraken2 does per base in
 * MinimizerScanner (mmscanner.cc lookup table: A/C/G/T in either case -> 0..3, anything else
 * ambiguous) is the whole arithmetic here.
 */
#include <immintrin.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include <thread>
#include <vector>

#include "../../include/nohuman_gpu.h"
#include "nh_internal.h"

namespace {

/* Software prefetch two pages ahead of the unit being packed: a core's demand misses alone reach about
 * 8 GB/s of host DRAM, with the prefetch 12-14 (tools/pack_bench.cc, PF_HINT / PF_DIST). */
constexpr uint64_t NH_PACK_PREFETCH = 8192;

/* one unit of 32 bases, scalar */
inline void pack_unit_scalar(const uint8_t *in, uint32_t n /* 1..32 valid input bytes */, uint8_t *codes8, uint32_t *valid) {
  uint32_t v = 0;
  memset(codes8, 0, 8);
  for (uint32_t j = 0; j < n; j++) {
    const uint8_t c = in[j], u = c & 0xDFu;
    const uint32_t code = ((c >> 1) ^ (c >> 2)) & 3u; /* A 0, C 1, G 2, T 3 */
    codes8[j >> 2] |= (uint8_t)(code << (6u - 2u * (j & 3u)));
    if (u == 'A' || u == 'C' || u == 'G' || u == 'T') v |= 1u << j;
  }
  *valid = v;
}

/* The loops exist twice: compiled for AVX2 (everything inlined, constants in registers across units) and
 * plain.  UNIT32(in, codes8, valid, keep) packs 32 bases that may be read in full; keep masks a partial unit. */
#define NH_PACK_LOOPS(SUFFIX, UNIT32)                                                                                          \
  void pack_range_##SUFFIX(const uint8_t *bases, const uint64_t *offsets, uint64_t total_bases, uint64_t s0, uint64_t s1,      \
                           uint8_t *codes, uint32_t *valid, const uint32_t *poff) {                                            \
    for (uint64_t s = s0; s < s1; s++) {                                                                                       \
      const uint8_t *in = bases + offsets[s];                                                                                  \
      const uint64_t len = offsets[s + 1] - offsets[s];                                                                        \
      uint8_t *c = codes + (uint64_t)poff[s] * 8;                                                                              \
      uint32_t *v = valid + poff[s];                                                                                           \
      uint64_t j = 0;                                                                                                          \
      for (; j + 32 <= len; j += 32) {                                                                                         \
        _mm_prefetch((const char *)(in + j + NH_PACK_PREFETCH), _MM_HINT_T0);                                                  \
        UNIT32(in + j, c + (j >> 2), v + (j >> 5), 0xFFFFFFFFu);                                                               \
      }                                                                                                                        \
      if (j < len) {                                                                                                           \
        const uint32_t rem = (uint32_t)(len - j);                                                                              \
        if (offsets[s] + j + 32 <= total_bases) { /* reading into the next sequence is harmless: its bits are masked */        \
          UNIT32(in + j, c + (j >> 2), v + (j >> 5), (1u << rem) - 1u);                                                        \
        } else {                                                                                                               \
          pack_unit_scalar(in + j, rem, c + (j >> 2), v + (j >> 5));                                                           \
        }                                                                                                                      \
      }                                                                                                                        \
    }                                                                                                                          \
  }

#define NH_UNIT32_SCALAR(in, c, v, keep) \
  do {                                   \
    pack_unit_scalar((in), 32, (c), (v)); \
    *(v) &= (keep);                      \
  } while (0)
NH_PACK_LOOPS(plain, NH_UNIT32_SCALAR)

#pragma GCC push_options
#pragma GCC target("avx2")
template <bool STREAM> /* STREAM: non-temporal stores — the whole-batch packer's output is read next by the copy engine, not by a core */
inline __attribute__((always_inline)) void pack_unit_avx2_t(const uint8_t *in, uint8_t *codes8, uint32_t *valid, uint32_t keep) {
  const __m256i v = _mm256_loadu_si256((const __m256i *)in);
  const __m256i s1 = _mm256_and_si256(_mm256_srli_epi16(v, 1), _mm256_set1_epi8(0x7F));
  const __m256i s2 = _mm256_and_si256(_mm256_srli_epi16(v, 2), _mm256_set1_epi8(0x3F));
  const __m256i code = _mm256_and_si256(_mm256_xor_si256(s1, s2), _mm256_set1_epi8(3));
  const __m256i letters = _mm256_setr_epi8('A', 'C', 'G', 'T', 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 'A', 'C', 'G', 'T', 0, 0, 0, 0, 0, 0, 0,
                                           0, 0, 0, 0, 0);
  const __m256i expect = _mm256_shuffle_epi8(letters, code);
  const __m256i ok = _mm256_cmpeq_epi8(_mm256_and_si256(v, _mm256_set1_epi8((char)0xDF)), expect);
  const uint32_t okbits = (uint32_t)_mm256_movemask_epi8(ok) & keep; /* keep: the bases of a last, partial unit */
  const __m256i p16 = _mm256_maddubs_epi16(code, _mm256_set1_epi16(0x0104)); /* c0*4 + c1 per 16-bit lane */
  const __m256i p32 = _mm256_madd_epi16(p16, _mm256_set1_epi32(0x00010010)); /* (c0*4+c1)*16 + (c2*4+c3) */
  const __m256i gather = _mm256_setr_epi8(0, 4, 8, 12, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1, 0, 4, 8, 12, -1, -1, -1, -1, -1, -1, -1,
                                          -1, -1, -1, -1, -1);
  const __m256i gq = _mm256_shuffle_epi8(p32, gather);
  const uint64_t both = (uint64_t)(uint32_t)_mm256_extract_epi32(gq, 0) | (uint64_t)(uint32_t)_mm256_extract_epi32(gq, 4) << 32;
  if (STREAM) {
    _mm_stream_si64((long long *)codes8, (long long)both);
    _mm_stream_si32((int *)valid, (int)okbits);
  } else {
    memcpy(codes8, &both, 8);
    *valid = okbits;
  }
}
#define NH_UNIT32_AVX2(in, c, v, keep) pack_unit_avx2_t<false>((in), (c), (v), (keep))
#define NH_UNIT32_AVX2_NT(in, c, v, keep) pack_unit_avx2_t<true>((in), (c), (v), (keep))
NH_PACK_LOOPS(avx2, NH_UNIT32_AVX2)
NH_PACK_LOOPS(avx2_nt, NH_UNIT32_AVX2_NT)
#pragma GCC pop_options

const bool g_avx2 = __builtin_cpu_supports("avx2");
const bool g_no_stream = getenv("NH_PACK_NO_STREAM") != nullptr; /* A/B switch: cached stores in nh_pack_reads */

}  // namespace

void nh_pack_range(const uint8_t *bases, const uint64_t *offsets, uint64_t total_bases, uint64_t s0, uint64_t s1, uint8_t *codes,
                   uint32_t *valid, const uint32_t *poff) {
  if (g_avx2 && !g_no_stream) {
    pack_range_avx2_nt(bases, offsets, total_bases, s0, s1, codes, valid, poff);
    _mm_sfence();
  } else if (g_avx2) {
    pack_range_avx2(bases, offsets, total_bases, s0, s1, codes, valid, poff);
  } else {
    pack_range_plain(bases, offsets, total_bases, s0, s1, codes, valid, poff);
  }
}

extern "C" int nh_packed_units(const uint64_t *offsets, uint64_t n_seqs, uint64_t *out_units) {
  if (!offsets || !out_units) return nh_set_error(NH_ERR_INVALID, "null argument");
  uint64_t u = 0;
  for (uint64_t s = 0; s < n_seqs; s++) u += (offsets[s + 1] - offsets[s] + 31) >> 5;
  *out_units = u;
  return NH_OK;
}

extern "C" int nh_pack_reads(const uint8_t *bases, const uint64_t *offsets, uint64_t n_seqs, uint8_t *codes, uint32_t *valid,
                             uint32_t *poff, int threads) {
  if (!offsets || !codes || !valid || !poff || (!bases && n_seqs && offsets[n_seqs] > offsets[0]))
    return nh_set_error(NH_ERR_INVALID, "null argument");
  if (threads < 1) threads = 1;
  if ((uint64_t)threads > n_seqs / 1024 + 1) threads = (int)(n_seqs / 1024 + 1);
  const uint64_t total = n_seqs ? offsets[n_seqs] : 0;
  /* two phases: every thread sums the units of its range of sequences, then, knowing where its range
   * starts, writes poff and packs */
  std::vector<uint64_t> part((size_t)threads + 1, 0);
  std::vector<std::thread> th;
  auto range = [&](int t, uint64_t *a, uint64_t *b) {
    *a = n_seqs * (uint64_t)t / (uint64_t)threads;
    *b = n_seqs * (uint64_t)(t + 1) / (uint64_t)threads;
  };
  for (int t = 0; t < threads; t++)
    th.emplace_back([&, t] {
      uint64_t a, b, u = 0;
      range(t, &a, &b);
      for (uint64_t s = a; s < b; s++) u += (offsets[s + 1] - offsets[s] + 31) >> 5;
      part[(size_t)t + 1] = u;
    });
  for (auto &x : th) x.join();
  th.clear();
  for (int t = 0; t < threads; t++) part[(size_t)t + 1] += part[(size_t)t];
  if (part[(size_t)threads] > 0xFFFFFFFFull) return nh_set_error(NH_ERR_CAPACITY, "batch too large for 32-bit unit offsets");
  for (int t = 0; t < threads; t++)
    th.emplace_back([&, t] {
      uint64_t a, b;
      range(t, &a, &b);
      uint64_t u = part[(size_t)t];
      for (uint64_t s = a; s < b; s++) {
        poff[s] = (uint32_t)u;
        u += (offsets[s + 1] - offsets[s] + 31) >> 5;
      }
      nh_pack_range(bases, offsets, total, a, b, codes, valid, poff);
    });
  for (auto &x : th) x.join();
  poff[n_seqs] = (uint32_t)part[(size_t)threads];
  return NH_OK;
}
