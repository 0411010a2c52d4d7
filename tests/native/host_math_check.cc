// Host-side check of nohuman_b200/csrc/nh_math.h against the oracle (tests only).
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <random>
#include "../../nohuman_b200/csrc/nh_math.h"
extern "C" {
#include "../../oracle/k2_oracle.h"
}

static int fails = 0;
#define CHECK(c, ...) do { if (!(c)) { if (fails < 10) { printf("FAIL %s:%d ", __FILE__, __LINE__); printf(__VA_ARGS__); printf("\n"); } fails++; } } while (0)

int main() {
  std::mt19937_64 rng(12345);
  // fmix64
  for (int i = 0; i < 100000; i++) { uint64_t x = rng(); CHECK(nh_fmix64(x) == k2o_fmix64(x), "fmix"); }
  CHECK(nh_fmix64(1) == 0xb456bcfc34c2cb2cULL, "fmix kat");
  // revcomp for all l, both versions
  for (int l = 1; l <= 31; l++)
    for (int i = 0; i < 20000; i++) {
      uint64_t x = rng() & ((1ULL << (2 * l)) - 1);
      CHECK(nh_revcomp(x, l, 1) == k2o_reverse_complement(x, l, 1), "revcomp v1 l=%d", l);
      CHECK(nh_revcomp(x, l, 0) == k2o_reverse_complement(x, l, 0), "revcomp v0 l=%d", l);
    }
  // pack4 on every byte value in every lane
  const char* valid = "ACGTacgt";
  for (int pos = 0; pos < 4; pos++)
    for (int b = 0; b < 256; b++) {
      uint8_t bytes[4] = {'A', 'c', 'G', 't'};
      bytes[pos] = (uint8_t)b;
      uint32_t v; memcpy(&v, bytes, 4);
      uint32_t amb; uint32_t p = nh_pack4(v, &amb);
      for (int i = 0; i < 4; i++) {
        bool ok = bytes[i] && strchr(valid, bytes[i]) != nullptr;
        CHECK(((amb >> i) & 1) == (ok ? 0u : 1u), "amb pos=%d b=%d i=%d amb=%x", pos, b, i, amb);
        if (ok) {
          int code = (bytes[i] | 0x20) == 'a' ? 0 : (bytes[i] | 0x20) == 'c' ? 1 : (bytes[i] | 0x20) == 'g' ? 2 : 3;
          CHECK(((p >> (6 - 2 * i)) & 3) == (uint32_t)code, "code pos=%d b=%d i=%d", pos, b, i);
        }
      }
    }
  // fastmod
  uint64_t ds[] = {1, 2, 3, 5, 7, 8, 1000003, (1ULL << 31), (1ULL << 31) - 1, (1ULL << 31) + 11, 3000000019ULL,
                   (1ULL << 33) + 7, (1ULL << 40) - 87, (1ULL << 62) + 3, (1ULL << 63) - 25, 0xFFFFFFFFULL, 0x100000000ULL};
  for (uint64_t d : ds) {
    nh_divisor dv = nh_make_divisor(d);
    uint64_t edge[] = {0, 1, d - 1, d, d + 1, 2 * d - 1, 2 * d, ~0ULL, ~0ULL - 1, (~0ULL / d) * d, (~0ULL / d) * d - 1};
    for (uint64_t a : edge) CHECK(nh_fastmod(a, d, dv.m, dv.sh1, dv.sh2) == a % d, "fastmod d=%llu a=%llu", (unsigned long long)d, (unsigned long long)a);
    for (int i = 0; i < 200000; i++) { uint64_t a = rng(); CHECK(nh_fastmod(a, d, dv.m, dv.sh1, dv.sh2) == a % d, "fastmod d=%llu a=%llu", (unsigned long long)d, (unsigned long long)a); }
  }
  for (int i = 0; i < 200000; i++) {
    uint64_t d = (rng() >> (rng() % 62)) | 1; if (d >> 63) d >>= 1;
    nh_divisor dv = nh_make_divisor(d);
    uint64_t a = rng();
    CHECK(nh_fastmod(a, d, dv.m, dv.sh1, dv.sh2) == a % d, "fastmod rnd d=%llu a=%llu", (unsigned long long)d, (unsigned long long)a);
  }
  // extract_lmer vs rolling
  for (int l = 1; l <= 31; l++) {
    const int N = 200;
    uint8_t codes[N]; for (int i = 0; i < N; i++) codes[i] = rng() & 3;
    uint32_t w[16] = {0};
    for (int i = 0; i < N; i++) w[i >> 4] |= (uint32_t)codes[i] << (30 - 2 * (i & 15));
    for (int t = 0; t + l <= N; t++) {
      uint64_t x = 0; for (int i = 0; i < l; i++) x = (x << 2) | codes[t + i];
      CHECK(nh_extract_lmer(w, t, l) == x, "extract l=%d t=%d", l, t);
    }
  }
  printf(fails ? "FAILED %d\n" : "OK\n", fails);
  return fails != 0;
}
