/*
 * k2_synth.c — CPU twin of the synthetic-workload generator in
 * nohuman_b200/csrc/nh_synth.cu (same genome formula, same block -> leaf
 * assignment, same read sampler), written against the oracle's scanner and
 * table.  BENCH / TEST INFRASTRUCTURE ONLY, like everything under oracle/.
 *
 * Why it exists: `bench.py --impl reference` times the CPU implementation of
 * the path (the kraken2 restatement in k2_oracle.c) and must not load the
 * CUDA library at all.  With this file the reference arm builds its own
 * HPRC.r2-sized table and samples the very same reads on the host cores.
 * Nothing in the reference (mbhall88/nohuman) corresponds to it: the HPRC
 * databases it downloads (config.toml:1-19) are not available offline.
 *
 * The table is built as kraken2-build builds one (build_db.cc ProcessSequence:
 * every distinct minimizer, value := LCA(existing, taxon)) but
 * by all cores at once, with the compare-and-swap the upstream code also uses.
 */
#include <math.h>
#include <omp.h>
#include <stdlib.h>
#include <string.h>

#include "k2_oracle.h"

#define K2S_TILE_POS 124 /* the GPU builder assigns taxa per tile of this many k-mer positions */

static inline uint32_t genome_code(uint64_t seed, uint64_t i) {
  const uint64_t w = k2o_fmix64(seed + ((i >> 5) + 1ULL) * 0x9E3779B97F4A7C15ULL);
  return (uint32_t)(w >> (2u * (uint32_t)(i & 31ULL))) & 3u;
}

static inline uint64_t umulhi64(uint64_t a, uint64_t b) { return (uint64_t)(((unsigned __int128)a * b) >> 64); }

void k2s_genome(uint8_t *out, uint64_t start, uint64_t n, uint64_t seed, int threads) {
  if (threads <= 0) threads = omp_get_max_threads();
  const uint64_t end = start + n;
  /* one hash word carries 32 bases */
#pragma omp parallel for num_threads(threads) schedule(static)
  for (int64_t wi = (int64_t)(start >> 5); wi <= (int64_t)((end - 1) >> 5); wi++) {
    if (n == 0) continue;
    const uint64_t w = k2o_fmix64(seed + ((uint64_t)wi + 1ULL) * 0x9E3779B97F4A7C15ULL);
    const uint64_t lo = (uint64_t)wi << 5;
    const uint64_t a = lo < start ? start : lo, b = lo + 32 < end ? lo + 32 : end;
    for (uint64_t i = a; i < b; i++) out[i - start] = (uint8_t)"ACGT"[(w >> (2u * (uint32_t)(i & 31ULL))) & 3u];
  }
}

/* compact_hash.cc CompareAndSet, collapsed to value := LCA(old, taxon); returns 1 if an empty cell was claimed */
static int insert_lca_atomic(k2o_cht *t, const k2o_taxonomy *tax, uint64_t key, uint32_t taxon) {
  const uint64_t hc = k2o_fmix64(key);
  const uint32_t vmask = (uint32_t)((1ULL << t->value_bits) - 1);
  const uint32_t ckey = (uint32_t)(hc >> (32 + t->value_bits));
  uint64_t idx = hc % t->capacity;
  for (uint64_t probes = 0; probes < t->capacity; probes++) {
    uint32_t cell = __atomic_load_n(&t->cells[idx], __ATOMIC_RELAXED);
    if ((cell & vmask) == 0) {
      const uint32_t want = (ckey << t->value_bits) | taxon;
      uint32_t expected = 0;
      if (__atomic_compare_exchange_n(&t->cells[idx], &expected, want, 0, __ATOMIC_RELAXED, __ATOMIC_RELAXED)) return 1;
      cell = expected; /* somebody else claimed it: examine what they wrote */
    }
    if ((cell >> t->value_bits) == ckey) {
      for (;;) {
        const uint32_t cur = cell & vmask;
        const uint32_t nv = (uint32_t)k2o_lca(tax, cur, taxon);
        if (nv == cur) return 0;
        const uint32_t want = (ckey << t->value_bits) | nv;
        if (__atomic_compare_exchange_n(&t->cells[idx], &cell, want, 0, __ATOMIC_RELAXED, __ATOMIC_RELAXED)) return 0;
      }
    }
    if (++idx >= t->capacity) idx = 0;
  }
  return 0;
}

/* t: allocated and zeroed (k2o_cht_alloc).  leaf_taxa: INTERNAL ids.  Mirrors nh_synth_build_db. */
int k2s_build_db(k2o_cht *t, const k2o_taxonomy *tax, const k2o_index_options *o, const uint32_t *leaf_taxa,
                 int n_leaves, double target_load, uint64_t genome_seed, uint64_t block_bases, double overlap_frac,
                 uint64_t max_genome_bases, int threads, uint64_t *genome_bases) {
  if (!t || !tax || !o || !leaf_taxa || n_leaves < 1) return -1;
  if (threads <= 0) threads = omp_get_max_threads();
  const uint64_t k = o->k;
  uint64_t chunk_max = t->capacity / 8;
  if (chunk_max > (128ULL << 20)) chunk_max = 128ULL << 20;
  if (chunk_max < (1ULL << 16)) chunk_max = 1ULL << 16;
  const uint64_t target = (uint64_t)(target_load * (double)t->capacity);
  const uint64_t overlap_start = (uint64_t)((1.0 - (overlap_frac < 0 ? 0 : overlap_frac)) * (double)block_bases);
  uint8_t *buf = (uint8_t *)malloc(chunk_max);
  if (!buf) return -2;
  uint64_t gpos = 0, size = 0;
  int stalls = 0;
  while (size < target) {
    const uint64_t need = target - size;
    uint64_t n = need * 3 + 4096; /* ~1 new cell per 3 bases (window of 5 l-mers) */
    if (n > chunk_max) n = chunk_max;
    if (max_genome_bases && gpos + n > max_genome_bases) {
      if (gpos + k >= max_genome_bases) break;
      n = max_genome_bases - gpos;
    }
    k2s_genome(buf, gpos, n, genome_seed, threads);
    const uint64_t npos = n >= k ? n - k + 1 : 0;
    const uint64_t n_tiles = (npos + K2S_TILE_POS - 1) / K2S_TILE_POS;
    uint64_t claimed = 0;
    /* The synthetic genome has no ambiguous base, so the scan is the plain rolling form of
     * MinimizerScanner::NextMinimizer (mmscanner.cc): forward and reverse-complement
     * l-mer, canonical, spaced seed, toggle, minimum over the window of k-l+1 l-mers.  The oracle's
     * own scanner gives the same stream (tests/test_synth_host.py) at a fiftieth of the speed. */
    const int l = (int)o->l, w = (int)(o->k - o->l + 1);
    const uint64_t lmask = l < 32 ? (1ULL << (2 * l)) - 1ULL : ~0ULL;
    const uint64_t smask = o->spaced_seed_mask ? o->spaced_seed_mask : lmask;
    const uint64_t toggle = o->toggle_mask & lmask;
    const int rev0 = o->revcom_version == 0;
#pragma omp parallel num_threads(threads) reduction(+ : claimed)
    {
      /* a thread takes runs of 512 tiles */
#pragma omp for schedule(dynamic, 1)
      for (int64_t run = 0; run < (int64_t)((n_tiles + 511) / 512); run++) {
        const uint64_t t0 = (uint64_t)run * 512, t1 = t0 + 512 < n_tiles ? t0 + 512 : n_tiles;
        const uint64_t p0 = t0 * K2S_TILE_POS, p1 = t1 * K2S_TILE_POS < npos ? t1 * K2S_TILE_POS : npos;
        uint64_t fwd = 0, rc = 0, ring[32], last = ~0ULL;
        for (int i = 0; i < 32; i++) ring[i] = ~0ULL;
        const uint8_t *sq = buf + p0;
        const uint64_t nb = p1 - p0 + k - 1;
        for (uint64_t i = 0; i < nb; i++) {
          const uint8_t ch = sq[i];
          const uint64_t c = ch == 'A' ? 0 : ch == 'C' ? 1 : ch == 'G' ? 2 : 3;
          fwd = ((fwd << 2) | c) & lmask;
          rc = (rc >> 2) | ((3 - c) << (2 * (l - 1)));
          uint64_t cand = ~0ULL;
          if (i + 1 >= (uint64_t)l) {
            const uint64_t rcv = rev0 ? (((rc << (64 - 2 * l)) | ((1ULL << (64 - 2 * l)) - 1ULL)) & lmask) : rc;
            cand = ((fwd < rcv ? fwd : rcv) & smask) ^ toggle;
          }
          ring[i % (uint64_t)w] = cand;
          if (i + 1 < k) continue;
          uint64_t m = ring[0];
          for (int j = 1; j < w; j++) m = ring[j] < m ? ring[j] : m;
          m ^= toggle;
          const uint64_t pos = p0 + (i + 1 - k);
          if (pos % K2S_TILE_POS == 0) last = ~0ULL; /* the GPU builder de-duplicates per tile */
          if (m != last) {
            last = m;
            const uint64_t g = gpos + (pos / K2S_TILE_POS) * K2S_TILE_POS;
            const uint64_t blk = g / block_bases, r = g % block_bases;
            claimed += (uint64_t)insert_lca_atomic(t, tax, m, leaf_taxa[blk % (uint64_t)n_leaves]);
            if (r >= overlap_start) claimed += (uint64_t)insert_lca_atomic(t, tax, m, leaf_taxa[(blk + 1) % (uint64_t)n_leaves]);
          }
        }
      }
    }
    gpos += n - (k - 1); /* chunks overlap by k-1 bases so every k-mer is seen once */
    if (claimed == 0 && ++stalls > 8) break;
    size += claimed;
  }
  free(buf);
  t->size = size;
  if (genome_bases) *genome_bases = gpos + (k - 1);
  return 0;
}

typedef struct {
  uint64_t seed, genome_seed, genome_bases;
  double human_frac, sub_rate, ins_rate, del_rate, n_rate;
  int32_t paired, reserved;
  double insert_mean, insert_sd;
} k2s_reads_params; /* = nh_synth_reads_params_t */

/* byte-for-byte what k_synth_reads writes (nh_synth.cu) */
void k2s_reads(uint8_t *bases, const uint64_t *off, uint64_t n_seqs, const k2s_reads_params *p, int threads) {
  if (threads <= 0) threads = omp_get_max_threads();
  const uint32_t human_thr = (uint32_t)(p->human_frac * 16777216.0);
  const double s1 = p->sub_rate, s2 = s1 + p->ins_rate, s3 = s2 + p->del_rate;
  const uint32_t sub_thr = (uint32_t)(s1 * 65536.0), ins_thr = (uint32_t)(s2 * 65536.0), del_thr = (uint32_t)(s3 * 65536.0);
  const uint32_t n_thr = (uint32_t)(p->n_rate * 16777216.0);
  const float insert_mean = (float)p->insert_mean, insert_sd = (float)p->insert_sd;
#pragma omp parallel for num_threads(threads) schedule(dynamic, 1024)
  for (int64_t si = 0; si < (int64_t)n_seqs; si++) {
    const uint64_t s = (uint64_t)si;
    const uint64_t o = off[s];
    const uint64_t len = off[s + 1] - o;
    if (len == 0) continue;
    const uint64_t u = p->paired ? (s >> 1) : s;
    const uint32_t mate = p->paired ? (uint32_t)(s & 1ULL) : 0u;
    const uint64_t hu = k2o_fmix64(p->seed ^ k2o_fmix64(u + 0x51ED27ULL));
    const int human = (uint32_t)(hu & 0xFFFFFFu) < human_thr;
    uint64_t frag_len = len + len / 8 + 64; /* slack for deletions */
    if (p->paired) {
      const uint64_t h2 = k2o_fmix64(hu + 1);
      float z = 0.f; /* ~normal from 4 uniforms */
      for (int i = 0; i < 4; i++) z = fmaf((float)((h2 >> (16 * i)) & 0xFFFFu), 1.0f / 65536.0f, z);
      z = (z - 2.0f) * 1.7320508f;
      long long fl = (long long)fmaf(insert_sd, z, insert_mean);
      const uint64_t other = s ^ 1ULL;
      const uint64_t len_other = off[other + 1] - off[other];
      const uint64_t lmax = len > len_other ? len : len_other;
      if (fl < (long long)lmax) fl = (long long)lmax;
      frag_len = (uint64_t)fl + lmax / 8 + 64;
    }
    const uint64_t h3 = k2o_fmix64(hu + 2);
    const uint64_t span = p->genome_bases > frag_len + 1 ? p->genome_bases - frag_len - 1 : 1;
    const uint64_t start = umulhi64(h3, span);
    const uint32_t strand = (uint32_t)(k2o_fmix64(hu + 3) & 1ULL);
    const int forward = (mate ^ strand) == 0u;
    const uint64_t hs = k2o_fmix64(hu ^ (0xA5A5ULL + mate));
    const int has_n = (uint32_t)(hs & 0xFFFFFFu) < n_thr;
    const uint64_t n_pos = umulhi64(k2o_fmix64(hs + 7), len);
    for (uint64_t c = 0; c * 32ULL < len; c++) {
      uint64_t rng = k2o_fmix64(hs + 0x1000ULL + c);
      uint64_t sp = c * 32ULL; /* source offset inside the fragment */
      const uint64_t jend = (c * 32ULL + 32ULL < len) ? c * 32ULL + 32ULL : len;
      for (uint64_t j = c * 32ULL; j < jend; j++) {
        rng = rng * 6364136223846793005ULL + 1442695040888963407ULL;
        const uint32_t x = (uint32_t)(rng >> 48);
        const uint32_t rb = (uint32_t)(rng >> 40) & 3u;
        uint32_t code = rb;
        if (human) {
          int take_src = 1;
          if (x < sub_thr) {
            sp++;
            take_src = 0;
          } else if (x < ins_thr) {
            take_src = 0;
          } else if (x < del_thr) {
            sp++;
          }
          if (take_src) {
            const uint64_t fo = sp < frag_len ? sp : frag_len - 1;
            code = forward ? genome_code(p->genome_seed, start + fo)
                           : 3u - genome_code(p->genome_seed, start + frag_len - 1 - fo);
            sp++;
          }
        }
        uint8_t ch = (uint8_t)"ACGT"[code];
        if (has_n && j == n_pos) ch = 'N';
        bases[o + j] = ch;
      }
    }
  }
}
