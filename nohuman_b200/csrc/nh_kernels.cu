/*
 * nh_kernels.cu — the classification path as hand-written sm_100a kernels.
 * Integer/byte work bound by issue slots (minimizer scan) and by random
 * 32-byte reads of the hash table, which on B200 are limited by address
 * translation; no tensor-core work exists on this path.
 *
 * Upstream units each kernel stands for (kraken2 @ reference Dockerfile:15,
 * 35-38; behavioural spec SURVEY.md Appendix A):
 *   k_plan_*             (none: tiles, lookup slots, tile roles, deferred units)
 *   k_stream_classify    mmscanner.cc MinimizerScanner::NextMinimizer/is_ambiguous,
 *                        the last_minimizer de-duplication of classify.cc
 *                        ClassifySequence, kv_store.h MurmurHash3, compact_hash.cc
 *                        CompactHashTable::Get (LINEAR_PROBING build), classify.cc
 *                        ResolveTree + taxonomy.cc IsAAncestorOfB /
 *                        LowestCommonAncestor — one kernel, the default path (A.3-A.5)
 *   k_scan_probe_score   the same work in three phases (NH_FUSED_KERNEL=phased)
 *   k_score, k_score_big ClassifySequence tail + ResolveTree for multi-tile units
 *                        (long reads) and units with many distinct taxa         (A.5)
 *   k_minimizers, k_probe  warp-per-tile scan and thread-per-lookup probe: any window
 *                        width k-l+1 <= 32, the synthetic builder, A/B runs     (A.3, A.4)
 *   k_gather_runs        per-read hit runs for kraken2's --output lines          (A.6)
 * and, in the reference itself, the keep/drop polarity of
 * src/main.rs:259-265 (--classified-out vs --unclassified-out).
 */
#include <stdlib.h>
#include <string.h>

#include "nh_kernels.cuh"

#define FULL_MASK 0xFFFFFFFFu

/* ------------------------------------------------------------------ */
/* small helpers                                                       */

__device__ __forceinline__ uint32_t lane_id() { return threadIdx.x & 31u; }

__device__ __forceinline__ uint32_t warp_sum_u32(uint32_t v) {
  return __reduce_add_sync(FULL_MASK, v);
}

/* exclusive scan over a block of up to 1024 threads; *total gets the sum.
 * The plan scans (tiles << 32 | k-mer positions) in one pass. */
__device__ __forceinline__ uint64_t block_excl_scan(uint64_t v, uint64_t *total,
                                                    uint64_t *s_warp /* [33] */) {
  uint32_t lane = lane_id(), warp = threadIdx.x >> 5;
  uint64_t inc = v;
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    uint64_t o = __shfl_up_sync(FULL_MASK, inc, d);
    if (lane >= (uint32_t)d) inc += o;
  }
  if (lane == 31) s_warp[warp] = inc;
  __syncthreads();
  if (warp == 0) {
    uint32_t nw = (blockDim.x + 31) >> 5;
    uint64_t ws = lane < nw ? s_warp[lane] : 0;
    uint64_t winc = ws;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      uint64_t o = __shfl_up_sync(FULL_MASK, winc, d);
      if (lane >= (uint32_t)d) winc += o;
    }
    s_warp[lane] = winc - ws; /* exclusive warp base */
    if (lane == 31) s_warp[32] = winc;
  }
  __syncthreads();
  uint64_t r = s_warp[warp] + inc - v;
  *total = s_warp[32];
  __syncthreads();
  return r;
}

/* ------------------------------------------------------------------ */
/* stage 0: plan — cut every sequence into tiles of <= tile_pos k-mers  */

#define PLAN_THREADS 1024

__device__ __forceinline__ uint32_t seq_npos(const uint64_t *__restrict__ off, uint32_t s, int k) {
  uint64_t len = off[s + 1] - off[s];
  return len < (uint64_t)k ? 0u : (uint32_t)(len - (uint64_t)k + 1);
}

__device__ __forceinline__ uint32_t npos_ntiles(uint32_t npos, int tile_pos) {
  return (npos + (uint32_t)tile_pos - 1u) / (uint32_t)tile_pos;
}

/* (tiles << 32) | positions of sequence s */
__device__ __forceinline__ uint64_t seq_work(const uint64_t *__restrict__ off, uint32_t s, int k,
                                             int tile_pos) {
  const uint32_t np = seq_npos(off, s, k);
  return ((uint64_t)npos_ntiles(np, tile_pos) << 32) | np;
}

__global__ void __launch_bounds__(PLAN_THREADS)
k_plan_count(const uint64_t *__restrict__ off, uint32_t n_seqs, int k, int tile_pos,
             uint64_t *__restrict__ block_sums) {
  __shared__ uint64_t s_warp[33];
  uint32_t s = blockIdx.x * PLAN_THREADS + threadIdx.x;
  uint64_t v = s < n_seqs ? seq_work(off, s, k, tile_pos) : 0;
  uint64_t total;
  block_excl_scan(v, &total, s_warp);
  if (threadIdx.x == 0) block_sums[blockIdx.x] = total;
}

/* single block: exclusive scan of the per-block sums, reset of the batch counters */
__global__ void __launch_bounds__(PLAN_THREADS)
k_plan_scan(uint64_t *__restrict__ block_sums, uint32_t nb, NhCounters *__restrict__ counters) {
  __shared__ uint64_t s_warp[33];
  __shared__ uint64_t s_running;
  if (threadIdx.x == 0) s_running = 0;
  __syncthreads();
  for (uint32_t base = 0; base < nb; base += PLAN_THREADS) {
    uint32_t i = base + threadIdx.x;
    uint64_t v = i < nb ? block_sums[i] : 0;
    uint64_t total;
    uint64_t ex = block_excl_scan(v, &total, s_warp);
    uint64_t run = s_running;
    if (i < nb) block_sums[i] = run + ex;
    __syncthreads();
    if (threadIdx.x == 0) s_running = run + total;
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    counters->n_tiles = (uint32_t)(s_running >> 32);
    counters->n_lookups = 0;
    counters->n_classified = 0;
    counters->n_kept = 0;
    counters->n_overflow = 0;
    counters->error = 0;
    counters->n_deferred = 0;
    counters->next_group = 0;
  }
}

/* Writes the tile descriptors.  With `fused` set it also decides, per unit,
 * whether the fused kernel can score it inside one warp (every mate at most
 * one tile, both tiles in the same group of 32) and queues the other units
 * for k_score. */
__global__ void __launch_bounds__(PLAN_THREADS)
k_plan_fill(const uint64_t *__restrict__ off, uint32_t n_seqs, int k, int tile_pos, int paired,
            int fused, const uint64_t *__restrict__ block_sums, uint32_t *__restrict__ tile_base,
            NhTile *__restrict__ tiles, uint32_t *__restrict__ deferred_units,
            NhCounters *__restrict__ counters) {
  __shared__ uint64_t s_warp[33];
  uint32_t s = blockIdx.x * PLAN_THREADS + threadIdx.x;
  uint64_t w = s < n_seqs ? seq_work(off, s, k, tile_pos) : 0;
  uint64_t total;
  uint64_t ex = block_excl_scan(w, &total, s_warp);
  if (s < n_seqs) {
    const uint64_t before = block_sums[blockIdx.x] + ex;
    const uint32_t tb = (uint32_t)(before >> 32), pb = (uint32_t)before;
    const uint32_t v = (uint32_t)(w >> 32);
    tile_base[s] = tb;
    uint32_t role = NH_ROLE_DEFERRED;
    if (fused) {
      const uint32_t mate = paired ? (s & 1u) : 0u;
      const uint32_t v_other = paired ? npos_ntiles(seq_npos(off, s ^ 1u, k), tile_pos) : 0u;
      bool simple = v <= 1u && v_other <= 1u && v + v_other >= 1u;
      if (v == 1u && v_other == 1u) {
        const uint32_t t0 = mate == 0u ? tb : tb - 1u;
        simple = simple && (t0 & 31u) != 31u;
      }
      if (simple && v == 1u)
        role = v_other == 0u ? NH_ROLE_LEADER : (mate == 0u ? NH_ROLE_LEADER2 : NH_ROLE_PARTNER);
      if (mate == 0u && !simple)
        deferred_units[atomicAdd(&counters->n_deferred, 1u)] = paired ? (s >> 1) : s;
    }
    for (uint32_t t = 0; t < v; t++) {
      NhTile d;
      d.seq = s;
      d.pos_begin = t * (uint32_t)tile_pos;
      d.slot = pb + d.pos_begin;
      d.role = role;
      tiles[tb + t] = d;
    }
  }
  if (s == 0) tile_base[n_seqs] = counters->n_tiles;
}

/* ------------------------------------------------------------------ */
/* stage 1: minimizers — one warp per tile                             */

struct __align__(16) MinWarpSmem {
  uint64_t st_min[NH_TILE_LMERS];      /* staged distinct-consecutive minimizers */
  uint32_t packed[16];                 /* 2-bit bases, 16 per word, first base in the MSBs */
  uint32_t ambits[8];                  /* 1 bit per base, LSB first */
  uint8_t st_start[NH_TILE_LMERS + 8]; /* ordinal (among non-ambiguous positions) of each run start */
};

__device__ __forceinline__ bool any_bits(const uint32_t *bits, uint32_t start, uint32_t n) {
  uint32_t j = start >> 5, o = start & 31u;
  uint32_t lo = __funnelshift_r(bits[j], bits[j + 1], o);
  if (n <= 32) return (lo & (n == 32 ? 0xFFFFFFFFu : ((1u << n) - 1u))) != 0;
  uint32_t hi = __funnelshift_r(bits[j + 1], bits[j + 2], o);
  n -= 32;
  return lo != 0 || (hi & (n >= 32 ? 0xFFFFFFFFu : ((1u << n) - 1u))) != 0;
}

__device__ __forceinline__ uint64_t shfl_carry(uint64_t cur, uint64_t prev, uint32_t d,
                                               uint32_t lane) {
  /* value of lane (lane - d); lanes that wrap take the previous iteration's value */
  uint64_t send = (lane >= 32u - d) ? prev : cur;
  return __shfl_sync(FULL_MASK, send, (lane - d) & 31u);
}

__device__ __forceinline__ uint64_t min_u64(uint64_t a, uint64_t b) { return a < b ? a : b; }

/* Returns the number of runs staged in sm (warp-uniform). */
template <int WT>
__device__ __forceinline__ uint32_t minimizer_tile(const NhDbParams &db, const NhBatchPtrs &b,
                                                   uint32_t tile, MinWarpSmem &sm,
                                                   uint32_t lane) {
  const int k = db.k, l = db.l;
  const int w = WT ? WT : db.w;
  const NhTile t = b.tiles[tile];
  const uint64_t so = b.offsets[t.seq];
  const uint32_t len = (uint32_t)(b.offsets[t.seq + 1] - so);
  const uint32_t npos_total = len - (uint32_t)k + 1u;
  uint32_t npos = npos_total - t.pos_begin;
  if (npos > (uint32_t)db.tile_pos) npos = (uint32_t)db.tile_pos;
  const uint32_t nb = npos + (uint32_t)k - 1u; /* bases of this tile */
  const uint32_t nq = npos + (uint32_t)w - 1u; /* l-mers of this tile */

  /* -- load + 2-bit pack: 8 bases per lane, 8-byte aligned loads -- */
  const uint8_t *g = b.bases + so + t.pos_begin;
  const uint32_t shift = (uint32_t)((uintptr_t)g & 7u);
  const uint8_t *ga = g - shift;
  const uint32_t nchunks = (shift + nb + 7u) >> 3;
  uint32_t half = 0, amb8 = 0;
  if (lane < nchunks) {
    uint2 v = __ldg(reinterpret_cast<const uint2 *>(ga) + lane);
    uint32_t a0, a1;
    uint32_t p0 = nh_pack4(v.x, &a0);
    uint32_t p1 = nh_pack4(v.y, &a1);
    half = (p0 << 8) | p1;
    amb8 = a0 | (a1 << 4);
    /* keep only ambiguity bits of bases inside [shift, shift + nb) */
    int lo = (int)shift - (int)(lane * 8u);
    int hi = (int)(shift + nb) - (int)(lane * 8u);
    lo = lo < 0 ? 0 : lo;
    hi = hi > 8 ? 8 : hi;
    uint32_t keep = (lo < 8 && hi > lo) ? (((1u << hi) - 1u) & ~((1u << lo) - 1u)) : 0u;
    amb8 &= keep;
  }
  reinterpret_cast<uint16_t *>(sm.packed)[lane ^ 1u] = (uint16_t)half;
  reinterpret_cast<uint8_t *>(sm.ambits)[lane] = (uint8_t)amb8;
  const bool any_amb = __ballot_sync(FULL_MASK, amb8 != 0) != 0;
  __syncwarp();

  uint64_t prev_lvl[6];
#pragma unroll
  for (int i = 0; i < 6; i++) prev_lvl[i] = NH_NONE64;
  uint64_t carry_last = NH_NONE64; /* last non-ambiguous minimizer seen in this tile */
  uint32_t nonamb_sofar = 0, n_runs = 0;
  const uint32_t lane_lt = (1u << lane) - 1u;
  const uint32_t iters = (nq + 31u) >> 5;

  for (uint32_t it = 0; it < iters; it++) {
    const uint32_t qi = it * 32u + lane; /* l-mer index in the tile */
    const uint32_t tb = shift + qi;      /* its first base in the packed stream */
    bool lmer_ok = qi < nq;
    if (any_amb && lmer_ok) lmer_ok = !any_bits(sm.ambits, tb, (uint32_t)l);
    uint64_t cand = NH_NONE64;
    {
      uint64_t x = nh_extract_lmer(sm.packed, tb, l);
      uint64_t rc = nh_revcomp(x, l, db.revcom_version);
      uint64_t c = min_u64(x, rc) & db.seed_mask;
      c ^= db.toggle;
      if (lmer_ok) cand = c;
    }
    /* sliding-window minimum over the last w l-mers (prefix doubling) */
    uint64_t m = cand;
    uint32_t span = 1;
#pragma unroll
    for (int lev = 0; lev < 5; lev++) {
      if ((int)(span * 2u) <= w) {
        uint64_t o = shfl_carry(m, prev_lvl[lev], span, lane);
        prev_lvl[lev] = m;
        m = min_u64(m, o);
        span *= 2u;
      }
    }
    if ((int)span < w) {
      uint64_t o = shfl_carry(m, prev_lvl[5], (uint32_t)w - span, lane);
      prev_lvl[5] = m;
      m = min_u64(m, o);
    }
    const uint64_t mz = m ^ db.toggle;

    const bool valid = qi >= (uint32_t)(w - 1) && qi < nq;
    const uint32_t pp = qi - (uint32_t)(w - 1); /* k-mer position within the tile */
    bool pos_amb = false;
    if (any_amb && valid)
      pos_amb = any_bits(sm.ambits, tb + (uint32_t)l - (uint32_t)db.amb_span,
                         (uint32_t)db.amb_span);
    const bool nonamb = valid && !pos_amb;

    /* de-duplicate against the previous non-ambiguous position */
    const uint32_t nb_mask = __ballot_sync(FULL_MASK, nonamb);
    const uint32_t lower = nb_mask & lane_lt;
    const uint32_t src = lower ? 31u - (uint32_t)__clz(lower) : lane;
    const uint64_t pm = __shfl_sync(FULL_MASK, mz, src);
    const uint64_t prev_m = lower ? pm : carry_last;
    const bool is_new = nonamb && (mz != prev_m);
    const uint32_t ordinal = nonamb_sofar + __popc(lower);
    const uint32_t new_mask = __ballot_sync(FULL_MASK, is_new);
    if (is_new) {
      uint32_t slot = n_runs + __popc(new_mask & lane_lt);
      sm.st_min[slot] = mz;
      sm.st_start[slot] = (uint8_t)ordinal;
    }
    n_runs += __popc(new_mask);
    if (nb_mask) carry_last = __shfl_sync(FULL_MASK, mz, 31u - (uint32_t)__clz(nb_mask));
    nonamb_sofar += __popc(nb_mask);

    if (b.dbg_pos_min != nullptr && valid) {
      uint64_t o = b.dbg_pos_offsets[t.seq] + t.pos_begin + pp;
      b.dbg_pos_min[o] = mz;
      b.dbg_pos_ambig[o] = pos_amb ? 1 : 0;
    }
  }
  if (lane == 0) sm.st_start[n_runs] = (uint8_t)nonamb_sofar;
  __syncwarp();
  return n_runs;
}

template <int WT>
__global__ void __launch_bounds__(NH_BLOCK_THREADS)
k_minimizers(const NhDbParams db, const NhBatchPtrs b) {
  __shared__ MinWarpSmem s_warp[NH_WARPS_PER_BLOCK];
  __shared__ uint32_t s_runs[2][NH_WARPS_PER_BLOCK];
  __shared__ uint32_t s_base[2];
  const uint32_t lane = lane_id(), warp = threadIdx.x >> 5;
  const uint32_t n_tiles = b.counters->n_tiles;
  uint32_t par = 0;
  for (uint32_t group = blockIdx.x; group * NH_WARPS_PER_BLOCK < n_tiles;
       group += gridDim.x, par ^= 1u) {
    const uint32_t tile = group * NH_WARPS_PER_BLOCK + warp;
    uint32_t n_runs = 0;
    if (tile < n_tiles) n_runs = minimizer_tile<WT>(db, b, tile, s_warp[warp], lane);
    if (lane == 0) s_runs[par][warp] = n_runs;
    __syncthreads();
    if (threadIdx.x == 0) {
      uint32_t tot = 0;
#pragma unroll
      for (int i = 0; i < NH_WARPS_PER_BLOCK; i++) tot += s_runs[par][i];
      s_base[par] = tot ? atomicAdd(&b.counters->n_lookups, tot) : 0u;
    }
    __syncthreads();
    if (tile < n_tiles) {
      uint32_t base = s_base[par];
      for (uint32_t i = 0; i < warp; i++) base += s_runs[par][i];
      const MinWarpSmem &sm = s_warp[warp];
      for (uint32_t r = lane; r < n_runs; r += 32u) {
        b.lk_min[base + r] = sm.st_min[r];
        b.lk_cnt[base + r] = (uint8_t)(sm.st_start[r + 1] - sm.st_start[r]);
      }
      if (lane == 0) {
        NhTileOut o;
        o.lk_off = base;
        o.lk_cnt = n_runs;
        b.tile_out[tile] = o;
      }
    }
    __syncwarp();
  }
}

/* ------------------------------------------------------------------ */
/* stage 2: compact hash table probe — one thread per lookup            */

__device__ __forceinline__ void ld_sector(const uint32_t *p, uint32_t (&c)[8]) {
  /* one 32-byte sector in one request (256-bit global load, sm_100+) */
#ifndef NH_LD_SECTOR_OP
#define NH_LD_SECTOR_OP "ld.global.nc.L1::no_allocate.L2::64B.v8.u32"
#endif
  asm volatile(NH_LD_SECTOR_OP " {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=r"(c[0]), "=r"(c[1]), "=r"(c[2]), "=r"(c[3]), "=r"(c[4]), "=r"(c[5]),
                 "=r"(c[6]), "=r"(c[7])
               : "l"(p));
}

__device__ __forceinline__ uint32_t cht_get(const NhDbParams &db, uint64_t key) {
  const uint64_t h = nh_fmix64(key);
  if (db.min_hash && h < db.min_hash) return 0;
  const uint32_t ckey = (uint32_t)(h >> (32u + db.value_bits));
  uint64_t idx = nh_fastmod(h, db.capacity, db.mod_m, db.mod_sh1, db.mod_sh2);
  uint64_t inspected = 0;
  for (;;) {
    const uint64_t sbase = idx & ~7ULL;
    uint32_t c[8];
    ld_sector(db.cells + sbase, c);
    const int start = (int)(idx & 7ULL);
    const uint64_t rem = db.capacity - sbase;
    const int limit = rem < 8 ? (int)rem : 8;
    int state = -1;
#pragma unroll
    for (int j = 7; j >= 0; j--) {
      const uint32_t val = c[j] & db.value_mask;
      const bool term = (val == 0) || ((c[j] >> db.value_bits) == ckey);
      if (j >= start && j < limit && term) state = (int)val;
    }
    if (state >= 0) return (uint32_t)state;
    inspected += (uint64_t)(limit - start);
    if (inspected >= db.capacity) return 0; /* table without an empty cell */
    idx = sbase + 8;
    if (idx >= db.capacity) idx = 0;
  }
}

__global__ void __launch_bounds__(256)
k_probe(const NhDbParams db, const uint64_t *__restrict__ keys, uint32_t *__restrict__ taxa,
        const uint32_t *__restrict__ n_dev, uint32_t n_fixed) {
  const uint32_t n = n_dev ? *n_dev : n_fixed;
  const uint32_t stride = gridDim.x * blockDim.x;
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride)
    taxa[i] = cht_get(db, keys[i]);
}

/* ------------------------------------------------------------------ */
/* stage 3: per-unit scoring, confidence walk-up, keep/drop             */

__device__ __forceinline__ uint32_t hash_slot(uint32_t taxon, uint32_t cap_mask) {
  return (taxon * 2654435761u >> 7) & cap_mask;
}

__device__ __forceinline__ bool hc_add(uint32_t *keys, uint32_t *cnts, uint32_t cap_mask,
                                       uint32_t taxon, uint32_t n) {
  uint32_t s = hash_slot(taxon, cap_mask);
  for (uint32_t p = 0; p <= cap_mask; p++) {
    uint32_t old = atomicCAS(&keys[s], 0u, taxon);
    if (old == 0u || old == taxon) {
      atomicAdd(&cnts[s], n);
      return true;
    }
    s = (s + 1u) & cap_mask;
  }
  return false;
}

__device__ __forceinline__ uint32_t hc_get(const uint32_t *keys, const uint32_t *cnts,
                                           uint32_t cap_mask, uint32_t taxon) {
  if (!taxon) return 0;
  uint32_t s = hash_slot(taxon, cap_mask);
  for (uint32_t p = 0; p <= cap_mask; p++) {
    uint32_t kx = keys[s];
    if (kx == taxon) return cnts[s];
    if (kx == 0u) return 0;
    s = (s + 1u) & cap_mask;
  }
  return 0;
}

__device__ __forceinline__ bool is_a_ancestor_of_b(const uint32_t *parent, uint32_t a,
                                                   uint32_t b) {
  if (!a || !b) return false;
  while (b > a) b = parent[b];
  return b == a;
}

__device__ __forceinline__ uint32_t lca(const uint32_t *parent, uint32_t a, uint32_t b) {
  if (!a || !b) return a ? a : b;
  while (a != b) {
    if (a > b)
      a = parent[a];
    else
      b = parent[b];
  }
  return a;
}

/* Classifies unit u with the warp; returns false if the taxon table overflowed. */
__device__ bool score_unit(const NhDbParams &db, const NhBatchPtrs &b, const NhScoreParams &sp,
                           const uint32_t *parent, uint32_t *keys, uint32_t *cnts,
                           uint32_t cap_mask, uint32_t u, uint32_t lane, uint32_t *classified,
                           uint32_t *kept, bool reprobe = false) {
  for (uint32_t s = lane; s <= cap_mask; s += 32u) {
    keys[s] = 0;
    cnts[s] = 0;
  }
  __syncwarp();
  const uint32_t nm = b.paired ? 2u : 1u;
  const uint32_t seq0 = u * nm;
  uint32_t total_kmers = 0;
  int groups = 0;
  bool ok = true;
  for (uint32_t mate = 0; mate < nm; mate++) {
    const uint32_t s = seq0 + mate;
    const uint64_t len = b.offsets[s + 1] - b.offsets[s];
    if (len >= (uint64_t)db.k) total_kmers += (uint32_t)(len - (uint64_t)db.k + 1);
    const uint32_t t0 = b.tile_base[s], t1 = b.tile_base[s + 1];
    if (b.tile_tab != nullptr && !reprobe) {
      /* streaming kernel: every tile left its own small table; lanes take tiles in parallel */
      bool any_ovf = false;
      for (uint32_t tile = t0 + lane; tile < t1; tile += 32u) {
        const NhTileSum ts = b.tile_sum[tile];
        groups += (int)ts.groups;
        if (ts.flags & NH_TILE_OVERFLOW) {
          any_ovf = true;
        } else {
          const NhTileTab tt = b.tile_tab[tile];
#pragma unroll
          for (int j = 0; j < NH_LANE_TAXA; j++)
            if (tt.keys[j]) ok &= hc_add(keys, cnts, cap_mask, tt.keys[j], tt.cnts[j]);
        }
        /* a tile starts with a fresh lookup even when its first minimizer equals the last one of
         * the tile before: upstream counts that as one group */
        if (tile > t0 && (ts.flags & NH_TILE_HAS) && (ts.flags & NH_TILE_FIRST_HIT)) {
          uint32_t p = tile;
          while (p > t0 && !(b.tile_sum[p - 1u].flags & NH_TILE_HAS)) p--;
          if (p > t0 && b.tile_sum[p - 1u].last_min == ts.first_min) groups--;
        }
      }
      if (__any_sync(FULL_MASK, any_ovf)) {
        /* rare: tiles whose table overflowed kept their lookups; count them with fresh probes */
        for (uint32_t tile = t0; tile < t1; tile++) {
          if (!(b.tile_sum[tile].flags & NH_TILE_OVERFLOW)) continue;
          const NhTileOut to = b.tile_out[tile];
          for (uint32_t j = lane; j < to.lk_cnt; j += 32u) {
            const uint32_t tx = cht_get(db, b.lk_min[to.lk_off + j]);
            if (tx) ok &= hc_add(keys, cnts, cap_mask, tx, b.lk_cnt[to.lk_off + j]);
          }
        }
      }
      continue;
    }
    uint64_t prev_last = NH_NONE64; /* ClassifySequence resets last_minimizer per mate */
    for (uint32_t tile = t0; tile < t1; tile++) {
      const NhTileOut to = b.tile_out[tile];
      for (uint32_t j = lane; j < to.lk_cnt; j += 32u) {
        const uint32_t tx = reprobe ? cht_get(db, b.lk_min[to.lk_off + j]) : b.lk_taxon[to.lk_off + j];
        if (tx) {
          groups++;
          ok &= hc_add(keys, cnts, cap_mask, tx, b.lk_cnt[to.lk_off + j]);
        }
      }
      if (to.lk_cnt && t1 - t0 > 1u) {
        /* a tile starts with a fresh lookup even when its first minimizer equals
         * the last one of the previous tile: upstream counts that as one group */
        if (lane == 0 && tile > t0 && b.lk_min[to.lk_off] == prev_last &&
            (reprobe ? cht_get(db, b.lk_min[to.lk_off]) : b.lk_taxon[to.lk_off]) != 0u)
          groups--;
        prev_last = b.lk_min[to.lk_off + to.lk_cnt - 1u];
      }
    }
  }
  __syncwarp();
  groups = (int)__reduce_add_sync(FULL_MASK, (uint32_t)groups);
  if (!__all_sync(FULL_MASK, ok)) return false;

  /* ResolveTree: root-to-leaf score of every hit taxon, ties fold to the LCA */
  uint32_t best_s = 0, best_t = 0;
  for (uint32_t s = lane; s <= cap_mask; s += 32u) {
    const uint32_t tx = keys[s];
    if (!tx) continue;
    uint32_t score = 0;
    for (uint32_t a = tx; a; a = parent[a]) score += hc_get(keys, cnts, cap_mask, a);
    if (score > best_s) {
      best_s = score;
      best_t = tx;
    } else if (score == best_s) {
      best_t = lca(parent, best_t, tx);
    }
  }
#pragma unroll
  for (int d = 16; d >= 1; d >>= 1) {
    const uint32_t os = __shfl_xor_sync(FULL_MASK, best_s, d);
    const uint32_t ot = __shfl_xor_sync(FULL_MASK, best_t, d);
    if (os > best_s) {
      best_s = os;
      best_t = ot;
    } else if (os == best_s) {
      best_t = lca(parent, best_t, ot);
    }
  }
  uint32_t max_taxon = best_t;
  uint32_t max_score = hc_get(keys, cnts, cap_mask, max_taxon);
  const uint32_t required = (uint32_t)ceil(__dmul_rn(sp.confidence, (double)total_kmers));
  while (max_taxon && max_score < required) {
    uint32_t part = 0;
    for (uint32_t s = lane; s <= cap_mask; s += 32u) {
      const uint32_t tx = keys[s];
      if (tx && is_a_ancestor_of_b(parent, max_taxon, tx)) part += cnts[s];
    }
    max_score = warp_sum_u32(part);
    if (max_score >= required) break;
    max_taxon = parent[max_taxon];
  }
  uint32_t call = max_taxon;
  if (call && groups < sp.min_hit_groups) call = 0;
  if (lane == 0) {
    const uint32_t is_cls = call != 0u;
    const uint32_t keep = sp.keep_human ? is_cls : !is_cls;
    if (b.out_call) b.out_call[u] = call ? db.ext_id[call] : 0u;
    if (b.out_keep) b.out_keep[u] = (uint8_t)keep;
    if (b.dbg_call) b.dbg_call[u] = call;
    if (b.dbg_total_kmers) b.dbg_total_kmers[u] = total_kmers;
    if (b.dbg_hit_groups) b.dbg_hit_groups[u] = (uint32_t)groups;
    *classified += is_cls;
    *kept += keep;
  }
  __syncwarp();
  return true;
}

template <bool SMEM_PARENT>
__global__ void __launch_bounds__(NH_BLOCK_THREADS)
k_score(const NhDbParams db, const NhBatchPtrs b, const NhScoreParams sp) {
  extern __shared__ uint32_t s_dyn[];
  uint32_t *s_parent = s_dyn;
  const uint32_t parent_words = SMEM_PARENT ? db.node_count : 0u;
  uint32_t *s_hash = s_dyn + parent_words;
  if (SMEM_PARENT) {
    for (uint32_t i = threadIdx.x; i < db.node_count; i += blockDim.x) s_parent[i] = db.parent[i];
    __syncthreads();
  }
  const uint32_t *parent = SMEM_PARENT ? s_parent : db.parent;
  const uint32_t lane = lane_id(), warp = threadIdx.x >> 5;
  uint32_t *keys = s_hash + warp * 2u * NH_WARP_HASH_SLOTS;
  uint32_t *cnts = keys + NH_WARP_HASH_SLOTS;
  uint32_t classified = 0, kept = 0;
  const uint32_t wstride = gridDim.x * NH_WARPS_PER_BLOCK;
  /* fused path: only the units k_plan_fill deferred; legacy path: every unit */
  const uint32_t n_todo = b.deferred_units ? b.counters->n_deferred : b.n_units;
  for (uint32_t i = blockIdx.x * NH_WARPS_PER_BLOCK + warp; i < n_todo; i += wstride) {
    const uint32_t u = b.deferred_units ? b.deferred_units[i] : i;
    if (!score_unit(db, b, sp, parent, keys, cnts, NH_WARP_HASH_SLOTS - 1u, u, lane, &classified,
                    &kept)) {
      if (lane == 0) b.overflow_units[atomicAdd(&b.counters->n_overflow, 1u)] = u;
    }
  }
  if (lane == 0) {
    if (classified) atomicAdd(&b.counters->n_classified, classified);
    if (kept) atomicAdd(&b.counters->n_kept, kept);
  }
}

/* overflow pass: units with more distinct taxa than the per-warp table holds;
 * one warp per block with a 16K-slot table */
__global__ void __launch_bounds__(32)
k_score_big(const NhDbParams db, const NhBatchPtrs b, const NhScoreParams sp) {
  extern __shared__ uint32_t s_dyn[];
  uint32_t *keys = s_dyn;
  uint32_t *cnts = s_dyn + NH_BIG_HASH_SLOTS;
  const uint32_t lane = lane_id();
  const uint32_t n = b.counters->n_overflow;
  uint32_t classified = 0, kept = 0;
  for (uint32_t i = blockIdx.x; i < n; i += gridDim.x) {
    const uint32_t raw = b.overflow_units[i];
    const uint32_t u = raw & ~NH_OVERFLOW_REPROBE;
    if (!score_unit(db, b, sp, db.parent, keys, cnts, NH_BIG_HASH_SLOTS - 1u, u, lane,
                    &classified, &kept, (raw & NH_OVERFLOW_REPROBE) != 0u)) {
      if (lane == 0) atomicExch(&b.counters->error, 1u);
    }
  }
  if (lane == 0) {
    if (classified) atomicAdd(&b.counters->n_classified, classified);
    if (kept) atomicAdd(&b.counters->n_kept, kept);
  }
}

/* ------------------------------------------------------------------ */
/* fused path: scan -> probe -> score in one kernel                     */
/*
 * One warp takes a group of 32 consecutive tiles.
 *   phase A  every lane scans ITS OWN tile base by base: rolling forward and
 *            reverse-complement l-mers, canonical/seed/toggle, window minimum
 *            over a register ring, run-length de-duplication; the distinct
 *            minimizers go to the tile's lookup slots in global memory (they
 *            stay in L2 for phase B).  ~1/3 of the instructions of the
 *            warp-per-tile kernel, because nothing is recomputed per l-mer.
 *   phase B  the group's lookups are flattened over the 32 lanes and probed
 *            in rounds.  A lane whose probe chain runs into the next sector
 *            does not loop on its own: it pushes the continuation onto a
 *            32-entry queue in shared memory and the next round hands the
 *            queue and fresh lookups out over all 32 lanes again, so every
 *            round has 32 independent DRAM requests in flight.  Results are
 *            folded straight into the owning unit's taxon table (shared-memory
 *            atomics); only tiles of deferred units write lk_taxon.
 *   phase C  lanes that lead a short unit (roles from k_plan_fill) run
 *            ResolveTree serially on their table; other units go to k_score.
 * Only the window width W is a template parameter (kraken2's default
 * k=35,l=31 gives W=5); other databases take the warp-per-tile kernels.
 */

#define NH_META_DEFERRED 0x80u

#define NH_RING_SLOTS 128u       /* look-ahead ring: two rounds of 32*U lookups */
#define NH_RING_SKIP 0x80000000u /* below minimum_acceptable_hash_value: no lookup, taxon 0 */

struct __align__(16) FusedWarpSmem {
  /* look-ahead ring: the group's lookups hashed 32 at a time (one per lane) */
  uint32_t r_unit[NH_RING_SLOTS];        /* aligned group of G sectors holding hash % capacity */
  uint32_t r_ckey[NH_RING_SLOTS];
  uint32_t r_slot[NH_RING_SLOTS];
  uint32_t r_aux[NH_RING_SLOTS];         /* owner tile | start cell in the group << 5 | NH_RING_SKIP */
  uint32_t q_unit[64];                   /* continuation queue: next group of sectors to read */
  uint32_t q_ckey[64];
  uint32_t q_slot[64];
  uint32_t q_aux[64];                    /* owner tile | groups visited << 5 */
  uint32_t prefix[33];                   /* exclusive scan of lookups per tile */
  uint32_t slot[32];                     /* first lookup slot of each tile */
  uint32_t keys[NH_LANE_TAXA * 32];      /* taxon tables, [slot][owner lane] */
  uint32_t cnts[NH_LANE_TAXA * 32];
  uint32_t groups[32];                   /* minimizer_hit_groups per owner lane */
  uint8_t meta[32];                      /* per tile: owner lane | NH_META_DEFERRED */
  uint32_t overflow;                     /* bit per owner lane: table overflowed */
};

template <typename SM>
__device__ __forceinline__ uint32_t lane_tab_get(const SM &sm, uint32_t lane, uint32_t n, uint32_t taxon) {
  for (uint32_t i = 0; i < n; i++)
    if (sm.keys[i * 32u + lane] == taxon) return sm.cnts[i * 32u + lane];
  return 0;
}

#ifndef NH_FUSED_MIN_BLOCKS
#define NH_FUSED_MIN_BLOCKS 4
#endif
#ifndef NH_PROBE_DEPTH_DEFAULT
#define NH_PROBE_DEPTH_DEFAULT 1
#endif
#ifndef NH_FUSED_STREAM_DEFAULT
#define NH_FUSED_STREAM_DEFAULT 1
#endif
template <int W, int G, int U>
__global__ void __launch_bounds__(NH_BLOCK_THREADS, NH_FUSED_MIN_BLOCKS)
k_scan_probe_score(const NhDbParams db, const NhBatchPtrs b, const NhScoreParams sp) {
  /* G lanes read the G adjacent sectors of one aligned 32*G-byte group with ONE warp
   * instruction: the memory system charges one request per instruction and 128-byte line
   * (profiles/r01_random_access_microbench.txt), so a probe chain that stays inside the
   * group costs nothing extra. */
  static_assert(G == 1 || G == 2 || G == 4, "a group is 32, 64 or 128 bytes");
  static_assert(U == 1 || U == 2, "sector reads in flight per lane");
  constexpr uint32_t NI = 32u * U / G;      /* lookups per round */
  constexpr uint32_t RMASK = NH_RING_SLOTS - 1u;
  constexpr uint32_t GCELLS = 8u * G;       /* cells per group */
  extern __shared__ __align__(16) uint32_t s_dyn[];
  uint32_t *s_parent = s_dyn;
  const bool smem_parent = db.node_count <= NH_SMEM_PARENT_MAX;
  const uint32_t parent_words = smem_parent ? db.node_count : 0u;
  FusedWarpSmem *s_warps = reinterpret_cast<FusedWarpSmem *>(s_dyn + ((parent_words + 3u) & ~3u));
  if (smem_parent) {
    for (uint32_t i = threadIdx.x; i < db.node_count; i += blockDim.x) s_parent[i] = db.parent[i];
    __syncthreads();
  }
  const uint32_t *parent = smem_parent ? s_parent : db.parent;
  const uint32_t lane = lane_id(), warp = threadIdx.x >> 5;
  const uint32_t lane_lt = (1u << lane) - 1u;
  FusedWarpSmem &sm = s_warps[warp];
  const uint32_t n_tiles = b.counters->n_tiles;
  const int k = db.k, l = db.l;
  const uint64_t lmask = (1ULL << (2 * l)) - 1ULL;
  const uint32_t rc_shift = 2u * (uint32_t)(l - 1);
  const uint64_t n_units = (db.capacity + GCELLS - 1ULL) / GCELLS;
  /* probe-chain guard for a table without any empty cell (never a real database) */
  const uint32_t max_visits = n_units + 1ULL < 0x7FFFFFFULL ? (uint32_t)(n_units + 1ULL) : 0x7FFFFFFu;
  uint32_t tot_lookups = 0, tot_classified = 0, tot_kept = 0;

  for (uint32_t group = blockIdx.x * NH_WARPS_PER_BLOCK + warp; group * 32u < n_tiles;
       group += gridDim.x * NH_WARPS_PER_BLOCK) {
    const uint32_t tile = group * 32u + lane;
    const bool have = tile < n_tiles;
    NhTile t;
    t.seq = 0; t.pos_begin = 0; t.slot = 0; t.role = NH_ROLE_DEFERRED;
    if (have) t = b.tiles[tile];
#pragma unroll
    for (int i = 0; i < NH_LANE_TAXA; i++) {
      sm.keys[i * 32 + lane] = 0;
      sm.cnts[i * 32 + lane] = 0;
    }
    sm.groups[lane] = 0;
    if (lane == 0) sm.overflow = 0;
    sm.meta[lane] = (uint8_t)(t.role == NH_ROLE_DEFERRED ? NH_META_DEFERRED
                              : (t.role == NH_ROLE_PARTNER ? lane - 1u : lane));

    /* ---------------- phase A: lane-serial minimizer scan ---------------- */
    uint32_t n_runs = 0;
    {
      uint64_t so = 0;
      uint32_t nb = 0; /* bases of this lane's tile */
      if (have) {
        so = b.offsets[t.seq];
        const uint32_t len = (uint32_t)(b.offsets[t.seq + 1] - so);
        uint32_t npos = len - (uint32_t)k + 1u - t.pos_begin;
        if (npos > (uint32_t)db.tile_pos) npos = (uint32_t)db.tile_pos;
        nb = npos + (uint32_t)k - 1u;
      }
      const uint8_t *g = b.bases + so + t.pos_begin;
      const uint32_t mis = (uint32_t)((uintptr_t)g & 3u);
      const uint32_t *q = reinterpret_cast<const uint32_t *>(g - mis);
      const uint32_t my_words = have ? (mis + nb + 3u) >> 2 : 0u;
      const uint32_t max_words = __reduce_max_sync(FULL_MASK, my_words);

      uint64_t fwd = 0, rc = 0;
      uint64_t ring[W > 1 ? W - 1 : 1];
#pragma unroll
      for (int i = 0; i < (W > 1 ? W - 1 : 1); i++) ring[i] = NH_NONE64;
      uint32_t c_run = 0;           /* consecutive unambiguous bases ending here */
      uint64_t last = NH_NONE64;    /* minimizer of the open run */
      uint32_t cnt = 0;             /* k-mer positions in the open run */
      uint64_t *out_min = b.lk_min + t.slot;
      uint8_t *out_cnt = b.lk_cnt + t.slot;
      const bool dbg = b.dbg_pos_min != nullptr;
      const uint64_t dbg_base = dbg && have ? b.dbg_pos_offsets[t.seq] + t.pos_begin : 0;
      const uint32_t first_pos = mis + (uint32_t)(k - 1); /* word-stream index of the first k-mer end */
      const uint32_t end_idx = mis + nb;

      /* W - 1 bases per inner iteration: the ring shift is then pure register renaming */
      static_assert(W == 5, "the inner loop consumes one 4-byte word per ring rotation");
      uint32_t codes = 0, ambs = 0;
      uint32_t w_next = my_words ? __ldg(q) : 0u; /* loaded one word ahead of its use */
      for (uint32_t base_i = 0; base_i < max_words * 4u; base_i += 4u) {
#pragma unroll
        for (int j = 0; j < 4; j++) {
          const uint32_t i = base_i + (uint32_t)j; /* index in the word-aligned stream */
          if (j == 0) {
            codes = nh_pack4(w_next, &ambs); /* first base in bits 7..6 */
            const uint32_t wi = (base_i >> 2) + 1u;
            w_next = wi < my_words ? __ldg(q + wi) : 0u;
          }
          const uint32_t sh = 6u - 2u * (i & 3u);
          const uint32_t c = (codes >> sh) & 3u;
          const bool inside = i >= mis && i < end_idx;
          /* bytes outside the tile count as ambiguous: they reset the l-mer and never reach a position */
          const bool amb = ((ambs >> (i & 3u)) & 1u) || !inside;
          fwd = ((fwd << 2) | c) & lmask;
          rc = (rc >> 2) | ((uint64_t)(3u - c) << rc_shift);
          c_run = amb ? 0u : c_run + 1u;
          uint64_t cand = NH_NONE64;
          if (c_run >= (uint32_t)l) {
            const uint64_t rcv = db.revcom_version == 0
                                     ? (((rc << (64 - 2 * l)) | ((1ULL << (64 - 2 * l)) - 1ULL)) & lmask)
                                     : rc;
            cand = ((fwd < rcv ? fwd : rcv) & db.seed_mask) ^ db.toggle;
          }
          uint64_t m = cand;
#pragma unroll
          for (int r = 0; r < W - 1; r++) m = min_u64(m, ring[r]);
#pragma unroll
          for (int r = W - 2; r > 0; r--) ring[r] = ring[r - 1];
          if (W > 1) ring[0] = cand;
          const bool at_pos = inside && i >= first_pos;
          const bool nonamb = at_pos && c_run >= (uint32_t)db.amb_span;
          const uint64_t mz = m ^ db.toggle;
          if (dbg && at_pos) {
            const uint64_t o = dbg_base + (i - first_pos);
            b.dbg_pos_min[o] = mz;
            b.dbg_pos_ambig[o] = nonamb ? 0 : 1;
          }
          const bool newrun = nonamb && mz != last;
          if (newrun && cnt) {
            out_min[n_runs] = last;
            out_cnt[n_runs] = (uint8_t)cnt;
            n_runs++;
          }
          cnt = newrun ? 1u : cnt + (nonamb ? 1u : 0u);
          last = newrun ? mz : last;
        }
      }
      if (cnt) {
        out_min[n_runs] = last;
        out_cnt[n_runs] = (uint8_t)cnt;
        n_runs++;
      }
      if (have) {
        NhTileOut o;
        o.lk_off = t.slot;
        o.lk_cnt = n_runs;
        b.tile_out[tile] = o;
      }
    }

    /* ---------------- phase B: probe the group's lookups ---------------- */
    uint32_t inc = n_runs;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      const uint32_t o = __shfl_up_sync(FULL_MASK, inc, d);
      if (lane >= (uint32_t)d) inc += o;
    }
    const uint32_t total = __shfl_sync(FULL_MASK, inc, 31);
    sm.prefix[lane] = inc - n_runs;
    sm.slot[lane] = t.slot;
    if (lane == 31) sm.prefix[32] = total;
    __syncwarp(); /* also orders the lookup slots written above before the reads below */
    tot_lookups += n_runs;

    uint32_t qn = 0, e_next = 0, e_ready = 0;
    /* hash the next (up to) 32*U lookups, U per lane, into the ring */
    auto prepare = [&]() {
      uint32_t n = total - e_ready;
      n = n < 32u * U ? n : 32u * U;
#pragma unroll
      for (int u = 0; u < U; u++) {
        const uint32_t e = e_ready + (uint32_t)u * 32u + lane;
        if (e < e_ready + n) {
          uint32_t o = 0; /* owner tile: largest o with prefix[o] <= e */
#pragma unroll
          for (int step = 16; step >= 1; step >>= 1)
            if (sm.prefix[o + step] <= e) o += (uint32_t)step;
          const uint32_t oslot = sm.slot[o] + (e - sm.prefix[o]);
          const uint64_t h = nh_fmix64(__ldcg(b.lk_min + oslot));
          uint32_t aux = o | NH_RING_SKIP;
          if (!(db.min_hash && h < db.min_hash)) {
            const uint64_t idx = nh_fastmod(h, db.capacity, db.mod_m, db.mod_sh1, db.mod_sh2);
            sm.r_unit[e & RMASK] = (uint32_t)(idx / GCELLS);
            sm.r_ckey[e & RMASK] = (uint32_t)(h >> (32u + db.value_bits));
            aux = o | ((uint32_t)(idx % GCELLS) << 5);
          }
          sm.r_slot[e & RMASK] = oslot;
          sm.r_aux[e & RMASK] = aux;
        }
      }
      e_ready += n;
    };
    prepare();
    __syncwarp();
    const uint32_t sub = lane % G;
    while (qn != 0u || e_next < total) {
      /* ---- every lane takes U work items: continuations first, then fresh lookups ---- */
      bool active[U], skip[U];
      uint32_t unit[U], ckey[U], oslot[U], aux[U], start[U];
      uint32_t c[U][8];
#pragma unroll
      for (int u = 0; u < U; u++) {
        const uint32_t it = (uint32_t)u * (32u / G) + lane / G;
        active[u] = false;
        skip[u] = false;
        unit[u] = ckey[u] = oslot[u] = aux[u] = start[u] = 0;
        if (it < qn) {
          unit[u] = sm.q_unit[it];
          ckey[u] = sm.q_ckey[it];
          oslot[u] = sm.q_slot[it];
          aux[u] = sm.q_aux[it];
          active[u] = true;
        } else {
          const uint32_t e = e_next + (it - qn);
          if (e < total) {
            active[u] = true;
            oslot[u] = sm.r_slot[e & RMASK];
            const uint32_t ra = sm.r_aux[e & RMASK];
            aux[u] = ra & 31u;
            if (ra & NH_RING_SKIP) {
              skip[u] = true;
            } else {
              unit[u] = sm.r_unit[e & RMASK];
              ckey[u] = sm.r_ckey[e & RMASK];
              start[u] = (ra >> 5) & 31u;
            }
          }
        }
      }
      e_next += NI - qn;
      if (e_next > total) e_next = total;
      /* ---- all sector reads of the round are issued before any is looked at ---- */
#pragma unroll
      for (int u = 0; u < U; u++)
        if (active[u] && !skip[u]) ld_sector(db.cells + ((uint64_t)unit[u] * G + sub) * 8ULL, c[u]);
      __syncwarp(); /* queue and ring fully read before they are refilled */
      if (e_ready < total && e_ready - e_next <= 32u * U) prepare(); /* its key reads overlap the sector reads */
      uint32_t q_fill = 0;
#pragma unroll
      for (int u = 0; u < U; u++) {
        const bool probing = active[u] && !skip[u];
        const uint64_t cell0 = ((uint64_t)unit[u] * G + sub) * 8ULL; /* first cell of this lane's sector */
        int state = -1;
        if (probing) {
          /* cells of this sector the chain may stop at: from `start` on, inside the table */
          const int lo = (int)start[u] - (int)(sub * 8u);
          uint32_t range = lo <= 0 ? 0xFFu : (lo >= 8 ? 0u : (0xFFu << lo) & 0xFFu);
          if (cell0 + 8ULL > db.capacity)
            range &= cell0 >= db.capacity ? 0u : (1u << (uint32_t)(db.capacity - cell0)) - 1u;
#pragma unroll
          for (int j = 7; j >= 0; j--) {
            const uint32_t val = c[u][j] & db.value_mask;
            const bool term = (val == 0u) || ((c[u][j] >> db.value_bits) == ckey[u]);
            if (term && ((range >> j) & 1u)) state = (int)val;
          }
        }
        bool done = skip[u];
        uint32_t result = 0;
        if (G == 1) {
          if (state >= 0) {
            done = true;
            result = (uint32_t)state;
          }
        } else {
          /* the first lane of the G that found a terminal cell has the answer */
          const uint32_t fmask = __ballot_sync(FULL_MASK, state >= 0);
          const uint32_t gbase = lane & ~(uint32_t)(G - 1);
          const uint32_t gbits = (fmask >> gbase) & ((1u << G) - 1u);
          const uint32_t any = (uint32_t)__shfl_sync(FULL_MASK, state, gbase + (gbits ? (uint32_t)__ffs(gbits) - 1u : 0u));
          if (gbits) {
            done = true;
            result = any;
          }
        }
        if (active[u] && !done) {
          const uint32_t visits = (aux[u] >> 5) + 1u;
          if (visits >= max_visits) {
            done = true; /* went round a table without an empty cell */
          } else {
            aux[u] = (aux[u] & 31u) | (visits << 5);
            unit[u] = (uint64_t)unit[u] + 1ULL >= n_units ? 0u : unit[u] + 1u;
          }
        }
        const bool leader = active[u] && sub == 0u;
        const bool cont = leader && !done;
        const uint32_t cmask = __ballot_sync(FULL_MASK, cont);
        if (cont) {
          const uint32_t pos = q_fill + __popc(cmask & lane_lt);
          sm.q_unit[pos] = unit[u];
          sm.q_ckey[pos] = ckey[u];
          sm.q_slot[pos] = oslot[u];
          sm.q_aux[pos] = aux[u];
        }
        q_fill += __popc(cmask);
        if (leader && done) {
          const uint32_t mt = sm.meta[aux[u] & 31u];
          if ((mt & NH_META_DEFERRED) || b.emit_all_taxa) b.lk_taxon[oslot[u]] = result;
          if (!(mt & NH_META_DEFERRED) && result) {
            const uint32_t own = mt;
            const uint32_t n = (uint32_t)__ldcg(b.lk_cnt + oslot[u]);
            atomicAdd(&sm.groups[own], 1u);
            int i = 0;
            for (; i < sp.lane_taxa; i++) {
              const uint32_t old = atomicCAS(&sm.keys[i * 32 + own], 0u, result);
              if (old == 0u || old == result) {
                atomicAdd(&sm.cnts[i * 32 + own], n);
                break;
              }
            }
            if (i == sp.lane_taxa) atomicOr(&sm.overflow, 1u << own);
          }
        }
      }
      qn = q_fill;
      __syncwarp();
    }

    /* ---------------- phase C: score short units in the warp ---------------- */
    if (have && (t.role == NH_ROLE_LEADER || t.role == NH_ROLE_LEADER2)) {
      const uint32_t u = b.paired ? (t.seq >> 1) : t.seq;
      if ((sm.overflow >> lane) & 1u) {
        /* more distinct taxa than a lane table holds: the big-table pass takes the unit.
         * Its taxa were folded, not stored, so the flag bit tells k_score_big to probe again. */
        b.overflow_units[atomicAdd(&b.counters->n_overflow, 1u)] = u | NH_OVERFLOW_REPROBE;
      } else {
        uint32_t ntab = 0;
        while (ntab < (uint32_t)sp.lane_taxa && sm.keys[ntab * 32u + lane] != 0u) ntab++;
        const int groups = (int)sm.groups[lane];
        const uint32_t s0 = b.paired ? (t.seq & ~1u) : t.seq;
        uint32_t total_kmers = 0;
        for (uint32_t mm = 0; mm < (b.paired ? 2u : 1u); mm++) {
          const uint64_t len = b.offsets[s0 + mm + 1] - b.offsets[s0 + mm];
          if (len >= (uint64_t)k) total_kmers += (uint32_t)(len - (uint64_t)k + 1);
        }
        /* ResolveTree */
        uint32_t best_s = 0, best_t = 0;
        for (uint32_t i = 0; i < ntab; i++) {
          const uint32_t tx = sm.keys[i * 32u + lane];
          uint32_t score = 0;
          for (uint32_t a = tx; a; a = parent[a]) score += lane_tab_get(sm, lane, ntab, a);
          if (score > best_s) {
            best_s = score;
            best_t = tx;
          } else if (score == best_s) {
            best_t = lca(parent, best_t, tx);
          }
        }
        uint32_t max_taxon = best_t;
        uint32_t max_score = max_taxon ? lane_tab_get(sm, lane, ntab, max_taxon) : 0u;
        const uint32_t required = (uint32_t)ceil(__dmul_rn(sp.confidence, (double)total_kmers));
        while (max_taxon && max_score < required) {
          uint32_t sum = 0;
          for (uint32_t i = 0; i < ntab; i++)
            if (is_a_ancestor_of_b(parent, max_taxon, sm.keys[i * 32u + lane]))
              sum += sm.cnts[i * 32u + lane];
          max_score = sum;
          if (max_score >= required) break;
          max_taxon = parent[max_taxon];
        }
        uint32_t call = max_taxon;
        if (call && groups < sp.min_hit_groups) call = 0;
        const uint32_t is_cls = call != 0u;
        const uint32_t keep = sp.keep_human ? is_cls : !is_cls;
        if (b.out_call) b.out_call[u] = call ? db.ext_id[call] : 0u;
        if (b.out_keep) b.out_keep[u] = (uint8_t)keep;
        if (b.dbg_call) b.dbg_call[u] = call;
        if (b.dbg_total_kmers) b.dbg_total_kmers[u] = total_kmers;
        if (b.dbg_hit_groups) b.dbg_hit_groups[u] = (uint32_t)groups;
        tot_classified += is_cls;
        tot_kept += keep;
      }
    }
    __syncwarp();
  }
  tot_lookups = warp_sum_u32(tot_lookups);
  tot_classified = warp_sum_u32(tot_classified);
  tot_kept = warp_sum_u32(tot_kept);
  if (lane == 0) {
    if (tot_lookups) atomicAdd(&b.counters->n_lookups, tot_lookups);
    if (tot_classified) atomicAdd(&b.counters->n_classified, tot_classified);
    if (tot_kept) atomicAdd(&b.counters->n_kept, tot_kept);
  }
}

/* ------------------------------------------------------------------ */
/* fused path, streaming form: the scan feeds the probe through shared memory */
/*
 * Same work as k_scan_probe_score, different schedule.  There the 32 lanes
 * scan their whole tiles first and park every lookup in global memory; each
 * of those scattered 8-byte and 1-byte stores is its own memory request, and
 * requests are what this path is short of.  Here a closed run
 * goes straight into a small queue in shared memory; whenever 32 lookups are
 * waiting the warp hashes them and issues their sector reads, then goes back
 * to scanning.  The sectors are looked at one probe round later, i.e. after
 * the warp has scanned further, so the scan's integer work hides the
 * table's latency inside one warp instead of relying on other warps.
 * Tiles of deferred units (long reads) and sessions with emit_runs still get
 * their lookups written to global memory for k_score / k_gather_runs.
 */

#ifndef NH_STREAM_TMA
#define NH_STREAM_TMA 1
#endif
#define NH_BCHUNK_WORDS 8u                          /* base words (4 bases each) per lane and chunk */
#define NH_BCHUNK_STRIDE (NH_BCHUNK_WORDS * 4u + 16u) /* + up to 12 bytes of 16-byte misalignment */

struct __align__(16) StreamWarpSmem {
  uint64_t pq_key[128];                  /* closed runs waiting to be probed (ring) */
  uint32_t pq_slot[128];
  uint16_t pq_meta[128];                 /* owner lane | k-mer count << 5 | first lookup of its tile << 13 */
  uint32_t q_unit[32];                   /* probe chains that continue into the next sector (at most one per lane) */
  uint32_t q_ckey[32];
  uint32_t q_slot[32];
  uint32_t q_aux[32];                    /* owner lane | k-mer count << 5 | first << 13 | sectors visited << 14 */
  uint64_t first_min[32], last_min[32];  /* first / last distinct minimizer of each lane's tile */
  uint32_t keys[NH_LANE_TAXA * 32];      /* taxon tables, [slot][owner lane] */
  uint32_t cnts[NH_LANE_TAXA * 32];
  uint32_t groups[32];                   /* minimizer_hit_groups per owner lane */
  uint8_t meta[32];                      /* per tile: owner lane (the lane before, for a second mate) */
  uint32_t overflow;                     /* bit per owner lane: table overflowed */
  uint32_t first_hit;                    /* bit per lane: the tile's first lookup hit */
#if NH_STREAM_TMA
  /* base staging: every lane's next NH_BCHUNK_WORDS words arrive by asynchronous 16-byte copies
   * (cp.async, completion on an mbarrier) into its own 48-byte window, double-buffered.  The scan then reads
   * bases with LDS only: a global base-word load shares its scoreboard with the table sector
   * loads (ptxas puts every LDG on SB5), so a lane waiting for 4 bases also waited ~1 us for the
   * sector read the probe round had just issued. */
  uint64_t bbar[2];
  __align__(16) uint8_t bchunk[2][32 * NH_BCHUNK_STRIDE];
  /* table sectors land here too (one 32-byte sector per lane and probe round): a register
   * destination would keep a load scoreboard busy, and ptxas drains every scoreboard at the first
   * potentially divergent branch, i.e. a few instructions after the round instead of one round later */
  uint64_t sbar;
  __align__(16) uint32_t sect[32 * 8];
#endif
};

#define NH_AUX_NONE 0xFFFFFFFFu

__device__ __forceinline__ uint32_t smem_addr(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
/* 16 bytes global -> shared without a register or a load scoreboard in between (LDGSTS, L2 only) */
__device__ __forceinline__ void cp_async16(uint32_t dst, const void *src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async16_l2_64(uint32_t dst, const void *src) {
  asm volatile("cp.async.cg.shared.global.L2::64B [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
}
/* the mbarrier gets this thread's arrival once all its earlier cp.async copies have landed */
__device__ __forceinline__ void cp_async_arrive(uint32_t bar) {
  asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "NH_WAIT:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra NH_DONE;\n"
      "bra NH_WAIT;\n"
      "NH_DONE:\n"
      "}" ::"r"(bar),
      "r"(parity)
      : "memory");
}
__device__ __forceinline__ uint32_t lds_u32(uint32_t addr) {
  uint32_t v;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(addr));
  return v;
}

/* 3 blocks of 8 warps per SM (up to 85 registers): with the overlap happening inside the warp,
 * registers are worth more than resident warps (measured 3.14 ms at 80 registers / 24 warps
 * against 3.35 ms at 64 / 32 and 3.38 ms at 103 / 16) */
#ifndef NH_STREAM_MIN_BLOCKS
#define NH_STREAM_MIN_BLOCKS 3
#endif
#ifndef NH_STREAM_PREFETCH
#define NH_STREAM_PREFETCH 1
#endif
template <int W, bool DBG, bool REV0>
__global__ void __launch_bounds__(NH_BLOCK_THREADS, NH_STREAM_MIN_BLOCKS)
k_stream_classify(const NhDbParams db, const NhBatchPtrs b, const NhScoreParams sp) {
  static_assert(W == 5, "the scan consumes one 4-byte word per ring rotation");
  extern __shared__ __align__(16) uint32_t s_dyn[];
  uint32_t *s_parent = s_dyn;
  const bool smem_parent = db.node_count <= NH_SMEM_PARENT_MAX;
  const uint32_t parent_words = smem_parent ? db.node_count : 0u;
  StreamWarpSmem *s_warps = reinterpret_cast<StreamWarpSmem *>(s_dyn + ((parent_words + 3u) & ~3u));
  if (smem_parent) {
    for (uint32_t i = threadIdx.x; i < db.node_count; i += blockDim.x) s_parent[i] = db.parent[i];
    __syncthreads();
  }
  const uint32_t *parent = smem_parent ? s_parent : db.parent;
  const uint32_t lane = lane_id(), warp = threadIdx.x >> 5;
  const uint32_t lane_lt = (1u << lane) - 1u;
  StreamWarpSmem &sm = s_warps[warp];
  const uint32_t n_tiles = b.counters->n_tiles;
  const int k = db.k, l = db.l;
  const uint64_t lmask = (1ULL << (2 * l)) - 1ULL;
  const uint32_t rc_shift = 2u * (uint32_t)(l - 1);
  const uint32_t n_sectors = (uint32_t)((db.capacity + 7ULL) >> 3); /* nh_fused_supported: capacity < 2^35 */
  const uint32_t last_sector = n_sectors - 1u;
  /* cells of the last sector that exist (the allocation is zero-padded past the table's end) */
  const uint32_t last_range = (db.capacity & 7ULL) ? (1u << (uint32_t)(db.capacity & 7ULL)) - 1u : 0xFFu;
  /* probe-chain guard for a table without any empty cell (never a real database) */
  const uint32_t max_visits = n_sectors + 1u < 0x3FFFFu ? n_sectors + 1u : 0x3FFFFu;
  uint32_t tot_lookups = 0, tot_classified = 0, tot_kept = 0;
#if NH_STREAM_TMA
  const uint32_t bar0 = smem_addr(&sm.bbar[0]);
  const uint32_t win0 = smem_addr(&sm.bchunk[0][lane * NH_BCHUNK_STRIDE]);
  uint32_t bar_par = 0; /* bit per buffer: parity of the phase its next chunk completes (warp-uniform) */
  const uint32_t sbar = smem_addr(&sm.sbar);
  const uint32_t sect0 = smem_addr(&sm.sect[0]);
  uint32_t sect_par = 0; /* parity of the phase the sectors in flight complete */
  if (lane == 0) {
    mbar_init(bar0, 32u); /* every lane arrives once per chunk, when its copies have landed */
    mbar_init(bar0 + 8u, 32u);
    mbar_init(sbar, 32u);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  __syncwarp();
#endif

  /* groups of 32 tiles are handed out through a counter: warps that draw short tiles take more */
  for (;;) {
    uint32_t group = 0;
    if (lane == 0) group = atomicAdd(&b.counters->next_group, 1u);
    group = __shfl_sync(FULL_MASK, group, 0);
    if ((uint64_t)group * 32ULL >= n_tiles) break;
    const uint32_t tile = group * 32u + lane;
    const bool have = tile < n_tiles;
    NhTile t;
    t.seq = 0; t.pos_begin = 0; t.slot = 0; t.role = NH_ROLE_DEFERRED;
    if (have) t = b.tiles[tile];
#pragma unroll
    for (int i = 0; i < NH_LANE_TAXA; i++) {
      sm.keys[i * 32 + lane] = 0;
      sm.cnts[i * 32 + lane] = 0;
    }
    sm.groups[lane] = 0;
    if (lane == 0) {
      sm.overflow = 0;
      sm.first_hit = 0;
    }
    /* every tile folds its hits into a table: a second mate into its leader's, all others into
     * their own (tiles of deferred units hand theirs to k_score through tile_tab) */
    sm.meta[lane] = (uint8_t)(t.role == NH_ROLE_PARTNER ? lane - 1u : lane);
    /* the per-read output needs every lookup in global memory */
    const bool spill_runs = b.emit_all_taxa != 0;
    __syncwarp();

    /* warp-uniform queue state */
    uint32_t pq_head = 0, pq_n = 0, cq_n = 0;
    /* the lookup this lane has in flight: its sector is in c[], looked at in the next round */
    uint32_t c[8];
    uint32_t f_unit = 0, f_ckey = 0, f_slot = 0, f_aux = NH_AUX_NONE, f_start = 0;
    bool any_inflight = false; /* warp-uniform */

    /* one probe round: finish the lookups in flight, then put up to 32 waiting ones in flight */
    auto probe_round = [&]() {
      /* ---- 1. look at the sectors issued last round ---- */
      if (any_inflight) {
        const bool active = f_aux != NH_AUX_NONE;
        bool done = false;
        uint32_t result = 0;
#if NH_STREAM_TMA
        mbar_wait(sbar, sect_par);
        sect_par ^= 1u;
        {
          const uint4 lo = *reinterpret_cast<const uint4 *>(&sm.sect[lane * 8u]);
          const uint4 hi = *reinterpret_cast<const uint4 *>(&sm.sect[lane * 8u + 4u]);
          c[0] = lo.x; c[1] = lo.y; c[2] = lo.z; c[3] = lo.w;
          c[4] = hi.x; c[5] = hi.y; c[6] = hi.z; c[7] = hi.w;
        }
#endif
        if (active) {
          uint32_t range = 0xFFu << f_start;
          if (f_unit == last_sector) range &= last_range;
          int state = -1;
#pragma unroll
          for (int j = 7; j >= 0; j--) {
            const uint32_t val = c[j] & db.value_mask;
            const bool term = (val == 0u) || ((c[j] >> db.value_bits) == f_ckey);
            if (term && ((range >> j) & 1u)) state = (int)val;
          }
          if (state >= 0) {
            done = true;
            result = (uint32_t)state;
          } else {
            const uint32_t visits = (f_aux >> 14) + 1u;
            if (visits >= max_visits) {
              done = true; /* went round a table without an empty cell */
            } else {
              f_aux = (f_aux & 0x3FFFu) | (visits << 14);
              f_unit = f_unit == last_sector ? 0u : f_unit + 1u;
            }
          }
        }
        const bool cont = active && !done;
        const uint32_t cmask = __ballot_sync(FULL_MASK, cont);
        if (cont) {
          const uint32_t pos = cq_n + __popc(cmask & lane_lt);
          sm.q_unit[pos] = f_unit;
          sm.q_ckey[pos] = f_ckey;
          sm.q_slot[pos] = f_slot;
          sm.q_aux[pos] = f_aux;
        }
        cq_n += __popc(cmask);
        if (active && done) {
          const uint32_t own_lane = f_aux & 31u;
          if (b.emit_all_taxa) b.lk_taxon[f_slot] = result;
          if (result) {
            const uint32_t own = sm.meta[own_lane];
            const uint32_t n = (f_aux >> 5) & 0xFFu;
            if (f_aux & 0x2000u) atomicOr(&sm.first_hit, 1u << own_lane);
            atomicAdd(&sm.groups[own], 1u);
            int i = 0;
            for (; i < sp.lane_taxa; i++) {
              const uint32_t old = atomicCAS(&sm.keys[i * 32 + own], 0u, result);
              if (old == 0u || old == result) {
                atomicAdd(&sm.cnts[i * 32 + own], n);
                break;
              }
            }
            if (i == sp.lane_taxa) atomicOr(&sm.overflow, 1u << own);
          }
        }
        __syncwarp(); /* continuation queue written before it is read below */
      }
      /* ---- 2. next 32 lookups: continuations first, then fresh runs ---- */
      const uint32_t n_cq = cq_n < 32u ? cq_n : 32u;
      const uint32_t room = 32u - n_cq;
      const uint32_t n_pq = pq_n < room ? pq_n : room;
      f_aux = NH_AUX_NONE;
      f_start = 0;
      if (lane < n_cq) {
        const uint32_t i = cq_n - 1u - lane; /* newest first: the queue stays a stack, no holes */
        f_unit = sm.q_unit[i];
        f_ckey = sm.q_ckey[i];
        f_slot = sm.q_slot[i];
        f_aux = sm.q_aux[i];
      } else if (lane - n_cq < n_pq) {
        const uint32_t i = (pq_head + (lane - n_cq)) & 127u;
        const uint64_t h = nh_fmix64(sm.pq_key[i]);
        f_slot = sm.pq_slot[i];
        const uint32_t meta = sm.pq_meta[i];
        if (db.min_hash && h < db.min_hash) {
          /* below minimum_acceptable_hash_value: kraken2 skips the lookup, taxon 0 */
          if (b.emit_all_taxa) b.lk_taxon[f_slot] = 0u;
        } else {
          const uint64_t idx = nh_fastmod(h, db.capacity, db.mod_m, db.mod_sh1, db.mod_sh2);
          f_unit = (uint32_t)(idx >> 3);
          f_start = (uint32_t)idx & 7u;
          f_ckey = (uint32_t)(h >> (32u + db.value_bits));
          f_aux = meta; /* owner | count << 5, zero sectors visited */
        }
      }
      cq_n -= n_cq;
      pq_head = (pq_head + n_pq) & 127u;
      pq_n -= n_pq;
      any_inflight = (n_cq + n_pq) != 0u;
#if NH_STREAM_TMA
      if (any_inflight) {
        /* lane pairs fetch the two 16-byte halves of one sector with ONE instruction, so a sector
         * stays one request (request = warp instruction x 128-byte line):
         * the first instruction covers the sectors of lanes 0-15, the second those of lanes 16-31 */
        const uint32_t mine = f_aux != NH_AUX_NONE ? f_unit : 0xFFFFFFFFu;
        const uint32_t half = lane & 1u, src_lane = lane >> 1;
        const uint32_t u0 = __shfl_sync(FULL_MASK, mine, src_lane);
        const uint32_t u1 = __shfl_sync(FULL_MASK, mine, 16u + src_lane);
        const uint32_t dst = sect0 + src_lane * 32u + half * 16u;
        if (u0 != 0xFFFFFFFFu) cp_async16_l2_64(dst, db.cells + (uint64_t)u0 * 8ULL + half * 4u);
        if (u1 != 0xFFFFFFFFu) cp_async16_l2_64(dst + 512u, db.cells + (uint64_t)u1 * 8ULL + half * 4u);
        cp_async_arrive(sbar);
      }
#else
      if (f_aux != NH_AUX_NONE) ld_sector(db.cells + (uint64_t)f_unit * 8ULL, c);
#endif
      __syncwarp(); /* queue slots just read may be overwritten by the next pushes */
    };

    /* ---------------- scan, feeding the probe ---------------- */
    /* Pass 0 is the real one.  Pass 1 runs only if a unit hit more distinct taxa than its
     * in-warp table holds (warp-uniform, rare): the lanes of such units scan again and write
     * their lookups to global memory, where k_score_big finds them. */
    uint32_t n_runs = 0;
    bool redo_lane = false;
    for (int pass = 0; pass < 2; pass++) {
      const bool rescan = pass == 1;
      const bool scanning = rescan ? redo_lane : have;
      const bool spill = rescan ? true : spill_runs;
      n_runs = 0;
      uint64_t so = 0;
      uint32_t nb = 0; /* bases of this lane's tile */
      if (scanning) {
        so = b.offsets[t.seq];
        const uint32_t len = (uint32_t)(b.offsets[t.seq + 1] - so);
        uint32_t npos = len - (uint32_t)k + 1u - t.pos_begin;
        if (npos > (uint32_t)db.tile_pos) npos = (uint32_t)db.tile_pos;
        nb = npos + (uint32_t)k - 1u;
      }
      const uint8_t *g = b.bases + so + t.pos_begin;
      const uint32_t mis = (uint32_t)((uintptr_t)g & 3u);
      const uint32_t *q = reinterpret_cast<const uint32_t *>(g - mis);
      const uint32_t my_words = scanning ? (mis + nb + 3u) >> 2 : 0u;
      const uint32_t max_words = __reduce_max_sync(FULL_MASK, my_words);

      uint64_t fwd = 0, rc = 0;
      /* window of 5 candidates c_i..c_{i-4}: a_i = min(c_i, c_{i-1}), m_i = min(a_i, a_{i-2}, c_{i-4});
       * ring[] holds c_{i-1}..c_{i-4}, pair[] holds a_{i-1}, a_{i-2} */
      uint64_t ring[4], pair[2];
#pragma unroll
      for (int i = 0; i < 4; i++) ring[i] = NH_NONE64;
      pair[0] = pair[1] = NH_NONE64;
      uint32_t c_run = 0;           /* consecutive unambiguous bases ending here */
      uint64_t last = NH_NONE64;    /* minimizer of the open run */
      uint32_t cnt = 0;             /* k-mer positions in the open run */
      const bool dbg = DBG && !rescan; /* per-position output for nh_debug_minimizers */
      const uint64_t dbg_base = dbg && scanning ? b.dbg_pos_offsets[t.seq] + t.pos_begin : 0;
      const uint32_t first_pos = mis + (uint32_t)(k - 1);
      const uint32_t end_idx = mis + nb;

      /* a closed run: to the shared queue (and to global memory when another kernel needs it) */
      auto emit = [&](bool pred, uint64_t key, uint32_t count) {
        const uint32_t slot = t.slot + n_runs;
        if (!rescan) {
          const uint32_t emask = __ballot_sync(FULL_MASK, pred);
          if (pred) {
            const uint32_t i = (pq_head + pq_n + __popc(emask & lane_lt)) & 127u;
            sm.pq_key[i] = key;
            sm.pq_slot[i] = slot;
            sm.pq_meta[i] = (uint16_t)(lane | (count << 5) | (n_runs == 0u ? 0x2000u : 0u));
          }
          pq_n += __popc(emask);
        }
        if (pred) {
          if (n_runs == 0u) sm.first_min[lane] = key;
          sm.last_min[lane] = key;
          if (spill) {
            b.lk_min[slot] = key;
            b.lk_cnt[slot] = (uint8_t)count;
          }
          n_runs++;
        }
      };

      /* four bases (one 4-byte word of the word-aligned stream, first base at stream index base_i) */
      auto scan_word = [&](const uint32_t word, const uint32_t base_i) {
        uint32_t ambs;
        const uint32_t codes = nh_pack4(word, &ambs); /* first base in bits 7..6 */
#pragma unroll
        for (int j = 0; j < 4; j++) {
          const uint32_t i = base_i + (uint32_t)j; /* index in the word-aligned stream */
          const uint32_t cc = (codes >> (6u - 2u * (uint32_t)j)) & 3u;
          const bool inside = i >= mis && i < end_idx;
          /* bytes outside the tile count as ambiguous: they reset the l-mer and never reach a position */
          const bool amb = ((ambs >> j) & 1u) || !inside;
          fwd = ((fwd << 2) | cc) & lmask;
          rc = (rc >> 2) | ((uint64_t)(3u - cc) << rc_shift);
          c_run = amb ? 0u : c_run + 1u;
          uint64_t cand = NH_NONE64;
          if (c_run >= (uint32_t)l) {
            /* revcom_version 0 (databases built before kraken2 2.0.8) keeps the un-shifted low bits */
            const bool rev0 = REV0 || (DBG && db.revcom_version == 0); /* the debug instantiation decides at run time */
            const uint64_t rcv = rev0 ? (((rc << (64 - 2 * l)) | ((1ULL << (64 - 2 * l)) - 1ULL)) & lmask) : rc;
            cand = ((fwd < rcv ? fwd : rcv) & db.seed_mask) ^ db.toggle;
          }
          const uint64_t a_i = min_u64(cand, ring[0]);
          const uint64_t m = min_u64(min_u64(a_i, pair[1]), ring[3]);
          pair[1] = pair[0];
          pair[0] = a_i;
#pragma unroll
          for (int r = 3; r > 0; r--) ring[r] = ring[r - 1];
          ring[0] = cand;
          const bool at_pos = inside && i >= first_pos;
          const bool nonamb = at_pos && c_run >= (uint32_t)db.amb_span;
          const uint64_t mz = m ^ db.toggle;
          if (dbg && at_pos) {
            const uint64_t o = dbg_base + (i - first_pos);
            b.dbg_pos_min[o] = mz;
            b.dbg_pos_ambig[o] = nonamb ? 0 : 1;
          }
          const bool newrun = nonamb && mz != last;
          emit(newrun && cnt != 0u, last, cnt);
          cnt = newrun ? 1u : cnt + (nonamb ? 1u : 0u);
          last = newrun ? mz : last;
#ifndef NH_STREAM_CHECK_MASK
#define NH_STREAM_CHECK_MASK 1 /* probe check after bases with (j & mask) == mask: 1 -> every 2nd base */
#endif
          if ((j & NH_STREAM_CHECK_MASK) == NH_STREAM_CHECK_MASK) {
#ifdef NH_STREAM_SINGLE_ROUND
            if (pq_n + cq_n >= 32u) probe_round();
            while (pq_n > 64u) probe_round();
#else
            while (pq_n + cq_n >= 32u) probe_round();
#endif
          }
        }
      };

#if NH_STREAM_TMA
      /* the lane's window for chunk c: bytes [32c, 32c + 48) from src0, the 16-byte block holding word 0 */
      const uint32_t a16 = (uint32_t)((uintptr_t)q & 15u);
      const uint8_t *src0 = reinterpret_cast<const uint8_t *>(q) - a16;
      const uint32_t my_bytes = my_words ? a16 + my_words * 4u : 0u;
      const uint32_t n_chunks = (max_words + NH_BCHUNK_WORDS - 1u) / NH_BCHUNK_WORDS;
      auto stage = [&](const uint32_t ch) {
        const uint32_t buf = ch & 1u;
        const uint32_t lo = ch * (NH_BCHUNK_WORDS * 4u);
        const uint32_t dst = win0 + buf * (32u * NH_BCHUNK_STRIDE);
#pragma unroll
        for (uint32_t pc = 0; pc < NH_BCHUNK_STRIDE; pc += 16u)
          if (lo + pc < my_bytes) cp_async16(dst + pc, src0 + lo + pc);
        cp_async_arrive(bar0 + buf * 8u);
      };
      if (n_chunks) stage(0u);
      for (uint32_t ch = 0; ch < n_chunks; ch++) {
        if (ch + 1u < n_chunks) stage(ch + 1u); /* its buffer was read two chunks ago (syncwarp below) */
        const uint32_t buf = ch & 1u;
        mbar_wait(bar0 + buf * 8u, (bar_par >> buf) & 1u);
        bar_par ^= 1u << buf;
        const uint32_t wbase = win0 + buf * (32u * NH_BCHUNK_STRIDE) + a16;
        const uint32_t w0 = ch * NH_BCHUNK_WORDS;
        const uint32_t wn = max_words - w0 < NH_BCHUNK_WORDS ? max_words - w0 : NH_BCHUNK_WORDS;
        uint32_t w_next = lds_u32(wbase);
#pragma unroll 1
        for (uint32_t w = 0; w < wn; w++) {
          const uint32_t word = w_next;
          w_next = lds_u32(wbase + (((w + 1u) & (NH_BCHUNK_WORDS - 1u)) << 2)); /* one word ahead; the wrap is a harmless re-read */
          scan_word(word, (w0 + w) * 4u);
        }
        __syncwarp(); /* every lane is done with this buffer before chunk ch + 2 lands in it */
      }
#else
      /* words are loaded NH_STREAM_PREFETCH iterations ahead of their use: a lane's 4-byte load is
       * its own request into a memory system kept busy by the random table reads */
      uint32_t w_q[NH_STREAM_PREFETCH];
#pragma unroll
      for (int pf = 0; pf < NH_STREAM_PREFETCH; pf++) w_q[pf] = (uint32_t)pf < my_words ? __ldg(q + pf) : 0u;
      for (uint32_t base_i = 0; base_i < max_words * 4u; base_i += 4u) {
        const uint32_t word = w_q[0];
#pragma unroll
        for (int pf = 0; pf + 1 < NH_STREAM_PREFETCH; pf++) w_q[pf] = w_q[pf + 1];
        const uint32_t wi = (base_i >> 2) + NH_STREAM_PREFETCH;
        w_q[NH_STREAM_PREFETCH - 1] = wi < my_words ? __ldg(q + wi) : 0u;
        scan_word(word, base_i);
      }
#endif
      emit(cnt != 0u, last, cnt);
      if (scanning) {
        NhTileOut o;
        o.lk_off = t.slot;
        o.lk_cnt = n_runs;
        b.tile_out[tile] = o;
      }
      if (rescan) break;
      tot_lookups += n_runs;
      /* drain: whatever is waiting or in flight */
      while (pq_n + cq_n != 0u || any_inflight) probe_round();
      const uint32_t ov = sm.overflow; /* final: every fold happened before the last __syncwarp */
      if (ov == 0u) break;
      redo_lane = have && ((t.role == NH_ROLE_PARTNER ? (ov >> (lane - 1u)) : (ov >> lane)) & 1u) && !spill_runs;
    }

    /* ---------------- tiles of deferred units: hand the tile's table to k_score ---------------- */
    if (have && t.role == NH_ROLE_DEFERRED) {
      const bool ovf = (sm.overflow >> lane) & 1u;
      NhTileSum ts;
      ts.first_min = n_runs ? sm.first_min[lane] : NH_NONE64;
      ts.last_min = n_runs ? sm.last_min[lane] : NH_NONE64;
      ts.groups = sm.groups[lane];
      ts.flags = (n_runs ? NH_TILE_HAS : 0u) | (((sm.first_hit >> lane) & 1u) ? NH_TILE_FIRST_HIT : 0u) |
                 (ovf ? NH_TILE_OVERFLOW : 0u);
      b.tile_sum[tile] = ts;
      if (!ovf) {
        NhTileTab tt;
#pragma unroll
        for (int i = 0; i < NH_LANE_TAXA; i++) {
          tt.keys[i] = sm.keys[i * 32 + lane];
          tt.cnts[i] = sm.cnts[i * 32 + lane];
        }
        b.tile_tab[tile] = tt;
      }
    }

    /* ---------------- score short units in the warp ---------------- */
    if (have && (t.role == NH_ROLE_LEADER || t.role == NH_ROLE_LEADER2)) {
      const uint32_t u = b.paired ? (t.seq >> 1) : t.seq;
      if ((sm.overflow >> lane) & 1u) {
        /* more distinct taxa than a lane table holds: k_score_big probes the unit again;
         * it needs the lookups in global memory, which only spilled tiles have */
        b.overflow_units[atomicAdd(&b.counters->n_overflow, 1u)] = u | NH_OVERFLOW_REPROBE;
      } else {
        uint32_t ntab = 0;
        while (ntab < (uint32_t)sp.lane_taxa && sm.keys[ntab * 32u + lane] != 0u) ntab++;
        const int groups = (int)sm.groups[lane];
        const uint32_t s0 = b.paired ? (t.seq & ~1u) : t.seq;
        uint32_t total_kmers = 0;
        for (uint32_t mm = 0; mm < (b.paired ? 2u : 1u); mm++) {
          const uint64_t len = b.offsets[s0 + mm + 1] - b.offsets[s0 + mm];
          if (len >= (uint64_t)k) total_kmers += (uint32_t)(len - (uint64_t)k + 1);
        }
        uint32_t best_s = 0, best_t = 0;
        for (uint32_t i = 0; i < ntab; i++) {
          const uint32_t tx = sm.keys[i * 32u + lane];
          uint32_t score = 0;
          for (uint32_t a = tx; a; a = parent[a]) score += lane_tab_get(sm, lane, ntab, a);
          if (score > best_s) {
            best_s = score;
            best_t = tx;
          } else if (score == best_s) {
            best_t = lca(parent, best_t, tx);
          }
        }
        uint32_t max_taxon = best_t;
        uint32_t max_score = max_taxon ? lane_tab_get(sm, lane, ntab, max_taxon) : 0u;
        const uint32_t required = (uint32_t)ceil(__dmul_rn(sp.confidence, (double)total_kmers));
        while (max_taxon && max_score < required) {
          uint32_t sum = 0;
          for (uint32_t i = 0; i < ntab; i++)
            if (is_a_ancestor_of_b(parent, max_taxon, sm.keys[i * 32u + lane]))
              sum += sm.cnts[i * 32u + lane];
          max_score = sum;
          if (max_score >= required) break;
          max_taxon = parent[max_taxon];
        }
        uint32_t call = max_taxon;
        if (call && groups < sp.min_hit_groups) call = 0;
        const uint32_t is_cls = call != 0u;
        const uint32_t keep = sp.keep_human ? is_cls : !is_cls;
        if (b.out_call) b.out_call[u] = call ? db.ext_id[call] : 0u;
        if (b.out_keep) b.out_keep[u] = (uint8_t)keep;
        if (b.dbg_call) b.dbg_call[u] = call;
        if (b.dbg_total_kmers) b.dbg_total_kmers[u] = total_kmers;
        if (b.dbg_hit_groups) b.dbg_hit_groups[u] = (uint32_t)groups;
        tot_classified += is_cls;
        tot_kept += keep;
      }
    }
    __syncwarp();
  }
  tot_lookups = warp_sum_u32(tot_lookups);
  tot_classified = warp_sum_u32(tot_classified);
  tot_kept = warp_sum_u32(tot_kept);
  if (lane == 0) {
    if (tot_lookups) atomicAdd(&b.counters->n_lookups, tot_lookups);
    if (tot_classified) atomicAdd(&b.counters->n_classified, tot_classified);
    if (tot_kept) atomicAdd(&b.counters->n_kept, tot_kept);
  }
}

/* ------------------------------------------------------------------ */
/* per-read output support: pack every tile's (external taxid, run length) */

__global__ void __launch_bounds__(NH_BLOCK_THREADS)
k_gather_runs(const NhDbParams db, const NhBatchPtrs b, uint32_t *__restrict__ run_ext,
              uint8_t *__restrict__ run_len, uint32_t *__restrict__ tile_run_off,
              uint32_t *__restrict__ cursor) {
  const uint32_t lane = lane_id(), warp = threadIdx.x >> 5;
  const uint32_t n_tiles = b.counters->n_tiles;
  for (uint32_t tile = blockIdx.x * NH_WARPS_PER_BLOCK + warp; tile < n_tiles;
       tile += gridDim.x * NH_WARPS_PER_BLOCK) {
    const NhTileOut to = b.tile_out[tile];
    uint32_t base = 0;
    if (lane == 0) base = to.lk_cnt ? atomicAdd(cursor, to.lk_cnt) : 0u;
    base = __shfl_sync(FULL_MASK, base, 0);
    if (lane == 0) tile_run_off[tile] = base;
    for (uint32_t j = lane; j < to.lk_cnt; j += 32u) {
      run_ext[base + j] = db.ext_id[b.lk_taxon[to.lk_off + j]];
      run_len[base + j] = b.lk_cnt[to.lk_off + j];
    }
  }
}

/* ------------------------------------------------------------------ */
/* roofline helper: uniformly random aligned 32-byte sector reads       */

__global__ void __launch_bounds__(256)
k_random_gather(const uint32_t *__restrict__ cells, uint64_t n_sectors, uint64_t n_reads,
                uint64_t seed, uint32_t *__restrict__ sink) {
  const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
  uint32_t acc = 0;
  for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_reads; i += stride) {
    const uint64_t h = nh_fmix64(i + seed);
    const uint64_t sec = __umul64hi(h, n_sectors);
    uint32_t c[8];
    ld_sector(cells + sec * 8ULL, c);
    acc ^= c[0] ^ c[1] ^ c[2] ^ c[3] ^ c[4] ^ c[5] ^ c[6] ^ c[7];
  }
  if (acc == 0x9E3779B9u) sink[0] = acc; /* keeps the loads alive */
}

/* ------------------------------------------------------------------ */
/* launchers                                                            */

static int g_big_smem_ok = 0;

cudaError_t nh_kernels_init(void) {
  cudaError_t e = cudaFuncSetAttribute(k_score_big, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       NH_BIG_HASH_SLOTS * 8);
  if (e != cudaSuccess) return e;
  e = cudaFuncSetAttribute(k_score<true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                           NH_SMEM_PARENT_MAX * 4 + NH_WARPS_PER_BLOCK * NH_WARP_HASH_SLOTS * 8);
  if (e != cudaSuccess) return e;
  const int fused_max = NH_SMEM_PARENT_MAX * 4 + NH_WARPS_PER_BLOCK * (int)sizeof(FusedWarpSmem);
  const int stream_max = NH_SMEM_PARENT_MAX * 4 + NH_WARPS_PER_BLOCK * (int)sizeof(StreamWarpSmem);
  e = cudaFuncSetAttribute(k_stream_classify<5, false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, stream_max);
  if (e != cudaSuccess) return e;
  e = cudaFuncSetAttribute(k_stream_classify<5, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, stream_max);
  if (e != cudaSuccess) return e;
  e = cudaFuncSetAttribute(k_stream_classify<5, true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, stream_max);
  if (e != cudaSuccess) return e;
  e = cudaFuncSetAttribute(k_scan_probe_score<5, 1, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, fused_max);
  if (e != cudaSuccess) return e;
  e = cudaFuncSetAttribute(k_scan_probe_score<5, 1, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, fused_max);
  if (e != cudaSuccess) return e;
  e = cudaFuncSetAttribute(k_scan_probe_score<5, 2, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, fused_max);
  if (e != cudaSuccess) return e;
  e = cudaFuncSetAttribute(k_scan_probe_score<5, 4, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, fused_max);
  if (e != cudaSuccess) return e;
  g_big_smem_ok = 1;
  return cudaSuccess;
}

int nh_launch_plan(const NhDbParams &db, const NhBatchPtrs &b, cudaStream_t st) {
  const uint32_t nb = (b.n_seqs + PLAN_THREADS - 1) / PLAN_THREADS;
  k_plan_count<<<nb, PLAN_THREADS, 0, st>>>(b.offsets, b.n_seqs, db.k, db.tile_pos, b.block_sums);
  k_plan_scan<<<1, PLAN_THREADS, 0, st>>>(b.block_sums, nb, b.counters);
  k_plan_fill<<<nb, PLAN_THREADS, 0, st>>>(b.offsets, b.n_seqs, db.k, db.tile_pos, b.paired,
                                           b.deferred_units != nullptr, b.block_sums, b.tile_base,
                                           b.tiles, b.deferred_units, b.counters);
  return 3;
}

bool nh_fused_supported(const NhDbParams &db) {
  /* window of 5 l-mers, u8 run lengths, u32 group indices (tables below 128 GiB) */
  return db.w == 5 && db.tile_pos <= 255 && db.capacity < (1ULL << 35) - 8ULL; /* sector index 0xFFFFFFFF is "none" */
}

static size_t fused_smem_bytes(const NhDbParams &db) {
  const uint32_t parent_words = db.node_count <= NH_SMEM_PARENT_MAX ? db.node_count : 0u;
  return (size_t)((parent_words + 3u) & ~3u) * 4 + NH_WARPS_PER_BLOCK * sizeof(FusedWarpSmem);
}

int nh_launch_fused(const NhDbParams &db, const NhBatchPtrs &b, const NhScoreParams &sp,
                    uint32_t tiles_upper, int sm_count, int form, cudaStream_t st) {
  uint32_t groups = (tiles_upper + 31u) / 32u;
  uint32_t blocks = (groups + NH_WARPS_PER_BLOCK - 1) / NH_WARPS_PER_BLOCK;
  uint32_t max_grid = (uint32_t)sm_count * NH_FUSED_MIN_BLOCKS;
  uint32_t grid = blocks < max_grid ? blocks : max_grid;
  if (grid == 0) grid = 1;
  /* NH_PROBE_LANES=1|2|4: lanes (adjacent sectors) per lookup.  2 and 4 cut the requests per
   * lookup from 1.41 to 1.24 / 1.12 but cost more issue slots than they save (measured:
   * 3.67 / 3.81 / 5.27 ms per 1 M pairs), so the default is 1.
   * NH_PROBE_DEPTH=1|2: sector reads in flight per lane and round. */
  static int lanes = 0, depth = 0;
  if (!lanes) {
    const char *e = getenv("NH_PROBE_LANES");
    lanes = e ? atoi(e) : 1;
    if (lanes != 1 && lanes != 2 && lanes != 4) lanes = 1;
    const char *d = getenv("NH_PROBE_DEPTH");
    depth = d ? atoi(d) : NH_PROBE_DEPTH_DEFAULT;
    if (depth != 1 && depth != 2) depth = NH_PROBE_DEPTH_DEFAULT;
  }
  const bool stream = form == 2;
  if (stream) {
    const uint32_t smax = (uint32_t)sm_count * NH_STREAM_MIN_BLOCKS;
    grid = blocks < smax ? blocks : smax;
    if (grid == 0) grid = 1;
    const uint32_t parent_words = db.node_count <= NH_SMEM_PARENT_MAX ? db.node_count : 0u;
    const size_t ssmem = (size_t)((parent_words + 3u) & ~3u) * 4 + NH_WARPS_PER_BLOCK * sizeof(StreamWarpSmem);
    if (b.dbg_pos_min != nullptr) /* nh_debug_minimizers */
      k_stream_classify<5, true, false><<<grid, NH_BLOCK_THREADS, ssmem, st>>>(db, b, sp);
    else if (db.revcom_version == 0)
      k_stream_classify<5, false, true><<<grid, NH_BLOCK_THREADS, ssmem, st>>>(db, b, sp);
    else
      k_stream_classify<5, false, false><<<grid, NH_BLOCK_THREADS, ssmem, st>>>(db, b, sp);
    return 1;
  }
  const size_t smem = fused_smem_bytes(db);
  if (lanes == 1 && depth == 2)
    k_scan_probe_score<5, 1, 2><<<grid, NH_BLOCK_THREADS, smem, st>>>(db, b, sp);
  else if (lanes == 1)
    k_scan_probe_score<5, 1, 1><<<grid, NH_BLOCK_THREADS, smem, st>>>(db, b, sp);
  else if (lanes == 2)
    k_scan_probe_score<5, 2, 1><<<grid, NH_BLOCK_THREADS, smem, st>>>(db, b, sp);
  else
    k_scan_probe_score<5, 4, 1><<<grid, NH_BLOCK_THREADS, smem, st>>>(db, b, sp);
  return 1;
}

int nh_launch_minimizers(const NhDbParams &db, const NhBatchPtrs &b, uint32_t tiles_upper,
                         int sm_count, cudaStream_t st) {
  uint32_t groups = (tiles_upper + NH_WARPS_PER_BLOCK - 1) / NH_WARPS_PER_BLOCK;
  uint32_t max_grid = (uint32_t)sm_count * 8u;
  uint32_t grid = groups < max_grid ? groups : max_grid;
  if (grid == 0) grid = 1;
  if (db.w == 5)
    k_minimizers<5><<<grid, NH_BLOCK_THREADS, 0, st>>>(db, b);
  else
    k_minimizers<0><<<grid, NH_BLOCK_THREADS, 0, st>>>(db, b);
  return 1;
}

int nh_launch_probe(const NhDbParams &db, const uint64_t *keys, uint32_t *taxa,
                    const uint32_t *n_dev, uint32_t n_upper, int sm_count, cudaStream_t st) {
  uint32_t blocks = (n_upper + 255u) / 256u;
  uint32_t max_grid = (uint32_t)sm_count * 8u;
  uint32_t grid = blocks < max_grid ? blocks : max_grid;
  if (grid == 0) grid = 1;
  k_probe<<<grid, 256, 0, st>>>(db, keys, taxa, n_dev, n_upper);
  return 1;
}

int nh_launch_score(const NhDbParams &db, const NhBatchPtrs &b, const NhScoreParams &sp,
                    int sm_count, cudaStream_t st) {
  uint32_t blocks = (b.n_units + NH_WARPS_PER_BLOCK - 1) / NH_WARPS_PER_BLOCK;
  uint32_t max_grid = (uint32_t)sm_count * 8u;
  uint32_t grid = blocks < max_grid ? blocks : max_grid;
  if (grid == 0) grid = 1;
  const size_t hash_bytes = NH_WARPS_PER_BLOCK * NH_WARP_HASH_SLOTS * 8;
  if (db.node_count <= NH_SMEM_PARENT_MAX)
    k_score<true><<<grid, NH_BLOCK_THREADS, db.node_count * 4 + hash_bytes, st>>>(db, b, sp);
  else
    k_score<false><<<grid, NH_BLOCK_THREADS, hash_bytes, st>>>(db, b, sp);
  k_score_big<<<sm_count, 32, NH_BIG_HASH_SLOTS * 8, st>>>(db, b, sp);
  return 2;
}

int nh_launch_gather_runs(const NhDbParams &db, const NhBatchPtrs &b, uint32_t tiles_upper,
                          uint32_t *run_ext, uint8_t *run_len, uint32_t *tile_run_off,
                          uint32_t *cursor, int sm_count, cudaStream_t st) {
  uint32_t blocks = (tiles_upper + NH_WARPS_PER_BLOCK - 1) / NH_WARPS_PER_BLOCK;
  uint32_t max_grid = (uint32_t)sm_count * 8u;
  uint32_t grid = blocks < max_grid ? blocks : max_grid;
  if (grid == 0) grid = 1;
  cudaMemsetAsync(cursor, 0, 4, st);
  k_gather_runs<<<grid, NH_BLOCK_THREADS, 0, st>>>(db, b, run_ext, run_len, tile_run_off, cursor);
  return 1;
}

int nh_launch_random_gather(const uint32_t *cells, uint64_t n_sectors, uint64_t n_reads,
                            uint64_t seed, uint32_t *sink, int sm_count, cudaStream_t st) {
  k_random_gather<<<sm_count * 8, 256, 0, st>>>(cells, n_sectors, n_reads, seed, sink);
  return 1;
}
