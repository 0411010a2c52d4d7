/*
 * nh_kernels.cuh — device-side parameter blocks and launch prototypes of the
 * four classification stages (sm_100a).  See DESIGN.md for the data layout.
 */
#ifndef NH_KERNELS_CUH
#define NH_KERNELS_CUH

#include <cuda_runtime.h>
#include <stdint.h>

#include "nh_math.h"

#ifndef NH_WARPS_PER_BLOCK
#define NH_WARPS_PER_BLOCK 8
#endif
#define NH_BLOCK_THREADS (NH_WARPS_PER_BLOCK * 32)
#define NH_TILE_LMERS 128           /* l-mers per minimizer tile of the warp-per-tile kernel (4 warp iterations) */
#define NH_FUSED_TILE_POS 512       /* k-mer positions per tile of the lane-serial kernel: 2x300 bp reads stay one tile */
#define NH_FUSED_TILE_POS_LONG 256  /* ... for batches of long reads: more, smaller groups balance better
                                     * (both multiples of 32: tiles of packed input start on a 32-base unit) */
#define NH_FUSED_TILE_POS_MAX 1023  /* run lengths travel in 10 bits */
#define NH_MAX_WINDOW 32            /* k - l + 1 must fit one warp */
#define NH_SMEM_PARENT_MAX 8192     /* taxonomy nodes staged in shared memory */
#define NH_WARP_HASH_SLOTS 64       /* per-read taxon->count table (fast path) */
#define NH_BIG_HASH_SLOTS 16384     /* overflow path: one warp per block */
#define NH_NONE64 0xFFFFFFFFFFFFFFFFULL

/* Database constants every kernel needs (passed by value). */
struct NhDbParams {
  const uint32_t *cells;     /* capacity x u32, resident in HBM */
  uint64_t capacity;
  uint64_t mod_m;            /* nh_fastmod constants for capacity */
  uint32_t mod_sh1, mod_sh2;
  uint32_t value_bits;
  uint32_t value_mask;
  int32_t k, l, w;           /* w = k - l + 1 */
  int32_t tile_pos;          /* k-mer positions per tile (warp-per-tile kernels: NH_TILE_LMERS - (w-1)) */
  int32_t legacy_tile_pos;   /* always NH_TILE_LMERS - (w-1): what minimizer_tile can stage */
  int32_t amb_span;          /* max(l, k-1): bases whose ambiguity voids a position */
  int32_t revcom_version;
  uint64_t seed_mask;        /* spaced_seed_mask, or the l-mer mask when it is 0 */
  uint64_t toggle;           /* toggle_mask & lmer_mask */
  uint64_t min_hash;         /* minimum_acceptable_hash_value */
  const uint32_t *parent;    /* node_count x u32 */
  const uint32_t *ext_id;    /* node_count x u32 */
  uint32_t node_count;
  /* taxonomies with more nodes than k_score_big's shared table could ever need: one table in global
   * memory per device, 2 x huge_slots words + a lock word, used by one unit at a time */
  uint32_t huge_slots;       /* power of two >= 2 * node_count, or 0 */
  uint32_t *huge_table;
};

struct __align__(16) NhTile {
  uint32_t seq;       /* sequence index in the batch */
  uint32_t pos_begin; /* first k-mer position of the tile */
  uint32_t slot;      /* first lookup slot of the tile = k-mer positions before it in the batch */
  uint32_t role;      /* who scores the tile's unit inside k_stream_classify, NH_ROLE_* */
};

/* role of a tile in the streaming kernel: a unit (read or pair) is scored inside the warp when all
 * its tiles sit in one group of 32; its first tile LEADs, the others are MEMBERs that fold their
 * hits into the leader's table */
#define NH_ROLE_DEFERRED 0u /* unit goes to k_score (more than 32 tiles, or split across groups) */
#define NH_ROLE_LEADER 1u   /* first tile of a unit scored in the warp */
#define NH_ROLE_MEMBER 2u   /* | (tile index - index of the unit's first tile) << 8 */
#define NH_ROLE_KIND(r) ((r) & 0xFFu)
#define NH_ROLE_DELTA(r) ((r) >> 8)
#define NH_LANE_TAXA 8      /* per-lane taxon->count slots in the streaming kernel */

struct NhTileOut {
  uint32_t lk_off; /* first lookup of the tile in the lookup arrays */
  uint32_t lk_cnt; /* number of lookups (distinct-consecutive minimizers) */
};

/* What the streaming kernel leaves behind for a tile of a deferred (multi-tile) unit instead of
 * its individual lookups: the tile's own taxon -> k-mer count table and what k_score needs to
 * stitch tiles together. */
struct __align__(16) NhTileTab {
  uint32_t keys[NH_LANE_TAXA]; /* internal taxids, 0-terminated */
  uint32_t cnts[NH_LANE_TAXA];
};
struct __align__(8) NhTileSum {
  uint64_t first_min, last_min; /* first / last distinct minimizer of the tile */
  uint32_t groups;              /* lookups of the tile that hit */
  uint32_t flags;               /* NH_TILE_* */
};
#define NH_TILE_HAS 1u       /* the tile has at least one lookup */
#define NH_TILE_FIRST_HIT 2u /* its first lookup hit (needed for the group count across tile borders) */
#define NH_TILE_OVERFLOW 4u  /* more distinct taxa than the table holds: k_score_big scans the unit again */

/* Device-side counters of one batch. */
struct NhCounters {
  uint32_t n_tiles;
  uint32_t n_lookups;    /* allocation cursor of the lookup arrays */
  uint32_t n_classified;
  uint32_t n_kept;
  uint32_t n_overflow;   /* units sent to the big-table scoring pass */
  uint32_t error;        /* nonzero: a unit exceeded even the big table */
  uint32_t n_deferred;   /* units k_score handles (not scored inside the fused kernel) */
  uint32_t next_group;   /* streaming kernel: next group of 32 tiles to hand out */
  uint32_t n_sector_reads; /* streaming kernel: table sectors requested (lookups + chain continuations) */
  uint32_t pad[3];
};

struct NhBatchPtrs {
  const uint8_t *bases;     /* ASCII, 1 byte per base; null when the batch came packed */
  const uint64_t *offsets;  /* n_seqs + 1 */
  /* packed input (nh_classify_batch_packed): every sequence starts on a unit of 32 bases */
  const uint8_t *codes;     /* 2-bit codes, 4 bases per byte, first base in the top bits: 8 bytes per unit */
  const uint32_t *valid;    /* 1 bit per base, LSB first, 1 = unambiguous base inside the sequence: one word per unit */
  const uint32_t *poff;     /* n_seqs + 1: first unit of each sequence */
  uint32_t n_seqs;
  uint32_t n_units;
  int32_t paired;
  /* plan */
  uint32_t *tile_base;      /* n_seqs + 1: first tile of each sequence */
  uint2 *seq_info;          /* null, or (batches of long reads) per sequence: lookup slot base, first tile of its in-warp unit */
  uint32_t tiles_upper;     /* host-side bound on the number of tiles */
  uint64_t *block_sums;
  NhTile *tiles;
  NhTileOut *tile_out;
  /* lookups */
  uint64_t *lk_min;
  uint16_t *lk_cnt;         /* k-mer positions that take this lookup's taxon */
  uint32_t *lk_taxon;
  /* results */
  uint32_t *out_call;       /* external taxid per unit (may be null) */
  uint8_t *out_keep;        /* may be null */
  uint32_t *dbg_call;       /* internal id per unit (may be null) */
  uint32_t *dbg_total_kmers;
  uint32_t *dbg_hit_groups;
  uint32_t *overflow_units;
  uint32_t *deferred_units; /* null: k_score walks every unit (legacy path) */
  int32_t emit_all_taxa;    /* streaming kernel: store lk_cnt / lk_taxon of every lookup (per-read output wanted) */
  NhTileTab *tile_tab;      /* streaming kernel: per-tile tables of deferred units (null: legacy path, lookups in lk_*) */
  NhTileSum *tile_sum;
  NhCounters *counters;
  /* per-position debug output of the minimizer kernel (may be null) */
  const uint64_t *dbg_pos_offsets;
  uint64_t *dbg_pos_min;
  uint8_t *dbg_pos_ambig;
};

struct NhScoreParams {
  double confidence;
  int32_t min_hit_groups;
  int32_t keep_human;
  int32_t lane_taxa;   /* fused kernel: taxon slots per unit before it overflows (<= NH_LANE_TAXA) */
  int32_t filter_mode; /* fused kernel, who asks the miss filter before the table: 0 nobody, 1 units without a hit so far,
                        * 2 every lookup, 3 units whose last NH_FILTER_RECENT lookups all missed */
  /* the miss filter (nh_kernels.cu, k_filter_build): one 32-byte record per block of 32 cells, or null.  It lives
   * here, in the LAST kernel parameter, on purpose: growing NhDbParams by these 16 bytes made ptxas rematerialise
   * addresses all over k_stream_classify (2496 -> 2616 SASS instructions, +12 % executed) */
  const uint32_t *filter;
  uint32_t n_filter_blocks;
};

/* launchers (nh_kernels.cu); each returns the number of kernels launched */
int nh_launch_plan(const NhDbParams &db, const NhBatchPtrs &b, cudaStream_t st);
/* sequence lengths -> base offsets and first units (sums: 2 * ceil(n_seqs / 1024) words of scratch); returns the launches */
/* builds the miss filter of db.cells: n_blocks = ceil(capacity / 32) records of 32 bytes */
void nh_launch_filter_build(const NhDbParams &db, uint32_t *filter, uint32_t n_blocks, cudaStream_t st);
int nh_launch_len_scan(const uint32_t *len, uint32_t n_seqs, uint64_t *sums, uint64_t *off, uint32_t *poff, cudaStream_t st);
/* streaming path: lane-serial minimizer scan feeding the probe, in-warp scoring */
bool nh_fused_supported(const NhDbParams &db);
int nh_launch_stream(const NhDbParams &db, const NhBatchPtrs &b, const NhScoreParams &sp,
                     uint32_t tiles_upper, int sm_count, cudaStream_t st);
int nh_launch_minimizers(const NhDbParams &db, const NhBatchPtrs &b, uint32_t tiles_upper,
                         int sm_count, cudaStream_t st);
int nh_launch_probe(const NhDbParams &db, const uint64_t *keys, uint32_t *taxa,
                    const uint32_t *n_dev, uint32_t n_upper, int sm_count, cudaStream_t st);
int nh_launch_score(const NhDbParams &db, const NhBatchPtrs &b, const NhScoreParams &sp,
                    int sm_count, cudaStream_t st);
/* per-tile runs (external taxid, k-mer count) packed densely for the per-read kraken output */
int nh_launch_gather_runs(const NhDbParams &db, const NhBatchPtrs &b, uint32_t tiles_upper,
                          uint32_t *run_ext, uint16_t *run_len, uint32_t *tile_run_off,
                          uint32_t *cursor, int sm_count, cudaStream_t st);
int nh_launch_probe_pattern(const uint32_t *cells, uint64_t n_sectors, int lanes, int depth, int blocks_per_sm,
                            uint32_t items_per_chain, uint32_t p_thresh, uint64_t seed, uint64_t sm_window_sectors,
                            unsigned long long *counters, uint32_t *sink, int sm_count, cudaStream_t st);
cudaError_t nh_kernels_init(void);

#endif
