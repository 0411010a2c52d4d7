/* nh_internal.h — host-side objects behind the opaque C-ABI handles. */
#ifndef NH_INTERNAL_H
#define NH_INTERNAL_H

#include <cuda_runtime.h>
#include <stdint.h>

#include <string>
#include <vector>

#include "../../include/nohuman_gpu.h"
#include "nh_kernels.cuh"

enum {
  EV_H2D0 = 0,
  EV_PLAN0,
  EV_MIN0,
  EV_PROBE0,
  EV_SCORE0,
  EV_SCORE1,
  EV_D2H1,
  NH_NUM_EVENTS
};

struct NhPackPool;
void nh_pack_pool_destroy(NhPackPool *p); /* defined next to the pool (nh_capi.cu) */

struct nh_db {
  nh_db_info_t info{};
  NhDbParams params{};
  uint32_t *d_cells = nullptr;
  bool owns_cells = false;
  uint32_t *d_parent = nullptr;
  uint32_t *d_ext = nullptr;
  uint32_t *d_filter = nullptr; /* miss filter: one 32-byte record per block of 32 cells (k_filter_build) */
  uint32_t n_filter_blocks = 0; /* 0 until the filter is built */
  uint32_t *d_huge = nullptr; /* global-memory taxon table for units with more distinct taxa than shared memory holds */
  std::vector<uint32_t> h_parent, h_ext;
  std::vector<uint64_t> h_ext64;
  std::vector<std::string> h_name, h_rank; /* taxo.k2d name / rank strings per node (reports) */
  int sm_count = 0;
};

struct nh_session {
  nh_db *db = nullptr;
  NhDbParams P{}; /* db->params with this session's tile size */
  nh_params_t params{};
  uint64_t cap_bases = 0, cap_seqs = 0, cap_tiles = 0, cap_lookups = 0;
  size_t device_bytes = 0;
  uint8_t *d_bases = nullptr;
  uint64_t *d_offsets = nullptr;
  uint32_t *d_tile_base = nullptr;
  uint2 *d_seq_info = nullptr;
  uint64_t *d_block_sums = nullptr;
  NhTile *d_tiles = nullptr;
  NhTileOut *d_tile_out = nullptr;
  uint64_t *d_lk_min = nullptr;
  uint16_t *d_lk_cnt = nullptr;
  uint32_t *d_lk_taxon = nullptr;
  uint32_t *d_out_call = nullptr;
  uint8_t *d_out_keep = nullptr;
  uint32_t *d_dbg_call = nullptr, *d_dbg_total = nullptr, *d_dbg_groups = nullptr;
  uint32_t *d_overflow = nullptr;
  uint32_t *d_deferred = nullptr;
  uint32_t *d_run_ext = nullptr, *d_tile_run_off = nullptr, *d_run_cursor = nullptr;
  uint16_t *d_run_len = nullptr;
  uint8_t *d_codes = nullptr;  /* packed input planes (nh_classify_batch_packed), allocated on first use */
  uint32_t *d_valid = nullptr, *d_poff = nullptr;
  bool packed_next = false;    /* the batch being enqueued came packed */
  /* nh_classify_batch_pack: the packer threads and the pinned planes they fill */
  NhPackPool *pack_pool = nullptr;
  int pack_threads = 0;
  uint8_t *h_codes = nullptr;
  uint32_t *h_valid = nullptr, *h_poff = nullptr, *h_len32 = nullptr;
  uint32_t *d_len32 = nullptr; /* sequence lengths as sent; k_len_* rebuild d_offsets and d_poff from them */
  uint64_t *d_len_sums = nullptr;
  cudaEvent_t ev_block = nullptr; /* blocking-sync event: the calling thread sleeps instead of spinning */
  uint64_t last_seqs = 0;
  bool use_fused = false, last_fused = false;
  int lane_taxa = NH_LANE_TAXA;
  int filter_mode = 3;
  int last_form = 0;       /* 0: warp-per-tile kernels, 2: k_stream_classify */
  int forced_tile_pos = 0; /* NH_FUSED_TILE_POS */
  NhTileTab *d_tile_tab = nullptr;
  NhTileSum *d_tile_sum = nullptr;
  NhCounters *d_counters = nullptr;
  NhCounters *h_counters = nullptr; /* pinned */
  cudaStream_t stream = nullptr;
  cudaEvent_t ev[NH_NUM_EVENTS] = {};
  bool pending = false;
  bool timed_copies = false;
  uint32_t last_launches = 0;
  uint64_t last_units = 0, last_bases = 0;
};

/* nh_pack.cc: sequences [s0, s1) of a batch into the planes (poff already holds their first units); AVX2 with
 * non-temporal stores when the CPU has it */
void nh_pack_range(const uint8_t *bases, const uint64_t *offsets, uint64_t total_bases, uint64_t s0, uint64_t s1, uint8_t *codes,
                   uint32_t *valid, const uint32_t *poff);

int nh_set_error(int code, const char *fmt, ...);
int nh_db_build_filter(nh_db *db); /* after the cells are on the device (every open path; the synthetic builder calls it itself) */
int nh_session_create_ex(nh_db *db, const nh_params_t *params, bool need_lookups, nh_session **out);
int nh_resolve_db_dir(const char *db_dir, std::string &out);

#endif
