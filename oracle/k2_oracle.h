/*
 * k2_oracle.h — CPU restatement of the kraken2 classification path that
 * nohuman shells out to.  TEST INFRASTRUCTURE ONLY.
 *
 * Who may use this: tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs.  The product (nohuman_b200/) never
 * links, imports or executes anything under oracle/.
 *
 * PARITY UNPINNED: the reference (/root/reference, mbhall88/nohuman v0.5.1)
 * contains no classification arithmetic; it exec()s `kraken2`
 * (src/lib.rs:22-23, src/main.rs:215-270).  The algorithm lives in the
 * third-party dependency DerrickWood/kraken2, pinned only at Dockerfile:15
 * (K2VER="2.17") / Dockerfile:35-38 (commit f885f832c986...).  That source
 * is not vendored and no kraken2 binary exists offline, and the reference's
 * own tests hold no golden vector for this path (src/lib.rs:153-222 only run
 * `ls`).  This file restates kraken2's published algorithm (upstream files
 * src/mmscanner.{h,cc}, kv_store.h, compact_hash.{h,cc}, taxonomy.{h,cc},
 * classify.cc, build_db.cc) as specified in SURVEY.md Appendix A/B.  The
 * known-answer vectors it is checked against are those of SURVEY.md
 * Appendix C plus hand-derived cases in tests/.
 */
#ifndef K2_ORACLE_H
#define K2_ORACLE_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ---- constants (upstream kraken2_data.h / mmscanner.h) ---- */
#define K2O_DEFAULT_TOGGLE_MASK 0xe37e28c4271b5a2dULL
#define K2O_TAXID_MAX ((uint64_t)-1)
#define K2O_MATE_PAIR_BORDER_TAXON (K2O_TAXID_MAX)
#define K2O_READING_FRAME_BORDER_TAXON (K2O_TAXID_MAX - 1)
#define K2O_AMBIGUOUS_SPAN_TAXON (K2O_TAXID_MAX - 2)
#define K2O_CURRENT_REVCOM_VERSION 1

/* ---- opts.k2d: struct IndexOptions, 64 bytes ---- */
typedef struct {
  uint64_t k;
  uint64_t l;
  uint64_t spaced_seed_mask;
  uint64_t toggle_mask;
  uint8_t dna_db;
  uint8_t pad0[7];
  uint64_t minimum_acceptable_hash_value;
  int32_t revcom_version;
  int32_t db_version;
  int32_t db_type;
  int32_t pad1;
} k2o_index_options;

/* ---- taxo.k2d: TaxonomyNode, 56 bytes ---- */
typedef struct {
  uint64_t parent_id;
  uint64_t first_child;
  uint64_t child_count;
  uint64_t name_offset;
  uint64_t rank_offset;
  uint64_t external_id;
  uint64_t godparent_id;
} k2o_taxonomy_node;

typedef struct {
  uint64_t node_count;
  uint64_t name_data_len;
  uint64_t rank_data_len;
  k2o_taxonomy_node *nodes;
  char *name_data;
  char *rank_data;
} k2o_taxonomy;

/* ---- hash.k2d: CompactHashTable ---- */
typedef struct {
  uint64_t capacity;
  uint64_t size;
  uint64_t key_bits;
  uint64_t value_bits;
  uint32_t *cells;
  int owns_cells;
} k2o_cht;

/* ---- primitives ---- */
uint64_t k2o_fmix64(uint64_t key);                                    /* kv_store.h MurmurHash3 */
uint64_t k2o_reverse_complement(uint64_t kmer, int n, int revcom_version); /* mmscanner.cc */
uint64_t k2o_canonical(uint64_t kmer, int n, int revcom_version);
/* seed template '1'*(l-2s) + '01'*s, each bit expanded to 2 bits (build-side) */
uint64_t k2o_spaced_seed_mask(int l, int spaces);

/* ---- MinimizerScanner (mmscanner.{h,cc}) ---- */
typedef struct {
  uint64_t candidate;
  int64_t pos;
} k2o_mmdata;

typedef struct {
  const char *str;
  size_t str_len;
  int64_t k, l;
  size_t str_pos, start, finish;
  uint64_t spaced_seed_mask;
  int dna;
  uint64_t toggle_mask;
  uint64_t lmer, lmer_mask, last_minimizer;
  int64_t loaded_ch;
  k2o_mmdata *queue; /* deque as ring buffer */
  int64_t qhead, qlen, qcap;
  int64_t queue_pos;
  uint64_t last_ambig;
  int revcom_version;
} k2o_scanner;

int k2o_scanner_init(k2o_scanner *s, int64_t k, int64_t l, uint64_t spaced_seed_mask,
                     int dna, uint64_t toggle_mask, int revcom_version);
void k2o_scanner_free(k2o_scanner *s);
void k2o_scanner_load(k2o_scanner *s, const char *seq, size_t len);
/* returns pointer to last_minimizer or NULL when exhausted */
uint64_t *k2o_scanner_next(k2o_scanner *s);
int k2o_scanner_is_ambiguous(const k2o_scanner *s);

/* per-position stream for tests: out_min[i], out_ambig[i] for every
 * NextMinimizer() return; returns number of returns (== max(0,len-k+1)). */
size_t k2o_scan_positions(const k2o_index_options *o, const char *seq, size_t len,
                          uint64_t *out_min, uint8_t *out_ambig, size_t cap);

/* ---- CompactHashTable (compact_hash.cc; LINEAR_PROBING build) ---- */
uint32_t k2o_cht_get(const k2o_cht *t, uint64_t key);
/* same, also reporting cells inspected and distinct 32-byte sectors touched */
uint32_t k2o_cht_get_stats(const k2o_cht *t, uint64_t key, uint64_t *cells, uint64_t *sectors);
/* build-side: value := LCA(existing, taxon) (build_db.cc ProcessSequence loop) */
int k2o_cht_insert_lca(k2o_cht *t, const k2o_taxonomy *tax, uint64_t key, uint32_t taxon);
int k2o_cht_alloc(k2o_cht *t, uint64_t capacity, uint64_t value_bits);
void k2o_cht_free(k2o_cht *t);

/* ---- Taxonomy (taxonomy.cc) ---- */
int k2o_is_a_ancestor_of_b(const k2o_taxonomy *t, uint64_t a, uint64_t b);
uint64_t k2o_lca(const k2o_taxonomy *t, uint64_t a, uint64_t b);
/* Build a kraken-style taxonomy from (ext_id, parent_ext_id) pairs: BFS from
 * the root (the node whose parent is itself or 0) assigns internal ids 1..n
 * so that parent < child; node 0 is the zeroed "unclassified" node. */
int k2o_taxonomy_build(k2o_taxonomy *out, size_t n, const uint64_t *ext_ids,
                       const uint64_t *parent_ext_ids, const char *const *names,
                       const char *const *ranks);
uint64_t k2o_taxonomy_internal_id(const k2o_taxonomy *t, uint64_t ext_id);
void k2o_taxonomy_free(k2o_taxonomy *t);

/* ---- on-disk formats ---- */
int k2o_load_opts(const char *path, k2o_index_options *out);
int k2o_save_opts(const char *path, const k2o_index_options *o);
int k2o_load_taxonomy(const char *path, k2o_taxonomy *out);
int k2o_save_taxonomy(const char *path, const k2o_taxonomy *t);
int k2o_load_cht(const char *path, k2o_cht *out);
int k2o_save_cht(const char *path, const k2o_cht *t);

/* ---- classify.cc: ResolveTree + ClassifySequence ---- */
typedef struct {
  uint64_t call;         /* internal taxid, 0 = unclassified */
  uint64_t ext_call;     /* external id of call */
  uint64_t total_kmers;  /* taxa.size() minus mate border */
  int64_t hit_groups;    /* minimizer_hit_groups */
  uint64_t lookups;      /* hash->Get calls */
  uint64_t cells;        /* cells inspected over those calls */
  uint64_t sectors;      /* 32-byte sectors touched over those calls */
} k2o_read_result;

typedef struct {
  uint64_t *taxon;
  uint32_t *count;
  size_t n, cap;
} k2o_hit_counts;

uint64_t k2o_resolve_tree(k2o_hit_counts *hc, const k2o_taxonomy *tax, uint64_t total_kmers,
                          double confidence);

typedef struct {
  const k2o_index_options *opts;
  const k2o_cht *cht;
  const k2o_taxonomy *tax;
  double confidence;
  int64_t minimum_hit_groups; /* kraken2 wrapper default: 2 */
} k2o_db;

/* taxa_out (optional): receives the per-k-mer `taxa` vector incl. border /
 * ambiguous sentinels; *taxa_n its length; caller frees with free(). */
void k2o_classify_sequence(const k2o_db *db, k2o_scanner *scanner, k2o_hit_counts *hc,
                           const char *seq1, size_t len1, const char *seq2, size_t len2,
                           int paired, k2o_read_result *res, uint64_t **taxa_out,
                           size_t *taxa_n);

/* Batch: sequences concatenated in `bases`, offsets[n_seqs+1]; paired means
 * sequences 2i and 2i+1 are the mates of unit i.  OpenMP over units with
 * `threads` threads (<=0: all).  Outputs are per unit; any may be NULL. */
int k2o_classify_batch(const k2o_db *db, const uint8_t *bases, const uint64_t *offsets,
                       uint64_t n_units, int paired, int threads, uint32_t *out_call_internal,
                       uint32_t *out_call_ext, uint32_t *out_total_kmers,
                       uint32_t *out_hit_groups, uint64_t *out_totals /* [3]: lookups,cells,sectors */);

/* hitlist string of kraken2's per-read output line (classify.cc AddHitlistString) */
char *k2o_hitlist_string(const k2o_taxonomy *tax, const uint64_t *taxa, size_t n);

/* build-side: add every non-ambiguous minimizer of seq under `taxon` (internal id) */
int k2o_build_add_sequence(k2o_cht *t, const k2o_taxonomy *tax, const k2o_index_options *o,
                           const char *seq, size_t len, uint32_t taxon);

#ifdef __cplusplus
}
#endif
#endif
