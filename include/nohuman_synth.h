/*
 * nohuman_synth.h — synthetic-workload tooling of libnohuman_gpu.so.
 *
 * The HPRC databases nohuman downloads (reference config.toml:1-19) are not
 * available offline, so benchmarks use a kraken2-format table built on the
 * GPU from a synthetic "pangenome" and synthetic reads sampled from it
 * (BASELINE.json north_star).  Nothing here is on the
 * classification path; the table it produces is an ordinary CompactHashTable
 * (hash.k2d cell layout) that the oracle can read back.
 */
#ifndef NOHUMAN_SYNTH_H
#define NOHUMAN_SYNTH_H

#include "nohuman_gpu.h"

#ifdef __cplusplus
extern "C" {
#endif

typedef struct {
  uint64_t capacity;       /* cells of the table to build */
  double target_load;      /* stop adding genome once size/capacity reaches this (kraken2-build: 0.7) */
  uint64_t genome_seed;
  uint64_t block_bases;    /* genome is cut into blocks; block b belongs to leaf b % n_leaves */
  double overlap_frac;     /* tail fraction of a block that also belongs to the next leaf */
  uint64_t max_genome_bases; /* 0: unlimited */
} nh_synth_db_params_t;

/* Build a database on `device` from the synthetic genome.  opts / taxo are
 * memory images of opts.k2d / taxo.k2d; leaf_taxa are INTERNAL taxonomy ids.
 * On return *genome_bases is the genome length that was inserted. */
int nh_synth_build_db(const void *opts, size_t opts_len, const void *taxo, size_t taxo_len,
                      const uint32_t *leaf_taxa, int n_leaves, const nh_synth_db_params_t *p,
                      int device, nh_db **out, uint64_t *genome_bases);

/* Copy the resident table back to the host (capacity x u32). */
int nh_db_download_cells(const nh_db *db, uint32_t *out_cells);

typedef struct {
  uint64_t seed;
  uint64_t genome_seed;
  uint64_t genome_bases;
  double human_frac;     /* fraction of units sampled from the genome; the rest are uniform random */
  double sub_rate, ins_rate, del_rate;
  double n_rate;         /* probability that a sequence carries one 'N' */
  int32_t paired;        /* sequences 2u, 2u+1 are mates of one fragment */
  int32_t reserved;
  double insert_mean, insert_sd;
} nh_synth_reads_params_t;

/* Fill d_bases (device) for the sequences described by d_offsets (device,
 * n_seqs+1).  Lengths are whatever the offsets say. */
int nh_synth_reads(int device, uint8_t *d_bases, const uint64_t *d_offsets, uint64_t n_seqs,
                   const nh_synth_reads_params_t *p, void *cuda_stream);

/* The synthetic genome itself (host copy of bases [start, start+n)). */
int nh_synth_genome(int device, uint64_t genome_seed, uint64_t start, uint64_t n, uint8_t *out_host);

#ifdef __cplusplus
}
#endif
#endif
