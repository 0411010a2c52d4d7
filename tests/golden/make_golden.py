#!/usr/bin/env python
"""Regenerates tests/golden/cfg1_tiny.npz: a tiny kraken2-format database
(cfg1 shape: synthetic human chr + 3 bacterial genomes, 23-node taxonomy,
k=35 l=31 s=7), seeded reads, and the per-read results the oracle gave for
them when the fixture was frozen.

The reference (mbhall88/nohuman) execs kraken2 and holds no vector for this
path, and no kraken2 binary exists offline, so these vectors pin the oracle
and the CUDA path to each other and to this commit's behaviour (regression
pin), not to upstream kraken2 — "parity unpinned", see oracle/k2_oracle.h.

    python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

import synth  # noqa: E402
from oracle import k2oracle  # noqa: E402


def main():
    genomes = synth.cfg1_genomes(seed=1, scale=0.002)  # 10 kb human + 4/6/8 kb bacteria
    tax = [k2oracle.TaxSpec(*t) for t in synth.TAXONOMY_CFG1]
    db = k2oracle.OracleDb.build([(t, bytes(g)) for t, g in genomes], tax)
    d = os.path.join(HERE, "_tmp_db")
    db.save(d)
    files = {n: np.frombuffer(open(os.path.join(d, n), "rb").read(), np.uint8)
             for n in ("hash.k2d", "opts.k2d", "taxo.k2d")}
    for n in files:
        os.remove(os.path.join(d, n))
    os.rmdir(d)
    out = {"hash_k2d": files["hash.k2d"], "opts_k2d": files["opts.k2d"], "taxo_k2d": files["taxo.k2d"]}
    rng = np.random.default_rng(99)
    se = synth.illumina_reads(genomes, 400, 150, seed=5, n_rate=0.1)
    se += [synth.random_genome(rng, int(L)) for L in (0, 1, 34, 35, 36, 123, 124, 125, 158, 159, 300)]
    pe = synth.illumina_reads(genomes, 300, 150, seed=6, paired=True, n_rate=0.05)
    ont = synth.ont_reads(genomes, 40, seed=7, n50=1500, max_len=6000)
    for name, seqs, paired, confs in (("se", se, False, (0.0, 0.1, 0.5)), ("pe", pe, True, (0.0, 0.5)),
                                      ("ont", ont, False, (0.0, 0.05))):
        bases, offsets = synth.pack(seqs)
        out[f"{name}_bases"] = bases
        out[f"{name}_offsets"] = offsets
        for conf in confs:
            db.confidence = conf
            r = db.classify_batch(bases, offsets, paired=paired)
            tag = f"{name}_c{int(round(conf * 100)):03d}"
            out[f"{tag}_ext"] = r["ext"]
            out[f"{tag}_call"] = r["call"]
            out[f"{tag}_hit_groups"] = r["hit_groups"]
            out[f"{tag}_total_kmers"] = r["total_kmers"]
            out[f"{tag}_lookups"] = np.array([r["lookups"]], np.uint64)
    # per-position minimizer stream of the first 20 single-end reads
    mins, ambs = [], []
    for s in se[:20]:
        m, a = k2oracle.scan_positions(db.opts, bytes(s))
        mins.append(m)
        ambs.append(a)
    out["se20_minimizers"] = np.concatenate(mins)
    out["se20_ambiguous"] = np.concatenate(ambs)
    # hitlist strings of kraken2's per-read output for a few reads
    db.confidence = 0.0
    out["se_hitlists"] = np.array([db.classify_one(bytes(s), want_taxa=True)["hitlist"] for s in se[:40]])
    np.savez_compressed(os.path.join(HERE, "cfg1_tiny.npz"), **out)
    print("wrote", os.path.join(HERE, "cfg1_tiny.npz"), os.path.getsize(os.path.join(HERE, "cfg1_tiny.npz")), "bytes")


if __name__ == "__main__":
    main()
