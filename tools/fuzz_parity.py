#!/usr/bin/env python
"""Randomised parity hunt: many small batches of odd shapes (lengths 0..20 kb around every tile / group border,
N runs, lower case, junk bytes, paired and single, random --conf, random tile sizes, shrunken taxon tables, every
policy of the miss filter, ASCII and packed transfer) through the C ABI against the CPU oracle on the cfg1-shaped database.  Prints the first mismatch with a reproducer seed.

    python tools/fuzz_parity.py [--seconds 120] [--seed 1]
"""
import argparse
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--seconds", type=float, default=120)
    ap.add_argument("--seed", type=int, default=1)
    args = ap.parse_args()
    import synth
    from nohuman_b200 import Database, Session
    from oracle import k2oracle as oracle
    import tempfile
    genomes = synth.cfg1_genomes(seed=1, scale=0.02)
    tax = [oracle.TaxSpec(*t) for t in synth.TAXONOMY_CFG1]
    odb = oracle.OracleDb.build([(t, bytes(g)) for t, g in genomes], tax)
    d = tempfile.mkdtemp(prefix="nh_fuzz_")
    odb.save(d)
    g = dict(genomes)
    keys = list(g)
    t_end = time.time() + args.seconds
    it = 0
    n_units = 0
    while time.time() < t_end:
        seed = args.seed * 1_000_003 + it
        rng = np.random.default_rng(seed)
        tile_pos = int(rng.choice([0, 0, 16, 33, 60, 100, 252, 508, 511, 1023]))
        lane_taxa = int(rng.choice([8, 8, 8, 1, 2, 3]))
        filter_mode = int(rng.choice([3, 3, 0, 1, 2]))   # who asks the miss filter
        entry = str(rng.choice(["ascii", "ascii", "pack"]))  # nh_classify_batch / nh_classify_batch_pack
        paired = bool(rng.integers(0, 2))
        conf = float(rng.choice([0.0, 0.01, 0.1, 0.25, 0.5, 0.9, 1.0]))
        keep_human = bool(rng.integers(0, 2))
        tp = tile_pos or 508
        seqs = []
        n = int(rng.integers(1, 400))
        for _ in range(n):
            kind = rng.integers(0, 10)
            if kind < 3:
                L = int(rng.integers(0, 400))
            elif kind < 6:
                m = int(rng.integers(1, 36))
                L = m * tp + 34 + int(rng.integers(-2, 3))
            elif kind < 8:
                L = int(rng.integers(400, 20000))
            else:
                L = int(rng.choice([0, 1, 30, 31, 34, 35, 36, 66, 150, 151]))
            L = max(0, min(L, 40000))
            src = g[keys[int(rng.integers(0, len(keys)))]]
            mode = rng.integers(0, 6)
            if mode == 0 or L == 0:
                r = synth.random_genome(rng, L)
            else:
                o = int(rng.integers(0, max(1, len(src) - L - 1)))
                r = src[o:o + L].copy()
                if len(r) < L:
                    r = np.concatenate([r, synth.random_genome(rng, L - len(r))])
                if mode == 2:
                    r = synth.mutate(rng, r, float(rng.choice([0.01, 0.05, 0.2])))
                if mode == 3 and L:
                    for _ in range(int(rng.integers(1, 6))):
                        p = int(rng.integers(0, L))
                        r[p:p + int(rng.integers(1, 40))] = ord("N")
                if mode == 4 and L:
                    r[int(rng.integers(0, L)):] |= 0x20  # lower case tail
                if mode == 5 and L > 3:
                    r[int(rng.integers(0, L))] = int(rng.choice([0, 10, 45, 82, 255]))
                    if rng.integers(0, 2):
                        r = synth.revcomp(r) if (r != 0).all() and (r < 128).all() and False else r
            seqs.append(r)
        if paired and len(seqs) % 2:
            seqs.append(seqs[0][:77].copy())
        bases, offsets = synth.pack(seqs)
        if tile_pos:
            os.environ["NH_FUSED_TILE_POS"] = str(tile_pos)
        else:
            os.environ.pop("NH_FUSED_TILE_POS", None)
        os.environ["NH_TEST_LANE_TAXA"] = str(lane_taxa)
        os.environ["NH_FILTER_MODE"] = str(filter_mode)
        odb.confidence = conf
        want = odb.classify_batch(bases, offsets, paired=paired)
        if os.environ.get("NH_FUZZ_VERBOSE"):
            print(f"it={it} seed={seed} tile_pos={tile_pos} lane_taxa={lane_taxa} filter_mode={filter_mode} entry={entry} paired={paired} "
                  f"conf={conf} n_seqs={len(seqs)} bases={len(bases)}", flush=True)
        with Database.open(d, 0) as db, Session(db, confidence=conf, paired=paired, keep_human=keep_human,
                                                max_batch_bases=len(bases) + 4096, max_batch_seqs=len(seqs) + 2) as sess:
            if entry == "pack":
                call, keep, st = sess.classify_pack(bases, offsets, threads=int(rng.integers(1, 7)))
            else:
                call, keep, st = sess.classify(bases, offsets)
            icall, tk, hg = sess.debug_last_batch(len(call))
        cls = (want["ext"] != 0).astype(np.uint8)
        ok = (np.array_equal(call, want["ext"]) and np.array_equal(tk, want["total_kmers"]) and np.array_equal(hg, want["hit_groups"])
              and np.array_equal(keep, cls if keep_human else 1 - cls))
        if not ok:
            bad = np.nonzero((call != want["ext"]) | (tk != want["total_kmers"]) | (hg != want["hit_groups"]))[0]
            print(f"MISMATCH seed={seed} it={it} tile_pos={tile_pos} lane_taxa={lane_taxa} filter_mode={filter_mode} entry={entry} "
                  f"paired={paired} conf={conf} units={bad[:10].tolist()}")
            for u in bad[:3]:
                nm = 2 if paired else 1
                print("  unit", int(u), "lens", [int(offsets[u * nm + j + 1] - offsets[u * nm + j]) for j in range(nm)],
                      "gpu", int(call[u]), int(tk[u]), int(hg[u]), "oracle", int(want["ext"][u]), int(want["total_kmers"][u]), int(want["hit_groups"][u]))
            return 1
        n_units += len(call)
        it += 1
    print(f"fuzz ok: {it} batches, {n_units} units, no mismatch (seed {args.seed})")
    return 0


if __name__ == "__main__":
    sys.exit(main())
