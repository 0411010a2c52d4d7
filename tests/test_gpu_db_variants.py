"""Databases that differ from kraken2-build's defaults — read unchanged from their
three files: other k / l / spaced seeds (window widths the fused kernel does not cover go
to the warp-per-tile kernels), k == l, minimum_acceptable_hash_value, revcom_version 0,
another toggle mask, a taxonomy too large for shared memory.  Per-read parity with the oracle."""
import numpy as np
import pytest

import synth

pytestmark = pytest.mark.gpu


def build(oracle, tmp_path, name, taxonomy=None, genomes=None, **opt_kw):
    rev = opt_kw.pop("revcom_version", 1)
    opts = oracle.default_options(**opt_kw)
    opts.revcom_version = rev
    genomes = genomes or synth.cfg1_genomes(seed=3, scale=0.004)
    tax = taxonomy or [oracle.TaxSpec(*t) for t in synth.TAXONOMY_CFG1]
    db = oracle.OracleDb.build([(t, bytes(g)) for t, g in genomes], tax, opts=opts)
    d = str(tmp_path / name)
    db.save(d)
    db.genomes = genomes
    return db, d


def check(oracle_db, path, fused_expected, paired=True, conf=0.2):
    from nohuman_b200 import Database, Session
    seqs = synth.illumina_reads(oracle_db.genomes, 800, 150, seed=8, paired=paired, n_rate=0.1)
    seqs += synth.ont_reads(oracle_db.genomes, 8, seed=9, n50=1500, max_len=4000)
    if paired and len(seqs) % 2:
        seqs.append(seqs[0][:90])
    bases, offsets = synth.pack(seqs)
    oracle_db.confidence = conf
    want = oracle_db.classify_batch(bases, offsets, paired=paired)
    with Database.open(path, 0) as db, Session(db, confidence=conf, paired=paired) as sess:
        call, keep, st = sess.classify(bases, offsets)
        icall, tk, hg = sess.debug_last_batch(len(call))
        mins, amb, pos_off = sess.debug_minimizers(bases[:int(offsets[40])], offsets[:41])
    import os
    assert bool(st.fused_kernel) == (fused_expected and os.environ.get("NH_LEGACY_KERNELS") != "1")
    np.testing.assert_array_equal(tk, want["total_kmers"])
    np.testing.assert_array_equal(hg, want["hit_groups"])
    np.testing.assert_array_equal(icall, want["call"])
    np.testing.assert_array_equal(call, want["ext"])
    assert (want["call"] != 0).mean() > 0.3
    return mins, amb, pos_off, seqs


@pytest.mark.parametrize("name,kw,fused", [
    ("k31_l31", dict(k=31, l=31, spaces=0), False),          # k == l: no window
    ("k35_l25_s4", dict(k=35, l=25, spaces=4), False),       # window of 11 l-mers
    ("k25_l21_s3", dict(k=25, l=21, spaces=3), True),        # window of 5 at another k
    ("no_spaced_seed", dict(spaces=0), True),
    ("other_toggle", dict(toggle=0x0123456789ABCDEF), True),
    ("min_hash", dict(min_hash=1 << 62), True),               # a quarter of the minimizers are never looked up
    ("revcom_v0", dict(revcom_version=0), True),              # pre-2.0.8 reverse complement
])
def test_option_variants(oracle, tmp_path, name, kw, fused):
    db, path = build(oracle, tmp_path, name, **kw)
    mins, amb, pos_off, seqs = check(db, path, fused)
    # per-position minimizer stream as well
    for i in range(40):
        wm, wa = oracle.scan_positions(db.opts, bytes(seqs[i]))
        lo, hi = int(pos_off[i]), int(pos_off[i + 1])
        np.testing.assert_array_equal(amb[lo:hi], wa)
        np.testing.assert_array_equal(mins[lo:hi][wa == 0], wm[wa == 0])


def test_taxonomy_larger_than_shared_memory(oracle, tmp_path):
    """9000-node taxonomy (value_bits 14): the parent array is read through L2 instead of shared memory"""
    rng = np.random.default_rng(5)
    tax = [oracle.TaxSpec(1, 1, "root", "no rank")]
    ids = [1]
    for i in range(2, 9001):
        parent = ids[int(rng.integers(max(0, len(ids) - 40), len(ids)))]  # deep, bushy tree
        tax.append(oracle.TaxSpec(i, parent, f"n{i}", "no rank"))
        ids.append(i)
    leaves = [int(x) for x in rng.choice(np.arange(4000, 9001), size=24, replace=False)]
    genomes = [(t, synth.random_genome(rng, 6000)) for t in leaves]
    shared = synth.random_genome(rng, 1500)
    for _, g in genomes[::2]:
        g[2000:3500] = shared  # LCA values high up the tree
    db, path = build(oracle, tmp_path, "bigtax", taxonomy=tax, genomes=genomes)
    assert db.tax.node_count > 8192 and int(db.cht.value_bits) == 14
    check(db, path, True, paired=False, conf=0.05)
    check(db, path, True, paired=True, conf=0.5)


def test_read_hitting_more_taxa_than_any_shared_table(oracle, tmp_path):
    """17 000 leaf taxa with one 60 bp genome each and ONE read that runs through all of them: more
    distinct taxa than the in-warp tables (8), k_score's (64) and k_score_big's shared table (16 384)
    hold.  The unit is classified from the device's global-memory table instead of failing the run."""
    from nohuman_b200 import Database, Session
    rng = np.random.default_rng(41)
    n_leaves = 17_000
    tax = [oracle.TaxSpec(1, 1, "root", "no rank"), oracle.TaxSpec(2, 1, "clade", "no rank")]
    tax += [oracle.TaxSpec(100 + i, 2, f"leaf{i}", "species") for i in range(n_leaves)]
    genomes = [(100 + i, synth.random_genome(rng, 60)) for i in range(n_leaves)]
    db = oracle.OracleDb.build([(t, bytes(g)) for t, g in genomes], tax, capacity=2_000_003)
    d = str(tmp_path / "many_taxa")
    db.save(d)
    everything = np.concatenate([g for _, g in genomes])
    seqs = [everything, everything[:300_000].copy(), genomes[5][1], np.concatenate([g for _, g in genomes[:40]])]
    seqs += synth.illumina_reads(genomes[:2000], 300, 60, seed=3)
    bases, offsets = synth.pack(seqs)
    for conf in (0.0, 0.5):
        db.confidence = conf
        want = db.classify_batch(bases, offsets)
        with Database.open(d, 0) as gdb, Session(gdb, confidence=conf, max_batch_bases=len(bases) + 4096) as sess:
            call, keep, st = sess.classify(bases, offsets)
            icall, tk, hg = sess.debug_last_batch(len(call))
        np.testing.assert_array_equal(hg, want["hit_groups"])
        np.testing.assert_array_equal(tk, want["total_kmers"])
        np.testing.assert_array_equal(call, want["ext"])
    assert want["hit_groups"][0] > 16_384


@pytest.mark.parametrize("name,load,cap_adjust", [
    ("load_0.95", 0.95, 0),        # long probe chains: filter records whose block has no free cell left
    ("load_0.99_odd", 0.99, 13),   # capacity not a multiple of 32 (and of 8): partial last block and sector
    ("load_0.5_prime", 0.5, -1),   # cap_adjust -1: the next prime, nothing divides it
])
@pytest.mark.parametrize("filter_mode", ["1", "2", "3", "0"])
def test_miss_filter_on_awkward_tables(oracle, tmp_path, name, load, cap_adjust, filter_mode, monkeypatch):
    """The miss filter (one record per block of 32 cells: occupancy + Bloom bits) must never change a call:
    tables so full that chains run through several blocks, capacities with a partial last block, every
    policy (0 never asked, 1 units without a hit so far, 2 every lookup, 3 units whose last lookups missed)."""
    monkeypatch.setenv("NH_FILTER_MODE", filter_mode)
    genomes = synth.cfg1_genomes(seed=5, scale=0.004)
    tax = [oracle.TaxSpec(*t) for t in synth.TAXONOMY_CFG1]
    probe = oracle.OracleDb.build([(t, bytes(g)) for t, g in genomes], tax, load_factor=load)
    cap = int(probe.cht.capacity)
    if cap_adjust == -1:
        cap |= 1
        while any(cap % p == 0 for p in range(3, 2000, 2)):
            cap += 2
    else:
        cap += cap_adjust
    db = oracle.OracleDb.build([(t, bytes(g)) for t, g in genomes], tax, capacity=cap)
    d = str(tmp_path / name)
    db.save(d)
    db.genomes = genomes
    from nohuman_b200 import Database
    with Database.open(d, 0) as gdb:
        assert (gdb.info.filter_bytes > 0) == True
        assert gdb.info.filter_bytes == (cap + 31) // 32 * 32
    check(db, d, True, paired=True, conf=0.1)
    check(db, d, True, paired=False, conf=0.0)
