"""Host side of the synthetic-database tooling (nohuman_b200/synth.py) on CPU: the
opts.k2d / taxo.k2d images it writes are what the oracle (and therefore any kraken2
reader of Appendix B's formats) reads back."""
import ctypes as C
import os

import numpy as np

from nohuman_b200 import synth


def test_opts_image_matches_index_options(oracle, tmp_path):
    img = synth.opts_image()
    assert len(img) == 64
    p = tmp_path / "opts.k2d"
    p.write_bytes(img)
    o = oracle.IndexOptions()
    assert oracle.lib().k2o_load_opts(str(p).encode(), C.byref(o)) == 0
    d = oracle.default_options()
    assert (o.k, o.l, o.spaced_seed_mask, o.toggle_mask, o.dna_db, o.revcom_version) == \
           (d.k, d.l, d.spaced_seed_mask, d.toggle_mask, 1, 1)
    assert synth.spaced_seed_mask(31, 7) == 0x3FFFFFFFF3333333 == oracle.lib().k2o_spaced_seed_mask(31, 7)
    img2 = synth.opts_image(k=31, l=25, spaces=3, min_hash=12345)
    p.write_bytes(img2)
    assert oracle.lib().k2o_load_opts(str(p).encode(), C.byref(o)) == 0
    assert (o.k, o.l, o.minimum_acceptable_hash_value) == (31, 25, 12345)
    assert o.spaced_seed_mask == oracle.lib().k2o_spaced_seed_mask(25, 3)


def test_taxonomy_image_matches_oracle_builder(oracle, tmp_path):
    nodes, leaves = synth.human_pangenome_taxonomy(n_super=3, n_hap_per_super=5)
    img, internal = synth.taxonomy_image(nodes)
    p = tmp_path / "taxo.k2d"
    p.write_bytes(img)
    tax = oracle.Taxonomy()
    assert oracle.lib().k2o_load_taxonomy(str(p).encode(), C.byref(tax)) == 0
    assert tax.node_count == len(nodes) + 1
    # same numbering as the oracle's own BFS builder: parent < child, root = 1
    specs = [oracle.TaxSpec(n.ext_id, n.parent_ext_id, n.name, n.rank) for n in nodes]
    ref = oracle.Taxonomy()
    n = len(specs)
    ext = (C.c_uint64 * n)(*[t.ext_id for t in specs])
    par = (C.c_uint64 * n)(*[t.parent_ext_id for t in specs])
    names = (C.c_char_p * n)(*[t.name.encode() for t in specs])
    ranks = (C.c_char_p * n)(*[t.rank.encode() for t in specs])
    assert oracle.lib().k2o_taxonomy_build(C.byref(ref), n, ext, par, names, ranks) == 0
    for i in range(1, tax.node_count):
        a, b = tax.nodes[i], ref.nodes[i]
        assert (a.parent_id, a.external_id, a.child_count, a.first_child) == \
               (b.parent_id, b.external_id, b.child_count, b.first_child), i
        assert a.parent_id < i
        assert C.string_at(tax.name_data + a.name_offset) == C.string_at(ref.name_data + b.name_offset)
        assert C.string_at(tax.rank_data + a.rank_offset) == C.string_at(ref.rank_data + b.rank_offset)
    for e, i in internal.items():
        assert oracle.lib().k2o_taxonomy_internal_id(C.byref(tax), e) == i
    assert len(leaves) == 15 and all(l in internal for l in leaves)
    # the human lineage is there, so a report line for 9606 reads "S ... Homo sapiens"
    assert internal[9606] and tax.nodes[internal[9606]].parent_id == internal[9605]


def test_cpu_twin_builds_a_valid_table(oracle):
    """oracle/k2_synth.c (what `bench.py --impl reference` builds its table with): every minimizer the
    oracle's own scanner finds in the synthetic genome is retrievable and carries the block's leaf or an
    ancestor of it; the load factor lands on the target; a second build gives the same lookups."""
    from oracle import k2synth
    cap = (1 << 18) + 7
    db, meta = k2synth.build_synthetic_db(cap, block_bases=1 << 13)
    hdr = meta["hash_header"]
    assert hdr[0] == cap and 0.69 < hdr[1] / cap < 0.72
    cells = db.cells()
    vmask = (1 << hdr[3]) - 1
    assert int(((cells & vmask) != 0).sum()) == hdr[1]
    g = k2synth.synth_genome(meta["genome_seed"], 0, 120_000)
    assert set(np.unique(g).tolist()) <= set(b"ACGT")
    mins, amb = oracle.scan_positions(db.opts, bytes(g))
    assert not amb.any()
    nodes, leaves = synth.human_pangenome_taxonomy()
    leaf_int = [meta["internal"][x] for x in leaves]
    block, tile = 1 << 13, 124
    n_lca = 0
    for p in range(0, len(mins), 29):
        v = db.get(int(mins[p]))
        assert v != 0, p
        leaf = leaf_int[((p // tile * tile) // block) % len(leaf_int)]
        assert oracle.lib().k2o_is_a_ancestor_of_b(C.byref(db.tax), v, leaf), (p, v, leaf)
        n_lca += v != leaf
    assert n_lca > 0
    db2, meta2 = k2synth.build_synthetic_db(cap, block_bases=1 << 13, threads=1)
    assert meta2["genome_bases"] == meta["genome_bases"] and meta2["hash_header"] == hdr
    assert all(db2.get(int(m)) == db.get(int(m)) for m in mins[::101])


def test_cpu_twin_reads_are_deterministic_and_half_human(oracle):
    from oracle import k2synth
    cap = (1 << 18) + 7
    db, meta = k2synth.build_synthetic_db(cap, block_bases=1 << 13)
    off = (np.arange(4001) * 150).astype(np.uint64)
    a = k2synth.synth_reads(off, meta["genome_seed"], 2 * cap, seed=3, paired=True, threads=1)
    b = k2synth.synth_reads(off, meta["genome_seed"], 2 * cap, seed=3, paired=True, threads=4)
    assert np.array_equal(a, b) and set(np.unique(a).tolist()) <= set(b"ACGTN")
    db.confidence = 0.5
    r = db.classify_batch(a, off, paired=True)
    assert 0.4 < (r["ext"] != 0).mean() < 0.6
