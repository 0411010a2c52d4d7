"""GPU synthetic-workload tooling (nohuman_b200/synth.py) checked with the oracle:
the table the GPU builder writes is a valid kraken2 CompactHashTable, and the
CUDA classification of GPU-sampled reads against it matches the oracle."""
import ctypes as C

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def sdb():
    from nohuman_b200 import synth
    s = synth.build_synthetic_db(capacity=(1 << 20) + 13, device=0, block_bases=1 << 14)
    yield s
    s.db.close()


@pytest.fixture(scope="module")
def odb(sdb, oracle):
    opts = oracle.IndexOptions.from_buffer_copy(sdb.opts)
    import os, tempfile
    d = tempfile.mkdtemp()
    sdb.save(d)
    db = oracle.OracleDb.load(d)
    return db


def test_table_is_valid(sdb, odb, oracle):
    from nohuman_b200 import synth
    info = sdb.db.info
    cells = odb.cells()
    vmask = (1 << int(info.value_bits)) - 1
    assert int(((cells & vmask) != 0).sum()) == int(info.size)
    load = info.size / info.capacity
    assert 0.69 < load < 0.72, load
    assert int((cells & vmask).max()) < info.node_count
    # every minimizer of the genome is retrievable and maps to the block's leaf or an ancestor
    g = synth.synth_genome(0, sdb.genome_seed, 0, 200_000)
    assert set(np.unique(g).tolist()) <= set(b"ACGT")
    mins, amb = oracle.scan_positions(odb.opts, bytes(g))
    assert not amb.any()
    nodes, leaves = synth.human_pangenome_taxonomy()
    leaf_int = [sdb.internal[x] for x in leaves]
    block = 1 << 14
    n_lca = 0
    for p in range(0, len(mins), 37):
        v = odb.get(int(mins[p]))
        assert v != 0, p
        leaf = leaf_int[(p // block) % len(leaf_int)]
        assert oracle.lib().k2o_is_a_ancestor_of_b(C.byref(odb.tax), v, leaf), (p, v, leaf)
        n_lca += v != leaf
    assert n_lca > 0  # overlaps / repeats produced LCA values


@pytest.mark.parametrize("paired,conf,ins", [(False, 0.0, 0.0), (True, 0.5, 0.0), (False, 0.1, 0.02)])
def test_synth_reads_classify_like_oracle(sdb, odb, paired, conf, ins):
    import torch
    from nohuman_b200 import Session, synth
    n_units = 5000
    n_seqs = n_units * (2 if paired else 1)
    rng = np.random.default_rng(1)
    lens = np.full(n_seqs, 150, np.int64) if paired else rng.integers(20, 600, size=n_seqs)
    offsets = np.zeros(n_seqs + 1, np.int64)
    np.cumsum(lens, out=offsets[1:])
    total = int(offsets[-1])
    d_off = torch.from_numpy(offsets).cuda()
    d_bases = torch.zeros(total + 64, dtype=torch.uint8, device="cuda")
    synth.synth_reads(0, d_bases.data_ptr(), d_off.data_ptr(), n_seqs, sdb.genome_seed,
                      sdb.genome_bases, seed=9, paired=paired, ins_rate=ins, del_rate=ins,
                      n_rate=0.05)
    torch.cuda.synchronize()
    bases = d_bases[:total].cpu().numpy()
    assert set(np.unique(bases).tolist()) <= set(b"ACGTN")
    d_call = torch.zeros(n_units, dtype=torch.int32, device="cuda")
    d_keep = torch.zeros(n_units, dtype=torch.uint8, device="cuda")
    with Session(sdb.db, confidence=conf, paired=paired, max_batch_bases=total + 1024,
                 max_batch_seqs=n_seqs) as sess:
        sess.classify_device(d_bases.data_ptr(), d_off.data_ptr(), n_seqs, total,
                             d_call.data_ptr(), d_keep.data_ptr())
        st = sess.sync()
        # host-buffer path gives the same answer
        call2, keep2, st2 = sess.classify(bases, offsets.astype(np.uint64))
    odb.confidence = conf
    want = odb.classify_batch(bases, offsets.astype(np.uint64), paired=paired)
    got = d_call.cpu().numpy().astype(np.uint32)
    np.testing.assert_array_equal(got, want["ext"])
    np.testing.assert_array_equal(call2, want["ext"])
    np.testing.assert_array_equal(d_keep.cpu().numpy(), (want["ext"] == 0).astype(np.uint8))
    frac = (want["ext"] != 0).mean()
    assert 0.3 < frac < 0.6, frac  # ~half the units are genome-derived
    assert st.n_classified == int((want["ext"] != 0).sum()) == st2.n_classified
