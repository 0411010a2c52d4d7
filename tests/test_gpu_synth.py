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


@pytest.fixture(scope="module")
def big_sdb():
    """2^26 + 5 cells (256 MiB, far beyond L2-resident probing of a toy table), built on the GPU"""
    from nohuman_b200 import synth
    s = synth.build_synthetic_db(capacity=(1 << 26) + 5, device=0)
    yield s
    s.db.close()


@pytest.fixture(scope="module")
def big_odb(big_sdb, oracle):
    import tempfile
    d = tempfile.mkdtemp()
    big_sdb.save(d)
    return oracle.OracleDb.load(d)


def _ont_lengths(rng, n, n50=10_000, lo=200, hi=100_000):
    sigma = 0.9
    mu = np.log(n50) - sigma * sigma  # length-weighted median of a log-normal = exp(mu + sigma^2)
    L = np.exp(rng.normal(mu, sigma, size=n)).astype(np.int64)
    L = np.clip(L, lo, hi)
    L[:6] = [hi, hi - 1, 50_000, 75_001, 16_290, 16_291]  # the longest reads and a few exact tile multiples
    return L


def test_ont_reads_up_to_100kb_match_oracle(big_sdb, big_odb):
    """BASELINE configs[2] shape at the scale it is quoted: log-normal lengths with N50 ~10 kb up to
    100 kb (hundreds of tiles, many 32-tile groups, k_score merging tile tables), 5 % errors,
    keep-human, against a GPU-built 2^26-cell table."""
    import torch
    from nohuman_b200 import Session, synth
    rng = np.random.default_rng(26)
    n = 3500
    lens = _ont_lengths(rng, n)
    offsets = np.zeros(n + 1, np.int64)
    np.cumsum(lens, out=offsets[1:])
    total = int(offsets[-1])
    assert total > 20_000_000 and lens.max() == 100_000
    d_off = torch.from_numpy(offsets).cuda()
    d_bases = torch.zeros(total + 64, dtype=torch.uint8, device="cuda")
    e = 0.05 / 3
    synth.synth_reads(0, d_bases.data_ptr(), d_off.data_ptr(), n, big_sdb.genome_seed, big_sdb.genome_bases, seed=4,
                      human_frac=0.5, sub_rate=e, ins_rate=e, del_rate=e, n_rate=0.02)
    torch.cuda.synchronize()
    bases = d_bases[:total].cpu().numpy()
    for conf in (0.0, 0.2):
        with Session(big_sdb.db, confidence=conf, keep_human=True, max_batch_bases=total + 1024, max_batch_seqs=n) as sess:
            call, keep, st = sess.classify(bases, offsets.astype(np.uint64))
            icall, tk, hg = sess.debug_last_batch(n)
        big_odb.confidence = conf
        want = big_odb.classify_batch(bases, offsets.astype(np.uint64), paired=False)
        np.testing.assert_array_equal(tk, want["total_kmers"])
        np.testing.assert_array_equal(hg, want["hit_groups"])
        np.testing.assert_array_equal(call, want["ext"])
        np.testing.assert_array_equal(keep, (want["ext"] != 0).astype(np.uint8))
        frac = float((call != 0).mean())
        # half the reads come from the genome; with 5 % errors only 0.95^35 = 17 % of their k-mers survive,
        # so --conf 0.2 calls few of them while --conf 0 calls nearly all
        assert (0.4 < frac < 0.6) if conf == 0.0 else (0.02 < frac < 0.4), (conf, frac)


def test_replica_adopted_from_device_memory(big_sdb, big_odb):
    """Database.from_memory(cells_on_device=True): the path ranks 1..N-1 take after the NCCL broadcast —
    the cell array is adopted in place (no copy).  The replica must answer exactly like the original."""
    import torch
    from nohuman_b200 import Database, Session, synth
    from nohuman_b200 import dist as nhd
    cap = int(big_sdb.db.info.capacity)
    nbytes = nhd.padded_cells(cap) * 4
    src = torch.as_tensor(_DevPtr(big_sdb.db.device_cells_ptr(), nbytes), device="cuda")
    replica_cells = src.clone()  # what dist.broadcast_table leaves on a non-source rank
    assert replica_cells.data_ptr() % 128 == 0
    rep = Database.from_memory(big_sdb.opts, big_sdb.taxo, big_sdb.hash_header(), replica_cells.data_ptr(), device=0,
                               cells_on_device=True)
    try:
        assert rep.device_cells_ptr() == replica_cells.data_ptr()
        n_pairs = 20000
        n_seqs, L = 2 * n_pairs, 150
        d_off = torch.arange(n_seqs + 1, dtype=torch.int64, device="cuda") * L
        d_bases = torch.zeros(n_seqs * L + 64, dtype=torch.uint8, device="cuda")
        synth.synth_reads(0, d_bases.data_ptr(), d_off.data_ptr(), n_seqs, big_sdb.genome_seed, big_sdb.genome_bases,
                          seed=77, paired=True, n_rate=0.01)
        torch.cuda.synchronize()
        bases = d_bases[:n_seqs * L].cpu().numpy()
        offsets = d_off.cpu().numpy().astype(np.uint64)
        with Session(rep, confidence=0.5, paired=True) as s1, Session(big_sdb.db, confidence=0.5, paired=True) as s0:
            c1, k1, _ = s1.classify(bases, offsets)
            c0, k0, _ = s0.classify(bases, offsets)
        big_odb.confidence = 0.5
        want = big_odb.classify_batch(bases, offsets, paired=True)
        np.testing.assert_array_equal(c1, want["ext"])
        np.testing.assert_array_equal(c0, c1)
        np.testing.assert_array_equal(k0, k1)
    finally:
        rep.close()
    # closing the replica must not free memory it does not own
    assert int(replica_cells[:4096].sum().item()) == int(src[:4096].sum().item())


class _DevPtr:
    def __init__(self, ptr, nbytes):
        self.__cuda_array_interface__ = {"shape": (nbytes,), "typestr": "|u1", "data": (ptr, False), "version": 3,
                                         "strides": None}


def test_cpu_twin_of_the_generator_matches_the_gpu(sdb, odb):
    """bench.py's reference arm builds its table and samples its reads with oracle/k2_synth.c instead of the
    CUDA library: same genome, byte-identical reads, and a table that answers like the GPU-built one."""
    import torch
    from nohuman_b200 import synth
    from oracle import k2synth
    cap = int(sdb.db.info.capacity)
    cdb, meta = k2synth.build_synthetic_db(cap, block_bases=1 << 14)
    assert abs(meta["genome_bases"] - sdb.genome_bases) < 4096
    assert abs(meta["hash_header"][1] - int(sdb.db.info.size)) < 64  # layout differs, occupancy does not
    g_gpu = synth.synth_genome(0, sdb.genome_seed, 12345, 100_000)
    g_cpu = k2synth.synth_genome(sdb.genome_seed, 12345, 100_000)
    assert np.array_equal(g_gpu, g_cpu)
    for paired, ins in ((True, 0.0), (False, 0.02)):
        n = 6000
        rng = np.random.default_rng(3)
        lens = np.full(n, 150, np.int64) if paired else rng.integers(20, 3000, size=n)
        off = np.zeros(n + 1, np.int64)
        np.cumsum(lens, out=off[1:])
        total = int(off[-1])
        d_off = torch.from_numpy(off).cuda()
        d_bases = torch.zeros(total + 64, dtype=torch.uint8, device="cuda")
        kw = dict(seed=11, human_frac=0.5, sub_rate=0.01, ins_rate=ins, del_rate=ins, n_rate=0.05, paired=paired,
                  insert_mean=350.0, insert_sd=50.0)
        synth.synth_reads(0, d_bases.data_ptr(), d_off.data_ptr(), n, sdb.genome_seed, 2 * cap, **kw)
        torch.cuda.synchronize()
        gpu = d_bases[:total].cpu().numpy()
        cpu = k2synth.synth_reads(off.astype(np.uint64), sdb.genome_seed, 2 * cap, **kw)
        assert np.array_equal(gpu, cpu), (paired, int((gpu != cpu).sum()))
        odb.confidence = cdb.confidence = 0.1
        a = odb.classify_batch(gpu, off.astype(np.uint64), paired=paired)
        b = cdb.classify_batch(cpu, off.astype(np.uint64), paired=paired)
        assert (a["ext"] != b["ext"]).mean() < 1e-3  # only compacted-key collisions of absent keys can differ


def test_probe_pattern_microbenchmark_counts(big_sdb):
    """The roofline helper (nh_bench_probe_pattern): requests = items x (1 + spill rate) for every fetch
    flavour, and more requests in flight are never slower by an order of magnitude (sanity, not a benchmark)."""
    for lanes, depth in ((1, 1), (1, 4), (2, 2), (4, 1), (0, 1), (0, 2)):
        items, req = big_sdb.db.probe_pattern(lanes=lanes, p_continue=0.4, items_per_chain=64, iters=1, depth=depth,
                                              blocks_per_sm=4)
        assert items > 1e9 and 1.3 < req / items < 1.5, (lanes, depth, items, req)
    items0, req0 = big_sdb.db.probe_pattern(lanes=1, p_continue=0.0, items_per_chain=64, iters=1)
    assert abs(req0 / items0 - 1.0) < 1e-6
