"""The miss filter's algorithm (kernels k_filter_build and the FILTER branch of k_stream_classify in
nohuman_b200/csrc/nh_kernels.cu), restated in numpy and checked on the CPU against the oracle's table:

* a key that IS in the table is never called a miss, and the cell the filter hands it to the table at is at or
  before the cell the key sits in (so the table probe that follows finds it);
* a key the filter calls a miss is a miss for CompactHashTable::Get;
* at load 0.7 the filter answers most missing keys with one record.

The GPU code is checked against the oracle end to end in tests/test_gpu_*.py; this pins the reasoning."""
import numpy as np
import pytest

import synth

M32 = 0xFFFFFFFF


def filter_hash(ck: np.ndarray) -> np.ndarray:
    h = (ck.astype(np.uint64) * 0x9E3779B1) & M32
    h ^= h >> 15
    h = (h * 0x85EBCA77) & M32
    h ^= h >> 13
    return h.astype(np.uint64)


def bloom_bits(h):
    i0, i1, i2, i3 = h & 63, (h >> 6) & 63, (h >> 12) & 63, (h >> 18) & 31
    return [(1 + (i0 >> 5), i0 & 31), (3 + (i1 >> 5), i1 & 31), (5 + (i2 >> 5), i2 & 31), (np.full_like(i3, 7), i3)]


def build_filter(cells: np.ndarray, capacity: int, value_bits: int) -> np.ndarray:
    n_blocks = (capacity + 31) // 32
    rec = np.zeros((n_blocks, 8), np.uint64)
    idx = np.arange(n_blocks * 32, dtype=np.int64)
    exists = idx < capacity
    c = np.zeros(n_blocks * 32, np.uint64)
    c[:capacity] = cells
    occupied = (~exists) | ((c & ((1 << value_bits) - 1)) != 0)
    np.bitwise_or.at(rec[:, 0], idx[occupied] // 32, np.uint64(1) << (idx[occupied] % 32).astype(np.uint64))
    stored = exists & occupied
    h = filter_hash(c[stored] >> np.uint64(value_bits))
    blk = idx[stored] // 32
    for w, b in bloom_bits(h):
        np.bitwise_or.at(rec, (blk, w.astype(np.int64)), np.uint64(1) << b)
    return rec


def ask(rec, n_blocks, home, ck):
    """-> ("miss", None, records asked) or ("table", cell to start at, records asked)"""
    h = int(filter_hash(np.array([ck], np.uint64))[0])
    i0, i1, i2, i3 = h & 63, (h >> 6) & 63, (h >> 12) & 63, (h >> 18) & 31
    probes = [(1 + (i0 >> 5), i0 & 31), (3 + (i1 >> 5), i1 & 31), (5 + (i2 >> 5), i2 & 31), (7, i3)]
    blk, off, asked = home // 32, home % 32, 0
    while True:
        asked += 1
        r = [int(x) for x in rec[blk]]
        maybe = all((r[w] >> b) & 1 for w, b in probes)
        free = (~r[0]) & M32 & ((M32 << off) & M32)
        if maybe:
            return "table", blk * 32 + off, asked
        if free:
            return "miss", None, asked
        blk, off = (blk + 1) % n_blocks, 0
        if asked > n_blocks:
            return "miss", None, asked


@pytest.mark.parametrize("load,cap_adjust", [(0.7, 0), (0.95, 13)])
def test_filter_never_loses_a_key(oracle, load, cap_adjust):
    genomes = synth.cfg1_genomes(seed=4, scale=0.002)
    tax = [oracle.TaxSpec(*t) for t in synth.TAXONOMY_CFG1]
    probe = oracle.OracleDb.build([(t, bytes(g)) for t, g in genomes], tax, load_factor=load)
    cap = int(probe.cht.capacity) + cap_adjust
    db = oracle.OracleDb.build([(t, bytes(g)) for t, g in genomes], tax, capacity=cap)
    vb = int(db.cht.value_bits)
    cells = db.cells().astype(np.uint64)
    rec = build_filter(cells, cap, vb)
    n_blocks = len(rec)
    # keys in the table: the minimizers of the genomes
    present = []
    for _, g in genomes:
        mins, amb = oracle.scan_positions(db.opts, bytes(g[:6000]))
        present += np.unique(mins[amb == 0]).tolist()
    present = present[:4000]
    for key in present:
        hk = oracle.fmix64(int(key))
        home, ck = hk % cap, hk >> (32 + vb)
        kind, cell, _ = ask(rec, n_blocks, home, ck)
        assert kind == "table", "a stored key was called a miss"
        assert db.get(int(key)) != 0
        # the key's own cell: first cell at or after home (cyclically) holding ck; the hand-over cell must not be past it
        pos = home
        while int(cells[pos]) >> vb != ck or int(cells[pos]) & ((1 << vb) - 1) == 0:
            pos = (pos + 1) % cap
        dist_key = (pos - home) % cap
        dist_start = (cell - home) % cap if cell < cap else (cell - home)
        assert dist_start <= dist_key
    # random keys: a "miss" from the filter is a miss of the table; most need one record
    rng = np.random.default_rng(5)
    one, total_miss = 0, 0
    for key in rng.integers(0, 1 << 62, 4000, dtype=np.uint64).tolist():
        hk = oracle.fmix64(int(key))
        home, ck = hk % cap, hk >> (32 + vb)
        kind, cell, asked = ask(rec, n_blocks, home, ck)
        if kind == "miss":
            assert db.get(int(key)) == 0
            total_miss += 1
            one += asked == 1
    assert total_miss > (3500 if load <= 0.7 else 2500)  # a fuller table: more Bloom false positives, longer chains
    if load <= 0.7:
        assert one / total_miss > 0.75
