"""CPU checks of the oracle (oracle/k2_oracle.c) against known-answer vectors
(SURVEY.md Appendix C), against an independent brute-force restatement of the
same definitions written here in plain Python, and against hand-derived
ResolveTree / ClassifySequence cases.  The reference repo holds no golden
vector for this path (src/lib.rs:153-222 only runs `ls`), so these
are the vectors the oracle is pinned to; see k2_oracle.h "PARITY UNPINNED".
"""
import ctypes as C
import math

import numpy as np
import pytest

import synth

MASK64 = (1 << 64) - 1
CODE = {ord("A"): 0, ord("C"): 1, ord("G"): 2, ord("T"): 3,
        ord("a"): 0, ord("c"): 1, ord("g"): 2, ord("t"): 3}


# ------------------------------------------------------------------ primitives
def py_fmix64(k):
    k ^= k >> 33
    k = (k * 0xFF51AFD7ED558CCD) & MASK64
    k ^= k >> 33
    k = (k * 0xC4CEB9FE1A85EC53) & MASK64
    k ^= k >> 33
    return k


def py_revcomp(x, n):
    """reverse complement of an n-mer in 2-bit code, from the definition"""
    r = 0
    for _ in range(n):
        r = (r << 2) | (3 - (x & 3))
        x >>= 2
    return r


def test_fmix64_known_answers(oracle):
    kat = {0: 0x0, 1: 0xB456BCFC34C2CB2C, 2: 0x3ABF2A20650683E7,
           0xDEADBEEF: 0xD24BD59F862A1DAC, MASK64: 0x64B5720B4B825F21}
    for x, want in kat.items():
        assert oracle.fmix64(x) == want
        assert py_fmix64(x) == want
    rng = np.random.default_rng(0)
    for x in rng.integers(0, 1 << 63, size=2000, dtype=np.uint64):
        assert oracle.fmix64(int(x)) == py_fmix64(int(x))


def test_seed_and_toggle_masks(oracle):
    L = oracle.lib()
    assert L.k2o_spaced_seed_mask(31, 7) == 0x3FFFFFFFF3333333
    assert L.k2o_spaced_seed_mask(31, 0) == (1 << 62) - 1
    assert 0xE37E28C4271B5A2D & ((1 << 62) - 1) == 0x237E28C4271B5A2D
    o = oracle.default_options()
    assert (o.k, o.l, o.dna_db, o.revcom_version) == (35, 31, 1, 1)
    assert C.sizeof(oracle.IndexOptions) == 64
    assert C.sizeof(oracle.TaxonomyNode) == 56


def test_reverse_complement(oracle):
    L = oracle.lib()
    enc = lambda s: int("".join(f"{CODE[ord(c)]:02b}" for c in s), 2)
    assert L.k2o_reverse_complement(enc("ACGT"), 4, 1) == enc("ACGT")
    assert L.k2o_reverse_complement(enc("AAAC"), 4, 1) == enc("GTTT")
    assert L.k2o_reverse_complement(enc("A" * 31), 31, 1) == enc("T" * 31)
    assert L.k2o_canonical(enc("TTTT"), 4, 1) == enc("AAAA")
    rng = np.random.default_rng(1)
    for n in (1, 2, 15, 16, 30, 31):
        for x in rng.integers(0, 1 << (2 * n), size=300, dtype=np.uint64):
            x = int(x)
            assert L.k2o_reverse_complement(x, n, 1) == py_revcomp(x, n)
            # revcom_version 0 (pre-2.0.8 databases): no shift back down
            full = py_revcomp(x, 32)  # x zero-extended to 32 bases
            assert L.k2o_reverse_complement(x, n, 0) == full & ((1 << (2 * n)) - 1)


# ------------------------------------------------------------------ scanner
def py_scan_positions(seq: bytes, k=35, l=31, seed_mask=0x3FFFFFFFF3333333,
                      toggle=0xE37E28C4271B5A2D):
    """Per-position (minimizer, ambiguous) from the pure-function restatement
    of MinimizerScanner, independent of the oracle's deque code."""
    lmask = (1 << (2 * l)) - 1
    toggle &= lmask
    n = len(seq)
    # cand[s] for the l-mer starting at s, None if it holds an ambiguous base
    cand = []
    for s in range(max(0, n - l + 1)):
        x = 0
        ok = True
        for ch in seq[s:s + l]:
            if ch not in CODE:
                ok = False
                break
            x = (x << 2) | CODE[ch]
        if not ok:
            cand.append(None)
            continue
        c = min(x, py_revcomp(x, l))
        if seed_mask:
            c &= seed_mask
        cand.append(c ^ toggle)
    out_m, out_a = [], []
    run = 0  # consecutive non-ambiguous bases ending at e
    runs = []
    for ch in seq:
        run = run + 1 if ch in CODE else 0
        runs.append(run)
    for e in range(k, n + 1):  # e = bases consumed
        c = runs[e - 1]
        if c < k - 1:
            out_m.append(None)
            out_a.append(1)
            continue
        n_lmers = min(c, k) - l + 1  # k-1 -> k-l l-mers (quirk), >=k -> k-l+1
        first = e - l - (n_lmers - 1)
        window = cand[first:e - l + 1]
        assert all(w is not None for w in window)
        out_m.append(min(window) ^ toggle)
        out_a.append(0)
    return out_m, out_a


@pytest.mark.parametrize("seed", range(6))
def test_scanner_matches_bruteforce(oracle, seed):
    rng = np.random.default_rng(seed)
    o = oracle.default_options()
    L = int(rng.integers(30, 260))
    s = synth.random_genome(rng, L)
    if seed % 2:  # sprinkle ambiguous bytes, lower case
        for p in rng.integers(0, L, size=3):
            s[p] = ord("N")
        s[L // 2:] |= 0x20
        s[s == (ord("N") | 0x20)] = ord("n")
    mins, amb = oracle.scan_positions(o, bytes(s))
    wm, wa = py_scan_positions(bytes(s))
    assert len(mins) == max(0, L - 35 + 1) == len(wm)
    assert amb.tolist() == wa
    for got, want, a in zip(mins.tolist(), wm, wa):
        if not a:
            assert got == want


def test_scanner_edge_lengths(oracle):
    o = oracle.default_options()
    rng = np.random.default_rng(9)
    for L in (0, 1, 30, 31, 34):
        m, a = oracle.scan_positions(o, bytes(synth.random_genome(rng, L)))
        assert len(m) == 0
    m, a = oracle.scan_positions(o, bytes(synth.random_genome(rng, 35)))
    assert len(m) == 1 and a[0] == 0
    m, a = oracle.scan_positions(o, b"N" * 100)
    assert len(m) == 66 and a.all()
    # one N in the middle of a 150-mer: positions whose last k-1 bases hold it are ambiguous
    s = synth.random_genome(rng, 150)
    s[75] = ord("N")
    m, a = oracle.scan_positions(o, bytes(s))
    e = np.arange(35, 151)  # bases consumed at each return
    want = ((e - 1 >= 75) & (e - 1 - 75 < 34)).astype(np.uint8)
    assert a.tolist() == want.tolist()
    # k == l short-circuit: every position returns its own canonical l-mer
    o2 = oracle.default_options(k=31, l=31, spaces=0)
    s = synth.random_genome(rng, 60)
    m, a = oracle.scan_positions(o2, bytes(s))
    wm, wa = py_scan_positions(bytes(s), k=31, l=31, seed_mask=0)
    assert m.tolist() == wm and not a.any()


# ------------------------------------------------------------------ hash table
def py_get(cells, cap, vbits, key):
    h = py_fmix64(key)
    ck = h >> (32 + vbits)
    idx = h % cap
    first = idx
    while True:
        cell = int(cells[idx])
        v = cell & ((1 << vbits) - 1)
        if v == 0:
            return 0
        if cell >> vbits == ck:
            return v
        idx = (idx + 1) % cap
        if idx == first:
            return 0


def test_compact_hash_table(oracle, small_db):
    L = oracle.lib()
    cells = small_db.cells()
    cap, vb = int(small_db.cht.capacity), int(small_db.cht.value_bits)
    assert int(small_db.cht.key_bits) + vb == 32
    assert (1 << vb) >= small_db.tax.node_count > (1 << (vb - 1))
    assert int(((cells & ((1 << vb) - 1)) != 0).sum()) == int(small_db.cht.size)
    assert 0.6 < small_db.cht.size / cap <= 0.7 + 1e-9
    rng = np.random.default_rng(2)
    taxid, g = small_db.genomes[0]
    mins, amb = oracle.scan_positions(small_db.opts, bytes(g[:3000]))
    internal = small_db.internal_id(taxid)
    for key in mins[amb == 0][::7].tolist():
        v = small_db.get(key)
        assert v == py_get(cells, cap, vb, key)
        assert v != 0 and L.k2o_is_a_ancestor_of_b(C.byref(small_db.tax), v, internal)
    hits = 0
    for key in rng.integers(0, 1 << 62, size=3000, dtype=np.uint64).tolist():
        v = small_db.get(key)
        assert v == py_get(cells, cap, vb, key)
        hits += v != 0
    # compacted-key collisions give a few false positives by design, not many
    assert hits < 300
    # stats variant agrees and counts sectors of the linear probe
    c, s = C.c_uint64(), C.c_uint64()
    key = int(mins[amb == 0][0])
    assert L.k2o_cht_get_stats(C.byref(small_db.cht), key, C.byref(c), C.byref(s)) == small_db.get(key)
    assert c.value >= 1 and 1 <= s.value <= (c.value + 14) // 8 + 1


def test_shared_segment_gets_lca(oracle, small_db):
    """the two Escherichia genomes share a segment: its minimizers carry the genus"""
    g_ecoli = dict(small_db.genomes)[562]
    lo = len(g_ecoli) // 4
    mins, amb = oracle.scan_positions(small_db.opts, bytes(g_ecoli[lo + 100:lo + 1100]))
    genus = small_db.internal_id(561)
    vals = [small_db.get(int(m)) for m in mins[amb == 0]]
    assert vals.count(genus) > 0.9 * len(vals)
    assert small_db.lca(small_db.internal_id(562), small_db.internal_id(564)) == genus
    assert small_db.lca(small_db.internal_id(562), small_db.internal_id(1423)) == small_db.internal_id(2)
    assert small_db.lca(small_db.internal_id(9606), small_db.internal_id(562)) == small_db.internal_id(131567)
    par = small_db.parents()
    assert par[1] == 0 and all(par[i] < i for i in range(2, len(par)))  # BFS ids: parent < child


# ------------------------------------------------------------------ ResolveTree
def _resolve(oracle, db, counts: dict, total, conf):
    hc = oracle.HitCounts()
    n = len(counts)
    tx = (C.c_uint64 * max(n, 1))(*[db.internal_id(t) for t in counts])
    ct = (C.c_uint32 * max(n, 1))(*counts.values())
    hc.taxon, hc.count, hc.n, hc.cap = tx, ct, n, max(n, 1)
    r = oracle.lib().k2o_resolve_tree(C.byref(hc), C.byref(db.tax), total, conf)
    ext = db.external_ids()
    return int(ext[r]) if r else 0


def test_resolve_tree_hand_cases(oracle, small_db):
    R = lambda counts, total, conf: _resolve(oracle, small_db, counts, total, conf)
    assert R({}, 116, 0.0) == 0
    assert R({9606: 10}, 116, 0.0) == 9606
    # leaf wins over its ancestors: root-to-leaf sum is largest at the leaf
    assert R({9606: 5, 9605: 3, 1: 20}, 116, 0.0) == 9606
    # two sibling species tie -> LCA (genus); a third hit on the genus does not break the tie
    assert R({562: 4, 564: 4}, 116, 0.0) == 561
    assert R({562: 4, 564: 4, 561: 2}, 116, 0.0) == 561
    assert R({562: 5, 564: 4}, 116, 0.0) == 562
    # confidence walk-up: required = ceil(conf * total)
    assert R({562: 5, 564: 4}, 100, 0.05) == 562          # needs 5, has 5
    assert R({562: 5, 564: 4}, 100, 0.06) == 561          # needs 6: genus clade has 9
    assert R({562: 5, 564: 4}, 100, 0.09) == 561
    assert R({562: 5, 564: 4}, 100, 0.10) == 0            # needs 10 > 9 everywhere
    assert R({562: 5, 564: 4, 1423: 1}, 100, 0.10) == 2   # Bacteria clade has 10
    assert R({9606: 58}, 116, 0.5) == 9606                # ceil(58.0) == 58
    assert R({9606: 57}, 116, 0.5) == 0
    # ceil is taken on the IEEE double product, as kraken2 does
    for conf in (0.1, 0.3, 0.7, float(np.float32(0.1))):
        for total in (10, 30, 116, 232, 1000):
            need = math.ceil(conf * total)
            assert R({9606: need}, total, conf) == 9606
            if need > 0:
                assert R({9606: need - 1}, total, conf) == 0


# ------------------------------------------------------------------ ClassifySequence
def test_classify_sequence_hand_cases(oracle, small_db):
    g = dict(small_db.genomes)
    human, ecoli = g[9606], g[562]
    small_db.confidence = 0.0
    r = small_db.classify_one(bytes(human[500:650]), want_taxa=True)
    assert r["ext_call"] == 9606 and r["total_kmers"] == 116 and r["hit_groups"] >= 2
    assert r["hitlist"] == "9606:116"
    assert r["lookups"] == r["hit_groups"]  # every distinct minimizer of a clean genome read hits
    # reverse complement classifies identically (canonical l-mers)
    r2 = small_db.classify_one(bytes(synth.revcomp(human[500:650])))
    assert (r2["ext_call"], r2["total_kmers"], r2["hit_groups"]) == (9606, 116, r["hit_groups"])
    # shorter than k: no k-mers at all
    r = small_db.classify_one(bytes(human[:34]), want_taxa=True)
    assert r["call"] == 0 and r["total_kmers"] == 0 and r["hitlist"] == "0:0"
    # random read: unclassified, all-zero hitlist
    rng = np.random.default_rng(4)
    r = small_db.classify_one(bytes(synth.random_genome(rng, 150)), want_taxa=True)
    assert r["call"] == 0 and r["hitlist"] == "0:116"
    # an N splits the hitlist with an ambiguous span of k-1 positions
    s = human[1000:1150].copy()
    s[75] = ord("N")
    r = small_db.classify_one(bytes(s), want_taxa=True)
    assert r["total_kmers"] == 116 and r["hitlist"] == "9606:41 A:34 9606:41"
    # paired: mates pooled, border marker, one call
    m1, m2 = bytes(human[2000:2150]), bytes(synth.revcomp(human[2200:2350]))
    r = small_db.classify_one(m1, m2, want_taxa=True)
    assert r["ext_call"] == 9606 and r["total_kmers"] == 232
    assert r["hitlist"] == "9606:116 |:| 9606:116"
    # minimum-hit-groups = 2 voids a call resting on a single minimizer group
    s = synth.random_genome(rng, 150)
    s[:36] = human[3000:3036]  # 2 k-mers -> usually 1-2 distinct minimizers
    r = small_db.classify_one(bytes(s))
    if r["hit_groups"] < 2:
        assert r["call"] == 0
    # chimeric human+bacterial read: LCA of the tie / larger side wins; confidence 1.0 kills it
    s = np.concatenate([human[4000:4075], ecoli[3000:3075]])
    r = small_db.classify_one(bytes(s), want_taxa=True)
    assert r["ext_call"] in (9606, 562, 131567)
    small_db.confidence = 1.0
    assert small_db.classify_one(bytes(s))["call"] == 0
    assert small_db.classify_one(bytes(human[500:650]))["ext_call"] == 9606
    small_db.confidence = 0.0


def test_batch_equals_single(oracle, small_db):
    seqs = synth.illumina_reads(small_db.genomes, 300, 150, seed=21, paired=True)
    bases, offsets = synth.pack(seqs)
    small_db.confidence = 0.3
    got = small_db.classify_batch(bases, offsets, paired=True, threads=2)
    lookups = 0
    for u in range(0, 300, 17):
        r = small_db.classify_one(bytes(seqs[2 * u]), bytes(seqs[2 * u + 1]))
        assert got["call"][u] == r["call"] and got["ext"][u] == r["ext_call"]
        assert got["total_kmers"][u] == r["total_kmers"] and got["hit_groups"][u] == r["hit_groups"]
    frac = (got["call"] != 0).mean()
    assert 0.5 < frac < 0.9  # 75 % of the fragments come from a genome in the database
    small_db.confidence = 0.0


def test_db_files_roundtrip(oracle, small_db, tmp_path):
    import os
    d = str(tmp_path / "db")
    small_db.save(d)
    assert os.path.getsize(os.path.join(d, "opts.k2d")) == 64
    assert os.path.getsize(os.path.join(d, "hash.k2d")) == 32 + 4 * int(small_db.cht.capacity)
    with open(os.path.join(d, "taxo.k2d"), "rb") as f:
        assert f.read(8) == b"K2TAXDAT"
    db2 = oracle.OracleDb.load(d)
    assert np.array_equal(db2.cells(), small_db.cells())
    assert np.array_equal(db2.parents(), small_db.parents())
    assert np.array_equal(db2.external_ids(), small_db.external_ids())
