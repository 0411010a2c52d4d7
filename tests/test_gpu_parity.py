"""Parity of the CUDA path (through the C ABI) against the CPU oracle.
Bit-exact: integer work only.  Run on the B200 box: pytest -m gpu."""
import numpy as np
import pytest

import synth

pytestmark = pytest.mark.gpu


def _compare_batch(oracle_db, sess, seqs, paired, conf):
    bases, offsets = synth.pack(seqs)
    oracle_db.confidence = conf
    want = oracle_db.classify_batch(bases, offsets, paired=paired)
    call, keep, st = sess.classify(bases, offsets)
    icall, tk, hg = sess.debug_last_batch(len(call))
    np.testing.assert_array_equal(tk, want["total_kmers"])
    np.testing.assert_array_equal(hg, want["hit_groups"])
    np.testing.assert_array_equal(icall, want["call"])
    np.testing.assert_array_equal(call, want["ext"])
    return call, keep, st, want


def test_minimizer_positions_match_scanner(small_db, gpu_db, oracle):
    from nohuman_b200 import Session
    rng = np.random.default_rng(7)
    seqs = synth.illumina_reads(small_db.genomes, 300, 150, seed=11, n_rate=0.3)
    # ragged lengths around k, l and the tile size; N runs; lower case; junk bytes
    for L in [0, 1, 30, 31, 34, 35, 36, 66, 123, 124, 125, 157, 158, 159, 160, 250, 283, 1000, 5000]:
        seqs.append(synth.random_genome(rng, L))
    s = synth.random_genome(rng, 400)
    s[50:60] = ord("N"); s[100] = ord("n"); s[135] = ord("R"); s[200:300] |= 0x20; s[399] = 0
    seqs.append(s)
    s = synth.random_genome(rng, 300)
    s[0] = ord("N"); s[34] = ord("N"); s[69] = ord("N"); s[299] = ord("N")
    seqs.append(s)
    seqs.append(np.full(200, ord("N"), np.uint8))
    seqs.append(np.full(200, ord("A"), np.uint8))
    bases, offsets = synth.pack(seqs)
    with Session(gpu_db) as sess:
        mins, amb, pos_off = sess.debug_minimizers(bases, offsets)
    for i, sq in enumerate(seqs):
        wm, wa = oracle.scan_positions(small_db.opts, bytes(sq))
        lo, hi = int(pos_off[i]), int(pos_off[i + 1])
        assert hi - lo == len(wm), f"seq {i} len {len(sq)}"
        np.testing.assert_array_equal(amb[lo:hi], wa, err_msg=f"ambig seq {i} len {len(sq)}")
        ok = wa == 0
        np.testing.assert_array_equal(mins[lo:hi][ok], wm[ok], err_msg=f"min seq {i} len {len(sq)}")


def test_probe_matches_get(small_db, gpu_db, oracle):
    from nohuman_b200 import Session
    rng = np.random.default_rng(3)
    present = []
    for _, g in small_db.genomes:
        m, a = oracle.scan_positions(small_db.opts, bytes(g[:20000]))
        present.append(m[a == 0])
    keys = np.concatenate(present + [rng.integers(0, 1 << 62, size=50000, dtype=np.uint64)])
    with Session(gpu_db) as sess:
        got = sess.debug_probe(keys)
    want = np.array([small_db.get(int(k)) for k in keys], dtype=np.uint32)
    np.testing.assert_array_equal(got, want)
    assert (want != 0).sum() > 10000


@pytest.mark.parametrize("conf", [0.0, 0.1, 0.5, 1.0])
def test_single_end_matches_oracle(small_db, gpu_db, conf):
    from nohuman_b200 import Session
    seqs = synth.illumina_reads(small_db.genomes, 4000, 150, seed=2)
    with Session(gpu_db, confidence=conf) as sess:
        call, keep, st, want = _compare_batch(small_db, sess, seqs, False, conf)
    np.testing.assert_array_equal(keep, (call == 0).astype(np.uint8))
    assert st.n_classified == int((want["call"] != 0).sum())
    assert st.n_units == 4000
    assert st.n_lookups == want["lookups"]  # 150 bp reads are single-tile: exact


@pytest.mark.parametrize("conf,keep_human", [(0.0, False), (0.5, False), (0.5, True)])
def test_paired_end_matches_oracle(small_db, gpu_db, conf, keep_human):
    from nohuman_b200 import Session
    seqs = synth.illumina_reads(small_db.genomes, 3000, 150, seed=3, paired=True)
    with Session(gpu_db, confidence=conf, paired=True, keep_human=keep_human) as sess:
        call, keep, st, want = _compare_batch(small_db, sess, seqs, True, conf)
    cls = (call != 0).astype(np.uint8)
    np.testing.assert_array_equal(keep, cls if keep_human else 1 - cls)
    assert st.n_kept == int(keep.sum())


def test_long_reads_match_oracle(small_db, gpu_db):
    from nohuman_b200 import Session
    seqs = synth.ont_reads(small_db.genomes, 300, seed=4, n50=4000, max_len=30000)
    with Session(gpu_db, confidence=0.05, keep_human=True) as sess:
        _compare_batch(small_db, sess, seqs, False, 0.05)


def test_edge_cases(small_db, gpu_db):
    from nohuman_b200 import Session
    rng = np.random.default_rng(5)
    g = small_db.genomes[0][1]
    seqs = [np.zeros(0, np.uint8), g[:34].copy(), g[:35].copy(), g[100:136].copy(),
            np.full(150, ord("N"), np.uint8), np.full(150, ord("A"), np.uint8)]
    r = g[1000:1150].copy(); r[75] = ord("N"); seqs.append(r)
    r = g[2000:2300].copy(); r[::40] = ord("N"); seqs.append(r)
    r = g[3000:3150].copy(); r |= 0x20; seqs.append(r)
    seqs += [synth.random_genome(rng, int(L)) for L in rng.integers(0, 400, size=200)]
    with Session(gpu_db) as sess:
        _compare_batch(small_db, sess, seqs, False, 0.0)
        # empty batch
        call, keep, st = sess.classify(np.zeros(0, np.uint8), np.zeros(1, np.uint64))
        assert len(call) == 0 and st.n_units == 0
    # paired with ragged mates (one mate shorter than k)
    pseqs = []
    for i in range(100):
        a = g[i * 500:i * 500 + 150].copy()
        b = synth.revcomp(g[i * 500 + 200:i * 500 + 350]) if i % 3 else g[:20].copy()
        pseqs += [a, b]
    with Session(gpu_db, paired=True, confidence=0.2) as sess:
        _compare_batch(small_db, sess, pseqs, True, 0.2)


@pytest.mark.parametrize("mode", ["legacy", "lane_taxa_1", "lane_taxa_2", "tile_pos_60", "tile_pos_256", "tile_pos_1023",
                                  "filter_mode_0", "filter_mode_2", "filter_mode_3"])
def test_kernel_path_variants_match_oracle(small_db, gpu_db, mode, monkeypatch):
    """The warp-per-tile kernels (NH_LEGACY_KERNELS=1), the streaming kernel with a shrunken
    in-warp taxon table (units overflow into k_score_big, which classifies them again from
    their bases), other tile sizes and the other two policies for the miss filter give the
    answers of the default path and of the oracle."""
    from nohuman_b200 import Session
    if mode == "legacy":
        monkeypatch.setenv("NH_LEGACY_KERNELS", "1")
    elif mode.startswith("filter_mode"):
        # the miss filter: never asked / asked by every lookup / by units whose last few lookups missed (default: by units without a hit so far)
        monkeypatch.setenv("NH_FILTER_MODE", mode[-1])
    elif mode.startswith("tile_pos"):
        monkeypatch.setenv("NH_FUSED_TILE_POS", mode.split("_")[-1])  # 60: every 150 bp read becomes a multi-tile unit
    else:
        monkeypatch.setenv("NH_TEST_LANE_TAXA", mode[-1])
    g = dict(small_db.genomes)
    seqs = synth.illumina_reads(small_db.genomes, 1500, 150, seed=31, paired=True)
    # chimeric pairs that hit several taxa at once
    for i in range(100):
        a = np.concatenate([g[9606][i * 40:i * 40 + 75], g[562][i * 30:i * 30 + 75]])
        b_ = np.concatenate([g[564][i * 50:i * 50 + 75], g[1423][i * 20:i * 20 + 75]])
        seqs += [a, b_]
    seqs += synth.ont_reads(small_db.genomes, 20, seed=5, n50=2000, max_len=8000)[:20]
    if len(seqs) % 2:
        seqs.append(seqs[-1][:100])
    with Session(gpu_db, confidence=0.1, paired=True) as sess:
        call, keep, st, want = _compare_batch(small_db, sess, seqs, True, 0.1)
        import os
        assert bool(st.fused_kernel) == (mode != "legacy" and os.environ.get("NH_LEGACY_KERNELS") != "1")
    assert st.n_classified == int((want["call"] != 0).sum())
    assert len(set(want["call"].tolist())) > 4


@pytest.mark.parametrize("tile_pos", [512, 256, 100])
def test_lengths_straddling_tile_and_group_borders(small_db, gpu_db, tile_pos, monkeypatch):
    """Reads whose k-mer positions end exactly at, one before and one after a tile border, units of
    2..33 tiles (scored in the warp when all tiles share a group of 32, by k_score otherwise), and
    pairs whose mates fall on either side of a group border."""
    from nohuman_b200 import Session
    monkeypatch.setenv("NH_FUSED_TILE_POS", str(tile_pos))
    tp = tile_pos
    k = 35
    g = dict(small_db.genomes)
    src = g[9606]
    rng = np.random.default_rng(17)
    lens = []
    for nt in (1, 2, 3, 5, 31, 32, 33):
        for d in (-1, 0, 1):
            lens.append(nt * tp + k - 1 + d)
    lens += [k - 1, k, k + 1, tp, tp + 1, 2 * tp - 1]
    seqs = []
    for rep in range(3):
        for L in lens:
            o = int(rng.integers(0, len(src) - L - 1))
            r = src[o:o + L].copy()
            if rep == 1 and L > 200:
                r[L // 2] = ord("N")  # an ambiguous stretch across a border keeps last_minimizer alive
                r[min(L - 1, tp + k - 2)] = ord("N")
            if rep == 2:
                r = synth.mutate(rng, r, 0.03)
            seqs.append(r)
    # filler of short reads so that the long units start at many different lanes
    short = synth.illumina_reads(small_db.genomes, 97, 150, seed=5)
    mixed = []
    for i, r in enumerate(seqs):
        mixed.append(r)
        mixed += short[(i * 3) % 90:(i * 3) % 90 + (i % 4)]
    with Session(gpu_db, confidence=0.05) as sess:
        _compare_batch(small_db, sess, mixed, False, 0.05)
    pairs = mixed + mixed[1:] + [mixed[0]]  # odd shift: mates of very different lengths, units across group borders
    if len(pairs) % 2:
        pairs.append(short[0])
    with Session(gpu_db, confidence=0.05, paired=True) as sess:
        _compare_batch(small_db, sess, pairs, True, 0.05)


def test_repeated_minimizers_across_tile_borders(small_db, gpu_db, monkeypatch):
    """Low-complexity sequence: the same minimizer runs across tile borders, so the hit-group count
    depends on the border repair (a tile's first lookup equals the previous tile's last one)."""
    from nohuman_b200 import Session
    monkeypatch.setenv("NH_FUSED_TILE_POS", "64")
    g = dict(small_db.genomes)
    src = g[9606]
    seqs = []
    for i in range(60):
        unit = src[i * 100:i * 100 + 40 + i % 7]
        seqs.append(np.concatenate([src[5000 + i * 10:5000 + i * 10 + 90], np.tile(unit, 12), src[9000:9100]]))
    for i in range(20):  # homopolymers and dinucleotide repeats: one minimizer for hundreds of positions
        seqs.append(np.concatenate([src[i * 50:i * 50 + 80], np.full(300, b"ACGT"[i % 4], np.uint8), src[700:800]]))
        seqs.append(np.tile(np.frombuffer(b"AC", np.uint8), 250))
    with Session(gpu_db, confidence=0.0) as sess:
        _compare_batch(small_db, sess, seqs, False, 0.0)


@pytest.mark.parametrize("mode", ["default", "lane_taxa_1", "tile_pos_96", "tile_pos_33"])
def test_packed_input_matches_ascii_and_oracle(small_db, gpu_db, mode, monkeypatch):
    """nh_classify_batch_packed (2-bit codes + validity bits packed on the host): same per-unit call, k-mer
    total and hit groups as the ASCII entry point and the oracle - ragged lengths, N runs, lower case, junk
    bytes, multi-tile reads, pairs, and (lane_taxa_1) units that overflow into k_score_big, which then scans
    the PACKED planes again."""
    from nohuman_b200 import Session
    if mode == "lane_taxa_1":
        monkeypatch.setenv("NH_TEST_LANE_TAXA", "1")
    if mode.startswith("tile_pos"):
        # 33: packed input rounds the tile size down to 32, so the batch has more tiles than the ASCII path
        # would cut (the tile arrays were once sized for 33: tools/fuzz_parity.py found the overrun)
        monkeypatch.setenv("NH_FUSED_TILE_POS", mode.split("_")[-1])
    rng = np.random.default_rng(23)
    g = dict(small_db.genomes)
    seqs = synth.illumina_reads(small_db.genomes, 1200, 150, seed=12, paired=True, n_rate=0.2)
    seqs += synth.ont_reads(small_db.genomes, 40, seed=6, n50=3000, max_len=20000)
    for L in [0, 1, 31, 32, 33, 34, 35, 36, 63, 64, 65, 511, 512, 513, 545, 546, 547, 1057, 1058, 16418, 16419]:
        o = int(rng.integers(0, len(g[9606]) - L - 1))
        seqs.append(g[9606][o:o + L].copy())
    r = g[562][500:1400].copy(); r[100:140] = ord("N"); r[300] = ord("n"); r[400:600] |= 0x20; r[700] = 0; seqs.append(r)
    for i in range(60):  # chimeras: several taxa per unit
        seqs.append(np.concatenate([g[9606][i * 40:i * 40 + 75], g[562][i * 30:i * 30 + 75], g[1423][i * 20:i * 20 + 90]]))
    if len(seqs) % 2:
        seqs.append(seqs[-1][:100].copy())
    bases, offsets = synth.pack(seqs)
    for paired, conf in ((True, 0.1), (False, 0.0)):
        small_db.confidence = conf
        want = small_db.classify_batch(bases, offsets, paired=paired)
        with Session(gpu_db, confidence=conf, paired=paired, max_batch_bases=len(bases) + 4096) as sess:
            call, keep, st = sess.classify_packed(bases, offsets, threads=3)
            icall, tk, hg = sess.debug_last_batch(len(call))
            call_a, keep_a, _ = sess.classify(bases, offsets)
            # ASCII in, packed inside the call by the session's pool (1 and 5 packing threads)
            for threads in (1, 5):
                call_p, keep_p, st_p = sess.classify_pack(bases, offsets, threads=threads)
                icall_p, tk_p, hg_p = sess.debug_last_batch(len(call_p))
                np.testing.assert_array_equal(call_p, want["ext"])
                np.testing.assert_array_equal(keep_p, keep_a)
                np.testing.assert_array_equal(tk_p, want["total_kmers"])
                np.testing.assert_array_equal(hg_p, want["hit_groups"])
                assert st_p.fused_kernel == 2
        np.testing.assert_array_equal(tk, want["total_kmers"])
        np.testing.assert_array_equal(hg, want["hit_groups"])
        np.testing.assert_array_equal(call, want["ext"])
        np.testing.assert_array_equal(call, call_a)
        np.testing.assert_array_equal(keep, keep_a)
        assert st.fused_kernel == 2


def test_pack_entry_with_a_million_short_sequences(small_db, gpu_db):
    """nh_classify_batch_pack sends 4-byte lengths and rebuilds offsets / first units on the device (k_len_*):
    1.3 M sequences (more than one pass of the single-block scan), lengths 0..90 with many empties, against the
    ASCII entry point for all of them and against the oracle for a prefix."""
    from nohuman_b200 import Session
    rng = np.random.default_rng(77)
    g = np.asarray(dict(small_db.genomes)[9606])
    n = 1_300_001
    lens = rng.integers(30, 91, n)
    lens[rng.random(n) < 0.05] = 0
    lens[rng.random(n) < 0.02] = 1
    lens[-1] = 0
    offsets = np.zeros(n + 1, np.uint64)
    np.cumsum(lens, out=offsets[1:])
    total = int(offsets[-1])
    # read i is the slice of the human genome starting at 37 i (wrapping), so most k-mers hit the table
    starts = (np.arange(n, dtype=np.int64) * 37) % (len(g) - 200)
    within = np.arange(total, dtype=np.int64) - np.repeat(offsets[:-1].astype(np.int64), lens)
    bases = g[np.repeat(starts, lens) + within].astype(np.uint8)
    bases[rng.random(total) < 0.002] = ord("N")
    small_db.confidence = 0.0
    with Session(gpu_db, confidence=0.0, paired=False, max_batch_bases=total + 4096, max_batch_seqs=n) as sess:
        call_a, keep_a, _ = sess.classify(bases, offsets)
        call_p, keep_p, st = sess.classify_pack(bases, offsets, threads=7)
        assert st.fused_kernel == 2
    np.testing.assert_array_equal(call_p, call_a)
    np.testing.assert_array_equal(keep_p, keep_a)
    m = 20_000
    want = small_db.classify_batch(bases[:int(offsets[m])], offsets[:m + 1], paired=False)
    np.testing.assert_array_equal(call_p[:m], want["ext"])
    assert (call_p != 0).sum() > n // 10
