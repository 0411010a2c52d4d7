"""Size-independent properties of the CUDA path at sizes the oracle is too slow for
(BASELINE-scale batches on a GPU-built table), plus a randomized ragged-input sweep against the oracle."""
import numpy as np
import pytest

import synth

pytestmark = pytest.mark.gpu
COMP = np.zeros(256, np.uint8)
COMP[:] = ord("N")
for a, b in zip(b"ACGTN", b"TGCAN"):
    COMP[a] = b


@pytest.fixture(scope="module")
def big():
    import torch
    from nohuman_b200 import synth as gsynth
    sdb = gsynth.build_synthetic_db(capacity=(1 << 26) + 7, device=0, block_bases=1 << 16)
    n_pairs, L = 1_000_000, 150
    n_seqs = 2 * n_pairs
    d_off = torch.arange(n_seqs + 1, dtype=torch.int64, device="cuda") * L
    d_bases = torch.zeros(n_seqs * L + 64, dtype=torch.uint8, device="cuda")
    gsynth.synth_reads(0, d_bases.data_ptr(), d_off.data_ptr(), n_seqs, sdb.genome_seed, sdb.genome_bases, seed=17,
                       paired=True, n_rate=0.02)
    torch.cuda.synchronize()
    bases = d_bases[:n_seqs * L].cpu().numpy()
    offsets = d_off.cpu().numpy().astype(np.uint64)
    yield sdb, bases, offsets, n_pairs, L
    sdb.db.close()


def classify(sdb, bases, offsets, **kw):
    from nohuman_b200 import Session
    n_seqs = len(offsets) - 1
    with Session(sdb.db, max_batch_bases=len(bases) + 4096, max_batch_seqs=n_seqs, **kw) as sess:
        call, keep, st = sess.classify(bases, offsets)
    return call, keep, st


def test_properties_at_scale(big):
    sdb, bases, offsets, n_pairs, L = big
    call, keep, st = classify(sdb, bases, offsets, paired=True, confidence=0.5)
    assert st.n_units == n_pairs and 0.45 < st.n_classified / n_pairs < 0.55
    np.testing.assert_array_equal(keep, (call == 0).astype(np.uint8))
    # determinism
    call2, _, _ = classify(sdb, bases, offsets, paired=True, confidence=0.5)
    np.testing.assert_array_equal(call, call2)
    # keep-human is the complement
    _, keep_h, st_h = classify(sdb, bases, offsets, paired=True, confidence=0.5, keep_human=True)
    np.testing.assert_array_equal(keep_h, 1 - keep)
    assert st_h.n_kept == n_pairs - st.n_kept
    # strand symmetry: canonical minimizers make a read and its reverse complement equivalent.
    # Only for reads without ambiguous bases: kraken2's ambiguity rule looks at the k-1 bases
    # BEFORE a k-mer's last base, which is not symmetric under reversal.
    seqs = bases.reshape(-1, L)
    clean = ~(seqs == ord("N")).any(axis=1).reshape(n_pairs, 2).any(axis=1)
    assert 0.9 < clean.mean() < 1.0
    rc = np.ascontiguousarray(COMP[seqs[:, ::-1]]).reshape(-1)
    call_rc, _, _ = classify(sdb, rc, offsets, paired=True, confidence=0.5)
    np.testing.assert_array_equal(call[clean], call_rc[clean])
    assert (call != call_rc).sum() < 50
    # mate order does not matter: hit counts are pooled over the pair
    swapped = np.ascontiguousarray(seqs.reshape(n_pairs, 2, L)[:, ::-1, :]).reshape(-1)
    call_sw, _, _ = classify(sdb, swapped, offsets, paired=True, confidence=0.5)
    np.testing.assert_array_equal(call, call_sw)
    # splitting the batch changes nothing (tiles, groups of 32 and deferred lists are batch-local)
    cut = 2 * 333_333
    a, _, _ = classify(sdb, bases[:cut * L], offsets[:cut + 1], paired=True, confidence=0.5)
    b, _, _ = classify(sdb, bases[cut * L:], offsets[cut:] - offsets[cut], paired=True, confidence=0.5)
    np.testing.assert_array_equal(call, np.concatenate([a, b]))
    # raising the confidence threshold only ever moves calls up the tree or to unclassified
    lo, _, _ = classify(sdb, bases, offsets, paired=True, confidence=0.0)
    hi, _, _ = classify(sdb, bases, offsets, paired=True, confidence=1.0)
    assert (lo != 0).sum() >= (call != 0).sum() >= (hi != 0).sum()
    assert not ((lo == 0) & (call != 0)).any()
    # single-end view of the same reads: every mate of a classified-at-conf-0 pair... at least the pair total is bounded
    se, _, st_se = classify(sdb, bases, offsets, paired=False, confidence=0.0)
    assert st_se.n_units == 2 * n_pairs


def test_random_ragged_inputs_match_oracle(small_db, gpu_db):
    """many small batches of adversarial shapes: lengths around k, l, the tile size and the 32-tile groups,
    N runs, lower case, junk bytes, empty reads; single and paired"""
    from nohuman_b200 import Session
    g = [x[1] for x in small_db.genomes]
    for seed in range(12):
        rng = np.random.default_rng(100 + seed)
        paired = bool(seed % 2)
        seqs = []
        for _ in range(int(rng.integers(1, 400))):
            kind = rng.integers(0, 6)
            if kind == 0:
                L = int(rng.choice([0, 1, 30, 31, 34, 35, 36, 66, 67, 68, 286, 287, 288, 289]))
            elif kind == 1:
                L = int(rng.integers(0, 700))
            else:
                L = int(rng.integers(100, 320))
            src = g[int(rng.integers(0, len(g)))]
            if rng.random() < 0.7 and L <= len(src):
                p = int(rng.integers(0, len(src) - L + 1))
                s = src[p:p + L].copy()
            else:
                s = synth.random_genome(rng, L)
            if L and rng.random() < 0.3:
                for q in rng.integers(0, L, size=int(rng.integers(1, 4))):
                    s[q:q + int(rng.integers(1, 6))] = rng.choice(np.frombuffer(b"NnRY-.\x00*", np.uint8))
            if L and rng.random() < 0.2:
                s |= 0x20  # lower case (changes N-like bytes too; still ambiguous)
            seqs.append(s)
        if paired and len(seqs) % 2:
            seqs.append(seqs[-1][:50])
        conf = float(rng.choice([0.0, 0.05, 0.33, 0.5, 1.0]))
        bases, offsets = synth.pack(seqs)
        small_db.confidence = conf
        want = small_db.classify_batch(bases, offsets, paired=paired)
        with Session(gpu_db, confidence=conf, paired=paired) as sess:
            call, keep, st = sess.classify(bases, offsets)
            icall, tk, hg = sess.debug_last_batch(len(call))
        np.testing.assert_array_equal(tk, want["total_kmers"], err_msg=f"seed {seed}")
        np.testing.assert_array_equal(hg, want["hit_groups"], err_msg=f"seed {seed}")
        np.testing.assert_array_equal(call, want["ext"], err_msg=f"seed {seed}")
    small_db.confidence = 0.0
