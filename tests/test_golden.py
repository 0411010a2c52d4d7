"""Committed fixture tests/golden/cfg1_tiny.npz (made by tests/golden/make_golden.py):
the oracle must still reproduce it (CPU), and the CUDA path through the C ABI
must reproduce it bit-exactly (GPU) — there with the database opened from the
fixture's own hash.k2d / opts.k2d / taxo.k2d bytes, read unchanged."""
import os

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
CASES = [("se", False, 0.0), ("se", False, 0.1), ("se", False, 0.5), ("pe", True, 0.0), ("pe", True, 0.5),
         ("ont", False, 0.0), ("ont", False, 0.05)]


@pytest.fixture(scope="module")
def golden():
    return np.load(os.path.join(HERE, "golden", "cfg1_tiny.npz"))


@pytest.fixture(scope="module")
def golden_db_dir(golden, tmp_path_factory):
    d = tmp_path_factory.mktemp("golden_db")
    for key, name in (("hash_k2d", "hash.k2d"), ("opts_k2d", "opts.k2d"), ("taxo_k2d", "taxo.k2d")):
        golden[key].tofile(os.path.join(d, name))
    return str(d)


def tag(name, conf):
    return f"{name}_c{int(round(conf * 100)):03d}"


@pytest.mark.parametrize("name,paired,conf", CASES)
def test_oracle_reproduces_fixture(oracle, golden, golden_db_dir, name, paired, conf):
    db = oracle.OracleDb.load(golden_db_dir)
    db.confidence = conf
    r = db.classify_batch(golden[f"{name}_bases"], golden[f"{name}_offsets"], paired=paired)
    t = tag(name, conf)
    for f in ("ext", "call", "hit_groups", "total_kmers"):
        np.testing.assert_array_equal(r[f], golden[f"{t}_{f}"], err_msg=f)
    assert r["lookups"] == int(golden[f"{t}_lookups"][0])
    assert (r["ext"] != 0).any() and (name == "ont" or (r["ext"] == 0).any())


def test_oracle_reproduces_minimizer_stream_and_hitlists(oracle, golden, golden_db_dir):
    db = oracle.OracleDb.load(golden_db_dir)
    off = golden["se_offsets"]
    bases = golden["se_bases"]
    mins, ambs = [], []
    for i in range(20):
        m, a = oracle.scan_positions(db.opts, bytes(bases[int(off[i]):int(off[i + 1])]))
        mins.append(m)
        ambs.append(a)
    np.testing.assert_array_equal(np.concatenate(ambs), golden["se20_ambiguous"])
    ok = golden["se20_ambiguous"] == 0
    np.testing.assert_array_equal(np.concatenate(mins)[ok], golden["se20_minimizers"][ok])
    for i, want in enumerate(golden["se_hitlists"]):
        got = db.classify_one(bytes(bases[int(off[i]):int(off[i + 1])]), want_taxa=True)["hitlist"]
        assert got == str(want)


@pytest.mark.gpu
@pytest.mark.parametrize("name,paired,conf", CASES)
def test_cuda_reproduces_fixture(golden, golden_db_dir, name, paired, conf):
    from nohuman_b200 import Database, Session
    t = tag(name, conf)
    with Database.open(golden_db_dir, 0) as db:
        info = db.info
        assert (info.k, info.l, info.spaced_seed_mask) == (35, 31, 0x3FFFFFFFF3333333)
        with Session(db, confidence=conf, paired=paired) as sess:
            call, keep, st = sess.classify(golden[f"{name}_bases"], golden[f"{name}_offsets"])
            icall, tk, hg = sess.debug_last_batch(len(call))
    np.testing.assert_array_equal(call, golden[f"{t}_ext"])
    np.testing.assert_array_equal(icall, golden[f"{t}_call"])
    np.testing.assert_array_equal(tk, golden[f"{t}_total_kmers"])
    np.testing.assert_array_equal(hg, golden[f"{t}_hit_groups"])
    np.testing.assert_array_equal(keep, (golden[f"{t}_ext"] == 0).astype(np.uint8))


@pytest.mark.gpu
def test_cuda_reproduces_minimizer_stream(golden, golden_db_dir):
    from nohuman_b200 import Database, Session
    off = golden["se_offsets"][:21]
    bases = golden["se_bases"][:int(off[-1])]
    with Database.open(golden_db_dir, 0) as db, Session(db) as sess:
        mins, amb, _ = sess.debug_minimizers(bases, off)
    np.testing.assert_array_equal(amb, golden["se20_ambiguous"])
    ok = amb == 0
    np.testing.assert_array_equal(mins[ok], golden["se20_minimizers"][ok])
