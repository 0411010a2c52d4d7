"""BASELINE.json configs[0] at its stated size: `nohuman -t 4` on 10 000 synthetic 150 bp single-end
reads against a kraken2-format database built from a synthetic human chromosome (5 Mbp) and three
bacterial genomes (2 / 3 / 4 Mbp, two sharing a genus and a 200 kb segment), k=35 l=31 s=7, load 0.7.
Per-read parity with the oracle through the batch API, and byte parity of the CLI's output file."""
import os
import subprocess

import numpy as np
import pytest

import synth

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def cfg1(oracle, tmp_path_factory):
    genomes = synth.cfg1_genomes(seed=1, scale=1.0)
    tax = [oracle.TaxSpec(*t) for t in synth.TAXONOMY_CFG1]
    db = oracle.OracleDb.build([(t, bytes(g)) for t, g in genomes], tax)
    d = str(tmp_path_factory.mktemp("db_cfg1_full"))
    db.save(d)
    reads = synth.illumina_reads(genomes, 10_000, 150, seed=2)  # 50 % human, 25 % bacterial, 25 % random, 1 % with an N
    return db, d, reads


def test_cfg1_batch_parity(cfg1):
    from nohuman_b200 import Database, Session
    db, path, reads = cfg1
    assert 0.69 < db.cht.size / db.cht.capacity <= 0.7001 and db.cht.capacity > 6_000_000
    bases, offsets = synth.pack(reads)
    want = db.classify_batch(bases, offsets)
    with Database.open(path, 0) as gdb, Session(gdb) as sess:
        call, keep, st = sess.classify(bases, offsets)
        icall, tk, hg = sess.debug_last_batch(len(call))
    np.testing.assert_array_equal(call, want["ext"])
    np.testing.assert_array_equal(icall, want["call"])
    np.testing.assert_array_equal(tk, want["total_kmers"])
    np.testing.assert_array_equal(hg, want["hit_groups"])
    assert st.n_lookups == want["lookups"]
    frac = (want["ext"] != 0).mean()
    assert 0.70 < frac < 0.78  # everything sampled from a genome in the database is dropped, bacteria included
    taxa = set(want["ext"].tolist())
    assert {0, 9606, 562, 564, 1423, 561} <= taxa  # species calls and genus-level LCA calls from the shared segment


def test_cfg1_cli_t4(cfg1, tmp_path):
    db, path, reads = cfg1
    inp = str(tmp_path / "cfg1.fastq")
    with open(inp, "wb") as f:
        for i, s in enumerate(reads):
            f.write(b"@read%d\n" % i + bytes(s) + b"\n+\n" + b"F" * len(s) + b"\n")
    r = subprocess.run([os.path.join(ROOT, "nohuman_b200", "bin", "nohuman"), "-t", "4", "--db", path, inp],
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    bases, offsets = synth.pack(reads)
    ext = db.classify_batch(bases, offsets)["ext"]
    want = b"".join(b"@read%d\n" % i + bytes(s) + b"\n+\n" + b"F" * len(s) + b"\n"
                    for i, s in enumerate(reads) if ext[i] == 0)
    assert open(str(tmp_path / "cfg1.nohuman.fq"), "rb").read() == want
    assert f"{int((ext != 0).sum())} / 10000" in r.stderr
