"""End to end through files on the GPU: nh_run_files (C ABI), the `nohuman`
CLI and the `kraken2` argv shim, checked against the oracle's per-read calls
and kraken2's output rules."""
import gzip
import os
import re
import subprocess

import numpy as np
import pytest

import synth

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BIN = os.path.join(ROOT, "nohuman_b200", "bin")


def fastq_bytes(seqs, prefix, suffix=""):
    out = bytearray()
    for i, s in enumerate(seqs):
        out += f"@{prefix}{i}{suffix} len={len(s)}\n".encode() + bytes(s) + b"\n+\n" + b"I" * len(s) + b"\n"
    return bytes(out)


def expected_fastq(seqs, prefix, suffix, keep, ext):
    out = bytearray()
    for i, s in enumerate(seqs):
        if not keep[i]:
            continue
        h = f"@{prefix}{i}{suffix} len={len(s)}".encode()
        if ext[i]:
            h += b" kraken:taxid|%d" % ext[i]
        out += h + b"\n" + bytes(s) + b"\n+\n" + b"I" * len(s) + b"\n"
    return bytes(out)


@pytest.fixture(scope="module")
def reads(small_db, tmp_path_factory):
    d = tmp_path_factory.mktemp("fq")
    se = synth.illumina_reads(small_db.genomes, 3000, 150, seed=41, n_rate=0.05)
    se += [np.zeros(0, np.uint8)[:0], synth.random_genome(np.random.default_rng(1), 20)]
    se = [s for s in se if len(s) > 0]  # FASTQ cannot carry an empty sequence line unambiguously here
    pe = synth.illumina_reads(small_db.genomes, 2500, 150, seed=42, paired=True)
    m1, m2 = pe[0::2], pe[1::2]
    paths = {"se": str(d / "sample.fastq.gz"), "m1": str(d / "s_1.fq"), "m2": str(d / "s_2.fq")}
    with open(paths["se"], "wb") as f:
        f.write(gzip.compress(fastq_bytes(se, "r")))
    with open(paths["m1"], "wb") as f:
        f.write(fastq_bytes(m1, "p", "/1"))
    with open(paths["m2"], "wb") as f:
        f.write(fastq_bytes(m2, "p", "/2"))
    return dict(se=se, m1=m1, m2=m2, pe=pe, paths=paths, dir=str(d))


def oracle_calls(db, seqs, paired, conf):
    bases, offsets = synth.pack(seqs)
    db.confidence = conf
    r = db.classify_batch(bases, offsets, paired=paired)
    db.confidence = 0.0
    return r["ext"]


def test_run_files_single_end_gzip(small_db, gpu_db, reads, tmp_path):
    from nohuman_b200 import Session
    ext = oracle_calls(small_db, reads["se"], False, 0.0)
    out = str(tmp_path / "out.fq.gz")
    with Session(gpu_db, threads=4) as sess:
        st = sess.run_files(reads["paths"]["se"], out, out_format="g")
    assert (st.total, st.classified) == (len(reads["se"]), int((ext != 0).sum()))
    assert st.unclassified == st.total - st.classified
    got = gzip.decompress(open(out, "rb").read())
    assert got == expected_fastq(reads["se"], "r", "", ext == 0, np.zeros_like(ext))
    assert 0.1 < st.unclassified / st.total < 0.5


def test_cli_paired_conf_and_default_names(small_db, reads):
    ext = oracle_calls(small_db, reads["pe"], True, 0.5)
    for f in ("s_1.nohuman.fq", "s_2.nohuman.fq"):
        p = os.path.join(reads["dir"], f)
        if os.path.exists(p):
            os.remove(p)
    r = subprocess.run([os.path.join(BIN, "nohuman"), "--db", small_db.path, "-t", "4", "--conf", "0.5",
                        reads["paths"]["m1"], reads["paths"]["m2"]], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    n_cls = int((ext != 0).sum())
    assert f"{n_cls} / {len(ext)} (" in r.stderr and "sequences classified as human" in r.stderr
    assert "Removing human reads..." in r.stderr and "Done." in r.stderr
    keep = ext == 0
    zero = np.zeros_like(ext)
    assert open(os.path.join(reads["dir"], "s_1.nohuman.fq"), "rb").read() == expected_fastq(reads["m1"], "p", "/1", keep, zero)
    assert open(os.path.join(reads["dir"], "s_2.nohuman.fq"), "rb").read() == expected_fastq(reads["m2"], "p", "/2", keep, zero)


def test_cli_keep_human_tags_headers(small_db, reads, tmp_path):
    ext = oracle_calls(small_db, reads["se"], False, 0.1)
    out = str(tmp_path / "human.fq")
    r = subprocess.run([os.path.join(BIN, "nohuman"), "-D", small_db.path, "-H", "-C", "0.1", "-o", out,
                        reads["paths"]["se"]], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    assert "Keeping human reads..." in r.stderr
    assert open(out, "rb").read() == expected_fastq(reads["se"], "r", "", ext != 0, ext)
    assert len(set(ext[ext != 0].tolist())) > 1  # several taxa in the tags, not only 9606


def test_kraken2_shim_with_nohumans_argv(small_db, reads, tmp_path):
    """the argv of src/main.rs:215-267, the '#' expansion and the stderr lines of src/lib.rs:61-97"""
    ext = oracle_calls(small_db, reads["pe"], True, 0.5)
    tmpl = str(tmp_path / "kraken_out#.fq")
    r = subprocess.run([os.path.join(BIN, "kraken2"), "--threads", "2", "--db", small_db.path, "--output", "/dev/null",
                        "--confidence", "0.5", "--paired", "--unclassified-out", tmpl,
                        reads["paths"]["m1"], reads["paths"]["m2"]], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    total = cls = uncls = None
    for line in r.stderr.splitlines():  # parse_kraken_stderr
        tok = line.split()[0].replace(",", "") if line.split() else "0"
        if "processed" in line:
            total = int(tok)
        elif "sequences classified" in line:
            cls = int(tok)
        elif "sequences unclassified" in line:
            uncls = int(tok)
    assert (total, cls, uncls) == (len(ext), int((ext != 0).sum()), int((ext == 0).sum()))
    keep, zero = ext == 0, np.zeros_like(ext)
    assert open(str(tmp_path / "kraken_out_1.fq"), "rb").read() == expected_fastq(reads["m1"], "p", "/1", keep, zero)
    assert open(str(tmp_path / "kraken_out_2.fq"), "rb").read() == expected_fastq(reads["m2"], "p", "/2", keep, zero)
    # missing '#' with --paired is an error, as in kraken2
    r = subprocess.run([os.path.join(BIN, "kraken2"), "--db", small_db.path, "--paired", "--unclassified-out",
                        str(tmp_path / "x.fq"), reads["paths"]["m1"], reads["paths"]["m2"]], capture_output=True, text=True)
    assert r.returncode != 0 and "#" in r.stderr


def test_cli_check_and_errors(small_db, tmp_path):
    r = subprocess.run([os.path.join(BIN, "nohuman"), "--check"], capture_output=True, text=True)
    assert r.returncode == 0 and "All dependencies are available" in r.stderr
    bad = tmp_path / "bad.fq"
    bad.write_bytes(b"not a fastq file\n")
    r = subprocess.run([os.path.join(BIN, "nohuman"), "--db", small_db.path, str(bad)], capture_output=True, text=True)
    assert r.returncode != 0 and "format not recognised" in r.stderr


def test_replicated_database_and_sharded_batches(small_db, gpu_db, reads, tmp_path):
    """nh_db_clone + nh_run_files_multi: replicas (on a second GPU when there is one, else on the
    same device) classify batches dealt to them; output and counts equal the single-session run."""
    from nohuman_b200 import Session, _ffi
    from nohuman_b200.api import run_files_multi
    n_dev = _ffi.lib().nh_device_count()
    ext = oracle_calls(small_db, reads["pe"], True, 0.5)
    replica = gpu_db.clone(1 if n_dev > 1 else 0)
    try:
        assert replica.info.capacity == gpu_db.info.capacity and replica.info.device == (1 if n_dev > 1 else 0)
        with Session(gpu_db, confidence=0.5, paired=True, threads=2) as s0, \
                Session(replica, confidence=0.5, paired=True, threads=2) as s1:
            # the replica answers like the original
            keys = np.arange(1, 5000, dtype=np.uint64) * np.uint64(0x9E3779B97F4A7C15)
            np.testing.assert_array_equal(s0.debug_probe(keys), s1.debug_probe(keys))
            o1, o2 = str(tmp_path / "m1.fq.gz"), str(tmp_path / "m2.fq.gz")
            st = run_files_multi([s0, s1], reads["paths"]["m1"], o1, reads["paths"]["m2"], o2, out_format="g")
        assert (st.total, st.classified) == (len(ext), int((ext != 0).sum()))
        keep, zero = ext == 0, np.zeros_like(ext)
        assert gzip.decompress(open(o1, "rb").read()) == expected_fastq(reads["m1"], "p", "/1", keep, zero)
        assert gzip.decompress(open(o2, "rb").read()) == expected_fastq(reads["m2"], "p", "/2", keep, zero)
    finally:
        replica.close()
    r = subprocess.run([os.path.join(BIN, "nohuman"), "--db", small_db.path, "--gpus", "all", "--conf", "0.5",
                        "-o", str(tmp_path / "c1.fq"), "-O", str(tmp_path / "c2.fq"),
                        reads["paths"]["m1"], reads["paths"]["m2"]], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    assert open(str(tmp_path / "c1.fq"), "rb").read() == expected_fastq(reads["m1"], "p", "/1", ext == 0, np.zeros_like(ext))
