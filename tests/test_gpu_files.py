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


def kraken_line(db, name, seqs, paired):
    r = db.classify_one(bytes(seqs[0]), bytes(seqs[1]) if paired else None, want_taxa=True)
    if paired and len(name) > 2 and name[-2] == "/" and name[-1] in "12":
        name = name[:-2]
    lens = "|".join(str(len(s)) for s in seqs)
    return f"{'C' if r['call'] else 'U'}\t{name}\t{r['ext_call']}\t{lens}\t{r['hitlist']}\n", r["call"]


def py_report(db, calls_internal, total):
    """kraken2 reports.cc ReportKrakenStyle restated (recalled upstream behaviour, see module docstring of the product)"""
    par = db.parents().tolist()
    ext = db.external_ids().tolist()
    n = len(par)
    import ctypes as C
    names = [C.string_at(db.tax.name_data + db.tax.nodes[i].name_offset).decode() if i else "" for i in range(n)]
    ranks = [C.string_at(db.tax.rank_data + db.tax.nodes[i].rank_offset).decode() if i else "" for i in range(n)]
    direct = [0] * n
    for c in calls_internal:
        if c:
            direct[c] += 1
    clade = list(direct)
    for i in range(n - 1, 1, -1):
        clade[par[i]] += clade[i]
    kids = [[] for _ in range(n)]
    for i in range(2, n):
        kids[par[i]].append(i)
    out = []
    uncls = total - sum(direct)
    fmt = lambda cl, d, rk, tid, name, depth: "%6.2f\t%d\t%d\t%s\t%d\t%s%s\n" % (100.0 * cl / total, cl, d, rk, tid, "  " * depth, name)
    if uncls:
        out.append(fmt(uncls, uncls, "U", 0, "unclassified", 0))
    codes = {"superkingdom": "D", "kingdom": "K", "phylum": "P", "class": "C", "order": "O", "family": "F", "genus": "G", "species": "S"}

    def dfs(t, code, rdepth, depth):
        if clade[t] == 0:
            return
        if ranks[t] in codes:
            code, rdepth = codes[ranks[t]], 0
        else:
            rdepth += 1
        out.append(fmt(clade[t], direct[t], code + (str(rdepth) if rdepth else ""), ext[t], names[t], depth))
        for c in sorted(kids[t], key=lambda x: -clade[x]):
            dfs(c, code, rdepth, depth + 1)
    dfs(1, "R", -1, 0)
    return "".join(out)


@pytest.mark.parametrize("paired", [False, True])
def test_kraken_output_lines_and_report(small_db, gpu_db, reads, tmp_path, paired):
    """--kraken-output (per-read lines with the run-length hitlist) and --kraken-report, rows (f)-2/3 of SURVEY §8"""
    from nohuman_b200 import Session
    from nohuman_b200.api import make_files
    from nohuman_b200._ffi import RunStats, check, lib
    import ctypes as C
    small_db.confidence = 0.1
    if paired:
        m1, m2 = reads["m1"][:600], reads["m2"][:600]
        # ragged cases: a mate shorter than k, a mate with N runs
        m1 = m1 + [reads["m1"][0][:20], reads["m1"][1]]
        bad = reads["m2"][1].copy(); bad[40:45] = ord("N"); bad[100] = ord("n")
        m2 = m2 + [reads["m2"][0], bad]
        p1, p2 = str(tmp_path / "k_1.fq"), str(tmp_path / "k_2.fq")
        open(p1, "wb").write(fastq_bytes(m1, "q", "/1"))
        open(p2, "wb").write(fastq_bytes(m2, "q", "/2"))
        units = [(f"q{i}/1", (m1[i], m2[i])) for i in range(len(m1))]
    else:
        se = reads["se"][:800] + synth.ont_reads(small_db.genomes, 6, seed=9, n50=1500, max_len=5000)
        p1, p2 = str(tmp_path / "k.fq"), None
        open(p1, "wb").write(fastq_bytes(se, "q"))
        units = [(f"q{i}", (se[i],)) for i in range(len(se))]
    want_lines, calls = [], []
    for name, seqs in units:
        line, call = kraken_line(small_db, name, seqs, paired)
        want_lines.append(line)
        calls.append(call)
    kout, krep = str(tmp_path / "kraken.out"), str(tmp_path / "kraken.report")
    f = make_files(p1, str(tmp_path / "o1.fq"), p2, str(tmp_path / "o2.fq") if paired else None)
    f.kraken_output, f.kraken_report = kout.encode(), krep.encode()
    st = RunStats()
    with Session(gpu_db, confidence=0.1, paired=paired, threads=2) as sess:
        check(lib().nh_run_files(sess._h, C.byref(f), C.byref(st)))
    got = open(kout).read().splitlines(keepends=True)
    assert len(got) == len(want_lines)
    for g, w in zip(got, want_lines):
        assert g == w
    assert any(" A:" in l for l in got) and (not paired or all("|:|" in l for l in got))
    assert open(krep).read() == py_report(small_db, calls, len(units))
    rep = open(krep).read().splitlines()
    assert rep[0].split("\t")[3] == "U" and any(l.split("\t")[3] == "S" and l.endswith("Homo sapiens") for l in rep)
    small_db.confidence = 0.0


def test_last_batch_runs_match_oracle_taxa(small_db, gpu_db):
    from nohuman_b200 import Session
    seqs = synth.illumina_reads(small_db.genomes, 200, 150, seed=51, n_rate=0.3)
    seqs += synth.ont_reads(small_db.genomes, 5, seed=52, n50=2000, max_len=6000)
    bases, offsets = synth.pack(seqs)
    with Session(gpu_db, emit_runs=True) as sess:
        sess.classify(bases, offsets)
        first, ext, ln = sess.last_batch_runs(len(seqs), int(offsets[-1]))
    ext_ids = small_db.external_ids()
    for i, s in enumerate(seqs):
        taxa = small_db.classify_one(bytes(s), want_taxa=True)["taxa"]
        # oracle positions without the ambiguous ones, run-length encoded on the external id
        want = []
        for t in taxa:
            if t >= (1 << 64) - 3:
                continue
            e = int(ext_ids[t])
            if want and want[-1][0] == e:
                want[-1][1] += 1
            else:
                want.append([e, 1])
        got = []
        for j in range(int(first[i]), int(first[i + 1])):
            if got and got[-1][0] == int(ext[j]):
                got[-1][1] += int(ln[j])
            else:
                got.append([int(ext[j]), int(ln[j])])
        assert got == want, i
