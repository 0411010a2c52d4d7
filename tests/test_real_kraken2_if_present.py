""""at run time, first probe `command -v kraken2`; if ever present, diff against it and
promote it to oracle".  No kraken2 binary exists in this image or on the GPU boxes (no network), so this
test skips here; on a machine that has the real kraken2 (the version nohuman pins is 2.17, reference
Dockerfile:15) it checks the oracle's per-read calls and hit lists against it on the same database
directory and reads — the step that would turn "parity unpinned" into pinned."""
import os
import shutil
import subprocess

import pytest

import synth


def real_kraken2():
    exe = shutil.which("kraken2")
    if not exe:
        return None
    try:
        v = subprocess.run([exe, "--version"], capture_output=True, text=True, timeout=20).stdout
    except Exception:
        return None
    if "Kraken version" not in v or "libnohuman_gpu" in v:  # our own argv shim does not count
        return None
    return exe


@pytest.mark.skipif(real_kraken2() is None, reason="no upstream kraken2 binary on PATH (parity stays unpinned)")
def test_oracle_matches_upstream_kraken2(small_db, tmp_path):
    exe = real_kraken2()
    reads = synth.illumina_reads(small_db.genomes, 2000, 150, seed=61, n_rate=0.1)
    reads += synth.ont_reads(small_db.genomes, 20, seed=62, n50=2000, max_len=8000)
    fq = str(tmp_path / "reads.fq")
    synth.write_fastq(fq, reads)
    for conf in (0.0, 0.1, 0.5):
        out = str(tmp_path / f"kraken_{conf}.out")
        r = subprocess.run([exe, "--db", small_db.path, "--threads", "2", "--confidence", str(conf), "--output", out, fq],
                           capture_output=True, text=True)
        assert r.returncode == 0, r.stderr
        small_db.confidence = conf
        with open(out) as f:
            for i, line in enumerate(f):
                flag, rid, taxid, length, hitlist = line.rstrip("\n").split("\t")
                want = small_db.classify_one(bytes(reads[i]), want_taxa=True)
                assert rid == f"r{i}" and int(length) == len(reads[i])
                assert int(taxid) == want["ext_call"], (conf, i)
                assert flag == ("C" if want["call"] else "U")
                assert hitlist == want["hitlist"], (conf, i)
    small_db.confidence = 0.0
