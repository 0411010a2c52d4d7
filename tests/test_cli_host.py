"""CLI drop-in surface that needs no GPU (hidden --plan prints what would run):
output naming and format precedence (reference src/main.rs:238-338), database
resolution (src/main.rs:393-434, src/lib.rs:119-141, src/download.rs:178-234 —
same cases as the reference's own tests at src/download.rs:508-549 and
src/main.rs:447-454), confidence parsing (src/lib.rs:203-221)."""
import gzip
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CLI = os.path.join(ROOT, "nohuman_b200", "bin", "nohuman")


@pytest.fixture(scope="module", autouse=True)
def built():
    if not os.path.exists(CLI):
        from nohuman_b200 import build
        build.build()
    assert os.path.exists(CLI)


def fake_db(path, version=None, added=None):
    os.makedirs(path, exist_ok=True)
    for n in ("hash.k2d", "opts.k2d", "taxo.k2d"):  # the reference only checks existence (src/download.rs:510-512)
        open(os.path.join(path, n), "wb").close()
    if version:
        # metadata sits in the install directory; the tarball may unpack into its db/ subdirectory
        meta_dir = os.path.dirname(path) if os.path.basename(path) == "db" else path
        with open(os.path.join(meta_dir, "nohuman-db.toml"), "w") as f:
            f.write(f'version = "{version}"\nadded = "{added}"\n')


def plan(args, env=None, cwd=None):
    e = dict(os.environ)
    e.pop("NOHUMAN_DB", None)
    e.update(env or {})
    r = subprocess.run([CLI, "--plan", *args], capture_output=True, text=True, env=e, cwd=cwd)
    kv = dict(l.split("=", 1) for l in r.stdout.splitlines() if "=" in l)
    return r, kv


def touch(p, data=b"@r\nACGT\n+\nIIII\n"):
    os.makedirs(os.path.dirname(p), exist_ok=True)
    with open(p, "wb") as f:
        f.write(data)
    return p


def test_default_output_names_and_format_precedence(tmp_path):
    db = str(tmp_path / "db")
    fake_db(db)
    d = str(tmp_path / "a" / "b")
    plain = touch(os.path.join(d, "in_1.fastq"))
    gz = touch(os.path.join(d, "in_1.fastq.gz"), gzip.compress(b"@r\nACGT\n+\nIIII\n"))
    gz2 = touch(os.path.join(d, "in_2.fastq.gz"), gzip.compress(b"@r\nACGT\n+\nIIII\n"))
    noext = touch(os.path.join(d, "reads"))
    # magic bytes of input[0] decide when neither -F nor --out1 is given
    _, kv = plan(["--db", db, plain])
    assert (kv["format"], kv["out1"]) == ("u", os.path.join(d, "in_1.nohuman.fq"))
    _, kv = plan(["--db", db, gz])
    assert (kv["format"], kv["out1"]) == ("g", os.path.join(d, "in_1.nohuman.fq.gz"))
    _, kv = plan(["--db", db, noext])
    assert kv["out1"] == os.path.join(d, "reads.nohuman.fq")
    # paired: each mate named from its own input, same format for both
    _, kv = plan(["--db", db, gz, gz2])
    assert kv["paired"] == "1"
    assert kv["out2"] == os.path.join(d, "in_2.nohuman.fq.gz")
    # -F wins over everything; --out1's extension wins over the input's magic; --out2's extension is never consulted
    _, kv = plan(["--db", db, "-F", "z", gz])
    assert (kv["format"], kv["out1"]) == ("z", os.path.join(d, "in_1.nohuman.fq.zst"))
    _, kv = plan(["--db", db, "-o", str(tmp_path / "x.fq.bz2"), "-O", str(tmp_path / "y.fq.xz"), gz, gz2])
    assert kv["format"] == "b" and kv["out2"].endswith("y.fq.xz")
    _, kv = plan(["--db", db, "-o", str(tmp_path / "x.fq"), gz])
    assert kv["format"] == "u"
    _, kv = plan(["--db", db, "-F", "G", plain])  # case-insensitive (src/compression.rs:396-417)
    assert kv["format"] == "g"
    r, _ = plan(["--db", db, "-F", "q", plain])
    assert r.returncode != 0 and "Invalid compression format" in r.stderr
    # a '.zstd' input: from_path says Zstd whose extension string is 'zst' != 'zstd', so only one suffix is stripped
    zs = touch(os.path.join(d, "x.fq.zstd"), b"\x28\xb5\x2f\xfd\x00")
    _, kv = plan(["--db", db, zs])
    assert kv["out1"] == os.path.join(d, "x.fq.nohuman.fq.zst")
    # inputs shorter than five bytes fail format detection (src/compression.rs:277-280)
    tiny = touch(os.path.join(d, "tiny.fq"), b"@r\n")
    r, _ = plan(["--db", db, tiny])
    assert r.returncode != 0 and "first five bytes" in r.stderr


def test_argument_validation(tmp_path):
    db = str(tmp_path / "db")
    fake_db(db)
    inp = touch(str(tmp_path / "r.fq"))
    r, _ = plan(["--db", db, str(tmp_path / "missing.fq")])
    assert r.returncode != 0 and "does not exist" in r.stderr
    r, _ = plan(["--db", db, inp, inp, inp])
    assert r.returncode != 0 and "Only one or two input files are allowed" in r.stderr
    r, _ = plan(["--db", db])
    assert r.returncode != 0 and "No input files provided" in r.stderr
    r, _ = plan(["--db", db, "-t", "0", inp])
    assert r.returncode != 0
    for bad in ("1.1", "-0.1", "abc"):
        r, _ = plan(["--db", db, f"--conf={bad}", inp])
        assert r.returncode != 0 and "Confidence score" in r.stderr
    for good, want in (("0.5", 0.5), ("1.0", 1.0), ("0.0", 0.0), ("0.1", 0.1), ("0.3", 0.3)):
        r, kv = plan(["--db", db, "-C", good, inp])
        assert r.returncode == 0 and float(kv["confidence"]) == want  # the double kraken2 would parse
    _, kv = plan(["--db", db, "-H", "-t", "7", inp])
    assert kv["keep_human"] == "1" and kv["threads"] == "7"


def test_database_resolution(tmp_path):
    inp = touch(str(tmp_path / "r.fq"))
    root = str(tmp_path / "root")
    # nothing installed
    r, _ = plan(["--db", root, inp])
    assert r.returncode != 0 and "Database does not exist" in r.stderr
    # --db pointing straight at a database, or at a directory with a db/ subdirectory
    fake_db(os.path.join(root, "direct"))
    fake_db(os.path.join(root, "nested", "db"))
    _, kv = plan(["--db", os.path.join(root, "direct"), inp])
    assert kv["db"] == os.path.join(root, "direct") and kv["version"] == ""
    _, kv = plan(["--db", os.path.join(root, "nested"), inp])
    assert kv["db"] == os.path.join(root, "nested", "db")
    # NOHUMAN_DB populates --db (src/main.rs:447-454); the flag wins over the variable
    _, kv = plan([inp], env={"NOHUMAN_DB": os.path.join(root, "direct")})
    assert kv["db"] == os.path.join(root, "direct")
    _, kv = plan(["--db", os.path.join(root, "nested"), inp], env={"NOHUMAN_DB": os.path.join(root, "direct")})
    assert kv["db"] == os.path.join(root, "nested", "db")
    # versioned installs: newest `added` wins, --db-version selects, 'all' is rejected
    inst = str(tmp_path / "installs")
    fake_db(os.path.join(inst, "HPRC.r1", "db"), "HPRC.r1", "2023-06-01")
    fake_db(os.path.join(inst, "HPRC.r2", "db"), "HPRC.r2", "2025-02-01")
    os.makedirs(os.path.join(inst, "junk"))
    _, kv = plan(["--db", inst, inp])
    assert kv["version"] == "HPRC.r2" and kv["db"] == os.path.join(inst, "HPRC.r2", "db")
    _, kv = plan(["--db", inst, "--db-version", "HPRC.r1", inp])
    assert kv["version"] == "HPRC.r1"
    r, _ = plan(["--db", inst, "--db-version", "HPRC.r9", inp])
    assert r.returncode != 0 and "is not installed" in r.stderr
    r, _ = plan(["--db", inst, "--db-version", "all", inp])
    assert r.returncode != 0 and "Cannot run with `--db-version all`" in r.stderr
    # legacy layout: database files directly in the root, no metadata (src/download.rs:508-516)
    legacy = str(tmp_path / "legacy")
    fake_db(legacy)
    _, kv = plan(["--db", legacy, "--db-version", "legacy", inp])
    assert kv["version"] == "legacy" and kv["db"] == legacy


def test_list_db_versions_reads_the_local_manifest(tmp_path):
    """src/main.rs:123-147 over download_config (src/download.rs:146-168): a config.toml in the working
    directory is used before any network access; the listing format is the reference's.  The manifest
    below is the reference's own config.toml (config.toml:1-19)."""
    (tmp_path / "config.toml").write_text(
        'default_version = "HPRC.r2"\n\n[[databases]]\nversion = "HPRC.r2"\n'
        'url = "https://ndownloader.figshare.com/files/59658306"\nmd5 = "bda8fb2ffb1a0b4cbeb880cbc4d79fc6"\nadded = "2025-11-19"\n\n'
        '[[databases]]\nversion = "HPRC.r1"\nurl = "https://zenodo.org/records/17626846/files/k2_HPRC_release1_20251110.tar.gz"\n'
        'md5 = "1cdf55d3739729fce4012519cc4706f1"\nadded = "2025-11-17"\n\n'
        '[[databases]]\nversion = "HPRC.r1.masked"\nurl = "https://zenodo.org/records/8339732/files/k2_HPRC_20230810.tar.gz"\n'
        'md5 = "87275d884181cfb6b46fdb883195dacb"\nadded = "2023-08-10"\n')
    r = subprocess.run([CLI, "--list-db-versions"], capture_output=True, text=True, cwd=tmp_path)
    assert r.returncode == 0, r.stderr
    assert r.stdout.splitlines() == [
        "Available databases:",
        "- HPRC.r2 (default) (added 2025-11-19) -> https://ndownloader.figshare.com/files/59658306",
        "- HPRC.r1 (added 2025-11-17) -> https://zenodo.org/records/17626846/files/k2_HPRC_release1_20251110.tar.gz",
        "- HPRC.r1.masked (added 2023-08-10) -> https://zenodo.org/records/8339732/files/k2_HPRC_20230810.tar.gz",
    ]
    # a release without a date, or with a malformed one, is refused like parse_added_date does (src/download.rs:290-292)
    bad = tmp_path / "bad"
    bad.mkdir()
    (bad / "config.toml").write_text('[[databases]]\nversion = "x"\nurl = "u"\nmd5 = "m"\nadded = "yesterday"\n')
    r = subprocess.run([CLI, "--list-db-versions"], capture_output=True, text=True, cwd=bad)
    assert r.returncode != 0 and "Failed to download database manifest" in r.stderr


def test_network_only_flags_say_so(tmp_path):
    r = subprocess.run([CLI, "--download"], capture_output=True, text=True, cwd=tmp_path)
    assert r.returncode != 0 and "network" in r.stderr
    # no local manifest: the reference would fetch it; this build says that it does not
    r = subprocess.run([CLI, "--list-db-versions"], capture_output=True, text=True, cwd=tmp_path)
    assert r.returncode != 0 and "network" in r.stderr
