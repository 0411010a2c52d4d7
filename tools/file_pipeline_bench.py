#!/usr/bin/env python
"""Wall-clock throughput of the FILE path (nohuman CLI -> nh_run_files) on BASELINE configs[1]-shaped
input: N pairs of 2x150 bp FASTQ, 50 % genome-derived, --conf 0.5, plain and gzip in/out.  The reads come
from the GPU sampler; the files are written with numpy.  Reports seconds and Gbp/s per variant.

    python tools/file_pipeline_bench.py [--pairs 2000000] [--threads 16] [--capacity-log2 28]
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def fastq_array(seqs2d: np.ndarray, mate: int) -> np.ndarray:
    n, L = seqs2d.shape
    hdr = np.frombuffer(b"@r000000000/%d\n" % mate, np.uint8).copy()
    rec = np.empty((n, len(hdr) + L + 3 + L + 1), np.uint8)
    rec[:, :len(hdr)] = hdr
    idx = np.arange(n)
    for d in range(9):
        rec[:, 2 + 8 - d] = 48 + (idx // 10 ** d) % 10
    o = len(hdr)
    rec[:, o:o + L] = seqs2d
    rec[:, o + L:o + L + 3] = np.frombuffer(b"\n+\n", np.uint8)
    rec[:, o + L + 3:o + 2 * L + 3] = ord("I")
    rec[:, -1] = ord("\n")
    return rec.reshape(-1)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--pairs", type=int, default=2_000_000)
    ap.add_argument("--threads", type=int, default=16)
    ap.add_argument("--capacity-log2", type=int, default=28)
    args = ap.parse_args()
    import torch
    from nohuman_b200 import synth
    torch.cuda.set_device(0)
    sdb = synth.build_synthetic_db(1 << args.capacity_log2, device=0)
    d = tempfile.mkdtemp(prefix="nh_filebench_")
    db_dir = os.path.join(d, "db")
    sdb.save(db_dir)
    n_seqs, L = 2 * args.pairs, 150
    d_off = torch.arange(n_seqs + 1, dtype=torch.int64, device="cuda") * L
    d_bases = torch.zeros(n_seqs * L + 64, dtype=torch.uint8, device="cuda")
    synth.synth_reads(0, d_bases.data_ptr(), d_off.data_ptr(), n_seqs, sdb.genome_seed, sdb.genome_bases, seed=3,
                      paired=True, n_rate=0.01)
    torch.cuda.synchronize()
    seqs = d_bases[:n_seqs * L].cpu().numpy().reshape(n_seqs, L)
    sdb.db.close()
    del d_bases
    torch.cuda.empty_cache()
    p1, p2 = os.path.join(d, "s_1.fq"), os.path.join(d, "s_2.fq")
    fastq_array(seqs[0::2], 1).tofile(p1)
    fastq_array(seqs[1::2], 2).tofile(p2)
    subprocess.check_call(["gzip", "-k", "-1", p1])
    subprocess.check_call(["gzip", "-k", "-1", p2])
    # blocked gzip (BGZF, what bgzip / bcl2fastq / this library write) of the same reads, made with the host-only hook
    from nohuman_b200.api import rewrite_files
    b1, b2 = os.path.join(d, "b_1.fq.gz"), os.path.join(d, "b_2.fq.gz")
    rewrite_files(np.ones(args.pairs, np.uint8), np.zeros(args.pairs, np.uint32), p1, b1, p2, b2, out_format="g",
                  threads=args.threads)
    cli = os.path.join(ROOT, "nohuman_b200", "bin", "nohuman")
    rows = []
    gbp = args.pairs * 2 * L / 1e9
    for name, a, b, fmt in (("plain in, plain out", p1, p2, "u"), ("plain in, gzip out", p1, p2, "g"),
                            ("gzip in, gzip out (configs[1])", p1 + ".gz", p2 + ".gz", "g"),
                            ("gzip in, plain out", p1 + ".gz", p2 + ".gz", "u"),
                            ("blocked gzip in, plain out", b1, b2, "u"), ("blocked gzip in, gzip out", b1, b2, "g")):
        o1, o2 = os.path.join(d, "o1"), os.path.join(d, "o2")
        t0 = time.perf_counter()
        r = subprocess.run([cli, "-v", "--db", db_dir, "-t", str(args.threads), "--conf", "0.5", "-F", fmt, "-o", o1, "-O", o2, a, b],
                           capture_output=True, text=True)
        dt = time.perf_counter() - t0
        assert r.returncode == 0, r.stderr
        line = [l for l in r.stderr.splitlines() if "classified as human" in l][0].split("] ", 1)[1]
        pipe_s = float([l for l in r.stderr.splitlines() if "DEBUG" in l and " s, " in l][0].split("] ", 1)[1].split(" s")[0])
        rows.append({"variant": name, "seconds": round(dt, 2), "gbp_s": round(gbp / dt, 3),
                     "pipeline_seconds": pipe_s, "pipeline_gbp_s": round(gbp / pipe_s, 3),
                     "out_bytes": os.path.getsize(o1) + os.path.getsize(o2), "summary": line})
        print(json.dumps(rows[-1]), file=sys.stderr)
    print(json.dumps({"pairs": args.pairs, "threads": args.threads, "host_cores": os.cpu_count(),
                      "note": "wall clock of the whole CLI process: CUDA context + database load + pipeline", "rows": rows}, indent=1))


if __name__ == "__main__":
    main()
