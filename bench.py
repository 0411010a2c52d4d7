#!/usr/bin/env python
"""bench.py — throughput of the nohuman hot path (kraken2-style classification
+ keep/drop) on B200.

    python bench.py --gpus N --steps K --warmup W            # this repo (CUDA)
    python bench.py --impl reference --gpus N ...            # CPU baseline arm

A step is one pass of the four kernels (plan -> minimizers -> hash probe ->
score/decide) over one batch of synthetic reads of BASELINE.json configs[1]'s
shape: 2x150 bp paired-end Illumina reads, 50 % sampled from the genome the
database was built from ("human-derived"), --conf 0.5, against a synthetic
kraken2-format table sized like HPRC.r2 (default 2^31 cells = 8 GiB, load 0.7).
`value` is measured with the batch already resident in HBM; `e2e` goes through
nh_classify_batch with HOST buffers (H2D + D2H inside the timed region).

The reference arm times the CPU implementation of the same path (the oracle
port of kraken2's classifier; the real kraken2 binary is unavailable offline)
on the GPU box's host cores.  oracle/ is imported ONLY in cpu_baseline() and
the reference arm.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

READ_LEN = 150
CONF = 0.5
METRIC = "classified+filtered throughput (Gbp/s)"
UNIT = "Gbp/s"


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--capacity-log2", type=int, default=31, help="hash table cells = 2^this")
    ap.add_argument("--pairs-per-step", type=int, default=1_000_000, help="read pairs per GPU per step")
    ap.add_argument("--ref-pairs-per-step", type=int, default=0,
                    help="reference arm: pairs per step (0 = sized for ~3 s per step)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    return ap.parse_args()


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed regions.  The
    sampler runs from before warm-up to the end of the bench; `mark()` brackets
    the timed regions and only samples whose timestamp falls inside one count
    (if a region is shorter than the sampling period, the samples taken under
    the same load just before it are used and `window` says so)."""

    FIELDS = ("timestamp,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
              "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
              "clocks_event_reasons.sw_power_cap")
    NAMES = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]

    def __init__(self, index: int):
        self.index = index
        self.proc = None
        self.lines = []
        self.windows = []
        self.t_first = None

    def start(self):
        import datetime
        self.t_first = datetime.datetime.now()
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.FIELDS}",
                 "--format=csv,noheader,nounits", "-lms", "50"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def mark(self):
        """returns a token; call close(token) at the end of the region"""
        import datetime
        self.windows.append([datetime.datetime.now(), None])
        return len(self.windows) - 1

    def close(self, tok):
        import datetime
        self.windows[tok][1] = datetime.datetime.now()

    def stop(self):
        import datetime
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi not sampled on this rank"]}
        time.sleep(0.12)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        rows = []
        for ln in self.lines:
            parts = [x.strip() for x in ln.split(",")]
            if len(parts) < 8:
                continue
            try:
                ts = datetime.datetime.strptime(parts[0], "%Y/%m/%d %H:%M:%S.%f")
                rows.append((ts, float(parts[1]), float(parts[2]), float(parts[3]), parts[4:8]))
            except ValueError:
                continue
        inside = [r for r in rows if any(w0 <= r[0] <= (w1 or r[0]) for w0, w1 in self.windows)]
        window = "timed regions"
        if not inside and self.windows:
            # regions shorter than the sampling period: the warm-up just before runs the same kernels
            w0 = self.windows[0][0] - datetime.timedelta(seconds=1.0)
            w1 = max(w[1] or w[0] for w in self.windows) + datetime.timedelta(seconds=0.1)
            inside = [r for r in rows if w0 <= r[0] <= w1]
            window = "timed regions + 1 s of warm-up before them"
        reasons = set()
        for r in inside:
            for nm, v in zip(self.NAMES, r[4]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        sm = [r[1] for r in inside]
        return {"sm_mhz": float(np.median(sm)) if sm else None,
                "sm_max_mhz": inside[0][2] if inside else (rows[0][2] if rows else None),
                "power_w_max": max((r[3] for r in inside), default=None),
                "reasons": sorted(reasons), "samples": len(sm), "samples_total": len(rows),
                "window": window}


class DevPtr:
    """Zero-copy torch view of a raw device pointer (for NCCL broadcast of the table)."""

    def __init__(self, ptr: int, nbytes: int):
        self.__cuda_array_interface__ = {"shape": (nbytes,), "typestr": "|u1", "data": (ptr, False),
                                         "version": 3, "strides": None}


def make_batch(torch, synth, device, sdb_meta, n_pairs, seed):
    n_seqs = 2 * n_pairs
    total = n_seqs * READ_LEN
    d_off = (torch.arange(n_seqs + 1, dtype=torch.int64, device="cuda") * READ_LEN)
    d_bases = torch.zeros(total + 64, dtype=torch.uint8, device="cuda")
    synth.synth_reads(device, d_bases.data_ptr(), d_off.data_ptr(), n_seqs, sdb_meta["genome_seed"],
                      sdb_meta["genome_bases"], seed=seed, human_frac=0.5, sub_rate=0.005,
                      n_rate=0.01, paired=True, insert_mean=350.0, insert_sd=50.0,
                      stream=torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    return d_bases, d_off, n_seqs, total


def cpu_baseline(cells, sdb_meta, opts_b, taxo_b, hdr, bases, offsets, target_s=12.0, threads=0):
    """The oracle port of kraken2's classifier on the host cores (bounded sample)."""
    from oracle import k2oracle  # the only place bench.py touches oracle/
    import ctypes as C
    import tempfile
    d = tempfile.mkdtemp(prefix="nh_bench_db_")
    with open(os.path.join(d, "opts.k2d"), "wb") as f:
        f.write(opts_b)
    with open(os.path.join(d, "taxo.k2d"), "wb") as f:
        f.write(taxo_b)
    L = k2oracle.lib()
    opts, tax = k2oracle.IndexOptions(), k2oracle.Taxonomy()
    assert L.k2o_load_opts(os.path.join(d, "opts.k2d").encode(), C.byref(opts)) == 0
    assert L.k2o_load_taxonomy(os.path.join(d, "taxo.k2d").encode(), C.byref(tax)) == 0
    odb = k2oracle.OracleDb.from_arrays(opts, tax, cells, hdr[0], hdr[1], hdr[3])
    odb.confidence = CONF
    cores = threads or os.cpu_count() or 1
    n_pairs_all = (len(offsets) - 1) // 2

    def run(n_pairs):
        o = offsets[:2 * n_pairs + 1]
        t0 = time.perf_counter()
        r = odb.classify_batch(bases[:int(o[-1])], o, paired=True, threads=cores)
        return time.perf_counter() - t0, r

    probe_n = min(n_pairs_all, 20000)
    t, _ = run(probe_n)
    rate = probe_n / max(t, 1e-6)
    n = int(min(n_pairs_all, max(probe_n, rate * target_s)))
    t, r = run(n)
    passes = 1
    if n == n_pairs_all and t < 0.6 * target_s:
        # the whole batch is faster than the sample budget: repeat it
        extra = int(min(30, max(1, round(target_s / max(t, 1e-3)) - 1)))
        for _ in range(extra):
            t += run(n)[0]
        passes += extra
    return odb, {"pairs": n, "passes": passes, "seconds": t, "gbp_s": passes * n * 2 * READ_LEN / t / 1e9,
                 "reads_s": passes * 2 * n / t, "cores": cores, "result": r}


def main():
    args = parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference" and rank != 0:
        return 0

    import torch
    import torch.distributed as dist
    from nohuman_b200 import Database, Session, synth

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (no CPU fallback exists)")
    torch.cuda.set_device(local_rank)
    dev = local_rank
    use_dist = world > 1 and args.impl == "ours"
    if use_dist:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    # ---------------- database: built on rank 0, replicated over NCCL/NVLink ----------------
    from nohuman_b200 import dist as nhd
    capacity = 1 << args.capacity_log2
    t_build0 = time.perf_counter()
    sdb = None
    cells = hdr = opts_b = taxo_b = None
    gmeta = torch.zeros(1, dtype=torch.int64, device="cuda")
    if rank == 0:
        sdb = synth.build_synthetic_db(capacity, device=dev)
        opts_b, taxo_b = sdb.opts, sdb.taxo
        hdr = sdb.hash_header()
        gmeta[0] = sdb.genome_bases
        db = sdb.db
    keep_alive = None
    if use_dist:
        dist.broadcast(gmeta, 0)
        if rank == 0:
            cells = torch.as_tensor(DevPtr(sdb.db.device_cells_ptr(), nhd.padded_cells(capacity) * 4), device="cuda")
        cells, hdr, opts_b, taxo_b = nhd.broadcast_table(cells, hdr, opts_b, taxo_b, torch.device("cuda", dev))
        torch.cuda.synchronize()
        if rank != 0:
            keep_alive = cells
            db = Database.from_memory(opts_b, taxo_b, hdr, cells.data_ptr(), device=dev, cells_on_device=True)
    sdb_meta = {"genome_seed": 0x5EED, "genome_bases": int(gmeta.item())}
    t_build = time.perf_counter() - t_build0

    n_pairs = args.pairs_per_step
    n_bufs = 2
    bufs = [make_batch(torch, synth, dev, sdb_meta, n_pairs, seed=1000 * (rank + 1) + b)
            for b in range(n_bufs)]
    n_seqs, total = bufs[0][2], bufs[0][3]

    if args.impl == "reference":
        return reference_arm(args, torch, sdb, sdb_meta, opts_b, taxo_b, hdr, bufs)

    # ---------------- device-resident timing (`value`) ----------------
    sess = Session(db, confidence=CONF, paired=True, max_batch_bases=total + 4096,
                   max_batch_seqs=n_seqs)
    d_call = torch.empty(n_pairs, dtype=torch.int32, device="cuda")
    d_keep = torch.empty(n_pairs, dtype=torch.uint8, device="cuda")
    ext = torch.cuda.ExternalStream(sess.stream)

    def step(i):
        b = bufs[i % n_bufs]
        sess.classify_device(b[0].data_ptr(), b[1].data_ptr(), n_seqs, total, d_call.data_ptr(),
                             d_keep.data_ptr())
        return sess.sync()

    sampler = ClockSampler(dev)
    if rank == 0:  # one nvidia-smi poller per run is enough; rank 0's GPU is the one reported
        sampler.start()
    for i in range(args.warmup):
        st = step(i)
    random_gbs = db.random_gather_gbs(1 << 27, 3) if rank == 0 else None

    def barrier():
        if use_dist:
            dist.barrier()
        torch.cuda.synchronize()

    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    stage = {"plan": 0.0, "minimizer": 0.0, "probe": 0.0, "score": 0.0}
    lookups = tiles = launches = classified = 0
    fused = False
    fused_form = 0
    barrier()
    tok = sampler.mark()
    torch.cuda.cudart().cudaProfilerStart()  # `ncu --profile-from-start off` sees the timed steps only
    ev0.record(ext)
    for i in range(args.steps):
        st = step(i)
        stage["plan"] += st.ms_plan
        stage["minimizer"] += st.ms_minimizer
        stage["probe"] += st.ms_probe
        stage["score"] += st.ms_score
        lookups += st.n_lookups
        tiles += st.n_tiles
        launches += st.gpu_launches
        classified += st.n_classified
        fused = bool(st.fused_kernel)
        fused_form = int(st.fused_kernel)
    ev1.record(ext)
    barrier()
    torch.cuda.cudart().cudaProfilerStop()
    sampler.close(tok)
    ms_total = ev0.elapsed_time(ev1)
    if use_dist:
        t = torch.tensor([ms_total], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_total = float(t.item())
    ms_step = ms_total / args.steps
    gbp_s = world * n_pairs * 2 * READ_LEN * args.steps / (ms_total * 1e-3) / 1e9
    reads_s = world * n_seqs * args.steps / (ms_total * 1e-3)

    # ---------------- e2e through the C ABI with host buffers ----------------
    e2e = None
    if not args.no_e2e:
        e2e = run_e2e(args, torch, dist if use_dist else None, db, bufs, n_pairs, n_seqs, total, world,
                      sampler)
    clocks = sampler.stop()

    if rank != 0:
        if use_dist:
            dist.destroy_process_group()
        return 0

    # ---------------- roofline of the dominant kernel + probe ----------------
    peak, peak_src = peaks()
    for kname in stage:
        stage[kname] /= args.steps
    lk_per_step = lookups / args.steps
    staging = lk_per_step * 13.0  # 8 B key + 1 B run length + 4 B taxon per lookup (L2-resident scratch)
    if fused:
        # one kernel: 1 B/base read + one 32 B sector per lookup + 5 B/unit of results
        fname = "stream_classify" if fused_form == 2 else "scan_probe_score"
        stage = {"plan": stage["plan"], fname: stage["minimizer"], "score_deferred": stage["score"]}
        alg = {fname: total * 1.0 + lk_per_step * 32.0 + n_pairs * 5.0, "score_deferred": 0.0, "plan": n_seqs * 8.0}
    else:
        alg = {
            # minimizer: 1 B/base read + 9 B per lookup written (8 B key + 1 B k-mer count) + 8 B/tile
            "minimizer": total * 1.0 + lk_per_step * 9.0 + (tiles / args.steps) * 16.0,
            # probe: one 32 B sector per lookup
            "probe": lk_per_step * 32.0,
            "score": lk_per_step * 5.0 + n_pairs * 5.0,
            "plan": n_seqs * 8.0,
        }
    dom = max(stage, key=lambda k_: stage[k_])
    probe_name = ("stream_classify" if fused_form == 2 else "scan_probe_score") if fused else "probe"

    traffic_tab = {}
    tpath = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tpath):
        with open(tpath) as f:
            tj = json.load(f)
        w = tj.get("workload", {})
        if (w.get("pairs_per_step_per_gpu") == n_pairs and w.get("table_cells") == capacity
                and w.get("read_len") == READ_LEN):
            traffic_tab = tj

    def roof(kname):
        ach = alg[kname] / (stage[kname] * 1e-3) / 1e9 if stage[kname] > 0 else 0.0
        tr = traffic_tab.get("k_" + kname, {}).get("dram_bytes_per_launch")
        return {"kernel": "k_" + kname, "bound": "hbm", "achieved": round(ach, 1), "peak": peak,
                "unit": "GB/s", "frac": round(ach / peak, 4), "traffic": tr, "peak_source": peak_src,
                "ms_per_launch": round(stage[kname], 4),
                "algorithmic_bytes_per_launch": int(alg[kname])}

    roofline = roof(dom)
    # the hash probe against the random-access rate measured on this box in this run
    probe_gbs = lk_per_step * 32.0 / (stage[probe_name] * 1e-3) / 1e9
    roofline_probe = {"kernel": "k_" + probe_name, "lookups_per_s": round(lk_per_step / (stage[probe_name] * 1e-3), 1),
                      "achieved_sector_gbs": round(probe_gbs, 1),
                      "note": "32 B per lookup over the kernel that holds the probe"
                              + (" (fused with the minimizer scan and scoring)" if fused else "")}
    if random_gbs:
        roofline_probe["random_sector_peak_gbs"] = round(random_gbs, 1)
        roofline_probe["frac_of_random_sector_peak"] = round(probe_gbs / random_gbs, 4)

    out = {
        "metric": METRIC, "value": round(gbp_s, 3), "unit": UNIT, "n_gpus": world,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": round(ms_step, 4),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u64",
        "data": "synthetic",
        "config": {
            "workload": "BASELINE configs[1] shape: 2x150 bp paired-end, 50% genome-derived, "
                        "--conf 0.5, synthetic HPRC.r2-sized kraken2 table",
            "pairs_per_step_per_gpu": n_pairs, "read_len": READ_LEN, "confidence": CONF,
            "table_cells": capacity, "table_gib": round(capacity * 4 / 2**30, 2),
            "table_load": round(hdr[1] / capacity, 4), "k": 35, "l": 31,
            "genome_bases": sdb_meta["genome_bases"], "db_build_s": round(t_build, 2),
            "l2_policy": "inputs larger than L2 (batch %.0f MB + random probes over the table)" % (total / 1e6),
            "parallelism": f"dp{world} (read batches sharded, table replicated via NCCL broadcast)",
        },
        "reads_per_s": round(reads_s, 1),
        "stage_ms": {k_: round(v, 4) for k_, v in stage.items()},
        "lookups_per_step": int(lk_per_step), "kernel_path": {0: "warp-per-tile", 1: "fused (phased)", 2: "fused (streaming)"}[fused_form],
        "classified_frac": round(classified / (args.steps * n_pairs), 4),
        "gpu_launches": int(launches),
        "clocks": clocks,
        "roofline": roofline,
        "roofline_probe": roofline_probe,
    }
    if e2e:
        out["e2e"] = e2e

    if world == 1 and not args.no_cpu_baseline:
        cells = sdb.download_cells()
        b = bufs[0]
        h_bases = b[0][:total].cpu().numpy()
        h_off = b[1].cpu().numpy().astype(np.uint64)
        odb, cb = cpu_baseline(cells, sdb_meta, opts_b, taxo_b, hdr, h_bases, h_off)
        # parity spot check of the timed configuration against the oracle
        n = cb["pairs"]
        got = d_call.cpu().numpy().astype(np.uint32)
        sess.classify_device(b[0].data_ptr(), b[1].data_ptr(), n_seqs, total, d_call.data_ptr(),
                             d_keep.data_ptr())
        sess.sync()
        got = d_call.cpu().numpy().astype(np.uint32)[:n]
        out["parity_vs_oracle"] = {"pairs_checked": n,
                                   "mismatches": int((got != cb["result"]["ext"][:n]).sum())}
        out["cpu_baseline"] = {
            "value": round(cb["gbp_s"], 5), "unit": UNIT, "cores": cb["cores"], "kind": "port",
            "sample": f"first {n} pairs of the step's batch x {cb['passes']} passes, {cb['seconds']:.1f} s, "
                      "kraken2 restatement (upstream binary unavailable offline), OpenMP",
            "reads_per_s": round(cb["reads_s"], 1),
            "oracle_lookups_per_pair": round(cb["result"]["lookups"] / n, 2),
            "oracle_sectors_per_lookup": round(cb["result"]["sectors"] / max(1, cb["result"]["lookups"]), 3),
        }
    print(json.dumps(out))
    if use_dist:
        dist.destroy_process_group()
    return 0


def run_e2e(args, torch, dist, db, bufs, n_pairs, n_seqs, total, world, sampler=None):
    """Same metric through nh_classify_batch with pinned HOST buffers: every step
    copies its bases + offsets H2D and its calls + keep mask D2H.  Two sessions
    on two host threads overlap one step's copies with the other's kernels."""
    from nohuman_b200 import Session
    n_workers = 2
    sessions = [Session(db, confidence=CONF, paired=True, max_batch_bases=total + 4096,
                        max_batch_seqs=n_seqs) for _ in range(n_workers)]
    host = []
    for b in bufs[:n_workers]:
        hb = torch.empty(total, dtype=torch.uint8).pin_memory()
        hb.copy_(b[0][:total])
        ho = torch.empty(n_seqs + 1, dtype=torch.int64).pin_memory()
        ho.copy_(b[1])
        hc = torch.empty(n_pairs, dtype=torch.int32).pin_memory()
        hk = torch.empty(n_pairs, dtype=torch.uint8).pin_memory()
        host.append((hb, ho, hc, hk))
    torch.cuda.synchronize()
    steps = max(args.steps, n_workers)

    def worker(w, n):
        hb, ho, hc, hk = host[w]
        for _ in range(n):
            sessions[w].classify_raw(hb.data_ptr(), ho.data_ptr(), n_seqs, hc.data_ptr(), hk.data_ptr())

    def run(n_total):
        ths = [threading.Thread(target=worker, args=(w, n_total // n_workers + (1 if w < n_total % n_workers else 0)))
               for w in range(n_workers)]
        for t in ths:
            t.start()
        for t in ths:
            t.join()

    run(max(args.warmup, n_workers))
    if dist:
        dist.barrier()
    torch.cuda.synchronize()
    tok = sampler.mark() if sampler else None
    t0 = time.perf_counter()
    run(steps)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    if sampler:
        sampler.close(tok)
    if dist:
        t = torch.tensor([dt], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dt = float(t.item())
    for s in sessions:
        s.close()
    return {"value": round(world * steps * n_pairs * 2 * READ_LEN / dt / 1e9, 3), "unit": UNIT,
            "h2d_bytes_per_step": int(total + (n_seqs + 1) * 8), "d2h_bytes_per_step": int(n_pairs * 5),
            "ms_per_step": round(dt / steps * 1e3, 4), "steps": steps,
            "api": "nh_classify_batch (host buffers, pinned), 2 sessions on 2 host threads"}


def reference_arm(args, torch, sdb, sdb_meta, opts_b, taxo_b, hdr, bufs):
    """CPU implementation of the path on the host cores: the oracle port of
    kraken2's classifier (kraken2 itself is not installable offline)."""
    cells = sdb.download_cells()
    b = bufs[0]
    total = b[3]
    h_bases = b[0][:total].cpu().numpy()
    h_off = b[1].cpu().numpy().astype(np.uint64)
    sdb.db.close()
    odb, probe = cpu_baseline(cells, sdb_meta, opts_b, taxo_b, hdr, h_bases, h_off, target_s=3.0)
    n = args.ref_pairs_per_step or probe["pairs"]
    n = min(n, (len(h_off) - 1) // 2)
    o = h_off[:2 * n + 1]
    bb = h_bases[:int(o[-1])]
    cores = probe["cores"]
    for _ in range(max(1, min(args.warmup, 2))):
        odb.classify_batch(bb, o, paired=True, threads=cores)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        odb.classify_batch(bb, o, paired=True, threads=cores)
    dt = time.perf_counter() - t0
    gbp_s = args.steps * n * 2 * READ_LEN / dt / 1e9
    out = {
        "impl": "reference", "metric": METRIC, "value": round(gbp_s, 5), "unit": UNIT,
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": round(dt / args.steps * 1e3, 3), "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "u64", "data": "synthetic",
        "config": {
            "workload": "BASELINE configs[1] shape: 2x150 bp paired-end, 50% genome-derived, "
                        "--conf 0.5, synthetic HPRC.r2-sized kraken2 table",
            "pairs_per_step": n, "read_len": READ_LEN, "confidence": CONF,
            "table_cells": hdr[0], "table_load": round(hdr[1] / hdr[0], 4),
        },
        "reads_per_s": round(args.steps * 2 * n / dt, 1),
        "cpu_baseline": {"value": round(gbp_s, 5), "unit": UNIT, "cores": cores, "kind": "port",
                         "sample": f"{n} pairs per step x {args.steps} steps; kraken2 restatement "
                                   "(upstream binary unavailable offline), OpenMP on all host cores"},
        "e2e": {"value": round(gbp_s, 5), "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(out))
    return 0


if __name__ == "__main__":
    sys.exit(main())
