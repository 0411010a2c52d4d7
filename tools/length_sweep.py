#!/usr/bin/env python
"""BASELINE.json configs[4]: read-length sweep 100 bp - 50 kb, single-end, fixed bases per batch,
device-resident timing of the whole classification step (plan + fused kernel + deferred scoring).
Long reads are multi-tile units: their scan and probe run in the fused kernel, their scoring in k_score.

    python tools/length_sweep.py [--capacity-log2 31] [--mbases 300] > profiles/rNN_length_sweep.json
"""
import argparse
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--capacity-log2", type=int, default=31)
    ap.add_argument("--mbases", type=int, default=300)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--ont-error", type=float, default=0.05)
    args = ap.parse_args()
    import torch
    from nohuman_b200 import Session, synth
    torch.cuda.set_device(0)
    sdb = synth.build_synthetic_db(1 << args.capacity_log2, device=0)
    rows = []
    for L in (100, 150, 250, 500, 1000, 2000, 5000, 10000, 20000, 50000):
        n = max(64, args.mbases * 1_000_000 // L)
        total = n * L
        d_off = torch.arange(n + 1, dtype=torch.int64, device="cuda") * L
        d_bases = torch.zeros(total + 64, dtype=torch.uint8, device="cuda")
        err = args.ont_error / 3 if L >= 1000 else 0.0
        synth.synth_reads(0, d_bases.data_ptr(), d_off.data_ptr(), n, sdb.genome_seed, sdb.genome_bases, seed=L,
                          human_frac=0.5, sub_rate=err if err else 0.005, ins_rate=err, del_rate=err, n_rate=0.01)
        torch.cuda.synchronize()
        d_call = torch.empty(n, dtype=torch.int32, device="cuda")
        d_keep = torch.empty(n, dtype=torch.uint8, device="cuda")
        with Session(sdb.db, confidence=0.0, keep_human=True, max_batch_bases=total + 4096, max_batch_seqs=n) as sess:
            ext = torch.cuda.ExternalStream(sess.stream)
            for _ in range(3):
                sess.classify_device(d_bases.data_ptr(), d_off.data_ptr(), n, total, d_call.data_ptr(), d_keep.data_ptr())
                st = sess.sync()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(ext)
            for _ in range(args.steps):
                sess.classify_device(d_bases.data_ptr(), d_off.data_ptr(), n, total, d_call.data_ptr(), d_keep.data_ptr())
                st = sess.sync()
            e1.record(ext)
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / args.steps
        rows.append({"read_len": L, "reads": n, "ms_per_step": round(ms, 3), "gbp_s": round(total / ms / 1e6, 2),
                     "reads_s": round(n / ms * 1e3, 1), "lookups": int(st.n_lookups), "tiles": int(st.n_tiles),
                     "stage_ms": {"plan": round(st.ms_plan, 3), "fused": round(st.ms_minimizer, 3), "score_deferred": round(st.ms_score, 3)},
                     "classified_frac": round(st.n_classified / n, 4)})
        print(json.dumps(rows[-1]), file=sys.stderr)
        del d_bases, d_off
    # BASELINE.json configs[2]: ONT-like reads, log-normal lengths tuned to N50 ~ 10 kb (200 bp .. 100 kb), 5 % error, keep-human
    rng = np.random.default_rng(4)
    sigma = 0.8
    lens = np.clip(rng.lognormal(np.log(10_000) - sigma ** 2, sigma, size=200_000), 200, 100_000).astype(np.int64)
    lens = lens[:int(np.searchsorted(np.cumsum(lens), args.mbases * 1_000_000))]
    srt = np.sort(lens)[::-1]
    n50 = int(srt[np.searchsorted(np.cumsum(srt), srt.sum() / 2)])
    offs = np.zeros(len(lens) + 1, np.int64)
    np.cumsum(lens, out=offs[1:])
    n, total = len(lens), int(offs[-1])
    d_off = torch.from_numpy(offs).cuda()
    d_bases = torch.zeros(total + 64, dtype=torch.uint8, device="cuda")
    e3 = args.ont_error / 3
    synth.synth_reads(0, d_bases.data_ptr(), d_off.data_ptr(), n, sdb.genome_seed, sdb.genome_bases, seed=4, human_frac=0.5,
                      sub_rate=e3, ins_rate=e3, del_rate=e3, n_rate=0.01)
    torch.cuda.synchronize()
    d_call = torch.empty(n, dtype=torch.int32, device="cuda")
    d_keep = torch.empty(n, dtype=torch.uint8, device="cuda")
    with Session(sdb.db, confidence=0.0, keep_human=True, max_batch_bases=total + 4096, max_batch_seqs=n) as sess:
        ext = torch.cuda.ExternalStream(sess.stream)
        for _ in range(3):
            sess.classify_device(d_bases.data_ptr(), d_off.data_ptr(), n, total, d_call.data_ptr(), d_keep.data_ptr())
            st = sess.sync()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(ext)
        for _ in range(args.steps):
            sess.classify_device(d_bases.data_ptr(), d_off.data_ptr(), n, total, d_call.data_ptr(), d_keep.data_ptr())
            st = sess.sync()
        e1.record(ext)
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / args.steps
    ont = {"workload": "configs[2] shape: ONT-like, log-normal lengths, 5% error (sub:ins:del 1:1:1), keep-human", "reads": n,
           "bases": total, "n50": n50, "max_len": int(lens.max()), "ms_per_step": round(ms, 3), "gbp_s": round(total / ms / 1e6, 2),
           "reads_s": round(n / ms * 1e3, 1), "lookups": int(st.n_lookups), "kept_frac": round(st.n_kept / n, 4),
           "stage_ms": {"plan": round(st.ms_plan, 3), "fused": round(st.ms_minimizer, 3), "score_deferred": round(st.ms_score, 3)}}
    print(json.dumps(ont), file=sys.stderr)
    print(json.dumps({"ont": ont, "workload": "read-length sweep, single-end, 50%% genome-derived, keep-human, synthetic 2^%d-cell table" % args.capacity_log2,
                      "rows": rows}, indent=1))


if __name__ == "__main__":
    main()
