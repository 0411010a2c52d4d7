#!/usr/bin/env python
"""A/B harness for kernel variants: one process per library build (NH_LIB_PATH), same synthetic
table and batches, device-resident timing of the whole classification step, and a checksum of the
calls so that variants can be compared with each other (the oracle parity lives in tests/ and bench.py).

    NH_LIB_PATH=nohuman_b200/variants/libnh_X.so python tools/kernel_ab.py --tag X [--capacity-log2 31]
"""
import argparse
import hashlib
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--tag", default="default")
    ap.add_argument("--capacity-log2", type=int, default=31)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--mbases", type=int, default=300)
    ap.add_argument("--cases", default="pe150,se100,se250,se500,se1000,se10000")
    args = ap.parse_args()
    import torch
    from nohuman_b200 import Session, synth
    torch.cuda.set_device(0)
    sdb = synth.build_synthetic_db(1 << args.capacity_log2, device=0)
    out = {"tag": args.tag, "lib": os.environ.get("NH_LIB_PATH", "in-tree"), "cases": {}}
    for case in args.cases.split(","):
        paired = case.startswith("pe")
        L = int(case[2:])
        n = max(64, args.mbases * 1_000_000 // L)
        n -= n & 1
        total = n * L
        d_off = torch.arange(n + 1, dtype=torch.int64, device="cuda") * L
        d_bases = torch.zeros(total + 64, dtype=torch.uint8, device="cuda")
        err = 0.05 / 3 if L >= 1000 else 0.0
        synth.synth_reads(0, d_bases.data_ptr(), d_off.data_ptr(), n, sdb.genome_seed, sdb.genome_bases, seed=1000 + L,
                          human_frac=0.5, sub_rate=err if err else 0.005, ins_rate=err, del_rate=err, n_rate=0.01,
                          paired=paired, insert_mean=350.0, insert_sd=50.0)
        torch.cuda.synchronize()
        nu = n // 2 if paired else n
        d_call = torch.empty(nu, dtype=torch.int32, device="cuda")
        d_keep = torch.empty(nu, dtype=torch.uint8, device="cuda")
        with Session(sdb.db, confidence=0.5 if paired else 0.0, paired=paired, keep_human=not paired,
                     max_batch_bases=total + 4096, max_batch_seqs=n) as sess:
            ext = torch.cuda.ExternalStream(sess.stream)
            for _ in range(3):
                sess.classify_device(d_bases.data_ptr(), d_off.data_ptr(), n, total, d_call.data_ptr(), d_keep.data_ptr())
                st = sess.sync()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            fused = 0.0
            e0.record(ext)
            for _ in range(args.steps):
                sess.classify_device(d_bases.data_ptr(), d_off.data_ptr(), n, total, d_call.data_ptr(), d_keep.data_ptr())
                st = sess.sync()
                fused += st.ms_minimizer
            e1.record(ext)
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / args.steps
        h = hashlib.sha1(d_call.cpu().numpy().tobytes() + d_keep.cpu().numpy().tobytes()).hexdigest()[:16]
        out["cases"][case] = {"ms_step": round(ms, 4), "ms_fused": round(fused / args.steps, 4), "gbp_s": round(total / ms / 1e6, 2),
                              "lookups": int(st.n_lookups), "glookups_s": round(st.n_lookups / (fused / args.steps) / 1e6, 2),
                              "classified": int(st.n_classified), "sha": h}
        del d_bases, d_off
    print(json.dumps(out))


if __name__ == "__main__":
    main()
