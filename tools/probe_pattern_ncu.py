#!/usr/bin/env python
"""Runs the roofline helper's patterns once each on the HPRC.r2-sized table, for an ncu launch list:

    ncu --metrics gpu__time_duration.sum,lts__t_requests_srcunit_tex.sum,lts__t_sectors_srcunit_tex.sum,dram__bytes_read.sum \
        --clock-control none -k regex:k_probe_pattern --csv --log-file profiles/rNN_probe_pattern_ncu.csv \
        python tools/probe_pattern_ncu.py
"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    import torch
    from nohuman_b200 import synth
    torch.cuda.set_device(0)
    sdb = synth.build_synthetic_db(1 << int(os.environ.get("CAPACITY_LOG2", "31")), device=0)
    rows = []
    for name, lanes, depth, bps, p, win in (("independent", 1, 2, 8, 0.0, 0), ("chain p=0.40", 1, 2, 8, 0.40, 0),
                                            ("chain p=0.59 (5% error long reads)", 1, 2, 8, 0.59, 0),
                                            ("chain, cp.async lane pairs", 0, 1, 3, 0.40, 0), ("2 lanes", 2, 2, 8, 0.24, 0),
                                            ("4 lanes", 4, 2, 8, 0.12, 0), ("chain, per-SM 64 MiB windows", 1, 1, 8, 0.40, 64 << 20)):
        items, req = sdb.db.probe_pattern(lanes=lanes, p_continue=p, sm_window_bytes=win, items_per_chain=4096 // (bps * depth),
                                          iters=1, depth=depth, blocks_per_sm=bps)
        rows.append({"pattern": name, "lanes": lanes, "depth": depth, "blocks_per_sm": bps, "p": p, "window": win,
                     "lookups_per_s": items, "requests_per_s": req})
    print(json.dumps(rows, indent=1))


if __name__ == "__main__":
    main()
