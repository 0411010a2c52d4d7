#!/usr/bin/env python
"""Instruction and stall budget of k_stream_classify per PHASE, from an ncu source-level export.

    ncu -i X.ncu-rep --page source --csv --print-source cuda,sass > src.csv
    python tools/phase_budget.py src.csv [nohuman_b200/csrc/nh_kernels.cu]

The phases are found by the marker comments in nh_kernels.cu (so the table follows the code), inlined
helpers from other files are attributed by name.  Prints a markdown table: share of executed warp
instructions, share of stall samples, top stall reasons."""
import csv
import re
import sys

src_csv = sys.argv[1]
cu = sys.argv[2] if len(sys.argv) > 2 else "nohuman_b200/csrc/nh_kernels.cu"
lines = open(cu).read().split("\n")


def find(pat, start=0):
    for i in range(start, len(lines)):
        if pat in lines[i]:
            return i + 1
    raise SystemExit(f"marker not found: {pat}")


k0 = find("k_stream_classify(const NhDbParams db")
marks = [
    ("kernel prologue, group hand-out, table reset", k0),
    ("probe round 1: wait for the sectors, scan 8 cells, continuation queue", find("/* ---- 1. look at the sectors", k0)),
    ("probe round 1b: fold the hit into the unit's taxon table (shared-memory atomics)", find("if (active && done) {", k0)),
    ("probe round 2: pop 32 lookups, fmix64 + hash % capacity, request the sectors (cp.async)", find("/* ---- 2. next 32 lookups", k0)),
    ("scan setup (tile geometry)", find("/* ---------------- scan, feeding the probe", k0)),
    ("emit: ballot / popc compaction of closed runs into the queue", find("auto emit = [&]", k0)),
    ("scan: 2-bit codes, rolling l-mers, canonical, window minimum, run logic (per base)", find("auto scan_quad = [&]", k0)),
    ("base staging (cp.async chunks, mbarrier) and word loop", find("/* the lane's window for chunk c", k0)),
    ("drain, tile borders, deferred-tile hand-off", find("const bool has_runs", k0)),
    ("in-warp ResolveTree + keep/drop (leader lanes)", find("/* ---------------- score the units that live in this warp", k0)),
    ("epilogue", find("tot_classified = warp_sum_u32(tot_classified);", k0)),
]
kend = find("/* per-read output support", k0)
helper_phase = {
    "nh_fmix64": 3, "nh_fastmod": 3, "nh_mulhi64": 3, "__umul64hi": 3, "min_u64": 6, "nh_pack4": 6,
    "atomicCAS": 2, "atomicAdd": 2, "atomicOr": 2, "__uAtomic": 2, "mbar_wait": None, "cp_async": None, "lds_u32": 7,
    "__ballot_sync": None, "__popc": None, "__shfl": 3, "__nvvm_vote": None, "__nvvm_bar_warp": None, "lane_tab_get": 9, "lca": 9,
    "is_a_ancestor_of_b": 9, "unit_total_kmers": 9, "warp_sum": 10,
}

rows = list(csv.reader(open(src_csv)))
agg = {}
cur_file = None
hdr = None
for r in rows:
    if r and r[0] == "File Path":
        cur_file = r[1]
        continue
    if len(r) > 5 and r[0] == "Line No":
        hdr = {h: i for i, h in enumerate(r)}
        ncol = len(r)
        continue
    if not hdr or len(r) != ncol or not r[0].strip().isdigit():
        continue
    ln = int(r[0])
    text = r[1]
    inst = int(r[hdr["Instructions Executed"]] or 0)
    smp = int(r[hdr["# Samples"]] or 0)
    stalls = {h[6:]: int(r[i] or 0) for h, i in hdr.items() if h.startswith("stall_") and "Not Issued" not in h}
    phase = None
    if cur_file and cur_file.endswith("nh_kernels.cu") and k0 <= ln < kend:
        for idx, (_, start) in enumerate(marks):
            if ln >= start:
                phase = idx
    if phase is None:
        for name, ph in helper_phase.items():
            if name in text:
                phase = ph
                break
    key = marks[phase][0] if phase is not None else "other inlined helpers (votes, waits, copies: attributed by the compiler to library headers)"
    a = agg.setdefault(key, {"inst": 0, "smp": 0, "stalls": {}})
    a["inst"] += inst
    a["smp"] += smp
    for k_, v in stalls.items():
        a["stalls"][k_] = a["stalls"].get(k_, 0) + v
ti = sum(a["inst"] for a in agg.values())
ts = sum(a["smp"] for a in agg.values())
print(f"total: {ti / 1e9:.3f} G warp instructions, {ts} stall samples\n")
print("| phase | warp instructions | share | stall samples | top stall reasons |")
print("|---|---|---|---|---|")
order = [m[0] for m in marks] + [k_ for k_ in agg if k_ not in [m[0] for m in marks]]
for k_ in order:
    if k_ not in agg:
        continue
    a = agg[k_]
    top = sorted(a["stalls"].items(), key=lambda x: -x[1])[:3]
    print(f"| {k_} | {a['inst'] / 1e6:.0f} M | {100 * a['inst'] / ti:.1f} % | {100 * a['smp'] / ts:.1f} % | "
          + ", ".join(f"{n} {100 * v / max(1, a['smp']):.0f} %" for n, v in top if v) + " |")
