#!/usr/bin/env python
"""k_stream_classify on the same 1 M pairs as ASCII and as packed input (kernel time from the session's CUDA events)."""
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    import torch
    from nohuman_b200 import Session, synth
    from nohuman_b200.api import pack_reads
    torch.cuda.set_device(0)
    cap = 1 << 31
    sdb = synth.build_synthetic_db(cap, device=0)
    n_pairs, L = 1_000_000, 150
    n_seqs = 2 * n_pairs
    d_off = torch.arange(n_seqs + 1, dtype=torch.int64, device="cuda") * L
    d_bases = torch.zeros(n_seqs * L + 64, dtype=torch.uint8, device="cuda")
    synth.synth_reads(0, d_bases.data_ptr(), d_off.data_ptr(), n_seqs, sdb.genome_seed, 2 * cap, seed=5, paired=True, n_rate=0.01)
    torch.cuda.synchronize()
    bases = d_bases[:n_seqs * L].cpu().numpy()
    offsets = d_off.cpu().numpy().astype(np.uint64)
    codes, valid, poff = pack_reads(bases, offsets, os.cpu_count() or 8)
    out = {}
    with Session(sdb.db, confidence=0.5, paired=True, max_batch_bases=n_seqs * L + 4096, max_batch_seqs=n_seqs) as s:
        call = np.zeros(n_pairs, np.uint32)
        keep = np.zeros(n_pairs, np.uint8)
        for name in ("ascii", "packed", "ascii", "packed"):
            ms = []
            for _ in range(6):
                if name == "ascii":
                    st = s.classify_raw(bases.ctypes.data, offsets.ctypes.data, n_seqs, call.ctypes.data, keep.ctypes.data)
                else:
                    st = s.classify_packed_raw(codes.ctypes.data, valid.ctypes.data, poff.ctypes.data, offsets.ctypes.data, n_seqs,
                                               call.ctypes.data, keep.ctypes.data)
                ms.append((st.ms_minimizer, st.ms_h2d))
            out.setdefault(name, []).append({"kernel_ms": round(min(m[0] for m in ms[1:]), 4), "h2d_ms": round(min(m[1] for m in ms[1:]), 3),
                                             "classified": int(st.n_classified), "checksum": int(call.astype(np.uint64).sum())})
    print(json.dumps(out))


if __name__ == "__main__":
    main()
