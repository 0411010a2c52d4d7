"""The N>1 host logic on CPU: two processes over gloo replicate a database
(header, opts, taxo, padded cell array), shard read batches round robin,
'classify' them with the oracle standing in for the GPU, reduce the three
counters and restore batch order — the same nohuman_b200.dist functions
bench.py uses over NCCL."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, db_dir, out_dir):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, HERE)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import ctypes as C
        import synth
        from nohuman_b200 import dist as nhd
        from oracle import k2oracle
        cpu = torch.device("cpu")
        # ---- rank 0 owns the database; everyone gets a replica ----
        cells = header = opts = taxo = None
        if rank == 0:
            raw = np.fromfile(os.path.join(db_dir, "hash.k2d"), dtype=np.uint8)
            header = [int(x) for x in raw[:32].view(np.uint64)]
            padded = np.zeros(nhd.padded_cells(header[0]) * 4, np.uint8)
            padded[:header[0] * 4] = raw[32:]
            cells = torch.from_numpy(padded)
            opts = open(os.path.join(db_dir, "opts.k2d"), "rb").read()
            taxo = open(os.path.join(db_dir, "taxo.k2d"), "rb").read()
        cells, header, opts, taxo = nhd.broadcast_table(cells, header, opts, taxo, cpu, chunk_bytes=1 << 16)
        assert cells.numel() % 128 == 0 and not cells[header[0] * 4:].any()
        rep = os.path.join(out_dir, f"replica{rank}")
        os.makedirs(rep, exist_ok=True)
        with open(os.path.join(rep, "hash.k2d"), "wb") as f:
            f.write(np.array(header, np.uint64).tobytes())
            f.write(cells.numpy()[:header[0] * 4].tobytes())
        open(os.path.join(rep, "opts.k2d"), "wb").write(opts)
        open(os.path.join(rep, "taxo.k2d"), "wb").write(taxo)
        db = k2oracle.OracleDb.load(rep)
        # ---- same seeded input on every rank; batches owned round robin ----
        genomes = synth.cfg1_genomes(seed=1, scale=0.02)
        seqs = synth.illumina_reads(genomes, 1100, 150, seed=77, paired=True)
        n_pairs, per_batch = len(seqs) // 2, 128
        n_batches = (n_pairs + per_batch - 1) // per_batch
        local, tot, cls = {}, 0, 0
        for b in range(n_batches):
            if nhd.batch_owner(b, world) != rank:
                continue
            part = seqs[2 * b * per_batch: 2 * min(n_pairs, (b + 1) * per_batch)]
            bases, offsets = synth.pack(part)
            ext = db.classify_batch(bases, offsets, paired=True)["ext"]
            local[b] = ext
            tot += len(ext)
            cls += int((ext != 0).sum())
        total, classified, unclassified = nhd.reduce_counts(tot, cls, tot - cls, cpu)
        merged = nhd.gather_in_batch_order(local, n_batches, world)
        if rank == 0:
            bases, offsets = synth.pack(seqs)
            want = db.classify_batch(bases, offsets, paired=True)["ext"]
            got = np.concatenate(merged)
            assert np.array_equal(got, want)
            assert (total, classified, unclassified) == (n_pairs, int((want != 0).sum()), int((want == 0).sum()))
            np.save(os.path.join(out_dir, "ok.npy"), got)
        # contiguous sharding helper: exact cover, balanced
        for n in (0, 1, 7, 1000, 1001):
            spans = [nhd.unit_shard(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(spans[i][1] == spans[i + 1][0] for i in range(world - 1))
            assert max(b - a for a, b in spans) - min(b - a for a, b in spans) <= 1
    finally:
        dist.destroy_process_group()


def test_two_ranks_replicate_shard_and_merge(small_db, tmp_path):
    port = _free_port()
    mp.spawn(_worker, args=(2, port, small_db.path, str(tmp_path)), nprocs=2, join=True)
    assert os.path.exists(tmp_path / "ok.npy")
    # both replicas are byte-identical to the source database
    for r in (0, 1):
        for n in ("hash.k2d", "opts.k2d", "taxo.k2d"):
            a = open(os.path.join(small_db.path, n), "rb").read()
            b = open(tmp_path / f"replica{r}" / n, "rb").read()
            assert a == b, (r, n)
