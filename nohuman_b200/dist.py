"""Multi-GPU plumbing: one process per GPU (torch.distributed), read batches
sharded, database replicated, no collective on the classification path.

The table is loaded (or built) once on rank 0 and broadcast — over NCCL /
NVLink on GPUs, over gloo in the CPU tests — together with the small
opts.k2d / taxo.k2d images; every rank then opens it in place with
nh_db_open_memory(cells_on_device=1).  Per-batch results carry their batch id,
so the writer restores input order (kraken2's ordered output queue,
SURVEY.md A.6); the only reduction is the three counters nohuman logs
(reference src/lib.rs:38-45).
"""
from __future__ import annotations

import numpy as np
import torch
import torch.distributed as dist

TABLE_ALIGN_CELLS = 32  # the resident table is padded to whole 128-byte lines


def padded_cells(capacity: int) -> int:
    return (capacity + TABLE_ALIGN_CELLS - 1) // TABLE_ALIGN_CELLS * TABLE_ALIGN_CELLS


def unit_shard(n_units: int, rank: int, world: int) -> tuple[int, int]:
    """Contiguous, balanced [start, stop) of units for `rank` (sizes differ by at most 1)."""
    base, extra = divmod(n_units, world)
    start = rank * base + min(rank, extra)
    return start, start + base + (1 if rank < extra else 0)


def batch_owner(batch_id: int, world: int) -> int:
    """Batch b of the input stream goes to rank b mod world."""
    return batch_id % world


def broadcast_bytes(data: bytes | None, src: int, device, group=None) -> bytes:
    n = torch.tensor([len(data) if data is not None else 0], dtype=torch.int64, device=device)
    dist.broadcast(n, src, group=group)
    buf = torch.empty(int(n.item()), dtype=torch.uint8, device=device)
    if dist.get_rank(group) == src:
        buf.copy_(torch.frombuffer(bytearray(data), dtype=torch.uint8))
    dist.broadcast(buf, src, group=group)
    return bytes(buf.cpu().numpy().tobytes())


def broadcast_table(cells: torch.Tensor | None, header: list[int] | None, opts: bytes | None,
                    taxo: bytes | None, device, src: int = 0, group=None, chunk_bytes: int = 1 << 30):
    """Replicates a database from `src`.  cells: uint8 view of the padded cell
    array on `src` (any tensor elsewhere is ignored).  Returns
    (cells_uint8_tensor, header[4], opts, taxo) on every rank; on `src` the
    tensor is the one passed in (no copy)."""
    rank = dist.get_rank(group)
    hdr = torch.zeros(4, dtype=torch.int64, device=device)
    if rank == src:
        hdr.copy_(torch.tensor([int(x) for x in header], dtype=torch.int64))
    dist.broadcast(hdr, src, group=group)
    header = [int(x) for x in hdr.cpu().tolist()]
    opts = broadcast_bytes(opts, src, device, group)
    taxo = broadcast_bytes(taxo, src, device, group)
    nbytes = padded_cells(header[0]) * 4
    if rank != src:
        cells = torch.zeros(nbytes, dtype=torch.uint8, device=device)
    assert cells.numel() == nbytes and cells.dtype == torch.uint8
    for lo in range(0, nbytes, chunk_bytes):  # bounded message size; one NVLink broadcast per chunk
        dist.broadcast(cells[lo:lo + chunk_bytes], src, group=group)
    return cells, header, opts, taxo


def reduce_counts(total: int, classified: int, unclassified: int, device, group=None):
    t = torch.tensor([total, classified, unclassified], dtype=torch.int64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
    return tuple(int(x) for x in t.cpu().tolist())


def gather_in_batch_order(local: dict[int, np.ndarray], n_batches: int, world: int, group=None) -> list[np.ndarray]:
    """Every rank holds the results of the batches it owned ({batch_id: array});
    rank 0 gets them back as a list in batch order (what the ordered writer consumes)."""
    objs = [None] * world
    dist.all_gather_object(objs, {int(k): np.asarray(v) for k, v in local.items()}, group=group)
    merged = {}
    for d in objs:
        merged.update(d)
    assert sorted(merged) == list(range(n_batches)), "a batch was lost or duplicated"
    return [merged[b] for b in range(n_batches)]
