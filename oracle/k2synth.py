"""CPU twin of the synthetic-workload generator (oracle/k2_synth.c <-> nohuman_b200/csrc/nh_synth.cu).
BENCH / TEST INFRASTRUCTURE ONLY: `bench.py --impl reference` builds its HPRC.r2-sized table and samples
its reads with this module, on the host cores, without ever loading the CUDA library."""
from __future__ import annotations

import ctypes as C
import os
import tempfile

import numpy as np

from . import k2oracle


class ReadsParams(C.Structure):
    _fields_ = [
        ("seed", C.c_uint64), ("genome_seed", C.c_uint64), ("genome_bases", C.c_uint64),
        ("human_frac", C.c_double), ("sub_rate", C.c_double), ("ins_rate", C.c_double),
        ("del_rate", C.c_double), ("n_rate", C.c_double), ("paired", C.c_int32),
        ("reserved", C.c_int32), ("insert_mean", C.c_double), ("insert_sd", C.c_double),
    ]


_bound = False


def _lib():
    global _bound
    L = k2oracle.lib()
    if not _bound:
        L.k2s_genome.restype = None
        L.k2s_genome.argtypes = [C.c_void_p, C.c_uint64, C.c_uint64, C.c_uint64, C.c_int]
        L.k2s_build_db.restype = C.c_int
        L.k2s_build_db.argtypes = [C.POINTER(k2oracle.Cht), C.POINTER(k2oracle.Taxonomy), C.POINTER(k2oracle.IndexOptions),
                                   C.c_void_p, C.c_int, C.c_double, C.c_uint64, C.c_uint64, C.c_double, C.c_uint64,
                                   C.c_int, C.POINTER(C.c_uint64)]
        L.k2s_reads.restype = None
        L.k2s_reads.argtypes = [C.c_void_p, C.c_void_p, C.c_uint64, C.POINTER(ReadsParams), C.c_int]
        _bound = True
    return L


def synth_genome(genome_seed: int, start: int, n: int, threads: int = 0) -> np.ndarray:
    out = np.empty(n, np.uint8)
    _lib().k2s_genome(out.ctypes.data, start, n, genome_seed, threads)
    return out


def build_synthetic_db(capacity: int, target_load: float = 0.7, genome_seed: int = 0x5EED,
                       block_bases: int = 1 << 20, overlap_frac: float = 0.1, max_genome_bases: int = 0,
                       k: int = 35, l: int = 31, spaces: int = 7, n_super: int = 5, n_hap_per_super: int = 4,
                       threads: int = 0):
    """Same arguments and defaults as nohuman_b200.synth.build_synthetic_db; returns (OracleDb, meta)
    with meta = {opts, taxo, internal, genome_seed, genome_bases, hash_header}."""
    from nohuman_b200 import synth as images  # pure-Python image builders; the CUDA library is not loaded
    nodes, leaves = images.human_pangenome_taxonomy(n_super, n_hap_per_super)
    taxo, internal = images.taxonomy_image(nodes)
    opts_b = images.opts_image(k, l, spaces)
    L = _lib()
    d = tempfile.mkdtemp(prefix="k2synth_")
    with open(os.path.join(d, "opts.k2d"), "wb") as f:
        f.write(opts_b)
    with open(os.path.join(d, "taxo.k2d"), "wb") as f:
        f.write(taxo)
    opts, tax, cht = k2oracle.IndexOptions(), k2oracle.Taxonomy(), k2oracle.Cht()
    assert L.k2o_load_opts(os.path.join(d, "opts.k2d").encode(), C.byref(opts)) == 0
    assert L.k2o_load_taxonomy(os.path.join(d, "taxo.k2d").encode(), C.byref(tax)) == 0
    value_bits = 1
    while (1 << value_bits) < tax.node_count:
        value_bits += 1
    if L.k2o_cht_alloc(C.byref(cht), capacity, value_bits):
        raise MemoryError(f"cannot allocate a table of {capacity} cells")
    leaf_ids = np.array([internal[x] for x in leaves], np.uint32)
    gb = C.c_uint64()
    rc = L.k2s_build_db(C.byref(cht), C.byref(tax), C.byref(opts), leaf_ids.ctypes.data, len(leaf_ids), target_load,
                        genome_seed, block_bases, overlap_frac, max_genome_bases, threads, C.byref(gb))
    if rc:
        raise RuntimeError(f"k2s_build_db failed ({rc})")
    db = k2oracle.OracleDb(opts, tax, cht)
    meta = {"opts": opts_b, "taxo": taxo, "internal": internal, "genome_seed": genome_seed, "genome_bases": int(gb.value),
            "hash_header": [int(cht.capacity), int(cht.size), int(cht.key_bits), int(cht.value_bits)]}
    return db, meta


def synth_reads(offsets: np.ndarray, genome_seed: int, genome_bases: int, seed: int = 1, human_frac: float = 0.5,
                sub_rate: float = 0.005, ins_rate: float = 0.0, del_rate: float = 0.0, n_rate: float = 0.01,
                paired: bool = False, insert_mean: float = 350.0, insert_sd: float = 50.0, threads: int = 0) -> np.ndarray:
    """The bytes nohuman_b200.synth.synth_reads writes on the GPU for the same arguments."""
    offsets = np.ascontiguousarray(offsets, dtype=np.uint64)
    n_seqs = len(offsets) - 1
    bases = np.zeros(int(offsets[-1]) + 64, np.uint8)
    p = ReadsParams(seed, genome_seed, genome_bases, human_frac, sub_rate, ins_rate, del_rate, n_rate, int(paired), 0,
                    insert_mean, insert_sd)
    _lib().k2s_reads(bases.ctypes.data, offsets.ctypes.data, n_seqs, C.byref(p), threads)
    return bases[:int(offsets[-1])]
