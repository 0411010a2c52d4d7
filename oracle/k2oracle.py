"""ctypes face of the CPU oracle (oracle/k2_oracle.c).  TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
--impl reference legs may import this module.  The product package
(nohuman_b200/) never does.  PARITY UNPINNED: see k2_oracle.h.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from dataclasses import dataclass

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "libk2oracle.so")


def build(force: bool = False) -> str:
    """Compile the oracle with the committed Makefile (gcc, seconds)."""
    deps = [os.path.join(_HERE, f) for f in ("k2_oracle.c", "k2_synth.c", "k2_oracle.h", "Makefile")]
    stale = (not os.path.exists(_LIB_PATH)) or any(
        os.path.getmtime(p) > os.path.getmtime(_LIB_PATH) for p in deps
    )
    if force or stale:
        subprocess.check_call(["make", "-C", _HERE, "-s", "-B"])
    return _LIB_PATH


class IndexOptions(C.Structure):
    _fields_ = [
        ("k", C.c_uint64),
        ("l", C.c_uint64),
        ("spaced_seed_mask", C.c_uint64),
        ("toggle_mask", C.c_uint64),
        ("dna_db", C.c_uint8),
        ("pad0", C.c_uint8 * 7),
        ("minimum_acceptable_hash_value", C.c_uint64),
        ("revcom_version", C.c_int32),
        ("db_version", C.c_int32),
        ("db_type", C.c_int32),
        ("pad1", C.c_int32),
    ]


class TaxonomyNode(C.Structure):
    _fields_ = [
        ("parent_id", C.c_uint64),
        ("first_child", C.c_uint64),
        ("child_count", C.c_uint64),
        ("name_offset", C.c_uint64),
        ("rank_offset", C.c_uint64),
        ("external_id", C.c_uint64),
        ("godparent_id", C.c_uint64),
    ]


class Taxonomy(C.Structure):
    _fields_ = [
        ("node_count", C.c_uint64),
        ("name_data_len", C.c_uint64),
        ("rank_data_len", C.c_uint64),
        ("nodes", C.POINTER(TaxonomyNode)),
        ("name_data", C.c_void_p),
        ("rank_data", C.c_void_p),
    ]


class Cht(C.Structure):
    _fields_ = [
        ("capacity", C.c_uint64),
        ("size", C.c_uint64),
        ("key_bits", C.c_uint64),
        ("value_bits", C.c_uint64),
        ("cells", C.POINTER(C.c_uint32)),
        ("owns_cells", C.c_int),
    ]


class ReadResult(C.Structure):
    _fields_ = [
        ("call", C.c_uint64),
        ("ext_call", C.c_uint64),
        ("total_kmers", C.c_uint64),
        ("hit_groups", C.c_int64),
        ("lookups", C.c_uint64),
        ("cells", C.c_uint64),
        ("sectors", C.c_uint64),
    ]


class HitCounts(C.Structure):
    _fields_ = [
        ("taxon", C.POINTER(C.c_uint64)),
        ("count", C.POINTER(C.c_uint32)),
        ("n", C.c_size_t),
        ("cap", C.c_size_t),
    ]


class Db(C.Structure):
    _fields_ = [
        ("opts", C.POINTER(IndexOptions)),
        ("cht", C.POINTER(Cht)),
        ("tax", C.POINTER(Taxonomy)),
        ("confidence", C.c_double),
        ("minimum_hit_groups", C.c_int64),
    ]


class Scanner(C.Structure):
    _fields_ = [("_opaque", C.c_uint8 * 256)]


_lib = None


def lib():
    global _lib
    if _lib is not None:
        return _lib
    build()
    L = C.CDLL(_LIB_PATH)
    u64, i32 = C.c_uint64, C.c_int
    L.k2o_fmix64.restype = u64
    L.k2o_fmix64.argtypes = [u64]
    L.k2o_reverse_complement.restype = u64
    L.k2o_reverse_complement.argtypes = [u64, i32, i32]
    L.k2o_canonical.restype = u64
    L.k2o_canonical.argtypes = [u64, i32, i32]
    L.k2o_spaced_seed_mask.restype = u64
    L.k2o_spaced_seed_mask.argtypes = [i32, i32]
    L.k2o_scan_positions.restype = C.c_size_t
    L.k2o_scan_positions.argtypes = [C.POINTER(IndexOptions), C.c_char_p, C.c_size_t,
                                     C.c_void_p, C.c_void_p, C.c_size_t]
    L.k2o_cht_get.restype = C.c_uint32
    L.k2o_cht_get.argtypes = [C.POINTER(Cht), u64]
    L.k2o_cht_get_stats.restype = C.c_uint32
    L.k2o_cht_get_stats.argtypes = [C.POINTER(Cht), u64, C.POINTER(u64), C.POINTER(u64)]
    L.k2o_cht_insert_lca.restype = i32
    L.k2o_cht_insert_lca.argtypes = [C.POINTER(Cht), C.POINTER(Taxonomy), u64, C.c_uint32]
    L.k2o_cht_alloc.restype = i32
    L.k2o_cht_alloc.argtypes = [C.POINTER(Cht), u64, u64]
    L.k2o_cht_free.argtypes = [C.POINTER(Cht)]
    L.k2o_is_a_ancestor_of_b.restype = i32
    L.k2o_is_a_ancestor_of_b.argtypes = [C.POINTER(Taxonomy), u64, u64]
    L.k2o_lca.restype = u64
    L.k2o_lca.argtypes = [C.POINTER(Taxonomy), u64, u64]
    L.k2o_taxonomy_build.restype = i32
    L.k2o_taxonomy_build.argtypes = [C.POINTER(Taxonomy), C.c_size_t, C.POINTER(u64),
                                     C.POINTER(u64), C.POINTER(C.c_char_p), C.POINTER(C.c_char_p)]
    L.k2o_taxonomy_internal_id.restype = u64
    L.k2o_taxonomy_internal_id.argtypes = [C.POINTER(Taxonomy), u64]
    L.k2o_taxonomy_free.argtypes = [C.POINTER(Taxonomy)]
    for name, typ in (("opts", IndexOptions), ("taxonomy", Taxonomy), ("cht", Cht)):
        f = getattr(L, f"k2o_load_{name}")
        f.restype = i32
        f.argtypes = [C.c_char_p, C.POINTER(typ)]
        f = getattr(L, f"k2o_save_{name}")
        f.restype = i32
        f.argtypes = [C.c_char_p, C.POINTER(typ)]
    L.k2o_resolve_tree.restype = u64
    L.k2o_resolve_tree.argtypes = [C.POINTER(HitCounts), C.POINTER(Taxonomy), u64, C.c_double]
    L.k2o_scanner_init.restype = i32
    L.k2o_scanner_init.argtypes = [C.POINTER(Scanner), C.c_int64, C.c_int64, u64, i32, u64, i32]
    L.k2o_scanner_free.argtypes = [C.POINTER(Scanner)]
    L.k2o_classify_sequence.restype = None
    L.k2o_classify_sequence.argtypes = [C.POINTER(Db), C.POINTER(Scanner), C.POINTER(HitCounts),
                                        C.c_char_p, C.c_size_t, C.c_char_p, C.c_size_t, i32,
                                        C.POINTER(ReadResult), C.POINTER(C.POINTER(u64)),
                                        C.POINTER(C.c_size_t)]
    L.k2o_classify_batch.restype = i32
    L.k2o_classify_batch.argtypes = [C.POINTER(Db), C.c_void_p, C.c_void_p, u64, i32, i32,
                                     C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
    L.k2o_hitlist_string.restype = C.c_void_p
    L.k2o_hitlist_string.argtypes = [C.POINTER(Taxonomy), C.POINTER(u64), C.c_size_t]
    L.k2o_build_add_sequence.restype = i32
    L.k2o_build_add_sequence.argtypes = [C.POINTER(Cht), C.POINTER(Taxonomy),
                                         C.POINTER(IndexOptions), C.c_char_p, C.c_size_t,
                                         C.c_uint32]
    _lib = L
    return L


_libc = C.CDLL(None)
_libc.free.argtypes = [C.c_void_p]


def default_options(k: int = 35, l: int = 31, spaces: int = 7,
                    toggle: int = 0xE37E28C4271B5A2D, min_hash: int = 0) -> IndexOptions:
    """IndexOptions as kraken2-build writes them for a nucleotide DB."""
    o = IndexOptions()
    o.k, o.l = k, l
    o.spaced_seed_mask = lib().k2o_spaced_seed_mask(l, spaces) if spaces else 0
    o.toggle_mask = toggle
    o.dna_db = 1
    o.minimum_acceptable_hash_value = min_hash
    o.revcom_version = 1
    return o


@dataclass
class TaxSpec:
    ext_id: int
    parent_ext_id: int
    name: str = ""
    rank: str = ""


class OracleDb:
    """A kraken2 database held by the oracle: options + taxonomy + hash table."""

    def __init__(self, opts: IndexOptions, tax: Taxonomy, cht: Cht,
                 confidence: float = 0.0, min_hit_groups: int = 2):
        self.opts, self.tax, self.cht = opts, tax, cht
        self.confidence = float(confidence)
        self.min_hit_groups = int(min_hit_groups)

    # -- construction ------------------------------------------------
    @staticmethod
    def build(genomes: list[tuple[int, bytes]], taxonomy: list[TaxSpec],
              opts: IndexOptions | None = None, load_factor: float = 0.7,
              capacity: int | None = None) -> "OracleDb":
        """kraken2-build restated: genomes = [(ext_taxid, sequence)]."""
        L = lib()
        opts = opts or default_options()
        tax = Taxonomy()
        n = len(taxonomy)
        ext = (C.c_uint64 * n)(*[t.ext_id for t in taxonomy])
        par = (C.c_uint64 * n)(*[t.parent_ext_id for t in taxonomy])
        names = (C.c_char_p * n)(*[t.name.encode() for t in taxonomy])
        ranks = (C.c_char_p * n)(*[t.rank.encode() for t in taxonomy])
        if L.k2o_taxonomy_build(C.byref(tax), n, ext, par, names, ranks):
            raise RuntimeError("taxonomy build failed")
        value_bits = 1
        while (1 << value_bits) < tax.node_count:
            value_bits += 1
        if capacity is None:
            parts = []
            for _, seq in genomes:
                mins, amb = scan_positions(opts, seq)
                parts.append(np.unique(mins[amb == 0]))
            distinct = np.unique(np.concatenate(parts)) if parts else np.zeros(0, np.uint64)
            capacity = max(8, int(np.ceil(len(distinct) / load_factor)))
        cht = Cht()
        if L.k2o_cht_alloc(C.byref(cht), capacity, value_bits):
            raise RuntimeError("cht alloc failed")
        for ext_id, seq in genomes:
            internal = L.k2o_taxonomy_internal_id(C.byref(tax), ext_id)
            if not internal:
                raise ValueError(f"taxid {ext_id} not in taxonomy")
            if L.k2o_build_add_sequence(C.byref(cht), C.byref(tax), C.byref(opts), seq,
                                        len(seq), internal):
                raise RuntimeError("table full")
        return OracleDb(opts, tax, cht)

    @staticmethod
    def load(db_dir: str) -> "OracleDb":
        L = lib()
        opts, tax, cht = IndexOptions(), Taxonomy(), Cht()
        if L.k2o_load_opts(os.path.join(db_dir, "opts.k2d").encode(), C.byref(opts)):
            raise IOError("opts.k2d")
        if L.k2o_load_taxonomy(os.path.join(db_dir, "taxo.k2d").encode(), C.byref(tax)):
            raise IOError("taxo.k2d")
        if L.k2o_load_cht(os.path.join(db_dir, "hash.k2d").encode(), C.byref(cht)):
            raise IOError("hash.k2d")
        return OracleDb(opts, tax, cht)

    @staticmethod
    def from_arrays(opts: IndexOptions, tax: Taxonomy, cells: np.ndarray, capacity: int,
                    size: int, value_bits: int) -> "OracleDb":
        """Wrap an existing uint32 cell array (e.g. copied back from the GPU builder)."""
        cht = Cht()
        cht.capacity, cht.size = capacity, size
        cht.value_bits, cht.key_bits = value_bits, 32 - value_bits
        cells = np.ascontiguousarray(cells, dtype=np.uint32)
        cht.cells = cells.ctypes.data_as(C.POINTER(C.c_uint32))
        cht.owns_cells = 0
        db = OracleDb(opts, tax, cht)
        db._keepalive = cells
        return db

    def save(self, db_dir: str) -> None:
        L = lib()
        os.makedirs(db_dir, exist_ok=True)
        assert L.k2o_save_opts(os.path.join(db_dir, "opts.k2d").encode(), C.byref(self.opts)) == 0
        assert L.k2o_save_taxonomy(os.path.join(db_dir, "taxo.k2d").encode(), C.byref(self.tax)) == 0
        assert L.k2o_save_cht(os.path.join(db_dir, "hash.k2d").encode(), C.byref(self.cht)) == 0

    # -- queries -----------------------------------------------------
    def _db(self) -> Db:
        d = Db()
        d.opts = C.pointer(self.opts)
        d.cht = C.pointer(self.cht)
        d.tax = C.pointer(self.tax)
        d.confidence = self.confidence
        d.minimum_hit_groups = self.min_hit_groups
        return d

    def get(self, key: int) -> int:
        return lib().k2o_cht_get(C.byref(self.cht), key)

    def cells(self) -> np.ndarray:
        return np.ctypeslib.as_array(self.cht.cells, shape=(self.cht.capacity,))

    def parents(self) -> np.ndarray:
        return np.array([self.tax.nodes[i].parent_id for i in range(self.tax.node_count)],
                        dtype=np.uint64)

    def external_ids(self) -> np.ndarray:
        return np.array([self.tax.nodes[i].external_id for i in range(self.tax.node_count)],
                        dtype=np.uint64)

    def internal_id(self, ext_id: int) -> int:
        return lib().k2o_taxonomy_internal_id(C.byref(self.tax), ext_id)

    def lca(self, a: int, b: int) -> int:
        return lib().k2o_lca(C.byref(self.tax), a, b)

    def classify_one(self, seq1: bytes, seq2: bytes | None = None, want_taxa: bool = False):
        L = lib()
        sc = Scanner()
        o = self.opts
        assert L.k2o_scanner_init(C.byref(sc), o.k, o.l, o.spaced_seed_mask, o.dna_db,
                                  o.toggle_mask, o.revcom_version) == 0
        hc = HitCounts()
        res = ReadResult()
        taxa_p = C.POINTER(C.c_uint64)()
        taxa_n = C.c_size_t(0)
        d = self._db()
        L.k2o_classify_sequence(C.byref(d), C.byref(sc), C.byref(hc), seq1, len(seq1),
                                seq2 if seq2 is not None else None,
                                len(seq2) if seq2 is not None else 0,
                                1 if seq2 is not None else 0, C.byref(res),
                                C.byref(taxa_p) if want_taxa else None,
                                C.byref(taxa_n) if want_taxa else None)
        L.k2o_scanner_free(C.byref(sc))
        _libc.free(C.cast(hc.taxon, C.c_void_p))
        _libc.free(C.cast(hc.count, C.c_void_p))
        out = {f: getattr(res, f) for f, _ in ReadResult._fields_}
        if want_taxa:
            n = taxa_n.value
            out["taxa"] = [taxa_p[i] for i in range(n)]
            sp = L.k2o_hitlist_string(C.byref(self.tax), taxa_p, n)
            out["hitlist"] = C.string_at(sp).decode()
            _libc.free(sp)
            _libc.free(C.cast(taxa_p, C.c_void_p))
        return out

    def classify_batch(self, bases: np.ndarray, offsets: np.ndarray, paired: bool = False,
                       threads: int = 0):
        """Returns dict(call, ext, total_kmers, hit_groups, lookups, cells, sectors)."""
        bases = np.ascontiguousarray(bases, dtype=np.uint8)
        offsets = np.ascontiguousarray(offsets, dtype=np.uint64)
        n_seqs = len(offsets) - 1
        n_units = n_seqs // 2 if paired else n_seqs
        call = np.zeros(n_units, np.uint32)
        ext = np.zeros(n_units, np.uint32)
        tk = np.zeros(n_units, np.uint32)
        hg = np.zeros(n_units, np.uint32)
        totals = np.zeros(3, np.uint64)
        d = self._db()
        rc = lib().k2o_classify_batch(C.byref(d), bases.ctypes.data, offsets.ctypes.data, n_units,
                                      int(paired), threads, call.ctypes.data, ext.ctypes.data,
                                      tk.ctypes.data, hg.ctypes.data, totals.ctypes.data)
        if rc:
            raise RuntimeError("oracle classify_batch failed")
        return dict(call=call, ext=ext, total_kmers=tk, hit_groups=hg,
                    lookups=int(totals[0]), cells=int(totals[1]), sectors=int(totals[2]))


def scan_positions(opts: IndexOptions, seq: bytes):
    """(minimizer[u64], ambiguous[u8]) for every NextMinimizer() return on seq."""
    cap = max(0, len(seq) - int(opts.k) + 1) + 1
    mins = np.zeros(cap, np.uint64)
    amb = np.zeros(cap, np.uint8)
    n = lib().k2o_scan_positions(C.byref(opts), seq, len(seq), mins.ctypes.data, amb.ctypes.data,
                                 cap)
    assert n <= cap
    return mins[:n], amb[:n]


def fmix64(x: int) -> int:
    return lib().k2o_fmix64(x)
