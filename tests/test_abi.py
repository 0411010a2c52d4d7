"""The C-ABI library loads on a CPU-only box and exports every symbol that
include/*.h declares; without a CUDA device every compute entry point fails
loudly (no CPU fallback)."""
import ctypes as C
import os
import re
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
INCLUDE = os.path.join(ROOT, "include")


def declared_symbols(header):
    src = open(os.path.join(INCLUDE, header)).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(nh_[a-z0-9_]+)\s*\(", src)))


@pytest.fixture(scope="module")
def nh():
    from nohuman_b200 import _ffi
    if not os.path.exists(_ffi.LIB_PATH):
        from nohuman_b200 import build
        build.build()
    return _ffi


@pytest.mark.parametrize("header", sorted(f for f in os.listdir(INCLUDE) if f.endswith(".h")))
def test_every_declared_symbol_is_exported(nh, header):
    L = C.CDLL(nh.LIB_PATH)
    names = declared_symbols(header)
    assert len(names) >= 4
    for n in names:
        assert hasattr(L, n), f"{n} declared in include/{header} but not exported"


def test_binding_table_covers_the_header(nh):
    from nohuman_b200 import synth
    bound = set(nh.SYMBOLS) | set(synth._SYNTH_SYMBOLS)
    for header in os.listdir(INCLUDE):
        if header.endswith(".h"):
            missing = set(declared_symbols(header)) - bound
            assert not missing, f"no ctypes binding for {sorted(missing)}"


def test_abi_version_and_struct_sizes(nh):
    L = nh.lib()
    assert L.nh_abi_version() == 3
    # struct layouts the ctypes mirror must agree with (sizes asserted against the C compiler)
    prog = r'''
#include <stdio.h>
#include "nohuman_gpu.h"
#include "nohuman_synth.h"
int main(void){printf("%zu %zu %zu %zu %zu %zu\n", sizeof(nh_db_info_t), sizeof(nh_params_t),
  sizeof(nh_batch_stats_t), sizeof(nh_run_stats_t), sizeof(nh_synth_db_params_t), sizeof(nh_synth_reads_params_t));return 0;}
'''
    import tempfile
    d = tempfile.mkdtemp()
    with open(os.path.join(d, "s.c"), "w") as f:
        f.write(prog)
    subprocess.check_call(["gcc", "-I", INCLUDE, "-o", os.path.join(d, "s"), os.path.join(d, "s.c")])
    sizes = [int(x) for x in subprocess.check_output([os.path.join(d, "s")]).split()]
    from nohuman_b200 import synth
    mirror = [nh.DbInfo, nh.Params, nh.BatchStats, nh.RunStats, synth.SynthDbParams, synth.SynthReadsParams]
    assert sizes == [C.sizeof(m) for m in mirror]


def test_no_cpu_fallback(nh, small_db):
    """On a box without a GPU the product path refuses to run."""
    L = nh.lib()
    if L.nh_device_count() > 0:
        pytest.skip("a CUDA device is present")
    from nohuman_b200 import Database, NhError
    with pytest.raises(NhError) as ei:
        Database.open(small_db.path, 0)
    assert ei.value.code == nh.NH_ERR_CUDA
    assert "no CPU path" in ei.value.message
    with pytest.raises(NhError) as ei:
        Database.open("/nonexistent/db", 0)
    assert ei.value.code == nh.NH_ERR_IO
    assert "hash.k2d" in ei.value.message


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "nohuman_b200")
    for dirpath, _, files in os.walk(pkg):
        for fn in files:
            if fn.endswith((".py", ".cu", ".cuh", ".h", ".cc", ".cpp")):
                txt = open(os.path.join(dirpath, fn), errors="replace").read()
                assert "k2_oracle" not in txt and "k2oracle" not in txt and "libk2oracle" not in txt, fn


def test_parse_confidence_score():
    """src/lib.rs:203-221 cases + the f32 -> shortest decimal -> double trip of src/main.rs:213"""
    from nohuman_b200 import parse_confidence_score as p
    assert p("0.5") == 0.5 and p("1.0") == 1.0 and p("0.0") == 0.0
    for bad in ("1.1", "-0.1", "abc"):
        with pytest.raises(ValueError):
            p(bad)
    assert p("0.1") == 0.1            # not 0.100000001490116 (the f32 widened)
    assert p("0.3") == 0.3
    assert p("0.123456789") == 0.12345679  # f32 rounding is visible
