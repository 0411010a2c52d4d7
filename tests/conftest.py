import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run with -m gpu on the B200 box)")


@pytest.fixture(scope="session", autouse=True)
def ensure_built():
    """The library and the CLI are built in-tree (python -m nohuman_b200.build); rebuild only if stale or missing."""
    from nohuman_b200 import build
    build.build()


@pytest.fixture(scope="session")
def oracle():
    from oracle import k2oracle
    k2oracle.lib()
    return k2oracle


@pytest.fixture(scope="session")
def small_db(oracle, tmp_path_factory):
    """cfg1-shaped database at 1/50 scale: synthetic human chr + 3 bacterial
    genomes, 23-node taxonomy, k=35 l=31 s=7, load factor 0.7."""
    import synth
    genomes = synth.cfg1_genomes(seed=1, scale=0.02)
    tax = [oracle.TaxSpec(*t) for t in synth.TAXONOMY_CFG1]
    db = oracle.OracleDb.build([(t, bytes(g)) for t, g in genomes], tax)
    d = tmp_path_factory.mktemp("db_cfg1_small")
    db.save(str(d))
    db.genomes = genomes
    db.path = str(d)
    return db


def have_gpu() -> bool:
    try:
        from nohuman_b200 import _ffi
        return _ffi.lib().nh_device_count() > 0
    except Exception:
        return False


@pytest.fixture(scope="session")
def gpu_db(small_db):
    from nohuman_b200 import Database
    db = Database.open(small_db.path, 0)
    yield db
    db.close()
