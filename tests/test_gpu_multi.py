"""nh_db_open_multi: one disk read, one replica per device (NCCL broadcast or the peer-copy tree).
The two-device cases need `gpurun --gpus 2`; on a one-GPU box they skip."""
import numpy as np
import pytest

import synth

pytestmark = pytest.mark.gpu


def _n_devices():
    from nohuman_b200 import _ffi
    return _ffi.lib().nh_device_count()


def test_open_multi_on_one_device_is_nh_db_open(small_db):
    from nohuman_b200 import Database, Session
    (db,) = Database.open_multi(small_db.path, [0])
    try:
        assert db.info.replicated_by == 0 and int(db.info.capacity) == small_db.cht.capacity
        seqs = synth.illumina_reads(small_db.genomes, 500, 150, seed=8)
        bases, offsets = synth.pack(seqs)
        with Session(db, confidence=0.1) as s:
            call, _, _ = s.classify(bases, offsets)
        small_db.confidence = 0.1
        np.testing.assert_array_equal(call, small_db.classify_batch(bases, offsets)["ext"])
    finally:
        db.close()
    from nohuman_b200 import NhError
    with pytest.raises(NhError):
        Database.open_multi(small_db.path, [0, 0])


@pytest.mark.parametrize("how", ["auto", "p2p"])
def test_open_multi_replicas_answer_alike(small_db, how, monkeypatch):
    if _n_devices() < 2:
        pytest.skip("needs two CUDA devices")
    from nohuman_b200 import Database, Session
    if how == "p2p":
        monkeypatch.setenv("NH_DB_REPLICATE", "p2p")
    devs = list(range(min(_n_devices(), 8)))
    dbs = Database.open_multi(small_db.path, devs)
    try:
        assert dbs[0].info.replicated_by == 0
        for d in dbs[1:]:
            assert d.info.replicated_by == (2 if how == "p2p" else d.info.replicated_by) and d.info.replicated_by in (1, 2)
        seqs = synth.illumina_reads(small_db.genomes, 3000, 150, seed=9, paired=True)
        bases, offsets = synth.pack(seqs)
        small_db.confidence = 0.2
        want = small_db.classify_batch(bases, offsets, paired=True)["ext"]
        for d in dbs:
            with Session(d, confidence=0.2, paired=True) as s:
                call, _, _ = s.classify(bases, offsets)
            np.testing.assert_array_equal(call, want)
    finally:
        for d in dbs:
            d.close()
