"""nohuman_b200 — B200-native (sm_100a) replacement for the kraken2
classification step that mbhall88/nohuman shells out to.

Only the hot path lives here: csrc/ (CUDA kernels + the C ABI declared in
include/nohuman_gpu.h) and a thin host-side mirror of how the reference
drives kraken2 (api.py).  There is no CPU fallback.
"""
from .api import BatchStats, Database, DbInfo, NhError, Session, parse_confidence_score  # noqa: F401

__all__ = ["Database", "Session", "NhError", "BatchStats", "DbInfo", "parse_confidence_score"]
