"""strainscan_b200 -- B200-native match+count engine for StrainScan's identification hot path.

Only what the path needs: csrc/ (sm_100a CUDA kernels + the C ABI of include/strainscan_b200.h),
engine.py (ctypes object layer), identify_shim.py / l2_shim.py (drop-in mirrors of the reference's
count adapters), dist.py (read sharding + NCCL sum).  No CPU fallback anywhere.
"""
from ._lib import LIB_PATH, SS_REC_IN_SET, SS_REC_IS_LAST, SS_REC_RAW_UPPER, StrainScanB200Error, SynthParams, Stats  # noqa: F401
from .engine import Engine, KmerSet, Reads  # noqa: F401

__all__ = ["Engine", "KmerSet", "Reads", "StrainScanB200Error", "SynthParams", "Stats", "LIB_PATH"]
