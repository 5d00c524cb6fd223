"""CPU ORACLE -- TEST INFRASTRUCTURE ONLY.  Builds and binds oracle/voxel_oracle.c (plain C)."""
from __future__ import annotations

import ctypes as C
import subprocess
from pathlib import Path

import numpy as np

_DIR = Path(__file__).resolve().parent
SRC = _DIR / "voxel_oracle.c"
LIB = _DIR / "_build" / "libvoxel_oracle.so"
_lib = None


def build(force: bool = False) -> Path:
    if force or not LIB.exists() or LIB.stat().st_mtime < SRC.stat().st_mtime:
        LIB.parent.mkdir(parents=True, exist_ok=True)
        subprocess.run(["gcc", "-O2", "-fPIC", "-shared", "-ffp-contract=off", "-o", str(LIB), str(SRC), "-lm"],
                       check=True)
    return LIB


def _load():
    global _lib
    if _lib is None:
        h = C.CDLL(str(build()))
        h.oracle_voxel_vote.restype = C.c_int64
        h.oracle_voxel_vote.argtypes = [C.c_void_p, C.c_int64, C.c_int, C.c_int, C.c_int, C.c_void_p]
        h.oracle_voxel_normalize.restype = None
        h.oracle_voxel_normalize.argtypes = [C.c_void_p, C.c_int64, C.c_void_p]
        _lib = h
    return _lib


def voxelize(features: np.ndarray, num_bins: int, height: int, width: int, normalize: bool = True):
    """-> (grid float32 [nb,H,W], dropped votes, stats (count, mean, std))."""
    ev = np.ascontiguousarray(features, dtype=np.float64)
    grid = np.empty((num_bins, height, width), dtype=np.float32)
    h = _load()
    dropped = h.oracle_voxel_vote(ev.ctypes.data, ev.shape[0], num_bins, height, width, grid.ctypes.data)
    stats = np.zeros(3, dtype=np.float64)
    if normalize:
        h.oracle_voxel_normalize(grid.ctypes.data, grid.size, stats.ctypes.data)
    return grid, int(dropped), stats
