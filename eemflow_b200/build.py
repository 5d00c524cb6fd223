"""Build recipe for libeemflow_b200.so (sm_100a only, in-tree).

`python -m eemflow_b200.build` compiles every `csrc/*.cu` with nvcc for
`-gencode arch=compute_100a,code=sm_100a` and links them into
`eemflow_b200/libeemflow_b200.so`.  nvcc cross-compiles without a GPU.  The
library links the static CUDA runtime and resolves `cuTensorMapEncodeTiled`
at run time through `cudaGetDriverEntryPoint`, so it has no link-time
dependency on libcuda or libtorch.
"""
from __future__ import annotations

import hashlib
import os
import shutil
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor
from pathlib import Path

PKG_DIR = Path(__file__).resolve().parent
CSRC = PKG_DIR / "csrc"
OBJ_DIR = PKG_DIR / "csrc" / "_obj"
LIB_PATH = PKG_DIR / "libeemflow_b200.so"

NVCC_FLAGS = [
    "-O3",
    "-std=c++17",
    "-lineinfo",
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-Xcompiler", "-fPIC,-fvisibility=hidden,-Wall,-Wno-unused-function",
    "-Xptxas", "-v",
]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and Path(cand).exists():
            return cand
    raise RuntimeError("nvcc not found; set NVCC or add /usr/local/cuda/bin to PATH")


def _sources() -> list[Path]:
    return sorted(CSRC.glob("*.cu"))


def _fingerprint(src: Path) -> str:
    h = hashlib.sha256()
    h.update(" ".join(NVCC_FLAGS).encode())
    h.update(src.read_bytes())
    for hdr in sorted(CSRC.glob("*.cuh")) + [PKG_DIR.parent / "include" / "eemflow_b200.h"]:
        h.update(hdr.read_bytes())
    return h.hexdigest()


def _compile_one(src: Path, verbose: bool) -> Path:
    OBJ_DIR.mkdir(parents=True, exist_ok=True)
    obj = OBJ_DIR / (src.stem + ".o")
    stamp = OBJ_DIR / (src.stem + ".sha")
    fp = _fingerprint(src)
    if obj.exists() and stamp.exists() and stamp.read_text() == fp:
        return obj
    cmd = [_nvcc(), *NVCC_FLAGS, "-c", str(src), "-o", str(obj)]
    res = subprocess.run(cmd, capture_output=True, text=True)
    log = OBJ_DIR / (src.stem + ".ptxas.log")
    log.write_text(res.stdout + res.stderr)
    if res.returncode != 0:
        raise RuntimeError(f"nvcc failed for {src.name}:\n{res.stdout}\n{res.stderr}")
    if verbose:
        sys.stderr.write(f"[eemflow_b200.build] compiled {src.name}\n")
    stamp.write_text(fp)
    return obj


def build(verbose: bool = True, force: bool = False) -> Path:
    """Compile and link the library; returns its path.  Incremental unless force=True."""
    if force and OBJ_DIR.exists():
        shutil.rmtree(OBJ_DIR)
    srcs = _sources()
    if not srcs:
        raise RuntimeError(f"no CUDA sources under {CSRC}")
    with ThreadPoolExecutor(max_workers=min(8, len(srcs))) as pool:
        objs = list(pool.map(lambda s: _compile_one(s, verbose), srcs))
    newest = max(o.stat().st_mtime for o in objs)
    if force or not LIB_PATH.exists() or LIB_PATH.stat().st_mtime < newest:
        cmd = [_nvcc(), "-shared", "-o", str(LIB_PATH), *map(str, objs),
               "-gencode", "arch=compute_100a,code=sm_100a", "-cudart", "static"]
        res = subprocess.run(cmd, capture_output=True, text=True)
        if res.returncode != 0:
            raise RuntimeError(f"link failed:\n{res.stdout}\n{res.stderr}")
        if verbose:
            sys.stderr.write(f"[eemflow_b200.build] linked {LIB_PATH}\n")
    return LIB_PATH


if __name__ == "__main__":
    build(verbose=True, force="--force" in sys.argv)
