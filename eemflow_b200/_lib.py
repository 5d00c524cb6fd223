"""ctypes binding of libeemflow_b200.so (the C ABI declared in include/eemflow_b200.h).

Loading the library creates no CUDA context (safe in fork/forkserver DataLoader workers, which
is where the reference voxelizes: utils/transformers.py:21-25, loader/MVSEC.py:46-50).  There is
deliberately no CPU fallback: if the library is missing, or a tensor is not on a CUDA device,
the call raises.
"""
from __future__ import annotations

import ctypes as C
import threading
from pathlib import Path

import torch

_PKG_DIR = Path(__file__).resolve().parent
LIB_PATH = _PKG_DIR / "libeemflow_b200.so"

EEM_OK = 0
EEM_ERR_BAD_ARG = -1
EEM_ERR_MISALIGNED = -2
EEM_ERR_WORKSPACE = -3
EEM_ERR_CUDA = -4
EEM_ERR_UNSUPPORTED = -5

VOXEL_ATOMIC, VOXEL_DETERMINISTIC = 0, 1
CORR_FP32, CORR_TF32 = 0, 1
PRECISIONS = ("fp32", "tf32", "tf32_f16")
WARP_EXACT, WARP_HALFPIX = 0, 1
MASK_NONE, MASK_GE1, MASK_9999 = 0, 1, 2

_vp, _i, _i64, _sz, _f = C.c_void_p, C.c_int, C.c_int64, C.c_size_t, C.c_float

# name -> (restype, argtypes); mirrors include/eemflow_b200.h one to one.
SIGNATURES = {
    "eem_version": (_i, []),
    "eem_last_error_string": (C.c_char_p, []),
    "eem_sm_count": (_i, []),
    "eem_launch_count": (C.c_longlong, []),
    "eem_voxelize_workspace_bytes": (_sz, [_i64, _i, _i, _i, _i, _i, _i]),
    "eem_voxelize": (_i, [_vp, _vp, _i, _i64, _i64, _i, _i, _i, _i, _i, _vp, _vp, _vp, _vp, _sz, _vp]),
    "eem_voxelize_scaled": (_i, [_vp, C.c_double, _vp, _i, _i64, _i64, _i, _i, _i, _i, _i, _vp, _vp, _vp, _vp, _sz, _vp]),
    "eem_voxelize_soa": (_i, [_vp, _i, _vp, _vp, _vp, _vp, _i, _i64, _i64, _i, _i, _i, _i, _i, _vp, _vp, _vp, _vp, _sz, _vp]),
    "eem_voxel_normalize_workspace_bytes": (_sz, [_i, _i64]),
    "eem_voxel_normalize": (_i, [_vp, _i, _i64, _vp, _vp, _sz, _vp]),
    "eem_corr_pyramid_workspace_bytes": (_sz, [_i, _i, _i, _i, _i]),
    "eem_corr_pyramid": (_i, [_vp, _vp, _i, _i, _i, _i, _i, C.POINTER(_vp), _i, _vp, _sz, _vp]),
    "eem_corr_packed_layout": (_i, [_i, _i, _i, C.POINTER(_i64), C.POINTER(_i64), C.POINTER(_i64)]),
    "eem_corr_pyramid_packed_workspace_bytes": (_sz, [_i, _i, _i, _i, _i]),
    "eem_corr_pyramid_packed": (_i, [_vp, _vp, _i, _i, _i, _i, _i, _vp, _vp, _sz, _vp]),
    "eem_corr_lookup_packed": (_i, [_vp, _i, _i, _i, _i, _i, _vp, _vp, _vp]),
    "eem_corr_pyramid_unpack": (_i, [_vp, _i, _i, _i, _i, C.POINTER(_vp), _vp]),
    "eem_avg_pool2x2": (_i, [_vp, _i64, _i, _i, _vp, _vp]),
    "eem_corr_lookup": (_i, [C.POINTER(_vp), _i, _i, _i, _i, _i, _vp, _vp, _vp]),
    "eem_local_corr": (_i, [_vp, _vp, _i, _i, _i, _i, _i, C.POINTER(_i), _i, _f, _vp, _vp]),
    "eem_local_corr_tf32": (_i, [_vp, _vp, _i, _i, _i, _i, _i, C.POINTER(_i), _i, _f, _vp, _vp]),
    "eem_local_corr_tf32_supported": (_i, [_i, _i, _i, _i, _i]),
    "eem_backwarp": (_i, [_vp, _vp, _i, _i, _i, _i, _i, _i, _vp, _vp, _vp]),
    "eem_upsample_flow_warp": (_i, [_vp, _i, _i, _f, _f, _vp, _i, _i, _i, _i, _vp, _vp, _vp]),
    "eem_blend_flow_warp": (_i, [_vp, _vp, _vp, _vp, _i, _i, _i, _i, _vp, _vp, _vp]),
    "eem_warp_blend": (_i, [_vp, _vp, _vp, _i, _i, _i, _vp, _vp]),
    "eem_bilinear_resize": (_i, [_vp, _i, _i, _i, _i, _vp, _i, _i, _i, _f, _f, _f, _vp]),
    "eem_bilinear_sample": (_i, [_vp, _vp, _i, _i, _i, _i, _i, _i, _vp, _vp, _vp]),
    "eem_bilinear_resize_multi": (_i, [C.POINTER(_vp), C.POINTER(_i), C.POINTER(_i), _i, _i, _i, C.POINTER(_vp), _i, _i, _i,
                                       C.POINTER(_f), C.POINTER(_f), _f, _vp]),
    "eem_scale_uv_inplace_multi": (_i, [C.POINTER(_vp), C.POINTER(_i), C.POINTER(_i), _i, _i, _i, C.POINTER(_f), C.POINTER(_f), _vp]),
    "eem_scale_uv_inplace": (_i, [_vp, _i, _i, _i, _i, _f, _f, _vp]),
    "eem_corr_lookup_backward": (_i, [_vp, _vp, _i, _i, _i, _i, _i, C.POINTER(_vp), _vp]),
    "eem_avg_pool2x2_backward": (_i, [_vp, _i64, _i, _i, _vp, _i, _vp]),
    "eem_batched_gemm_f32": (_i, [_vp, _vp, _vp, _i, _i, _i, _i, _i64, _i64, _i64, _i64, _i64, _i64, _i, _f, _i, _vp]),
    "eem_batched_gemm_tf32": (_i, [_vp, _vp, _vp, _i, _i, _i, _i, _i64, _i64, _i64, _i64, _i64, _i64, _i, _f, _i, _vp]),
    "eem_batched_gemm_tf32_multi": (_i, [C.POINTER(_vp), C.POINTER(_vp), _vp, _i, _i, _i, _i, C.POINTER(_i), C.POINTER(_i64),
                                         C.POINTER(_i64), _i64, C.POINTER(_i64), C.POINTER(_i64), _i64, _i, _f, _i, _vp]),
    "eem_batched_gemm_tf32_supported": (_i, [_i, _i, _i, _i, _i64, _i64, _i64, _i64, _i]),
    "eem_local_corr_backward": (_i, [_vp, _vp, _vp, _i, _i, _i, _i, _i, C.POINTER(_i), _i, _f, _vp, _vp, _vp]),
    "eem_backwarp_backward": (_i, [_vp, _vp, _vp, _i, _i, _i, _i, _i, _i, _vp, _vp, _vp]),
    "eem_bilinear_sample_backward": (_i, [_vp, _vp, _vp, _i, _i, _i, _i, _i, _i, _vp, _vp, _vp]),
    "eem_bilinear_resize_backward": (_i, [_vp, _i, _i, _i, _i, _i, _i, _i, _f, _f, _f, _vp, _vp]),
    "eem_event_mask": (_i, [_vp, _vp, _i, _i64, _i, _i, _vp, _vp]),
    "eem_voxel_bin_sum": (_i, [_vp, _i64, _i, _i, _i, _vp, _vp]),
    "eem_flow_error": (_i, [_vp, _vp, _vp, _i, _i, _i, _i, _vp, _vp]),
    "eem_motion_propagate": (_i, [_vp, _i, _i, _i, _i, _i, _vp, _vp]),
    "eem_replicate_pad": (_i, [_vp, _i, _i, _i, _i, _i, _i, _i, _i, _vp, _vp]),
}

_lock = threading.Lock()
_lib = None


class EemError(RuntimeError):
    """A call into libeemflow_b200.so failed; .code holds the eem_status."""

    def __init__(self, code: int, msg: str):
        super().__init__(f"libeemflow_b200: {msg} (status {code})")
        self.code = code


def lib() -> C.CDLL:
    """Load (once) and return the shared library; raises if it has not been built."""
    global _lib
    if _lib is None:
        with _lock:
            if _lib is None:
                if not LIB_PATH.exists():
                    raise ImportError(
                        f"{LIB_PATH} is missing. Build it with `python -m eemflow_b200.build` "
                        "(nvcc, sm_100a). There is no CPU fallback for this path."
                    )
                h = C.CDLL(str(LIB_PATH))
                for name, (res, args) in SIGNATURES.items():
                    fn = getattr(h, name)
                    fn.restype = res
                    fn.argtypes = args
                _lib = h
    return _lib


def check(status: int) -> None:
    """Translate an eem_status into the exception the reference would have raised."""
    if status == EEM_OK:
        return
    msg = lib().eem_last_error_string().decode("utf-8", "replace")
    if status == EEM_ERR_BAD_ARG:
        # the reference asserts on bad sizes (utils/transformers.py:51-54)
        raise AssertionError(msg)
    if status == EEM_ERR_UNSUPPORTED:
        raise NotImplementedError(msg)
    raise EemError(status, msg)


def stream_ptr(device: torch.device) -> int:
    return torch.cuda.current_stream(device).cuda_stream


def ptr(t: torch.Tensor | None) -> int | None:
    return None if t is None else t.data_ptr()


def require_cuda(t: torch.Tensor, name: str, dtype: torch.dtype = torch.float32) -> torch.Tensor:
    """The product path runs on the GPU only; anything else is a caller error, not a fallback."""
    if not t.is_cuda:
        raise RuntimeError(
            f"eemflow_b200: `{name}` must be a CUDA tensor (got device {t.device}); this package has no CPU path"
        )
    if t.dtype != dtype:
        t = t.to(dtype)
    return t.contiguous()


def ptr_array(tensors) -> "C.Array":
    arr = (_vp * len(tensors))()
    for k, t in enumerate(tensors):
        arr[k] = t.data_ptr() if t is not None and t.numel() > 0 else None
    return arr


class Workspace:
    """Grow-only scratch buffer per (device, stream, op family) handed to the C ABI (no hidden cudaMalloc in the
    library).  Keyed by the stream as well so that callers running the ops on several streams never share scratch."""

    def __init__(self):
        self._buf: dict[tuple, torch.Tensor] = {}

    def get(self, device: torch.device, nbytes: int, tag: str = "") -> torch.Tensor | None:
        if nbytes <= 0:
            return None
        index = device.index if device.index is not None else torch.cuda.current_device()
        key = (index, torch.cuda.current_stream(index).cuda_stream, tag)
        buf = self._buf.get(key)
        if buf is None or buf.numel() < nbytes:
            buf = torch.empty(max(nbytes, 1 << 20), dtype=torch.uint8, device=device)
            self._buf[key] = buf
        return buf


workspace = Workspace()
