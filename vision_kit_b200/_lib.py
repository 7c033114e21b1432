"""ctypes binding of libvk_b200.so (include/vk_b200.h).

There is no fallback: if the shared library is missing or lacks a declared symbol this
module raises, and every entry point raises ``VkError`` on a non-zero return code.
"""
from __future__ import annotations

import ctypes as C
import os
import re

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("VK_B200_LIB") or os.path.join(HERE, "libvk_b200.so")   # (override: profiling builds)
HEADER = os.path.join(os.path.dirname(HERE), "include", "vk_b200.h")

VK_MAX_LEVELS = 4
VK_MAX_ANCHORS = 8
VK_MAX_DET = 1024
VK_MAX_SEGMENTS = 2048
VK_LB_F32_NCHW, VK_LB_BF16_NCHW, VK_LB_U8_NHWC = 0, 1, 2
VK_HEAD_V5, VK_HEAD_V7 = 0, 1
VK_F32, VK_F16, VK_BF16 = 0, 1, 2
VK_FILTER_AUTO, VK_FILTER_SPARSE, VK_FILTER_DENSE, VK_FILTER_DENSE_ONEPASS = 0, 1, 2, 3
VK_CONV_TILE, VK_CONV_PERSISTENT = 0, 1
VK_HIST_BINS = 1024
VK_CTRL_WORDS = 4
VK_FLAG_LIST = 2


class VkError(RuntimeError):
    def __init__(self, fn: str, code: int, msg: str):
        super().__init__(f"{fn} failed with code {code}: {msg}")
        self.code = code


class VkLbGeom(C.Structure):
    _fields_ = [("ratio", C.c_double), ("pad_w", C.c_double), ("pad_h", C.c_double),
                ("new_w", C.c_int32), ("new_h", C.c_int32),
                ("top", C.c_int32), ("bottom", C.c_int32), ("left", C.c_int32), ("right", C.c_int32),
                ("out_h", C.c_int32), ("out_w", C.c_int32),
                ("needs_resize", C.c_int32), ("reserved", C.c_int32)]


class VkLbDesc(C.Structure):
    _fields_ = [("src", C.c_void_p), ("pitch", C.c_int64),
                ("src_h", C.c_int32), ("src_w", C.c_int32),
                ("new_h", C.c_int32), ("new_w", C.c_int32),
                ("top", C.c_int32), ("left", C.c_int32),
                ("reserved0", C.c_int32), ("reserved1", C.c_int32)]


class VkHeadCfg(C.Structure):
    _fields_ = [("variant", C.c_int32), ("nl", C.c_int32), ("na", C.c_int32), ("nc", C.c_int32),
                ("ny", C.c_int32 * VK_MAX_LEVELS), ("nx", C.c_int32 * VK_MAX_LEVELS),
                ("stride", C.c_float * VK_MAX_LEVELS),
                ("anchors", (C.c_float * (2 * VK_MAX_ANCHORS)) * VK_MAX_LEVELS)]


class VkCandBuf(C.Structure):
    _fields_ = [("cand", C.c_void_p), ("boxes", C.c_void_p), ("ctrl", C.c_void_p),
                ("seg_count", C.c_void_p), ("list", C.c_void_p),
                ("cap", C.c_int32), ("rows", C.c_int32), ("segs", C.c_int32), ("nc", C.c_int32),
                ("list_cap", C.c_int32), ("reserved", C.c_int32)]


assert C.sizeof(VkLbDesc) == 48 and C.sizeof(VkLbGeom) == 64

_P = C.c_void_p
_PROTOS = {
    "vk_version": (C.c_int, []),
    "vk_last_error": (C.c_char_p, []),
    "vk_launch_count": (C.c_uint64, []),
    "vk_build_arch": (C.c_int, []),
    "vk_letterbox_geometry": (C.c_int, [C.c_int] * 8 + [C.POINTER(VkLbGeom)]),
    "vk_dataset_geometry": (C.c_int, [C.c_int] * 4 + [C.POINTER(VkLbGeom)]),
    "vk_letterbox_workspace_bytes": (C.c_size_t, [C.c_int] * 3),
    "vk_letterbox_batch": (C.c_int, [_P, _P, C.c_int, C.c_int, C.c_int, C.c_int, C.c_uint32,
                                     C.c_int, _P, _P, C.c_size_t, _P]),
    "vk_head_rows": (C.c_int, [C.POINTER(VkHeadCfg)]),
    "vk_detect_decode": (C.c_int, [C.POINTER(VkHeadCfg), _P, C.c_int, C.c_int, _P, _P, _P]),
    "vk_cand_tile_slots": (C.c_int, [C.c_int, C.c_int]),
    "vk_filter_segments": (C.c_int, [C.c_int]),
    "vk_decode_filter_segments": (C.c_int, [C.POINTER(VkHeadCfg)]),
    "vk_filter_pred": (C.c_int, [_P, C.c_int, C.c_int, C.c_int, C.c_int, C.c_float, C.c_int, _P, C.c_int,
                                 C.POINTER(VkCandBuf), _P]),
    "vk_decode_filter": (C.c_int, [C.POINTER(VkHeadCfg), _P, C.c_int, C.c_int, C.c_float, C.c_int, _P, C.c_int,
                                   C.POINTER(VkCandBuf), _P]),
    "vk_conv_decode_filter": (C.c_int, [C.POINTER(VkHeadCfg), _P, _P, _P, _P, C.c_int, C.c_float, C.c_int, _P, C.c_int,
                                        C.POINTER(VkCandBuf), _P, _P]),
    "vk_nms_batched": (C.c_int, [C.POINTER(VkCandBuf), C.c_int, C.c_double, C.c_int,
                                 C.c_int, C.c_int, C.c_float, _P, _P, _P, _P, _P]),
    "vk_scale_coords": (C.c_int, [_P, C.c_int, C.c_int, C.c_float, C.c_float, C.c_float, C.c_int,
                                  C.c_float, C.c_float, _P]),
    "vk_cxcywh_to_xyxy": (C.c_int, [_P, _P, C.c_int, _P]),
    "vk_eval_match_smem_bytes": (C.c_size_t, [C.c_int, C.c_int]),
    "vk_eval_match": (C.c_int, [_P, _P, C.c_int, C.c_int, _P, _P, C.c_int, _P, C.c_int, C.c_int, C.c_int, _P, C.c_int,
                                _P, _P, _P, _P]),
}

_lib = None


def declared_symbols() -> list[str]:
    """Every function name include/vk_b200.h declares."""
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(vk_[a-z0-9_]+)\s*\(", src)))


def lib() -> C.CDLL:
    """Loads the library once; raises if it or any declared symbol is missing."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            f"{LIB_PATH} not found: build it with `python __graft_entry__.py` "
            "(vision_kit_b200 has no CPU fallback)")
    handle = C.CDLL(LIB_PATH)
    names = declared_symbols() if os.path.exists(HEADER) else list(_PROTOS)
    for name in names:
        if not hasattr(handle, name):
            raise RuntimeError(f"{LIB_PATH} does not export {name}")
    for name, (res, args) in _PROTOS.items():
        fn = getattr(handle, name)
        fn.restype = res
        fn.argtypes = args
    _lib = handle
    return handle


def check(fn: str, rc: int) -> None:
    if rc != 0:
        raise VkError(fn, rc, lib().vk_last_error().decode(errors="replace"))


def launch_count() -> int:
    return int(lib().vk_launch_count())


def require_cuda(t, what: str):
    """Product entry points take CUDA tensors only; anything else is an error."""
    if not t.is_cuda:
        raise RuntimeError(f"{what}: expected a CUDA tensor, got {t.device} "
                           "(vision_kit_b200 has no CPU path)")
    return t


def stream_ptr():
    import torch
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)
