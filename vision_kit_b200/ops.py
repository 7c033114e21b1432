"""Tensor-level wrappers over the C-ABI (include/vk_b200.h).

PyTorch is used here for device memory and streams only; every computation is a kernel of
libvk_b200.so enqueued on the current CUDA stream.  No function in this module
synchronises except where stated.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass
from typing import Optional, Sequence

import numpy as np
import torch

from . import _lib
from ._lib import (VK_BF16, VK_CONV_PERSISTENT, VK_CONV_TILE, VK_CTRL_WORDS, VK_F16, VK_F32, VK_FILTER_AUTO,
                   VK_FILTER_DENSE, VK_FILTER_DENSE_ONEPASS, VK_FILTER_SPARSE, VK_HEAD_V5, VK_HEAD_V7, VK_LB_BF16_NCHW,
                   VK_LB_F32_NCHW, VK_LB_U8_NHWC, VK_MAX_ANCHORS, VK_MAX_LEVELS, VkCandBuf, VkHeadCfg, VkLbDesc,
                   VkLbGeom)

MAX_WH = 7680          # utils/image_proc.py:107
LIST_CAP = 8192        # entries of the per-image top list the NMS kernel sorts from
_DTYPE = {torch.float32: VK_F32, torch.float16: VK_F16, torch.bfloat16: VK_BF16}
_KERNEL = {"auto": VK_FILTER_AUTO, "sparse": VK_FILTER_SPARSE, "dense": VK_FILTER_DENSE,
           "dense_onepass": VK_FILTER_DENSE_ONEPASS,
           None: VK_FILTER_AUTO, VK_FILTER_AUTO: VK_FILTER_AUTO, VK_FILTER_SPARSE: VK_FILTER_SPARSE,
           VK_FILTER_DENSE: VK_FILTER_DENSE, VK_FILTER_DENSE_ONEPASS: VK_FILTER_DENSE_ONEPASS}


def expects_dense(kernel, conf_thres: float) -> bool:
    """The library's rule for the filter kernel (and for whether a candidate buffer wants the top list of
    vk_nms_batched's selection pass: eval thresholds leave ~240 k candidates per image)."""
    k = _KERNEL[kernel]
    return k in (VK_FILTER_DENSE, VK_FILTER_DENSE_ONEPASS) or (k == VK_FILTER_AUTO and float(conf_thres) < 0.05)


def _ptr(t: Optional[torch.Tensor]) -> C.c_void_p:
    return C.c_void_p(0 if t is None else t.data_ptr())


# ---------------------------------------------------------------------------- letterbox
def letterbox_geometry(src_h: int, src_w: int, img_sz, stride: int = 32, letterbox: bool = True,
                       scaleup: bool = True, auto: bool = False) -> VkLbGeom:
    """Scalar set-up of utils/image_proc.py:22-55, computed by the library in float64."""
    if isinstance(img_sz, int):
        img_sz = (img_sz, img_sz)                     # image_proc.py:24
    g = VkLbGeom()
    _lib.check("vk_letterbox_geometry",
               _lib.lib().vk_letterbox_geometry(int(src_h), int(src_w), int(img_sz[0]), int(img_sz[1]),
                                                int(stride), int(bool(letterbox)), int(bool(scaleup)),
                                                int(bool(auto)), C.byref(g)))
    return g


def pack_color(color: Sequence[int]) -> int:
    c = [int(v) & 255 for v in color]
    return c[0] | (c[1] << 8) | (c[2] << 16)


_FMT = {torch.float32: VK_LB_F32_NCHW, torch.bfloat16: VK_LB_BF16_NCHW, torch.uint8: VK_LB_U8_NHWC}


def dataset_geometry(src_h: int, src_w: int, img_sz) -> VkLbGeom:
    """load_resized_image + PadIfNeeded geometry (data/datasets/yolo.py:144-160, data/augmentations.py:197-200)."""
    if isinstance(img_sz, int):
        img_sz = (img_sz, img_sz)
    g = VkLbGeom()
    _lib.check("vk_dataset_geometry", _lib.lib().vk_dataset_geometry(
        int(src_h), int(src_w), int(img_sz[0]), int(img_sz[1]), C.byref(g)))
    return g


class LetterboxPlan:
    """Host-side descriptors of one batch: geometry per source, the device descriptor array
    and the coefficient-table workspace.  Re-usable while the source pointers stay valid."""

    def __init__(self, srcs: Sequence[torch.Tensor], img_sz=(640, 640), stride: int = 32,
                 letterbox: bool = True, scaleup: bool = True, auto: bool = False,
                 upload: bool = True, mode: str = "letterbox"):
        """mode "letterbox": utils/image_proc.py `resize`; "dataset": the eval loader's
        load_resized_image + PadIfNeeded (stride / letterbox / scaleup / auto unused)."""
        if mode not in ("letterbox", "dataset"):
            raise ValueError(f"mode {mode!r}")
        if isinstance(img_sz, int):
            img_sz = (img_sz, img_sz)
        self.batch = len(srcs)
        self.geoms = []
        self.descs = (VkLbDesc * max(self.batch, 1))()
        out_hw = None
        for i, s in enumerate(srcs):
            _lib.require_cuda(s, "letterbox source")
            if s.dtype != torch.uint8 or s.dim() != 3 or s.shape[2] != 3 or s.stride(2) != 1 or s.stride(1) != 3:
                raise ValueError("letterbox source must be a uint8 HWC tensor with packed pixels")
            h, w = int(s.shape[0]), int(s.shape[1])
            g = (letterbox_geometry(h, w, img_sz, stride, letterbox, scaleup, auto) if mode == "letterbox"
                 else dataset_geometry(h, w, img_sz))
            if out_hw is None:
                out_hw = (g.out_h, g.out_w)
            elif out_hw != (g.out_h, g.out_w):
                raise ValueError("all images of a batch must share one canvas size (auto=True gives "
                                 f"{(g.out_h, g.out_w)} vs {out_hw})")
            d = self.descs[i]
            d.src, d.pitch = s.data_ptr(), int(s.stride(0))
            d.src_h, d.src_w, d.new_h, d.new_w, d.top, d.left = h, w, g.new_h, g.new_w, g.top, g.left
            self.geoms.append(g)
        self.out_h, self.out_w = out_hw if out_hw else (int(img_sz[0]), int(img_sz[1]))
        self._srcs = list(srcs)          # keep the sources alive
        dev = srcs[0].device if srcs else torch.device("cuda")
        nbytes = _lib.lib().vk_letterbox_workspace_bytes(max(self.batch, 1), self.out_h, self.out_w)
        self.ws = torch.empty(nbytes, dtype=torch.uint8, device=dev)
        self.descs_dev = None
        if upload and self.batch:
            host = torch.frombuffer(bytearray(bytes(self.descs)), dtype=torch.uint8)[: self.batch * C.sizeof(VkLbDesc)]
            self.descs_dev = host.to(dev)

    def ratio_pads(self):
        """[(ratio, (dw, dh))] as the reference returns them (image_proc.py:55,60)."""
        return [(g.ratio, (g.pad_w, g.pad_h)) for g in self.geoms]

    def run(self, out: torch.Tensor, swap_rb: bool = False, color=(114, 114, 114)) -> torch.Tensor:
        fmt = _FMT[out.dtype]
        shape = ((self.batch, self.out_h, self.out_w, 3) if fmt == VK_LB_U8_NHWC
                 else (self.batch, 3, self.out_h, self.out_w))
        if tuple(out.shape) != shape or not out.is_contiguous() or not out.is_cuda:
            raise ValueError(f"letterbox output must be a contiguous CUDA tensor of shape {shape}")
        _lib.check("vk_letterbox_batch", _lib.lib().vk_letterbox_batch(
            C.cast(self.descs, C.c_void_p), _ptr(self.descs_dev), self.batch, self.out_h, self.out_w,
            int(bool(swap_rb)), pack_color(color), fmt, _ptr(out), _ptr(self.ws), self.ws.numel(),
            _lib.stream_ptr()))
        return out


def dataset_batch(srcs: Sequence[torch.Tensor], img_sz=(640, 640), color=(114, 114, 114), dtype=torch.float32):
    """Eval ingest of one batch: RGB uint8 HWC CUDA sources of any size ->
    what `validation_step` feeds the model, `(B,3,H,W)` float32/bf16 in [0,1]
    (load_resized_image -> PadIfNeeded -> collate -> permute/float//255, core/train/det_trainer.py:72-75),
    or the (B,H,W,3) uint8 batch the collate produces for dtype=torch.uint8.
    Returns (tensor, [(orig_h, orig_w)], [(resized_h, resized_w)]) like the loader's shapes."""
    plan = LetterboxPlan(srcs, img_sz, upload=False, mode="dataset")
    dev = srcs[0].device
    shape = (plan.batch, plan.out_h, plan.out_w, 3) if dtype == torch.uint8 else (plan.batch, 3, plan.out_h, plan.out_w)
    out = torch.empty(shape, dtype=dtype, device=dev)
    plan.run(out, swap_rb=False, color=color)
    return out, [(int(s.shape[0]), int(s.shape[1])) for s in srcs], [(g.new_h, g.new_w) for g in plan.geoms]


def letterbox_batch(srcs: Sequence[torch.Tensor], img_sz=(640, 640), stride: int = 32,
                    letterbox: bool = True, scaleup: bool = True, auto: bool = False,
                    color=(114, 114, 114), swap_rb: bool = False, dtype=torch.float32):
    """uint8 HWC CUDA sources (any sizes) -> (B,3,H,W) ``dtype`` normalised (or (B,H,W,3)
    uint8) + [(ratio, pad)]."""
    plan = LetterboxPlan(srcs, img_sz, stride, letterbox, scaleup, auto, upload=False)
    dev = srcs[0].device
    if dtype == torch.uint8:
        out = torch.empty((plan.batch, plan.out_h, plan.out_w, 3), dtype=dtype, device=dev)
    else:
        out = torch.empty((plan.batch, 3, plan.out_h, plan.out_w), dtype=dtype, device=dev)
    plan.run(out, swap_rb=swap_rb, color=color)
    return out, plan.ratio_pads()


# ---------------------------------------------------------------------------- Detect head
def head_cfg(variant: str, nc: int, anchors_px, strides, grids) -> VkHeadCfg:
    """anchors_px: per level, flat (w0,h0,w1,h1,...) in pixels; grids: [(ny, nx)]."""
    cfg = VkHeadCfg()
    nl = len(grids)
    if nl > VK_MAX_LEVELS:
        raise ValueError(f"{nl} detection levels > {VK_MAX_LEVELS}")
    cfg.variant = {"v5": VK_HEAD_V5, "v7": VK_HEAD_V7}[variant]
    cfg.nl, cfg.nc = nl, int(nc)
    na = len(anchors_px[0]) // 2
    if na > VK_MAX_ANCHORS:
        raise ValueError(f"{na} anchors per level > {VK_MAX_ANCHORS}")
    cfg.na = na
    for l in range(nl):
        cfg.ny[l], cfg.nx[l] = int(grids[l][0]), int(grids[l][1])
        cfg.stride[l] = float(strides[l])
        for k, v in enumerate(anchors_px[l]):
            cfg.anchors[l][k] = float(v)
    return cfg


def head_rows(cfg: VkHeadCfg) -> int:
    r = _lib.lib().vk_head_rows(C.byref(cfg))
    if r < 0:
        _lib.check("vk_head_rows", r)
    return r


def _level_ptrs(levels: Sequence[torch.Tensor], cfg: VkHeadCfg):
    """Validates the head's conv outputs and returns (pointer array, batch, VK_F32 / VK_F16 / VK_BF16)."""
    no = cfg.nc + 5
    arr = (C.c_void_p * VK_MAX_LEVELS)()
    bs = int(levels[0].shape[0])
    dt = levels[0].dtype
    if dt not in _DTYPE:
        raise ValueError(f"Detect levels must be float32, float16 or bfloat16, got {dt}")
    for l, t in enumerate(levels):
        _lib.require_cuda(t, "Detect level")
        if t.dtype != dt or not t.is_contiguous():
            raise ValueError("Detect levels must be contiguous tensors of one dtype")
        if tuple(t.shape) != (bs, cfg.na * no, cfg.ny[l], cfg.nx[l]):
            raise ValueError(f"level {l}: shape {tuple(t.shape)} != {(bs, cfg.na * no, cfg.ny[l], cfg.nx[l])}")
        arr[l] = t.data_ptr()
    return arr, bs, _DTYPE[dt]


def detect_decode(cfg: VkHeadCfg, levels: Sequence[torch.Tensor], want_raw: bool = False):
    """(B, na*no, ny, nx) x nl (float32 / float16 / bfloat16) -> pred (B, rows, no) float32
    [+ raw (B, na, ny, nx, no) float32 per level]."""
    arr, bs, dt = _level_ptrs(levels, cfg)
    no = cfg.nc + 5
    pred = torch.empty((bs, head_rows(cfg), no), dtype=torch.float32, device=levels[0].device)
    raws, raw_arr = None, None
    if want_raw:
        raws = [torch.empty((bs, cfg.na, cfg.ny[l], cfg.nx[l], no), dtype=torch.float32,
                            device=levels[0].device) for l in range(cfg.nl)]
        raw_arr = (C.c_void_p * VK_MAX_LEVELS)()
        for l, t in enumerate(raws):
            raw_arr[l] = t.data_ptr()
    _lib.check("vk_detect_decode", _lib.lib().vk_detect_decode(
        C.byref(cfg), C.cast(arr, C.c_void_p), dt, bs, _ptr(pred),
        C.cast(raw_arr, C.c_void_p) if raw_arr is not None else C.c_void_p(0), _lib.stream_ptr()))
    return (pred, raws) if want_raw else pred


# ---------------------------------------------------------------------------- candidates
@dataclass
class CandBuf:
    """Caller-owned device buffers of one candidate set (include/vk_b200.h `VkCandBuf`)."""
    cand: torch.Tensor       # int64 (B, cap): low 32 = score bits, high 32 = row*nc + cls; segment s owns slots [s*T, (s+1)*T)
    boxes: torch.Tensor      # float32 (B, rows, 4)
    ctrl: torch.Tensor       # int32 (4 * B): counts | flags | list entries | bound
    seg_count: torch.Tensor  # int32 (B, segs)
    list: Optional[torch.Tensor]   # int64 (B, list_cap): ordered score << 32 | ~(row*nc + cls), unordered; None = no selection pass
    cap: int
    rows: int
    segs: int
    nc: int
    list_cap: int

    @staticmethod
    def alloc(batch: int, rows: int, segs: int, nc: int, cap: int, device, top_list: bool = False,
              list_cap: int = LIST_CAP) -> "CandBuf":
        """top_list=True adds the list of vk_nms_batched's selection pass: for buffers that will hold more
        candidates per image than a stage of the NMS kernel (eval thresholds)."""
        list_cap = int(list_cap) if top_list else 0
        return CandBuf(torch.empty((batch, cap), dtype=torch.int64, device=device),
                       torch.empty((batch, rows, 4), dtype=torch.float32, device=device),
                       torch.zeros((VK_CTRL_WORDS * batch,), dtype=torch.int32, device=device),
                       torch.empty((batch, segs), dtype=torch.int32, device=device),
                       torch.empty((batch, list_cap), dtype=torch.int64, device=device) if list_cap > 0 else None,
                       int(cap), int(rows), int(segs), int(nc), list_cap)

    def c_struct(self) -> VkCandBuf:
        s = VkCandBuf()
        s.cand, s.boxes, s.ctrl = self.cand.data_ptr(), self.boxes.data_ptr(), self.ctrl.data_ptr()
        s.seg_count = self.seg_count.data_ptr()
        s.list = self.list.data_ptr() if self.list is not None else None
        s.cap, s.rows, s.segs, s.nc, s.list_cap = self.cap, self.rows, self.segs, self.nc, self.list_cap
        return s

    @property
    def batch(self) -> int:
        return int(self.seg_count.shape[0])

    @property
    def counts(self) -> torch.Tensor:
        """int32 (B,): candidates per image."""
        return self.ctrl[: self.batch]

    @property
    def tile_slots(self) -> torch.Tensor:
        """int32 (B,): slots each 64-row tile owns, as the filter kernel recorded it."""
        return (self.ctrl[self.batch: 2 * self.batch] >> 8) << 6


def class_mask(classes, nc: int, device) -> Optional[torch.Tensor]:
    """Bitmap for the `classes` filter (utils/image_proc.py:150-151); None = keep all."""
    if classes is None:
        return None
    words = np.zeros(((nc + 31) // 32,), np.uint32)
    for c in classes:
        c = int(c)
        if 0 <= c < nc:
            words[c >> 5] |= np.uint32(1 << (c & 31))
    return torch.from_numpy(words.view(np.int32)).to(device)


def default_cap(segs: int, nc: int, multi_label: bool) -> int:
    """Slots per image: every 64-row tile owns a fixed range (include/vk_b200.h)."""
    return segs * _lib.lib().vk_cand_tile_slots(int(nc), int(bool(multi_label)))


def filter_pred(pred: torch.Tensor, conf_thres: float, multi_label: bool = False, classes=None,
                cap: Optional[int] = None, buf: Optional[CandBuf] = None, kernel="auto") -> CandBuf:
    """Candidates of an existing (B, rows, 5+nc) prediction tensor (float32 / float16 / bfloat16)."""
    _lib.require_cuda(pred, "prediction")
    if pred.dtype not in _DTYPE or not pred.is_contiguous() or pred.dim() != 3:
        raise ValueError("prediction must be a contiguous float32/float16/bfloat16 (B, rows, 5+nc) tensor")
    bs, rows, no = (int(v) for v in pred.shape)
    nc = no - 5
    segs = _lib.lib().vk_filter_segments(rows)
    if buf is None:
        buf = CandBuf.alloc(bs, rows, segs, nc, cap or default_cap(segs, nc, multi_label), pred.device,
                            top_list=expects_dense(kernel, conf_thres))
    mask = class_mask(classes, nc, pred.device)
    cs = buf.c_struct()
    _lib.check("vk_filter_pred", _lib.lib().vk_filter_pred(
        _ptr(pred), _DTYPE[pred.dtype], bs, rows, nc, float(conf_thres), int(bool(multi_label)), _ptr(mask),
        _KERNEL[kernel], C.byref(cs), _lib.stream_ptr()))
    buf._mask = mask
    return buf


def decode_filter(cfg: VkHeadCfg, levels: Sequence[torch.Tensor], conf_thres: float,
                  multi_label: bool = False, classes=None, cap: Optional[int] = None,
                  buf: Optional[CandBuf] = None, kernel="auto") -> CandBuf:
    arr, bs, dt = _level_ptrs(levels, cfg)
    rows = head_rows(cfg)
    segs = _lib.lib().vk_decode_filter_segments(C.byref(cfg))
    if buf is None:
        buf = CandBuf.alloc(bs, rows, segs, cfg.nc, cap or default_cap(segs, cfg.nc, multi_label),
                            levels[0].device, top_list=expects_dense(kernel, conf_thres))
    mask = class_mask(classes, cfg.nc, levels[0].device)
    cs = buf.c_struct()
    _lib.check("vk_decode_filter", _lib.lib().vk_decode_filter(
        C.byref(cfg), C.cast(arr, C.c_void_p), dt, bs, float(conf_thres), int(bool(multi_label)),
        _ptr(mask), _KERNEL[kernel], C.byref(cs), _lib.stream_ptr()))
    buf._mask = mask
    return buf


def conv_decode_filter(cfg: VkHeadCfg, feats: Sequence[torch.Tensor], weights: Sequence[torch.Tensor],
                       biases: Optional[Sequence[Optional[torch.Tensor]]], conf_thres: float,
                       multi_label: bool = False, classes=None, buf: Optional[CandBuf] = None,
                       persistent: bool = True) -> CandBuf:
    """vk_conv_decode_filter: Detect 1x1 conv (tcgen05, TF32) + decode + filter in one kernel.
    feats[l] (B, cin, ny, nx) float32; weights[l] (na*no, cin) or (na*no, cin, 1, 1); biases[l] (na*no) or None.
    persistent=True (default): the warp-specialised kernel on CTA pairs (tcgen05 cta_group::2);
    False: the one-tile-per-CTA kernel (identical results, ~15% slower)."""
    nl = cfg.nl
    bs = int(feats[0].shape[0])
    dev = feats[0].device
    keep = []
    fp = (C.c_void_p * VK_MAX_LEVELS)(); wp = (C.c_void_p * VK_MAX_LEVELS)(); bp = (C.c_void_p * VK_MAX_LEVELS)()
    cin = (C.c_int32 * VK_MAX_LEVELS)()
    for l in range(nl):
        f = feats[l]
        _lib.require_cuda(f, "features")
        if f.dtype != torch.float32 or not f.is_contiguous() or f.dim() != 4 or int(f.shape[0]) != bs or \
                (int(f.shape[2]), int(f.shape[3])) != (cfg.ny[l], cfg.nx[l]):
            raise ValueError(f"feats[{l}] must be a contiguous float32 (B, cin, {cfg.ny[l]}, {cfg.nx[l]}) tensor")
        w = weights[l].detach().reshape(weights[l].shape[0], -1).contiguous().float()
        if tuple(w.shape) != (cfg.na * (cfg.nc + 5), int(f.shape[1])):
            raise ValueError(f"weights[{l}] must have shape ({cfg.na * (cfg.nc + 5)}, {int(f.shape[1])})")
        bb = None if biases is None or biases[l] is None else biases[l].detach().contiguous().float()
        keep += [w, bb]
        fp[l], wp[l], bp[l], cin[l] = f.data_ptr(), w.data_ptr(), (bb.data_ptr() if bb is not None else None), int(f.shape[1])
    rows = head_rows(cfg)
    segs = _lib.lib().vk_decode_filter_segments(C.byref(cfg))
    if buf is None:
        buf = CandBuf.alloc(bs, rows, segs, cfg.nc, default_cap(segs, cfg.nc, multi_label), dev,
                            top_list=float(conf_thres) < 0.05)
    mask = class_mask(classes, cfg.nc, dev)
    fault = torch.zeros(1, dtype=torch.int32, device=dev)
    cs = buf.c_struct()
    _lib.check("vk_conv_decode_filter", _lib.lib().vk_conv_decode_filter(
        C.byref(cfg), C.cast(fp, C.c_void_p), C.cast(cin, C.c_void_p), C.cast(wp, C.c_void_p), C.cast(bp, C.c_void_p),
        bs, float(conf_thres), int(bool(multi_label)), _ptr(mask), VK_CONV_PERSISTENT if persistent else VK_CONV_TILE,
        C.byref(cs), _ptr(fault), _lib.stream_ptr()))
    buf._mask, buf._keep, buf.fault = mask, keep, fault
    return buf


@dataclass
class NmsOut:
    dets: torch.Tensor        # float32 (B, max_det, 6), rows >= count zero
    counts: torch.Tensor      # int32 (B,)
    keep: Optional[torch.Tensor]   # int64 (B, max_det), -1 padded
    status: torch.Tensor      # int32 (B,), always 0 (a candidate buffer sized by default_cap cannot overflow)
    fault: Optional[torch.Tensor] = None   # int32 (1,) of the fused conv head: 1 = a tensor-core wait timed out

    def check(self) -> None:
        """Raises if the fused conv head reported a fault (synchronises; `to_list` calls it)."""
        if self.fault is not None and int(self.fault.item()) != 0:
            raise RuntimeError("vk_conv_decode_filter: a tensor-core completion wait timed out; detections are incomplete")


def nms_batched(buf: CandBuf, iou_thres: float, agnostic: bool = False, max_nms: int = 30000,
                max_det: int = 300, max_wh: float = MAX_WH, want_keep: bool = False,
                out: Optional[NmsOut] = None) -> NmsOut:
    dev = buf.cand.device
    bs = buf.batch
    if out is None:
        out = NmsOut(torch.empty((bs, max_det, 6), dtype=torch.float32, device=dev),
                     torch.empty((bs,), dtype=torch.int32, device=dev),
                     torch.empty((bs, max_det), dtype=torch.int64, device=dev) if want_keep else None,
                     torch.empty((bs,), dtype=torch.int32, device=dev))
    out.fault = getattr(buf, "fault", None)
    cs = buf.c_struct()
    _lib.check("vk_nms_batched", _lib.lib().vk_nms_batched(
        C.byref(cs), bs, float(iou_thres), int(bool(agnostic)), int(max_nms), int(max_det),
        float(max_wh), _ptr(out.dets), _ptr(out.counts), _ptr(out.keep), _ptr(out.status), _lib.stream_ptr()))
    return out


# ---------------------------------------------------------------------------- small ops
def scale_coords_(coords: torch.Tensor, pad_w: float, pad_h: float, gain: float,
                  subtract_pad: bool = True, clip_hw=None) -> torch.Tensor:
    """In place on a (n, >=4) float32 CUDA tensor/view whose rows are `stride(0)` floats apart."""
    _lib.require_cuda(coords, "coords")
    if coords.dtype != torch.float32 or coords.dim() != 2 or coords.shape[1] < 4 or \
            (coords.shape[0] > 0 and coords.stride(1) != 1):
        raise ValueError("coords must be a float32 (n, >=4) tensor with unit column stride")
    n = int(coords.shape[0])
    cw, ch = (-1.0, -1.0) if clip_hw is None else (float(clip_hw[1]), float(clip_hw[0]))
    _lib.check("vk_scale_coords", _lib.lib().vk_scale_coords(
        _ptr(coords), n, int(coords.stride(0)) if n else 4, float(pad_w), float(pad_h), float(gain),
        int(bool(subtract_pad)), cw, ch, _lib.stream_ptr()))
    return coords


def cxcywh_to_xyxy(b: torch.Tensor) -> torch.Tensor:
    _lib.require_cuda(b, "boxes")
    src = b.contiguous().float().view(-1, 4)
    out = torch.empty_like(src)
    _lib.check("vk_cxcywh_to_xyxy", _lib.lib().vk_cxcywh_to_xyxy(
        _ptr(src), _ptr(out), int(src.shape[0]), _lib.stream_ptr()))
    return out.view(b.shape)


# ---------------------------------------------------------------------------- evaluator matching
@dataclass
class EvalMatch:
    predn: torch.Tensor       # float32 (B, max_det, 6): detections in original-image pixels, clipped
    labeln: torch.Tensor      # float32 (n_labels, 5): cls, x1, y1, x2, y2 in original-image pixels
    correct: torch.Tensor     # bool (B, max_det, niou): true-positive matrix (rows >= count are False)


def eval_match(dets: torch.Tensor, counts: torch.Tensor, labels: torch.Tensor, label_offsets: torch.Tensor,
               max_labels: int, img0_hw, img1_hw, iouv: torch.Tensor, prescaled: bool = False) -> EvalMatch:
    """vk_eval_match: un-letterbox + box_iou + process_batch for a batch in one launch.
    dets (B, max_det, 6) / counts (B,) = NMS output; labels (n, 6) = image, cls, cx, cy, w, h in canvas
    pixels grouped by image; label_offsets int32 (B+1,); img0_hw int32 (B, 2).
    prescaled: process_batch alone -- labels carry x1, y1, x2, y2, nothing is rescaled."""
    _lib.require_cuda(dets, "dets")
    B, max_det = int(dets.shape[0]), int(dets.shape[1])
    dev = dets.device
    dets = dets.contiguous().float()
    labels = labels.contiguous().float().view(-1, 6)
    iouv = iouv.to(device=dev, dtype=torch.float32).contiguous()
    niou = int(iouv.numel())
    predn = torch.empty((B, max_det, 6), dtype=torch.float32, device=dev)
    labeln = torch.empty((labels.shape[0], 5), dtype=torch.float32, device=dev)
    correct = torch.empty((B, max_det, niou), dtype=torch.uint8, device=dev)
    _lib.check("vk_eval_match", _lib.lib().vk_eval_match(
        _ptr(dets), _ptr(counts), B, max_det, _ptr(labels) if labels.numel() else C.c_void_p(0),
        _ptr(label_offsets), int(max_labels), _ptr(img0_hw) if img0_hw is not None else C.c_void_p(0),
        int(img1_hw[0]), int(img1_hw[1]), int(bool(prescaled)), _ptr(iouv), niou, _ptr(predn), _ptr(labeln) if labels.numel() else C.c_void_p(0), _ptr(correct),
        _lib.stream_ptr()))
    return EvalMatch(predn, labeln, correct.bool())
