"""Batched device pipeline of the hot path with pre-allocated buffers: what a serving or
eval loop calls once per batch.

    pipe = DetectPipeline("v5", batch=64)
    x = pipe.preprocess(srcs)            # uint8 HWC CUDA sources -> (B,3,640,640) input tensor
    feats = model_backbone_and_convs(x)  # the reference's PyTorch modules (not part of this repo)
    out = pipe.postprocess(feats)        # fused decode + filter + NMS -> NmsOut (device tensors)
    dets = pipe.to_list(out)             # the reference's list[(k,6)] (one synchronisation)

Replaces, per batch, demo/processing.py:45-52 + models/heads/*.py eval forward +
utils/image_proc.py:83-187, with no host synchronisation between the kernels.

A loop whose buffers stay put (a staging buffer for the sources, the model's static output
tensors) can `capture()` the step once and `replay()` it: one CUDA-graph launch per batch
instead of three C calls and their stream bookkeeping.
"""
from __future__ import annotations

import ctypes as C
from typing import List, Optional, Sequence

import torch

from . import _lib, ops

# default anchors of the reference's heads (models/heads/yolov5.py:24-28, yolov7.py:23-27), pixels
V5_ANCHORS = [[10, 13, 16, 30, 33, 23], [30, 61, 62, 45, 59, 119], [116, 90, 156, 198, 373, 326]]
V7_ANCHORS = [[12, 16, 19, 36, 40, 28], [36, 75, 76, 55, 72, 146], [142, 110, 192, 243, 459, 401]]


class DetectPipeline:
    def __init__(self, variant: str = "v5", nc: int = 80, img_sz=(640, 640), batch: int = 64,
                 conf_thres: float = 0.25, iou_thres: float = 0.45, classes=None, agnostic: bool = False,
                 multi_label: bool = False, max_det: int = 300, max_nms: int = 30000,
                 anchors=None, strides=(8, 16, 32), dtype=torch.float32, color=(114, 114, 114),
                 swap_rb: bool = True, device=None, cand_cap: Optional[int] = None, want_keep: bool = False,
                 overlap: bool = False, filter_kernel="auto", list_cap: int = ops.LIST_CAP,
                 fork_preprocess: bool = False, nvtx: bool = False, nms_fork: str = "auto",
                 side_priority: int = 0):
        self.device = torch.device(device if device is not None else f"cuda:{torch.cuda.current_device()}")
        if isinstance(img_sz, int):
            img_sz = (img_sz, img_sz)
        self.img_sz, self.batch, self.dtype, self.color, self.swap_rb = tuple(img_sz), batch, dtype, color, swap_rb
        self.conf_thres, self.iou_thres, self.classes = conf_thres, iou_thres, classes
        self.agnostic, self.multi_label, self.max_det, self.max_nms = agnostic, multi_label, max_det, max_nms
        if anchors is None:
            anchors = V5_ANCHORS if variant == "v5" else V7_ANCHORS
        grids = [(img_sz[0] // int(s), img_sz[1] // int(s)) for s in strides]
        self.cfg = ops.head_cfg(variant, nc, anchors, strides, grids)
        self.rows = ops.head_rows(self.cfg)
        segs = _lib.lib().vk_decode_filter_segments(C.byref(self.cfg))
        ml = bool(multi_label) and nc > 1
        cap = cand_cap or ops.default_cap(segs, nc, ml)
        self._kernel = ops._KERNEL[filter_kernel]
        # two sets of candidate / output buffers: with overlap=True the NMS of batch k runs on a
        # side stream while the letterbox and filter kernels of batch k+1 run on the main one
        # nvtx=True brackets the three C calls with NVTX ranges (vk/letterbox, vk/decode_filter, vk/nms) so that a
        # timeline tool shows where a batch is inside the pipeline; off by default (two calls per range)
        self._nvtx = bool(nvtx)
        self.overlap = bool(overlap)
        # captured graphs only: the letterbox on a third branch, beside the filter (they share no data)
        self.fork_preprocess = bool(fork_preprocess)
        # captured overlapped graphs: where the NMS branch of the previous batch forks off.  "start": beside the letterbox
        # and the filter; "after_preprocess": beside the filter only.  The copy-peak letterbox loses more to NMS CTAs on
        # its SMs than the sector-rate-bound sparse filter does (config 2: 117.6 -> 113.4 us per step), while the eval
        # NMS's select pass is better started early (config 3: 271 vs 292 us).
        if nms_fork not in ("auto", "start", "after_preprocess"):
            raise ValueError(f"nms_fork={nms_fork!r}")
        if nms_fork == "auto":
            nms_fork = "start" if ops.expects_dense(filter_kernel, conf_thres) else "after_preprocess"
        self.nms_fork = nms_fork
        nsets = 2 if self.overlap else 1
        if agnostic and list_cap == ops.LIST_CAP:
            list_cap *= 2        # agnostic NMS goes deeper: most candidates it meets are other classes of rows already decided
        self._cands = [ops.CandBuf.alloc(batch, self.rows, segs, nc, cap, self.device, list_cap=list_cap,
                                         top_list=ops.expects_dense(filter_kernel, conf_thres)) for _ in range(nsets)]
        self._outs = [ops.NmsOut(
            torch.empty((batch, max_det, 6), dtype=torch.float32, device=self.device),
            torch.empty((batch,), dtype=torch.int32, device=self.device),
            torch.empty((batch, max_det), dtype=torch.int64, device=self.device) if want_keep else None,
            torch.empty((batch,), dtype=torch.int32, device=self.device)) for _ in range(nsets)]
        self._set = 0
        self.cand, self.out = self._cands[0], self._outs[0]
        if self.fork_preprocess:
            self.pre_stream = torch.cuda.Stream(device=self.device)
        if self.overlap:
            self.side = torch.cuda.Stream(device=self.device, priority=int(side_priority))
            self._ev_filter = [torch.cuda.Event() for _ in range(2)]
            self._ev_nms = [torch.cuda.Event() for _ in range(2)]
            self._nms_pending = [False, False]
        self.input = torch.empty((batch, 3, img_sz[0], img_sz[1]), dtype=dtype, device=self.device)
        self.plan: Optional[ops.LetterboxPlan] = None
        # Pre-marshalled ctypes arguments: a step of the loop is three C calls with constant arguments
        self._lib = _lib.lib()
        self._mask = ops.class_mask(classes, nc, self.device)
        self._mask_p = ops._ptr(self._mask)
        self._cs = [c.c_struct() for c in self._cands]
        self._cfg_ref = C.byref(self.cfg)
        self._ml = int(ml)
        self._conf = C.c_float(float(conf_thres))
        self._nms_args = [(C.byref(cs), batch, C.c_double(float(iou_thres)), int(bool(agnostic)),
                           int(max_nms), int(max_det), C.c_float(ops.MAX_WH), ops._ptr(o.dets), ops._ptr(o.counts),
                           ops._ptr(o.keep), ops._ptr(o.status))
                          for cs, o in zip(self._cs, self._outs)]
        self._lv_key = None
        self._lv_arr = None
        self._lv_dt = 0
        self._lb_args = None
        self._graphs: List[torch.cuda.CUDAGraph] = []
        self._graph_feats = None
        self._graph_step = 0

    # -- letterbox + normalise
    def plan_sources(self, srcs: Sequence[torch.Tensor]) -> ops.LetterboxPlan:
        """Builds (and uploads) the descriptors for a set of source buffers.  The plan stays
        valid while those buffers do: a loop that refills a fixed staging buffer plans once."""
        if len(srcs) != self.batch:
            raise ValueError(f"expected {self.batch} sources, got {len(srcs)}")
        self.plan = ops.LetterboxPlan(srcs, self.img_sz, upload=True)
        pl = self.plan
        if (pl.out_h, pl.out_w) != self.img_sz:
            raise ValueError(f"letterbox canvas {(pl.out_h, pl.out_w)} != pipeline image size {self.img_sz}")
        self._lb_args = (C.cast(pl.descs, C.c_void_p), ops._ptr(pl.descs_dev), pl.batch, pl.out_h, pl.out_w,
                         int(bool(self.swap_rb)), ops.pack_color(self.color), ops._FMT[self.input.dtype],
                         ops._ptr(self.input), ops._ptr(pl.ws), C.c_size_t(pl.ws.numel()))
        return self.plan

    def preprocess(self, srcs: Optional[Sequence[torch.Tensor]] = None) -> torch.Tensor:
        if srcs is not None:
            self.plan_sources(srcs)
        if self._nvtx:
            torch.cuda.nvtx.range_push("vk/letterbox")
        rc = self._lib.vk_letterbox_batch(*self._lb_args, C.c_void_p(torch.cuda.current_stream().cuda_stream))
        if self._nvtx:
            torch.cuda.nvtx.range_pop()
        if rc:
            _lib.check("vk_letterbox_batch", rc)
        return self.input

    def ratio_pads(self):
        return self.plan.ratio_pads()

    # -- fused Detect decode + confidence filter + NMS
    def filter(self, feats: Sequence[torch.Tensor]) -> ops.CandBuf:
        if self.overlap:
            self._set ^= 1
            self.cand, self.out = self._cands[self._set], self._outs[self._set]
            if self._nms_pending[self._set]:        # the NMS that last read this buffer set
                torch.cuda.current_stream().wait_event(self._ev_nms[self._set])
        # the marshalled pointer array is reused while the same tensors (address, dtype, shape, layout) come back
        key = tuple((t.data_ptr(), t.dtype, t.shape, t.stride()) for t in feats)
        if key != self._lv_key:
            self._lv_arr, bs, self._lv_dt = ops._level_ptrs(feats, self.cfg)
            if bs != self.batch:
                raise ValueError(f"expected a batch of {self.batch}, got {bs}")
            self._lv_key = key
        if self._nvtx:
            torch.cuda.nvtx.range_push("vk/decode_filter")
        rc = self._lib.vk_decode_filter(self._cfg_ref, C.cast(self._lv_arr, C.c_void_p), self._lv_dt, self.batch,
                                        self._conf, self._ml, self._mask_p, self._kernel, C.byref(self._cs[self._set]),
                                        C.c_void_p(torch.cuda.current_stream().cuda_stream))
        if self._nvtx:
            torch.cuda.nvtx.range_pop()
        if rc:
            _lib.check("vk_decode_filter", rc)
        return self.cand

    def nms(self) -> ops.NmsOut:
        if self._nvtx:
            torch.cuda.nvtx.range_push("vk/nms")
        try:
            return self._nms()
        finally:
            if self._nvtx:
                torch.cuda.nvtx.range_pop()

    def _nms(self) -> ops.NmsOut:
        k = self._set
        if not self.overlap:
            rc = self._lib.vk_nms_batched(*self._nms_args[k], C.c_void_p(torch.cuda.current_stream().cuda_stream))
            if rc:
                _lib.check("vk_nms_batched", rc)
            return self.out
        # the kernels take the stream as an argument: no need to switch torch's current stream
        self._ev_filter[k].record()
        self.side.wait_event(self._ev_filter[k])
        rc = self._lib.vk_nms_batched(*self._nms_args[k], C.c_void_p(self.side.cuda_stream))
        if rc:
            _lib.check("vk_nms_batched", rc)
        self._ev_nms[k].record(self.side)
        self._nms_pending[k] = True
        return self.out

    def join(self) -> None:
        """Makes the current stream wait for every NMS issued on the side stream."""
        if self.overlap:
            for k in range(2):
                if self._nms_pending[k]:
                    torch.cuda.current_stream().wait_event(self._ev_nms[k])

    def postprocess(self, feats: Sequence[torch.Tensor], join: bool = True) -> ops.NmsOut:
        """join=False leaves the NMS running on the side stream (overlap=True): the caller must
        call join() (or wait on the side stream) before reading the returned tensors."""
        self.filter(feats)
        out = self.nms()
        if join:
            self.join()
        return out

    # -- the same step as one CUDA-graph launch
    def capture(self, feats: Sequence[torch.Tensor], with_preprocess: bool = True) -> None:
        """Captures preprocess (optional; needs plan_sources) + filter + NMS on THESE tensors into a CUDA
        graph.  The source buffers of the plan and `feats` must keep their addresses; refill them in place
        between `replay()` calls.  With overlap=True two graphs alternate over the two buffer sets and
        every launch runs the NMS of the previous batch beside the letterbox and filter of the current one."""
        if with_preprocess and self._lb_args is None:
            raise RuntimeError("capture(with_preprocess=True) needs plan_sources() first")
        for _ in range(2):                       # eager warm-up: kernel attributes, validation, caches
            if with_preprocess:
                self.preprocess()
            self.postprocess(feats)
        torch.cuda.synchronize(self.device)
        self._graph_feats = list(feats)          # keep them alive
        self._graphs = []
        if not self.overlap:
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                if with_preprocess:
                    self.preprocess()
                self.filter(feats)
                self.nms()
            self._graphs = [g]
        else:
            # graph s: [letterbox, filter -> set s] beside [NMS of set s^1]
            self._nms_pending = [False, False]
            for s in (0, 1):
                g = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g):
                    main = torch.cuda.current_stream()
                    late = with_preprocess and not self.fork_preprocess and self.nms_fork == "after_preprocess"
                    if late:
                        self.preprocess()
                    self.side.wait_stream(main)
                    rc = self._lib.vk_nms_batched(*self._nms_args[s ^ 1], C.c_void_p(self.side.cuda_stream))
                    if rc:
                        _lib.check("vk_nms_batched", rc)
                    if late:
                        pass
                    elif with_preprocess and self.fork_preprocess:
                        self.pre_stream.wait_stream(main)
                        rc = self._lib.vk_letterbox_batch(*self._lb_args, C.c_void_p(self.pre_stream.cuda_stream))
                        if rc:
                            _lib.check("vk_letterbox_batch", rc)
                    elif with_preprocess:
                        self.preprocess()
                    rc = self._lib.vk_decode_filter(self._cfg_ref, C.cast(self._lv_arr, C.c_void_p), self._lv_dt,
                                                    self.batch, self._conf, self._ml, self._mask_p, self._kernel,
                                                    C.byref(self._cs[s]), C.c_void_p(main.cuda_stream))
                    if rc:
                        _lib.check("vk_decode_filter", rc)
                    main.wait_stream(self.side)
                    if with_preprocess and self.fork_preprocess:
                        main.wait_stream(self.pre_stream)
                self._graphs.append(g)
            self._set = 1                        # the eager warm-up left valid candidates in both sets
        self._graph_step = 0

    def replay(self) -> ops.NmsOut:
        """One launch of the captured step.  Without overlap: the NmsOut of this batch.  With overlap: the
        NmsOut of the PREVIOUS batch (its NMS ran inside this launch); call `flush()` after the last batch."""
        if not self._graphs:
            raise RuntimeError("replay() before capture()")
        if not self.overlap:
            self._graphs[0].replay()
            return self.out
        s = self._graph_step & 1
        self._graph_step += 1
        self._graphs[s].replay()
        self._set = s
        self.cand = self._cands[s]
        self.out = self._outs[s ^ 1]
        return self.out

    def flush(self) -> ops.NmsOut:
        """overlap=True graphs: runs the NMS of the last replayed batch (eagerly, on the current stream)."""
        s = self._set
        rc = self._lib.vk_nms_batched(*self._nms_args[s], C.c_void_p(torch.cuda.current_stream().cuda_stream))
        if rc:
            _lib.check("vk_nms_batched", rc)
        self.out = self._outs[s]
        return self.out

    # -- drop-in layout: decode to the (B, rows, no) tensor, then filter + NMS from it
    def postprocess_materialised(self, feats: Sequence[torch.Tensor]):
        pred = ops.detect_decode(self.cfg, feats)
        buf = ops.filter_pred(pred, self.conf_thres, self.multi_label, self.classes)
        return pred, ops.nms_batched(buf, self.iou_thres, self.agnostic, self.max_nms, self.max_det)

    @staticmethod
    def to_list(out: ops.NmsOut) -> List[torch.Tensor]:
        out.check()
        counts = out.counts.cpu().tolist()
        return [out.dets[i, :k] for i, k in enumerate(counts)]
