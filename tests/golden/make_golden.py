"""Generates tests/golden/*.npz by running the UNMODIFIED reference
(/root/reference, imported live through oracle/live.py) on seeded synthetic
inputs.  Dev-container only -- the GPU box has no reference tree; the fixtures
it writes are committed and travel instead.

    python tests/golden/make_golden.py

Inputs are regenerated from seeds by vision_kit_b200/synth.py (bit-exact on
any host), so the fixtures store outputs only, except where noted.
"""
from __future__ import annotations

import hashlib
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from oracle import live                      # noqa: E402
from tests import synth                      # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))

# (name, src_h, src_w, img_sz, kwargs, store_full)
LETTERBOX_CASES = [
    ("s_down_wide", 45, 80, (64, 64), {}, True),
    ("s_down_tall", 101, 67, (64, 64), {}, True),
    ("s_exact2x", 128, 128, (64, 64), {}, True),
    ("s_up", 20, 30, (64, 64), {}, True),
    ("s_identity", 64, 64, (64, 64), {}, True),
    ("s_rect_target", 33, 47, (96, 64), {}, True),
    ("s_auto", 100, 37, (64, 64), {"auto": True}, True),
    ("s_noletterbox", 45, 80, (64, 64), {"letterbox": False}, True),
    ("s_noscaleup", 20, 30, (64, 64), {"scaleup": False}, True),
    ("s_big_down", 300, 200, (64, 64), {}, True),
    ("s_int_sz", 50, 70, 64, {}, True),
    ("s_color", 45, 80, (64, 64), {"color": (0, 128, 255)}, True),
    ("l_bus_shape", 1080, 810, (640, 640), {}, False),
    ("l_cat_shape", 375, 500, (640, 640), {}, False),
    ("l_zidane_shape", 720, 1280, (640, 640), {}, False),
    ("l_identity", 640, 640, (640, 640), {}, False),
    ("l_up480", 480, 480, (640, 640), {}, False),
    ("l_exact2x", 1280, 1280, (640, 640), {}, False),
    ("l_odd", 1279, 853, (640, 640), {}, False),
    ("l_huge", 2000, 3000, (640, 640), {}, False),
    ("l_auto", 1080, 810, (640, 640), {"auto": True}, False),
]

# (name, rows, batch, mode, clusters, kwargs)
NMS_CASES = [
    ("s_default", 252, 3, "demo", 4, {}),
    ("s_multi", 252, 3, "eval", 4, {"conf_thres": 0.001, "iou_thres": 0.6, "multi_label": True}),
    ("s_agnostic", 252, 3, "demo", 4, {"agnostic": True}),
    ("s_classes", 252, 3, "demo", 4, {"classes": [1, 3, 5, 7, 11]}),
    ("s_maxdet", 252, 3, "eval", 4, {"conf_thres": 0.01, "multi_label": True, "max_det": 5}),
    ("s_conf0", 252, 2, "demo", 2, {"conf_thres": 0.0, "iou_thres": 0.3}),
    ("l_demo", 25200, 2, "demo", 30, {}),
    ("l_eval", 25200, 2, "eval", 30, {"conf_thres": 0.001, "iou_thres": 0.6, "multi_label": True}),
    ("l_eval_agnostic", 25200, 1, "eval", 30,
     {"conf_thres": 0.001, "iou_thres": 0.6, "multi_label": True, "agnostic": True}),
]


def sha(a: np.ndarray) -> str:
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def case_seed(name: str) -> int:
    return int(hashlib.sha256(name.encode()).hexdigest()[:8], 16)


def gen_letterbox(ref):
    blob, meta = {}, {}
    for name, h, w, sz, kw, full in LETTERBOX_CASES:
        img = synth.image_u8(h, w, case_seed(name))
        out, (ratio, pad) = ref.image_proc.resize(sz, img.copy(), **kw)
        ip = ref.ImageProcessor(img_sz=sz, **kw)
        out2, (ratio2, pad2) = ip.resize(img.copy())
        assert np.array_equal(out, out2) and ratio == ratio2 and tuple(pad) == tuple(pad2)
        ten, _ = ip.preprocess(img.copy(), is_BGR=True)
        meta[name] = dict(h=h, w=w, img_sz=sz, kw=kw, shape=list(out.shape), sha=sha(out),
                          ratio=float(ratio), pad=[float(pad[0]), float(pad[1])],
                          pre_sha=sha(ten.numpy()), pre_shape=list(ten.shape))
        if full:
            blob[name] = out
    for asset in ("bus", "cat", "zidane"):
        import cv2
        img = cv2.imread(os.path.join(live.REF_ROOT, "assets", asset + ".jpg"))
        ip = ref.ImageProcessor(auto=False)
        out, (ratio, pad) = ip.resize(img.copy())
        ten, _ = ip.preprocess(img.copy())
        meta["asset_" + asset] = dict(h=img.shape[0], w=img.shape[1], src_sha=sha(img),
                                      shape=list(out.shape), sha=sha(out), ratio=float(ratio),
                                      pad=[float(pad[0]), float(pad[1])], pre_sha=sha(ten.numpy()))
    np.savez_compressed(os.path.join(OUT, "letterbox.npz"), **blob)
    with open(os.path.join(OUT, "letterbox.json"), "w") as f:
        json.dump(meta, f, indent=1, sort_keys=True)


def gen_decode():
    blob = {}
    for variant in ("v5", "v7"):
        lv = [torch.from_numpy(x) for x in synth.head_logits(2, seed=11, img=64)]
        pred, raws = live.head_decode(variant, lv)
        blob[f"{variant}_small_pred"] = pred.numpy()
        blob[f"{variant}_small_raw0"] = raws[0].numpy()
        lv = [torch.from_numpy(x) for x in synth.head_logits(1, seed=12, img=640)]
        pred, _ = live.head_decode(variant, lv)
        blob[f"{variant}_640_rows"] = pred.numpy()[:, ::97].copy()
    np.savez_compressed(os.path.join(OUT, "decode.npz"), **blob)


def gen_nms(ref):
    blob, meta = {}, {}
    for name, rows, batch, mode, clusters, kw in NMS_CASES:
        p = torch.from_numpy(synth.prediction(batch, rows, seed=case_seed(name), mode=mode,
                                              clusters=clusters, img=64 if rows == 252 else 640))
        before = p.clone()
        with live.capture_keep() as cap:
            out = ref.image_proc.nms(p, **kw)
        assert torch.equal(p, before), "reference mutated its input"
        ki = 0
        ns = []
        for i, d in enumerate(out):
            blob[f"{name}_dets{i}"] = d.numpy()
            if d.shape[0]:
                blob[f"{name}_keep{i}"] = cap.keeps[ki][: kw.get("max_det", 300)].numpy()
                ns.append(int(d.shape[0]))
                ki += 1
            else:
                blob[f"{name}_keep{i}"] = np.zeros((0,), np.int64)
                ns.append(0)
        meta[name] = dict(rows=rows, batch=batch, mode=mode, clusters=clusters, kw=kw, counts=ns)
        # demo ImageProcessor copy (max_nms = 10000), demo/processing.py:107-199
        ipkw = dict(conf_thres=kw.get("conf_thres", 0.25), iou_thres=kw.get("iou_thres", 0.45),
                    filtered_classes=kw.get("classes"), agnostic=kw.get("agnostic", False),
                    multi_label=kw.get("multi_label", False), max_det=kw.get("max_det", 300))
        ip = ref.ImageProcessor(**ipkw)
        out2 = ip.nms(p)
        for i, d in enumerate(out2):
            blob[f"{name}_ipdets{i}"] = d.numpy()
    np.savez_compressed(os.path.join(OUT, "nms.npz"), **blob)
    with open(os.path.join(OUT, "nms.json"), "w") as f:
        json.dump(meta, f, indent=1, sort_keys=True)


def gen_scale_coords(ref):
    blob = {}
    rng = np.random.Generator(np.random.PCG64(5))
    for name, img0 in (("bus", (1080, 810)), ("wide", (375, 1242)), ("same", (640, 640))):
        c = (rng.random((40, 6), dtype=np.float32) * np.float32(700) - np.float32(30))
        t = torch.from_numpy(c.copy())
        ret = ref.image_proc.scale_coords((640, 640), t[:, :4], img0)
        blob[f"{name}_in"] = c
        blob[f"{name}_inplace"] = t.numpy()
        blob[f"{name}_ret"] = ret.numpy()
        ip = ref.ImageProcessor()
        img = synth.image_u8(img0[0], img0[1], 3)
        ip.resize(img)
        t2 = torch.from_numpy(c.copy())
        blob[f"{name}_demo"] = ip.scale_coords(t2).numpy()
    np.savez_compressed(os.path.join(OUT, "scale_coords.npz"), **blob)


def gen_eval():
    """``DetEvaluator.evaluate`` of the unmodified reference (core/eval/det_evaluator.py:129-182):
    stats = (correct, conf, pred_cls, target_cls) per image, plus the vstacked predn / targetn."""
    from vision_kit.core.eval.det_evaluator import DetEvaluator
    blob = {}
    for name, n_img, canvas, _, _ in synth.EVAL_CASES:
        preds, targets, shapes, _ = synth.eval_inputs(name)
        ev = DetEvaluator([str(i) for i in range(5)], img_size=canvas)
        img = torch.zeros((n_img, 3, canvas[0], canvas[1]))
        pn, tn = ev.evaluate(img, shapes, list(range(n_img)), [torch.from_numpy(p.copy()) for p in preds],
                             torch.from_numpy(targets.copy()))
        blob[f"{name}_predn"] = pn.numpy()
        blob[f"{name}_targetn"] = tn.numpy()
        blob[f"{name}_nstats"] = np.int64(len(ev.stats))
        for j, st in enumerate(ev.stats):
            blob[f"{name}_correct{j}"] = st[0].numpy()
            blob[f"{name}_conf{j}"] = st[1].numpy()
            blob[f"{name}_pcls{j}"] = st[2].numpy()
            blob[f"{name}_tcls{j}"] = st[3].numpy()
        import warnings
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")                      # np.trapz deprecation
            m50, m95, _, _ = ev.summarize()
        blob[f"{name}_map"] = np.asarray([m50, m95, ev.mp, ev.mr], np.float64)
        blob[f"{name}_prf"] = np.stack([ev.precision, ev.recall, ev.f1]).astype(np.float64)
    np.savez_compressed(os.path.join(OUT, "eval.npz"), **blob)


if __name__ == "__main__":
    assert live.available(), "needs /root/reference"
    torch.manual_seed(0)
    ref = live.load()
    gen_letterbox(ref)
    gen_decode()
    gen_nms(ref)
    gen_scale_coords(ref)
    gen_eval()
    for f in sorted(os.listdir(OUT)):
        print(f, os.path.getsize(os.path.join(OUT, f)))
