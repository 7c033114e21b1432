"""CPU tests of the host side: the C-ABI library loads and exports every symbol the header
declares, the host-only geometry entry point equals the oracle, argument validation, the
division-free /255, and the world_size-2 sharding + all-gather logic over gloo."""
import ctypes as C
import os
import socket
from fractions import Fraction

import numpy as np
import pytest
import torch
import torch.multiprocessing as mp

from oracle import restate


def test_library_exports_every_declared_symbol(vk_lib):
    from vision_kit_b200 import _lib
    names = _lib.declared_symbols()
    assert len(names) >= 17 and set(_lib._PROTOS) <= set(names)
    for n in names:
        assert hasattr(vk_lib, n), n
    assert vk_lib.vk_version() >= 100 and vk_lib.vk_build_arch() == 100
    assert vk_lib.vk_last_error() is not None


def test_struct_layouts_match_header():
    from vision_kit_b200 import _lib
    assert C.sizeof(_lib.VkLbDesc) == 48
    assert C.sizeof(_lib.VkLbGeom) == 64
    assert C.sizeof(_lib.VkHeadCfg) == 4 * 4 + 4 * 4 * 2 + 4 * 4 + 4 * 16 * 4
    assert C.sizeof(_lib.VkCandBuf) == 5 * 8 + 6 * 4


def test_geometry_equals_oracle(vk_lib):
    from vision_kit_b200 import ops
    rng = np.random.Generator(np.random.PCG64(3))
    cases = [(1080, 810, (640, 640)), (375, 500, (640, 640)), (720, 1280, (640, 640)), (640, 640, 640),
             (1, 1, (32, 32)), (3000, 17, (640, 640)), (45, 80, (64, 64))]
    for _ in range(3000):
        cases.append((int(rng.integers(1, 2500)), int(rng.integers(1, 2500)),
                      (int(rng.integers(16, 1500)), int(rng.integers(16, 1500)))))
    for h, w, sz in cases:
        for lb in (True, False):
            for su in (True, False):
                for au in (True, False):
                    e = restate.letterbox_geometry(h, w, sz, 32, lb, su, au)
                    if e["new_w"] < 1 or e["new_h"] < 1:
                        continue
                    g = ops.letterbox_geometry(h, w, sz, 32, lb, su, au)
                    got = (g.ratio, g.pad_w, g.pad_h, g.new_w, g.new_h, g.top, g.bottom, g.left, g.right,
                           g.out_h, g.out_w)
                    exp = (e["ratio"], float(e["pad"][0]), float(e["pad"][1]), e["new_w"], e["new_h"], e["top"],
                           e["bottom"], e["left"], e["right"], e["out_h"], e["out_w"])
                    assert got == exp, (h, w, sz, lb, su, au)


def test_argument_validation_without_gpu(vk_lib):
    from vision_kit_b200 import _lib, ops
    g = _lib.VkLbGeom()
    assert vk_lib.vk_letterbox_geometry(0, 5, 64, 64, 32, 1, 1, 0, C.byref(g)) == -1
    assert b"bad size" in vk_lib.vk_last_error()
    cfg = ops.head_cfg("v5", 80, [[1, 2, 3, 4, 5, 6]] * 3, (8, 16, 32), [(80, 80), (40, 40), (20, 20)])
    assert ops.head_rows(cfg) == 25200
    assert vk_lib.vk_decode_filter_segments(C.byref(cfg)) == 3 * (100 + 25 + 7)
    assert vk_lib.vk_filter_segments(25200) == 394
    cfg.nl = 9
    with pytest.raises(_lib.VkError):
        ops.head_rows(cfg)
    assert vk_lib.vk_letterbox_workspace_bytes(64, 640, 640) == 3072 + 64 * (640 * 8 + 640 * 16 + 80 * 16)   # descs, xtab, ytab, tile table
    # null pointers / out-of-range thresholds are rejected before any launch
    assert vk_lib.vk_nms_batched(None, 1, 0.5, 0, 30000, 300, 7680.0, None, None, None, None, None) == -1
    assert vk_lib.vk_filter_pred(None, 0, 1, 100, 80, 0.25, 0, None, 0, None, None) == -1
    assert vk_lib.vk_scale_coords(None, 0, 4, 0.0, 0.0, 1.0, 1, -1.0, -1.0, None) == 0   # n = 0 is a no-op


def test_shims_refuse_cpu_tensors(vk_lib):
    from vision_kit_b200 import bboxes, image_proc
    with pytest.raises(RuntimeError, match="no CPU path"):
        image_proc.nms(torch.zeros(1, 10, 85))
    with pytest.raises(RuntimeError, match="no CPU path"):
        bboxes.cxcywh_to_xyxy(torch.zeros(4, 4))
    with pytest.raises(AssertionError):
        image_proc.nms(torch.zeros(1, 10, 85), iou_thres=2.0)


def test_norm255_split_reciprocal_is_correctly_rounded():
    # letterbox.cu norm255(): fma(v, hi, fl(v * lo)) with hi = RN(1/255), lo = RN(1/255 - hi).
    # Exact rational arithmetic, each step rounded to float32 -> equals RN(v/255) for all 256 inputs.
    def rn_exact(x: Fraction) -> Fraction:   # round-to-nearest-even of an exact rational to float32
        if x == 0:
            return Fraction(0)
        f = np.float32(float(x))             # float(x) is the correctly rounded double; doubles of these
        lo, hi = np.nextafter(f, np.float32(-np.inf)), np.nextafter(f, np.float32(np.inf))
        best = min((abs(Fraction(float(c)) - x), i, c) for i, c in enumerate((f, lo, hi)))
        return Fraction(float(best[2]))

    hi = Fraction(float(np.float32(1.0) / np.float32(255.0)))
    assert float(hi) == float(np.float32(0.003921568859368563))
    lo = rn_exact(Fraction(1, 255) - hi)
    assert float(lo) == float(np.float32(-2.319175823606301e-10))
    for v in range(256):
        q = rn_exact(Fraction(v) * hi + rn_exact(Fraction(v) * lo))
        assert float(q) == float(np.float32(v) / np.float32(255.0)), v


def test_iou_threshold_rounding_rule():
    # vk_nms_batched compares against the largest float32 <= the double threshold
    for thr in (0.6, 0.45, 0.5, 0.0, 1.0, float(np.float32(0.6))):
        f = np.float32(thr)
        if float(f) > thr:
            f = np.nextafter(f, np.float32(-np.inf))
        for x in (np.nextafter(f, np.float32(-np.inf)), f, np.nextafter(f, np.float32(np.inf))):
            assert (float(x) > thr) == (x > f)


# ------------------------------------------------------------------ world_size-2 over gloo
def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, n_images, ret):
    import torch.distributed as dist
    from vision_kit_b200 import dist as vkd
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    lo, hi = vkd.shard_range(n_images, rank, world)
    g = torch.Generator().manual_seed(0)
    all_d = torch.rand((n_images, 7, 6), generator=g)
    all_c = torch.randint(0, 8, (n_images,), generator=g, dtype=torch.int32)
    d, c = vkd.allgather_detections(all_d[lo:hi].clone(), all_c[lo:hi].clone(), n_images)
    ok = torch.equal(d, all_d) and torch.equal(c, all_c)
    ret[rank] = bool(ok)
    dist.destroy_process_group()


@pytest.mark.parametrize("n_images", [8, 7])
def test_shard_and_allgather_world2(n_images):
    from vision_kit_b200 import dist as vkd
    assert vkd.shard_sizes(7, 2) == [4, 3] and vkd.shard_range(7, 1, 2) == (4, 7)
    assert sum(vkd.shard_sizes(512, 8)) == 512 and vkd.shard_range(256, 3, 8) == (96, 128)
    ctx = mp.get_context("spawn")
    ret = ctx.Manager().dict()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, n_images, ret)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    assert ret[0] and ret[1]


def test_dataset_geometry_equals_oracle(vk_lib):
    """vk_dataset_geometry (C, host) against oracle/restate.py::dataset_geometry on random sizes,
    including r == 1 and the int() truncation cases."""
    from vision_kit_b200 import ops
    from oracle import restate
    rng = np.random.Generator(np.random.PCG64(4))
    cases = [(640, 640), (480, 640), (1280, 720), (333, 500), (1279, 853), (2000, 3000), (640, 1)]
    cases += [(int(rng.integers(1, 3000)), int(rng.integers(1, 3000))) for _ in range(300)]
    for h, w in cases:
        r, (nh, nw), (top, bottom, left, right) = restate.dataset_geometry(h, w, (640, 640))
        if nh <= 0 or nw <= 0:
            continue
        g = ops.dataset_geometry(h, w, (640, 640))
        assert (g.new_h, g.new_w, g.top, g.bottom, g.left, g.right) == (nh, nw, top, bottom, left, right), (h, w)
        assert g.ratio == r and (g.out_h, g.out_w) == (640, 640)
    from vision_kit_b200 import _lib
    g = _lib.VkLbGeom()
    assert vk_lib.vk_dataset_geometry(1000, 500, 512, 640, C.byref(g)) == -1      # 640 tall on a 512 canvas


def test_filter_and_nms_argument_validation(vk_lib):
    """Entry points reject bad arguments before touching the device: unknown kernel selector / dtype,
    thresholds outside [0, 1], inconsistent candidate buffers."""
    from vision_kit_b200 import _lib, ops
    cfg = ops.head_cfg("v5", 80, [[10, 13, 16, 30, 33, 23]] * 3, (8, 16, 32), [(80, 80), (40, 40), (20, 20)])
    lv = (C.c_void_p * 4)(256, 256, 256, 0)
    cb = _lib.VkCandBuf()
    cb.cand = cb.boxes = cb.ctrl = cb.seg_count = 256
    cb.cap, cb.rows, cb.segs, cb.nc = 396 * 64, 25200, 396, 80
    args = lambda dtype, conf, kernel: (C.byref(cfg), C.cast(lv, C.c_void_p), dtype, 2, C.c_float(conf), 0, None, kernel,
                                        C.byref(cb), None)
    assert vk_lib.vk_decode_filter(*args(0, 0.25, 7)) == -1 and b"kernel 7" in vk_lib.vk_last_error()
    assert vk_lib.vk_decode_filter(*args(5, 0.25, 0)) == -1 and b"dtype 5" in vk_lib.vk_last_error()
    assert vk_lib.vk_decode_filter(*args(0, 1.5, 0)) == -1 and b"conf_thres" in vk_lib.vk_last_error()
    cb.list_cap = 100                                                  # list_cap without a list
    assert vk_lib.vk_decode_filter(*args(0, 0.25, 0)) == -1 and b"list" in vk_lib.vk_last_error()
    cb.list_cap = 0
    cb.cap = 100                                                       # too few slots: every tile owns 64
    assert vk_lib.vk_decode_filter(*args(0, 0.25, 0)) == -1 and b"cap 100" in vk_lib.vk_last_error()
    assert vk_lib.vk_nms_batched(C.byref(cb), 2, C.c_double(1.5), 0, 30000, 300, C.c_float(7680.0), 256, 256, None, None,
                                 None) == -1 and b"iou_thres" in vk_lib.vk_last_error()
    assert vk_lib.vk_nms_batched(C.byref(cb), 2, C.c_double(0.5), 0, 30000, 5000, C.c_float(7680.0), 256, 256, None, None,
                                 None) == -3                            # VK_E_LIMIT: max_det
    assert vk_lib.vk_eval_match_smem_bytes(10, 100) == 100 * 6 * 4 + 10 * 100 * 4
    assert vk_lib.vk_eval_match_smem_bytes(0, 100) == 0
    assert ops.expects_dense("auto", 0.001) and not ops.expects_dense("auto", 0.25) and ops.expects_dense("dense", 0.25)
    assert ops.expects_dense("dense_onepass", 0.25) and ops._KERNEL["dense_onepass"] == _lib.VK_FILTER_DENSE_ONEPASS == 3
    assert vk_lib.vk_decode_filter(*args(0, 0.25, 4)) == -1 and b"kernel 4" in vk_lib.vk_last_error()
    assert _lib.C.sizeof(_lib.VkCandBuf) == 5 * 8 + 6 * 4


def test_bench_reference_arm_prints_the_contract_line():
    """`bench.py --impl reference` (the driver's CPU arm) runs without a GPU and prints ONE JSON line with the keys of
    the bench contract: same metric / unit / config as our arm, `impl`, a `cpu_baseline` describing the run and an
    `e2e` object that repeats the value with zero copies."""
    import json, os, subprocess, sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    env = dict(os.environ, CUDA_VISIBLE_DEVICES="")
    r = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "1"],
                       capture_output=True, text=True, timeout=600, env=env, cwd=root)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    baseline = json.load(open(os.path.join(root, "BASELINE.json")))
    assert d["impl"] == "reference" and d["metric"] == baseline["metric"] and d["unit"] == "images/s"
    assert d["higher_is_better"] is True and d["n_gpus"] == 1 and d["steps"] == 1
    assert d["config"]["workload"].startswith("configs[1]") and d["config"]["batch_per_gpu"] == 64
    cb = d["cpu_baseline"]
    assert cb["kind"] in ("reference", "port") and cb["cores"] >= 1 and cb["value"] == d["value"] > 0
    assert d["e2e"] == {"value": d["value"], "unit": "images/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}


def test_pairs_kernel_pretest_error_budget():
    """The dense multi-label filter drops a logit x without evaluating its sigmoid when x <= logit_floor(conf, obj)
    (csrc/decode.cu).  That is only allowed if such a logit can never pass `fl(sigmoid_approx(x) * obj) > conf`,
    where the approximate sigmoid is within 4e-7 (relative) of the true one and the product rounds once.  The floor is
    restated here in float32 arithmetic step by step (lg2.approx modelled with a 2^-22 absolute error in either
    direction) and the claim is checked in float64 for logits AT the floor -- the worst case, sigma is increasing --
    with every error at its unfavourable end, over thresholds and objectness values from barely above conf to 1."""
    f32 = np.float32
    rng = np.random.default_rng(7)
    worst = -np.inf
    for conf in (1e-6, 1e-3, 0.01, 0.05, 0.25, 0.5, 0.9, 0.999):
        obj = np.concatenate([conf * (1 + 10.0 ** rng.uniform(-7, 0, 4000)), rng.uniform(conf, 1, 4000), [1.0]])
        obj = np.unique(np.minimum(obj, 1.0).astype(f32))
        obj = obj[obj > f32(conf)]
        s = np.minimum((f32(conf) / obj).astype(f32) * f32(0.999996), f32(0.999)).astype(f32)
        u = ((f32(1.0) / s).astype(f32) - f32(1.0)).astype(f32)
        for lg_err in (-2.0 ** -22, 2.0 ** -22):
            l = (np.log2(u.astype(np.float64)) + lg_err).astype(f32)
            floor = ((l * f32(-0.6931471805599453)).astype(f32) - f32(1e-3)).astype(f32)
            x = floor.astype(np.float64)
            sig = 1.0 / (1.0 + np.exp(-x))
            p_max = sig * (1 + 4e-7) * obj.astype(np.float64) * (1 + 2.0 ** -24)     # the largest the kernel could compute
            worst = max(worst, float(np.max(p_max / np.float64(f32(conf)))))
            assert np.all(p_max <= np.float64(f32(conf))), (conf, float(np.max(p_max / conf)))
    assert worst < 1.0 and worst > 0.99           # a margin, but not a loose one: the floor sits right under the cut
