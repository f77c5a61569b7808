"""Generate the golden fixtures in tests/golden/ by RUNNING THE REFERENCE's own code.

Run once in the build container (needs /root/reference; the GPU box does not have it):

    python tests/golden/make_golden.py

What is executed (all paths relative to /root/reference/minddet/models):
  * pointpillars/src/core/nms.py:85-112      nms_jit(dets, thresh, eps=0.0)
  * pointpillars/src/core/nms.py:7-41        apply_nms(boxes, scores, thres, max_boxes)  (+1 areas)
  * pointpillars/src/core/box_np_ops.py:639-679   iou_jit(boxes, query, eps=1.0)
  * pointpillars/src/core/target_assigner.py:29-166  create_target_np(..., positive_fraction=None)
  * pointpillars/src/core/box_np_ops.py:453-523   create_anchors_3d_stride (grid order)
  * pointpillars/src/core/box_np_ops.py:639-679   iou_jit(boxes, query, eps=0.0)  (offset-0 IoU)
  * pointpillars/src/core/eval_utils.py:118-165   image_box_overlap(boxes, query, criterion=-1)
  * pointpillars/src/data/kitti_common.py:10-73    iou(boxes1, boxes2, add1)
  * pointpillars/src/core/nms.py:66-83            topk_(matrix, K, axis=0)
  * pointpillars/src/predict.py:98-112            softmax(x, axis=1), sigmoid(x)
  * centerpoint/det3d_ms/core/utils/center_utils.py:97-131  bilinear_interpolate_torch (4-tap weights)
  * centerpoint/det3d_ms/ops/test_custom_pytorch/iou3d_nms_kernel.cu:42,347-358  EPS and iou_normal, cut out of the
    file where it lies by oracle/Makefile and compiled for the host (oracle/ref_cu_device_harness.cpp ->
    oracle/_ref/iou3d_device_ref.so); the pair loop / greedy sweep around it restate nms_normal_kernel :361-405 and
    the host reduce :526-536
  * centerpoint/det3d_ms/ops/iou-bev-nms-org.cpp:237-283  boxes_iou_nms_cpu, compiled by
    oracle/Makefile into oracle/_ref/nms_fast_ref.so and called through the 7-argument aot ABI.
`mindspore` is not installed, so a stub module is injected; none of the functions above touch it
except apply_nms's `.asnumpy()` calls, which a tiny wrapper provides.
"""
import ctypes
import os
import sys
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(os.path.dirname(HERE))
REF = "/root/reference/minddet/models"


def _stub_mindspore():
    ms = types.ModuleType("mindspore")
    ms.ops = types.ModuleType("mindspore.ops")
    ms.nn = types.ModuleType("mindspore.nn")

    class _T:  # the two MindSpore calls create_anchors_3d_stride makes, with numpy semantics
        @staticmethod
        def from_numpy(a):
            return _AsNumpy(a)

    ms.Tensor = _T
    ms.ops.Tensor = _T
    ms.ops.meshgrid = lambda *ts, indexing="xy": tuple(
        _AsNumpy(m) for m in np.meshgrid(*[t.asnumpy() for t in ts], indexing=indexing))
    sys.modules["mindspore"] = ms
    sys.modules["mindspore.ops"] = ms.ops
    sys.modules["mindspore.nn"] = ms.nn
    ms.dtype = types.ModuleType("mindspore.dtype")               # imported (never touched) by pointpillars/src/predict.py
    ms.numpy = types.ModuleType("mindspore.numpy")
    sys.modules["mindspore.dtype"] = ms.dtype
    sys.modules["mindspore.numpy"] = ms.numpy
    ms.common = types.ModuleType("mindspore.common")           # imported (never called) by circle_nms_jit.py
    ms.common.dtype = types.ModuleType("mindspore.common.dtype")
    sys.modules["mindspore.common"] = ms.common
    sys.modules["mindspore.common.dtype"] = ms.common.dtype


class _AsNumpy:
    def __init__(self, a):
        self.a = a

    def asnumpy(self):
        return self.a


def rand_boxes(rng, n, img_w=1344.0, img_h=800.0, smin=8.0, smax=400.0, cluster=None):
    if cluster is None:
        cx = rng.uniform(0, img_w, n)
        cy = rng.uniform(0, img_h, n)
    else:
        ctr = rng.uniform([0, 0], [img_w, img_h], (cluster, 2))
        pick = rng.integers(0, cluster, n)
        cx = ctr[pick, 0] + rng.normal(0, 12, n)
        cy = ctr[pick, 1] + rng.normal(0, 12, n)
    w = np.exp(rng.uniform(np.log(smin), np.log(smax), n))
    h = np.exp(rng.uniform(np.log(smin), np.log(smax), n))
    b = np.stack([cx - w / 2, cy - h / 2, cx + w / 2, cy + h / 2], 1)
    b[:, 0::2] = np.clip(b[:, 0::2], 0, img_w - 1)
    b[:, 1::2] = np.clip(b[:, 1::2], 0, img_h - 1)
    return b.astype(np.float32)


def main():
    _stub_mindspore()
    sys.path.insert(0, os.path.join(REF, "pointpillars"))
    from src.core import box_np_ops, nms as ref_nms, target_assigner

    rng = np.random.default_rng(20261018)
    out = {}

    # ---- nms_jit (offset 0, inclusive >=, no union guard) ------------------------------------
    for tag, n, cl in (("a", 300, None), ("b", 500, 12), ("c", 64, 3)):
        boxes = rand_boxes(rng, n, cluster=cl)
        if tag == "c":  # exact duplicates and degenerate boxes
            boxes[10] = boxes[3]
            boxes[20, 2] = boxes[20, 0]
        scores = rng.permutation(n).astype(np.float32) / n
        dets = np.concatenate([boxes, scores[:, None]], 1).astype(np.float32)
        for thr in (0.3, 0.7):
            keep = np.array(ref_nms.nms_jit(dets, np.float32(thr), eps=0.0), dtype=np.int64)
            out[f"nmsjit_{tag}_{thr}_dets"] = dets
            out[f"nmsjit_{tag}_{thr}_keep"] = keep

    # ---- apply_nms (+1 areas, keeps ovr <= thr i.e. suppresses strict >) ---------------------
    for tag, n, cl in (("a", 256, None), ("b", 400, 10)):
        boxes = np.round(rand_boxes(rng, n, cluster=cl))
        scores = rng.permutation(n).astype(np.float32) / n
        for thr in (0.5, 0.7):
            keep = ref_nms.apply_nms(_AsNumpy(boxes), _AsNumpy(scores), np.float32(thr), 10 ** 9)
            out[f"applynms_{tag}_{thr}_boxes"] = boxes
            out[f"applynms_{tag}_{thr}_scores"] = scores
            out[f"applynms_{tag}_{thr}_keep"] = np.asarray(keep, np.int64)

    # ---- iou_jit(eps=1) ------------------------------------------------------------------------
    a = rand_boxes(rng, 700)
    g = rand_boxes(rng, 23, smin=16, smax=512)
    out["iou_boxes"] = a
    out["iou_gts"] = g
    out["iou_mat"] = box_np_ops.iou_jit(a, g, eps=1.0)

    # ---- create_target_np (positive_fraction=None: deterministic) -------------------------------
    def sim(anchors, gts):
        return box_np_ops.iou_jit(anchors, gts, eps=1.0).astype(np.float32)

    def enc(gts, anchors):
        return np.zeros((anchors.shape[0], 4), np.float32)

    for tag, n, ng, pos, neg in (("a", 3000, 9, 0.7, 0.3), ("b", 2000, 17, 0.5, 0.5), ("c", 1500, 5, 0.6, 0.45)):
        anchors = np.round(rand_boxes(rng, n, smin=16, smax=300))
        gts = np.round(rand_boxes(rng, ng, smin=24, smax=400))
        if tag == "b":  # duplicated anchors -> ties for a gt's best anchor; one far-away gt
            anchors[100:110] = anchors[50:60]
            gts[3] = anchors[55]
            gts[4] = np.array([5000, 5000, 5100, 5100], np.float32)
        r = target_assigner.create_target_np(
            anchors, gts, sim, enc, matched_threshold=float(pos), unmatched_threshold=float(neg),
            positive_fraction=None, box_code_size=4)
        gt_ids = np.full(n, -1, np.int32)
        gt_ids[r["assigned_anchors_inds"]] = r["positive_gt_id"]
        out[f"assign_{tag}_anchors"] = anchors
        out[f"assign_{tag}_gts"] = gts
        out[f"assign_{tag}_thr"] = np.array([pos, neg], np.float32)
        out[f"assign_{tag}_labels"] = r["labels"].astype(np.int32)
        out[f"assign_{tag}_gtids"] = gt_ids

    # ---- grid order of the reference anchor generator -------------------------------------------
    # (create_anchors_3d_range :526-568 is broken under numpy 2 -- tuple assignment -- so the
    #  strided generator :453-523 is used; ms.ops.meshgrid is stubbed with np.meshgrid.)
    anc = box_np_ops.create_anchors_3d_stride([1, 5, 7], sizes=(1.0, 2.0, 3.0), rotations=(0.0, 1.0),
                                              anchor_range=[0.0, 0.0, 0.0, 6 * 8.0, 4 * 8.0, 0.0])
    out["grid_ref"] = np.asarray(anc, np.float32)  # [1,5,7,1,2,7]: x fastest, per-cell variants innermost

    # ---- compiled reference aot CPU op ----------------------------------------------------------
    so = os.path.join(REPO, "oracle", "_ref", "nms_fast_ref.so")
    lib = ctypes.CDLL(so)
    n = 1000  # the reference hard-codes N=1000 (iou-bev-nms-org.cpp:244)
    ctr = rng.uniform([0, -40], [70, 40], (40, 2))
    pick = rng.integers(0, 40, n)
    boxes = np.zeros((n, 7), np.float32)
    boxes[:, 0] = ctr[pick, 0] + rng.normal(0, 1.0, n)
    boxes[:, 1] = ctr[pick, 1] + rng.normal(0, 1.0, n)
    boxes[:, 3] = rng.uniform(1.5, 4.5, n)
    boxes[:, 4] = rng.uniform(1.2, 2.2, n)
    boxes[:, 5] = 1.5
    boxes[:, 6] = rng.uniform(-np.pi, np.pi, n)
    boxes[-20:, 3] = 0.0  # zero-area padding rows are pre-removed (:253-255)
    thr = np.array([0.2], np.float32)
    keep = np.zeros(n, np.int32)
    cnt = np.zeros(1, np.int32)
    params = (ctypes.c_void_p * 4)(boxes.ctypes.data, thr.ctypes.data, keep.ctypes.data, cnt.ctypes.data)
    rc = lib.boxes_iou_nms_cpu(4, params, None, None, None, None, None)
    assert rc == 0
    assert lib.boxes_iou_nms_cpu(3, params, None, None, None, None, None) == 1
    out["rotnms_boxes"] = boxes
    out["rotnms_thr"] = thr
    out["rotnms_keep"] = keep
    out["rotnms_count"] = cnt

    # ---- added in round 2: independent generator, so every array above keeps its bytes -------------
    rng2 = np.random.default_rng(20261019)

    # iou_jit(eps=0) and eval_utils.image_box_overlap(criterion=-1): the offset-0 IoU (no +1)
    from src.core import eval_utils
    a0 = rand_boxes(rng2, 400)
    g0 = rand_boxes(rng2, 31, smin=16, smax=512)
    a0[7] = g0[3]                                  # identical pair (IoU 1)
    a0[8] = g0[4] + np.float32(2000.0)             # disjoint
    a0[9, 0] = g0[5, 2]; a0[9, 2] = a0[9, 0] + 40  # touching edge: iw == 0 -> 0
    out["iou0_boxes"] = a0
    out["iou0_gts"] = g0
    out["iou0_mat_iou_jit"] = box_np_ops.iou_jit(a0, g0, eps=0.0)
    out["iou0_mat_image_box_overlap"] = eval_utils.image_box_overlap(a0, g0, criterion=-1)

    # bilinear_interpolate_torch (centerpoint/det3d_ms/core/utils/center_utils.py:97-131): the 4-tap
    # weights on interior points.  The file is executed from where it lies through a synthetic package
    # (its own package __init__ imports MindSpore), so its relative import of circle_nms_jit resolves.
    import importlib
    import torch
    pkg = types.ModuleType("_ref_center_utils_pkg")
    pkg.__path__ = [os.path.join(REF, "centerpoint", "det3d_ms", "core", "utils")]
    sys.modules["_ref_center_utils_pkg"] = pkg
    cu = importlib.import_module("_ref_center_utils_pkg.center_utils")
    Hh, Ww, Cc = 19, 27, 5
    im = rng2.uniform(-1, 1, (Hh, Ww, Cc)).astype(np.float32)
    n = 600
    x = rng2.uniform(0.0, Ww - 1.0, n).astype(np.float32)
    y = rng2.uniform(0.0, Hh - 1.0, n).astype(np.float32)
    x[:8] = np.arange(8, dtype=np.float32)          # exactly on lattice columns
    y[4:12] = np.arange(8, dtype=np.float32) + 2     # exactly on lattice rows
    x = np.minimum(x, np.float32(Ww - 1.001)); y = np.minimum(y, np.float32(Hh - 1.001))
    val = cu.bilinear_interpolate_torch(torch.from_numpy(im), torch.from_numpy(x), torch.from_numpy(y)).numpy()
    out["bilinear_im"] = im
    out["bilinear_x"] = x
    out["bilinear_y"] = y
    out["bilinear_val"] = val.astype(np.float32)

    # ---- iou_normal (iou3d_nms_kernel.cu:347-358), cut out of the reference's CUDA file by oracle/Makefile and run on
    # the host (oracle/ref_cu_device_harness.cpp): the arithmetic of the DEFAULT NMS mode (offset 0, strict >, union
    # guard EPS).  Boxes on a 0.25-px lattice: xyxy <-> (centre, size) converts without rounding, so the fixture pins
    # bits, not tolerances; one off-lattice case with a threshold margin checks the general behaviour.
    nl = ctypes.CDLL(os.path.join(REPO, "oracle", "_ref", "iou3d_device_ref.so"))
    nl.ref_cu_eps.restype = ctypes.c_float
    vp = ctypes.c_void_p
    nl.ref_cu_pair_matrix.argtypes = [ctypes.c_int, vp, ctypes.c_int, vp, ctypes.c_int, vp]
    nl.ref_cu_nms.argtypes = [ctypes.c_int, vp, ctypes.c_int, ctypes.c_float, vp, ctypes.POINTER(ctypes.c_int)]
    out["ioun_eps"] = np.array([nl.ref_cu_eps()], np.float32)

    def to7(xyxy):
        b7 = np.zeros((len(xyxy), 7), np.float32)
        b7[:, 0] = (xyxy[:, 0] + xyxy[:, 2]) / np.float32(2)
        b7[:, 1] = (xyxy[:, 1] + xyxy[:, 3]) / np.float32(2)
        b7[:, 3] = xyxy[:, 2] - xyxy[:, 0]
        b7[:, 4] = xyxy[:, 3] - xyxy[:, 1]
        b7[:, 5] = 1.0
        return b7

    def ref_matrix(a7, b7):
        m = np.zeros((len(a7), len(b7)), np.float32)
        nl.ref_cu_pair_matrix(0, a7.ctypes.data, len(a7), b7.ctypes.data, len(b7), m.ctypes.data)
        return m

    def ref_keep(b7, thr):
        keep = np.zeros(len(b7), np.int64)
        cnt = ctypes.c_int(0)
        nl.ref_cu_nms(0, b7.ctypes.data, len(b7), thr, keep.ctypes.data, ctypes.byref(cnt))
        return keep[:cnt.value].copy()

    for tag, n, cl, lattice in (("a", 600, None, True), ("b", 900, 25, True), ("c", 64, 4, True), ("d", 700, 20, False)):
        xy = rand_boxes(rng2, n, cluster=cl)
        if lattice:
            xy = (np.round(xy * 4) / 4).astype(np.float32)
            xy[:, 2] = np.maximum(xy[:, 2], xy[:, 0] + 0.25)
            xy[:, 3] = np.maximum(xy[:, 3], xy[:, 1] + 0.25)
            if tag == "c":
                xy[10] = xy[3]                       # exact duplicate
                xy[20, 2] = xy[20, 0]                # zero-width box: area 0, union guard
                xy[21] = xy[20]
            b7 = to7(xy)
            assert np.array_equal(b7[:, 0] - b7[:, 3] / 2, xy[:, 0]) and np.array_equal(b7[:, 0] + b7[:, 3] / 2, xy[:, 2])
            assert np.array_equal(b7[:, 1] - b7[:, 4] / 2, xy[:, 1]) and np.array_equal(b7[:, 1] + b7[:, 4] / 2, xy[:, 3])
        else:
            b7 = to7(xy)
        out[f"ioun_{tag}_xyxy"] = xy                 # rows are already in score order (row 0 = best)
        out[f"ioun_{tag}_box7"] = b7
        m = ref_matrix(b7, b7)
        out[f"ioun_{tag}_iou"] = m[:96].copy()       # first 96 rows against all boxes (keeps the fixture small)
        for thr in (0.3, 0.7):
            t = float(np.float32(thr))
            if not lattice:
                assert np.abs(m[np.triu_indices(n, 1)] - t).min() > 1e-5   # no pair on the threshold
            out[f"ioun_{tag}_{thr}_keep"] = ref_keep(b7, t)

    # ---- topk_ (pointpillars/src/core/nms.py:66-83), the reference's numpy top-k (argpartition + argsort; it returns the
    # K-1 best: `K = K - 1` on its first line).  Unique scores, so tie order cannot matter.
    sc = (rng2.permutation(20000).astype(np.float32) / np.float32(20000) - np.float32(0.5)) * np.float32(12)
    assert len(np.unique(sc)) == len(sc)
    vals, idx = ref_nms.topk_(sc[:, None], 1001, axis=0)
    out["topk_scores"] = sc
    out["topk_ref_vals"] = vals[:, 0].astype(np.float32)
    out["topk_ref_idx"] = idx[:, 0].astype(np.int64)

    # ---- kitti_common.iou (pointpillars/src/data/kitti_common.py:10-73): vectorised numpy IoU, with and without the +1
    for name in ("skimage", "skimage.io"):          # imported at the top of kitti_common.py (image loading), never called here
        sys.modules.setdefault(name, types.ModuleType(name))
    from src.data import kitti_common
    out["iou0_mat_kitti"] = kitti_common.iou(a0, g0, add1=False).astype(np.float32)
    out["iou1_mat_kitti"] = kitti_common.iou(a0, g0, add1=True).astype(np.float32)

    # ---- sigmoid / softmax (pointpillars/src/predict.py:98-112): the score activations of the post-process rows
    from src import predict
    ax = rng2.normal(0, 3, (300, 16)).astype(np.float32)
    ax[0, :8] = [-30, -20, -10, -1e-3, 0, 1e-3, 10, 20]
    ax[1] = 0.0
    ax[2] = np.linspace(-12, 12, 16, dtype=np.float32)
    out["act_x"] = ax
    out["act_sigmoid_ref"] = predict.sigmoid(ax).astype(np.float32)
    out["act_softmax_ref"] = predict.softmax(ax, axis=1).astype(np.float32)

    path = sys.argv[1] if len(sys.argv) > 1 else os.path.join(HERE, "reference_golden.npz")   # argv[1]: tests/test_golden_regenerates.py
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes;", len(out), "arrays")


if __name__ == "__main__":
    main()
