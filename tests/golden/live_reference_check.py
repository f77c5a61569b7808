"""Randomised LIVE comparison of the CPU oracle with the reference's own code -- many seeds, not the few committed fixtures.

Runs only where /root/reference exists (the build container); driven as a subprocess by tests/test_oracle_live_reference.py
because it injects a stub `mindspore` into sys.modules (see make_golden.py).  Usage:
    python tests/golden/live_reference_check.py [n_seeds]
Exits non-zero on the first disagreement.  What is compared, per seed:
  * nms_jit (pointpillars/src/core/nms.py:85-112)         vs oracle NMS, mode offset 0 / >= / no union guard   keep lists equal
  * apply_nms (nms.py:7-41)                               vs oracle NMS, mode offset 1 / >                     keep lists equal
  * iou_normal, host-run (iou3d_nms_kernel.cu:347-358)    vs oracle NMS, DEFAULT mode, lattice boxes            keep lists equal
  * create_target_np (target_assigner.py:84-134)          vs oracle assign mode 1 (labels, matched gt ids)     equal;
    and mode 0 differs from it only on anchors tying one gt's best IoU while their argmax is another gt
  * iou_jit eps 0 / 1 (box_np_ops.py:639-679)             vs oracle IoU matrix                                 <= 2e-6
  * topk_ (nms.py:66-83)                                  vs oracle top-k on unique scores                      equal
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, HERE)
sys.path.insert(0, REPO)

import make_golden as mg  # noqa: E402


def main(n_seeds):
    mg._stub_mindspore()
    sys.path.insert(0, os.path.join(mg.REF, "pointpillars"))
    from src.core import box_np_ops, nms as ref_nms, target_assigner
    import oracle as O
    from oracle import bev

    def keep_of(mask):
        return np.nonzero(mask)[0]

    checks = 0
    for seed in range(n_seeds):
        rng = np.random.default_rng(977 + seed)
        n = int(rng.integers(40, 700))
        cl = None if seed % 3 == 0 else int(rng.integers(2, 30))
        thr = float(np.float32(rng.choice([0.1, 0.3, 0.45, 0.5, 0.7, 0.9])))
        boxes = mg.rand_boxes(rng, n, cluster=cl)
        if seed % 4 == 1:
            boxes = np.round(boxes)                       # integer corners: exact ties at the threshold become possible
        boxes[n // 2] = boxes[0]
        scores = rng.permutation(n).astype(np.float32) / n
        order = np.argsort(-scores, kind="stable")
        dets = np.concatenate([boxes, scores[:, None]], 1).astype(np.float32)

        ref = np.asarray(ref_nms.nms_jit(dets, thr, eps=0.0))
        got = order[keep_of(O.nms(dets[order], np.float32(thr), off=0.0, inclusive=True, union_eps=0.0))]
        assert np.array_equal(got, ref), ("nms_jit", seed)

        ref = np.asarray(ref_nms.apply_nms(mg._AsNumpy(boxes), mg._AsNumpy(scores), np.float32(thr), 10 ** 9), np.int64)
        got = order[keep_of(O.nms(boxes[order], np.float32(thr), off=1.0, inclusive=False, union_eps=0.0))]
        assert np.array_equal(got, ref), ("apply_nms", seed)

        lat = (np.round(boxes * 4) / 4).astype(np.float32)
        lat[:, 2] = np.maximum(lat[:, 2], lat[:, 0] + 0.25)
        lat[:, 3] = np.maximum(lat[:, 3], lat[:, 1] + 0.25)
        b7 = np.zeros((n, 7), np.float32)
        b7[:, 0], b7[:, 1] = (lat[:, 0] + lat[:, 2]) / 2, (lat[:, 1] + lat[:, 3]) / 2
        b7[:, 3], b7[:, 4] = lat[:, 2] - lat[:, 0], lat[:, 3] - lat[:, 1]
        ref = bev.ref_cu_nms(b7, thr, rotated=False)
        got = keep_of(O.nms(lat, np.float32(thr), off=0.0, inclusive=False, union_eps=1e-8))
        assert np.array_equal(got, ref), ("iou_normal nms", seed)
        assert np.array_equal(bev.iou_normal(b7[:32], b7), bev.ref_cu_pairs("iou_normal", b7[:32], b7)), ("iou_normal", seed)

        na, ng = int(rng.integers(300, 2500)), int(rng.integers(1, 20))
        anchors = np.round(mg.rand_boxes(rng, na, smin=16, smax=300))
        gts = np.round(mg.rand_boxes(rng, ng, smin=24, smax=400))
        if seed % 2:
            anchors[10:20] = anchors[40:50]               # duplicated anchors: ties for a gt's best anchor
            gts[0] = anchors[45]
        pos = float(rng.choice([0.5, 0.6, 0.7]))
        neg = float(rng.choice([0.3, 0.4, pos]))
        r = target_assigner.create_target_np(
            anchors, gts, lambda a, g: box_np_ops.iou_jit(a, g, eps=1.0).astype(np.float32),
            lambda g, a: np.zeros((a.shape[0], 4), np.float32), matched_threshold=pos, unmatched_threshold=neg,
            positive_fraction=None, box_code_size=4)
        labels = r["labels"].astype(np.int32)
        gt_ids = np.full(na, -1, np.int32)
        gt_ids[r["assigned_anchors_inds"]] = r["positive_gt_id"]
        a1, _, am = O.assign(anchors, gts, pos, neg, 0.0, off=1.0, mode=1)
        assert np.array_equal(a1 > 0, labels > 0) and np.array_equal(a1 == 0, labels == 0), ("assign labels", seed)
        fg = labels > 0
        assert np.array_equal(a1[fg] - 1, gt_ids[fg]), ("assign gt ids", seed)
        a0, _, _ = O.assign(anchors, gts, pos, neg, 1e-30, off=1.0, mode=0)
        iou = O.iou_matrix(anchors, gts, off=1.0)
        gmax = iou.max(0)
        ties = (iou == gmax[None, :]) & (gmax[None, :] > 0)
        forced = ties.any(1)
        last = np.where(forced, ties.shape[1] - 1 - np.argmax(ties[:, ::-1], 1), -1)
        differ = forced & (last != am)
        assert np.array_equal(a0[~differ], a1[~differ]), ("mode 0 vs 1", seed)
        assert np.array_equal(a0[differ], last[differ] + 1), ("mode 0 forced", seed)

        for eps in (0.0, 1.0):
            refm = box_np_ops.iou_jit(anchors[:400], gts, eps=eps)
            gotm = O.iou_matrix(anchors[:400], gts, off=eps)
            assert np.array_equal(gotm == 0, refm == 0) and np.allclose(gotm, refm, rtol=2e-6, atol=1e-7), ("iou_jit", eps, seed)

        m = int(rng.integers(2000, 30000))
        sc = (rng.permutation(m).astype(np.float32) / np.float32(m) - np.float32(0.5)) * np.float32(9)
        k = int(rng.integers(2, min(m, 2049)))
        vals, idx = ref_nms.topk_(sc[:, None], k + 1, axis=0)
        v, i = O.topk(sc, k)
        assert np.array_equal(i, idx[:, 0]) and np.array_equal(v, vals[:, 0]), ("topk_", seed)
        checks += 8
    print("live reference check OK:", n_seeds, "seeds,", checks, "comparisons")


if __name__ == "__main__":
    main(int(sys.argv[1]) if len(sys.argv) > 1 else 12)
