"""Pin the CPU oracle against fixtures produced by RUNNING the reference's own code
(tests/golden/make_golden.py).  CPU only."""
import numpy as np

import oracle as O


def _keep_from_mask(m):
    return np.nonzero(m)[0]


def test_nms_mode_nms_jit_bit_exact(golden):
    # pointpillars/src/core/nms.py:85-112: offset 0, suppress iou >= thr, no union guard.
    for tag in "abc":
        for thr in (0.3, 0.7):
            dets = golden[f"nmsjit_{tag}_{thr}_dets"]
            ref_keep = golden[f"nmsjit_{tag}_{thr}_keep"]
            order = np.argsort(-dets[:, 4], kind="stable")  # scores are a permutation: no ties
            mask = O.nms(dets[order], np.float32(thr), off=0.0, inclusive=True, union_eps=0.0)
            assert np.array_equal(order[_keep_from_mask(mask)], ref_keep), (tag, thr)


def test_nms_mode_apply_nms_bit_exact(golden):
    # pointpillars/src/core/nms.py:7-41: +1 areas, keeps ovr <= thr (suppress strict >).
    for tag in "ab":
        for thr in (0.5, 0.7):
            boxes = golden[f"applynms_{tag}_{thr}_boxes"]
            scores = golden[f"applynms_{tag}_{thr}_scores"]
            ref_keep = golden[f"applynms_{tag}_{thr}_keep"]
            order = np.argsort(-scores, kind="stable")
            mask = O.nms(boxes[order], np.float32(thr), off=1.0, inclusive=False, union_eps=0.0)
            assert np.array_equal(order[_keep_from_mask(mask)], ref_keep), (tag, thr)


def test_iou_plus_one_matches_iou_jit(golden):
    # pointpillars/src/core/box_np_ops.py:639-679 with eps=1.0 (numba evaluates in float64).
    got = O.iou_matrix(golden["iou_boxes"], golden["iou_gts"], off=1.0)
    ref = golden["iou_mat"]
    assert np.array_equal(got == 0, ref == 0)
    np.testing.assert_allclose(got, ref, rtol=1e-6, atol=1e-7)


def test_assign_mode1_matches_create_target_np(golden):
    # pointpillars/src/core/target_assigner.py:84-134 (positive_fraction=None).
    for tag in "abc":
        anchors, gts = golden[f"assign_{tag}_anchors"], golden[f"assign_{tag}_gts"]
        pos, neg = golden[f"assign_{tag}_thr"]
        labels, gtids = golden[f"assign_{tag}_labels"], golden[f"assign_{tag}_gtids"]
        a, _, _ = O.assign(anchors, gts, pos, neg, 0.0, off=1.0, mode=1)
        assert np.array_equal(a > 0, labels > 0), tag
        assert np.array_equal(a == 0, labels == 0), tag
        assert np.array_equal(a == -1, labels == -1), tag
        fg = labels > 0
        assert fg.sum() > 0
        assert np.array_equal(a[fg] - 1, gtids[fg]), tag


def test_anchor_grid_order_matches_reference_generator(golden):
    # pointpillars/src/core/box_np_ops.py:453-523: output [D,H,W,S,R,7]; x fastest, per-cell innermost.
    ref = golden["grid_ref"]  # (1,5,7,1,2,7), centres on an 8-px lattice
    H, W, A = 5, 7, 2
    base = np.array([[-1, -1, 1, 1], [-2, -2, 2, 2]], np.float32)
    mine = O.anchor_grid(base, H, W, 8.0).reshape(H, W, A, 4)
    cx = (mine[..., 0] + mine[..., 2]) * 0.5
    cy = (mine[..., 1] + mine[..., 3]) * 0.5
    np.testing.assert_allclose(cx, ref[0, :, :, 0, :, 0], atol=1e-5)
    np.testing.assert_allclose(cy, ref[0, :, :, 0, :, 1], atol=1e-5)
    # per-cell variants are the innermost axis in both
    assert np.array_equal(ref[0, 2, 3, 0, :, 6], np.array([0.0, 1.0], np.float32))
    assert np.array_equal(mine[2, 3, :, 2] - mine[2, 3, :, 0], np.array([2.0, 4.0], np.float32))


def test_iou_offset0_matches_iou_jit_and_image_box_overlap(golden):
    # pointpillars/src/core/box_np_ops.py:639-679 with eps=0.0 and eval_utils.py:118-165 (criterion -1):
    # the no-offset IoU the NMS modes and the YOLO / RCNN post-process use.
    got = O.iou_matrix(golden["iou0_boxes"], golden["iou0_gts"], off=0.0)
    for name in ("iou0_mat_iou_jit", "iou0_mat_image_box_overlap"):
        ref = golden[name]
        assert np.array_equal(got == 0, ref == 0), name          # same pairs overlap (strict iw > 0, ih > 0)
        np.testing.assert_allclose(got, ref, rtol=2e-6, atol=1e-7, err_msg=name)
    assert got[7, 3] == 1.0 and got[8, 4] == 0.0 and got[9, 5] == 0.0


def bilinear_rois(x, y):
    """One-bin, one-sample RoIs whose single sample point is (x, y) on a stride-1 map."""
    h = np.float32(0.5)
    return np.stack([np.zeros_like(x), x - h, y - h, x + h, y + h], 1).astype(np.float32)


def test_roialign_taps_match_bilinear_interpolate_torch(golden):
    # centerpoint/det3d_ms/core/utils/center_utils.py:97-131 on interior points: a 1x1-bin, 1-sample RoIAlign
    # of a stride-1 map is exactly one bilinear read.  (At x >= W-1 that function's clamped x1 makes its weights
    # negative; the RoIAlign edge rules of CONVENTIONS #16 differ there by design, so the fixture stays interior.)
    im, x, y, ref = golden["bilinear_im"], golden["bilinear_x"], golden["bilinear_y"], golden["bilinear_val"]
    feat = np.ascontiguousarray(im.transpose(2, 0, 1)[None])
    rois = bilinear_rois(x, y)
    got = O.roialign_fwd([feat], (1,), rois, P=1, S=1, lvl=np.zeros(len(x), np.int32))
    np.testing.assert_allclose(got.reshape(len(x), -1), ref, rtol=0, atol=2e-5)


def test_assign_mode0_differs_from_the_pinned_mode1_only_where_the_rules_say(golden):
    """Mode 0 (the benchmarked, unpinned rule order) against mode 1 (bit-exact to the reference's create_target_np,
    target_assigner.py:84-134) on the reference-run fixtures.  With F(a) = the gts whose best IoU anchor a ties
    (max > 0): mode 1 gives such an anchor its OWN argmax gt, mode 0 the last gt of F(a) (CONVENTIONS #11 / #12).
    Everywhere else -- positives by threshold, negatives, ignored -- the two modes must be identical, so the
    reference pins mode 0 on every anchor outside that set."""
    total = differing = 0
    # a hand-made case that exercises the differing set: anchor 0 equals gt 0 and is also the only anchor touching gt 1
    crafted = (np.array([[0, 0, 99, 99], [300, 300, 340, 340], [0, 0, 99, 49]], np.float32),
               np.array([[0, 0, 99, 99], [90, 90, 109, 109]], np.float32), np.float32(0.6), np.float32(0.3))
    cases = [(t, golden[f"assign_{t}_anchors"], golden[f"assign_{t}_gts"], *golden[f"assign_{t}_thr"]) for t in "abc"]
    for tag, anchors, gts, pos, neg in cases + [("crafted",) + crafted]:
        a1, _, am = O.assign(anchors, gts, pos, neg, 0.0, off=1.0, mode=1)
        a0, _, _ = O.assign(anchors, gts, pos, neg, 1e-30, off=1.0, mode=0)   # min_pos_iou: "max > 0", as mode 1
        iou = O.iou_matrix(anchors, gts, off=1.0)
        gmax = iou.max(0)
        ties = (iou == gmax[None, :]) & (gmax[None, :] > 0)
        forced = ties.any(1)
        last = np.where(forced, ties.shape[1] - 1 - np.argmax(ties[:, ::-1], 1), -1)
        differ = forced & (last != am)
        assert np.array_equal(a0[~differ], a1[~differ]), tag
        assert np.array_equal(a0[differ], last[differ] + 1) and np.array_equal(a1[differ], am[differ] + 1), tag
        if tag == "crafted":
            assert differ.tolist() == [True, False, False] and a0.tolist() == [2, 0, -1] and a1.tolist() == [1, 0, -1]
            continue
        # on the reference-run fixtures themselves mode 0 reproduces the reference's labels and matched gt ids
        labels, gtids = golden[f"assign_{tag}_labels"], golden[f"assign_{tag}_gtids"]
        fg = (labels > 0) & ~differ
        assert np.array_equal(a0 > 0, labels > 0) and np.array_equal(a0 == 0, labels == 0), tag
        assert np.array_equal(a0[fg] - 1, gtids[fg]), tag
        total += len(a0)
        differing += int(differ.sum())
    assert total > 1000 and differing < total // 20


def test_default_nms_mode_matches_the_references_iou_normal(golden):
    """The DEFAULT NMS mode (offset 0, strict >, union guard: what `Proposal`, the YOLO / RCNN post-process and
    `NmsNormalGpu` run) against the reference's own `iou_normal` (iou3d_nms_kernel.cu:347-358) executed on the host
    (oracle/ref_cu_device_harness.cpp).  Lattice cases a-c: keep lists AND IoU values bit-identical; the off-lattice
    case d (centre/size <-> corner conversion rounds) keeps the same boxes and agrees to 1e-6."""
    from oracle import bev
    eps = float(golden["ioun_eps"][0])
    assert eps == float(np.float32(1e-8))
    for tag in "abcd":
        xy, b7, ref_iou = golden[f"ioun_{tag}_xyxy"], golden[f"ioun_{tag}_box7"], golden[f"ioun_{tag}_iou"]
        got = bev.iou_normal(b7[:96], b7)
        if tag != "d":
            assert np.array_equal(got, ref_iou), tag                       # numpy restatement used by the BEV tests
        else:
            np.testing.assert_allclose(got, ref_iou, rtol=1e-6, atol=1e-7)
        for thr in (0.3, 0.7):
            ref_keep = golden[f"ioun_{tag}_{thr}_keep"]
            mask = O.nms(xy, np.float32(thr), off=0.0, inclusive=False, union_eps=eps)
            assert np.array_equal(_keep_from_mask(mask), ref_keep), (tag, thr)
            # the numpy oracle the NmsNormalGpu tests use (tests/test_gpu_bev.py)
            assert np.array_equal(bev.greedy_from_iou(bev.iou_normal(b7, b7), np.float32(thr)), ref_keep), (tag, thr)


def test_reference_gpu_file_and_cpu_file_agree_bit_for_bit():
    """The rotated-BEV symbols are pinned to the reference's CPU file (iou-bev-nms-org.cpp, compiled as it lies).  The symbols
    they replace live in the GPU file (iou3d_nms_kernel.cu): its __device__ functions, cut out and run on the host, give the
    same IoU bits and the keep lists the fixtures hold for the strict '>' symbols -- one arithmetic, two files."""
    import os
    import pytest
    from oracle import bev
    if not (bev.have_ref() and bev.have_ref_cu()):
        pytest.skip("oracle/_ref not built (needs /root/reference once)")
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "bev_golden.npz"))
    a, b = g["iou_a"], g["iou_b"]
    iou_cu = bev.ref_cu_pairs("iou_bev", a, b)
    assert np.array_equal(iou_cu, g["iou_ref"]) and np.array_equal(iou_cu, bev.ref_iou_bev(a, b))
    assert (iou_cu > 0).sum() > 1000
    ov = bev.ref_cu_pairs("box_overlap", a, b)
    sa, sb = a[:, 3] * a[:, 4], b[:, 3] * b[:, 4]
    assert np.array_equal(ov / np.maximum(sa[:, None] + sb[None] - ov, np.float32(1e-8)), iou_cu)
    for tag in "abcd":
        boxes, thr = g[f"nms_{tag}_boxes"], g[f"nms_{tag}_thr"][0]
        n = int(g[f"nms_{tag}_count_gt"][0])
        assert np.array_equal(bev.ref_cu_nms(boxes, thr, rotated=True), g[f"nms_{tag}_keep_gt"][:n]), tag


def test_topk_matches_the_references_numpy_topk(golden):
    # pointpillars/src/core/nms.py:66-83 (`topk_`: argpartition + argsort, returns the K-1 best) on unique scores:
    # values and indices of the sorted top-k (a3); the tie order stays a decision (CONVENTIONS #3).
    sc, rv, ri = golden["topk_scores"], golden["topk_ref_vals"], golden["topk_ref_idx"]
    v, i = O.topk(sc, len(ri))
    assert len(ri) == 1000 and np.array_equal(i, ri) and np.array_equal(v, rv)


def test_iou_both_offsets_match_kitti_common_iou(golden):
    # pointpillars/src/data/kitti_common.py:10-73 (vectorised numpy, add1 False / True) on the iou0 fixture boxes
    a, g = golden["iou0_boxes"], golden["iou0_gts"]
    for off, name in ((0.0, "iou0_mat_kitti"), (1.0, "iou1_mat_kitti")):
        got, ref = O.iou_matrix(a, g, off=off), golden[name]
        assert np.array_equal(got == 0, ref == 0), name
        np.testing.assert_allclose(got, ref, rtol=2e-6, atol=1e-7, err_msg=name)


def test_score_activations_match_the_references_numpy_functions(golden):
    # pointpillars/src/predict.py:98-112: sigmoid = 1 / (1 + 1 / exp(x)), softmax = exp(x) / sum exp(x) (no max shift).
    # The oracle's own exp (CONVENTIONS #6) and its max-shifted softmax agree to a few ulp.
    x = golden["act_x"]
    np.testing.assert_allclose(O.sigmoid(x), golden["act_sigmoid_ref"], rtol=1e-6, atol=0)
    np.testing.assert_allclose(O.softmax_rows(x), golden["act_softmax_ref"], rtol=4e-6, atol=0)
