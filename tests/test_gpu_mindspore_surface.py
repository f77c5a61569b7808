"""minddet_b200/mindspore_ops.py executed end to end: the cells a minddet maintainer would drop into the MindSpore graph,
run on the GPU through a stub ``mindspore`` (tests/ms_stub.py) whose ``ops.Custom`` makes the real aot call, and checked
against the oracle.  VERDICT r1 'next 8': the file used to be exercised for syntax only."""
import importlib

import numpy as np
import pytest
import torch

import oracle as O
from minddet_b200 import synth

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def M():
    from tests import ms_stub
    ms_stub.install()
    import minddet_b200.mindspore_ops as m
    m = importlib.reload(m)
    assert m.HAVE_MINDSPORE
    return m


def dev(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


def host(t):
    return t.detach().cpu().numpy()


def test_anchor_decode_encode_topk_nms(M):
    rng = np.random.default_rng(1)
    gen = M.AnchorGenerator(8, [8], [0.5, 1.0, 2.0])
    anchors = host(gen.grid_anchors((25, 42), 8))
    assert np.array_equal(anchors, O.anchor_grid(gen.base_anchors, 25, 42, 8))
    deltas = rng.normal(0, 0.3, anchors.shape).astype(np.float32)
    boxes = host(M.BoundingBoxDecode((200, 336))(dev(anchors), dev(deltas)))
    assert np.array_equal(boxes, O.decode(anchors, deltas, 200, 336))
    enc = host(M.BoundingBoxEncode()(dev(anchors), dev(boxes + 1)))
    np.testing.assert_allclose(enc, O.encode(anchors, boxes + 1), rtol=1e-5, atol=1e-6)
    scores = rng.normal(0, 1, (2, 3, 25, 42)).astype(np.float32)
    vals, idx = M.TopKPerLevel(300, apply_sigmoid=True)(dev(scores))
    for b in range(2):
        rv, ri = O.topk(O.level_scores(scores[b]), 300)
        assert np.array_equal(host(idx[b]), ri) and np.array_equal(host(vals[b]), rv)
    sb = np.concatenate([synth.rand_boxes(rng, 500, cluster=20), np.sort(rng.uniform(0, 1, 500))[::-1, None].astype(np.float32)], 1)
    keep_idx, mask, count = M.NMSWithMask(0.5)(dev(sb))
    rk = O.nms(sb, 0.5)                                     # keep mask over the score-sorted rows
    assert int(count) == int(rk.sum()) and np.array_equal(host(mask).astype(np.uint8), rk)


def test_proposal_assign_extract(M):
    B, C = 2, 32
    strides, shapes = synth.STRIDES, [(25, 42), (13, 21), (7, 11), (4, 6), (2, 3)]
    bases = synth.base_anchor_sets(strides)
    logits, deltas = synth.rpn_head_outputs(B, shapes, 3, seed=7)
    gts, labels, gvalid = synth.gt_boxes(B, G=16, max_valid=6, img_h=100, img_w=168, seed=7)
    prop = M.Proposal(B, (100, 168), strides, bases, nms_pre=300, max_num=256)
    props, pmask = prop(tuple(dev(x) for x in logits), tuple(dev(x) for x in deltas))
    pcfg = O.proposal_cfg(100, 168, nms_pre=300, max_num=256)
    refs = [O.proposal_image([(logits[l][b], deltas[l][b], bases[l], strides[l]) for l in range(5)], pcfg) for b in range(B)]
    for b in range(B):
        assert np.array_equal(host(props[b]), refs[b]["props"]) and np.array_equal(host(pmask[b]).astype(np.uint8), refs[b]["mask"])
    # RPN targets on the anchors of all levels: two calls = steps 0 and 1 of the member seed tensor
    anchors = np.concatenate([O.anchor_grid(b_, h, w, s) for b_, (h, w), s in zip(bases, shapes, strides)])
    rpn = M.BboxAssignSample(0.7, 0.3, 0.3, 16, 32, 32, seed=11)
    valid = torch.ones(anchors.shape[0], dtype=torch.bool, device="cuda")
    for step in range(2):
        out = rpn(dev(gts), dev(gvalid).bool(), dev(anchors), valid)
        acfg = O.assign_cfg(0.7, 0.3, 0.3, 16, 32, 32, seed=11, step=step)
        for b in range(B):
            r = O.assign_sample_rpn(anchors, gts[b], gvalid[b], acfg, b)
            assert np.array_equal(host(out[0][b]), r["assigned"]) and np.array_equal(host(out[3][b]), r["neg_idx"]), (step, b)
    rcnn = M.BboxAssignSampleForRcnn(0.5, 0.5, 0.5, 16, 48, 64, seed=3)
    out = rcnn(dev(gts), dev(labels), pmask, props, dev(gvalid).bool())
    rcfg = O.assign_cfg(0.5, 0.5, 0.5, 16, 48, 64, stds=(0.1, 0.1, 0.2, 0.2), seed=3)
    for b in range(B):
        r = O.assign_sample_rcnn(refs[b]["props"][:, :4], refs[b]["mask"], gts[b], labels[b], gvalid[b], rcfg, b)
        assert np.array_equal(host(out[5][b]), r["sel_idx"]) and np.array_equal(host(out[2][b]), r["labels"])
    rois = out[0].reshape(-1, 5)
    feats = synth.features(B, shapes[:4], C, seed=7)
    ext = M.SingleRoIExtractor(7, 2, strides[:4], 56)
    ft = [dev(f).requires_grad_(True) for f in feats]
    y = ext(rois, *ft)
    rois_np = host(rois)
    np.testing.assert_allclose(host(y), O.roialign_fwd(feats, strides[:4], rois_np), rtol=1e-5, atol=1e-6)
    assert np.array_equal(host(ext.map_roi_levels(rois)), O.roi_levels(rois_np, 56.0, 4))
    dout = np.random.default_rng(2).uniform(-1, 1, tuple(y.shape)).astype(np.float32)
    y.backward(dev(dout))                       # -> the cell's bprop -> MdRoiAlignBwd
    dref = O.roialign_bwd([f.shape for f in feats], strides[:4], rois_np, dout)
    for l in range(4):
        np.testing.assert_allclose(host(ft[l].grad), dref[l], rtol=1e-5, atol=1e-5 * max(1.0, np.abs(dref[l]).max()))
    # the two-op bprop: plan from the RoIs (+ feature shapes), then the planned backward
    plan = M.RoIAlignGradPlan(strides[:4], 56)(rois, *[dev(f) for f in feats])
    grads = M.RoIAlignGradPlanned([f.shape for f in feats], strides[:4], 56)(rois, dev(dout), plan)
    for l in range(4):
        np.testing.assert_allclose(host(grads[l]), dref[l], rtol=1e-5, atol=1e-5 * max(1.0, np.abs(dref[l]).max()))


def test_next_row_cells(M):
    rng = np.random.default_rng(4)
    shapes, strides = [(8, 8), (4, 4), (2, 2)], (8, 16, 32)
    A = sum(h * w for h, w in shapes)
    pred = rng.normal(0, 1, (2, 64 + 5, A)).astype(np.float32)
    pred[:, 64:] -= 1.0
    dets, keep, cnt = M.YoloV8PostProcess(shapes, strides, conf_thr=0.2, nms_pre=64, max_det=20)(dev(pred))
    for b in range(2):
        ro, ri, rc = O.yolo_nms(O.yolo_decode(pred[b], shapes, strides), 0.2, 64, 0.7, False, 20)
        assert int(cnt[b]) == rc and np.array_equal(host(keep[b]), ri) and np.array_equal(host(dets[b]), ro)
    masks = (rng.uniform(0, 1, (1, 3, 60, 80)) > 0.5).astype(np.uint8)
    rois = np.array([[0, 5, 5, 40, 30], [0, 20, 10, 70, 55]], np.float32)
    gi = np.array([1, 2], np.int32)
    mt = M.MaskTargets(28, 2)(dev(masks).bool(), dev(rois), dev(gi))
    assert np.array_equal(host(mt).astype(np.uint8), O.mask_targets(masks[0], rois[:, 1:], gi, 28, 2))
    P, nc1 = 50, 4
    rb = synth.rand_boxes(rng, P, img_w=336.0, img_h=200.0, smin=8, smax=120)[None]
    logits = rng.normal(0, 2, (1, P, nc1)).astype(np.float32)
    dl = rng.normal(0, 0.5, (1, P, nc1 * 4)).astype(np.float32)
    rv = np.ones((1, P), np.uint8)
    dets, keep, cnt = M.RcnnPostProcess((200, 336), score_thr=0.05, iou_thr=0.5, max_det=30, nms_pre=128)(dev(rb), dev(rv).bool(), dev(logits), dev(dl))
    ro = O.rcnn_post(rb[0], rv[0], logits[0], dl[0], 200, 336, score_thr=0.05, iou_thr=0.5, nms_pre=128, max_det=30)
    assert int(cnt[0]) == ro[2] and np.array_equal(host(dets[0]), ro[0])
