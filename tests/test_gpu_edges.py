"""Edge cases through the C-ABI: empty inputs, nothing-to-keep, everything-invalid.  The library must neither fault
nor leave outputs uninitialised."""
import numpy as np
import pytest
import torch

import oracle as O
from minddet_b200 import (BboxAssignSample, NMSWithMask, Proposal, RcnnPostProcess, SingleRoIExtractor, TopKPerLevel,
                          YoloV8PostProcess, synth)
from minddet_b200.bev_ops import BoxesIouBevGpu, NmsBevGpu, NumGpu

pytestmark = pytest.mark.gpu


def dev(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


def test_zero_rois_forward_and_backward():
    feats = [torch.rand(1, 8, h, w, device="cuda") for h, w in synth.level_shapes()[:4]]
    ext = SingleRoIExtractor()
    rois = torch.zeros(0, 5, device="cuda")
    out = ext(rois, *feats)
    assert out.shape == (0, 8, 7, 7)
    grads = ext._backward(rois, torch.zeros(0, 8, 7, 7, device="cuda"), [tuple(f.shape) for f in feats])
    torch.cuda.synchronize()
    assert all(float(g.abs().sum()) == 0.0 for g in grads)        # dX is still zero-filled


def test_rois_with_bad_batch_index_are_ignored_by_reference_semantics():
    # RoIs that map entirely outside the map produce zeros (every sample invalid), never a fault
    feats = [torch.rand(2, 8, h, w, device="cuda") for h, w in synth.level_shapes()[:4]]
    rois = dev(np.array([[0, 5000, 5000, 5100, 5100], [1, -900, -900, -800, -800]], np.float32))
    out = SingleRoIExtractor()(rois, *feats)
    assert float(out.abs().sum()) == 0.0


def test_no_valid_gt_and_no_valid_anchor():
    anchors = np.concatenate([O.anchor_grid(b, h, w, s) for b, (h, w), s in
                              zip(synth.base_anchor_sets()[2:], synth.level_shapes()[2:], synth.STRIDES[2:])])
    N = anchors.shape[0]
    gts = np.zeros((2, 4, 4), np.float32)
    gts[1, 0] = [100, 100, 300, 300]
    gvalid = np.array([[0, 0, 0, 0], [1, 0, 0, 0]], np.uint8)
    valid = np.ones(N, np.uint8)
    op = BboxAssignSample(0.7, 0.3, 0.3, 16, 32, 32, seed=5)
    out = op(dev(gts), dev(gvalid).bool(), dev(anchors), dev(valid).bool())
    cfg = O.assign_cfg(0.7, 0.3, 0.3, 16, 32, 32, seed=5)
    for b in range(2):
        ref = O.assign_sample_rpn(anchors, gts[b], gvalid[b], cfg, b, valid=valid)
        assert np.array_equal(out["assigned"][b].cpu().numpy(), ref["assigned"])
        assert np.array_equal(out["pos_idx"][b].cpu().numpy(), ref["pos_idx"]) and int(out["num_pos"][b]) == ref["num_pos"]
        assert np.array_equal(out["neg_idx"][b].cpu().numpy(), ref["neg_idx"])
    assert int(out["num_pos"][0]) == 0
    # every anchor invalid -> everything ignored, no samples
    out2 = op(dev(gts), dev(gvalid).bool(), dev(anchors), dev(np.zeros(N, np.uint8)).bool())
    assert int((out2["assigned"] != -1).sum()) == 0 and int(out2["pos_valid"].sum()) == 0 and int(out2["neg_valid"].sum()) == 0


def test_nms_single_box_and_identical_boxes():
    one = dev(np.array([[10, 10, 50, 50, 0.9]], np.float32))
    keep, mask, cnt = NMSWithMask(0.5)(one)
    assert int(cnt) == 1 and bool(mask[0]) and int(keep[0]) == 0
    same = dev(np.tile(np.array([[10, 10, 50, 50, 0.9]], np.float32), (130, 1)))
    keep, mask, cnt = NMSWithMask(0.5)(same)
    assert int(cnt) == 1 and int(mask.sum()) == 1            # the first box suppresses the other 129 across 3 mask words


def test_topk_k_larger_than_n_and_all_equal_scores():
    x = torch.zeros(2, 3, 4, 5, device="cuda")
    vals, idx = TopKPerLevel(2000)(x)
    assert idx.shape == (2, 60)
    assert np.array_equal(idx[0].cpu().numpy(), np.arange(60))       # ties -> lower index first


def test_postprocess_with_nothing_to_keep():
    shapes = [(8, 8), (4, 4)]
    pred = torch.full((2, 64 + 5, 80), -20.0, device="cuda")        # every score ~ 2e-9 < conf
    out, keep_idx, count = YoloV8PostProcess(shapes, (8, 16), conf_thr=0.25, nms_pre=64, max_det=10)(pred)
    assert int(count.sum()) == 0 and float(out.abs().sum()) == 0.0 and bool((keep_idx == -1).all())
    rois = torch.rand(2, 50, 4, device="cuda") * 100
    logits = torch.zeros(2, 50, 4, device="cuda")
    logits[:, :, 0] = 30.0                                           # background everywhere
    out, keep_idx, count = RcnnPostProcess((200, 200), max_det=10, nms_pre=64)(rois, torch.ones(2, 50, dtype=torch.bool, device="cuda"),
                                                                              logits, torch.zeros(2, 50, 16, device="cuda"))
    assert int(count.sum()) == 0 and bool((keep_idx == -1).all())
    out, keep_idx, count = RcnnPostProcess((200, 200), score_thr=0.0, max_det=10, nms_pre=64)(
        rois, torch.zeros(2, 50, dtype=torch.bool, device="cuda"), torch.zeros(2, 50, 4, device="cuda"), torch.zeros(2, 50, 16, device="cuda"))
    assert int(count.sum()) == 0                                     # every RoI invalid


def test_bev_degenerate_inputs():
    z = torch.zeros(5, 7, device="cuda")
    iou = BoxesIouBevGpu()(z, z)
    assert bool(torch.isfinite(iou).all()) and float(iou.abs().sum()) == 0.0     # zero-area boxes: 0 / eps
    keep, cnt = NmsBevGpu()(z, torch.tensor([0.2], device="cuda"))
    assert int(cnt) == 0                                             # all zero-area boxes are dropped first
    keep64, num = NumGpu()(z, torch.tensor([0.2], device="cuda"))
    assert int(num) == 5                                             # the '>' symbol keeps them (IoU 0 is not > thr)


def test_proposal_fewer_anchors_than_nms_pre():
    shapes = [(3, 4), (2, 2)]
    strides = (8, 16)
    bases = synth.base_anchor_sets(strides)
    rng = np.random.default_rng(3)
    logits = [rng.normal(0, 1, (1, 3, h, w)).astype(np.float32) for h, w in shapes]
    deltas = [rng.normal(0, 0.1, (1, 12, h, w)).astype(np.float32) for h, w in shapes]
    prop = Proposal((32, 32), strides, bases, nms_pre=100, max_num=64)
    props, pmask = prop([dev(x) for x in logits], [dev(x) for x in deltas])
    cfg = O.proposal_cfg(32, 32, nms_pre=100, max_num=64)
    ref = O.proposal_image([(logits[l][0], deltas[l][0], bases[l], strides[l]) for l in range(2)], cfg)
    assert np.array_equal(props[0].cpu().numpy(), ref["props"]) and np.array_equal(pmask[0].cpu().numpy().astype(np.uint8), ref["mask"])


def test_proposal_single_level_takes_the_one_lane_path():
    """L = 1: no helper stream, one top-k + NMS launch (the two-lane split needs at least two levels)."""
    shapes, strides = [(25, 42)], (16,)
    bases = synth.base_anchor_sets(strides)
    logits, deltas = synth.rpn_head_outputs(2, shapes, 3, 11)
    prop = Proposal((400, 672), strides, bases, nms_pre=300, max_num=200)
    props, pmask = prop([dev(x) for x in logits], [dev(x) for x in deltas])
    cfg = O.proposal_cfg(400, 672, nms_pre=300, max_num=200)
    for b in range(2):
        ref = O.proposal_image([(logits[0][b], deltas[0][b], bases[0], strides[0])], cfg)
        assert np.array_equal(props[b].cpu().numpy(), ref["props"])
        assert np.array_equal(pmask[b].cpu().numpy().astype(np.uint8), ref["mask"])


def test_proposal_two_lanes_inside_a_cuda_graph():
    """MdProposal forks onto the workspace's helper stream and joins back with events: the same call must be capturable
    (the bench replays it inside a CUDA graph) and the replay must reproduce the eager result bit for bit."""
    shapes, strides = [(50, 84), (25, 42), (13, 21)], (8, 16, 32)
    bases = synth.base_anchor_sets(strides)
    logits, deltas = synth.rpn_head_outputs(2, shapes, 3, 12)
    prop = Proposal((400, 672), strides, bases, nms_pre=1000, max_num=1000)
    ls, ds = [dev(x) for x in logits], [dev(x) for x in deltas]
    side = torch.cuda.Stream()
    with torch.cuda.stream(side):
        eager_props, eager_mask = prop(ls, ds)          # warm-up on the capture stream: workspace + helper stream exist
        side.synchronize()
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph, stream=side):
            g_props, g_mask = prop(ls, ds)
        g_props.zero_()
        g_mask.zero_()
        graph.replay()
        side.synchronize()
    assert torch.equal(g_props, eager_props) and torch.equal(g_mask, eager_mask)
    cfg = O.proposal_cfg(400, 672, nms_pre=1000, max_num=1000)
    ref = O.proposal_image([(logits[l][0], deltas[l][0], bases[l], strides[l]) for l in range(3)], cfg)
    assert np.array_equal(g_props[0].cpu().numpy(), ref["props"])
