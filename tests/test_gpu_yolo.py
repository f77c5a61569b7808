"""GPU parity of the YOLOv8 post-process (SURVEY.md 8(f) row 1 / BASELINE config 5) against the CPU oracle:
decoded boxes, scores and labels bit-exact (every op individually rounded on both sides); NMS keep indices,
counts and candidate order bit-exact, including the dense-crowd stress set of section 8(d)."""
import numpy as np
import pytest
import torch

import oracle as O
from minddet_b200 import YoloV8PostProcess

pytestmark = pytest.mark.gpu
SHAPES, STRIDES = [(80, 80), (40, 40), (20, 20)], (8, 16, 32)


def dev(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


def crowd_pred(rng, B, nc=80, objects=300, sigma=8.0):
    """raw head output whose decoded boxes cluster around `objects` centres (long suppression chains)"""
    A = sum(h * w for h, w in SHAPES)
    pred = rng.normal(0.0, 1.0, (B, 64 + nc, A)).astype(np.float32)
    pred[:, 64:] = rng.normal(-6.0, 1.0, (B, nc, A)).astype(np.float32)
    ctr = rng.uniform(40, 600, (B, objects, 2))
    cls = rng.integers(0, nc, (B, objects))
    a0 = 0
    for (h, w), s in zip(SHAPES, STRIDES):
        ys, xs = np.meshgrid(np.arange(h), np.arange(w), indexing="ij")
        cx, cy = (xs.reshape(-1) + 0.5) * s, (ys.reshape(-1) + 0.5) * s
        for b in range(B):
            d2 = (cx[:, None] - ctr[b, :, 0][None]) ** 2 + (cy[:, None] - ctr[b, :, 1][None]) ** 2
            near = d2.argmin(1)
            hit = d2.min(1) < (3 * sigma) ** 2
            idx = np.nonzero(hit)[0]
            pred[b, 64 + cls[b, near[idx]], a0 + idx] = rng.normal(2.0, 1.5, len(idx)).astype(np.float32)
            # sharpen the DFL bins so that boxes of one object overlap heavily
            for side in range(4):
                k = rng.integers(2, 6, len(idx))
                pred[b, side * 16 + k, a0 + idx] += 6.0
        a0 += h * w
    return pred


@pytest.mark.parametrize("nc,B", [(80, 3), (1, 2), (7, 1)])
def test_decode_bit_exact(nc, B):
    rng = np.random.default_rng(500 + nc)
    A = sum(h * w for h, w in SHAPES)
    pred = rng.normal(0.0, 2.5, (B, 64 + nc, A)).astype(np.float32)
    pred[:, :64, :100] *= 12.0               # saturating softmax inputs
    op = YoloV8PostProcess(SHAPES, STRIDES)
    got = op.decode(dev(pred)).cpu().numpy()
    for b in range(B):
        assert np.array_equal(got[b], O.yolo_decode(pred[b], SHAPES, STRIDES)), b


def test_decode_scalar_path_and_float64_tolerance():
    rng = np.random.default_rng(9)
    shapes = [(5, 7), (3, 3)]                # A = 44 + ... not a multiple of 4 after slicing -> 1-anchor path
    A = 35 + 9
    pred = rng.normal(0.0, 2.0, (2, 64 + 3, A + 1)).astype(np.float32)[:, :, :A].copy()
    pred = np.ascontiguousarray(pred[:, :, :43])
    shapes = [(5, 7), (2, 4)]
    op = YoloV8PostProcess(shapes, (8, 16))
    got = op.decode(dev(pred)).cpu().numpy()
    ref = O.yolo_decode(pred[0], shapes, (8, 16))
    assert np.array_equal(got[0], ref)
    x = pred[0, :16, 0].astype(np.float64)
    p = np.exp(x - x.max()); p /= p.sum()
    assert abs((0.5 - (p * np.arange(16)).sum()) * 8 - got[0, 0, 0]) <= 1e-5 * max(1.0, abs(got[0, 0, 0]))   # north_star: 1e-5 relative


@pytest.mark.parametrize("conf,agnostic,nms_pre,max_det", [(0.25, False, 2048, 300), (0.001, False, 2048, 300),
                                                           (0.25, True, 1000, 100), (0.9999, False, 64, 10)])
def test_postprocess_dense_crowd_bit_exact(conf, agnostic, nms_pre, max_det):
    rng = np.random.default_rng(77)
    B = 4
    pred = crowd_pred(rng, B)
    op = YoloV8PostProcess(SHAPES, STRIDES, conf_thr=conf, iou_thr=0.7, agnostic=agnostic, nms_pre=nms_pre, max_det=max_det)
    out, keep_idx, count = op(dev(pred))
    out, keep_idx, count = out.cpu().numpy(), keep_idx.cpu().numpy(), count.cpu().numpy()
    kept_total = 0
    for b in range(B):
        dets = O.yolo_decode(pred[b], SHAPES, STRIDES)
        ro, ri, rc = O.yolo_nms(dets, conf, nms_pre, 0.7, agnostic, max_det)
        assert count[b] == rc, (b, count[b], rc)
        assert np.array_equal(keep_idx[b], ri), b
        assert np.array_equal(out[b], ro), b
        kept_total += rc
    if conf <= 0.25:
        assert kept_total > 50            # the stress set really exercises suppression chains


@pytest.mark.parametrize("nc,conf", [(1, 0.25), (3, 0.05), (80, 1.5)])
def test_postprocess_label_grouping_edges(nc, conf):
    """The label-major NMS path at its edges: one class only (a single diagonal block), three classes (long runs that
    cross many 64-row tiles), and a threshold nothing passes (zero candidates: dynamic row count 0)."""
    rng = np.random.default_rng(78 + nc)
    B = 3
    pred = crowd_pred(rng, B, nc=nc)
    op = YoloV8PostProcess(SHAPES, STRIDES, conf_thr=conf, iou_thr=0.7, nms_pre=2048, max_det=300)
    out, keep_idx, count = op(dev(pred))
    out, keep_idx, count = out.cpu().numpy(), keep_idx.cpu().numpy(), count.cpu().numpy()
    for b in range(B):
        dets = O.yolo_decode(pred[b], SHAPES, STRIDES)
        ro, ri, rc = O.yolo_nms(dets, conf, 2048, 0.7, False, 300)
        assert count[b] == rc, (b, count[b], rc)
        assert np.array_equal(keep_idx[b], ri), b
        assert np.array_equal(out[b], ro), b
    if conf > 1.0:
        assert count.sum() == 0


def test_full_batch_properties():
    """config 5 size (B=64): idempotence -- running NMS on its own survivors keeps every one of them."""
    rng = np.random.default_rng(5)
    pred = crowd_pred(rng, 8)
    pred = np.concatenate([pred] * 8)       # 64 images
    op = YoloV8PostProcess(SHAPES, STRIDES, conf_thr=0.25, nms_pre=2048, max_det=300)
    dets = op.decode(dev(pred))
    out, keep_idx, count = op.nms(dets)
    assert torch.equal(out[:8], out[8:16]) and torch.equal(count[:8], count[56:])
    pad = torch.zeros(64, 300, 6, device="cuda")
    pad[:, :, :] = out
    out2, _, count2 = op.nms(pad)
    assert torch.equal(count2, count)
    n = int(count.max())
    assert torch.equal(out2[:, :n], out[:, :n])
