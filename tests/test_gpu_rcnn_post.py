"""RCNN-head post-process + BoundingBoxEncode ('next' row 4) against the CPU oracle: boxes / scores / labels /
keep indices / counts bit-exact; encode within the stated FP tolerance (logf)."""
import numpy as np
import pytest
import torch

import oracle as O
from minddet_b200 import BoundingBoxEncode, RcnnPostProcess, synth

pytestmark = pytest.mark.gpu


def dev(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


def head_outputs(rng, B, P, nc):
    rois = np.stack([synth.rand_boxes(rng, P, cluster=25, sigma=10.0) for _ in range(B)])
    logits = rng.normal(0, 1.0, (B, P, nc + 1)).astype(np.float32)
    logits[:, :, 0] += 3.0                                   # background dominates ...
    hot = rng.uniform(0, 1, (B, P)) < 0.35                   # ... except on a third of the RoIs
    cls = rng.integers(1, nc + 1, (B, P))
    bi, pi = np.nonzero(hot)
    logits[bi, pi, cls[bi, pi]] += rng.uniform(3, 9, len(bi)).astype(np.float32)
    deltas = rng.normal(0, 0.6, (B, P, (nc + 1) * 4)).astype(np.float32)
    valid = (rng.uniform(0, 1, (B, P)) > 0.1).astype(np.uint8)
    return rois.astype(np.float32), valid, logits, deltas


@pytest.mark.parametrize("nc,P,score_thr,nms_pre,max_det", [(80, 1000, 0.05, 2048, 100), (80, 1000, 0.001, 2048, 100),
                                                            (3, 300, 0.3, 128, 20), (1, 64, 0.5, 2048, 100)])
def test_rcnn_post_bit_exact(nc, P, score_thr, nms_pre, max_det):
    rng = np.random.default_rng(700 + nc + P)
    B = 3
    rois, valid, logits, deltas = head_outputs(rng, B, P, nc)
    op = RcnnPostProcess((800, 1344), score_thr=score_thr, iou_thr=0.5, max_det=max_det, nms_pre=nms_pre)
    out, keep_idx, count = op(dev(rois), dev(valid).bool(), dev(logits), dev(deltas))
    out, keep_idx, count = out.cpu().numpy(), keep_idx.cpu().numpy(), count.cpu().numpy()
    total = 0
    for b in range(B):
        ro, ri, rc = O.rcnn_post(rois[b], valid[b], logits[b], deltas[b], 800, 1344, score_thr=score_thr, nms_pre=nms_pre,
                                 iou_thr=0.5, max_det=max_det)
        assert count[b] == rc, (b, count[b], rc)
        assert np.array_equal(keep_idx[b], ri), b
        assert np.array_equal(out[b], ro), b
        total += rc
    assert total > 0
    # rois given as (B,P,5) [batch,x1,y1,x2,y2] (what BboxAssignSampleForRcnn / Proposal hand over) give the same result
    rois5 = np.concatenate([np.zeros((B, P, 1), np.float32), rois], 2)
    out5, _, count5 = op(dev(rois5), dev(valid).bool(), dev(logits), dev(deltas))
    assert np.array_equal(out5.cpu().numpy(), out) and np.array_equal(count5.cpu().numpy(), count)


def test_softmax_scores_sum_to_one_and_match_float64():
    rng = np.random.default_rng(3)
    x = rng.normal(0, 4, (500, 81)).astype(np.float32)
    p = O.softmax_rows(x)
    e = np.exp(x.astype(np.float64) - x.max(1, keepdims=True))
    np.testing.assert_allclose(p, e / e.sum(1, keepdims=True), rtol=1e-5, atol=1e-9)


def test_encode_matches_oracle_and_inverts_decode():
    rng = np.random.default_rng(4)
    K = 5000
    props = synth.rand_boxes(rng, K)
    gts = synth.rand_boxes(rng, K)
    enc = BoundingBoxEncode(means=(0.0, 0.0, 0.0, 0.0), stds=(0.1, 0.1, 0.2, 0.2))
    got = enc(dev(props), dev(gts)).cpu().numpy()
    ref = O.encode(props, gts, means=(0, 0, 0, 0), stds=(0.1, 0.1, 0.2, 0.2))
    np.testing.assert_allclose(got, ref, rtol=1e-5, atol=1e-5)     # logf: libm vs CUDA, FP tolerance (CONVENTIONS #9)
    # decode(encode(gt)) == gt up to rounding when no clamp / clip binds
    from minddet_b200 import BoundingBoxDecode
    ok = (np.abs(got[:, 2:]) * 0.2 < 4.0).all(1)
    dec = BoundingBoxDecode((800, 1344), stds=(0.1, 0.1, 0.2, 0.2))(dev(props), dev(got)).cpu().numpy()
    np.testing.assert_allclose(dec[ok], gts[ok], rtol=1e-4, atol=2e-2)
