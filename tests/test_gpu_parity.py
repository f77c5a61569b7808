"""GPU parity tests proper: every call goes through the C-ABI (libmdregion.so aot symbols) and is
compared with the CPU oracle on the same seeded inputs.  Integer outputs bit-exact; decoded boxes and
RoIAlign forward bit-exact (same rounded op sequence); encode targets and RoIAlign backward within
the stated FP tolerance.  Run with `pytest -m gpu` on the B200 box."""
import numpy as np
import pytest
import torch

import oracle as O
from minddet_b200 import (AnchorGenerator, BboxAssignSample, BboxAssignSampleForRcnn, BoundingBoxDecode,
                          NMSWithMask, Proposal, SingleRoIExtractor, TopKPerLevel)
from minddet_b200 import synth

pytestmark = pytest.mark.gpu


def dev(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


def host(t):
    return t.detach().cpu().numpy()


# ---------------------------------------------------------------------------------------------- a1
def test_anchor_grid_bit_exact():
    for stride, (h, w) in zip(synth.STRIDES, synth.level_shapes()):
        g = AnchorGenerator(stride, [8], [0.5, 1.0, 2.0])
        got = host(g.grid_anchors((h, w), stride))
        ref = O.anchor_grid(g.base_anchors, h, w, stride)
        assert got.shape == ref.shape and np.array_equal(got, ref)
    g = AnchorGenerator(16, [2, 4], [1.0, 3.0], scale_major=False)
    assert np.array_equal(host(g.grid_anchors((3, 5), 16)), O.anchor_grid(g.base_anchors, 3, 5, 16))


# ---------------------------------------------------------------------------------------------- a2
def test_decode_rows_and_level_bit_exact():
    rng = np.random.default_rng(10)
    bases = synth.base_anchor_sets()
    dec = BoundingBoxDecode((800, 1344), means=(0.0, 0.01, 0.0, -0.02), stds=(1.0, 0.5, 1.0, 2.0))
    for stride, base, (h, w) in zip(synth.STRIDES[1:], bases[1:], synth.level_shapes()[1:]):
        B, A = 2, base.shape[0]
        d = rng.normal(0, 1.0, (B, 4 * A, h, w)).astype(np.float32)
        d[:, 2::4] *= 3.0   # exercise the wh clamp
        got = host(dec.decode_level(dev(d), dev(base), stride))
        for b in range(B):
            ref = O.decode_level_nchw(base, h, w, stride, d[b], 800, 1344, means=(0.0, 0.01, 0.0, -0.02), stds=(1.0, 0.5, 1.0, 2.0))
            assert np.array_equal(got[b], ref), (stride, b)
    anc = O.anchor_grid(bases[2], 50, 84, 16)[:5000]
    dl = rng.normal(0, 0.5, (5000, 4)).astype(np.float32)
    got = host(dec(dev(anc), dev(dl)))
    ref = O.decode(anc, dl, 800, 1344, means=(0.0, 0.01, 0.0, -0.02), stds=(1.0, 0.5, 1.0, 2.0))
    assert np.array_equal(got, ref)
    # and against the float64 formula (tolerance stated by north_star: 1e-5 relative)
    a, dd = anc.astype(np.float64), dl.astype(np.float64) * np.array([1, 0.5, 1, 2.0]) + np.array([0, 0.01, 0, -0.02])
    pw, ph = a[:, 2] - a[:, 0] + 1, a[:, 3] - a[:, 1] + 1
    px, py = (a[:, 0] + a[:, 2]) / 2, (a[:, 1] + a[:, 3]) / 2
    mr = abs(np.log(0.016))
    gw, gh = pw * np.exp(np.clip(dd[:, 2], -mr, mr)), ph * np.exp(np.clip(dd[:, 3], -mr, mr))
    gx, gy = px + pw * dd[:, 0], py + ph * dd[:, 1]
    r64 = np.stack([np.clip(gx - gw / 2 + 0.5, 0, 1343), np.clip(gy - gh / 2 + 0.5, 0, 799),
                    np.clip(gx + gw / 2 - 0.5, 0, 1343), np.clip(gy + gh / 2 - 0.5, 0, 799)], 1)
    np.testing.assert_allclose(got, r64, rtol=1e-5, atol=1e-3)


# ---------------------------------------------------------------------------------------------- a3
@pytest.mark.parametrize("shape,k", [((3, 13, 21), 2000), ((3, 50, 84), 2000), ((3, 200, 336), 2000), ((1, 7, 9), 10)])
def test_topk_head_layout_bit_exact(shape, k):
    rng = np.random.default_rng(11)
    B = 3
    x = rng.normal(-4, 2, (B,) + shape).astype(np.float32)
    vals, idx = TopKPerLevel(k, apply_sigmoid=True)(dev(x))
    vals, idx = host(vals), host(idx)
    for b in range(B):
        flat = O.level_scores(x[b], True)
        rv, ri = O.topk(flat, k)
        assert np.array_equal(idx[b], ri), (shape, b)
        assert np.array_equal(vals[b], rv)


def test_topk_ties_and_special_values():
    rng = np.random.default_rng(12)
    B, N = 4, 70000
    x = np.zeros((B, N), np.float32)
    x[0] = 0.25                                   # everything ties: must return indices 0..K-1
    x[1] = rng.integers(0, 7, N).astype(np.float32)   # 7 distinct values, the K-th straddles a tie
    x[2] = rng.normal(0, 1, N).astype(np.float32)
    x[2, ::3] = -0.0
    x[2, 1::3] = 0.0
    x[3] = rng.normal(0, 1, N).astype(np.float32)
    x[3, 100] = np.inf
    x[3, 200] = -np.inf
    for k in (1, 37, 2000, 2048):
        vals, idx = TopKPerLevel(k)(dev(x))
        vals, idx = host(vals), host(idx)
        for b in range(B):
            rv, ri = O.topk(x[b], k)
            assert np.array_equal(idx[b], ri), (k, b)
            assert np.array_equal(vals[b].view(np.uint32), rv.view(np.uint32))
    assert list(host(TopKPerLevel(5)(dev(x[:1]))[1])[0]) == [0, 1, 2, 3, 4]


# ---------------------------------------------------------------------------------------------- a4
def _sorted_dets(rng, n, cluster, sigma=12.0):
    b = synth.rand_boxes(rng, n, cluster=cluster, sigma=sigma)
    s = np.sort(rng.uniform(0, 1, n).astype(np.float32))[::-1]
    return np.concatenate([b, s[:, None]], 1).astype(np.float32)


@pytest.mark.parametrize("n", [1, 2, 63, 64, 65, 819, 2000, 2048, 2049, 3000, 4096])      # > 2048: the 64-word sweep
def test_nms_default_mode_bit_exact(n):
    rng = np.random.default_rng(100 + n)
    dets = np.stack([_sorted_dets(rng, n, cluster=max(1, n // 25)) for _ in range(3)])
    dets[1, n // 2] = dets[1, 0]                 # exact duplicate
    keep_idx, mask, count = NMSWithMask(0.7)(dev(dets))
    keep_idx, mask, count = host(keep_idx), host(mask), host(count)
    for b in range(3):
        ref = O.nms(dets[b], 0.7)
        assert np.array_equal(mask[b].astype(np.uint8), ref), (n, b)
        kept = np.nonzero(ref)[0]
        assert count[b] == len(kept)
        assert np.array_equal(keep_idx[b, :len(kept)], kept.astype(np.int32))
        assert not keep_idx[b, len(kept):].any()


def test_nms_dense_crowd_and_modes():
    rng = np.random.default_rng(7)
    dets = _sorted_dets(rng, 2000, cluster=30, sigma=8.0)     # long suppression chains
    for thr, off, inc, eps in ((0.5, 0.0, False, 1e-8), (0.3, 0.0, True, 0.0), (0.7, 1.0, False, 0.0), (0.0, 0.0, True, 1e-8)):
        _, mask, _ = NMSWithMask(thr, off, inc, eps)(dev(dets))
        ref = O.nms(dets, np.float32(thr), off=off, inclusive=inc, union_eps=eps)
        assert np.array_equal(host(mask).astype(np.uint8), ref), (thr, off, inc, eps)
    # single 2-D (K,5) input like the reference's NmsNormalGpu
    _, mask, cnt = NMSWithMask(0.7)(dev(dets))
    assert mask.shape == (2000,) and int(cnt[0]) == int(host(mask).sum())


def test_nms_matches_reference_golden_vectors(golden):
    # the CUDA path against outputs of the reference's own nms_jit / apply_nms (tests/golden)
    for tag in "abc":
        for thr in (0.3, 0.7):
            dets = golden[f"nmsjit_{tag}_{thr}_dets"]
            order = np.argsort(-dets[:, 4], kind="stable")
            _, mask, _ = NMSWithMask(float(np.float32(thr)), 0.0, True, 0.0)(dev(dets[order]))
            assert np.array_equal(order[np.nonzero(host(mask))[0]], golden[f"nmsjit_{tag}_{thr}_keep"])
    for tag in "ab":
        for thr in (0.5, 0.7):
            boxes, scores = golden[f"applynms_{tag}_{thr}_boxes"], golden[f"applynms_{tag}_{thr}_scores"]
            order = np.argsort(-scores, kind="stable")
            _, mask, _ = NMSWithMask(float(np.float32(thr)), 1.0, False, 0.0)(dev(boxes[order]))
            assert np.array_equal(order[np.nonzero(host(mask))[0]], golden[f"applynms_{tag}_{thr}_keep"])


def test_nms_idempotent_full_size():
    rng = np.random.default_rng(8)
    dets = np.stack([_sorted_dets(rng, 2000, cluster=60) for _ in range(8)])
    keep_idx, mask, count = NMSWithMask(0.7)(dev(dets))
    m = host(mask)
    for b in range(8):
        kept = dets[b][m[b]]
        pad = np.zeros((2000 - len(kept), 5), np.float32)
        again = NMSWithMask(0.7)(dev(np.concatenate([kept, pad])[None]))
        assert host(again[1])[0, :len(kept)].all()      # NMS of the survivors keeps all of them


# ---------------------------------------------------------------------------------------- a3..a6
def _proposal_case(B, shapes, strides, nms_pre, max_num, seed):
    bases = synth.base_anchor_sets(strides)
    logits, deltas = synth.rpn_head_outputs(B, shapes, 3, seed)
    prop = Proposal((800, 1344), strides, bases, nms_pre=nms_pre, max_num=max_num)
    props, mask = prop([dev(x) for x in logits], [dev(x) for x in deltas])
    topk_idx, keep = prop.last_debug
    cfg = O.proposal_cfg(800, 1344, nms_pre=nms_pre, max_num=max_num)
    props, mask, topk_idx, keep = host(props), host(mask), host(topk_idx), host(keep)
    for b in range(B):
        ref = O.proposal_image([(logits[l][b], deltas[l][b], bases[l], strides[l]) for l in range(len(shapes))], cfg)
        off = 0
        for l, (h, w) in enumerate(shapes):
            K = min(nms_pre, 3 * h * w)
            assert np.array_equal(topk_idx[b, l, :K], ref["idx"][off:off + K]), ("topk", b, l)
            assert (topk_idx[b, l, K:] == -1).all()
            assert np.array_equal(keep[b, l, :K].astype(np.uint8), ref["keep"][off:off + K]), ("keep", b, l)
            assert not keep[b, l, K:].any()
            off += K
        assert np.array_equal(mask[b].astype(np.uint8), ref["mask"]), ("mask", b)
        assert np.array_equal(props[b], ref["props"]), ("props", b)
    return props, mask


def test_proposal_small_pyramid_bit_exact():
    _proposal_case(2, [(25, 42), (13, 21), (7, 11)], (16, 32, 64), nms_pre=500, max_num=600, seed=1)
    # max_num larger than everything available -> zero padding
    _proposal_case(1, [(7, 11), (4, 6)], (32, 64), nms_pre=100, max_num=400, seed=2)


def test_proposal_full_size_bit_exact():
    _proposal_case(2, synth.level_shapes(), synth.STRIDES, nms_pre=2000, max_num=2000, seed=3)


# ---------------------------------------------------------------------------------------- a7/a8
def _anchors_all():
    bases = synth.base_anchor_sets()
    return np.concatenate([O.anchor_grid(b, h, w, s) for b, (h, w), s in zip(bases, synth.level_shapes(), synth.STRIDES)])


@pytest.mark.parametrize("mode,full", [(0, False), (1, False), (0, True)])
def test_assign_sample_rpn_bit_exact(mode, full):
    """full=True forces the samplers' full-scan cluster select (the path taken when the short candidate list is
    not provably exact); both paths must give the oracle's sample."""
    B = 3
    anchors = _anchors_all()
    N = anchors.shape[0]
    gts, _, gvalid = synth.gt_boxes(B, G=128, seed=21)
    rng = np.random.default_rng(22)
    valid = (rng.uniform(0, 1, N) > 0.05).astype(np.uint8)
    op = BboxAssignSample(0.7, 0.3, 0.3, 128, 256, 256, seed=0x1234567890, mode=mode, force_full_scan=full)
    out = op(dev(gts), dev(gvalid).bool(), dev(anchors), dev(valid).bool())
    cfg = O.assign_cfg(0.7, 0.3, 0.3, 128, 256, 256, seed=0x1234567890, mode=mode)
    for b in range(B):
        ref = O.assign_sample_rpn(anchors, gts[b], gvalid[b], cfg, b, valid=valid)
        assert np.array_equal(host(out["assigned"][b]), ref["assigned"]), b
        for k in ("pos_idx", "neg_idx", "pos_gt"):
            assert np.array_equal(host(out[k][b]), ref[k]), (k, b)
        for k in ("pos_valid", "neg_valid"):
            assert np.array_equal(host(out[k][b]).astype(np.uint8), ref[k]), (k, b)
        assert int(out["num_pos"][b]) == ref["num_pos"]
        np.testing.assert_allclose(host(out["pos_target"][b]), ref["pos_target"], rtol=1e-5, atol=1e-6)


def test_assign_matches_reference_golden_vectors(golden):
    # CUDA mode 1 against outputs of the reference's own create_target_np (tests/golden)
    for tag in "abc":
        anchors, gts = golden[f"assign_{tag}_anchors"], golden[f"assign_{tag}_gts"]
        pos, neg = [float(v) for v in golden[f"assign_{tag}_thr"]]
        labels, gtids = golden[f"assign_{tag}_labels"], golden[f"assign_{tag}_gtids"]
        G = gts.shape[0]
        op = BboxAssignSample(pos, neg, 0.0, 16, 16, 32, mode=1)
        out = op(dev(gts[None]), torch.ones(1, G, dtype=torch.bool, device="cuda"), dev(anchors),
                 torch.ones(anchors.shape[0], dtype=torch.bool, device="cuda"))
        a = host(out["assigned"][0])
        assert np.array_equal(a > 0, labels > 0) and np.array_equal(a == 0, labels == 0) and np.array_equal(a == -1, labels == -1)
        fg = labels > 0
        assert np.array_equal(a[fg] - 1, gtids[fg])


def test_assign_edge_cases():
    anchors = _anchors_all()[:5000]
    gts = np.zeros((2, 4, 4), np.float32)
    gts[1, 0] = [10, 10, 60, 60]
    gvalid = np.array([[0, 0, 0, 0], [1, 0, 0, 0]], np.uint8)    # image 0 has no gt at all
    op = BboxAssignSample(0.7, 0.3, 0.3, 8, 16, 16, seed=5)
    out = op(dev(gts), dev(gvalid).bool(), dev(anchors), torch.ones(5000, dtype=torch.bool, device="cuda"))
    cfg = O.assign_cfg(0.7, 0.3, 0.3, 8, 16, 16, seed=5)
    for b in range(2):
        ref = O.assign_sample_rpn(anchors, gts[b], gvalid[b], cfg, b)
        assert np.array_equal(host(out["assigned"][b]), ref["assigned"])
        assert np.array_equal(host(out["pos_idx"][b]), ref["pos_idx"]) and np.array_equal(host(out["neg_idx"][b]), ref["neg_idx"])
        assert np.array_equal(host(out["pos_valid"][b]).astype(np.uint8), ref["pos_valid"])
    assert int(out["num_pos"][0]) == 0


@pytest.mark.parametrize("full", [False, True])
def test_assign_sample_rcnn_bit_exact(full):
    B, P = 3, 2000
    rng = np.random.default_rng(31)
    gts, labels, gvalid = synth.gt_boxes(B, G=128, seed=32)
    props = np.zeros((B, P, 5), np.float32)
    pmask = np.zeros((B, P), np.uint8)
    for b in range(B):
        nv = int(gvalid[b].sum())
        jit = gts[b, rng.integers(0, nv, 600)] + rng.normal(0, 6, (600, 4)).astype(np.float32)
        props[b, :600, :4] = jit
        props[b, 600:, :4] = synth.rand_boxes(rng, P - 600)
        props[b, :, 4] = np.sort(rng.uniform(0, 1, P))[::-1]
        pmask[b, :1700] = 1
    op = BboxAssignSampleForRcnn(0.5, 0.5, 0.5, 128, 384, 512, seed=99, force_full_scan=full)
    out = op(dev(gts), dev(labels), dev(pmask).bool(), dev(props), dev(gvalid).bool())
    cfg = O.assign_cfg(0.5, 0.5, 0.5, 128, 384, 512, stds=(0.1, 0.1, 0.2, 0.2), seed=99)
    for b in range(B):
        ref = O.assign_sample_rcnn(props[b, :, :4], pmask[b], gts[b], labels[b], gvalid[b], cfg, b)
        assert np.array_equal(host(out["assigned"][b]), ref["assigned"]), b
        assert np.array_equal(host(out["sel_idx"][b]), ref["sel_idx"]), b
        assert np.array_equal(host(out["labels"][b]), ref["labels"])
        assert np.array_equal(host(out["mask"][b]).astype(np.uint8), ref["mask"])
        assert np.array_equal(host(out["rois"][b])[:, 1:], ref["rois"])
        assert (host(out["rois"][b])[:, 0] == b).all()
        np.testing.assert_allclose(host(out["deltas"][b]), ref["deltas"], rtol=1e-5, atol=1e-5)
        assert int(out["num_pos"][b]) == ref["num_pos"]
    # upstream 640-slot variant (128 + 512, total 512)
    op2 = BboxAssignSampleForRcnn(0.5, 0.5, 0.5, 128, 512, 512, seed=99)
    out2 = op2(dev(gts), dev(labels), dev(pmask).bool(), dev(props), dev(gvalid).bool())
    cfg2 = O.assign_cfg(0.5, 0.5, 0.5, 128, 512, 512, stds=(0.1, 0.1, 0.2, 0.2), seed=99)
    ref2 = O.assign_sample_rcnn(props[0, :, :4], pmask[0], gts[0], labels[0], gvalid[0], cfg2, 0)
    assert np.array_equal(host(out2["sel_idx"][0]), ref2["sel_idx"]) and int(host(out2["mask"][0]).sum()) <= 512


# ---------------------------------------------------------------------------------------- a9..a11
def _rois(rng, R, B):
    b = synth.rand_boxes(rng, R, smin=4, smax=900)
    return np.concatenate([rng.integers(0, B, (R, 1)).astype(np.float32), b], 1).astype(np.float32)


def test_roi_levels_bit_exact():
    rng = np.random.default_rng(41)
    rois = _rois(rng, 50000, 4)
    # park many RoIs exactly on the level boundaries (sqrt(area) = 112, 224, 448)
    for i, s in enumerate((112, 224, 448)):
        rois[i * 100:(i + 1) * 100, 1:] = [10, 10, 10 + s - 1, 10 + s - 1]
        rois[300 + i * 100:400 + i * 100, 3] = rois[300 + i * 100:400 + i * 100, 1] + s - 1 + rng.integers(-1, 2, 100)
    ext = SingleRoIExtractor()
    assert np.array_equal(host(ext.map_roi_levels(dev(rois))), O.roi_levels(rois, 56.0, 4))


@pytest.mark.parametrize("exact", [False, True])
@pytest.mark.parametrize("P,S,C", [(7, 2, 16), (14, 2, 8), (7, 1, 4), (7, 2, 37), (7, 2, 64)])
def test_roialign_fwd_bwd_vs_oracle(P, S, C, exact):
    rng = np.random.default_rng(42)
    B = 2
    shapes = synth.level_shapes()[:4]
    strides = synth.STRIDES[:4]
    feats = [rng.uniform(-1, 1, (B, C, h, w)).astype(np.float32) for h, w in shapes]
    rois = _rois(rng, 160, B)
    rois[0, 1:] = [-30, -30, 40, 50]
    rois[1, 1:] = [1300, 760, 1343, 799]
    rois[2, 1:] = [20, 20, 20.4, 20.2]
    rois[3, 1:] = [0, 0, 1343, 799]
    rois[4, 1:] = [100, 5, 130, 790]      # tall and thin -> large footprint on a fine level (split into bin-row groups)
    rois[5, 1:] = [5, 100, 1300, 130]     # wide and flat -> wider than the largest TMA box (gather path)
    rois[6, 1:] = [-500, -500, -300, -300]  # entirely outside: every sample invalid -> zeros
    rois[7, 1:] = [1342, 798, 1400, 900]
    rois[8, 1:] = [300, 300, 301, 301]    # one-pixel RoI: every bin on the same two rows / columns (dense row weights)
    rois[9, 1:] = [40, 40, 67, 67]        # bin size exactly 1 feature pixel on level 0: samples on x.25 / x.75
    rois[10, 1:] = [64, 64, 175, 175]     # sqrt(area) = 112: first RoI of level 1
    rois[11, 1:] = [10, 10, 450, 120]     # 110 columns on level 1 (3 column chunks of the channel-lane kernel)
    rois[12, 0] = -1                       # batch index out of range: no data (zeros forward, nothing backward)
    rois[13, 0] = B
    ext = SingleRoIExtractor(P, S, strides, 56, exact=exact)
    ft = [dev(f).requires_grad_(True) for f in feats]
    out = ext(dev(rois), *ft)
    ref = O.roialign_fwd(feats, strides, rois, P=P, S=S)
    if exact:
        assert np.array_equal(host(out), ref)          # gather kernels: oracle's op order, bit-identical
    else:
        # TMA separable path: different summation order + FMA.  north_star: 1e-5 relative; atol 1e-6 covers
        # cancellation in near-zero averages of U(-1,1) features.
        np.testing.assert_allclose(host(out), ref, rtol=1e-5, atol=1e-6)
    dout = rng.uniform(-1, 1, ref.shape).astype(np.float32)
    out.backward(dev(dout))
    dref = O.roialign_bwd([f.shape for f in feats], strides, rois, dout, P=P, S=S)
    for l in range(4):
        # float atomics / L2 reductions reorder the sums: tolerance 1e-5 relative to the gradient scale
        np.testing.assert_allclose(host(ft[l].grad), dref[l], rtol=1e-5, atol=1e-5 * max(1.0, np.abs(dref[l]).max()))


@pytest.mark.parametrize("exact", [False, True])
def test_roialign_taps_vs_reference_bilinear_golden(golden, exact):
    """The CUDA path against outputs of the REFERENCE's own bilinear_interpolate_torch
    (centerpoint/det3d_ms/core/utils/center_utils.py:97-131; fixture made by tests/golden/make_golden.py):
    a 1x1-bin, 1-sample RoIAlign of a stride-1 map is one bilinear read at the RoI centre."""
    im, x, y, ref = golden["bilinear_im"], golden["bilinear_x"], golden["bilinear_y"], golden["bilinear_val"]
    feat = np.ascontiguousarray(im.transpose(2, 0, 1)[None])
    h = np.float32(0.5)
    rois = np.stack([np.zeros_like(x), x - h, y - h, x + h, y + h], 1).astype(np.float32)
    ext = SingleRoIExtractor(1, 1, (1,), 56, exact=exact)
    got = host(ext(dev(rois), dev(feat))).reshape(len(x), -1)
    np.testing.assert_allclose(got, ref, rtol=0, atol=2e-5)
    assert np.array_equal(got, O.roialign_fwd([feat], (1,), rois, P=1, S=1).reshape(len(x), -1)) or not exact


def test_roialign_full_size_vs_oracle():
    """config-2 shapes (C=256, 4 levels), 128 RoIs against the oracle."""
    rng = np.random.default_rng(44)
    B, C = 2, 256
    shapes = synth.level_shapes()[:4]
    strides = synth.STRIDES[:4]
    feats = [rng.uniform(-1, 1, (B, C, h, w)).astype(np.float32) for h, w in shapes]
    rois = _rois(rng, 128, B)
    ext = SingleRoIExtractor()
    ft = [dev(f).requires_grad_(True) for f in feats]
    out = ext(dev(rois), *ft)
    ref = O.roialign_fwd(feats, strides, rois)
    np.testing.assert_allclose(host(out), ref, rtol=1e-5, atol=1e-6)
    dout = rng.uniform(-1, 1, ref.shape).astype(np.float32)
    out.backward(dev(dout))
    dref = O.roialign_bwd([f.shape for f in feats], strides, rois, dout)
    for l in range(4):
        np.testing.assert_allclose(host(ft[l].grad), dref[l], rtol=1e-5, atol=1e-5 * max(1.0, np.abs(dref[l]).max()))


@pytest.mark.parametrize("C", [32, 256])
def test_roialign_channel_lane_kernels_vs_oracle(C, monkeypatch):
    """The opt-in channel-per-lane kernels (roialign_ch.cu, MD_ROI_CH=1; the launcher reads the switch on every call):
    compact, tall, wide (in-kernel gather), edge-clamped, outside and bad-batch RoIs on all four levels, several
    (RoI, channel group) items per persistent warp."""
    monkeypatch.setenv("MD_ROI_CH", "1")
    rng = np.random.default_rng(46)
    B = 2
    shapes = synth.level_shapes()[:4]
    strides = synth.STRIDES[:4]
    feats = [rng.uniform(-1, 1, (B, C, h, w)).astype(np.float32) for h, w in shapes]
    n = 3000 if C == 32 else 700              # > 4 items per warp of the persistent grid
    rois = _rois(rng, n, B)
    rois[0, 1:] = [-30, -30, 40, 50]
    rois[1, 1:] = [1300, 760, 1343, 799]
    rois[2, 1:] = [20, 20, 20.4, 20.2]
    rois[3, 1:] = [0, 0, 1343, 799]
    rois[4, 1:] = [100, 5, 130, 790]
    rois[5, 1:] = [5, 100, 1300, 130]
    rois[6, 1:] = [-500, -500, -300, -300]
    rois[7, 1:] = [300, 300, 301, 301]
    rois[8, 1:] = [10, 10, 450, 120]
    rois[9, 1:] = [200, 100, 255, 160]
    rois[10, 0] = -1
    rois[11, 0] = B
    ext = SingleRoIExtractor()
    ft = [dev(f).requires_grad_(True) for f in feats]
    out = ext(dev(rois), *ft)
    ref = O.roialign_fwd(feats, strides, rois)
    np.testing.assert_allclose(host(out), ref, rtol=1e-5, atol=1e-6)
    dout = rng.uniform(-1, 1, ref.shape).astype(np.float32)
    out.backward(dev(dout))
    dref = O.roialign_bwd([f.shape for f in feats], strides, rois, dout)
    for l in range(4):
        np.testing.assert_allclose(host(ft[l].grad), dref[l], rtol=1e-5, atol=1e-5 * max(1.0, np.abs(dref[l]).max()))

@pytest.mark.parametrize("C,S,acc", [(32, 2, False), (256, 2, False), (64, 2, True), (32, 1, False)])
def test_roialign_tile_backward_vs_oracle(C, S, acc, monkeypatch):
    """The tile-stationary backward (roialign_tile.cu: every dX byte written once, no zero-fill): compact, tall, wide, edge-clamped,
    outside and bad-batch RoIs on all four levels, many RoIs per tile; the output tensors start as garbage (nothing may rely on
    a zero-fill).  acc: the accumulating form (MD_ROI_TILE_ACC=1) through MdRoiAlignBwdAcc.  S = 1: the plan declines every
    RoI, so the tile kernel writes zeros and the gather kernel adds everything."""
    monkeypatch.setenv("MD_ROI_TILE", "1")
    if acc:
        monkeypatch.setenv("MD_ROI_TILE_ACC", "1")
    rng = np.random.default_rng(47)
    B = 2
    shapes = synth.level_shapes()[:4]
    strides = synth.STRIDES[:4]
    n = 3000 if C <= 64 else 600
    rois = _rois(rng, n, B)
    rois[0, 1:] = [-30, -30, 40, 50]
    rois[1, 1:] = [1300, 760, 1343, 799]
    rois[2, 1:] = [20, 20, 20.4, 20.2]
    rois[3, 1:] = [0, 0, 1343, 799]
    rois[4, 1:] = [100, 5, 130, 790]
    rois[5, 1:] = [5, 100, 1300, 130]
    rois[6, 1:] = [-500, -500, -300, -300]
    rois[7, 1:] = [300, 300, 301, 301]
    rois[8, 1:] = [10, 10, 450, 120]
    rois[9, 1:] = [40, 40, 67, 67]
    rois[10, 0] = -1
    rois[11, 0] = B
    rois[12, 1:] = [1342, 798, 1400, 900]
    rois[13:40, 1:] = rois[13:40, 1:] * 0.05 + np.array([600, 400, 600, 400], np.float32)   # a crowd on a few tiles
    ext = SingleRoIExtractor(7, S, strides, 56)
    dout = rng.uniform(-1, 1, (n, C, 7, 7)).astype(np.float32)
    dref = O.roialign_bwd([(B, C, h, w) for h, w in shapes], strides, rois, dout, S=S)
    if acc:
        base = [rng.uniform(-1, 1, (B, C, h, w)).astype(np.float32) for h, w in shapes]
        got = ext._backward_into(dev(rois), dev(dout), [dev(b) for b in base])
    else:
        base = [np.zeros((B, C, h, w), np.float32) for h, w in shapes]
        junk = [torch.full((B, C, h, w), float("nan"), device="cuda") for h, w in shapes]   # recycled by the caching allocator
        del junk
        got = ext._backward(dev(rois), dev(dout), [(B, C, h, w) for h, w in shapes])
    for l in range(4):
        tol = 1e-5 * max(1.0, np.abs(dref[l]).max())
        np.testing.assert_allclose(host(got[l]), base[l] + dref[l], rtol=1e-5, atol=2 * tol if acc else tol)


@pytest.mark.parametrize("chunk", ["32767", "4"])
def test_roialign_tile_backward_repeats(chunk, monkeypatch):
    """Visits go in RoI order, so with one work item per tile (MD_TILE_CHUNK large) two runs give bit-identical gradients.
    A crowded tile is cut into several items that add their partial sums at L2 (default 12 visits per item; 4 here so that
    many tiles are cut): those sums may differ in the last bits from run to run and must stay inside the tolerance."""
    monkeypatch.setenv("MD_TILE_CHUNK", chunk)
    rng = np.random.default_rng(48)
    B, C = 2, 64
    shapes = synth.level_shapes()[:4]
    strides = synth.STRIDES[:4]
    rois_h = _rois(rng, 2000, B)
    rois_h[:300, 1:] = rois_h[:300, 1:] * 0.03 + np.array([500, 300, 500, 300], np.float32)     # a crowd: ~300 RoIs on a few tiles
    rois = dev(rois_h)
    dout_h = rng.uniform(-1, 1, (2000, C, 7, 7)).astype(np.float32)
    dout = dev(dout_h)
    ext = SingleRoIExtractor()
    a = ext._backward(rois, dout, [(B, C, h, w) for h, w in shapes])
    b = ext._backward(rois, dout, [(B, C, h, w) for h, w in shapes])
    dref = O.roialign_bwd([(B, C, h, w) for h, w in shapes], strides, rois_h, dout_h)
    for l, (x, y) in enumerate(zip(a, b)):
        tol = 1e-5 * max(1.0, np.abs(dref[l]).max())
        np.testing.assert_allclose(host(x), dref[l], rtol=1e-5, atol=tol)
        if chunk == "32767":
            assert torch.equal(x, y)
        else:
            np.testing.assert_allclose(host(y), host(x), rtol=1e-5, atol=tol)


def test_roialign_two_op_backward_vs_oracle():
    """MdRoiAlignBwdPrepare (on another stream, beside other work) + MdRoiAlignBwdPlanned == the oracle; the plan tensor is
    owned by the caller and can be consumed twice (the backward re-arms its ticket)."""
    rng = np.random.default_rng(49)
    B, C = 2, 64
    shapes = synth.level_shapes()[:4]
    strides = synth.STRIDES[:4]
    rois_h = _rois(rng, 1500, B)
    rois_h[0, 1:] = [5, 100, 1300, 130]
    rois_h[1, 0] = -1
    rois_h[2:200, 1:] = rois_h[2:200, 1:] * 0.03 + np.array([500, 300, 500, 300], np.float32)     # a crowd on a few tiles
    dout_h = rng.uniform(-1, 1, (1500, C, 7, 7)).astype(np.float32)
    dref = O.roialign_bwd([(B, C, h, w) for h, w in shapes], strides, rois_h, dout_h)
    ext = SingleRoIExtractor()
    rois, dout = dev(rois_h), dev(dout_h)
    feats = [torch.empty(B, C, h, w, device="cuda") for h, w in shapes]
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        plan = ext.prepare_backward(rois, feats)
    torch.cuda.current_stream().wait_stream(side)
    fshapes = [(B, C, h, w) for h, w in shapes]
    for _ in range(2):
        got = ext._backward_planned(rois, dout, fshapes, plan)
        for l in range(4):
            np.testing.assert_allclose(host(got[l]), dref[l], rtol=1e-5, atol=1e-5 * max(1.0, np.abs(dref[l]).max()))
    # a plan that was prepared for another RoI count is refused on the device (signature mismatch): nothing is read through it
    ext._backward_planned(rois[:1400].contiguous(), dout[:1400].contiguous(), fshapes, plan)
    torch.cuda.synchronize()
    # shapes the tile kernel does not take are refused (return code 4), never silently mis-handled
    ext14 = SingleRoIExtractor(14, 2, strides, 56)
    with pytest.raises(Exception):
        ext14._backward_planned(rois, dev(rng.uniform(-1, 1, (1500, C, 14, 14)).astype(np.float32)), fshapes, plan)


@pytest.mark.parametrize("shapes,strides", [([(7, 45), (3, 9)], (4, 16)), ([(5, 30)], (8,)), ([(64, 64), (33, 31), (17, 16), (8, 8), (4, 4)], (4, 8, 16, 32, 64))])
def test_roialign_tile_backward_odd_pyramids(shapes, strides):
    """Tile kernel on pyramids that are nothing like config 2: maps smaller than one tile, widths that are not multiples of 4
    (scalar read-out), one level, five levels, one image, RoIs hanging over every border."""
    rng = np.random.default_rng(50)
    B, C, n = 1, 32, 300
    img_w, img_h = shapes[0][1] * strides[0], shapes[0][0] * strides[0]
    b = synth.rand_boxes(rng, n, img_w=float(img_w), img_h=float(img_h), smin=2, smax=float(max(img_w, img_h)))
    b += rng.uniform(-20, 20, b.shape).astype(np.float32)                      # some of them outside the image
    rois = np.concatenate([np.zeros((n, 1), np.float32), b], 1).astype(np.float32)
    ext = SingleRoIExtractor(7, 2, strides, 56)
    dout = rng.uniform(-1, 1, (n, C, 7, 7)).astype(np.float32)
    fshapes = [(B, C, h, w) for h, w in shapes]
    dref = O.roialign_bwd(fshapes, strides, rois, dout)
    got = ext._backward(dev(rois), dev(dout), fshapes)
    plan = ext.prepare_backward(dev(rois), [torch.empty(s, device="cuda") for s in fshapes])
    got2 = ext._backward_planned(dev(rois), dev(dout), fshapes, plan)
    for l in range(len(shapes)):
        tol = 1e-5 * max(1.0, np.abs(dref[l]).max())
        np.testing.assert_allclose(host(got[l]), dref[l], rtol=1e-5, atol=tol)
        np.testing.assert_allclose(host(got2[l]), dref[l], rtol=1e-5, atol=tol)


def test_roialign_bwd_accumulates_into_caller_tensors():
    """MdRoiAlignBwdAcc: acc_l += ROIAlignGrad(dout).  Starting from zeros it is the plain bprop (oracle), starting
    from an existing gradient it adds to it; the self-contained MdRoiAlignBwd stays the zero-filling form."""
    rng = np.random.default_rng(45)
    B, C = 2, 32
    shapes = synth.level_shapes()[:4]
    strides = synth.STRIDES[:4]
    rois = _rois(rng, 96, B)
    rois[0, 1:] = [5, 100, 1300, 130]       # declined by the TMA path -> gather kernel, also accumulating
    ext = SingleRoIExtractor()
    dout = rng.uniform(-1, 1, (96, C, 7, 7)).astype(np.float32)
    dref = O.roialign_bwd([(B, C, h, w) for h, w in shapes], strides, rois, dout)
    zeros = [torch.zeros(B, C, h, w, device="cuda") for h, w in shapes]
    got = ext._backward_into(dev(rois), dev(dout), zeros)
    base = [rng.uniform(-1, 1, (B, C, h, w)).astype(np.float32) for h, w in shapes]
    acc = ext._backward_into(dev(rois), dev(dout), [dev(b) for b in base])
    for l in range(4):
        tol = 1e-5 * max(1.0, np.abs(dref[l]).max())
        np.testing.assert_allclose(host(got[l]), dref[l], rtol=1e-5, atol=tol)
        np.testing.assert_allclose(host(acc[l]), base[l] + dref[l], rtol=1e-5, atol=2 * tol)


def test_roialign_full_size_properties():
    """config-2 sizes: linearity f(a*x+y) = a*f(x)+f(y) (tolerance) and adjointness <f(x),d> = <x,f^T(d)>."""
    rng = np.random.default_rng(43)
    B, C = 2, 256
    shapes = synth.level_shapes()[:4]
    rois = dev(_rois(rng, 512, B))
    ext = SingleRoIExtractor()
    x = [torch.rand(B, C, h, w, device="cuda") * 2 - 1 for h, w in shapes]
    y = [torch.rand(B, C, h, w, device="cuda") * 2 - 1 for h, w in shapes]
    fx, fy = ext(rois, *x), ext(rois, *y)
    fxy = ext(rois, *[2.0 * a + b for a, b in zip(x, y)])
    torch.testing.assert_close(fxy, 2.0 * fx + fy, rtol=1e-4, atol=1e-5)
    d = torch.rand_like(fx)
    xs = [a.clone().requires_grad_(True) for a in x]
    ext(rois, *xs).backward(d)
    lhs = (fx.double() * d.double()).sum().item()
    rhs = sum((a.double() * g.grad.double()).sum().item() for a, g in zip(x, xs))
    # fp32 rounding in both operators: each of the ~6.4M products is off by ~1e-7 relative, so the two sums differ by
    # ~1e-7 * sqrt(n) * |term| ~ 5e-5 (seen: 1e-5 .. 6e-5, the reduce-add order is not deterministic) whatever |lhs|
    # happens to be after cancellation; one dropped RoI would move them apart by ~30.  Scale: the sum of |terms|.
    scale = (fx.double() * d.double()).abs().sum().item()
    assert abs(lhs - rhs) <= 2e-8 * scale
