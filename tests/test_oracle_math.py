"""Accuracy pins for the oracle's deterministic fp32 math and independent cross-checks against
torchvision / torch (NOT the reference; conventions coincide where stated).  CPU only."""
import numpy as np
import pytest
import torch

import oracle as O


def _ulp_err(got, ref64):
    ref32 = ref64.astype(np.float32)
    ulp = np.spacing(np.abs(ref32)).astype(np.float64)
    return np.abs(got.astype(np.float64) - ref64) / ulp


def test_o_exp_within_2ulp():
    x = np.concatenate([np.linspace(-87, 88, 400001), np.linspace(-4.2, 4.2, 200001)]).astype(np.float32)
    got = O.exp(x)
    assert _ulp_err(got, np.exp(x.astype(np.float64))).max() <= 2.0


def test_o_sigmoid_accuracy_and_monotone():
    x = np.linspace(-30, 30, 600001).astype(np.float32)
    got = O.sigmoid(x)
    ref = 1.0 / (1.0 + np.exp(-x.astype(np.float64)))
    assert _ulp_err(got, ref).max() <= 4.0
    assert np.all(np.diff(got) >= 0)


def test_philox_known_answer():
    # Random123 KAT for philox4x32-10: ctr=0, key=0 -> 6627e8d5 e169c58d bc57ac4c 9b00dbd8
    assert int(O.philox_key([0], 0, 0, 0)[0]) == 0x6627E8D5
    # ctr = (243f6a88, 85a308d3, 13198a2e, 03707344) cannot be expressed (4th word fixed to 0); check
    # determinism + dispersion instead.
    k = O.philox_key(np.arange(4096), 1, 2, 3)
    assert len(np.unique(k)) > 4090


def test_topk_vs_torch_on_unique_scores():
    rng = np.random.default_rng(0)
    s = rng.permutation(50000).astype(np.float32) / 7.0 - 1000.0
    v, i = O.topk(s, 2000)
    tv, ti = torch.topk(torch.from_numpy(s), 2000, sorted=True)
    assert np.array_equal(i, ti.numpy().astype(np.int32))
    assert np.array_equal(v, tv.numpy())


def test_topk_ties_lower_index_first():
    s = np.zeros(1000, np.float32)
    s[[5, 900, 17]] = 1.0
    s[100:200] = 0.5
    v, i = O.topk(s, 60)
    assert list(i[:3]) == [5, 17, 900]
    assert list(i[3:]) == list(range(100, 157))
    # -0.0 sorts below +0.0
    s2 = np.array([-0.0, 0.0, -0.0, 0.0], np.float32)
    assert list(O.topk(s2, 4)[1]) == [1, 3, 0, 2]


def test_nms_default_matches_torchvision():
    tv = pytest.importorskip("torchvision")
    rng = np.random.default_rng(1)
    for n, cl in ((500, None), (2000, 40)):
        ctr = rng.uniform([0, 0], [1344, 800], (cl or n, 2))
        pick = rng.integers(0, cl or n, n)
        c = ctr[pick] + rng.normal(0, 10, (n, 2))
        wh = np.exp(rng.uniform(np.log(8), np.log(300), (n, 2)))
        b = np.concatenate([c - wh / 2, c + wh / 2], 1).astype(np.float32)
        sc = np.sort(rng.permutation(n).astype(np.float32))[::-1].copy()
        keep_tv = tv.ops.nms(torch.from_numpy(b), torch.from_numpy(sc), 0.7).numpy()
        mask = O.nms(b, 0.7)
        assert np.array_equal(np.nonzero(mask)[0], np.sort(keep_tv))


def _feats(rng, B, C, shapes):
    return [rng.uniform(-1, 1, (B, C, h, w)).astype(np.float32) for h, w in shapes]


def test_roialign_matches_torchvision_fwd_bwd():
    tv = pytest.importorskip("torchvision")
    rng = np.random.default_rng(2)
    B, C = 2, 8
    shapes = [(50, 84), (25, 42), (13, 21), (7, 11)]
    strides = [4, 8, 16, 32]
    feats = _feats(rng, B, C, shapes)
    R = 64
    c = rng.uniform([0, 0], [336, 200], (R, 2))
    wh = np.exp(rng.uniform(np.log(4), np.log(300), (R, 2)))
    rois = np.concatenate([rng.integers(0, B, (R, 1)), c - wh / 2, c + wh / 2], 1).astype(np.float32)
    rois[0, 1:] = [-20, -20, 30, 30]      # sticks out of the image
    rois[1, 1:] = [300, 150, 400, 260]
    rois[2, 1:] = [10, 10, 10.2, 10.3]    # tiny -> roi_w clamps to 1
    lvl = O.roi_levels(rois, 56.0, 4)
    got = O.roialign_fwd(feats, strides, rois, P=7, S=2, lvl=lvl)
    dout = rng.uniform(-1, 1, got.shape).astype(np.float32)
    dgot = O.roialign_bwd([f.shape for f in feats], strides, rois, dout, P=7, S=2, lvl=lvl)
    for l in range(4):
        sel = np.nonzero(lvl == l)[0]
        if len(sel) == 0:
            continue
        ft = torch.from_numpy(feats[l]).requires_grad_(True)
        ref = tv.ops.roi_align(ft, torch.from_numpy(rois[sel]), (7, 7), 1.0 / strides[l], 2, aligned=False)
        np.testing.assert_allclose(got[sel], ref.detach().numpy(), rtol=1e-5, atol=1e-6)
        ref.backward(torch.from_numpy(dout[sel]))
        np.testing.assert_allclose(dgot[l], ft.grad.numpy(), rtol=1e-4, atol=1e-5)


def test_roi_levels_equals_floor_log2():
    rng = np.random.default_rng(3)
    wh = np.exp(rng.uniform(np.log(2), np.log(1500), (20000, 2)))
    rois = np.concatenate([np.zeros((20000, 3)), wh - 1], 1).astype(np.float32)
    rois[:, 0] = 0
    got = O.roi_levels(rois, 56.0, 4)
    s = np.sqrt((rois[:, 3].astype(np.float64) + 1) * (rois[:, 4].astype(np.float64) + 1))
    ref = np.clip(np.floor(np.log2(s / 56 + 1e-6)), 0, 3).astype(np.int32)
    # identical away from the power-of-two boundaries; at them only fp32 rounding may differ
    away = np.abs(np.log2(s / 56 + 1e-6) - np.round(np.log2(s / 56 + 1e-6))) > 1e-5
    assert np.array_equal(got[away], ref[away])
    assert (got != ref).mean() < 1e-3


def test_decode_matches_float64_formula():
    rng = np.random.default_rng(4)
    n = 12600
    base = np.array([[-22, -10, 25, 13], [-14, -14, 17, 17], [-10, -22, 13, 25]], np.float32)
    anc = O.anchor_grid(base, 50, 84, 16.0)[:n]
    d = np.concatenate([rng.normal(0, 0.3, (n, 2)), rng.normal(0, 1.5, (n, 2))], 1).astype(np.float32)
    got = O.decode(anc, d, 800, 1344)
    a, dd = anc.astype(np.float64), d.astype(np.float64)
    pw, ph = a[:, 2] - a[:, 0] + 1, a[:, 3] - a[:, 1] + 1
    px, py = (a[:, 0] + a[:, 2]) / 2, (a[:, 1] + a[:, 3]) / 2
    mr = abs(np.log(0.016))
    gw, gh = pw * np.exp(np.clip(dd[:, 2], -mr, mr)), ph * np.exp(np.clip(dd[:, 3], -mr, mr))
    gx, gy = px + pw * dd[:, 0], py + ph * dd[:, 1]
    ref = np.stack([np.clip(gx - gw / 2 + 0.5, 0, 1343), np.clip(gy - gh / 2 + 0.5, 0, 799),
                    np.clip(gx + gw / 2 - 0.5, 0, 1343), np.clip(gy + gh / 2 - 0.5, 0, 799)], 1)
    np.testing.assert_allclose(got, ref, rtol=1e-5, atol=1e-3)


def test_assign_mode0_rules_small_case():
    gts = np.array([[0, 0, 9, 9], [20, 20, 39, 39], [100, 100, 109, 109]], np.float32)
    boxes = np.array([
        [0, 0, 9, 9],        # IoU 1 with gt0 -> pos (1)
        [0, 0, 9, 4],        # IoU .5 with gt0 -> between thresholds -> ignore (-1)
        [22, 22, 41, 41],    # IoU .68 with gt1: best for gt1 and >= min_pos -> force (2)
        [200, 200, 210, 210],  # no overlap -> neg (0)
        [100, 100, 101, 101],  # IoU .04 with gt2: best for gt2 but < min_pos_iou -> neg (0)
        [0, 0, 9, 9],        # invalid anchor -> -1
    ], np.float32)
    valid = np.array([1, 1, 1, 1, 1, 0], np.uint8)
    a, m, am = O.assign(boxes, gts, 0.7, 0.3, 0.3, valid=valid, mode=0)
    assert list(a) == [1, -1, 2, 0, 0, -1]
    # invalid gts are skipped
    a2, _, _ = O.assign(boxes, gts, 0.7, 0.3, 0.3, valid=valid, gt_valid=np.array([0, 1, 1], np.uint8), mode=0)
    assert list(a2) == [0, 0, 2, 0, 0, -1]


def test_sample_is_k_smallest_philox_keys():
    rng = np.random.default_rng(5)
    assigned = rng.integers(-1, 3, 5000).astype(np.int32)
    idx, cnt = O.sample(assigned, True, 7, 3, 99, 128)
    cands = np.nonzero(assigned > 0)[0]
    assert cnt == len(cands)
    keys = O.philox_key(cands, 7, 3, 99).astype(np.uint64) << np.uint64(32) | cands.astype(np.uint64)
    ref = cands[np.argsort(keys)[:128]]
    assert np.array_equal(idx, ref.astype(np.int32))
    idx2, cnt2 = O.sample(np.array([0, 1, 0, 0], np.int32), True, 0, 0, 0, 8)
    assert cnt2 == 1 and list(idx2) == [1, 0, 0, 0, 0, 0, 0, 0]
