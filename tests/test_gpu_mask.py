"""Mask R-CNN additions (SURVEY.md 8(f) row 2 / BASELINE config 4): 28x28 mask-target crop bit-exact vs the oracle;
14x14 mask RoIAlign forward/backward is covered by test_gpu_parity.py::test_roialign_fwd_bwd_vs_oracle[P=14]."""
import numpy as np
import pytest
import torch

import oracle as O
from minddet_b200 import MaskTargets, synth

pytestmark = pytest.mark.gpu


def dev(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


def blob_masks(rng, B, G, H, W, gts):
    m = np.zeros((B, G, H, W), np.uint8)
    yy, xx = np.mgrid[0:H, 0:W]
    for b in range(B):
        for g in range(G):
            x1, y1, x2, y2 = gts[b, g]
            if x2 <= x1 or y2 <= y1:
                continue
            cx, cy, rx, ry = (x1 + x2) / 2, (y1 + y2) / 2, max((x2 - x1) / 2, 1), max((y2 - y1) / 2, 1)
            m[b, g] = (((xx - cx) / rx) ** 2 + ((yy - cy) / ry) ** 2 <= 1.0).astype(np.uint8)
    return m


@pytest.mark.parametrize("M,S", [(28, 2), (14, 1), (28, 3)])
def test_mask_targets_bit_exact(M, S):
    rng = np.random.default_rng(600 + M + S)
    B, G, H, W = 2, 6, 200, 336
    gts, _, gvalid = synth.gt_boxes(B, G=G, max_valid=G, img_h=H, img_w=W, seed=61)
    masks = blob_masks(rng, B, G, H, W, gts)
    R = 96
    rois = np.zeros((R, 5), np.float32)
    gt_idx = np.zeros(R, np.int32)
    for r in range(R):
        b = r % B
        g = int(rng.integers(0, max(1, int(gvalid[b].sum()))))
        rois[r, 0] = b
        rois[r, 1:] = gts[b, g] + rng.normal(0, 6, 4)
        gt_idx[r] = g
    rois[0, 1:] = [-20, -20, 30, 40]            # partly outside the image
    rois[1, 1:] = [330, 190, 400, 260]
    rois[2, 1:] = [50, 50, 50.3, 50.2]          # degenerate: 1-pixel minimum size rule
    gt_idx[3] = -1                              # no gt -> empty target
    gt_idx[4] = G + 3                           # out of range -> empty target
    got = MaskTargets(M, S)(dev(masks).bool(), dev(rois), dev(gt_idx)).cpu().numpy().astype(np.uint8)
    fg = 0
    for b in range(B):
        sel = np.nonzero(rois[:, 0] == b)[0]
        ref = O.mask_targets(masks[b], rois[sel, 1:], gt_idx[sel], M, S)
        assert np.array_equal(got[sel], ref), b
        fg += int(ref.sum())
    assert fg > 1000 and not got[3].any() and not got[4].any()


def test_mask_targets_full_size():
    """config-4 size: 8 images x 128 positive RoIs on 800x1344 masks; identity crop reproduces the mask."""
    B, G, H, W = 2, 4, 800, 1344
    rng = np.random.default_rng(7)
    gts, _, _ = synth.gt_boxes(B, G=G, max_valid=G, seed=71)
    masks = blob_masks(rng, B, G, H, W, gts)
    # an RoI that spans exactly 28 pixels with S=2 samples at quarter-pixel offsets of a constant region -> all ones
    rois = np.array([[0, gts[0, 0, 0] * 0.5 + gts[0, 0, 2] * 0.5 - 2, gts[0, 0, 1] * 0.5 + gts[0, 0, 3] * 0.5 - 2,
                      gts[0, 0, 0] * 0.5 + gts[0, 0, 2] * 0.5 + 2, gts[0, 0, 1] * 0.5 + gts[0, 0, 3] * 0.5 + 2]], np.float32)
    got = MaskTargets(28, 2)(dev(masks).bool(), dev(rois), dev(np.zeros(1, np.int32)))
    assert bool(got.all())                      # centre of the blob: every sample is foreground
    R = 256
    rois = np.concatenate([rng.integers(0, B, (R, 1)).astype(np.float32), synth.rand_boxes(rng, R)], 1).astype(np.float32)
    gi = rng.integers(0, G, R).astype(np.int32)
    got = MaskTargets(28, 2)(dev(masks).bool(), dev(rois), dev(gi)).cpu().numpy().astype(np.uint8)
    for b in range(B):
        sel = np.nonzero(rois[:, 0] == b)[0]
        assert np.array_equal(got[sel], O.mask_targets(masks[b], rois[sel, 1:], gi[sel], 28, 2))
