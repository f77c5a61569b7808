"""RegionPath: the Faster R-CNN region path of one training step, wired from the host classes the way
the upstream graph wires it (Proposal -> BboxAssignSample (RPN targets) -> BboxAssignSampleForRcnn ->
SingleRoIExtractor fwd -> ROIAlignGrad).  Template caller in the reference:
centerpoint/det3d_ms/models/bbox_heads/center_head.py:398-463 (top-k -> custom NMS -> gather)."""
import numpy as np
import torch

from . import synth
from .ops import (AnchorGenerator, BboxAssignSample, BboxAssignSampleForRcnn, Proposal, SingleRoIExtractor)

KERNELS_PER_STEP = 28   # 2 lanes x (select, nms_mask, nms_sweep), merge | gtmax, label, prefilter, select(list), select(full scan,
                        # returns at once), finalize | gtmax (+ gt head), label, prefilter, select x2, finalize |
                        # roialign fwd (stream + gather for declined RoIs) | roialign bwd (tile-stationary: plan, offsets, scatter,
                        # sort / zero, tile kernel, gather); memset nodes (small counters) are not counted.  bench.py counts the
                        # kernel nodes of the captured graph itself.


class RegionPath:
    def __init__(self, img_shape=(synth.IMG_H, synth.IMG_W), strides=synth.STRIDES, roi_levels=4,
                 nms_pre=2000, max_num=2000, nms_thr=0.7, roi_pos=128, roi_neg=384, roi_total=512,
                 out_size=7, sample_num=2, seed=0, device="cuda", advance=True):
        self.img_shape, self.strides, self.device = img_shape, tuple(strides), device
        self.shapes = synth.level_shapes(img_shape[0], img_shape[1], strides)
        self.bases = synth.base_anchor_sets(strides)
        self.generators = [AnchorGenerator(s, [8], [0.5, 1.0, 2.0]) for s in strides]
        self.proposal = Proposal(img_shape, strides, self.bases, nms_pre=nms_pre, max_num=max_num, nms_thr=nms_thr)
        self.rpn_targets = BboxAssignSample(0.7, 0.3, 0.3, 128, 256, 256, seed=seed, advance=advance)
        self.rcnn_targets = BboxAssignSampleForRcnn(0.5, 0.5, 0.5, roi_pos, roi_neg, roi_total, seed=seed, advance=advance)
        self.extractor = SingleRoIExtractor(out_size, sample_num, strides[:roi_levels], 56)
        self.roi_levels = roi_levels
        self._anchors = None
        self._anchor_valid = None

    def anchors(self):
        """All-level anchors (N,4), generated once on the device and cached (the reference caches its
        anchors once too: pointpillars/src/data/dataset.py:27-40)."""
        if self._anchors is None:
            parts = [g.grid_anchors(hw, s, self.device) for g, hw, s in zip(self.generators, self.shapes, self.strides)]
            self._anchors = torch.cat(parts).contiguous()
            self._anchor_valid = torch.ones(self._anchors.shape[0], dtype=torch.bool, device=self.device)
        return self._anchors, self._anchor_valid

    def forward(self, cls_scores, bbox_preds, feats, gts, gt_labels, gt_valid):
        anchors, avalid = self.anchors()
        props, pmask = self.proposal(cls_scores, bbox_preds)
        rpn = self.rpn_targets(gts, gt_valid, anchors, avalid)
        rcnn = self.rcnn_targets(gts, gt_labels, pmask, props, gt_valid)
        rois = rcnn["rois"].reshape(-1, 5)
        roi_feats = self.extractor._forward(rois, feats[:self.roi_levels])
        return dict(props=props, pmask=pmask, rpn=rpn, rcnn=rcnn, rois=rois, roi_feats=roi_feats)

    def backward(self, rois, dout, feat_shapes):
        return self.extractor._backward(rois, dout, feat_shapes)

    def step(self, cls_scores, bbox_preds, feats, gts, gt_labels, gt_valid, dout):
        out = self.forward(cls_scores, bbox_preds, feats, gts, gt_labels, gt_valid)
        out["dfeats"] = self.backward(out["rois"], dout, [tuple(f.shape) for f in feats[:self.roi_levels]])
        return out


def make_inputs(B, C=256, seed=0xD37, roi_levels=4, slots=512, P=7, shapes=None, pin=False):
    """Host (numpy -> torch CPU, optionally pinned) synthetic inputs of config 2."""
    shapes = shapes or synth.level_shapes()
    logits, deltas = synth.rpn_head_outputs(B, shapes, 3, seed)
    gts, labels, valid = synth.gt_boxes(B, G=128, seed=seed)
    feats = synth.features(B, shapes[:roi_levels], C, seed)
    rng = np.random.default_rng(seed + 3)
    dout = rng.uniform(-1, 1, (B * slots, C, P, P)).astype(np.float32)

    def t(a):
        x = torch.from_numpy(np.ascontiguousarray(a))
        return x.pin_memory() if pin else x

    return dict(cls_scores=[t(x) for x in logits], bbox_preds=[t(x) for x in deltas], feats=[t(x) for x in feats],
                gts=t(gts), gt_labels=t(labels), gt_valid=t(valid.astype(bool)), dout=t(dout))


def to_device(inp, device="cuda", non_blocking=True):
    out = {}
    for k, v in inp.items():
        out[k] = [x.to(device, non_blocking=non_blocking) for x in v] if isinstance(v, list) else v.to(device, non_blocking=non_blocking)
    return out


def input_bytes(inp):
    n = 0
    for v in inp.values():
        for x in (v if isinstance(v, list) else [v]):
            n += x.numel() * x.element_size()
    return n
