"""Host-side mirror of the operator surface north_star names -- ``AnchorGenerator``, ``Proposal``,
``BboxAssignSample``, ``BboxAssignSampleForRcnn``, ``SingleRoIExtractor`` -- each a thin
``nn.Cell``-shaped class (``construct`` + ``__call__``) whose whole body is ONE aot call into
libmdregion.so, the way the reference wires its custom NMS into a graph
(centerpoint/det3d_ms/models/bbox_heads/center_head.py:435-459, ops/test_custom_pytorch/iou_gpu.py:46-80).

In the reference checkout these classes are empty stubs (minddet/models/faster_rcnn.py:1-3,
minddet/models/heads/roi_head.py:1-3); argument meaning follows SURVEY.md section 8(a).
Tensors are torch CUDA tensors here because MindSpore is not installable; `mindspore_ops.py` holds the
same classes over real ``mindspore.ops.Custom``.
"""
import math

import numpy as np
import torch

from ._aot import LIB_PATH, Custom

MAX_RATIO = float(np.float32(abs(math.log(0.016))))


def _so(symbol):
    return f"{LIB_PATH}:{symbol}"


class _Cell:
    def __call__(self, *a, **k):
        return self.construct(*a, **k)

    def _cfg(self, values, device, dtype=torch.float32):
        device = torch.device(device)
        if device.type == "cuda" and device.index is None:
            device = torch.device("cuda", torch.cuda.current_device())
        key = (str(device), dtype, tuple(values))
        cache = self.__dict__.setdefault("_cfg_cache", {})
        if key not in cache:
            cache[key] = torch.tensor(values, dtype=dtype, device=device)
        return cache[key]


class AnchorGenerator(_Cell):
    """a1.  mmdet-v1 style anchor generator (oracle/CONVENTIONS.md #7).  Base anchors are host/numpy
    (float64 -> round -> fp32), the grid is written by ``MdAnchorGrid``."""

    def __init__(self, base_size, scales, ratios, scale_major=True, ctr=None):
        self.base_size = base_size
        self.scales = np.asarray(scales, dtype=np.float64)
        self.ratios = np.asarray(ratios, dtype=np.float64)
        self.scale_major = scale_major
        self.ctr = ctr
        self.base_anchors = self.gen_base_anchors()
        self._grid = Custom(_so("MdAnchorGrid"), lambda base, cfg: None, torch.float32)

    @property
    def num_base_anchors(self):
        return self.base_anchors.shape[0]

    def gen_base_anchors(self):
        w = h = float(self.base_size)
        if self.ctr is None:
            x_ctr, y_ctr = 0.5 * (w - 1), 0.5 * (h - 1)
        else:
            x_ctr, y_ctr = self.ctr
        h_ratios = np.sqrt(self.ratios)
        w_ratios = 1 / h_ratios
        if self.scale_major:
            ws = (w * w_ratios[:, None] * self.scales[None, :]).reshape(-1)
            hs = (h * h_ratios[:, None] * self.scales[None, :]).reshape(-1)
        else:
            ws = (w * self.scales[:, None] * w_ratios[None, :]).reshape(-1)
            hs = (h * self.scales[:, None] * h_ratios[None, :]).reshape(-1)
        base = np.stack([x_ctr - 0.5 * (ws - 1), y_ctr - 0.5 * (hs - 1),
                         x_ctr + 0.5 * (ws - 1), y_ctr + 0.5 * (hs - 1)], axis=-1).round()
        return base.astype(np.float32)

    def base_tensor(self, device):
        cache = self.__dict__.setdefault("_base_cache", {})
        if str(device) not in cache:
            cache[str(device)] = torch.from_numpy(self.base_anchors).to(device)
        return cache[str(device)]

    def grid_anchors(self, featmap_size, stride=16, device="cuda"):
        feat_h, feat_w = featmap_size
        base = self.base_tensor(device)
        A = base.shape[0]
        self._grid.out_shape = lambda b, c: (feat_h, feat_w, A, 4)
        return self._grid(base, self._cfg([float(stride)], device)).reshape(-1, 4)

    construct = grid_anchors


def decode_cfg(img_shape, means, stds, max_ratio):
    return [float(img_shape[0]), float(img_shape[1])] + [float(m) for m in means] + [float(s) for s in stds] + [float(max_ratio)]


class BoundingBoxDecode(_Cell):
    """a2 on gathered rows (``MdDecodeClip``) or on a whole level from the head layout (``MdDecodeLevel``)."""

    def __init__(self, max_shape, means=(0.0, 0.0, 0.0, 0.0), stds=(1.0, 1.0, 1.0, 1.0), wh_ratio_clip=0.016):
        self.cfg_values = decode_cfg(max_shape, means, stds, float(np.float32(abs(math.log(wh_ratio_clip)))))
        self._rows = Custom(_so("MdDecodeClip"), lambda a, d, c: a, torch.float32)
        self._level = Custom(_so("MdDecodeLevel"), lambda d, b, c: (d[0], d[2] * d[3] * b[0], 4), torch.float32)

    def construct(self, anchors, deltas):
        return self._rows(anchors, deltas, self._cfg(self.cfg_values, anchors.device))

    def decode_level(self, deltas_nchw, base_anchors, stride):
        return self._level(deltas_nchw, base_anchors, self._cfg(self.cfg_values + [float(stride)], deltas_nchw.device))


class TopKPerLevel(_Cell):
    """a3.  ``TopK(sorted=True)`` over each image of one level; ties -> lower index first."""

    def __init__(self, k, apply_sigmoid=False):
        self.k = k
        self.apply_sigmoid = apply_sigmoid
        self._op = Custom(_so("MdTopKPerLevel"), None, (torch.float32, torch.int32))

    def construct(self, scores):
        B = scores.shape[0]
        n = scores[0].numel()
        k = min(self.k, n)
        self._op.out_shape = lambda s, c: ((B, k), (B, k))
        return self._op(scores, self._cfg([1.0 if self.apply_sigmoid else 0.0], scores.device))


class NMSWithMask(_Cell):
    """a4.  Greedy NMS on score-sorted boxes; returns (keep_idx, mask, count) like the reference's
    NmsNormalGpu returns (keep[N], num[1]) (iou_gpu.py:69-80) plus the NMSWithMask validity mask
    (pointpillars/src/core/nms.py:115-120)."""

    def __init__(self, iou_threshold=0.5, offset=0.0, inclusive=False, union_eps=1e-8):
        self.cfg_values = [float(iou_threshold), float(offset), 1.0 if inclusive else 0.0, float(union_eps)]
        self._op = Custom(_so("MdNms"), None, (torch.int32, torch.bool, torch.int32))

    def construct(self, boxes):
        lead = boxes.shape[:-1]
        B = boxes.shape[0] if boxes.dim() == 3 else 1
        self._op.out_shape = lambda b, c: (tuple(lead), tuple(lead), (B,))
        return self._op(boxes, self._cfg(self.cfg_values, boxes.device))


class Proposal(_Cell):
    """a3..a6.  ``construct(cls_scores, bbox_preds, anchor_generators)`` -> proposals (B,max_num,5), mask.

    cls_scores[l]: (B,A,H_l,W_l) logits (sigmoid applied inside when use_sigmoid_cls);
    bbox_preds[l]: (B,4A,H_l,W_l).  Anchors are regenerated on the device from the base anchors.
    """

    def __init__(self, img_shape, strides, base_anchors, nms_pre=2000, max_num=2000, nms_thr=0.7,
                 means=(0.0, 0.0, 0.0, 0.0), stds=(1.0, 1.0, 1.0, 1.0), use_sigmoid_cls=True,
                 nms_offset=0.0, nms_inclusive=False, union_eps=1e-8, max_ratio=MAX_RATIO):
        self.strides = [float(s) for s in strides]
        self.base_anchors = [np.asarray(b, np.float32) for b in base_anchors]
        self.nms_pre, self.max_num = nms_pre, max_num
        self.cfg_values = decode_cfg(img_shape, means, stds, max_ratio) + [
            float(nms_thr), float(nms_offset), 1.0 if nms_inclusive else 0.0, float(union_eps),
            1.0 if use_sigmoid_cls else 0.0] + self.strides
        self._op = Custom(_so("MdProposal"), None, (torch.float32, torch.bool, torch.int32, torch.bool))
        self.last_debug = None

    def _bases(self, device):
        cache = self.__dict__.setdefault("_base_cache", {})
        if str(device) not in cache:
            cache[str(device)] = [torch.from_numpy(b).to(device) for b in self.base_anchors]
        return cache[str(device)]

    def construct(self, cls_scores, bbox_preds):
        L = len(cls_scores)
        B = cls_scores[0].shape[0]
        dev = cls_scores[0].device
        self._op.out_shape = lambda *s: ((B, self.max_num, 5), (B, self.max_num), (B, L, self.nms_pre), (B, L, self.nms_pre))
        props, mask, topk_idx, keep = self._op(*cls_scores, *bbox_preds, *self._bases(dev), self._cfg(self.cfg_values, dev))
        self.last_debug = (topk_idx, keep)
        return props, mask


def _seed_values(seed, advance):
    """int32 seed tensor of the samplers: {seed_lo, seed_hi[, step]}.  With ``advance`` (default) the tensor has the third
    word: the op uses it as the per-call Philox counter and increments it ON THE DEVICE after every call, so consecutive
    calls -- and consecutive replays of a captured CUDA graph -- draw fresh samples, like the reference's per-call
    ``npr.choice`` (pointpillars/src/core/target_assigner.py:116-128).  ``advance=False`` freezes the draw (tests)."""
    s = int(seed) & 0xFFFFFFFFFFFFFFFF
    v = [np.int32(np.uint32(s & 0xFFFFFFFF)).item(), np.int32(np.uint32(s >> 32)).item()]
    return v + [0] if advance else v


class BboxAssignSample(_Cell):
    """a7/a8, RPN flavour.  ``construct(gt_bboxes, gt_valids, bboxes, valid_mask)``.

    Returns dict(assigned, pos_idx, pos_valid, neg_idx, neg_valid, pos_gt, pos_target, num_pos).
    """

    def __init__(self, pos_iou_thr=0.7, neg_iou_thr=0.3, min_pos_iou=0.3, num_expected_pos=128,
                 num_expected_neg=256, num_expected_total=256, means=(0.0, 0.0, 0.0, 0.0), stds=(1.0, 1.0, 1.0, 1.0),
                 seed=0, iou_offset=1.0, mode=0, force_full_scan=False, advance=True):
        self.Sp, self.Sn = num_expected_pos, num_expected_neg
        self.cfg_values = [float(pos_iou_thr), float(neg_iou_thr), float(min_pos_iou), float(iou_offset), float(mode),
                           float(num_expected_total)] + [float(m) for m in means] + [float(s) for s in stds] + [
                               1.0 if force_full_scan else 0.0, 0.0]
        self.seed_values = _seed_values(seed, advance)
        self._op = Custom(_so("MdAssignSample"), None,
                          (torch.int32, torch.int32, torch.bool, torch.int32, torch.bool, torch.int32, torch.float32, torch.int32))

    def seed_tensor(self, device):
        """the persistent device tensor {seed_lo, seed_hi[, step]}; ``seed_tensor(dev)[2]`` is the step the NEXT call uses"""
        return self._cfg(self.seed_values, device, torch.int32)

    def construct(self, gt_bboxes, gt_valids, bboxes, valid_mask):
        B, G = gt_bboxes.shape[:2]
        N = bboxes.shape[-2]
        Sp, Sn = self.Sp, self.Sn
        dev = gt_bboxes.device
        self._op.out_shape = lambda *s: ((B, N), (B, Sp), (B, Sp), (B, Sn), (B, Sn), (B, Sp), (B, Sp, 4), (B,))
        out = self._op(bboxes, valid_mask, gt_bboxes, gt_valids, self._cfg(self.cfg_values, dev),
                       self._cfg(self.seed_values, dev, torch.int32))
        return dict(zip(("assigned", "pos_idx", "pos_valid", "neg_idx", "neg_valid", "pos_gt", "pos_target", "num_pos"), out))


class BboxAssignSampleForRcnn(_Cell):
    """a8, stage-2 flavour: gts are prepended to the proposals as candidates.
    ``construct(gt_bboxes, gt_labels, proposal_mask, proposals, gt_valids)`` ->
    dict(rois (B,S,5), deltas, labels, mask, assigned, sel_idx, pos_gt, num_pos)."""

    def __init__(self, pos_iou_thr=0.5, neg_iou_thr=0.5, min_pos_iou=0.5, num_expected_pos=128,
                 num_expected_neg=384, num_expected_total=512, means=(0.0, 0.0, 0.0, 0.0), stds=(0.1, 0.1, 0.2, 0.2),
                 seed=0, iou_offset=1.0, mode=0, force_full_scan=False, advance=True):
        self.Sp, self.Sn = num_expected_pos, num_expected_neg
        self.cfg_values = [float(pos_iou_thr), float(neg_iou_thr), float(min_pos_iou), float(iou_offset), float(mode),
                           float(num_expected_total)] + [float(m) for m in means] + [float(s) for s in stds] + [
                               1.0 if force_full_scan else 0.0, 0.0]
        self.seed_values = _seed_values(seed, advance)
        self._op = Custom(_so("MdAssignSampleRcnn"), None,
                          (torch.float32, torch.float32, torch.int32, torch.bool, torch.int32, torch.int32, torch.int32, torch.int32))

    def seed_tensor(self, device):
        """the persistent device tensor {seed_lo, seed_hi[, step]}; ``seed_tensor(dev)[2]`` is the step the NEXT call uses"""
        return self._cfg(self.seed_values, device, torch.int32)

    def construct(self, gt_bboxes, gt_labels, proposal_mask, proposals, gt_valids):
        B, G = gt_bboxes.shape[:2]
        P = proposals.shape[1]
        Sp, S = self.Sp, self.Sp + self.Sn
        dev = gt_bboxes.device
        self._op.out_shape = lambda *s: ((B, S, 5), (B, S, 4), (B, S), (B, S), (B, G + P), (B, S), (B, Sp), (B,))
        out = self._op(proposals, proposal_mask, gt_bboxes, gt_labels, gt_valids, self._cfg(self.cfg_values, dev),
                       self._cfg(self.seed_values, dev, torch.int32))
        return dict(zip(("rois", "deltas", "labels", "mask", "assigned", "sel_idx", "pos_gt", "num_pos"), out))


class _RoIAlignFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, ext, rois, *feats):
        ctx.ext, ctx.shapes = ext, [tuple(f.shape) for f in feats]
        ctx.save_for_backward(rois)
        return ext._forward(rois, feats)

    @staticmethod
    def backward(ctx, dout):
        (rois,) = ctx.saved_tensors
        grads = ctx.ext._backward(rois, dout.contiguous(), ctx.shapes)
        return (None, None) + tuple(grads)


class SingleRoIExtractor(_Cell):
    """a9..a12.  ``construct(rois, feat1, ..., featL)`` -> (R,C,P,P); rois (R,5) = [batch,x1,y1,x2,y2].
    Only the level each RoI maps to is read.  The bprop (ROIAlignGrad) is ``MdRoiAlignBwd``."""

    def __init__(self, out_size=7, sample_num=2, featmap_strides=(4, 8, 16, 32), finest_scale=56, roi_end_mode=0,
                 exact=False):
        """exact=True selects the gather kernels only (forward bit-identical to the oracle's op order);
        the default is the TMA-staged separable path (rtol 1e-5 / atol 1e-6)."""
        self.P = out_size
        self.strides = [float(s) for s in featmap_strides]
        self.cfg_values = [float(finest_scale), float(sample_num), float(roi_end_mode), 0.0] + self.strides
        self._fwd = Custom(_so("MdRoiAlignFwdExact" if exact else "MdRoiAlignFwd"), None, torch.float32)
        self._bwd = Custom(_so("MdRoiAlignBwdExact" if exact else "MdRoiAlignBwd"), None, torch.float32)
        self._bwd_acc = None if exact else Custom(_so("MdRoiAlignBwdAcc"), lambda *s: (1,), torch.int32)
        self._lvl = Custom(_so("MdRoiLevels"), lambda r, c: (r[0],), torch.int32)
        self._prep = None if exact else Custom(_so("MdRoiAlignBwdPrepare"), None, torch.int32)
        self._planned = None if exact else Custom(_so("MdRoiAlignBwdPlanned"), None, torch.float32)

    def map_roi_levels(self, rois):
        return self._lvl(rois, self._cfg([self.cfg_values[0], float(len(self.strides))], rois.device))

    def _forward(self, rois, feats):
        R, C, P = rois.shape[0], feats[0].shape[1], self.P
        self._fwd.out_shape = lambda *s: (R, C, P, P)
        return self._fwd(rois, *feats, self._cfg(self.cfg_values, rois.device))

    def _backward(self, rois, dout, feat_shapes):
        self._bwd.out_shape = lambda *s: tuple(feat_shapes)
        out = self._bwd(rois, dout, self._cfg(self.cfg_values, rois.device))
        return out if isinstance(out, tuple) else (out,)

    def plan_words(self, R, feat_shapes):
        """int32 words of the plan tensor of the two-op backward (``MdRoiAlignPlanBytes``)."""
        import ctypes
        from ._aot import load_library
        L = len(feat_shapes)
        H = (ctypes.c_int * L)(*[int(s[2]) for s in feat_shapes])
        W = (ctypes.c_int * L)(*[int(s[3]) for s in feat_shapes])
        n = load_library().MdRoiAlignPlanBytes(int(R), int(feat_shapes[0][0]), int(feat_shapes[0][1]), L, H, W)
        if n < 0:
            raise ValueError("MdRoiAlignPlanBytes: bad arguments")
        return (int(n) + 3) // 4

    def prepare_backward(self, rois, feats):
        """Two-op bprop, first half (``MdRoiAlignBwdPrepare``): everything the tile-stationary backward derives from the RoIs, into a
        tensor the caller owns.  Needs the RoIs and the level shapes only, so it can run beside the forward on another stream."""
        if self._prep is None:
            raise RuntimeError("the exact RoIAlign variant has no two-op backward")
        n = self.plan_words(rois.shape[0], [tuple(f.shape) for f in feats])
        self._prep.out_shape = lambda *s: (n,)
        return self._prep(rois, *feats, self._cfg(self.cfg_values, rois.device))

    def _backward_planned(self, rois, dout, feat_shapes, plan):
        """Two-op bprop, second half (``MdRoiAlignBwdPlanned``): the backward proper on a prepared plan; every dX byte written."""
        self._planned.out_shape = lambda *s: tuple(feat_shapes)
        out = self._planned(rois, dout, self._cfg(self.cfg_values, rois.device), plan)
        return out if isinstance(out, tuple) else (out,)

    def _backward_into(self, rois, dout, grads):
        """Accumulating bprop (``MdRoiAlignBwdAcc``): grads_l += ROIAlignGrad(dout); the caller has initialised ``grads``
        (zeros for a plain bprop -- as a node of its own that zero-fill can run early, beside other work)."""
        if self._bwd_acc is None:
            raise RuntimeError("the exact RoIAlign variant has no accumulating backward")
        self._bwd_acc(rois, dout, self._cfg(self.cfg_values, rois.device), *grads)
        return tuple(grads)

    def construct(self, rois, *feats):
        if any(f.requires_grad for f in feats):
            return _RoIAlignFn.apply(self, rois, *feats)
        return self._forward(rois, feats)


class YoloV8PostProcess(_Cell):
    """a14 ("next" row 1).  ``construct(pred)``: pred (B, 64+nc, A) raw head output, anchors level after level ->
    (dets (B,max_det,6) [x1,y1,x2,y2,score,label], keep_idx (B,max_det), count (B)).  DFL decode of every anchor
    (``MdYoloDecode``) + class-aware NMS (``MdYoloNms``)."""

    def __init__(self, level_shapes, strides=(8, 16, 32), conf_thr=0.25, iou_thr=0.7, agnostic=False, nms_pre=2048,
                 max_det=300):
        self.level_shapes = [tuple(s) for s in level_shapes]
        self.nms_pre, self.max_det = nms_pre, max_det
        self.dec_cfg = [float(len(strides))] + [float(v) for (h, w), s in zip(self.level_shapes, strides) for v in (h, w, s)]
        self.nms_cfg = [float(conf_thr), float(iou_thr), 1.0 if agnostic else 0.0]
        self._dec = Custom(_so("MdYoloDecode"), lambda p, c: (p[0], p[2], 6), torch.float32)
        self._nms = Custom(_so("MdYoloNms"), None, (torch.float32, torch.int32, torch.int32, torch.int32))
        self.last_candidates = None

    def decode(self, pred):
        return self._dec(pred, self._cfg(self.dec_cfg, pred.device))

    def nms(self, dets):
        B = dets.shape[0]
        self._nms.out_shape = lambda d, c: ((B, self.max_det, 6), (B, self.max_det), (B,), (B, self.nms_pre))
        out, keep_idx, count, cand = self._nms(dets, self._cfg(self.nms_cfg, dets.device))
        self.last_candidates = cand
        return out, keep_idx, count

    def construct(self, pred):
        return self.nms(self.decode(pred))


class MaskTargets(_Cell):
    """a13 ("next" row 2).  ``construct(gt_masks (B,G,H,W) bool/uint8, rois (R,5), gt_idx (R) int32)`` ->
    (R,M,M) bool mask targets: the assigned gt's mask cropped to the RoI and resized by RoIAlign, >= 0.5."""

    def __init__(self, mask_size=28, sample_num=2):
        self.M = mask_size
        self.cfg_values = [float(sample_num)]
        self._op = Custom(_so("MdMaskTargets"), None, torch.bool)

    def construct(self, gt_masks, rois, gt_idx):
        R, M = rois.shape[0], self.M
        self._op.out_shape = lambda *s: (R, M, M)
        return self._op(gt_masks, rois, gt_idx, self._cfg(self.cfg_values, rois.device))


class BoundingBoxEncode(_Cell):
    """a2 inverse (legacy +1 bbox2delta) on rows: ``construct(proposals (K,4), gts (K,4))`` -> deltas (K,4)."""

    def __init__(self, means=(0.0, 0.0, 0.0, 0.0), stds=(1.0, 1.0, 1.0, 1.0)):
        self.cfg_values = [float(m) for m in means] + [float(s) for s in stds]
        self._op = Custom(_so("MdEncode"), lambda p, g, c: p, torch.float32)

    def construct(self, proposals, gts):
        return self._op(proposals, gts, self._cfg(self.cfg_values, proposals.device))


class RcnnPostProcess(_Cell):
    """"next" row 4: ``construct(rois (B,P,4|5), roi_valid (B,P), cls_logits (B,P,nc+1), bbox_deltas (B,P,(nc+1)*4))`` ->
    (dets (B,max_det,6) [x1,y1,x2,y2,score,label], keep_idx, count).  Softmax, per-class decode, score threshold,
    class-aware NMS, top max_det -- one aot call."""

    def __init__(self, img_shape, score_thr=0.05, iou_thr=0.5, max_det=100, nms_pre=2048, means=(0.0, 0.0, 0.0, 0.0),
                 stds=(0.1, 0.1, 0.2, 0.2), max_ratio=MAX_RATIO):
        self.max_det, self.nms_pre = max_det, nms_pre
        self.cfg_values = decode_cfg(img_shape, means, stds, max_ratio) + [float(score_thr), float(iou_thr)]
        self._op = Custom(_so("MdRcnnPostProcess"), None, (torch.float32, torch.int32, torch.int32, torch.int32))
        self.last_candidates = None

    def construct(self, rois, roi_valid, cls_logits, bbox_deltas):
        B = rois.shape[0]
        self._op.out_shape = lambda *s: ((B, self.max_det, 6), (B, self.max_det), (B,), (B, self.nms_pre))
        out, keep_idx, count, cand = self._op(rois, roi_valid, cls_logits, bbox_deltas, self._cfg(self.cfg_values, rois.device))
        self.last_candidates = cand
        return out, keep_idx, count
