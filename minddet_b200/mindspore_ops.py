"""The operator surface over REAL ``mindspore.ops.Custom(func_type="aot")`` -- what a minddet maintainer drops into the
graph: ``AnchorGenerator``, ``BoundingBoxDecode``, ``BoundingBoxEncode``, ``TopKPerLevel``, ``NMSWithMask``, ``Proposal``,
``BboxAssignSample``, ``BboxAssignSampleForRcnn``, ``SingleRoIExtractor`` (+ bprop), ``RoIAlignGradPlan`` /
``RoIAlignGradPlanned`` (the two-op bprop), ``MaskTargets``,
``YoloV8PostProcess``, ``RcnnPostProcess``.  Every cell is ``nn.Cell``-shaped and its ``construct`` is one or two aot calls
with static output shapes computed from the input shapes -- the call convention the reference already uses
(centerpoint/det3d_ms/ops/test_custom_pytorch/iou_gpu.py:46-80, ops/nms_cpu.py:10-27).

MindSpore is not installable in the build environment.  The module only needs the five names below from ``mindspore``;
``tests/ms_stub.py`` provides a stand-in whose ``ops.Custom`` marshals exactly what MindSpore's runtime passes (through
``minddet_b200._aot``) so that ``tests/test_gpu_mindspore_surface.py`` EXECUTES every cell of this file on the GPU and
checks it against the oracle.  See INTEGRATION.md.
"""
import math

import numpy as np

from ._aot import LIB_PATH

try:
    import mindspore as ms                                  # noqa: F401
    from mindspore import Tensor, nn, ops
    from mindspore import dtype as mstype
    HAVE_MINDSPORE = True
except ImportError:  # pragma: no cover
    HAVE_MINDSPORE = False

MAX_RATIO = float(np.float32(abs(math.log(0.016))))


def _so(symbol):
    return f"{LIB_PATH}:{symbol}"


def _decode_cfg(img_shape, means, stds, max_ratio):
    return [float(img_shape[0]), float(img_shape[1])] + [float(m) for m in means] + [float(s) for s in stds] + [float(max_ratio)]


def _seed_words(seed, advance):
    s = int(seed) & 0xFFFFFFFFFFFFFFFF
    v = [np.int32(np.uint32(s & 0xFFFFFFFF)).item(), np.int32(np.uint32(s >> 32)).item()]
    return v + [0] if advance else v


if HAVE_MINDSPORE:

    def _f32(values):
        return Tensor(np.asarray(values, np.float32))

    class AnchorGenerator(nn.Cell):
        """a1.  ``grid_anchors(featmap_size, stride)`` -> (H*W*A, 4); base anchors are host numpy (float64 -> round -> fp32)."""

        def __init__(self, base_size, scales, ratios, scale_major=True, ctr=None):
            super().__init__()
            w = h = float(base_size)
            x_ctr, y_ctr = (0.5 * (w - 1), 0.5 * (h - 1)) if ctr is None else ctr
            scales, ratios = np.asarray(scales, np.float64), np.asarray(ratios, np.float64)
            hr = np.sqrt(ratios)
            wr = 1 / hr
            if scale_major:
                ws, hs = (w * wr[:, None] * scales[None, :]).reshape(-1), (h * hr[:, None] * scales[None, :]).reshape(-1)
            else:
                ws, hs = (w * scales[:, None] * wr[None, :]).reshape(-1), (h * scales[:, None] * hr[None, :]).reshape(-1)
            base = np.stack([x_ctr - 0.5 * (ws - 1), y_ctr - 0.5 * (hs - 1), x_ctr + 0.5 * (ws - 1), y_ctr + 0.5 * (hs - 1)], -1).round()
            self.base_anchors = base.astype(np.float32)
            self.base = Tensor(self.base_anchors)

        def grid_anchors(self, featmap_size, stride=16):
            fh, fw = featmap_size
            A = self.base_anchors.shape[0]
            op = ops.Custom(_so("MdAnchorGrid"), out_shape=lambda b, c: (fh, fw, A, 4), out_dtype=mstype.float32, func_type="aot")
            return op(self.base, _f32([stride])).reshape(-1, 4)

        def construct(self, featmap_size, stride=16):
            return self.grid_anchors(featmap_size, stride)

    class BoundingBoxDecode(nn.Cell):
        """a2.  ``construct(anchors (K,4), deltas (K,4))`` -> boxes (K,4), legacy +1 delta2bbox + clip."""

        def __init__(self, max_shape, means=(0., 0., 0., 0.), stds=(1., 1., 1., 1.), wh_ratio_clip=0.016):
            super().__init__()
            self.cfg = _f32(_decode_cfg(max_shape, means, stds, float(np.float32(abs(math.log(wh_ratio_clip))))))
            self.op = ops.Custom(_so("MdDecodeClip"), out_shape=lambda a, d, c: a, out_dtype=mstype.float32, func_type="aot")

        def construct(self, anchors, deltas):
            return self.op(anchors, deltas, self.cfg)

    class BoundingBoxEncode(nn.Cell):
        """``construct(proposals (K,4), gts (K,4))`` -> deltas (K,4), legacy +1 bbox2delta."""

        def __init__(self, means=(0., 0., 0., 0.), stds=(1., 1., 1., 1.)):
            super().__init__()
            self.cfg = _f32([*means, *stds])
            self.op = ops.Custom(_so("MdEncode"), out_shape=lambda p, g, c: p, out_dtype=mstype.float32, func_type="aot")

        def construct(self, proposals, gts):
            return self.op(proposals, gts, self.cfg)

    class TopKPerLevel(nn.Cell):
        """a3.  ``construct(scores (B,A,H,W) | (B,N))`` -> (values (B,k), indices (B,k) int32); ties -> lower index."""

        def __init__(self, k, apply_sigmoid=False):
            super().__init__()
            self.k = k
            self.cfg = _f32([1.0 if apply_sigmoid else 0.0])

        def construct(self, scores):
            shp = tuple(scores.shape)
            n = int(np.prod(shp[1:]))
            k = min(self.k, n)
            op = ops.Custom(_so("MdTopKPerLevel"), out_shape=lambda s, c: ((s[0], k), (s[0], k)),
                            out_dtype=(mstype.float32, mstype.int32), func_type="aot")
            return op(scores, self.cfg)

    class NMSWithMask(nn.Cell):
        """a4.  ``construct(boxes (K,5) | (B,K,5) score-sorted)`` -> (keep_idx, mask, count)."""

        def __init__(self, iou_threshold=0.5, offset=0.0, inclusive=False, union_eps=1e-8):
            super().__init__()
            self.cfg = _f32([iou_threshold, offset, 1.0 if inclusive else 0.0, union_eps])
            self.op = ops.Custom(_so("MdNms"),
                                 out_shape=lambda b, c: (tuple(b[:-1]), tuple(b[:-1]), (b[0] if len(b) == 3 else 1,)),
                                 out_dtype=(mstype.int32, mstype.bool_, mstype.int32), func_type="aot")

        def construct(self, boxes):
            return self.op(boxes, self.cfg)

    class Proposal(nn.Cell):
        """a3..a6.  ``construct(cls_scores: tuple, bbox_preds: tuple)`` -> (proposals (B,max_num,5), mask (B,max_num))"""

        def __init__(self, batch_size, img_shape, strides, base_anchors, nms_pre=2000, max_num=2000, nms_thr=0.7,
                     means=(0., 0., 0., 0.), stds=(1., 1., 1., 1.), use_sigmoid_cls=True):
            super().__init__()
            L = len(strides)
            self.bases = tuple(Tensor(np.asarray(b, np.float32)) for b in base_anchors)
            self.cfg = _f32(_decode_cfg(img_shape, means, stds, MAX_RATIO) + [nms_thr, 0.0, 0.0, 1e-8, 1.0 if use_sigmoid_cls else 0.0,
                                                                              *[float(s) for s in strides]])
            B = batch_size
            self.op = ops.Custom(
                _so("MdProposal"),
                out_shape=lambda *s: ((B, max_num, 5), (B, max_num), (B, L, nms_pre), (B, L, nms_pre)),
                out_dtype=(mstype.float32, mstype.bool_, mstype.int32, mstype.bool_), func_type="aot")

        def construct(self, cls_scores, bbox_preds):
            out = self.op(*cls_scores, *bbox_preds, *self.bases, self.cfg)
            return out[0], out[1]

    class BboxAssignSample(nn.Cell):
        """a7/a8 (RPN).  ``construct(gt_bboxes (B,G,4), gt_valids (B,G), bboxes (N,4), valid_mask (N))`` ->
        (assigned (B,N), pos_idx, pos_valid, neg_idx, neg_valid, pos_gt, pos_target, num_pos).  The seed tensor is a
        member ({lo, hi, step}): the op advances ``step`` on the device after every call (md_region_aot.h)."""

        def __init__(self, pos_iou_thr=0.7, neg_iou_thr=0.3, min_pos_iou=0.3, num_expected_pos=128, num_expected_neg=256,
                     num_expected_total=256, means=(0., 0., 0., 0.), stds=(1., 1., 1., 1.), seed=0, iou_offset=1.0, mode=0,
                     advance=True):
            super().__init__()
            Sp, Sn = num_expected_pos, num_expected_neg
            self.cfg = _f32([pos_iou_thr, neg_iou_thr, min_pos_iou, iou_offset, mode, num_expected_total, *means, *stds, 0.0, 0.0])
            self.seed = Tensor(np.asarray(_seed_words(seed, advance), np.int32))
            self.op = ops.Custom(
                _so("MdAssignSample"),
                out_shape=lambda bx, vm, gt, gv, c, s: ((gt[0], bx[-2]), (gt[0], Sp), (gt[0], Sp), (gt[0], Sn), (gt[0], Sn),
                                                        (gt[0], Sp), (gt[0], Sp, 4), (gt[0],)),
                out_dtype=(mstype.int32, mstype.int32, mstype.bool_, mstype.int32, mstype.bool_, mstype.int32, mstype.float32,
                           mstype.int32), func_type="aot")

        def construct(self, gt_bboxes, gt_valids, bboxes, valid_mask):
            return self.op(bboxes, valid_mask, gt_bboxes, gt_valids, self.cfg, self.seed)

    class BboxAssignSampleForRcnn(nn.Cell):
        """a8 (stage 2).  ``construct(gt_bboxes, gt_labels, proposal_mask, proposals (B,P,5), gt_valids)`` ->
        (rois (B,S,5), deltas, labels, mask, assigned (B,G+P), sel_idx, pos_gt, num_pos)."""

        def __init__(self, pos_iou_thr=0.5, neg_iou_thr=0.5, min_pos_iou=0.5, num_expected_pos=128, num_expected_neg=384,
                     num_expected_total=512, means=(0., 0., 0., 0.), stds=(0.1, 0.1, 0.2, 0.2), seed=0, iou_offset=1.0, mode=0,
                     advance=True):
            super().__init__()
            Sp, S = num_expected_pos, num_expected_pos + num_expected_neg
            self.cfg = _f32([pos_iou_thr, neg_iou_thr, min_pos_iou, iou_offset, mode, num_expected_total, *means, *stds, 0.0, 0.0])
            self.seed = Tensor(np.asarray(_seed_words(seed, advance), np.int32))
            self.op = ops.Custom(
                _so("MdAssignSampleRcnn"),
                out_shape=lambda pr, pm, gt, gl, gv, c, s: ((pr[0], S, 5), (pr[0], S, 4), (pr[0], S), (pr[0], S),
                                                            (pr[0], gt[1] + pr[1]), (pr[0], S), (pr[0], Sp), (pr[0],)),
                out_dtype=(mstype.float32, mstype.float32, mstype.int32, mstype.bool_, mstype.int32, mstype.int32, mstype.int32,
                           mstype.int32), func_type="aot")

        def construct(self, gt_bboxes, gt_labels, proposal_mask, proposals, gt_valids):
            return self.op(proposals, proposal_mask, gt_bboxes, gt_labels, gt_valids, self.cfg, self.seed)

    class SingleRoIExtractor(nn.Cell):
        """a9..a12.  ``construct(rois (R,5), feat1..featL)`` -> (R,C,P,P); bprop = ``MdRoiAlignBwd`` (ROIAlignGrad)."""

        def __init__(self, out_size=7, sample_num=2, featmap_strides=(4, 8, 16, 32), finest_scale=56, roi_end_mode=0):
            super().__init__()
            self.cfg = _f32([finest_scale, sample_num, roi_end_mode, 0.0, *[float(s) for s in featmap_strides]])
            self.lvl_cfg = _f32([finest_scale, len(featmap_strides)])
            P, L = out_size, len(featmap_strides)

            def bprop(rois, *args):                    # args = feat_1..feat_L, cfg, out, dout
                feats, cfg_t, dout = args[:L], args[L], args[-1]
                # the aot signature has no "shape only" inputs: the feature shapes travel through out_shape's closure
                op = ops.Custom(_so("MdRoiAlignBwd"), out_shape=lambda r, d, c: tuple(tuple(f.shape) for f in feats),
                                out_dtype=tuple(mstype.float32 for _ in range(L)), func_type="aot")
                grads = op(rois, dout, cfg_t)
                grads = grads if isinstance(grads, tuple) else (grads,)
                return (ops.zeros_like(rois),) + tuple(grads) + (ops.zeros_like(cfg_t),)

            self.op = ops.Custom(_so("MdRoiAlignFwd"), out_shape=lambda r, *rest: (r[0], rest[0][1], P, P),
                                 out_dtype=mstype.float32, func_type="aot", bprop=bprop)
            self.lvl = ops.Custom(_so("MdRoiLevels"), out_shape=lambda r, c: (r[0],), out_dtype=mstype.int32, func_type="aot")

        def map_roi_levels(self, rois):
            return self.lvl(rois, self.lvl_cfg)

        def construct(self, rois, *feats):
            return self.op(rois, *feats, self.cfg)

    class RoIAlignGradPlan(nn.Cell):
        """a11, two-op form, first half (``MdRoiAlignBwdPrepare``).  ``construct(rois (R,5), feat1..featL)`` -> plan (int32): the
        per-RoI plans and per-tile RoI lists of the tile-stationary backward, in a tensor the graph owns.  Reads the RoIs and the
        SHAPES of the feature maps only, so the executor can place it beside the forward.  7x7 / 2 samples / C % 32 == 0."""

        def __init__(self, featmap_strides=(4, 8, 16, 32), finest_scale=56, roi_end_mode=0):
            super().__init__()
            self.cfg = _f32([finest_scale, 2, roi_end_mode, 0.0, *[float(s) for s in featmap_strides]])

        @staticmethod
        def plan_words(R, feat_shapes):
            import ctypes
            lib = ctypes.CDLL(LIB_PATH)
            lib.MdRoiAlignPlanBytes.restype = ctypes.c_int64
            L = len(feat_shapes)
            n = lib.MdRoiAlignPlanBytes(int(R), int(feat_shapes[0][0]), int(feat_shapes[0][1]), L,
                                        (ctypes.c_int * L)(*[int(s[2]) for s in feat_shapes]), (ctypes.c_int * L)(*[int(s[3]) for s in feat_shapes]))
            if n < 0:
                raise ValueError("MdRoiAlignPlanBytes: bad arguments")
            return (int(n) + 3) // 4

        def construct(self, rois, *feats):
            n = self.plan_words(rois.shape[0], [tuple(f.shape) for f in feats])
            op = ops.Custom(_so("MdRoiAlignBwdPrepare"), out_shape=lambda *s: (n,), out_dtype=mstype.int32, func_type="aot")
            return op(rois, *feats, self.cfg)

    class RoIAlignGradPlanned(nn.Cell):
        """a11, two-op form, second half (``MdRoiAlignBwdPlanned``).  ``construct(rois, dout (R,C,7,7), plan)`` -> per-level
        gradients, every byte written; ``feat_shapes`` are the (B,C,H_l,W_l) of the forward's inputs."""

        def __init__(self, feat_shapes, featmap_strides=(4, 8, 16, 32), finest_scale=56, roi_end_mode=0):
            super().__init__()
            self.cfg = _f32([finest_scale, 2, roi_end_mode, 0.0, *[float(s) for s in featmap_strides]])
            shapes = tuple(tuple(int(v) for v in s) for s in feat_shapes)
            self.op = ops.Custom(_so("MdRoiAlignBwdPlanned"), out_shape=lambda *s: shapes,
                                 out_dtype=tuple(mstype.float32 for _ in shapes), func_type="aot")

        def construct(self, rois, dout, plan):
            grads = self.op(rois, dout, self.cfg, plan)
            return grads if isinstance(grads, tuple) else (grads,)

    class MaskTargets(nn.Cell):
        """a13.  ``construct(gt_masks (B,G,H,W) bool, rois (R,5), gt_idx (R) int32)`` -> (R,M,M) bool."""

        def __init__(self, mask_size=28, sample_num=2):
            super().__init__()
            M = mask_size
            self.cfg = _f32([sample_num])
            self.op = ops.Custom(_so("MdMaskTargets"), out_shape=lambda m, r, g, c: (r[0], M, M), out_dtype=mstype.bool_, func_type="aot")

        def construct(self, gt_masks, rois, gt_idx):
            return self.op(gt_masks, rois, gt_idx, self.cfg)

    class YoloV8PostProcess(nn.Cell):
        """a14.  ``construct(pred (B, 64+nc, A))`` -> (dets (B,max_det,6), keep_idx (B,max_det), count (B))."""

        def __init__(self, level_shapes, strides=(8, 16, 32), conf_thr=0.25, iou_thr=0.7, agnostic=False, nms_pre=2048, max_det=300):
            super().__init__()
            self.dec_cfg = _f32([len(strides)] + [v for (h, w), s in zip(level_shapes, strides) for v in (h, w, s)])
            self.nms_cfg = _f32([conf_thr, iou_thr, 1.0 if agnostic else 0.0])
            self.dec = ops.Custom(_so("MdYoloDecode"), out_shape=lambda p, c: (p[0], p[2], 6), out_dtype=mstype.float32, func_type="aot")
            self.nms = ops.Custom(_so("MdYoloNms"), out_shape=lambda d, c: ((d[0], max_det, 6), (d[0], max_det), (d[0],), (d[0], nms_pre)),
                                  out_dtype=(mstype.float32, mstype.int32, mstype.int32, mstype.int32), func_type="aot")

        def construct(self, pred):
            out = self.nms(self.dec(pred, self.dec_cfg), self.nms_cfg)
            return out[0], out[1], out[2]

    class RcnnPostProcess(nn.Cell):
        """``construct(rois (B,P,4|5), roi_valid (B,P), cls_logits (B,P,nc+1), bbox_deltas (B,P,(nc+1)*4))`` ->
        (dets (B,max_det,6), keep_idx, count)."""

        def __init__(self, img_shape, score_thr=0.05, iou_thr=0.5, max_det=100, nms_pre=2048, means=(0., 0., 0., 0.),
                     stds=(0.1, 0.1, 0.2, 0.2)):
            super().__init__()
            self.cfg = _f32(_decode_cfg(img_shape, means, stds, MAX_RATIO) + [score_thr, iou_thr])
            self.op = ops.Custom(_so("MdRcnnPostProcess"),
                                 out_shape=lambda r, v, l, d, c: ((r[0], max_det, 6), (r[0], max_det), (r[0],), (r[0], nms_pre)),
                                 out_dtype=(mstype.float32, mstype.int32, mstype.int32, mstype.int32), func_type="aot")

        def construct(self, rois, roi_valid, cls_logits, bbox_deltas):
            out = self.op(rois, roi_valid, cls_logits, bbox_deltas, self.cfg)
            return out[0], out[1], out[2]
