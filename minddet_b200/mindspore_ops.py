"""The same operator surface over REAL ``mindspore.ops.Custom(func_type="aot")`` -- what a minddet
maintainer drops into the graph.  Import-guarded: MindSpore is not installed in the build
environment, so this module is exercised only for syntax here; the call convention it relies on is
the one the reference already uses (centerpoint/det3d_ms/ops/test_custom_pytorch/iou_gpu.py:46-80).
See INTEGRATION.md.
"""
import math

import numpy as np

from ._aot import LIB_PATH

try:  # pragma: no cover - MindSpore is absent in this environment
    import mindspore as ms
    from mindspore import Tensor, nn, ops
    from mindspore import dtype as mstype
    HAVE_MINDSPORE = True
except ImportError:  # pragma: no cover
    HAVE_MINDSPORE = False

MAX_RATIO = float(np.float32(abs(math.log(0.016))))

if HAVE_MINDSPORE:  # pragma: no cover

    def _so(symbol):
        return f"{LIB_PATH}:{symbol}"

    class Proposal(nn.Cell):
        """construct(cls_scores: tuple, bbox_preds: tuple) -> (proposals (B,max_num,5), mask (B,max_num))"""

        def __init__(self, batch_size, img_shape, strides, base_anchors, nms_pre=2000, max_num=2000, nms_thr=0.7,
                     means=(0., 0., 0., 0.), stds=(1., 1., 1., 1.), use_sigmoid_cls=True):
            super().__init__()
            L = len(strides)
            self.bases = tuple(Tensor(np.asarray(b, np.float32)) for b in base_anchors)
            cfg = [img_shape[0], img_shape[1], *means, *stds, MAX_RATIO, nms_thr, 0.0, 0.0, 1e-8,
                   1.0 if use_sigmoid_cls else 0.0, *strides]
            self.cfg = Tensor(np.asarray(cfg, np.float32))
            B = batch_size
            self.op = ops.Custom(
                _so("MdProposal"),
                out_shape=lambda *s: ((B, max_num, 5), (B, max_num), (B, L, nms_pre), (B, L, nms_pre)),
                out_dtype=(mstype.float32, mstype.bool_, mstype.int32, mstype.bool_), func_type="aot")

        def construct(self, cls_scores, bbox_preds):
            out = self.op(*cls_scores, *bbox_preds, *self.bases, self.cfg)
            return out[0], out[1]

    class SingleRoIExtractor(nn.Cell):
        """construct(rois (R,5), feat1..featL) -> (R,C,P,P); bprop -> MdRoiAlignBwd"""

        def __init__(self, num_rois, channels, feat_shapes, out_size=7, sample_num=2, featmap_strides=(4, 8, 16, 32),
                     finest_scale=56, roi_end_mode=0):
            super().__init__()
            cfg = [finest_scale, sample_num, roi_end_mode, 0.0, *featmap_strides]
            self.cfg = Tensor(np.asarray(cfg, np.float32))
            P = out_size
            bwd = ops.Custom(_so("MdRoiAlignBwd"), out_shape=lambda *s: tuple(tuple(f) for f in feat_shapes),
                             out_dtype=tuple(mstype.float32 for _ in feat_shapes), func_type="aot")
            cfg_t = self.cfg

            def bprop(rois, *args):
                dout = args[-1]
                grads = bwd(rois, dout, cfg_t)
                return (ops.zeros_like(rois),) + tuple(grads) + (ops.zeros_like(cfg_t),)

            self.op = ops.Custom(_so("MdRoiAlignFwd"), out_shape=lambda *s: (num_rois, channels, P, P),
                                 out_dtype=mstype.float32, func_type="aot", bprop=bprop)

        def construct(self, rois, *feats):
            return self.op(rois, *feats, self.cfg)
