"""minddet_b200 -- B200-native (sm_100a) Faster R-CNN region path behind the minddet operator surface.

Scope: anchors -> RPN decode/clip -> per-level top-k -> batched NMS -> MaxIoU assign/sample ->
FPN RoIAlign fwd/bwd, as ONE C-ABI library (lib/libmdregion.so, MindSpore ops.Custom "aot"
signature) plus the thin host classes that call it.  No CPU path, no fallback: importing the ops
without the built library raises.
"""
from .ops import (AnchorGenerator, BoundingBoxEncode, MaskTargets, RcnnPostProcess, BboxAssignSample, BboxAssignSampleForRcnn, BoundingBoxDecode,  # noqa: F401
                  NMSWithMask, Proposal, SingleRoIExtractor, TopKPerLevel, YoloV8PostProcess)
from ._aot import LIB_PATH, SYMBOLS, AotError, Custom, call_aot, load_library  # noqa: F401

__version__ = "0.1.0"
