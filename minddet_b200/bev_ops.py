"""Host cells for the reference's own custom-op surface (rotated BEV IoU / NMS), same class names, argument
order and output conventions as minddet/models/centerpoint/det3d_ms/ops/test_custom_pytorch/iou_gpu.py:14-80
(`BoxesIouBevGpu`, `BoxesOverlapBevGpu`, `NumGpu` -> NmsGpu, `NmsNormalGpu`) and .../ops/nms_cpu.py:7-27
(`NmsCpu` -> here `NmsBevGpu`, the device twin).  Each body is ONE aot call into libmdregion.so."""
import torch

from ._aot import LIB_PATH, Custom


def _so(symbol):
    return f"{LIB_PATH}:{symbol}"


class _Cell:
    def __call__(self, *a, **k):
        return self.construct(*a, **k)


class BoxesIouBevGpu(_Cell):
    """in: boxes_a (N,7), boxes_b (M,7) f32 [x,y,z,dx,dy,dz,heading] -> ans_iou (N,M) f32  (iou_gpu.py:14-26)"""

    def __init__(self):
        self.boxes_iou = Custom(_so("BoxesIouBevGpu"), out_shape=lambda a, b: (a[0], b[0]), out_dtype=torch.float32)

    def construct(self, a, b):
        return self.boxes_iou(a, b)


class BoxesOverlapBevGpu(_Cell):
    """in: boxes_a (N,7), boxes_b (M,7) -> overlap area (N,M) f32  (iou_gpu.py:32-44)"""

    def __init__(self):
        self.boxes_overlap = Custom(_so("BoxesOverlapBevGpu"), out_shape=lambda a, b: (a[0], b[0]), out_dtype=torch.float32)

    def construct(self, a, b):
        return self.boxes_overlap(a, b)


class NumGpu(_Cell):
    """NmsGpu: boxes (N,7) score-sorted, thresh f32[1] -> keep (N) int64 zero padded, num_to_keep int32[1]
    (iou_gpu.py:50-66; the reference names the cell NumGpu)"""

    def __init__(self):
        self.nms_gpu = Custom(_so("NmsGpu"), out_shape=lambda boxes, thresh: ((boxes[0],), tuple(thresh)),
                              out_dtype=(torch.int64, torch.int32))

    def construct(self, boxes, thresh):
        return self.nms_gpu(boxes, thresh)


class NmsNormalGpu(_Cell):
    """axis-aligned variant on the same boxes (iou_gpu.py:72-80)"""

    def __init__(self):
        self.nms_normal_gpu = Custom(_so("NmsNormalGpu"), out_shape=lambda boxes, thresh: ((boxes[0],), tuple(thresh)),
                                     out_dtype=(torch.int64, torch.int32))

    def construct(self, boxes, thresh):
        return self.nms_normal_gpu(boxes, thresh)


class NmsBevGpu(_Cell):
    """Device twin of the reference's CPU aot op (`NmsCpu`, nms_cpu.py:7-27 -> boxes_iou_nms_cpu): keep (N) int32 +
    count int32[1]; IoU >= thresh suppresses; zero-area (padding) boxes are dropped first."""

    def __init__(self):
        self.nms = Custom(_so("BoxesIouNmsGpu"), out_shape=lambda x, _: ((x[0],), (1,)), out_dtype=(torch.int32, torch.int32))

    def construct(self, boxes, thresh):
        return self.nms(boxes, thresh)
