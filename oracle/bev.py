"""Oracle side of the rotated-BEV ops (TEST INFRASTRUCTURE ONLY).

* rotated IoU / NMS: NOT restated -- the reference's own file
  (minddet/models/centerpoint/det3d_ms/ops/iou-bev-nms-org.cpp) is compiled from where it lies into
  oracle/_ref/nms_fast_ref.so by oracle/Makefile and called here (kind: "reference", parity PINNED).
* the GPU file's own __device__ functions (iou_bev, box_overlap, iou_normal of .../test_custom_pytorch/
  iou3d_nms_kernel.cu) are cut out of it and run on the host (ref_cu_pairs / ref_cu_nms below): on the fixtures they give
  the same bits as the CPU file, so the pin against the CPU file is a pin against the GPU file's arithmetic too.
* axis-aligned IoU on 7-float boxes (`iou_normal`, .../test_custom_pytorch/iou3d_nms_kernel.cu:347-358) needs a GPU +
  libtorch to run in the reference as a whole, so it is restated in numpy fp32 with the same operation order; the
  restatement is pinned bit-exact (lattice boxes, tests/test_oracle_golden.py) to that one function cut out of the
  reference file and compiled for the host (oracle/ref_cu_device_harness.cpp -> oracle/_ref/iou3d_device_ref.so).
"""
import ctypes
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
REF_SO = os.path.join(_HERE, "_ref", "nms_fast_ref.so")
_f32p = ctypes.POINTER(ctypes.c_float)


def have_ref():
    return os.path.exists(REF_SO)


def ref_iou_bev(a, b):
    """boxes_iou_bev_cpu (iou-bev-nms-org.cpp:226-233) -> (N,M) fp32"""
    lib = ctypes.CDLL(REF_SO)
    fn = getattr(lib, "_Z17boxes_iou_bev_cpuPKfiS0_iPf")
    fn.argtypes = [_f32p, ctypes.c_int, _f32p, ctypes.c_int, _f32p]
    a, b = np.ascontiguousarray(a, np.float32), np.ascontiguousarray(b, np.float32)
    out = np.zeros((a.shape[0], b.shape[0]), np.float32)
    fn(a.ctypes.data_as(_f32p), a.shape[0], b.ctypes.data_as(_f32p), b.shape[0], out.ctypes.data_as(_f32p))
    return out


def ref_nms_cpu(boxes, thr):
    """boxes_iou_nms_cpu (iou-bev-nms-org.cpp:237-283) through the aot ABI; boxes must be (1000,7)."""
    assert boxes.shape == (1000, 7), "the reference hard-codes N = 1000 (iou-bev-nms-org.cpp:244)"
    lib = ctypes.CDLL(REF_SO)
    boxes = np.ascontiguousarray(boxes, np.float32)
    t = np.array([thr], np.float32)
    keep, cnt = np.zeros(1000, np.int32), np.zeros(1, np.int32)
    params = (ctypes.c_void_p * 4)(boxes.ctypes.data, t.ctypes.data, keep.ctypes.data, cnt.ctypes.data)
    assert lib.boxes_iou_nms_cpu(4, params, None, None, None, None, None) == 0
    return keep, int(cnt[0])


REF_CU_SO = os.path.join(_HERE, "_ref", "iou3d_device_ref.so")


def have_ref_cu():
    return os.path.exists(REF_CU_SO)


def _ref_cu():
    lib = ctypes.CDLL(REF_CU_SO)
    vp = ctypes.c_void_p
    lib.ref_cu_pair_matrix.argtypes = [ctypes.c_int, vp, ctypes.c_int, vp, ctypes.c_int, vp]
    lib.ref_cu_nms.argtypes = [ctypes.c_int, vp, ctypes.c_int, ctypes.c_float, vp, ctypes.POINTER(ctypes.c_int)]
    return lib


def ref_cu_pairs(which, a, b):
    """The reference's GPU-side __device__ functions run on the host (oracle/ref_cu_device_harness.cpp):
    which = "iou_normal" (iou3d_nms_kernel.cu:347-358), "iou_bev" (:258-265) or "box_overlap" (:135-256) -> (N,M) fp32"""
    a, b = np.ascontiguousarray(a, np.float32), np.ascontiguousarray(b, np.float32)
    out = np.zeros((a.shape[0], b.shape[0]), np.float32)
    _ref_cu().ref_cu_pair_matrix({"iou_normal": 0, "iou_bev": 1, "box_overlap": 2}[which], a.ctypes.data, a.shape[0],
                                 b.ctypes.data, b.shape[0], out.ctypes.data)
    return out


def ref_cu_nms(boxes, thr, rotated=True):
    """greedy sweep (strict >) over the reference's own iou_bev / iou_normal: NmsGpu / NmsNormalGpu keep lists"""
    boxes = np.ascontiguousarray(boxes, np.float32)
    keep, cnt = np.zeros(boxes.shape[0], np.int64), ctypes.c_int(0)
    _ref_cu().ref_cu_nms(1 if rotated else 0, boxes.ctypes.data, boxes.shape[0], float(np.float32(thr)), keep.ctypes.data,
                         ctypes.byref(cnt))
    return keep[:cnt.value].copy()


def greedy_from_iou(iou, thr, inclusive=False):
    """keep list of the bitmask NMS (iou3d_nms_kernel.cu:300-344 + host reduce :526-536) given pairwise IoUs"""
    n = iou.shape[0]
    alive = np.ones(n, bool)
    kept = []
    for i in range(n):
        if alive[i]:
            kept.append(i)
            sup = iou[i, i + 1:] >= thr if inclusive else iou[i, i + 1:] > thr
            alive[i + 1:] &= ~sup
    return np.asarray(kept, np.int64)


def iou_normal(a, b):
    """iou_normal (iou3d_nms_kernel.cu:347-358), fp32, same operation order; a (N,7), b (M,7) -> (N,M)"""
    a, b = np.asarray(a, np.float32), np.asarray(b, np.float32)
    h = np.float32(2)
    al, ar = a[:, 0] - a[:, 3] / h, a[:, 0] + a[:, 3] / h
    at, ab = a[:, 1] - a[:, 4] / h, a[:, 1] + a[:, 4] / h
    bl, br = b[:, 0] - b[:, 3] / h, b[:, 0] + b[:, 3] / h
    bt, bb = b[:, 1] - b[:, 4] / h, b[:, 1] + b[:, 4] / h
    left, right = np.maximum(al[:, None], bl[None]), np.minimum(ar[:, None], br[None])
    top, bottom = np.maximum(at[:, None], bt[None]), np.minimum(ab[:, None], bb[None])
    w = np.maximum(right - left, np.float32(0))
    hgt = np.maximum(bottom - top, np.float32(0))
    inter = (w * hgt).astype(np.float32)
    sa, sb = (a[:, 3] * a[:, 4]).astype(np.float32), (b[:, 3] * b[:, 4]).astype(np.float32)
    union = np.maximum((sa[:, None] + sb[None]).astype(np.float32) - inter, np.float32(1e-8))
    return (inter / union).astype(np.float32)
