"""Golden fixtures for the rotated-BEV ops, produced by RUNNING THE REFERENCE's own CPU implementation:
/root/reference/minddet/models/centerpoint/det3d_ms/ops/iou-bev-nms-org.cpp compiled by oracle/Makefile into
oracle/_ref/nms_fast_ref.so (boxes_iou_nms_cpu :237-283 through the aot ABI; boxes_iou_bev_cpu :226-233 through its
mangled C++ name).  Run once in the build container:  python tests/golden/make_bev_golden.py
"""
import ctypes
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(os.path.dirname(HERE))
f32p = ctypes.POINTER(ctypes.c_float)


def make_boxes(rng, n, nclusters, sigma, zero_tail=0):
    ctr = rng.uniform([0, -40], [70, 40], (nclusters, 2))
    pick = rng.integers(0, nclusters, n)
    b = np.zeros((n, 7), np.float32)
    b[:, 0] = ctr[pick, 0] + rng.normal(0, sigma, n)
    b[:, 1] = ctr[pick, 1] + rng.normal(0, sigma, n)
    b[:, 3] = rng.uniform(1.5, 4.5, n)
    b[:, 4] = rng.uniform(1.2, 2.2, n)
    b[:, 5] = 1.5
    b[:, 6] = rng.uniform(-np.pi, np.pi, n)
    if zero_tail:
        b[-zero_tail:, 3] = 0.0
    return b


def main():
    lib = ctypes.CDLL(os.path.join(REPO, "oracle", "_ref", "nms_fast_ref.so"))
    iou_fn = getattr(lib, "_Z17boxes_iou_bev_cpuPKfiS0_iPf")
    iou_fn.argtypes = [f32p, ctypes.c_int, f32p, ctypes.c_int, f32p]
    rng = np.random.default_rng(20261018)
    out = {}
    # (N,M) IoU matrix
    a, b = make_boxes(rng, 300, 12, 1.5), make_boxes(rng, 200, 12, 1.5)
    b[:60] = a[:60] + rng.normal(0, 0.15, (60, 7)).astype(np.float32)      # heavy overlaps
    b[:, 3:5] = np.abs(b[:, 3:5])
    iou = np.zeros((300, 200), np.float32)
    iou_fn(a.ctypes.data_as(f32p), 300, b.ctypes.data_as(f32p), 200, iou.ctypes.data_as(f32p))
    out["iou_a"], out["iou_b"], out["iou_ref"] = a, b, iou
    # NMS cases (N = 1000: hard-coded in the reference)
    for tag, (ncl, sigma, thr, tail) in {"a": (40, 1.0, 0.2, 20), "b": (15, 0.8, 0.01, 0), "c": (60, 1.5, 0.5, 100),
                                         "d": (5, 0.5, 0.1, 7)}.items():
        for attempt in range(50):
            boxes = make_boxes(rng, 1000, ncl, sigma, tail)
            # full IoU matrix from the reference (for the strict '>' symbols and for the borderline check)
            m = np.zeros((1000, 1000), np.float32)
            iou_fn(boxes.ctypes.data_as(f32p), 1000, boxes.ctypes.data_as(f32p), 1000, m.ctypes.data_as(f32p))
            margin = np.abs(m[np.triu_indices(1000, 1)] - thr).min()
            if margin > 2e-5:       # no pair sits on the threshold: sinf/cosf ulp differences cannot flip a decision
                break
        t = np.array([thr], np.float32)
        keep = np.zeros(1000, np.int32)
        cnt = np.zeros(1, np.int32)
        params = (ctypes.c_void_p * 4)(boxes.ctypes.data, t.ctypes.data, keep.ctypes.data, cnt.ctypes.data)
        assert lib.boxes_iou_nms_cpu(4, params, None, None, None, None, None) == 0
        # greedy NMS with the strict '>' of the GPU symbols on the reference's own IoU values (NmsGpu semantics,
        # iou3d_nms_kernel.cu:300-344 + :526-536)
        alive = np.ones(1000, bool)
        kept = []
        for i in range(1000):
            if alive[i]:
                kept.append(i)
                alive[i + 1:] &= ~(m[i, i + 1:] > thr)
        k64 = np.zeros(1000, np.int64)
        k64[:len(kept)] = kept
        out[f"nms_{tag}_boxes"], out[f"nms_{tag}_thr"] = boxes, t
        out[f"nms_{tag}_keep_cpu"], out[f"nms_{tag}_count_cpu"] = keep, cnt
        out[f"nms_{tag}_keep_gt"], out[f"nms_{tag}_count_gt"] = k64, np.array([len(kept)], np.int32)
        out[f"nms_{tag}_margin"] = np.array([margin], np.float32)
        print(tag, "kept(cpu >=)", int(cnt[0]), "kept(>)", len(kept), "margin", margin, "attempts", attempt + 1)
    path = sys.argv[1] if len(sys.argv) > 1 else os.path.join(HERE, "bev_golden.npz")   # argv[1]: tests/test_golden_regenerates.py
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
