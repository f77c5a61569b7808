"""CPU-only checks of the drop-in boundary: the C-ABI library loads, exports every symbol the public
header declares, validates its arguments before touching CUDA, and the host classes fail loudly
instead of falling back.  Also proves the ctypes harness speaks the real MindSpore aot ABI by calling
the REFERENCE's own compiled CPU op through it."""
import ctypes
import os
import re

import numpy as np
import pytest
import torch

import minddet_b200 as M
from minddet_b200 import _aot

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    hdr = open(os.path.join(ROOT, "include", "md_region_aot.h")).read()
    declared = re.findall(r"MD_API\s+int\s+(\w+)\(MD_AOT_ARGS\)", hdr)
    assert sorted(declared) == sorted(_aot.SYMBOLS) and len(declared) == 26
    assert "MdRoiAlignPlanBytes" in hdr and hasattr(M.load_library(), "MdRoiAlignPlanBytes")
    lib = M.load_library()
    for s in declared:
        assert hasattr(lib, s), s
    assert b"sm_100a" in lib.MdVersion()
    # the product library must not depend on torch, the oracle or any CPU path
    import subprocess
    needed = subprocess.run(["ldd", M.LIB_PATH], capture_output=True, text=True).stdout
    assert "torch" not in needed and "oracle" not in needed


def test_wrong_nparam_returns_1_like_the_reference():
    # iou-bev-nms-org.cpp:238: `if (nparam != 4) return 1;`
    lib = M.load_library()
    for s in _aot.SYMBOLS:
        assert getattr(lib, s)(0, None, None, None, None, None, None) == 1, s


def test_bad_dtype_or_shape_returns_2_without_touching_cuda():
    lib = M.load_library()
    n = 5
    params = (ctypes.c_void_p * n)(*[0] * n)
    ndims = (ctypes.c_int * n)(2, 1, 1, 1, 1)
    sh = [(ctypes.c_int64 * 2)(16, 5), (ctypes.c_int64 * 2)(4, 0), (ctypes.c_int64 * 2)(16, 0),
          (ctypes.c_int64 * 2)(16, 0), (ctypes.c_int64 * 2)(1, 0)]
    shapes = (ctypes.POINTER(ctypes.c_int64) * n)(*[ctypes.cast(a, ctypes.POINTER(ctypes.c_int64)) for a in sh])
    good = [b"float32", b"float32", b"int32", b"bool", b"int32"]
    bad = [b"float16", b"float32", b"int32", b"bool", b"int32"]
    assert lib.MdNms(n, params, ndims, shapes, (ctypes.c_char_p * n)(*bad), None, None) == 2
    sh[0][1] = 3   # fewer than 4 coordinates
    assert lib.MdNms(n, params, ndims, shapes, (ctypes.c_char_p * n)(*good), None, None) == 2
    sh[0][1] = 5
    sh[0][0] = 4097   # K > 4096 (64 mask words per row) is refused, not silently truncated
    sh[2][0] = sh[3][0] = 4097
    assert lib.MdNms(n, params, ndims, shapes, (ctypes.c_char_p * n)(*good), None, None) == 4


def test_host_classes_refuse_cpu_tensors_no_fallback():
    op = M.NMSWithMask(0.7)
    with pytest.raises(M.AotError):
        op(torch.zeros(8, 5))
    with pytest.raises(ImportError):
        _aot.load_library(os.path.join(ROOT, "minddet_b200", "lib", "does_not_exist.so"))


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "minddet_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert "import oracle" not in src and "from oracle" not in src and "liboracle" not in src, f


def test_anchor_generator_base_anchors_known_values():
    # mmdet-v1 formula, worked by hand: ws=16/sqrt(.5)*8=181.02 -> 7.5-0.5*180.02=-82.51 -> -83, ...
    g = M.AnchorGenerator.__new__(M.AnchorGenerator)
    g.base_size, g.scales, g.ratios, g.scale_major, g.ctr = 16, np.array([8.0]), np.array([0.5, 1.0, 2.0]), True, None
    assert np.array_equal(g.gen_base_anchors(), np.array([[-83, -37, 98, 52], [-56, -56, 71, 71], [-37, -83, 52, 98]], np.float32))


def test_harness_speaks_the_reference_aot_abi(golden):
    """Call the reference's compiled aot CPU NMS (oracle/_ref, built from
    centerpoint/det3d_ms/ops/iou-bev-nms-org.cpp) with the same 7-argument convention call_aot uses."""
    so = os.path.join(ROOT, "oracle", "_ref", "nms_fast_ref.so")
    if not os.path.exists(so):
        pytest.skip("oracle/_ref not built (reference absent)")
    lib = ctypes.CDLL(so)
    boxes = np.ascontiguousarray(golden["rotnms_boxes"])
    thr = np.ascontiguousarray(golden["rotnms_thr"])
    keep = np.zeros(1000, np.int32)
    cnt = np.zeros(1, np.int32)
    params = (ctypes.c_void_p * 4)(boxes.ctypes.data, thr.ctypes.data, keep.ctypes.data, cnt.ctypes.data)
    ndims = (ctypes.c_int * 4)(2, 1, 1, 1)
    assert lib.boxes_iou_nms_cpu(4, params, ndims, None, None, None, None) == 0
    assert cnt[0] == golden["rotnms_count"][0] and np.array_equal(keep, golden["rotnms_keep"])
    assert lib.boxes_iou_nms_cpu(3, params, ndims, None, None, None, None) == 1


def test_plan_bytes_is_a_host_function():
    """MdRoiAlignPlanBytes (size of the two-op backward's plan tensor) needs no GPU: positive, grows with R, -1 on bad arguments."""
    import ctypes
    lib = M.load_library()
    H = (ctypes.c_int * 4)(200, 100, 50, 25)
    W = (ctypes.c_int * 4)(336, 168, 84, 42)
    a = lib.MdRoiAlignPlanBytes(512, 1, 256, 4, H, W)
    b = lib.MdRoiAlignPlanBytes(4096, 8, 256, 4, H, W)
    assert 0 < a < b < (64 << 20)
    assert lib.MdRoiAlignPlanBytes(4096, 8, 256, 0, H, W) == -1
    assert lib.MdRoiAlignPlanBytes(-1, 8, 256, 4, H, W) == -1
    assert lib.MdRoiAlignPlanBytes(4096, 8, 256, 4, None, W) == -1
