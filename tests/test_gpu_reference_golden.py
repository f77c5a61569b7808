"""The CUDA path against fixtures made by RUNNING THE REFERENCE's own code (tests/golden/make_golden.py):
  * the default NMS mode and the reference's own symbol NmsNormalGpu against `iou_normal` (iou3d_nms_kernel.cu:347-358, cut
    out of the file where it lies and compiled for the host: oracle/ref_cu_device_harness.cpp) -- keep lists bit-exact;
  * MdTopKPerLevel against the reference's numpy top-k (pointpillars/src/core/nms.py:66-83).
The CPU side of these pins (oracle == fixture) is tests/test_oracle_golden.py; the other golden-vector GPU tests
(nms_jit / apply_nms modes, assign mode 1, bilinear taps, rotated BEV) live in test_gpu_parity.py / test_gpu_bev.py.
Green on a B200: profiles/r2_gpu_tests_head.log."""
import numpy as np
import pytest
import torch

from minddet_b200 import NMSWithMask
from minddet_b200.bev_ops import NmsNormalGpu

pytestmark = pytest.mark.gpu


def dev(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


@pytest.mark.parametrize("tag", ["a", "b", "c", "d"])
def test_default_nms_mode_vs_reference_iou_normal_golden(golden, tag):
    # NMSWithMask default mode = offset 0, strict >, union guard 1e-8 (CONVENTIONS #1-2); rows are in score order
    xy = golden[f"ioun_{tag}_xyxy"]
    for thr in (0.3, 0.7):
        ref_keep = golden[f"ioun_{tag}_{thr}_keep"]
        _, mask, _ = NMSWithMask(float(np.float32(thr)), 0.0, False, 1e-8)(dev(xy))
        assert np.array_equal(np.nonzero(mask.cpu().numpy())[0], ref_keep), (tag, thr)


@pytest.mark.parametrize("tag", ["a", "b", "c", "d"])
def test_nms_normal_gpu_vs_reference_iou_normal_golden(golden, tag):
    # the reference's own symbol (NmsNormalGpu, iou3d_nms_kernel.cu:548-601) on the 7-float boxes
    b7 = golden[f"ioun_{tag}_box7"]
    for thr in (0.3, 0.7):
        ref_keep = golden[f"ioun_{tag}_{thr}_keep"]
        keep, num = NmsNormalGpu()(dev(b7), dev(np.array([thr], np.float32)))
        got = keep.cpu().numpy()
        assert int(num) == len(ref_keep), (tag, thr)
        assert np.array_equal(got[:len(ref_keep)], ref_keep) and not got[len(ref_keep):].any(), (tag, thr)


def test_topk_vs_reference_numpy_topk_golden(golden):
    # MdTopKPerLevel against outputs of the reference's own numpy top-k (pointpillars/src/core/nms.py:66-83), unique scores
    from minddet_b200 import TopKPerLevel
    sc, rv, ri = golden["topk_scores"], golden["topk_ref_vals"], golden["topk_ref_idx"]
    vals, idx = TopKPerLevel(len(ri))(dev(sc[None]))
    assert np.array_equal(idx[0].cpu().numpy(), ri.astype(np.int32))
    assert np.array_equal(vals[0].cpu().numpy(), rv)
