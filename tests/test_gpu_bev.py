"""GPU parity of the rotated-BEV symbols against outputs of the reference's OWN CPU
implementation (tests/golden/bev_golden.npz, made by tests/golden/make_bev_golden.py from the compiled
iou-bev-nms-org.cpp) and, when oracle/_ref travelled to the box, against the compiled reference live."""
import os

import numpy as np
import pytest
import torch

from minddet_b200.bev_ops import BoxesIouBevGpu, BoxesOverlapBevGpu, NmsBevGpu, NmsNormalGpu, NumGpu
from oracle import bev as OB

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def bevg():
    return np.load(os.path.join(ROOT, "tests", "golden", "bev_golden.npz"))


def dev(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


def test_iou_and_overlap_matrices_vs_reference(bevg):
    a, b, ref = bevg["iou_a"], bevg["iou_b"], bevg["iou_ref"]
    iou = BoxesIouBevGpu()(dev(a), dev(b)).cpu().numpy()
    # trig functions differ in the last ulp between glibc and CUDA: tolerance, not bit-exactness, for the fp values
    np.testing.assert_allclose(iou, ref, rtol=1e-4, atol=2e-6)
    assert (ref > 0.3).sum() > 50            # the fixture really contains heavy overlaps
    ov = BoxesOverlapBevGpu()(dev(a), dev(b)).cpu().numpy()
    sa, sb = a[:, 3] * a[:, 4], b[:, 3] * b[:, 4]
    ov_ref = ref * (sa[:, None] + sb[None]) / (1.0 + ref)          # iou = s / (sa + sb - s)
    np.testing.assert_allclose(ov, ov_ref, rtol=1e-4, atol=1e-5)


@pytest.mark.parametrize("tag", ["a", "b", "c", "d"])
def test_rotated_nms_keep_indices_vs_reference(bevg, tag):
    boxes, thr = bevg[f"nms_{tag}_boxes"], bevg[f"nms_{tag}_thr"]
    # device twin of boxes_iou_nms_cpu: '>=' + zero-area pre-removal, int32 keep + count (bit-exact)
    keep, cnt = NmsBevGpu()(dev(boxes), dev(thr))
    assert int(cnt) == int(bevg[f"nms_{tag}_count_cpu"][0])
    assert np.array_equal(keep.cpu().numpy(), bevg[f"nms_{tag}_keep_cpu"])
    # NmsGpu: strict '>' on the reference's IoU values, int64 keep zero padded + num_to_keep
    keep64, num = NumGpu()(dev(boxes), dev(thr))
    assert keep64.dtype == torch.int64 and int(num) == int(bevg[f"nms_{tag}_count_gt"][0])
    assert np.array_equal(keep64.cpu().numpy(), bevg[f"nms_{tag}_keep_gt"])


def test_rotated_nms_live_against_compiled_reference():
    if not OB.have_ref():
        pytest.skip("oracle/_ref/nms_fast_ref.so did not travel")
    rng = np.random.default_rng(77)
    from tests.golden.make_bev_golden import make_boxes
    for thr in (0.2, 0.35):
        boxes = make_boxes(rng, 1000, 30, 1.2, 13)
        m = OB.ref_iou_bev(boxes, boxes)
        if np.abs(m[np.triu_indices(1000, 1)] - thr).min() < 2e-5:
            continue                          # a pair sits on the threshold: ulp noise could legitimately flip it
        keep_ref, cnt_ref = OB.ref_nms_cpu(boxes, thr)
        keep, cnt = NmsBevGpu()(dev(boxes), dev(np.array([thr], np.float32)))
        assert int(cnt) == cnt_ref and np.array_equal(keep.cpu().numpy(), keep_ref)


@pytest.mark.parametrize("n", [1, 63, 64, 65, 1000, 2048, 2100, 4096])
def test_normal_nms_vs_numpy_restatement(n):
    rng = np.random.default_rng(100 + n)
    from tests.golden.make_bev_golden import make_boxes
    boxes = make_boxes(rng, n, max(1, n // 25), 1.0)
    thr = np.array([0.25], np.float32)
    keep, num = NmsNormalGpu()(dev(boxes), dev(thr))
    ref = OB.greedy_from_iou(OB.iou_normal(boxes, boxes), thr[0])
    assert int(num) == len(ref)
    got = keep.cpu().numpy()
    assert np.array_equal(got[:len(ref)], ref) and not got[len(ref):].any()


def test_bev_nms_error_codes_and_sizes():
    from minddet_b200 import AotError
    boxes = torch.zeros(4097, 7, device="cuda")
    with pytest.raises(AotError):
        NumGpu()(boxes, torch.tensor([0.2], device="cuda"))          # > 4096 boxes: unsupported size (rc 4)
    with pytest.raises(AotError):
        BoxesIouBevGpu()(torch.zeros(4, 6, device="cuda"), torch.zeros(4, 7, device="cuda"))   # bad shape (rc 2)
