"""The committed fixtures ARE outputs of the reference: where /root/reference exists (the build container), running the
committed generator scripts again reproduces every array bit for bit.  Skipped on the GPU box (no reference there)."""
import os
import subprocess
import sys

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)

pytestmark = pytest.mark.skipif(not os.path.isdir("/root/reference/minddet"), reason="the reference checkout is not on this box")


@pytest.mark.parametrize("script,fixture", [("make_golden.py", "reference_golden.npz"), ("make_bev_golden.py", "bev_golden.npz")])
def test_fixture_regenerates_bit_identical(tmp_path, script, fixture):
    import oracle
    oracle.build()                                   # oracle/_ref: the reference's files compiled from where they lie
    out = str(tmp_path / fixture)
    subprocess.check_call([sys.executable, os.path.join(HERE, "golden", script), out], cwd=ROOT,
                          stdout=subprocess.DEVNULL, timeout=600)
    new, old = np.load(out), np.load(os.path.join(HERE, "golden", fixture))
    assert sorted(new.files) == sorted(old.files)
    for k in old.files:
        assert new[k].dtype == old[k].dtype and new[k].shape == old[k].shape and np.array_equal(new[k], old[k]), k
