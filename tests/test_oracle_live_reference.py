"""Oracle vs the reference's own code, LIVE and randomised (many seeds beyond the committed fixtures): see
tests/golden/live_reference_check.py for what is compared.  Needs /root/reference, so it runs in the build container only."""
import os
import subprocess
import sys

import pytest

HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.mark.skipif(not os.path.isdir("/root/reference/minddet"), reason="the reference checkout is not on this box")
def test_oracle_agrees_with_the_reference_on_random_seeds():
    import oracle
    oracle.build()
    out = subprocess.run([sys.executable, "-W", "ignore", os.path.join(HERE, "golden", "live_reference_check.py"), "40"],
                         cwd=os.path.dirname(HERE), capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-2000:]
    assert "live reference check OK: 40 seeds" in out.stdout
