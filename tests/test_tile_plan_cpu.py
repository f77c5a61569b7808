"""Host logic of the tile-stationary RoIAlign backward (minddet_b200/csrc/roialign_tile_plan.h: per-RoI separable plans,
tile binning, visit clipping), compiled with g++ and accumulated on the CPU exactly the way the kernel walks its tiles,
against the oracle's o_roialign_bwd (1e-5 of the gradient scale).  No GPU needed."""
import os
import shutil
import subprocess

import pytest

import oracle as O

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("seed", [1, 2])
def test_tile_plan_logic_vs_oracle(tmp_path, seed):
    if shutil.which("g++") is None:
        pytest.skip("no g++")
    O.cpu.build()
    build = os.path.join(ROOT, "oracle", "_build")
    exe = str(tmp_path / "tile_plan_check")
    subprocess.check_call(["g++", "-O2", "-ffp-contract=off", "-o", exe, os.path.join(ROOT, "scripts", "tile_plan_check.cpp"),
                           os.path.join(build, "liboracle.so"), "-Wl,-rpath," + build])
    out = subprocess.run([exe, "1500", str(seed)], capture_output=True, text=True)
    assert out.returncode == 0, out.stdout + out.stderr
    assert "-> OK" in out.stdout
