"""Side outputs of the C++ 2D driver (off by default, NSB_OUTPUT=1): the reference's file names
(NavierStokes2D.cpp:669-692, :622-636).  Kept last in the run order: it is the least important check."""
import os
import subprocess

import pytest

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EXE = os.path.join(ROOT, "navierstokes_project_nm4pde_b200", "csrc", "host", "bin", "navier_stokes2D")


def test_driver_side_outputs(tmp_path):
    env = dict(os.environ, NSB_MAX_STEPS="2", NSB_OUTPUT="1")
    r = subprocess.run([EXE, "gen:cylinder2d:1"], capture_output=True, text=True, env=env, cwd=tmp_path, timeout=600)
    assert r.returncode == 0, r.stdout[-1500:] + r.stderr[-1500:]
    for step in (0, 1, 2):
        assert (tmp_path / "output2D_1" / f"output-navier-stokes-2D_{step:03d}.vtu").stat().st_size > 1000
    gm = (tmp_path / "gmres.csv").read_text().strip().splitlines()
    co = (tmp_path / "coeff_2.csv").read_text().strip().splitlines()
    assert len(gm) == 2 and len(co) == 3 and all(len(ln.split(",")) == 3 for ln in gm + co)
