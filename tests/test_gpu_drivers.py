"""The three drop-in driver scripts (drivers/*.py) executed as a user would -- `python <script>.py` in a working
directory -- on the GPU, and their result files / progress lines compared with what the unmodified upstream scripts
produce (goldens from the reference: tests/golden/)."""
import hashlib
import os
import re
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = os.path.join(ROOT, "tests", "golden")
X_AXIS_MD5 = "fcc9f2c9d5dbc0d99c4cb1aa25229a06"     # the four shipped hist_x_axis_* files are byte-identical


def run_driver(name, workdir, *args):
    r = subprocess.run([sys.executable, os.path.join(ROOT, "drivers", name), *args], cwd=str(workdir), capture_output=True,
                       text=True, timeout=1500)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    return r.stdout


def md5(path):
    return hashlib.md5(open(path, "rb").read()).hexdigest()


def test_cube_driver_full_run(tmp_path, oracle, cube_cfg, cube_init):
    """Open_Air_Cube_MC.py as shipped: 500 timesteps, serial sweep.  The run is chaotic: the reference's NumPy-scalar
    `v**2` (libm pow) differs from `v*v` in the last bit for 0.07 % of the inputs, which changes the collision sequence
    after ~100 steps, so the eight result files are compared byte for byte with the ones the CPU oracle writes in the
    same plain arithmetic, and with the unmodified script's own output where that is meaningful: the
    first steps' collision counts exactly, the totals statistically."""
    from argon_monte_carlo_b200 import outputs
    from oracle import steps
    out = run_driver("Open_Air_Cube_MC.py", tmp_path)
    assert out.splitlines()[0].strip() == "24627"
    cols = [int(v) for v in re.findall(r"^\s+(\d+)\s+collisions$", out, flags=re.M)]
    assert len(cols) == 500 and cols[:10] == [31, 41, 36, 50, 44, 45, 47, 49, 60, 53]      # printed by the reference
    assert abs(sum(cols) - 24382) < 0.02 * 24382                                            # reference: 24,382
    n_paths = int(re.search(r"Num of collisions total: (\d+)", out).group(1))
    assert abs(n_paths - 27448) < 0.02 * 27448                                              # reference: 27,448
    mfp = float(re.search(r"Simulation 1 mean free path: ([0-9.e+-]+)", out).group(1))
    assert abs(mfp - 3.5298513644857337e-07) < 0.03 * 3.53e-07
    # the same 500 steps through the oracle (plain arithmetic = the kernels' arithmetic)
    st = oracle.ParticleState(*cube_init)
    sink = oracle.PathSink()
    ocols = [steps.cube_step(st, cube_cfg, sink)["pp_collisions"] for _ in range(500)]
    assert cols == ocols and n_paths == len(sink)
    ref_dir = tmp_path / "oracle"
    ref_dir.mkdir()
    counts = [oracle.histogram(a) for a in sink.arrays()]
    import numpy as np
    outputs.write_histograms(np.array(counts, dtype=np.uint64), str(ref_dir))
    for ax in ("total", "x", "y", "z"):
        assert md5(tmp_path / ("hist_x_axis_%s_data.txt" % ax)) == X_AXIS_MD5
        assert md5(tmp_path / ("hist_y_axis_%s_data.txt" % ax)) == md5(ref_dir / ("hist_y_axis_%s_data.txt" % ax))


def test_pore_driver_three_steps(tmp_path):
    """Open_Air_Pore_MC.py, first three timesteps: the progress lines carry the reference's counts
    (tests/golden/ref_pore_3steps.json: 1069 / 261 / 414 collisions)."""
    out = run_driver("Open_Air_Pore_MC.py", tmp_path, "--steps", "3")
    assert [int(v) for v in re.findall(r"^\s+(\d+)\s+collisions from this timestep", out, flags=re.M)] == [1069, 261, 414]
    assert "There are 0 particles out of bounds after initialization." in out
    for ax in ("total", "x", "y", "z"):
        assert md5(tmp_path / ("hist_x_axis_%s_data.txt" % ax)) == X_AXIS_MD5
        assert os.path.getsize(tmp_path / ("hist_y_axis_%s_data.txt" % ax)) > 0


@pytest.mark.parametrize("rng", ["host", "device"])
def test_temp_driver_two_steps(tmp_path, rng):
    """Temperature_Pore_MC.py: --rng host reproduces rows 0-1 of the shipped momentum_energy.csv digit for digit and
    the reference's collision counts (1402 / 563); --rng device gives the same file layout from the device stream."""
    out = run_driver("Temperature_Pore_MC.py", tmp_path, "--steps", "2", "--rng", rng)
    cols = [int(v) for v in re.findall(r"^\s+(\d+)\s+collisions from this timestep", out, flags=re.M)]
    rows = open(tmp_path / "momentum_energy.csv").read().split("\n")
    assert rows[0] == ",Momentum,EnergyCold,EnergyHot" and len([r for r in rows if r]) == 3
    if rng == "host":
        assert cols == [1402, 563]
        assert rows[:3] == open(os.path.join(GOLD, "shipped_momentum_energy.csv")).read().split("\n")[:3]
    else:
        assert len(cols) == 2 and 1300 < cols[0] < 1500
        for r in rows[1:3]:
            k, p, ec, eh = r.split(",")
            assert float(p) != 0 and float(ec) < 0 and float(eh) < 0
    for ax in ("total", "x", "y", "z"):
        assert md5(tmp_path / ("hist_x_axis_%s_data.txt" % ax)) == X_AXIS_MD5
