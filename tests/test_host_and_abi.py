"""Host-side logic and the C-ABI library, without a GPU: constants, exported symbols, loud failure
when no device is present, the product never touching the oracle, output formats."""
import ctypes
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = os.path.join(ROOT, "tests", "golden")


def test_constants_match_the_reference_scripts(pore_cfg, temp_cfg, cube_cfg):
    assert pore_cfg.num_molecules == temp_cfg.num_molecules == 557649 and cube_cfg.num_molecules == 24627
    assert pore_cfg.collision_range == 3.385137501286538e-10
    assert (pore_cfg.open_air_particles, pore_cfg.hot_pore_particles, pore_cfg.gap_particles,
            pore_cfg.cold_pore_particles, pore_cfg.remaining_particles) == (174079, 2088, 2683, 204717, 3)
    assert pore_cfg.geom.z_gb == 1.3000000000000003e-07          # open_air_height + hot_coating_height
    assert abs(pore_cfg.dt - 1.84808e-13) < 1e-18 and abs(temp_cfg.dt - 1.84895e-13) < 1e-18
    assert abs(cube_cfg.dt - 7.39234e-12) < 1e-17
    assert pore_cfg.grid.nc == (14, 14, 148) and pore_cfg.grid.c0 == (-7, -7, 0)
    assert cube_cfg.grid.nc == (15, 15, 15) and cube_cfg.min_num_particles_per_cell == 7 and cube_cfg.remaining_particles == 1002
    assert float(temp_cfg.surface_energy_cold) == 1.7463480823716586e-21
    assert float(temp_cfg.surface_energy_hot) == 3.2454458503944502e-21


def test_grid_tables_use_the_reference_expressions(pore_cfg):
    g, c = pore_cfg.grid, pore_cfg
    for grp in range(2):
        for layer in range(7):
            k = 2 * layer + grp
            assert g.lo[0][k] == (2 * layer + grp - c.num_x_subdivions) * c.dx - c.collision_range
            assert g.edge[0][k + 1] == (2 * layer + grp - c.num_x_subdivions + 1) * c.dx
        for layer in range(74):
            k = 2 * layer + grp
            assert g.lo[2][k] == (2 * layer + grp) * c.dz - c.collision_range
            assert g.edge[2][k + 1] == (2 * layer + grp + 1) * c.dz


def test_scaled_configs_keep_density_and_even_grids():
    from argon_monte_carlo_b200 import config
    base = config.pore_config(True)
    for scale in (0.5, 2.82, 5.64):
        c = config.pore_config(True, scale=scale)
        assert all(n % 2 == 0 for n in c.grid.nc)
        assert abs(c.num_molecules / base.num_molecules / scale**3 - 1) < 1e-3
        assert 18e-9 < c.dx < 26e-9 and 18e-9 < c.dz < 26e-9


def test_overlap_threshold_is_the_exact_sqrt_boundary(pore_cfg):
    import math
    from argon_monte_carlo_b200 import amc
    cr = float(pore_cfg.collision_range)
    t = amc.overlap_threshold(cr)
    assert math.sqrt(t) >= cr and math.sqrt(math.nextafter(t, 0.0)) < cr
    rng = np.random.default_rng(0)
    d2 = t * (1 + rng.uniform(-4, 4, 20000) * 2.220446049250313e-16)
    assert np.array_equal(np.sqrt(d2) < cr, d2 < t)


def test_library_exports_every_symbol_of_the_header():
    from argon_monte_carlo_b200 import amc, build
    header = open(os.path.join(ROOT, "include", "amc.h")).read()
    declared = sorted(set(re.findall(r"\b(amc_[a-z_0-9]+)\s*\(", header)))
    assert declared == sorted(amc.EXPORTS)
    lib = ctypes.CDLL(build.build_library())
    for name in declared:
        assert hasattr(lib, name), name
    assert lib.amc_abi_version() == amc.ABI_VERSION


def test_ctypes_structs_match_the_header_layout():
    from argon_monte_carlo_b200 import amc
    header = open(os.path.join(ROOT, "include", "amc.h")).read()
    geom = re.search(r"typedef struct amc_geom \{(.*?)\} amc_geom;", header, re.S).group(1)
    names = re.findall(r"\b([A-Za-z_0-9]+)\s*[,;]", re.sub(r"/\*.*?\*/", "", geom, flags=re.S))
    assert tuple(n for n in names if n != "double") == amc.GEOM_FIELDS
    assert ctypes.sizeof(amc.AmcGeom) == 8 * len(amc.GEOM_FIELDS)
    assert ctypes.sizeof(amc.AmcStepStats) == 8 * (10 + 10 + 3)


def test_no_cpu_fallback_without_a_device(pore_cfg):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    from argon_monte_carlo_b200 import amc
    with pytest.raises(amc.AmcError):
        amc.Simulation(pore_cfg)


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "argon_monte_carlo_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(dirpath, f)).read()
                assert "oracle" not in text.lower().replace("the cpu oracle", "").replace("oracle)", "") or \
                    "import oracle" not in text and "from oracle" not in text, f
                assert "from oracle" not in text and "import oracle" not in text, f
    for f in os.listdir(os.path.join(ROOT, "drivers")):
        text = open(os.path.join(ROOT, "drivers", f)).read()
        assert "from oracle" not in text and "import oracle" not in text, f


def test_host_rng_consumes_the_streams_like_the_reference():
    """random_inbounds_direction (Temperature_Pore_MC.py:132-141) draws np.random.uniform,
    random.uniform, np.random.choice per attempt; check the stream positions and the cone."""
    import random
    from argon_monte_carlo_b200 import host_rng
    np.random.seed(3); random.seed(3)
    norm = np.array([0.0, 0.0, -1.0])
    d = host_rng.inbound_direction(norm)
    assert abs(np.dot(d, d) - 1) < 1e-15 and np.dot(d, norm) >= host_rng.COS85
    after_np, after_py = np.random.uniform(), random.random()
    # replay by hand
    np.random.seed(3); random.seed(3)
    import math
    while True:
        ct = np.random.uniform(low=-1.0, high=1.0); phi = random.uniform(0, math.pi); th = math.acos(ct)
        v = np.array([math.cos(phi) * math.sin(th), math.sin(phi) * math.sin(th) * np.random.choice([-1, 1]), math.cos(th)])
        if abs(np.dot(v, norm)) < host_rng.COS85:
            continue
        if np.dot(v, norm) < host_rng.COS85:
            v = -v
        break
    assert np.array_equal(v, d) and np.random.uniform() == after_np and random.random() == after_py


def test_output_writers_formats(tmp_path):
    from argon_monte_carlo_b200 import outputs
    counts = np.zeros((4, 200), dtype=np.int64)
    counts[:, 0], counts[:, 3] = 7, 1
    outputs.write_histograms(counts, str(tmp_path))
    x = open(os.path.join(str(tmp_path), "hist_x_axis_z_data.txt")).read()
    assert x == open(os.path.join(GOLD, "shipped_hist_x_axis_z_data.txt")).read()
    y = open(os.path.join(str(tmp_path), "hist_y_axis_total_data.txt")).read()
    assert y.startswith("[1.75e+08 0.00e+00 0.00e+00 2.50e+07") and y.endswith("]")
    ref = np.histogram([1e-9] * 7 + [1.6e-8], bins=200, range=(0, 10 ** -6), density=True)[0]
    assert np.array_equal(outputs.density(counts[0]), ref)
    outputs.write_momentum_energy_csv([-4.06031315107044e-22, 0], [-1.81678728388749e-18, 0], [-1.4427412763244e-19, 0],
                                      os.path.join(str(tmp_path), "m.csv"))
    lines = open(os.path.join(str(tmp_path), "m.csv")).read().split("\n")
    shipped = open(os.path.join(GOLD, "shipped_momentum_energy.csv")).read().split("\n")
    assert lines[0] == shipped[0] and lines[1] == shipped[1] and lines[2] == "1,0,0,0"


def test_shipped_histograms_are_consistent_with_the_density_rule():
    """The four shipped hist_y files all come from 30,828 completed paths."""
    from argon_monte_carlo_b200 import outputs
    for s in ("total", "x", "y", "z"):
        dens = np.array(open(os.path.join(GOLD, "shipped_hist_y_axis_%s_data.txt" % s)).read().strip("[]").split(), dtype=float)
        assert len(dens) == 200
        counts = np.rint(dens * 5e-9 * 30828).astype(np.int64)
        assert counts.sum() == 30828
        assert np.allclose(outputs.density(counts), dens, rtol=2e-8)
