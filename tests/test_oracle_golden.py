"""The CPU oracle against the reference's own outputs (committed golden vectors, made by
oracle/make_golden.py from the unmodified Python reference in the build container), and against
the result files the reference ships.  No GPU needed."""
import hashlib
import json
import os

import numpy as np
import pytest

GOLD = os.path.join(os.path.dirname(__file__), "golden")
KEYS = ("x", "y", "z", "vx", "vy", "vz", "dist", "dist_x", "dist_y", "dist_z", "flag")


def digest(a):
    a = np.ascontiguousarray(a)
    return hashlib.sha256(a.tobytes()).hexdigest()


def unhex(lst):
    return np.array([float.fromhex(v) for v in lst])


def load(name):
    with open(os.path.join(GOLD, name)) as f:
        return json.load(f)


# ---------------------------------------------------------------------------------------------- init
@pytest.mark.parametrize("kind", ["pore", "temp"])
def test_initial_state_matches_reference(kind, pore_cfg, temp_cfg, pore_init, temp_init):
    g = load("ref_init.json")[kind]
    init = pore_init if kind == "pore" else temp_init
    assert len(init[0]) == g["n"] == 557649
    for k, a in zip(("x", "y", "z", "vx", "vy", "vz"), init):
        assert digest(a) == g[k], k


def test_init_leaves_host_rng_where_the_reference_does(temp_cfg):
    import random
    from argon_monte_carlo_b200 import init_state
    init_state.pore_initial_state(temp_cfg)
    g = load("ref_init.json")["temp"]
    assert float(np.random.uniform()).hex() == g["next_np_uniform"]
    assert float(random.random()).hex() == g["next_py_random"]


# ---------------------------------------------------------------------------------------------- whole steps
@pytest.mark.parametrize("kind", ["pore", "temp"])
def test_three_reference_steps_bit_exact(kind, oracle, pore_cfg, temp_cfg):
    """Full 557,649-particle state after each of the first three timesteps of the unmodified
    scripts: SHA-256 of every array, collision counters, completed paths, per-step series."""
    from argon_monte_carlo_b200 import init_state
    from oracle import steps
    gold = load("ref_%s_3steps.json" % kind)
    cfg = pore_cfg if kind == "pore" else temp_cfg
    oracle.set_ref_mode(True)
    try:
        st = oracle.ParticleState(*init_state.pore_initial_state(cfg))
        sink = oracle.PathSink()
        series = []
        for k in range(3):
            r = steps.pore_step(st, cfg, sink) if kind == "pore" else steps.temp_step_host_rng(st, cfg, sink)
            rec = gold["steps"][k]["end"]
            assert r["collisions"] == rec["ncol"]
            for key in KEYS:
                assert digest(getattr(st, key)) == rec[key], "step %d %s" % (k, key)
            series.append(r)
        fin = gold["final"]
        for got, key in zip(sink.arrays(), ("completed", "completed_x", "completed_y", "completed_z")):
            assert np.array_equal(np.sort(got), unhex(fin[key]))
        if kind == "temp":
            assert [float(r["dpz"]).hex() for r in series] == fin["momentum"]
            assert [float(r["e_cold"]).hex() for r in series] == fin["e_cold"]
            assert [float(r["e_hot"]).hex() for r in series] == fin["e_hot"]
    finally:
        oracle.set_ref_mode(False)


def test_shipped_momentum_energy_csv_rows(oracle, temp_cfg):
    """Rows 0-1 of the momentum_energy.csv the reference ships, digit for digit (the file holds
    str(mpf): 15 significant digits); row 2 to 1e-13 (the author's NumPy/libm differ in the last ulp)."""
    import mpmath
    from argon_monte_carlo_b200 import init_state
    from oracle import steps
    rows = open(os.path.join(GOLD, "shipped_momentum_energy.csv")).read().split("\n")
    assert rows[0] == ",Momentum,EnergyCold,EnergyHot" and len(rows) >= 251
    oracle.set_ref_mode(True)
    try:
        st = oracle.ParticleState(*init_state.pore_initial_state(temp_cfg))
        for k in range(3):
            r = steps.temp_step_host_rng(st, temp_cfg)
            line = "%d,%s,%s,%s" % (k, mpmath.mpf(r["dpz"]), mpmath.mpf(r["e_cold"]), mpmath.mpf(r["e_hot"]))
            if k < 2:
                assert line == rows[1 + k]
            else:
                ref = [float(v) for v in rows[1 + k].split(",")[1:]]
                got = [float(r["dpz"]), float(r["e_cold"]), float(r["e_hot"])]
                assert np.allclose(got, ref, rtol=1e-13, atol=0)
    finally:
        oracle.set_ref_mode(False)


def test_plain_and_reference_arithmetic_agree(oracle, pore_cfg, pore_init):
    """The arithmetic the CUDA path implements (v*v, unfused dot) against NumPy-scalar arithmetic
    (libm pow, FMA-chain dot): identical collision-pair sets and wall flags, state within 1e-12."""
    from oracle import steps
    out = []
    for mode in (True, False):
        oracle.set_ref_mode(mode)
        st = oracle.ParticleState(*pore_init)
        pairs = oracle.PairSink()
        bits = []
        for k in range(2):
            bits.append(steps.pore_step(st, pore_cfg, None, pairs, want_bits=True)["hit_bits"])
        out.append((st, sorted(zip(*[a.tolist() for a in pairs.arrays()])), bits))
    oracle.set_ref_mode(False)
    (a, pa, ba), (b, pb, bb) = out
    assert pa == pb
    assert all(np.array_equal(x, y) for x, y in zip(ba, bb))
    for key in KEYS[:-1]:
        x, y = getattr(a, key), getattr(b, key)
        assert np.allclose(x, y, rtol=1e-11, atol=0), key


# ---------------------------------------------------------------------------------------------- operators
def test_pairwise_particles_in_cell_fixtures(oracle):
    """pairwise_particles_in_cell (Open_Air_Pore_MC.py:160-255) on dense random cells with chained
    collisions: positions, velocities, paths, flags and completed paths bit for bit."""
    from argon_monte_carlo_b200 import config
    cfg = config.pore_config(False)
    oracle.set_ref_mode(True)
    try:
        for rec in load("ref_operators.json")["pairwise"]:
            n = rec["n"]
            pos, vel, paths = unhex(rec["pos"]).reshape(3, n), unhex(rec["vel"]).reshape(3, n), unhex(rec["paths"]).reshape(4, n)
            st = oracle.ParticleState(pos[0], pos[1], pos[2], vel[0], vel[1], vel[2], paths[0], paths[1], paths[2], paths[3],
                                      np.array(rec["flag"], dtype=np.uint8))
            grid = config.Grid(nc=(1, 1, 1), c0=(0, 0, 0), d=(1.0, 1.0, 1.0), band=(1.0, 1.0, 1.0))   # one cell holding everything
            sink = oracle.PathSink()
            ncol, checks, errs = oracle.cube_pp_sweep(st, grid, cfg.collision_range, cfg.argon_mass, sink)
            assert ncol == rec["ncol"] and checks == n * (n - 1) // 2 and errs == 0
            for got, exp in zip((st.x, st.y, st.z), rec["out_pos"]):
                assert np.array_equal(got, unhex(exp))
            for got, exp in zip((st.vx, st.vy, st.vz), rec["out_vel"]):
                assert np.array_equal(got, unhex(exp))
            for got, exp in zip((st.dist, st.dist_x, st.dist_y, st.dist_z), rec["out_paths"]):
                assert np.array_equal(got, unhex(exp))
            assert np.array_equal(st.flag.astype(bool), np.array(rec["out_flag"], dtype=bool))
            for got, exp in zip(sink.arrays(), rec["completed"]):
                assert np.array_equal(got, unhex(exp))       # same order too: one cell, ascending index
    finally:
        oracle.set_ref_mode(False)


def _state_from_snapshot(oracle, snap, n):
    pos, vel, paths = unhex(snap["pos"]).reshape(3, n), unhex(snap["vel"]).reshape(3, n), unhex(snap["paths"]).reshape(4, n)
    st = oracle.ParticleState(pos[0], pos[1], pos[2], vel[0], vel[1], vel[2], paths[0], paths[1], paths[2], paths[3],
                              np.array(snap["flag"], dtype=np.uint8))
    st.px[:], st.py[:], st.pz[:] = pos[0], pos[1], pos[2]
    return st


def _assert_snapshot(st, snap, n):
    pos, vel, paths = unhex(snap["pos"]).reshape(3, n), unhex(snap["vel"]).reshape(3, n), unhex(snap["paths"]).reshape(4, n)
    for got, exp in zip((st.x, st.y, st.z, st.vx, st.vy, st.vz, st.dist, st.dist_x, st.dist_y, st.dist_z),
                        list(pos) + list(vel) + list(paths)):
        assert np.array_equal(got, exp)
    assert np.array_equal(st.flag.astype(bool), np.array(snap["flag"], dtype=bool))


def test_pore_wall_operator_fixtures(oracle, pore_cfg):
    """hit_cylinder_side_wall / hit_vertical_wall (Open_Air_Pore_MC.py:257-348): the fixture states
    are built so that exactly case 1 (side) or case 2a (plane z=0) fires for every particle."""
    oracle.set_ref_mode(True)
    try:
        for rec in load("ref_operators.json")["pore_walls"]:
            n = rec["n"]
            st = _state_from_snapshot(oracle, rec["before"], n)
            sink = oracle.PathSink()
            counts, errs, _ = oracle.pore_walls(st, pore_cfg.geom, sink)
            assert errs == 0 and counts.sum() == n and counts[0 if rec["op"] == "side" else 1] == n
            _assert_snapshot(st, rec["after"], n)
            for got, exp in zip(sink.arrays(), rec["completed"]):
                assert np.array_equal(got, unhex(exp))
    finally:
        oracle.set_ref_mode(False)


def test_temp_energized_operator_fixtures(oracle, temp_cfg):
    """hit_vertical_coated_wall / hit_cylinder_coated_side_wall / hit_cylinder_gap_side_wall
    (Temperature_Pore_MC.py:349-553) with random_inbounds_direction pinned to a deterministic
    function of the normal; momentum / energy sums bit for bit; surface_energy_gap by mpmath."""
    from argon_monte_carlo_b200 import config

    def fixed_direction(norm):
        d = np.array([0.3, -0.2, 0.0]) + 0.9 * np.asarray(norm, dtype=float)
        return d / np.sqrt(np.dot(d, d))
    import copy
    oracle.set_ref_mode(True)
    try:
        for rec in load("ref_operators.json")["temp_walls"]:
            n = rec["n"]
            st = _state_from_snapshot(oracle, rec["before"], n)
            g = copy.copy(temp_cfg.geom)
            idx = np.arange(n, dtype=np.int64)
            sink = oracle.PathSink()
            if rec["op"] == "plane":        # surface_energy_cold, plane z = 0, inbound +z  == case 3-cold arithmetic
                g.zc3 = 0.0
                case, surf = 3, None
                normals = np.tile([0.0, 0.0, 1.0], (n, 1))
            elif rec["op"] == "side":       # surface_energy_hot on the open-air collision radius == case 6-hot arithmetic
                g.R_p_c = g.R_oa_c
                case, surf = 8, None
            else:                           # gap wall, case 4
                case = 5
            if rec["op"] != "plane":
                # detect() evaluates the case mask; the fixtures call the operator with an all-true mask,
                # so compute the contact normals the same way but for every particle
                R = g.R_p_c if case == 8 else g.R_g_c
                normals = np.zeros((n, 3))
                colz = np.zeros(n)
                for k in range(n):
                    x, y, z, vx, vy, vz = st.x[k], st.y[k], st.z[k], st.vx[k], st.vy[k], st.vz[k]
                    a = (-vx)**2 + (-vy)**2
                    b = 2 * (x * (-vx) + y * (-vy))
                    c = x**2 + y**2 - R**2
                    t = np.min([(-b + np.sqrt(b**2 - 4 * a * c)) / (2 * a), (-b - np.sqrt(b**2 - 4 * a * c)) / (2 * a)])
                    normals[k] = -(np.array([x - vx * t, y - vy * t, 0]) / R)
                    colz[k] = z - vz * t
                surf = np.array([float(config.surface_energy_gap(temp_cfg, cz)) for cz in colz]) if case == 5 else None
            dirs = np.array([fixed_direction(nv) for nv in normals])
            dpz, de, errs = oracle.temp_case_apply(st, g, case, idx, dirs, surf, sink)
            assert errs == 0
            _assert_snapshot(st, rec["after"], n)
            assert float(dpz).hex() == rec["dpz"]
            if "de" in rec:
                assert float(de).hex() == rec["de"]
            for got, exp in zip(sink.arrays(), rec["completed"]):
                assert np.array_equal(got, unhex(exp))
            if rec["op"] == "gap":
                for zh, eh in rec["surface_energy_gap"].items():
                    assert float(config.surface_energy_gap(temp_cfg, float.fromhex(zh))).hex() == eh
    finally:
        oracle.set_ref_mode(False)


def test_recapture_fixtures(oracle, pore_cfg, temp_cfg):
    for rec in load("ref_operators.json")["recapture"]:
        n = rec["n"]
        pos = unhex(rec["before"]).reshape(3, n)
        st = oracle.ParticleState(pos[0], pos[1], pos[2], np.zeros(n), np.zeros(n), np.zeros(n))
        if rec["kind"] == "pore":
            assert oracle.pore_recapture(st, pore_cfg.geom) == rec["count"]
        else:
            assert oracle.temp_oob_count(st, temp_cfg.geom) == rec["report"]
            assert oracle.temp_recapture(st, temp_cfg.geom) == rec["count"]
        after = unhex(rec["after"]).reshape(3, n)
        assert np.array_equal(st.x, after[0]) and np.array_equal(st.y, after[1]) and np.array_equal(st.z, after[2])


# ---------------------------------------------------------------------------------------------- cube, end to end
def test_cube_full_run_reproduces_reference_outputs(oracle, cube_cfg, cube_init, tmp_path):
    """The whole 500-step Open_Air_Cube_MC.py run: collisions per step, completed paths, the printed
    mean free paths (all 17 digits) and the MD5 of the eight result files, as produced by the
    unmodified script in the build container."""
    from argon_monte_carlo_b200 import outputs
    from oracle import steps
    oracle.set_ref_mode(True)
    try:
        st = oracle.ParticleState(*cube_init)
        sink = oracle.PathSink()
        ncol = [steps.cube_step(st, cube_cfg, sink)["pp_collisions"] for _ in range(500)]
    finally:
        oracle.set_ref_mode(False)
    assert ncol[:10] == [31, 41, 36, 50, 44, 45, 47, 49, 60, 53] and sum(ncol) == 24382
    total, cx, cy, cz = sink.arrays()
    assert len(total) == 27448
    assert str(np.average(total)) == "3.5298513644857337e-07"
    assert str(np.average(cx)) == "1.7567313741196367e-07"
    assert str(np.average(cy)) == "1.772052223449425e-07"
    assert str(np.average(cz)) == "1.7713843307953472e-07"
    counts = [oracle.histogram(a) for a in (total, cx, cy, cz)]
    for c, a in zip(counts, (total, cx, cy, cz)):
        assert np.array_equal(c, np.histogram(a, bins=200, range=(0, 10 ** -6))[0])
    outputs.write_histograms(counts, str(tmp_path))
    md5 = lambda f: hashlib.md5(open(os.path.join(str(tmp_path), f), "rb").read()).hexdigest()
    assert md5("hist_y_axis_total_data.txt") == "8f339ccafbc1c7cb736bbca21ebd4900"
    assert md5("hist_y_axis_x_data.txt") == "4257653f7b7ec37e0ab2cdb80d04d7f4"
    assert md5("hist_y_axis_y_data.txt") == "cd290f91715f6deee6c438a8b7086c2a"
    assert md5("hist_y_axis_z_data.txt") == "7f8d7ec75901a75481005e440a0cb818"
    for s in ("total", "x", "y", "z"):      # the x-axis files are byte-identical to the shipped ones
        assert open(os.path.join(str(tmp_path), "hist_x_axis_%s_data.txt" % s)).read() == \
            open(os.path.join(GOLD, "shipped_hist_x_axis_%s_data.txt" % s)).read()


# ---------------------------------------------------------------------------------------------- small pieces
def test_histogram_rule_matches_numpy(oracle):
    rng = np.random.default_rng(5)
    edges = np.linspace(0, 10 ** -6, 201)
    v = np.concatenate([rng.exponential(8e-8, 20000), edges, np.nextafter(edges, 0), np.nextafter(edges, 1),
                        [0.0, 1e-6, 1.0000001e-6, -1e-12, np.nan]])
    assert np.array_equal(oracle.histogram(v), np.histogram(v[~np.isnan(v)], bins=200, range=(0, 10 ** -6))[0])
    assert np.array_equal(oracle.histogram(np.zeros(0)), np.zeros(200, dtype=np.int64))


def test_philox_known_answers(oracle):
    """Philox4x32-10 known-answer vectors (Random123 kat_vectors)."""
    import ctypes as C
    L = oracle.lib()

    def ph(ctr, key):
        out = (C.c_uint32 * 4)()
        L.orc_philox4x32_10((C.c_uint32 * 4)(*ctr), (C.c_uint32 * 2)(*key), out)
        return [int(v) for v in out]
    assert ph([0, 0, 0, 0], [0, 0]) == [0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8]
    assert ph([0xffffffff] * 4, [0xffffffff] * 2) == [0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd]
    assert ph([0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344], [0xa4093822, 0x299f31d0]) == \
        [0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1]


def test_philox_directions_are_isotropic_in_the_cone(oracle, temp_cfg):
    g = temp_cfg.geom
    norm = np.array([0.0, 0.0, 1.0])
    d = np.array([oracle.philox_direction(17, i, 3, 3, norm, g.cos85) for i in range(4000)])
    assert np.allclose(np.linalg.norm(d, axis=1), 1.0, atol=1e-14)
    assert (d @ norm >= g.cos85).all()
    # isotropic within the accepted hemisphere-cap: cos(theta) uniform on [cos85, 1]
    from scipy import stats
    u = (d @ norm - g.cos85) / (1 - g.cos85)
    assert stats.kstest(u, "uniform").pvalue > 1e-3
    assert abs(d[:, 0].mean()) < 0.05 and abs(d[:, 1].mean()) < 0.05


def test_gap_energy_chebyshev_matches_mpmath(oracle, temp_cfg):
    from argon_monte_carlo_b200 import config
    cheb = config.gap_energy_chebyshev(temp_cfg, 16)
    for z in np.linspace(temp_cfg.gap_bottom_height, temp_cfg.gap_top_height, 11):
        exact = float(config.surface_energy_gap(temp_cfg, z))
        assert abs(oracle.cheb_eval(cheb, z) - exact) <= 1e-13 * exact
    assert abs(float(config.surface_energy_gap(temp_cfg, temp_cfg.gap_bottom_height)) - 4.5806845213987e-20) < 1e-33
