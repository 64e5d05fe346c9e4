"""Phase-level entry points against the oracle one phase at a time, and the edge cases of the C ABI
(empty and tiny inputs, particles outside every cell, capacity errors, bad configs)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu
KEYS = ("x", "y", "z", "vx", "vy", "vz", "dist", "dist_x", "dist_y", "dist_z")


def same(got, st):
    for k in KEYS:
        assert np.array_equal(got[k], getattr(st, k)), k
    assert np.array_equal(got["flag"].astype(bool), st.flag.astype(bool))


def test_pore_phase_by_phase(oracle, pore_cfg, pore_init):
    """amc_drift / amc_walls / amc_recapture / amc_pairs == the oracle's phases (Pore:427-437, 442-485,
    354-375, 522-549), state compared after every phase, for three consecutive steps."""
    from argon_monte_carlo_b200 import amc
    st = oracle.ParticleState(*pore_init)
    sim = amc.Simulation(pore_cfg, taps=amc.TAP_WALL_BITS)
    sim.set_state(*pore_init)
    for step in range(3):
        oracle.drift(st, pore_cfg.dt, True)
        sim.drift()
        same(sim.get_state(), st)
        counts, errs, bits = oracle.pore_walls(st, pore_cfg.geom, None, True)
        g = sim.walls()
        assert np.array_equal(g["wall_hits"][:9], counts) and g["errors"] == errs
        assert np.array_equal(sim.wall_bits(), bits)
        same(sim.get_state(), st)
        assert sim.recapture() == oracle.pore_recapture(st, pore_cfg.geom)
        same(sim.get_state(), st)
        ncol, checks, perr = oracle.pp_groups(st, pore_cfg.grid, pore_cfg.collision_range, pore_cfg.argon_mass)
        g = sim.pairs()
        assert g["pp_collisions"] == ncol and g["pair_checks_ref"] == checks
        assert g["pair_checks_exec"] < checks            # hashed-bin detection + filtered visits of the flagged cells only
        same(sim.get_state(), st)
        assert sim.recapture() == oracle.pore_recapture(st, pore_cfg.geom)
        same(sim.get_state(), st)
    sim.close()


def test_empty_and_tiny_states(pore_cfg):
    from argon_monte_carlo_b200 import amc
    sim = amc.Simulation(pore_cfg, max_particles=16)
    z = np.zeros(0)
    sim.set_state(z, z, z, z, z, z)
    st = sim.step(2)
    assert all(s["collisions"] == 0 for s in st) and len(sim.get_state()["x"]) == 0
    # one particle, then two overlapping particles in the bottom end cap
    sim.set_state([1e-8], [0.0], [5e-8], [100.0], [0.0], [0.0])
    assert sim.step(1)[0]["collisions"] == 0
    cr = pore_cfg.collision_range
    sim.set_state([1e-8, 1e-8 + 0.5 * cr], [0.0, 0.0], [5e-8, 5e-8], [100.0, -100.0], [0.0, 0.0], [0.0, 0.0])
    s = sim.step(1)[0]
    assert s["pp_collisions"] == 1
    g = sim.get_state()
    assert abs(g["vx"][0] + 100.0) < 1e-9 and abs(g["vx"][1] - 100.0) < 1e-9   # head-on equal masses exchange velocities
    assert g["flag"].tolist() == [1, 1]
    sim.close()


def test_particles_outside_every_cell_are_skipped_and_recaptured(oracle, pore_cfg):
    from argon_monte_carlo_b200 import amc
    from oracle import steps
    g = pore_cfg.geom
    x = np.array([0.0, 2 * g.R_oa, 0.0, 1e-9, np.nextafter(pore_cfg.grid.edge[0][-1], 1.0)])
    y = np.array([0.0, 0.0, 0.0, 1e-9, 0.0])
    z = np.array([-5e-9, 5e-8, g.H + 3e-9, 5e-8, 5e-8])
    v = np.zeros(5)
    st = oracle.ParticleState(x, y, z, v + 1.0, v, v + 1.0)
    sim = amc.Simulation(pore_cfg, max_particles=8)
    sim.set_state(x, y, z, v + 1.0, v, v + 1.0)
    for _ in range(2):
        r = steps.pore_step(st, pore_cfg)
        s = sim.step(1)[0]
        assert s["oob_after_walls"] == r["oob_after_walls"] and s["collisions"] == r["collisions"]
        same(sim.get_state(), st)
    sim.close()


def test_cell_capacity_overflow_is_reported_not_ignored(pore_cfg):
    from argon_monte_carlo_b200 import amc
    rng = np.random.default_rng(0)
    n = 700                                   # > AMC_MAX_MEMBERS (512) particles in one reference cell
    x = 1e-8 + rng.uniform(0, 5e-9, n); y = 1e-8 + rng.uniform(0, 5e-9, n); z = 5e-8 + rng.uniform(0, 5e-9, n)
    sim = amc.Simulation(pore_cfg, max_particles=n)
    sim.set_state(x, y, z, np.zeros(n), np.zeros(n), np.zeros(n))
    with pytest.raises(amc.AmcError, match="AMC_MAX_MEMBERS"):
        sim.step(1)
    sim.close()


def test_bad_arguments_fail_loudly(pore_cfg, cube_cfg):
    from argon_monte_carlo_b200 import amc, config
    sim = amc.Simulation(pore_cfg, max_particles=4)
    with pytest.raises(amc.AmcError, match="max_particles"):
        sim.set_state(np.zeros(5), np.zeros(5), np.zeros(5), np.zeros(5), np.zeros(5), np.zeros(5))
    with pytest.raises(amc.AmcError):
        sim.pair_list()                        # tap not enabled
    with pytest.raises(amc.AmcError):
        sim.wall_case(0)                       # not a Temp handle
    sim.close()
    odd = config.Grid(nc=(15, 14, 148), c0=(-7, -7, 0), d=(1e-8, 1e-8, 1e-8), band=(1e-10,) * 3)
    with pytest.raises(amc.AmcError, match="even"):
        amc.Simulation(pore_cfg, grid=odd)
    host = amc.Simulation(config.pore_config(True), rng_mode=amc.RNG_HOST, max_particles=4)
    with pytest.raises(amc.AmcError, match="phase"):
        host.step(1)                           # host-RNG handles are stepped phase by phase
    host.close()


def test_cube_geometry_with_colour_groups(oracle):
    """BASELINE config 4: cube walls + the 8-colour-group pair schedule (the serial sweep is not
    shardable); oracle = cube walls + pp_groups."""
    from argon_monte_carlo_b200 import amc, config, init_state
    cfg = config.cube_config(scale=2.0, n_sub=10)                      # 200 nm cube, 20 nm cells
    cfg.grid = config.Grid(nc=(10, 10, 10), c0=(0, 0, 0), d=(cfg.dx, cfg.dy, cfg.dz), band=(cfg.collision_range,) * 3)
    n = 200000
    init = init_state.synthetic_cube_state(cfg, n, seed=9)
    st = oracle.ParticleState(*init)
    sim = amc.Simulation(cfg, kind=amc.KIND_CUBE, pp_mode=amc.PP_GROUPS, max_particles=n)
    sim.set_state(*init)
    for _ in range(4):
        oracle.drift(st, cfg.dt, False)
        oracle.cube_walls(st, cfg.cube_x, cfg.cube_y, cfg.cube_z)
        ncol, _, _ = oracle.pp_groups(st, cfg.grid, cfg.collision_range, cfg.argon_mass)
        s = sim.step(1)[0]
        assert s["pp_collisions"] == ncol
        same(sim.get_state(), st)
    sim.close()


def _cube_groups_run(oracle, scale, n_sub, n, seed, steps):
    from argon_monte_carlo_b200 import amc, config, init_state
    cfg = config.cube_config(scale=scale, n_sub=n_sub)
    cfg.grid = config.Grid(nc=(n_sub,) * 3, c0=(0, 0, 0), d=(cfg.dx, cfg.dy, cfg.dz), band=(cfg.collision_range,) * 3)
    init = init_state.synthetic_cube_state(cfg, n, seed=seed)
    st = oracle.ParticleState(*init)
    sim = amc.Simulation(cfg, kind=amc.KIND_CUBE, pp_mode=amc.PP_GROUPS, max_particles=n)
    sim.set_state(*init)
    total = 0
    for _ in range(steps):
        oracle.drift(st, cfg.dt, False)
        oracle.cube_walls(st, cfg.cube_x, cfg.cube_y, cfg.cube_z)
        ncol, checks, _ = oracle.pp_groups(st, cfg.grid, cfg.collision_range, cfg.argon_mass)
        s = sim.step(1)[0]
        assert s["pp_collisions"] == ncol and s["pair_checks_ref"] == checks
        same(sim.get_state(), st)
        total += ncol
    sim.close()
    return total


def test_cells_with_more_candidates_than_the_detection_pass_holds(oracle):
    """25 nm cells with ~400 members: more candidates than a CTA of k_detect takes (384), so every cell is flagged
    unseen and the ordered kernel searches all of them itself -- the conservative fallback, end to end."""
    assert _cube_groups_run(oracle, 2.0, 8, 204800, 11, 3) > 100


def test_tiny_cells_dense_gas_many_chained_collisions(oracle):
    """1 nm cells (3 collision ranges: 2-3 search bins per axis) and 25 x the gas density: overlapping pairs
    everywhere, several collisions per cell visit, moved particles that change cell all the time.  The ordered
    resolution, the escaped lists and the activation of the cells a particle enters or leaves have to reproduce the
    reference sweep exactly."""
    assert _cube_groups_run(oracle, 0.2, 20, 5000, 12, 6) > 300


def test_checkpoint_resume_is_exact(temp_cfg, temp_init, tmp_path):
    """A run continued from a checkpoint in a fresh handle equals the uninterrupted run: state,
    histograms, free-path sums and the device-RNG stream (keyed by the step index)."""
    from argon_monte_carlo_b200 import amc
    a = amc.Simulation(temp_cfg, seed=9)
    a.set_state(*temp_init)
    a.step(8)
    ref_state, ref_hist = a.get_state(), a.histograms()
    a.close()
    b = amc.Simulation(temp_cfg, seed=9)
    b.set_state(*temp_init)
    b.step(5)
    b.checkpoint(str(tmp_path / "ck.npz"))
    b.close()
    c = amc.Simulation(temp_cfg, seed=9)
    c.restore(str(tmp_path / "ck.npz"))
    c.step(3)
    got_state, got_hist = c.get_state(), c.histograms()
    c.close()
    for k in KEYS + ("flag",):
        assert np.array_equal(got_state[k], ref_state[k]), k
    assert np.array_equal(got_hist[0], ref_hist[0]) and got_hist[1] == ref_hist[1] and np.array_equal(got_hist[2], ref_hist[2])
