"""Slab decomposition: the N-rank run must be bit-identical to the single-domain run.
All ranks live on one GPU here (LocalTransport: device copies instead of NCCL), which exercises every
device-side piece of the protocol -- migration, ghost copies, per-group hand-over of particles moved
by a collision at a cut -- deterministically; the NCCL transport itself is covered by
tests/test_slab_transport.py (gloo, CPU) and by bench.py on several GPUs."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

KEYS = ("x", "y", "z", "vx", "vy", "vz", "dist", "dist_x", "dist_y", "dist_z")


def run_single(cfg, init, steps, **kw):
    from argon_monte_carlo_b200 import amc
    sim = amc.Simulation(cfg, taps=amc.TAP_PAIRS, **kw)
    sim.set_state(*init)
    stats = sim.step(steps)
    out = sim.get_state(), stats, sim.pair_list(), sim.histograms()
    sim.close()
    return out


def run_slabs(cfg, init, steps, nranks, cuts=None, **kw):
    from argon_monte_carlo_b200 import amc, slab
    sim = slab.SlabSimulation(cfg, nranks, init[2], taps=amc.TAP_PAIRS, cuts=cuts, **kw)
    sim.debug_counts = True
    sim.set_state(*init)
    stats = sim.step(steps)
    out = sim.get_state(), stats, sim.pair_list(), sim.histograms(), sim.cuts
    run_slabs.last_exchanged = dict(sim.exchanged)
    sim.close()
    return out


def compare(single, slabs, n):
    (st1, stats1, pairs1, hist1), (st2, stats2, pairs2, hist2, cuts) = single, slabs
    assert st2["_owned_total"] == n, "particles lost or duplicated across ranks (cuts %s)" % cuts
    for a, b in zip(stats1, stats2):
        for k in ("wall_collisions", "pp_collisions", "pair_checks_ref", "oob_after_walls", "oob_after_pp", "errors",
                  "completed_paths"):
            assert a[k] == b[k], (k, a[k], b[k])
        assert np.array_equal(a["wall_hits"], b["wall_hits"])
        for k in ("dpz", "e_cold", "e_hot"):
            assert abs(a[k] - b[k]) <= 1e-14 * abs(a[k])
    for k in KEYS:
        bad = np.nonzero(st1[k] != st2[k])[0]
        assert len(bad) == 0, "%s differs for %d particles (first id %d), cuts %s" % (k, len(bad), bad[0], cuts)
    assert np.array_equal(st1["flag"], st2["flag"])
    key = lambda p: sorted(zip(p[0].tolist(), p[1].tolist(), p[2].tolist()))
    assert key(pairs1) == key(pairs2)
    assert np.array_equal(hist1[0], hist2[0]) and hist1[1] == hist2[1]


@pytest.mark.parametrize("nranks,cuts", [(2, None), (3, None), (4, [0, 2, 5, 144, 148])])
def test_pore_slabs_bit_identical(pore_cfg, pore_init, nranks, cuts):
    single = run_single(pore_cfg, pore_init, 4)
    slabs = run_slabs(pore_cfg, pore_init, 4, nranks, cuts)
    compare(single, slabs, len(pore_init[0]))


def test_temp_slabs_bit_identical_device_rng(temp_cfg, temp_init):
    single = run_single(temp_cfg, temp_init, 4, seed=17)
    slabs = run_slabs(temp_cfg, temp_init, 4, 3, seed=17)
    compare(single, slabs, len(temp_init[0]))


def test_dense_cut_many_boundary_collisions(oracle):
    """A small, dense synthetic pore (overlapping start) cut through its end caps: thousands of
    collisions per step, many of them at the cuts."""
    from argon_monte_carlo_b200 import config, init_state
    cfg = config.pore_config(False, scale=0.5)
    init = init_state.synthetic_pore_state(cfg, seed=11)
    single = run_single(cfg, init, 6)
    nz = cfg.grid.nc[2]   # the end caps are the z layers 0-2 and nz-3..nz-1: cut inside them, one-layer slabs included
    total = {"xfer": 0, "boundary": 0}
    for nranks, cuts in ((2, [0, 2, nz]), (3, [0, 1, nz - 2, nz]), (4, [0, 1, 2, nz - 1, nz])):
        compare(single, run_slabs(cfg, init, 6, nranks, cuts), len(init[0]))
        ex = run_slabs.last_exchanged
        print("exchanged records:", ex)
        assert ex["xfer"] > 500, ex
        total = {k: total[k] + ex[k] for k in total}
    assert total["boundary"] >= 10, total   # collisions at the cuts were actually handed over between ranks


def test_long_run_many_handovers():
    """30 steps of the dense small pore over 4 ranks with every cut inside an end cap: dozens of
    particles are moved by a collision while a neighbour holds a copy (or needs one afterwards)."""
    from argon_monte_carlo_b200 import config, init_state
    cfg = config.pore_config(False, scale=0.5)
    init = init_state.synthetic_pore_state(cfg, seed=23)
    nz = cfg.grid.nc[2]
    single = run_single(cfg, init, 30)
    compare(single, run_slabs(cfg, init, 30, 4, [0, 1, 2, nz - 1, nz]), len(init[0]))
    print("exchanged records:", run_slabs.last_exchanged)
    assert run_slabs.last_exchanged["boundary"] >= 10


def test_cube_geometry_slabs_bit_identical():
    """BASELINE config 4 at small scale: Maxwellian gas in a cube with specular walls, colour-group pair
    schedule, slab-decomposed along z -- identical to the single-domain run."""
    from argon_monte_carlo_b200 import amc, config, init_state, slab
    cfg = config.cube_config(scale=2.0, n_sub=10)
    grid = config.Grid(nc=(10, 10, 10), c0=(0, 0, 0), d=(cfg.dx, cfg.dy, cfg.dz), band=(cfg.collision_range,) * 3)
    n = 200000
    init = init_state.synthetic_cube_state(cfg, n, seed=4)
    one = amc.Simulation(cfg, kind=amc.KIND_CUBE, pp_mode=amc.PP_GROUPS, grid=grid, max_particles=n, taps=amc.TAP_PAIRS)
    one.set_state(*init)
    stats1 = one.step(5)
    single = (one.get_state(), stats1, one.pair_list(), one.histograms())
    one.close()
    sim = slab.SlabSimulation(cfg, 3, init[2], kind=amc.KIND_CUBE, grid=grid, taps=amc.TAP_PAIRS)
    sim.set_state(*init)
    stats2 = sim.step(5)
    slabs = (sim.get_state(), stats2, sim.pair_list(), sim.histograms(), sim.cuts)
    sim.close()
    compare(single, slabs, n)


def test_device_resident_stepping_single_rank(temp_cfg, temp_init):
    """amc_slab_step keeps the particle count on the device and never returns to the host inside a step; with one
    rank (no peers) it must reproduce the single-domain run exactly -- the multi-rank version of this check needs
    one GPU per rank (tests/test_gpu_nccl_slab.py)."""
    from argon_monte_carlo_b200 import amc, slab
    one = amc.Simulation(temp_cfg, seed=17)
    one.set_state(*temp_init)
    ref = one.step(6)
    ref_digest = one.state_digest()
    one.close()
    sim = slab.SlabSimulation(temp_cfg, 1, temp_init[2], seed=17, p2p=True)
    sim.set_state(*temp_init)
    got = sim.step_fused(3) + sim.step_fused(3)
    for a, b in zip(ref, got):
        for k in ("wall_collisions", "pp_collisions", "pair_checks_ref", "oob_after_walls", "oob_after_pp", "errors", "completed_paths"):
            assert a[k] == b[k], (k, a[k], b[k])
        assert np.array_equal(a["wall_hits"], b["wall_hits"])
    assert sim.state_digest() == ref_digest
    sim.close()
