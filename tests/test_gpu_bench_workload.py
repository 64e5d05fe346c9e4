"""Parity on the benchmarked configuration itself: the exact workload bench.py times (energized pore scaled
to 12.5 M particles, seed 17, device RNG) is run for two timesteps on the GPU and through the CPU oracle;
state, counters and collision-pair set must be identical.  (The oracle needs ~10 s per step here.)"""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

KEYS = ("x", "y", "z", "vx", "vy", "vz", "dist", "dist_x", "dist_y", "dist_z")


def test_bench_workload_two_steps_bit_exact(oracle):
    import bench
    from argon_monte_carlo_b200 import amc, config, init_state
    from oracle import steps
    cfg, _ = bench.scaled_temp_config(12_500_000)
    assert cfg.num_molecules == 12_499_989
    cheb = config.gap_energy_chebyshev(cfg, 16)
    state = init_state.synthetic_pore_state(cfg, seed=17)
    n = len(state[0])
    st = oracle.ParticleState(*state)
    sim = amc.Simulation(cfg, seed=17, cheb=cheb, max_particles=n, taps=amc.TAP_PAIRS, pair_capacity=1 << 20)
    sim.set_state(*state)
    pairs = oracle.PairSink()
    for k in range(2):
        r = steps.temp_step_philox(st, cfg, 17, k, cheb, None, pairs)
        g = sim.step(1)[0]
        assert np.array_equal(g["wall_hits"], r["wall_counts"]), k
        assert g["collisions"] == r["collisions"] and g["pp_collisions"] == r["pp_collisions"], k
        assert g["pair_checks_ref"] == r["checks"], k
        assert g["oob_after_walls"] == r["oob_after_walls"] and g["oob_after_pp"] == r["oob_after_pp"], k
        assert g["errors"] == r["errors"], k
        for key in ("dpz", "e_cold", "e_hot"):
            assert abs(g[key] - r[key]) <= 1e-12 * abs(r[key]), key   # fixed-point accumulation vs sequential sum
    got = sim.get_state()
    for key in KEYS:
        bad = np.nonzero(got[key] != getattr(st, key))[0]
        assert len(bad) == 0, "%s differs for %d of %d particles (first id %d)" % (key, len(bad), n, bad[0])
    assert np.array_equal(got["flag"].astype(bool), st.flag.astype(bool))
    hi, lo, grp, cell = sim.pair_list()
    ohi, olo, ogrp, ocell = pairs.arrays()
    assert len(hi) == len(ohi) > 20000
    assert sorted(zip(hi.tolist(), lo.tolist(), grp.tolist(), cell.tolist())) == \
        sorted(zip(ohi.tolist(), olo.tolist(), ogrp.tolist(), ocell.tolist()))
    # the device-side checksum bench.py prints equals the checksum of the oracle's state
    assert sim.state_digest() == amc.digest_of_arrays(np.arange(n), st.arrays())
    sim.close()
