"""Whole-timestep drivers of the CPU oracle, mirroring the reference's step loops
(Open_Air_Pore_MC.py:416-557, Temperature_Pore_MC.py:662-853, Open_Air_Cube_MC.py:175-338).
Test infrastructure only."""
from __future__ import annotations

import numpy as np

from . import oracle as O  # noqa: E402  (package-relative when imported as oracle.steps)

ENERGIZED = (3, 4, 5, 6, 7, 8, 9)
COLD_CASES, HOT_CASES = (3, 7, 9), (4, 6, 8)


def pore_step(st, cfg, sink=None, pairs=None, want_bits=False):
    """One Open_Air_Pore_MC timestep. Returns dict of the counters the script prints."""
    O.drift(st, cfg.dt, True)
    counts, errs, bits = O.pore_walls(st, cfg.geom, sink, want_bits)
    oob_walls = O.pore_recapture(st, cfg.geom)
    ncol, checks, perr = O.pp_groups(st, cfg.grid, cfg.collision_range, cfg.argon_mass, sink, pairs)
    oob_pp = O.pore_recapture(st, cfg.geom)
    return dict(wall_counts=counts, wall_hits=int(counts.sum()), pp_collisions=ncol, checks=checks,
                errors=errs + perr, oob_after_walls=oob_walls, oob_after_pp=oob_pp, hit_bits=bits,
                collisions=int(counts.sum()) + ncol)


def temp_step_host_rng(st, cfg, sink=None, pairs=None, surface_energy_gap=None, want_bits=False):
    """One Temperature_Pore_MC timestep with the reference's host RNG (parity mode)."""
    from argon_monte_carlo_b200 import host_rng
    from argon_monte_carlo_b200.config import surface_energy_gap as seg
    O.drift(st, cfg.dt, True)
    dpz = 0
    e_hot = 0
    e_cold = 0
    counts = np.zeros(10, dtype=np.int64)
    errs = 0
    bits = np.zeros(st.n, dtype=np.uint16) if want_bits else None
    for case in range(10):
        idx, normal, colz = O.temp_case_detect(st, cfg.geom, case)
        counts[case] = len(idx)
        if want_bits:
            bits[idx] |= np.uint16(1 << case)
        dirs = surf = None
        if case in ENERGIZED:
            if case == 5:
                # reference order inside the loop: direction first (Temp:514), then the gap energy (Temp:519)
                dirs = np.zeros((len(idx), 3))
                surf = np.zeros(len(idx))
                for k in range(len(idx)):
                    if normal[k, 0] == normal[k, 0]:
                        dirs[k] = host_rng.inbound_direction(normal[k])
                        surf[k] = float((surface_energy_gap or (lambda z: seg(cfg, z)))(colz[k]))
            else:
                dirs = host_rng.directions_for_hits(normal)
        p, e, er = O.temp_case_apply(st, cfg.geom, case, idx, dirs, surf, sink)
        errs += er
        if case in ENERGIZED:
            dpz = dpz + p
            if case in COLD_CASES:
                e_cold = e_cold + e
            elif case in HOT_CASES:
                e_hot = e_hot + e
    oob_walls = O.temp_oob_count(st, cfg.geom)
    O.temp_recapture(st, cfg.geom)
    ncol, checks, perr = O.pp_groups(st, cfg.grid, cfg.collision_range, cfg.argon_mass, sink, pairs)
    oob_pp = O.temp_oob_count(st, cfg.geom)
    O.temp_recapture(st, cfg.geom)
    wall_hits = int(counts[3:].sum())
    return dict(wall_counts=counts, wall_hits=wall_hits, pp_collisions=ncol, checks=checks, errors=errs + perr,
                oob_after_walls=oob_walls, oob_after_pp=oob_pp, dpz=dpz, e_hot=e_hot, e_cold=e_cold,
                hit_bits=bits, collisions=wall_hits + ncol)


def temp_step_philox(st, cfg, seed, step, cheb, sink=None, pairs=None, want_bits=False):
    """One Temperature_Pore_MC timestep with the counter-based device RNG rule."""
    O.drift(st, cfg.dt, True)
    counts, sums, errs, bits = O.temp_walls_philox(st, cfg.geom, seed, step, cheb, sink, want_bits)
    oob_walls = O.temp_oob_count(st, cfg.geom)
    O.temp_recapture(st, cfg.geom)
    ncol, checks, perr = O.pp_groups(st, cfg.grid, cfg.collision_range, cfg.argon_mass, sink, pairs)
    oob_pp = O.temp_oob_count(st, cfg.geom)
    O.temp_recapture(st, cfg.geom)
    wall_hits = int(counts[3:].sum())
    return dict(wall_counts=counts, wall_hits=wall_hits, pp_collisions=ncol, checks=checks, errors=errs + perr,
                oob_after_walls=oob_walls, oob_after_pp=oob_pp, dpz=sums[0], e_cold=sums[1], e_hot=sums[2],
                hit_bits=bits, collisions=wall_hits + ncol)


def cube_step(st, cfg, sink=None, pairs=None):
    """One Open_Air_Cube_MC timestep (no prior positions, no MFP bookkeeping at walls)."""
    O.drift(st, cfg.dt, False)
    counts = O.cube_walls(st, cfg.cube_x, cfg.cube_y, cfg.cube_z)
    ncol, checks, perr = O.cube_pp_sweep(st, cfg.grid, cfg.collision_range, cfg.argon_mass, sink, pairs)
    return dict(wall_counts=counts, pp_collisions=ncol, checks=checks, errors=perr, collisions=ncol)
