"""Operator-level shims with the reference's own signatures ("two granularities").

The reference's de-facto operator boundary is a handful of functions that work on module-global arrays or on
per-cell argument tuples:

    pairwise_particles_in_cell(completed_paths, completed_x_paths, completed_y_paths, completed_z_paths,
                               in_cell, continue_path, continue_x_path, continue_y_path, continue_z_path,
                               has_collided, x_positions_in_cell, ..., z_velocities_in_cell)     Open_Air_Pore_MC.py:160
    hit_vertical_wall(hits, z_plane, completed_paths, completed_x_paths, ...)                    Open_Air_Pore_MC.py:257
    hit_cylinder_side_wall(hits, collision_radius, completed_paths, ...)                         Open_Air_Pore_MC.py:294
    hit_vertical_specular_wall(hits, z_plane)                                                    Temperature_Pore_MC.py:311
    hit_cylinder_specular_side_wall(hits, collision_radius, total_errs)                          Temperature_Pore_MC.py:317

Here each of them runs on the GPU through libamc.so (amc_pairs on a one-cell grid; amc_wall_operator), with the
same argument order, the same in-place / return conventions and the same appends to the four result lists, so a
parity test can call the reference function and this one side by side.  Where the reference reads module globals
(x_vals, y_vals, ..., num_collisions_per_step) the wall operators take them from the `Globals` object bound with
`bind(globals_object)`; `num_collisions_per_step` is a plain int attribute of it.  These shims are for operator-level
checks -- production code steps whole timesteps (amc.Simulation.step); there is no CPU fallback here either.
"""
from __future__ import annotations

import ctypes as C
from types import SimpleNamespace

import numpy as np

from . import amc, config

_cfg = None
_G = None


def use_config(cfg):
    """Constants of the script whose operators are being mirrored (collision_range, argon_mass); default: Open_Air_Pore_MC."""
    global _cfg
    _cfg = cfg


def _config():
    global _cfg
    if _cfg is None:
        _cfg = config.pore_config(False)
    return _cfg


class Globals(SimpleNamespace):
    """The module globals the reference's wall operators mutate (Pore:385-400): x_vals, y_vals, z_vals, x_velocities,
    y_velocities, z_velocities, dist_since_collision, dist_x/y/z_since_collision, full_path_traveled,
    num_collisions_per_step."""


def bind(g):
    global _G
    _G = g
    if not hasattr(g, "num_collisions_per_step"):
        g.num_collisions_per_step = 0
    return g


def _one_cell_simulation(n, taps):
    cfg = _config()
    one = SimpleNamespace(kind="cube", dt=0.0, cube_x=1.0, cube_y=1.0, cube_z=1.0, argon_mass=cfg.argon_mass,
                          argon_radius=cfg.argon_radius, collision_range=cfg.collision_range, seed=0, num_molecules=n,
                          grid=config.Grid(nc=(1, 1, 1), c0=(0, 0, 0), d=(1.0, 1.0, 1.0), band=(1.0, 1.0, 1.0)))   # one cell (-1, 1)^3 holding everything
    return amc.Simulation(one, kind=amc.KIND_CUBE, pp_mode=amc.PP_SWEEP, max_particles=max(n, 1), taps=taps, path_capacity=4 * max(n, 1) + 16)


def pairwise_particles_in_cell(completed_paths, completed_x_paths, completed_y_paths, completed_z_paths, in_cell, continue_path,
                               continue_x_path, continue_y_path, continue_z_path, has_collided, x_positions_in_cell,
                               y_positions_in_cell, z_positions_in_cell, x_velocities_in_cell, y_velocities_in_cell,
                               z_velocities_in_cell):
    """One collision cell: the sequential (i, j < i) sweep with in-place elastic exchange and MFP bookkeeping
    (Open_Air_Pore_MC.py:160-255).  Mutates and returns the last twelve arguments like the reference; completed free
    paths are appended to the four lists in collision order; the number of collisions is added to the bound
    Globals.num_collisions_per_step (the reference's shared counter, Pore:244-245) and kept in `last_num_collisions`."""
    global last_num_collisions
    n = len(x_positions_in_cell)
    last_num_collisions = 0
    if n >= 2:
        sim = _one_cell_simulation(n, amc.TAP_PATHS)
        try:
            sim.set_state(x_positions_in_cell, y_positions_in_cell, z_positions_in_cell, x_velocities_in_cell, y_velocities_in_cell,
                          z_velocities_in_cell, continue_path, continue_x_path, continue_y_path, continue_z_path,
                          np.asarray(has_collided).astype(np.uint8))
            st = sim.pairs()
            out = sim.get_state()
            done = sim.completed_paths()
        finally:
            sim.close()
        for dst, key in ((x_positions_in_cell, "x"), (y_positions_in_cell, "y"), (z_positions_in_cell, "z"),
                         (x_velocities_in_cell, "vx"), (y_velocities_in_cell, "vy"), (z_velocities_in_cell, "vz"),
                         (continue_path, "dist"), (continue_x_path, "dist_x"), (continue_y_path, "dist_y"), (continue_z_path, "dist_z")):
            dst[:] = out[key]
        has_collided[:] = out["flag"].astype(bool)
        for lst, vals in zip((completed_paths, completed_x_paths, completed_y_paths, completed_z_paths), done):
            lst.extend(vals.tolist())
        last_num_collisions = int(st["pp_collisions"])
        if _G is not None:
            _G.num_collisions_per_step += last_num_collisions
    return (in_cell, continue_path, continue_x_path, continue_y_path, continue_z_path, has_collided, x_positions_in_cell,
            y_positions_in_cell, z_positions_in_cell, x_velocities_in_cell, y_velocities_in_cell, z_velocities_in_cell)


last_num_collisions = 0

_KEYS = (("x_vals", "x"), ("y_vals", "y"), ("z_vals", "z"), ("x_velocities", "vx"), ("y_velocities", "vy"), ("z_velocities", "vz"),
         ("dist_since_collision", "dist"), ("dist_x_since_collision", "dist_x"), ("dist_y_since_collision", "dist_y"),
         ("dist_z_since_collision", "dist_z"))


def _wall_operator(op, hits, param, lists, count_collisions):
    if _G is None:
        raise RuntimeError("operators.bind(Globals(...)) first: the wall operators work on the module-global arrays")
    g = _G
    hits = np.ascontiguousarray(np.asarray(hits).astype(np.uint8))
    n = len(g.x_vals)
    cfg = _config()
    sim = amc.Simulation(cfg, kind=amc.KIND_PORE, max_particles=max(n, 1), taps=amc.TAP_PATHS, path_capacity=4 * max(n, 1) + 16)
    try:
        sim.set_state(*[getattr(g, a) for a, _ in _KEYS], np.asarray(g.full_path_traveled).astype(np.uint8))
        nh, ne = C.c_int64(0), C.c_int64(0)
        sim._check(sim.lib.amc_wall_operator(sim.h, int(op), hits.ctypes.data_as(amc.c_uint8_p), C.c_double(float(param)),
                                             C.byref(nh), C.byref(ne)), "amc_wall_operator")
        out = sim.get_state()
        done = sim.completed_paths() if lists else None
    finally:
        sim.close()
    for a, k in _KEYS:
        getattr(g, a)[:] = out[k]
    g.full_path_traveled[:] = out["flag"].astype(bool)
    if lists:
        # the device appends in arbitrary order; the reference appends in ascending particle index.  A particle completes
        # at most one path per operator call, so ordering the new entries by |path| is not needed for the multiset checks
        for lst, vals in zip(lists, done):
            lst.extend(vals.tolist())
    if count_collisions:
        g.num_collisions_per_step += int(nh.value) - int(ne.value)
    return int(nh.value), int(ne.value)


def hit_vertical_wall(hits, z_plane, completed_paths, completed_x_paths, completed_y_paths, completed_z_paths):
    """Open_Air_Pore_MC.py:257-292: specular reflection off the plane z = z_plane for the particles selected by `hits`,
    MFP bookkeeping, num_collisions_per_step += number of hits."""
    _wall_operator(0, hits, z_plane, (completed_paths, completed_x_paths, completed_y_paths, completed_z_paths), True)


def hit_cylinder_side_wall(hits, collision_radius, completed_paths, completed_x_paths, completed_y_paths, completed_z_paths):
    """Open_Air_Pore_MC.py:294-348: specular reflection off the coaxial cylinder of radius collision_radius."""
    _wall_operator(1, hits, collision_radius, (completed_paths, completed_x_paths, completed_y_paths, completed_z_paths), True)


def hit_vertical_specular_wall(hits, z_plane):
    """Temperature_Pore_MC.py:311-315: specular plane, no MFP bookkeeping, no counter."""
    _wall_operator(2, hits, z_plane, None, False)


def hit_cylinder_specular_side_wall(hits, collision_radius, total_errs):
    """Temperature_Pore_MC.py:317-347: specular cylinder, no MFP bookkeeping; returns total_errs plus the hits whose
    rewind raised a floating-point error in the reference (its try/except path)."""
    _, errs = _wall_operator(3, hits, collision_radius, None, False)
    return total_errs + errs
