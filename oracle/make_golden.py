"""Generate the committed golden vectors under tests/golden/ (build container only).

  python oracle/make_golden.py run   -- runs the UNMODIFIED reference scripts for 3 steps through
                                        oracle/run_reference.py (about 9 minutes on 8 cores) into
                                        /tmp/refrun_{pore,temp}
  python oracle/make_golden.py pack  -- condenses those dumps into small JSON fixtures: per step and
                                        per array a SHA-256 of the raw float64 bytes of all 557,649
                                        particles, the collision counters, the completed-path lists
                                        and (Temp) the per-step momentum / energy series
  python oracle/make_golden.py cells -- calls the reference's own functions (imported under the
                                        matplotlib stub) on small seeded inputs and stores inputs
                                        and outputs: pairwise_particles_in_cell, hit_vertical_wall,
                                        hit_cylinder_side_wall, the Temp coated/gap operators with
                                        a fixed direction, num_out_of_bounds / recapture

The shipped result files of the reference (momentum_energy.csv, hist_*_data.txt) are copied
verbatim as data fixtures by `pack`.
"""
import hashlib
import json
import os
import shutil
import subprocess
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
GOLD = os.path.join(ROOT, "tests", "golden")
REF = os.environ.get("AMC_REFERENCE_DIR", "/root/reference")
KEYS = ("x", "y", "z", "vx", "vy", "vz", "dist", "dist_x", "dist_y", "dist_z", "flag")


def digest(a):
    a = np.ascontiguousarray(a)
    if a.dtype == np.bool_:
        a = a.astype(np.uint8)
    return hashlib.sha256(a.tobytes()).hexdigest()


def run():
    for kind in ("temp", "pore"):
        subprocess.check_call([sys.executable, os.path.join(HERE, "run_reference.py"), kind, "3", "/tmp/refrun_" + kind])


def pack():
    os.makedirs(GOLD, exist_ok=True)
    for kind in ("pore", "temp"):
        d = "/tmp/refrun_" + kind
        out = {"source": "unmodified %s run for 3 steps via oracle/run_reference.py" %
               {"pore": "Open_Air_Pore_MC.py", "temp": "Temperature_Pore_MC.py"}[kind],
               "numpy": np.__version__, "steps": []}
        for k in range(3):
            rec = {}
            for tag in ("walls", "end"):
                z = np.load(os.path.join(d, "%s_%d.npz" % (tag, k)))
                rec[tag] = {key: digest(z[key]) for key in KEYS}
                rec[tag]["ncol"] = int(z["ncol"])
                rec[tag]["sample_idx"] = list(range(0, len(z["x"]), 50021))
                rec[tag]["sample"] = {key: [float(v) for v in z[key][::50021]] for key in KEYS}
            out["steps"].append(rec)
        fin = np.load(os.path.join(d, "final.npz"))
        out["final"] = {k: [float(v).hex() for v in np.sort(fin[k]) if True] if k.startswith("completed")
                        else [float(v).hex() for v in fin[k]] for k in fin.files}
        with open(os.path.join(GOLD, "ref_%s_3steps.json" % kind), "w") as f:
            json.dump(out, f, indent=1)
    for name in sorted(os.listdir(REF)):
        if name.endswith(".txt") or name.endswith(".csv"):
            shutil.copy(os.path.join(REF, name), os.path.join(GOLD, "shipped_" + name))


def cells():
    sys.path.insert(0, ROOT)
    from multiprocessing import Value
    from oracle import ref_import
    rng = np.random.default_rng(20261018)
    P = ref_import.load("Open_Air_Pore_MC")
    T = ref_import.load("Temperature_Pore_MC")
    P.init_globals(Value("i", 0))
    T.init_globals(Value("i", 0))
    out = {"pairwise": [], "pore_walls": [], "temp_walls": [], "recapture": []}
    cr = float(P.collision_range)

    def hexl(a):
        return [float(v).hex() for v in np.asarray(a, dtype=np.float64).ravel()]

    # --- pairwise_particles_in_cell on dense random cells (many overlaps, chained collisions)
    for n, box in ((2, 1.0), (3, 1.2), (8, 2.5), (40, 5.0), (120, 9.0)):
        for rep in range(3):
            pos = rng.uniform(0, box * cr, (3, n)) + 1e-7
            vel = rng.normal(0, 250.0, (3, n))
            paths = rng.uniform(0, 1e-7, (4, n))
            flag = rng.random(n) < 0.5
            args = [pos[0].copy(), pos[1].copy(), pos[2].copy(), vel[0].copy(), vel[1].copy(), vel[2].copy()]
            lists = [[], [], [], []]
            P.num_collisions_per_step.value = 0
            res = P.pairwise_particles_in_cell(*lists, np.ones(n, bool), paths[0].copy(), paths[1].copy(),
                                               paths[2].copy(), paths[3].copy(), flag.copy(), *args)
            out["pairwise"].append({
                "n": n, "pos": hexl(pos), "vel": hexl(vel), "paths": hexl(paths), "flag": [int(v) for v in flag],
                "ncol": int(P.num_collisions_per_step.value),
                "out_paths": [hexl(res[k]) for k in (1, 2, 3, 4)], "out_flag": [int(v) for v in res[5]],
                "out_pos": [hexl(res[k]) for k in (6, 7, 8)], "out_vel": [hexl(res[k]) for k in (9, 10, 11)],
                "completed": [hexl(l) for l in lists]})

    # --- wall operators driven through the module globals
    def set_state(M, n, region):
        M.num_molecules = n
        if region == "side":       # just outside the open-air radius
            r = float(M.open_air_radius) * (1 + rng.uniform(1e-6, 2e-3, n))
            th = rng.uniform(0, 2 * np.pi, n)
            M.x_vals, M.y_vals = r * np.cos(th), r * np.sin(th)
            M.z_vals = rng.uniform(0, float(M.total_height), n)
        else:                       # just past a plane
            r = float(M.open_air_radius) * np.sqrt(rng.uniform(0, 0.9, n))
            th = rng.uniform(0, 2 * np.pi, n)
            M.x_vals, M.y_vals = r * np.cos(th), r * np.sin(th)
            M.z_vals = -rng.uniform(1e-12, 1e-10, n)
        v = rng.normal(0, 250.0, (3, n))
        if region == "side":        # moving outward
            rad = np.stack([M.x_vals, M.y_vals]) / np.hypot(M.x_vals, M.y_vals)
            sp = np.abs(rng.normal(0, 250.0, n)) + 1.0
            tv = sp * rng.uniform(-0.5, 0.5, n)  # tangential part small enough that the path crosses R - a
            v[0], v[1] = rad[0] * sp - rad[1] * tv, rad[1] * sp + rad[0] * tv
        else:
            v[2] = -np.abs(v[2]) - 1.0
        M.x_velocities, M.y_velocities, M.z_velocities = v[0].copy(), v[1].copy(), v[2].copy()
        p = rng.uniform(0, 1e-7, (4, n))
        M.dist_since_collision, M.dist_x_since_collision = p[0].copy(), p[1].copy()
        M.dist_y_since_collision, M.dist_z_since_collision = p[2].copy(), p[3].copy()
        M.full_path_traveled = rng.random(n) < 0.5
        M.prior_x_vals, M.prior_y_vals, M.prior_z_vals = M.x_vals.copy(), M.y_vals.copy(), M.z_vals.copy()

    def snap(M):
        return {"pos": hexl([M.x_vals, M.y_vals, M.z_vals]),
                "vel": hexl([M.x_velocities, M.y_velocities, M.z_velocities]),
                "paths": hexl([M.dist_since_collision, M.dist_x_since_collision, M.dist_y_since_collision,
                               M.dist_z_since_collision]),
                "flag": [int(v) for v in M.full_path_traveled]}

    n = 24
    for region, fn in (("side", lambda l: P.hit_cylinder_side_wall(np.ones(n, bool), P.open_air_collision_radius, *l)),
                       ("plane", lambda l: P.hit_vertical_wall(np.ones(n, bool), 0, *l))):
        set_state(P, n, region)
        before = snap(P)
        lists = [[], [], [], []]
        fn(lists)
        out["pore_walls"].append({"op": region, "n": n, "before": before, "after": snap(P),
                                  "completed": [hexl(l) for l in lists]})

    # Temp energized operators with the random direction pinned (random_inbounds_direction patched to
    # a deterministic function of the normal so no RNG state is involved)
    def fixed_direction(norm):
        d = np.array([0.3, -0.2, 0.0]) + 0.9 * np.asarray(norm, dtype=float)
        return d / np.sqrt(np.dot(d, d))
    T.random_inbounds_direction = fixed_direction
    for region in ("side", "plane"):
        set_state(T, n, region)
        before = snap(T)
        lists = [[], [], [], []]
        if region == "side":
            normals_in = None
            dp, de, errs = T.hit_cylinder_coated_side_wall(np.ones(n, bool), T.surface_energy_hot,
                                                          T.pore_collision_radius * 0 + T.open_air_collision_radius,
                                                          *lists, 0)
        else:
            dp, de = T.hit_vertical_coated_wall(np.ones(n, bool), T.surface_energy_cold, 0.0, 1, *lists)
        out["temp_walls"].append({"op": region, "n": n, "before": before, "after": snap(T),
                                  "dpz": float(dp).hex(), "de": float(de).hex(),
                                  "completed": [hexl(l) for l in lists]})
    # gap wall: surface_energy_gap per hit
    set_state(T, 6, "side")
    scale = float(T.gap_collision_radius) / float(T.open_air_radius)
    T.x_vals *= scale; T.y_vals *= scale
    T.z_vals = rng.uniform(float(T.gap_bottom_height), float(T.gap_top_height), 6)
    before = snap(T)
    lists = [[], [], [], []]
    dp, errs = T.hit_cylinder_gap_side_wall(np.ones(6, bool), T.gap_collision_radius, *lists, 0)
    out["temp_walls"].append({"op": "gap", "n": 6, "before": before, "after": snap(T), "dpz": float(dp).hex(),
                              "completed": [hexl(l) for l in lists],
                              "surface_energy_gap": {float(z).hex(): float(T.surface_energy_gap(z)).hex()
                                                     for z in np.linspace(float(T.gap_bottom_height), float(T.gap_top_height), 9)}})

    # recapture / census
    for M, name in ((P, "pore"), (T, "temp")):
        n2 = 400
        M.x_vals = rng.uniform(-1.1, 1.1, n2) * float(M.open_air_radius)
        M.y_vals = rng.uniform(-1.1, 1.1, n2) * float(M.open_air_radius)
        M.z_vals = rng.uniform(-0.02, 1.02, n2) * float(M.total_height)
        M.prior_x_vals, M.prior_y_vals, M.prior_z_vals = M.x_vals.copy(), M.y_vals.copy(), M.z_vals.copy()
        before = hexl([M.x_vals, M.y_vals, M.z_vals])
        rec = {"kind": name, "n": n2, "before": before}
        if name == "pore":
            rec["count"] = int(M.num_out_of_bounds())
        else:
            import contextlib, io
            with contextlib.redirect_stdout(io.StringIO()):
                rec["report"] = int(M.num_out_of_bounds())
            rec["count"] = int(M.recapture_out_of_bounds())
        rec["after"] = hexl([M.x_vals, M.y_vals, M.z_vals])
        out["recapture"].append(rec)
    os.makedirs(GOLD, exist_ok=True)
    with open(os.path.join(GOLD, "ref_operators.json"), "w") as f:
        json.dump(out, f)


def init():
    """SHA-256 of the initial positions / velocities produced by the reference's own
    init_positions() / init_velocities() under its seeds."""
    sys.path.insert(0, ROOT)
    import random
    from oracle import ref_import
    out = {}
    for name, kind in (("Open_Air_Pore_MC", "pore"), ("Temperature_Pore_MC", "temp")):
        M = ref_import.load(name)          # seeds both generators at import (Pore:89-90)
        x, y, z = M.init_positions()
        vx, vy, vz = M.init_velocities()
        out[kind] = {k: digest(v) for k, v in zip(("x", "y", "z", "vx", "vy", "vz"), (x, y, z, vx, vy, vz))}
        out[kind]["n"] = int(len(x))
        out[kind]["next_np_uniform"] = float(np.random.uniform()).hex()
        out[kind]["next_py_random"] = float(random.random()).hex()
    with open(os.path.join(GOLD, "ref_init.json"), "w") as f:
        json.dump(out, f, indent=1)


if __name__ == "__main__":
    {"run": run, "pack": pack, "cells": cells, "init": init}[sys.argv[1]]()
