#!/usr/bin/env python
"""CPU baseline of SURVEY section 8(d): the UNMODIFIED upstream scripts timed on the host cores.

  python baseline/stage.py stage            copy the five upstream files from /root/reference into the git-ignored
                                            baseline/_ref/ (the GPU box only receives /root/repo; build container only)
  python baseline/stage.py run {pore|temp|cube} [K]
                                            run the staged script for K timesteps and print one JSON line

The only edit is the loop bound (`range(num_timesteps)` -> `range(K)`, Open_Air_Pore_MC.py:416 /
Temperature_Pore_MC.py:662 / Open_Air_Cube_MC.py:175), applied to a scratch copy under baseline/_ref/_run/;
matplotlib (absent from the image, imported by every script at line 1) comes from the 30-line stub in
oracle/stubs.  Times are the scripts' own prints "Wall Step Runtime" (Pore:517 / Temp:810) and
"Particle-Particle step Runtime" (Pore:554 / Temp:849); the cube script prints no timings, so its step loop is
timed as the wall-clock difference of a 2K-step and a K-step run (same init and post-processing in both).  The pore scripts start cpu_count()+1
worker processes (Pore:86), so `cores` is os.cpu_count().

Nothing here is imported by the product package; bench.py calls run() for its cpu_baseline / --impl reference legs.
"""
from __future__ import annotations

import json
import os
import re
import shutil
import subprocess
import sys
import time

HERE = os.path.dirname(os.path.abspath(__file__))
REF_DIR = os.path.join(HERE, "_ref")
STUBS = os.path.join(os.path.dirname(HERE), "oracle", "stubs")
FILES = ("Open_Air_Cube_MC.py", "Open_Air_Pore_MC.py", "Temperature_Pore_MC.py", "utils.py", "graph_sim_data.py")
SCRIPT = {"cube": "Open_Air_Cube_MC.py", "pore": "Open_Air_Pore_MC.py", "temp": "Temperature_Pore_MC.py"}
PARTICLES = {"cube": 24627, "pore": 557649, "temp": 557649}


def stage(src="/root/reference"):
    if not os.path.isdir(src):
        return False
    os.makedirs(REF_DIR, exist_ok=True)
    for f in FILES:
        shutil.copyfile(os.path.join(src, f), os.path.join(REF_DIR, f))
    return True


def staged():
    return all(os.path.isfile(os.path.join(REF_DIR, f)) for f in FILES)


def _run_script(kind, k, timeout):
    work = os.path.join(REF_DIR, "_run", "%s_%d_%d" % (kind, k, os.getpid()))
    os.makedirs(work, exist_ok=True)
    src = open(os.path.join(REF_DIR, SCRIPT[kind])).read()
    if src.count("range(num_timesteps)") != 1:
        raise RuntimeError("unexpected loop header in " + SCRIPT[kind])
    with open(os.path.join(work, SCRIPT[kind]), "w") as f:
        f.write(src.replace("range(num_timesteps)", "range(%d)" % k))
    shutil.copyfile(os.path.join(REF_DIR, "utils.py"), os.path.join(work, "utils.py"))
    env = dict(os.environ, PYTHONPATH=STUBS)
    env.pop("OMP_NUM_THREADS", None)      # torchrun sets it to 1 for its children; the reference never limits NumPy
    t0 = time.perf_counter()
    p = subprocess.run([sys.executable, "-W", "ignore", SCRIPT[kind]], cwd=work, env=env, capture_output=True, text=True,
                       timeout=timeout)
    dt = time.perf_counter() - t0
    shutil.rmtree(work, ignore_errors=True)
    if p.returncode != 0:
        raise RuntimeError("%s failed: %s" % (SCRIPT[kind], p.stderr[-800:]))
    return p.stdout, dt


def run(kind="temp", steps=1, timeout=3600):
    """Time `steps` timesteps of the staged upstream script; returns the cpu_baseline dict of bench.py."""
    if not staged():
        raise RuntimeError("baseline/_ref is not staged (python baseline/stage.py stage, build container only)")
    n = PARTICLES[kind]
    out, total = _run_script(kind, steps, timeout)
    if kind == "cube":
        _, total2 = _run_script(kind, 2 * steps, timeout)
        per_step = [(total2 - total) / steps]
        detail = {"run_%d_steps_s" % steps: total, "run_%d_steps_s" % (2 * steps): total2}
        cores = 1
    else:
        wall = [float(v) for v in re.findall(r"Wall Step Runtime: ([0-9.eE+-]+) seconds", out)]
        pvp = [float(v) for v in re.findall(r"Particle-Particle step Runtime: ([0-9.eE+-]+) seconds", out)]
        if len(wall) != steps or len(pvp) != steps:
            raise RuntimeError("could not parse the timing prints of " + SCRIPT[kind])
        per_step = [a + b for a, b in zip(wall, pvp)]
        cols = [int(v) for v in re.findall(r"(\d+)\s+collisions from this timestep", out)]
        detail = {"wall_step_s": wall, "pp_step_s": pvp, "collisions_per_step": cols, "whole_run_s": total}
        cores = os.cpu_count()
    use = per_step[1:] if len(per_step) > 1 else per_step     # drop the initial-overlap transient when there is more than one step
    sec = sum(use) / len(use)
    if kind == "cube":
        how = "serial; per-step time = (run of %d steps - run of %d steps) / %d, wall clock" % (2 * steps, steps, steps)
    else:
        how = "cpu_count()+1 = %d worker processes; times = the script's own prints%s" % (
            cores + 1, ", step 0 (initial-overlap transient) dropped" if steps > 1 else "")
    return {"value": n / sec, "unit": "particle-steps/s", "cores": cores, "kind": "reference",
            "sample": "unmodified %s, loop bound patched to %d step%s, N=%d, %s" % (SCRIPT[kind], steps, "" if steps == 1 else "s", n, how),
            "ms_per_step": sec * 1e3, "detail": detail}


if __name__ == "__main__":
    if len(sys.argv) >= 2 and sys.argv[1] == "stage":
        print("staged" if stage() else "no /root/reference here")
    elif len(sys.argv) >= 3 and sys.argv[1] == "run":
        print(json.dumps(run(sys.argv[2], int(sys.argv[3]) if len(sys.argv) > 3 else 1)))
    else:
        print(__doc__)
