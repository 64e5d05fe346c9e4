"""Workload for compute-sanitizer (memcheck / racecheck / synccheck): every kernel family once, small enough to
finish under the tool.   compute-sanitizer --tool memcheck python tools/sanitize_target.py [pore|temp|slab|cube]..."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from argon_monte_carlo_b200 import amc, config, init_state, slab  # noqa: E402

which = sys.argv[1:] or ["pore", "temp", "slab", "cube"]
scale = float(os.environ.get("AMC_SAN_SCALE", "0.5"))
if "pore" in which:
    cfg = config.pore_config(False, scale=scale)
    st = init_state.synthetic_pore_state(cfg, seed=11)          # overlapping start: thousands of collisions per step
    sim = amc.Simulation(cfg, max_particles=len(st[0]), taps=amc.TAP_PAIRS | amc.TAP_PATHS)
    sim.set_state(*st)
    s = sim.step(3)
    print("pore", len(st[0]), [x["collisions"] for x in s], sim.state_digest())
    sim.close()
if "temp" in which:
    cfg = config.pore_config(True, scale=scale)
    st = init_state.synthetic_pore_state(cfg, seed=3)
    sim = amc.Simulation(cfg, seed=3, max_particles=len(st[0]))
    sim.set_state(*st)
    s = sim.step(3)
    print("temp", len(st[0]), [x["collisions"] for x in s])
    sim.get_state()
    sim.close()
if "slab" in which:
    cfg = config.pore_config(True, scale=scale)
    st = init_state.synthetic_pore_state(cfg, seed=23)
    nz = cfg.grid.nc[2]
    sim = slab.SlabSimulation(cfg, 4, st[2], cuts=[0, 1, 2, nz - 1, nz], seed=23)     # every cut inside a dense end cap
    sim.set_state(*st)
    s = sim.step(4)
    print("slab x4", len(st[0]), [x["collisions"] for x in s], sim.state_digest())
    sim.close()
if "cube" in which:
    cfg = config.cube_config()
    st = init_state.cube_initial_state(cfg)
    sim = amc.Simulation(cfg)
    sim.set_state(*st)
    s = sim.step(2)
    print("cube", len(st[0]), [x["pp_collisions"] for x in s])
    sim.close()
