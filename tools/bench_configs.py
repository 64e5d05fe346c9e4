#!/usr/bin/env python
"""Measured lines for the BASELINE.json configurations that bench.py's default run does not cover, one GPU:
  cfg1  Open_Air_Cube_MC.py as shipped (24,627 particles, serial sweep)
  cfg2  Open_Air_Pore_MC.py (557,649 particles, specular walls)
  cfg3  Temperature_Pore_MC.py (557,649 particles): device-RNG mode and host-RNG parity mode
One JSON line per configuration (CUDA-event device time per step; the L2 is flushed between timed steps because
these states fit into it).   python tools/bench_configs.py [cfg1 cfg2 cfg3 cfg3host] [--steps K]"""
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
from argon_monte_carlo_b200 import amc, config, init_state  # noqa: E402


def timed_steps(sim, k, warmup=3):
    flush = torch.empty(512 << 20, dtype=torch.uint8, device="cuda")
    sim.step_quiet(warmup)
    tot, launches = np.zeros(5), 0
    for _ in range(k):
        flush.fill_(1)
        torch.cuda.synchronize()
        sim.step_quiet(1)
        ms, launches = sim.last_timing()
        tot += np.array(ms)
    return tot / k, launches


def line(name, n, ms, launches, extra=None):
    d = {"config": name, "particles": n, "ms_per_step": ms[4], "value": n / (ms[4] * 1e-3), "unit": "particle-steps/s",
         "phases_ms": {"advect_or_keys": ms[0], "sort": ms[1], "pairs": ms[2], "recapture": ms[3]}, "launches_per_step": launches,
         "l2": "flushed between timed steps"}
    d.update(extra or {})
    print(json.dumps(d), flush=True)


def main():
    steps = int(sys.argv[sys.argv.index("--steps") + 1]) if "--steps" in sys.argv else 50
    args = [a for a in sys.argv[1:] if a.startswith("cfg")]
    which = args or ["cfg1", "cfg2", "cfg3", "cfg3host"]
    if "cfg1" in which:
        cfg = config.cube_config()
        st = init_state.cube_initial_state(cfg)
        sim = amc.Simulation(cfg)
        sim.set_state(*st)
        ms, l = timed_steps(sim, steps)
        line("cfg1 Open_Air_Cube_MC.py as shipped (serial sweep, k_sweep_detect + k_sweep_events)", len(st[0]), ms, l)
        sim.close()
    if "cfg2" in which:
        cfg = config.pore_config(False)
        st = init_state.pore_initial_state(cfg)
        sim = amc.Simulation(cfg)
        sim.set_state(*st)
        ms, l = timed_steps(sim, steps)
        line("cfg2 Open_Air_Pore_MC.py (specular walls)", len(st[0]), ms, l, {"pair_detect_ms": sim.last_detect_ms()})
        sim.close()
    if "cfg3" in which or "cfg3host" in which:
        cfg = config.pore_config(True)
        st = init_state.pore_initial_state(cfg)
        if "cfg3" in which:
            sim = amc.Simulation(cfg, seed=17)
            sim.set_state(*st)
            ms, l = timed_steps(sim, steps)
            line("cfg3 Temperature_Pore_MC.py, device RNG (Philox)", len(st[0]), ms, l, {"pair_detect_ms": sim.last_detect_ms()})
            sim.close()
        if "cfg3host" in which:
            sim = amc.Simulation(cfg, rng_mode=amc.RNG_HOST)
            sim.set_state(*st)
            sim.step_host_rng()
            t0 = time.perf_counter()
            k = 3
            for _ in range(k):
                s = sim.step_host_rng()
            torch.cuda.synchronize()
            dt = (time.perf_counter() - t0) / k
            print(json.dumps({"config": "cfg3 Temperature_Pore_MC.py, host-RNG parity mode (Mersenne-Twister draws + mpmath on the host, "
                                        "one device round trip per wall case)", "particles": len(st[0]), "ms_per_step": dt * 1e3,
                              "value": len(st[0]) / dt, "unit": "particle-steps/s", "timing": "wall clock (host-bound)",
                              "wall_hits_last_step": int(s["wall_collisions"])}), flush=True)
            sim.close()


if __name__ == "__main__":
    main()
