#!/usr/bin/env python
"""A/B of libamc.so under different environments in ONE short process (no torch import: a run costs seconds of GPU time).

  python tools/ab_env.py [--particles 12500000,557649] [--steps 30] [--warmup 4] ARM [ARM ...]
  ARM = name[:VAR=val[,VAR=val ...]]        e.g.  base  pdl1:AMC_PDL=1  pdl3:AMC_PDL=3

The knobs of libamc.so are read when a handle is created, so every arm makes its own handle on the same synthetic
energized pore (device-side initialiser, seed 17), runs warm-up + timed steps in one amc_step call and prints one JSON
line: CUDA-event phase times per step, the duration of the detection and scatter launches, and the state checksum, which
must be the same in every arm.  Arms are run twice in alternating order (a b c c b a) and the better time is kept."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from argon_monte_carlo_b200 import amc, config, init_state  # noqa: E402


def run_arm(n_particles, env, steps, warmup):
    saved = {k: os.environ.get(k) for k in env}
    os.environ.update(env)
    try:
        cfg = config.pore_config(True, scale=(n_particles / 557649) ** (1.0 / 3.0))
        n = cfg.num_molecules
        sim = amc.Simulation(cfg, seed=17, device=0, max_particles=n)
        sim.init_synthetic(init_state.pore_spec(cfg, seed=17))
        sim.step_quiet(warmup)
        stats = sim.step(steps)
        ms, launches = sim.last_timing()
        out = {"particles": n, "ms_per_step": ms[4] / steps, "keys": ms[0] / steps, "scan_scatter": ms[1] / steps,
               "pairs": ms[2] / steps, "recapture": ms[3] / steps, "detect": sim.last_detect_ms() / steps,
               "scatter": sim.last_scatter_ms() / steps, "collisions": int(sum(s["collisions"] for s in stats)),
               "digest": ["%016x" % d for d in sim.state_digest()[:2]]}
        sim.close()
        return out
    finally:
        for k, v in saved.items():
            if v is None:
                os.environ.pop(k, None)
            else:
                os.environ[k] = v


def main():
    argv = sys.argv[1:]
    opts = {"--particles": "12500000", "--steps": "30", "--warmup": "4"}
    arms = []
    while argv:
        a = argv.pop(0)
        if a in opts:
            opts[a] = argv.pop(0)
        else:
            name, _, spec = a.partition(":")
            arms.append((name, dict(kv.split("=", 1) for kv in spec.split(",") if kv)))
    steps, warmup = int(opts["--steps"]), int(opts["--warmup"])
    for n in [int(v) for v in opts["--particles"].split(",")]:
        best = {}
        for name, env in arms + arms[::-1]:
            r = run_arm(n, env, steps, warmup)
            if name not in best or r["ms_per_step"] < best[name]["ms_per_step"]:
                best[name] = r
        for name, _ in arms:
            r = best[name]
            r["arm"] = name
            print(json.dumps({k: (round(v, 5) if isinstance(v, float) else v) for k, v in r.items()}), flush=True)


if __name__ == "__main__":
    main()
