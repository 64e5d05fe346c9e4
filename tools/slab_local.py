"""Emulate an N-rank slab run on ONE GPU with the local transport (debug aid):
python tools/slab_local.py NRANKS PARTICLES_TOTAL STEPS"""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from argon_monte_carlo_b200 import config, init_state, slab

nranks, total, steps = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
cfg = config.pore_config(True, scale=(total / 557649) ** (1 / 3))
ids, *state = init_state.synthetic_pore_chunked(cfg, 17)
_, _, _, zs, _, _, _ = init_state.synthetic_pore_chunk(cfg, 17, 0, 1 << 22)
cuts = slab.balanced_cuts(zs, cfg.grid.edge[2], nranks)
print("cuts", cuts, "cells", cfg.grid.nc)
sim = slab.SlabSimulation(cfg, nranks, zs, cuts=cuts, n_total=cfg.num_molecules, seed=17)
print("xfer_capacity", sim.ranks[0].xfer_send.shape)
sim.debug_counts = True
sim.set_state(*state)
for k in range(steps):
    before = dict(sim.exchanged)
    st = sim.step(1)[0]
    print(k, "collisions", st["collisions"], "per rank", sim.particles_per_rank(),
          "xfer", sim.exchanged["xfer"] - before["xfer"], "boundary", sim.exchanged["boundary"] - before["boundary"])
sim.close()
