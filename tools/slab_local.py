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
print("xfer buffer doubles", sim.ranks[0].xfer_send.shape)
sim.debug_counts = True
sim.set_state(*state)
for k in range(steps):
    before = dict(sim.exchanged)
    st = sim.step(1)[0]
    print(k, "collisions", st["collisions"], "per rank", sim.particles_per_rank(),
          "xfer", sim.exchanged["xfer"] - before["xfer"], "boundary", sim.exchanged["boundary"] - before["boundary"])
sim.close()

# ---- pure compute time of the phases per rank (ranks run one after the other here, so no waiting is included)
import torch, ctypes as C
sim = slab.SlabSimulation(cfg, nranks, zs, cuts=cuts, n_total=cfg.num_molecules, seed=17)
sim.set_state(*state)
sim.step(3)
T, R = sim.transport, sim.ranks
def timed(fn):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); fn(); e1.record(); torch.cuda.synchronize(); return e0.elapsed_time(e1)
acc = np.zeros((nranks, 5))
grp = np.zeros((nranks, 8))
for it in range(3):
    for i, r in enumerate(R): acc[i, 0] += timed(lambda: r.call("amc_slab_advect"))
    T.alltoall(R)
    for i, r in enumerate(R): acc[i, 1] += timed(lambda: r.call("amc_slab_sort", None))
    for i, r in enumerate(R): acc[i, 2] += timed(lambda: r.call("amc_slab_pairs_begin", C.c_int32(1)))
    T.neighbors(R)
    for i, r in enumerate(R): acc[i, 3] += timed(lambda: r.call("amc_slab_apply", C.c_int32(-1)))
    for g in range(8):
        for i, r in enumerate(R):
            t = timed(lambda: r.call("amc_slab_group", C.c_int32(g)))
            acc[i, 2] += t; grp[i, g] += t
        T.neighbors(R)
        for i, r in enumerate(R): acc[i, 3] += timed(lambda: r.call("amc_slab_apply", C.c_int32(g)))
    for i, r in enumerate(R):
        from argon_monte_carlo_b200 import amc
        st = amc.AmcStepStats()
        acc[i, 4] += timed(lambda: r.call("amc_slab_finish", C.byref(st)))
print("per rank ms/step: advect, sort, pairs(begin+8 groups+pack), apply(9 rounds), finish ; particles")
for i in range(nranks):
    print(i, np.round(acc[i] / 3, 3), sim.particles_per_rank()[i])
sim.close()

print("per rank x colour group: pair kernel + pack (ms)")
print(np.round(grp / 3, 3))
print("sum over groups of the slowest rank: %.3f ms; slowest rank's own sum: %.3f ms" % ((grp / 3).max(axis=0).sum(), (grp / 3).sum(axis=1).max()))
