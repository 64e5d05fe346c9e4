"""Probe the cost of the boundary apply kernel per rank and round (all ranks emulated on one GPU)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, ctypes as C
from argon_monte_carlo_b200 import config, init_state, slab, amc
nranks, total = int(sys.argv[1]), int(sys.argv[2])
cfg = config.pore_config(True, scale=(total / 557649) ** (1 / 3))
ids, *state = init_state.synthetic_pore_chunked(cfg, 17)
_, _, _, zs, _, _, _ = init_state.synthetic_pore_chunk(cfg, 17, 0, 1 << 22)
cuts = slab.balanced_cuts(zs, cfg.grid.edge[2], nranks)
sim = slab.SlabSimulation(cfg, nranks, zs, cuts=cuts, n_total=cfg.num_molecules, seed=17)
sim.set_state(*state); sim.step(3)
T, R = sim.transport, sim.ranks
def timed(fn):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); fn(); e1.record(); torch.cuda.synchronize(); return e0.elapsed_time(e1)
for r in R: r.call("amc_slab_advect")
T.alltoall(R)
for r in R: r.call("amc_slab_sort", None)
for r in R: r.call("amc_slab_pairs_begin", C.c_int32(int(sim.pre_round)))
for g in range(8):
    for r in R: r.call("amc_slab_group", C.c_int32(g))
    T.neighbors(R)
    row = []
    for r in R:
        nu, nd = int(r.bnd_recv_up[0, 0].item()), int(r.bnd_recv_down[0, 0].item())
        t = timed(lambda: r.call("amc_slab_apply", C.c_int32(g)))
        row.append("r%d: up%3d dn%3d %6.1fus" % (r.rank, nu, nd, t * 1e3))
    print("group", g, " | ".join(row))
for r in R:
    st = amc.AmcStepStats(); r.call("amc_slab_finish", C.byref(st))
sim.close()
