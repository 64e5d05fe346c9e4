"""Small fixed workload for ncu captures: python tools/profile_target.py {pore_ref|temp_scaled|slab1} STEPS [PARTICLES]
slab1 = the energized pore through the device-resident slab path (amc_slab_step) with a single rank: the kernels of
the multi-GPU step without the peer-to-peer transfers (ncu cannot wrap a multi-rank command)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from argon_monte_carlo_b200 import amc, config, init_state

kind, steps = sys.argv[1], int(sys.argv[2])
if kind == "pore_ref":
    cfg = config.pore_config(False)
    state = init_state.pore_initial_state(cfg)
else:
    n = int(sys.argv[3]) if len(sys.argv) > 3 else 12_500_000
    cfg = config.pore_config(True, scale=(n / 557649) ** (1 / 3))
    state = init_state.synthetic_pore_state(cfg, seed=17)
if kind == "slab1":
    from argon_monte_carlo_b200 import slab
    ss = slab.SlabSimulation(cfg, 1, state[2], seed=17, p2p=True)
    ss.set_state(*state)
    for k in range(steps):
        st = ss.step_fused(1)[0]
    print(kind, len(state[0]), "phase ms", ss.phase_ms, "collisions", st["collisions"])
    ss.close()
    sys.exit(0)
sim = amc.Simulation(cfg, max_particles=len(state[0]))
sim.set_state(*state)
for k in range(steps):
    st = sim.step(1)[0]
ms, launches = sim.last_timing()
print(kind, len(state[0]), "detect ms", sim.last_detect_ms(), "last step ms", ms, "collisions", st["collisions"], "checks exec/ref", st["pair_checks_exec"], st["pair_checks_ref"])
sim.close()
