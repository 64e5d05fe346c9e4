"""Where does a cell visit of the pair kernel spend its cycles?  Builds a debug library with
-DAMC_PHASE_CLOCK and prints thread 0's average cycles per phase.
python tools/phase_clocks.py {pore_ref|temp_scaled} [PARTICLES]"""
import ctypes as C, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from argon_monte_carlo_b200 import build
dbg = os.path.join(ROOT, "argon_monte_carlo_b200", "libamc_phaseclock.so")
if not os.path.isfile(dbg) or os.environ.get("REBUILD"):
    build.build_library(force=True, defines=("AMC_PHASE_CLOCK",), output=dbg)
os.environ["AMC_LIBRARY"] = dbg
from argon_monte_carlo_b200 import amc, config, init_state
kind = sys.argv[1]
if kind == "pore_ref":
    cfg = config.pore_config(False); state = init_state.pore_initial_state(cfg)
else:
    n = int(sys.argv[2]) if len(sys.argv) > 2 else 12_500_000
    cfg = config.pore_config(True, scale=(n / 557649) ** (1 / 3)); state = init_state.synthetic_pore_state(cfg, seed=17)
sim = amc.Simulation(cfg, max_particles=len(state[0])); sim.set_state(*state)
sim.step(4)
out = (C.c_ulonglong * 16)()
sim.lib.amc_debug_phase_clocks(out, 1)
sim.step(4)
sim.lib.amc_debug_phase_clocks(out, 0)
ms, _ = sim.last_timing()
names = ["header", "gather:barrier", "chain:barrier", "-", "-", "walk:barrier+zero", "pick", "retest+tail", "resolve_pair", "chain:own", "walk:own", "gather:issue", "gather:wait+store"]
visits = out[15]
print(kind, len(state[0]), "visits/step", visits / 4, "pairs ms/step", ms[2] / 4)
tot = 0
for k, nm in enumerate(names):
    if nm != "-":
        print(f"  {nm:14s} {out[k] / max(visits, 1):9.0f} cycles/visit"); tot += out[k] / max(visits, 1)
print(f"  {'sum':14s} {tot:9.0f} cycles/visit = {tot / 1.965e3:.1f} us at 1965 MHz")
