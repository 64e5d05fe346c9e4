"""argon_monte_carlo_b200 -- B200 (sm_100a) implementation of Argon_Monte_Carlo's per-timestep
hard-sphere collision loop behind a C ABI (include/amc.h), with the host-side pieces the three
drop-in driver scripts need (constants, initial state, host RNG parity mode, output writers)."""
from . import config  # noqa: F401

__all__ = ["config", "amc", "init_state", "host_rng", "build"]
