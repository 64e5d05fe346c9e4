import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def oracle():
    """The CPU oracle (test infrastructure). Plain arithmetic mode by default."""
    from oracle import oracle as O
    O.lib()
    O.set_ref_mode(False)
    return O


@pytest.fixture(scope="session")
def pore_cfg():
    from argon_monte_carlo_b200 import config
    return config.pore_config(False)


@pytest.fixture(scope="session")
def temp_cfg():
    from argon_monte_carlo_b200 import config
    return config.pore_config(True)


@pytest.fixture(scope="session")
def cube_cfg():
    from argon_monte_carlo_b200 import config
    return config.cube_config()


@pytest.fixture(scope="session")
def pore_init(pore_cfg):
    from argon_monte_carlo_b200 import init_state
    return init_state.pore_initial_state(pore_cfg)


@pytest.fixture(scope="session")
def temp_init(temp_cfg):
    from argon_monte_carlo_b200 import init_state
    return init_state.pore_initial_state(temp_cfg)


@pytest.fixture(scope="session")
def cube_init(cube_cfg):
    from argon_monte_carlo_b200 import init_state
    return init_state.cube_initial_state(cube_cfg)
