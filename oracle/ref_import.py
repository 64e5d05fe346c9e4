"""Import the upstream Python reference (read-only at /root/reference) as modules.

Only used in the build container to validate the oracle and to generate the committed golden
vectors (tests/golden/); /root/reference does not exist on the GPU box.  Open_Air_Cube_MC.py has
no __main__ guard (importing it runs all 500 steps), so only the two pore scripts are imported;
the cube script is run as a subprocess by make_golden.py when needed."""
import importlib.util
import os
import sys

REFERENCE_DIR = os.environ.get("AMC_REFERENCE_DIR", "/root/reference")
_STUBS = os.path.join(os.path.dirname(os.path.abspath(__file__)), "stubs")


def available():
    return os.path.isfile(os.path.join(REFERENCE_DIR, "Open_Air_Pore_MC.py"))


def load(name):
    """name: 'Open_Air_Pore_MC' or 'Temperature_Pore_MC'. Returns a fresh module object."""
    if not available():
        raise FileNotFoundError(REFERENCE_DIR)
    for p in (_STUBS, REFERENCE_DIR):
        if p not in sys.path:
            sys.path.insert(0, p)
    spec = importlib.util.spec_from_file_location("_amc_ref_" + name, os.path.join(REFERENCE_DIR, name + ".py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod
