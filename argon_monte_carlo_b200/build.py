"""In-tree build of libamc.so with nvcc for sm_100a (the .so travels to the GPU box with the repo
snapshot; it is git-ignored).  -fmad=false: the reference never fuses a multiply with an add and
bit-exact flags / pair sets depend on that."""
from __future__ import annotations

import os
import shutil
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
_CSRC = os.path.join(_HERE, "csrc")
_INCLUDE = os.path.join(os.path.dirname(_HERE), "include")
SOURCES = ("amc_api.cu",)
HEADERS = ("amc_kernels.cuh", "amc_device.cuh")
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-fmad=false", "-std=c++17",
              "-shared", "-Xcompiler", "-fPIC"]


def library_path() -> str:
    return os.environ.get("AMC_LIBRARY") or os.path.join(_HERE, "libamc.so")


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.isfile(cand):
            return cand
    raise RuntimeError("nvcc not found: libamc.so cannot be built (there is no CPU fallback)")


def is_stale() -> bool:
    lib = library_path()
    if not os.path.isfile(lib):
        return True
    t = os.path.getmtime(lib)
    deps = [os.path.join(_CSRC, f) for f in SOURCES + HEADERS] + [os.path.join(_INCLUDE, "amc.h")]
    return any(os.path.getmtime(d) > t for d in deps)


def build_library(force: bool = False, verbose: bool = False, defines=(), output=None) -> str:
    lib = output or library_path()
    if not force and not is_stale() and not output:
        return lib
    env = dict(os.environ)
    env.pop("CC", None)   # the image exports a CC that is not a usable nvcc host compiler
    env.pop("CXX", None)
    cmd = [_nvcc(), *NVCC_FLAGS, *["-D" + d for d in defines], "-I", _INCLUDE, "-o", lib] + [os.path.join(_CSRC, f) for f in SOURCES]
    if verbose:
        cmd.insert(1, "-Xptxas")
        cmd.insert(2, "-v")
    subprocess.check_call(cmd, env=env)
    return lib


if __name__ == "__main__":
    print(build_library(force=True, verbose=True))
