"""ctypes front end of the CPU oracle (oracle/amc_oracle.c).

TEST INFRASTRUCTURE ONLY -- imported by tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs; never by the product package.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "libamc_oracle.so")
_lib = None

c_double_p = C.POINTER(C.c_double)
c_int64_p = C.POINTER(C.c_int64)


class State(C.Structure):
    _fields_ = [("n", C.c_int64)] + [(k, c_double_p) for k in
                                     ("x", "y", "z", "vx", "vy", "vz", "dist", "dist_x", "dist_y", "dist_z")] + \
               [("flag", C.POINTER(C.c_uint8))] + [(k, c_double_p) for k in ("px", "py", "pz")]


class Paths(C.Structure):
    _fields_ = [("count", C.c_int64), ("cap", C.c_int64)] + [(k, c_double_p) for k in ("total", "cx", "cy", "cz")]


class PairLog(C.Structure):
    _fields_ = [("count", C.c_int64), ("cap", C.c_int64), ("hi", c_int64_p), ("lo", c_int64_p),
                ("group", C.POINTER(C.c_int32)), ("cell", C.POINTER(C.c_int32))]


GEOM_FIELDS = ("argon_mass", "argon_radius", "collision_range", "R_oa", "R_oa_c", "R_p", "R_p_c", "R_g", "R_g_c",
               "H", "oah", "z_cold", "z_gb", "z_gt_pore", "z_gt", "ten_a", "R_oa_sq", "R_g_sq", "R_p_sq",
               "zc3", "zh3", "zgt_m", "zgb_p", "R_g_c_sq", "R_p_c_sq", "recap_lo", "recap_hi",
               "E_cold", "E_hot", "alpha_c", "alpha_g", "cos85")


class Geom(C.Structure):
    _fields_ = [(k, C.c_double) for k in GEOM_FIELDS]


class GridS(C.Structure):
    _fields_ = [("nc", C.c_int32 * 3), ("c0", C.c_int32 * 3), ("edge", c_double_p * 3), ("lo", c_double_p * 3)]


def build(force=False):
    if force or not os.path.isfile(_LIB_PATH) or \
            os.path.getmtime(_LIB_PATH) < os.path.getmtime(os.path.join(_HERE, "amc_oracle.c")):
        subprocess.check_call(["make", "-C", _HERE, "-s", "CC=gcc"])
    return _LIB_PATH


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_LIB_PATH)
        L.orc_pore_walls.restype = C.c_int64
        L.orc_pore_recapture.restype = C.c_int64
        L.orc_temp_case_detect.restype = C.c_int64
        L.orc_temp_case_apply.restype = C.c_int64
        L.orc_temp_recapture.restype = C.c_int64
        L.orc_temp_oob_count.restype = C.c_int64
        L.orc_temp_walls_philox.restype = C.c_int64
        L.orc_pp_groups.restype = C.c_int64
        L.orc_cube_pp_sweep.restype = C.c_int64
        L.orc_cheb_eval.restype = C.c_double
        L.orc_cheb_eval.argtypes = [c_double_p, C.c_int, C.c_double, C.c_double, C.c_double]
        _lib = L
    return _lib


def set_ref_mode(on: bool):
    """True: NumPy-scalar arithmetic of the Python reference (pow / FMA-chain dot);
    False (default): the plain arithmetic the CUDA path implements."""
    lib().orc_set_ref_mode(int(bool(on)))


def _dp(a):
    return a.ctypes.data_as(c_double_p)


STATE_KEYS = ("x", "y", "z", "vx", "vy", "vz", "dist", "dist_x", "dist_y", "dist_z")


class ParticleState:
    """The reference's module-global particle arrays (Pore:385-400) as one object."""

    def __init__(self, x, y, z, vx, vy, vz, dist=None, dist_x=None, dist_y=None, dist_z=None, flag=None):
        n = len(x)
        f = lambda a: np.ascontiguousarray(np.array(a, dtype=np.float64, copy=True))
        z0 = lambda a: f(a) if a is not None else np.zeros(n)
        self.x, self.y, self.z, self.vx, self.vy, self.vz = f(x), f(y), f(z), f(vx), f(vy), f(vz)
        self.dist, self.dist_x, self.dist_y, self.dist_z = z0(dist), z0(dist_x), z0(dist_y), z0(dist_z)
        self.flag = np.ascontiguousarray(np.array(flag, dtype=np.uint8, copy=True)) if flag is not None \
            else np.zeros(n, dtype=np.uint8)
        self.px, self.py, self.pz = np.zeros(n), np.zeros(n), np.zeros(n)
        self.n = n

    def copy(self):
        s = ParticleState(self.x, self.y, self.z, self.vx, self.vy, self.vz, self.dist, self.dist_x, self.dist_y,
                          self.dist_z, self.flag)
        s.px, s.py, s.pz = self.px.copy(), self.py.copy(), self.pz.copy()
        return s

    def c(self):
        st = State()
        st.n = self.n
        for k in STATE_KEYS + ("px", "py", "pz"):
            setattr(st, k, _dp(getattr(self, k)))
        st.flag = self.flag.ctypes.data_as(C.POINTER(C.c_uint8))
        return st

    def arrays(self):
        return {k: getattr(self, k) for k in STATE_KEYS + ("flag",)}


def make_geom(g):
    s = Geom()
    for k in GEOM_FIELDS:
        setattr(s, k, float(getattr(g, k)))
    return s


def make_grid(grid):
    s = GridS()
    keep = []
    for a in range(3):
        s.nc[a], s.c0[a] = grid.nc[a], grid.c0[a]
        e, lo = np.ascontiguousarray(grid.edge[a]), np.ascontiguousarray(grid.lo[a])
        keep += [e, lo]
        s.edge[a], s.lo[a] = _dp(e), _dp(lo)
    s._keep = keep
    return s


class PathSink:
    def __init__(self):
        self.s = Paths()

    def arrays(self):
        n = self.s.count
        if n == 0:
            return tuple(np.zeros(0) for _ in range(4))
        return tuple(np.ctypeslib.as_array(getattr(self.s, k), shape=(n,)).copy() for k in ("total", "cx", "cy", "cz"))

    def __len__(self):
        return self.s.count

    def __del__(self):
        if _lib is not None:
            _lib.orc_paths_free(C.byref(self.s))


class PairSink:
    def __init__(self):
        self.s = PairLog()

    def arrays(self):
        n = self.s.count
        if n == 0:
            return (np.zeros(0, np.int64), np.zeros(0, np.int64), np.zeros(0, np.int32), np.zeros(0, np.int32))
        return tuple(np.ctypeslib.as_array(getattr(self.s, k), shape=(n,)).copy() for k in ("hi", "lo", "group", "cell"))

    def __len__(self):
        return self.s.count

    def __del__(self):
        if _lib is not None:
            _lib.orc_pairlog_free(C.byref(self.s))


# ----------------------------------------------------------------------------- phases
def drift(st: ParticleState, dt, save_prior=True):
    s = st.c()
    lib().orc_drift(C.byref(s), C.c_double(dt), int(save_prior))


def pore_walls(st, geom, sink: PathSink | None = None, want_bits=False):
    s, g = st.c(), make_geom(geom)
    counts = (C.c_int64 * 9)()
    bits = np.zeros(st.n, dtype=np.uint16) if want_bits else None
    errs = lib().orc_pore_walls(C.byref(s), C.byref(g), C.byref(sink.s) if sink is not None else None,
                                bits.ctypes.data_as(C.POINTER(C.c_uint16)) if want_bits else None, counts)
    return np.array(counts[:], dtype=np.int64), int(errs), bits


def pore_recapture(st, geom):
    s, g = st.c(), make_geom(geom)
    return int(lib().orc_pore_recapture(C.byref(s), C.byref(g)))


def temp_case_detect(st, geom, case):
    s, g = st.c(), make_geom(geom)
    n = st.n
    idx = np.zeros(max(n, 1), dtype=np.int64)
    normal = np.zeros(3 * max(n, 1))
    colz = np.zeros(max(n, 1))
    k = int(lib().orc_temp_case_detect(C.byref(s), C.byref(g), int(case), idx.ctypes.data_as(c_int64_p), _dp(normal), _dp(colz)))
    return idx[:k].copy(), normal[:3 * k].reshape(k, 3).copy(), colz[:k].copy()


def temp_case_apply(st, geom, case, idx, dirs=None, surf_e=None, sink: PathSink | None = None):
    s, g = st.c(), make_geom(geom)
    k = len(idx)
    idx = np.ascontiguousarray(idx, dtype=np.int64)
    dirs = np.ascontiguousarray(dirs if dirs is not None else np.zeros((max(k, 1), 3)), dtype=np.float64)
    surf_e = np.ascontiguousarray(surf_e if surf_e is not None else np.zeros(max(k, 1)), dtype=np.float64)
    sums = (C.c_double * 2)()
    errs = lib().orc_temp_case_apply(C.byref(s), C.byref(g), int(case), C.c_int64(k), idx.ctypes.data_as(c_int64_p),
                                     _dp(dirs), _dp(surf_e), C.byref(sink.s) if sink is not None else None, sums)
    return float(sums[0]), float(sums[1]), int(errs)


def temp_recapture(st, geom):
    s, g = st.c(), make_geom(geom)
    return int(lib().orc_temp_recapture(C.byref(s), C.byref(g)))


def temp_oob_count(st, geom):
    s, g = st.c(), make_geom(geom)
    return int(lib().orc_temp_oob_count(C.byref(s), C.byref(g)))


def temp_walls_philox(st, geom, seed, step, cheb, sink: PathSink | None = None, want_bits=False):
    coef, zmid, inv_half = cheb
    coef = np.ascontiguousarray(coef, dtype=np.float64)
    s, g = st.c(), make_geom(geom)
    counts = (C.c_int64 * 10)()
    sums = (C.c_double * 3)()
    bits = np.zeros(st.n, dtype=np.uint16) if want_bits else None
    errs = lib().orc_temp_walls_philox(C.byref(s), C.byref(g), C.c_uint64(seed), C.c_int64(step), _dp(coef),
                                       len(coef), C.c_double(zmid), C.c_double(inv_half),
                                       C.byref(sink.s) if sink is not None else None,
                                       bits.ctypes.data_as(C.POINTER(C.c_uint16)) if want_bits else None, counts, sums)
    return np.array(counts[:], dtype=np.int64), np.array(sums[:]), int(errs), bits


def pp_groups(st, grid, cr, mass, sink: PathSink | None = None, pairs: PairSink | None = None):
    s, g = st.c(), make_grid(grid)
    checks, errs = C.c_int64(0), C.c_int64(0)
    n = lib().orc_pp_groups(C.byref(s), C.byref(g), C.c_double(cr), C.c_double(mass),
                            C.byref(sink.s) if sink is not None else None, C.byref(pairs.s) if pairs is not None else None,
                            C.byref(checks), C.byref(errs))
    return int(n), int(checks.value), int(errs.value)


def cube_walls(st, cube_x, cube_y, cube_z):
    s = st.c()
    counts = (C.c_int64 * 6)()
    lib().orc_cube_walls(C.byref(s), C.c_double(cube_x), C.c_double(cube_y), C.c_double(cube_z), counts)
    return np.array(counts[:], dtype=np.int64)


def cube_pp_sweep(st, grid, cr, mass, sink: PathSink | None = None, pairs: PairSink | None = None):
    s, g = st.c(), make_grid(grid)
    checks, errs = C.c_int64(0), C.c_int64(0)
    n = lib().orc_cube_pp_sweep(C.byref(s), C.byref(g), C.c_double(cr), C.c_double(mass),
                                C.byref(sink.s) if sink is not None else None, C.byref(pairs.s) if pairs is not None else None,
                                C.byref(checks), C.byref(errs))
    return int(n), int(checks.value), int(errs.value)


def histogram(values, nbins=200, first=0.0, last=10 ** -6):
    v = np.ascontiguousarray(values, dtype=np.float64)
    edges = np.linspace(first, last, nbins + 1)
    counts = np.zeros(nbins, dtype=np.int64)
    lib().orc_histogram(_dp(v), C.c_int64(len(v)), int(nbins), C.c_double(first), C.c_double(last), _dp(edges),
                        counts.ctypes.data_as(c_int64_p))
    return counts


def philox_direction(seed, pid, step, case, norm, cos85):
    out = (C.c_double * 3)()
    nv = (C.c_double * 3)(*[float(v) for v in norm])
    lib().orc_philox_direction(C.c_uint64(seed), C.c_int64(pid), C.c_int64(step), int(case), nv, C.c_double(cos85), out)
    return np.array(out[:])


def cheb_eval(cheb, z):
    coef, zmid, inv_half = cheb
    coef = np.ascontiguousarray(coef, dtype=np.float64)
    return float(lib().orc_cheb_eval(_dp(coef), len(coef), zmid, inv_half, float(z)))


def num_threads():
    return int(lib().orc_num_threads())


def set_num_threads(n):
    lib().orc_set_num_threads(int(n))
    return num_threads()
