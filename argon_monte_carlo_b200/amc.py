"""ctypes binding of libamc.so (include/amc.h) and the host-side mirror of the reference's
per-timestep interface.

The reference keeps its particle state in module globals and advances it with a script-level
loop (Open_Air_Pore_MC.py:416-557, Temperature_Pore_MC.py:662-853, Open_Air_Cube_MC.py:175-338);
`Simulation` owns the same state on one B200 and exposes that loop body as `step()`, plus the
phase-level operators the parity tests compare one by one.  There is no CPU fallback: if the CUDA
library is missing or no device is present, construction raises.
"""
from __future__ import annotations

import ctypes as C
import math
import os

import numpy as np

from . import build as _build
from .config import HIST_RANGE, NUM_BINS

KIND_CUBE, KIND_PORE, KIND_TEMP = 0, 1, 2
PP_GROUPS, PP_SWEEP = 0, 1
RNG_DEVICE, RNG_HOST = 0, 1
TAP_PAIRS, TAP_WALL_BITS, TAP_PATHS = 1, 2, 4
NUM_CASES = 10
ABI_VERSION = 1

GEOM_FIELDS = ("argon_mass", "argon_radius", "collision_range", "R_oa", "R_oa_c", "R_p", "R_p_c", "R_g", "R_g_c",
               "H", "oah", "z_cold", "z_gb", "z_gt_pore", "z_gt", "ten_a", "R_oa_sq", "R_g_sq", "R_p_sq",
               "zc3", "zh3", "zgt_m", "zgb_p", "R_g_c_sq", "R_p_c_sq", "recap_lo", "recap_hi",
               "E_cold", "E_hot", "alpha_c", "alpha_g", "cos85")

c_double_p = C.POINTER(C.c_double)
c_int64_p = C.POINTER(C.c_int64)
c_uint8_p = C.POINTER(C.c_uint8)


class AmcGeom(C.Structure):
    _fields_ = [(k, C.c_double) for k in GEOM_FIELDS]


class AmcConfig(C.Structure):
    _fields_ = [("abi_version", C.c_int32), ("kind", C.c_int32), ("pp_mode", C.c_int32), ("rng_mode", C.c_int32),
                ("taps", C.c_int32), ("dt", C.c_double), ("cube", C.c_double * 3), ("geom", AmcGeom),
                ("nc", C.c_int32 * 3), ("c0", C.c_int32 * 3), ("edge", c_double_p * 3), ("lo", c_double_p * 3),
                ("overlap_sq", C.c_double), ("seed", C.c_uint64), ("cheb_coef", c_double_p), ("cheb_n", C.c_int32),
                ("cheb_zmid", C.c_double), ("cheb_inv_half", C.c_double), ("hist_first", C.c_double),
                ("hist_last", C.c_double), ("hist_edges", c_double_p), ("max_particles", C.c_int64),
                ("pair_capacity", C.c_int64), ("path_capacity", C.c_int64)]


class AmcStepStats(C.Structure):
    _fields_ = [("wall_hits", C.c_int64 * NUM_CASES), ("wall_collisions", C.c_int64), ("pp_collisions", C.c_int64),
                ("pair_checks_ref", C.c_int64), ("pair_checks_exec", C.c_int64), ("oob_after_walls", C.c_int64),
                ("oob_after_pp", C.c_int64), ("oob_after_walls_recapture", C.c_int64),
                ("oob_after_pp_recapture", C.c_int64), ("errors", C.c_int64), ("completed_paths", C.c_int64),
                ("dpz", C.c_double), ("e_cold", C.c_double), ("e_hot", C.c_double)]

    def as_dict(self):
        d = {k: getattr(self, k) for k, _ in self._fields_ if k != "wall_hits"}
        d["wall_hits"] = np.array(self.wall_hits[:], dtype=np.int64)
        d["collisions"] = d["wall_collisions"] + d["pp_collisions"]
        return d


INIT_MAX_REGIONS = 8


class AmcInitSpec(C.Structure):
    """struct amc_init_spec (include/amc.h): synthetic Maxwellian state generated on the device."""
    _fields_ = [("n_total", C.c_int64), ("seed", C.c_uint64), ("n_regions", C.c_int32), ("shape", C.c_int32),
                ("cum_weight", C.c_double * INIT_MAX_REGIONS), ("radius", C.c_double * INIT_MAX_REGIONS),
                ("bx", C.c_double * INIT_MAX_REGIONS), ("by", C.c_double * INIT_MAX_REGIONS),
                ("z_lo", C.c_double * INIT_MAX_REGIONS), ("z_hi", C.c_double * INIT_MAX_REGIONS),
                ("sigma", C.c_double), ("keep_z_lo", C.c_double), ("keep_z_hi", C.c_double)]


class AmcError(RuntimeError):
    pass


_lib = None

# every exported symbol of include/amc.h (tests check the built library against this list)
EXPORTS = ("amc_create", "amc_destroy", "amc_last_error", "amc_abi_version", "amc_set_state", "amc_get_state",
           "amc_num_particles", "amc_step", "amc_drift", "amc_walls", "amc_recapture", "amc_pairs", "amc_wall_case",
           "amc_wall_hits_pending", "amc_wall_apply_directions", "amc_get_histograms", "amc_get_pair_list",
           "amc_get_wall_bits", "amc_get_completed_paths", "amc_clear_taps", "amc_set_step_index", "amc_last_timing",
           "amc_last_detect_ms", "amc_init_synthetic",
           "amc_get_outputs_raw", "amc_set_outputs_raw", "amc_get_step_index",
           "amc_slab_enable", "amc_set_stream", "amc_set_ids", "amc_slab_advect", "amc_slab_sort", "amc_slab_pairs_begin",
           "amc_slab_group", "amc_slab_apply", "amc_slab_finish", "amc_slab_get_owned", "amc_state_digest",
           "amc_slab_p2p_setup", "amc_slab_p2p_connect", "amc_slab_step", "amc_wall_operator", "amc_seed_relax", "amc_last_scatter_ms")


def load_library():
    """dlopen libamc.so (building it in-tree with nvcc first if it is missing)."""
    global _lib
    if _lib is None:
        path = _build.library_path()
        if not os.environ.get("AMC_LIBRARY"):
            try:
                _build.build_library()       # returns at once unless a source is newer than the library
            except Exception:
                if not os.path.isfile(path):
                    raise                    # no nvcc and no library: nothing to run (there is no CPU fallback)
        L = C.CDLL(path)
        L.amc_last_error.restype = C.c_char_p
        L.amc_last_error.argtypes = [C.c_void_p]
        L.amc_num_particles.restype = C.c_int64
        L.amc_num_particles.argtypes = [C.c_void_p]
        L.amc_get_step_index.restype = C.c_int64
        L.amc_get_step_index.argtypes = [C.c_void_p]
        if L.amc_abi_version() != ABI_VERSION:
            raise AmcError("libamc.so ABI version mismatch")
        _lib = L
    return _lib


def overlap_threshold(collision_range: float) -> float:
    """Smallest double t with sqrt(t) >= collision_range, so that the reference's predicate
    ``sqrt(d2) < collision_range`` (Pore:173-174) equals ``d2 < t`` exactly (sqrt is correctly
    rounded and monotone)."""
    cr = float(collision_range)
    t = cr * cr
    while math.sqrt(t) >= cr:
        t = math.nextafter(t, 0.0)
    while math.sqrt(t) < cr:
        t = math.nextafter(t, math.inf)
    return t


def _dp(a):
    return a.ctypes.data_as(c_double_p) if a is not None else None


_M64 = np.uint64(0xFFFFFFFFFFFFFFFF)


def _mix64(z):
    z = z + np.uint64(0x9E3779B97F4A7C15)
    z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
    z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
    return z ^ (z >> np.uint64(31))


def digest_of_arrays(ids, state):
    """NumPy restatement of amc_state_digest (include/amc.h): (sum0, sum1, count) over the given particles.
    ids: global particle indices; state: mapping with the ten float64 arrays and 'flag'."""
    keys = ("x", "y", "z", "vx", "vy", "vz", "dist", "dist_x", "dist_y", "dist_z")
    ids = np.asarray(ids).astype(np.uint64)
    s0 = s1 = np.uint64(0)
    with np.errstate(over="ignore"):
        for f, k in enumerate(keys + ("flag",)):
            if k == "flag":
                bits = (np.asarray(state[k]).astype(np.uint64) & np.uint64(1))
            else:
                bits = np.ascontiguousarray(state[k], dtype=np.float64).view(np.uint64)
            a = _mix64(bits ^ _mix64(np.uint64(16) * ids + np.uint64(f)))
            s0 = s0 + a.sum(dtype=np.uint64)
            s1 = s1 + _mix64(a + np.uint64(0xD6E8FEB86659FD93)).sum(dtype=np.uint64)
    return int(s0), int(s1), int(len(ids))


def combine_digests(parts):
    """Sum of per-rank digests (mod 2^64 per component)."""
    m = (1 << 64) - 1
    return tuple(sum(p[i] for p in parts) & m for i in range(3))


class Simulation:
    """One simulation domain resident on one GPU.

    cfg: a namespace from ``config.cube_config`` / ``config.pore_config``.
    kind / pp_mode default to the script the config belongs to (cube: six planes + serial sweep;
    pore / temp: pore walls + 8 colour groups)."""

    def __init__(self, cfg, *, kind=None, pp_mode=None, rng_mode=RNG_DEVICE, taps=0, device=0, seed=None,
                 max_particles=None, pair_capacity=1 << 20, path_capacity=1 << 22, grid=None, cheb=None):
        self.lib = load_library()
        self.cfg = cfg
        default_kind = {"cube": KIND_CUBE, "pore": KIND_PORE, "temp": KIND_TEMP}[cfg.kind]
        self.kind = default_kind if kind is None else kind
        self.pp_mode = (PP_SWEEP if cfg.kind == "cube" else PP_GROUPS) if pp_mode is None else pp_mode
        self.rng_mode = rng_mode
        grid = grid or cfg.grid
        c = AmcConfig()
        c.abi_version = ABI_VERSION
        c.kind, c.pp_mode, c.rng_mode, c.taps = self.kind, self.pp_mode, rng_mode, taps
        c.dt = float(cfg.dt)
        self._keep = []
        if cfg.kind == "cube":
            c.cube[0], c.cube[1], c.cube[2] = cfg.cube_x, cfg.cube_y, cfg.cube_z
            c.geom.argon_mass, c.geom.argon_radius = cfg.argon_mass, cfg.argon_radius
            c.geom.collision_range = cfg.collision_range
        else:
            for k in GEOM_FIELDS:
                setattr(c.geom, k, float(getattr(cfg.geom, k)))
        for a in range(3):
            c.nc[a], c.c0[a] = grid.nc[a], grid.c0[a]
            e, lo = np.ascontiguousarray(grid.edge[a], dtype=np.float64), np.ascontiguousarray(grid.lo[a], dtype=np.float64)
            self._keep += [e, lo]
            c.edge[a], c.lo[a] = _dp(e), _dp(lo)
        c.overlap_sq = overlap_threshold(cfg.collision_range)
        c.seed = int(cfg.seed if seed is None else seed)
        if self.kind == KIND_TEMP and rng_mode == RNG_DEVICE:
            if cheb is None:
                from .config import gap_energy_chebyshev
                cheb = gap_energy_chebyshev(cfg, 16)
            coef = np.ascontiguousarray(cheb[0], dtype=np.float64)
            self._keep.append(coef)
            c.cheb_coef, c.cheb_n, c.cheb_zmid, c.cheb_inv_half = _dp(coef), len(coef), cheb[1], cheb[2]
        self.cheb = cheb
        edges = np.linspace(HIST_RANGE[0], HIST_RANGE[1], NUM_BINS + 1)
        self._keep.append(edges)
        self.hist_edges = edges
        c.hist_first, c.hist_last, c.hist_edges = float(HIST_RANGE[0]), float(HIST_RANGE[1]), _dp(edges)
        c.max_particles = int(max_particles if max_particles is not None else cfg.num_molecules)
        c.pair_capacity, c.path_capacity = int(pair_capacity), int(path_capacity)
        self._c = c
        self.h = C.c_void_p()
        rc = self.lib.amc_create(C.byref(c), int(device), C.byref(self.h))
        if rc != 0:
            msg = self.lib.amc_last_error(None)
            self.h = None
            raise AmcError("amc_create failed (%d): %s" % (rc, msg.decode() if msg else "?"))
        self.n = 0

    # ------------------------------------------------------------------ plumbing
    def _check(self, rc, what):
        if rc != 0:
            msg = self.lib.amc_last_error(self.h)
            raise AmcError("%s failed (%d): %s" % (what, rc, msg.decode() if msg else "?"))

    def close(self):
        if getattr(self, "h", None):
            self.lib.amc_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    # ------------------------------------------------------------------ state
    def set_state(self, x, y, z, vx, vy, vz, dist=None, dist_x=None, dist_y=None, dist_z=None, flag=None):
        f = lambda a: None if a is None else np.ascontiguousarray(a, dtype=np.float64)
        arrs = [f(a) for a in (x, y, z, vx, vy, vz, dist, dist_x, dist_y, dist_z)]
        fl = None if flag is None else np.ascontiguousarray(np.asarray(flag).astype(np.uint8))
        n = len(arrs[0])
        rc = self.lib.amc_set_state(self.h, C.c_int64(n), *[_dp(a) for a in arrs],
                                    fl.ctypes.data_as(c_uint8_p) if fl is not None else None)
        self._check(rc, "amc_set_state")
        self.n = n

    def init_synthetic(self, spec):
        """Generate the synthetic Maxwellian state described by an AmcInitSpec on the device (init_state.pore_spec /
        cube_spec build one); returns the number of particles this handle holds afterwards."""
        n = C.c_int64(0)
        self._check(self.lib.amc_init_synthetic(self.h, C.byref(spec), C.byref(n)), "amc_init_synthetic")
        self.n = int(n.value)
        return self.n

    def seed_relax(self, max_rounds=8):
        """Overlap-free seeding after init_synthetic (amc_seed_relax): returns (positions re-drawn, particles still
        overlapping a neighbour afterwards)."""
        a, b = C.c_int64(0), C.c_int64(0)
        self._check(self.lib.amc_seed_relax(self.h, C.c_int32(max_rounds), C.byref(a), C.byref(b)), "amc_seed_relax")
        return int(a.value), int(b.value)

    def get_state(self, out=None):
        """dict of the ten float64 arrays and the uint8 flag, original particle index order.
        `out`: optional dict of preallocated (e.g. pinned) arrays to fill."""
        keys = ("x", "y", "z", "vx", "vy", "vz", "dist", "dist_x", "dist_y", "dist_z")
        if out is None:
            out = {k: np.empty(self.n, dtype=np.float64) for k in keys}
            out["flag"] = np.empty(self.n, dtype=np.uint8)
        rc = self.lib.amc_get_state(self.h, *[_dp(out[k]) for k in keys], out["flag"].ctypes.data_as(c_uint8_p))
        self._check(rc, "amc_get_state")
        return out

    # ------------------------------------------------------------------ stepping
    def step(self, n_steps=1):
        """n_steps whole timesteps on the device; returns a list of per-step counter dicts."""
        st = (AmcStepStats * n_steps)()
        self._check(self.lib.amc_step(self.h, C.c_int32(n_steps), st), "amc_step")
        return [s.as_dict() for s in st]

    def step_quiet(self, n_steps=1):
        self._check(self.lib.amc_step(self.h, C.c_int32(n_steps), None), "amc_step")

    def drift(self):
        self._check(self.lib.amc_drift(self.h), "amc_drift")

    def walls(self):
        st = AmcStepStats()
        self._check(self.lib.amc_walls(self.h, C.byref(st)), "amc_walls")
        return st.as_dict()

    def recapture(self, with_after=False):
        cnt, after = C.c_int64(0), C.c_int64(0)
        self._check(self.lib.amc_recapture(self.h, C.byref(cnt), C.byref(after)), "amc_recapture")
        return (int(cnt.value), int(after.value)) if with_after else int(cnt.value)

    def pairs(self):
        st = AmcStepStats()
        self._check(self.lib.amc_pairs(self.h, C.byref(st)), "amc_pairs")
        return st.as_dict()

    def set_step_index(self, step):
        self._check(self.lib.amc_set_step_index(self.h, C.c_int64(step)), "amc_set_step_index")

    # ------------------------------------------------------------------ host-RNG parity mode
    def wall_case(self, case):
        n = C.c_int64(0)
        self._check(self.lib.amc_wall_case(self.h, int(case), C.byref(n)), "amc_wall_case")
        return int(n.value)

    def wall_hits_pending(self, case, cap=None):
        cap = int(cap or max(self.n, 1))
        idx = np.zeros(cap, dtype=np.int64)
        nrm = np.zeros(3 * cap)
        colz = np.zeros(cap)
        n = C.c_int64(0)
        rc = self.lib.amc_wall_hits_pending(self.h, int(case), C.c_int64(cap), C.byref(n), idx.ctypes.data_as(c_int64_p),
                                            _dp(nrm), _dp(colz))
        self._check(rc, "amc_wall_hits_pending")
        k = int(n.value)
        return idx[:k].copy(), nrm[:3 * k].reshape(k, 3).copy(), colz[:k].copy()

    def wall_apply_directions(self, case, idx, dirs, surf_e=None):
        k = len(idx)
        idx = np.ascontiguousarray(idx, dtype=np.int64)
        dirs = np.ascontiguousarray(dirs, dtype=np.float64).reshape(-1) if k else np.zeros(3)
        se = np.ascontiguousarray(surf_e, dtype=np.float64) if surf_e is not None else None
        dpz, de = np.zeros(max(k, 1)), np.zeros(max(k, 1))
        errs = C.c_int64(0)
        rc = self.lib.amc_wall_apply_directions(self.h, int(case), C.c_int64(k), idx.ctypes.data_as(c_int64_p), _dp(dirs),
                                                _dp(se), _dp(dpz), _dp(de), C.byref(errs))
        self._check(rc, "amc_wall_apply_directions")
        return dpz[:k], de[:k], int(errs.value)

    def step_host_rng(self, surface_energy_gap=None):
        """One Temperature_Pore_MC timestep with the reference's host Mersenne-Twister draws
        (Temp:662-853): walls case by case, directions drawn on the host in the reference's order,
        per-hit contributions summed sequentially like the reference's mpf accumulators."""
        from . import host_rng
        from .config import surface_energy_gap as seg
        cfg = self.cfg
        self.drift()
        dpz = e_hot = e_cold = 0
        counts = np.zeros(NUM_CASES, dtype=np.int64)
        errors = 0
        for case in range(NUM_CASES):
            if case <= 2:
                counts[case] = self.wall_case(case)
                continue
            idx, nrm, colz = self.wall_hits_pending(case)
            counts[case] = len(idx)
            surf = None
            if case == 5:
                dirs = np.zeros((len(idx), 3))
                surf = np.zeros(len(idx))
                for k in range(len(idx)):
                    if nrm[k, 0] == nrm[k, 0]:
                        dirs[k] = host_rng.inbound_direction(nrm[k])
                        surf[k] = float((surface_energy_gap or (lambda z: seg(cfg, z)))(colz[k]))
            else:
                dirs = host_rng.directions_for_hits(nrm)
            p, e, er = self.wall_apply_directions(case, idx, dirs, surf)
            errors += er
            sp = se = 0
            for k in range(len(idx)):      # sequential sums, ascending particle index (Temp:385,389)
                sp = sp + p[k]
                se = se + e[k]
            dpz = dpz + sp
            if case in (3, 7, 9):
                e_cold = e_cold + se
            elif case in (4, 6, 8):
                e_hot = e_hot + se
        oob_walls, oob_walls_after = self.recapture(True)
        st = self.pairs()
        oob_pp, oob_pp_after = self.recapture(True)
        wall_hits = int(counts[3:].sum())
        st.update(wall_hits=counts, wall_collisions=wall_hits, oob_after_walls=oob_walls, oob_after_pp=oob_pp,
                  oob_after_walls_recapture=oob_walls_after, oob_after_pp_recapture=oob_pp_after,
                  dpz=dpz, e_cold=e_cold, e_hot=e_hot, collisions=wall_hits + st["pp_collisions"],
                  errors=st["errors"] + errors)
        return st

    # ------------------------------------------------------------------ outputs
    def histograms(self):
        """(counts[4, 200] uint64 in the order total/x/y/z, number of completed paths, sums[4])."""
        counts = np.zeros((4, NUM_BINS), dtype=np.uint64)
        n = C.c_uint64(0)
        sums = np.zeros(4)
        rc = self.lib.amc_get_histograms(self.h, counts.ctypes.data_as(C.POINTER(C.c_uint64)), C.byref(n), _dp(sums))
        self._check(rc, "amc_get_histograms")
        return counts, int(n.value), sums

    def pair_list(self, cap=1 << 20):
        hi, lo = np.zeros(cap, dtype=np.int64), np.zeros(cap, dtype=np.int64)
        grp, cell = np.zeros(cap, dtype=np.int32), np.zeros(cap, dtype=np.int32)
        n = C.c_int64(0)
        rc = self.lib.amc_get_pair_list(self.h, C.c_int64(cap), C.byref(n), hi.ctypes.data_as(c_int64_p),
                                        lo.ctypes.data_as(c_int64_p), grp.ctypes.data_as(C.POINTER(C.c_int32)),
                                        cell.ctypes.data_as(C.POINTER(C.c_int32)))
        self._check(rc, "amc_get_pair_list")
        k = int(n.value)
        return hi[:k].copy(), lo[:k].copy(), grp[:k].copy(), cell[:k].copy()

    def wall_bits(self):
        bits = np.zeros(max(self.n, 1), dtype=np.uint16)
        self._check(self.lib.amc_get_wall_bits(self.h, bits.ctypes.data_as(C.POINTER(C.c_uint16))), "amc_get_wall_bits")
        return bits[:self.n]

    def completed_paths(self, cap=1 << 22):
        arrs = [np.zeros(cap) for _ in range(4)]
        n = C.c_int64(0)
        rc = self.lib.amc_get_completed_paths(self.h, C.c_int64(cap), C.byref(n), *[_dp(a) for a in arrs])
        self._check(rc, "amc_get_completed_paths")
        k = int(n.value)
        return tuple(a[:k].copy() for a in arrs)

    # ------------------------------------------------------------------ checkpoint / resume
    def checkpoint(self, path=None):
        """Everything needed to continue the run exactly: particle state, histograms, free-path sums
        (raw fixed-point limbs) and the step index that keys the device RNG.  Returns the dict and, if
        `path` is given, writes it with np.savez."""
        st = self.get_state()
        counts = np.zeros((4, NUM_BINS), dtype=np.uint64)
        n = C.c_uint64(0)
        limbs = np.zeros(8, dtype=np.uint64)
        rc = self.lib.amc_get_outputs_raw(self.h, counts.ctypes.data_as(C.POINTER(C.c_uint64)), C.byref(n),
                                          limbs.ctypes.data_as(C.POINTER(C.c_uint64)))
        self._check(rc, "amc_get_outputs_raw")
        st.update(hist_counts=counts, n_paths=np.uint64(n.value), path_limbs=limbs,
                  step_index=np.int64(self.lib.amc_get_step_index(self.h)))
        if self.rng_mode == RNG_HOST:
            # parity mode draws from the two global Mersenne-Twister streams (host_rng.py): part of the state
            import pickle
            import random
            st["host_rng"] = np.frombuffer(pickle.dumps((np.random.get_state(), random.getstate())), dtype=np.uint8)
        if path is not None:
            np.savez(path, **st)
        return st

    def restore(self, ck):
        """ck: dict from checkpoint() or the path of an .npz written by it."""
        if isinstance(ck, (str, bytes)) or hasattr(ck, "__fspath__"):
            ck = dict(np.load(ck))
        self.set_state(*[ck[k] for k in ("x", "y", "z", "vx", "vy", "vz", "dist", "dist_x", "dist_y", "dist_z")], flag=ck["flag"])
        counts = np.ascontiguousarray(ck["hist_counts"], dtype=np.uint64)
        limbs = np.ascontiguousarray(ck["path_limbs"], dtype=np.uint64)
        rc = self.lib.amc_set_outputs_raw(self.h, counts.ctypes.data_as(C.POINTER(C.c_uint64)), C.c_uint64(int(ck["n_paths"])),
                                          limbs.ctypes.data_as(C.POINTER(C.c_uint64)))
        self._check(rc, "amc_set_outputs_raw")
        self.set_step_index(int(ck["step_index"]))
        if self.rng_mode == RNG_HOST:
            if "host_rng" not in ck:
                raise AmcError("checkpoint of a host-RNG run without the Mersenne-Twister states: exact resume impossible")
            import pickle
            import random
            np_state, py_state = pickle.loads(np.asarray(ck["host_rng"], dtype=np.uint8).tobytes())
            np.random.set_state(np_state)
            random.setstate(py_state)

    def state_digest(self):
        """(sum0, sum1, count): order-independent checksum of the owned particle state, computed on the device
        (amc_state_digest); equals digest_of_arrays(arange(n), get_state()) for a single-domain handle."""
        out = (C.c_uint64 * 3)()
        self._check(self.lib.amc_state_digest(self.h, out), "amc_state_digest")
        return int(out[0]), int(out[1]), int(out[2])

    def clear_taps(self):
        self._check(self.lib.amc_clear_taps(self.h), "amc_clear_taps")

    def last_timing(self):
        """(ms[5], launches) of the last step() call: advect+walls, cell sort, pair kernels,
        recapture, total -- CUDA events on the handle's stream."""
        ms = (C.c_double * 5)()
        n = C.c_int64(0)
        self._check(self.lib.amc_last_timing(self.h, ms, C.byref(n)), "amc_last_timing")
        return list(ms), int(n.value)

    def last_scatter_ms(self):
        """Device time (ms) of k_scatter_advect alone, summed over the steps of the last step() call."""
        v = C.c_double(0.0)
        self._check(self.lib.amc_last_scatter_ms(self.h, C.byref(v)), "amc_last_scatter_ms")
        return float(v.value)

    def last_detect_ms(self):
        """Device time (ms) of the detection kernel alone, summed over the steps of the last step() call."""
        v = C.c_double(0.0)
        self._check(self.lib.amc_last_detect_ms(self.h, C.byref(v)), "amc_last_detect_ms")
        return float(v.value)
