"""Initial particle state ("L1" of the reference) on the host.

Reproduces the reference's initial positions and velocities *bit for bit*, which means consuming
the two global Mersenne-Twister streams (``np.random`` and ``random``) in the reference's order:
  Open_Air_Pore_MC.py:106-158 / Temperature_Pore_MC.py:154-213 (pore), Open_Air_Cube_MC.py:145-172.
The per-particle Python loop of ``init_velocities`` is replaced by block draws of the same raw
32-bit words (one ``uniform`` = two words, one ``choice([-1, 1])`` = one word masked to a bit),
after which the global ``np.random`` state is advanced to where the loop would have left it, so
later host draws (the energized walls' parity mode) continue the same stream.

Also provides the synthetic Maxwellian states of BASELINE.json configs 4-5 (not reference
states: uniform positions per region, Gaussian velocity components).
"""
from __future__ import annotations

import math
import random as _pyrandom

import numpy as np
from scipy.stats import maxwell


def _isotropic_components(speeds: np.ndarray):
    """random_components(v) for every v (Pore:97-104), same stream order, vectorised."""
    n = len(speeds)
    st = np.random.get_state()
    bg = np.random.MT19937()
    bg.state = {"bit_generator": "MT19937", "state": {"key": st[1], "pos": st[2]}}
    raw = bg.random_raw(3 * n).reshape(n, 3)
    a, b = raw[:, 0] >> np.uint64(5), raw[:, 1] >> np.uint64(6)
    u = (a.astype(np.float64) * 67108864.0 + b.astype(np.float64)) / 9007199254740992.0
    costheta = -1.0 + 2.0 * u                       # np.random.uniform(low=-1.0, high=1.0)
    sign = np.where((raw[:, 2] & np.uint64(1)) == 1, 1.0, -1.0)  # np.random.choice([-1, 1])
    s2 = bg.state["state"]
    np.random.set_state(("MT19937", s2["key"], s2["pos"], st[3], st[4]))
    phi = [_pyrandom.uniform(0, math.pi) for _ in range(n)]
    theta = [math.acos(c) for c in costheta]
    cphi = np.array([math.cos(p) for p in phi])
    sphi = np.array([math.sin(p) for p in phi])
    sth = np.array([math.sin(t) for t in theta])
    cth = np.array([math.cos(t) for t in theta])
    fx = speeds * cphi * sth
    fy = speeds * sphi * sth * sign
    fz = speeds * cth
    return fx, fy, fz


def _maxwell_velocities(cfg):
    speeds = maxwell.rvs(loc=0, scale=cfg.a_shape, size=cfg.num_molecules)
    return _isotropic_components(np.asarray(speeds, dtype=np.float64))


def pore_initial_state(cfg, seed: bool = True):
    """(x, y, z, vx, vy, vz) of the pore scripts. Seeds both generators with cfg.seed (Pore:89-90)."""
    if seed:
        np.random.seed(cfg.seed)
        _pyrandom.seed(cfg.seed)
    n, a = cfg.num_molecules, cfg.argon_radius
    theta = np.random.uniform(0, 2 * np.pi, n)
    rand_radius = np.random.uniform(0, 1, n)
    cos_t = np.array([math.cos(t) for t in theta])
    sin_t = np.array([math.sin(t) for t in theta])
    root = np.sqrt(rand_radius)
    oa, hot, gap, cold = cfg.open_air_particles, cfg.hot_pore_particles, cfg.gap_particles, cfg.cold_pore_particles
    bounds = np.cumsum([0, oa, hot, gap, cold])
    radii = [cfg.open_air_radius - a, cfg.pore_coated_radius - a, cfg.gap_radius - a,
             cfg.pore_coated_radius - a, cfg.open_air_radius - a]
    h = cfg.open_air_height
    z_ranges = [
        (0 + a, h - a),
        (h, h + cfg.hot_coating_height),
        (h + cfg.hot_coating_height + a, h + cfg.hot_coating_height + cfg.gap_height - a),
        (h + cfg.hot_coating_height + cfg.gap_height,
         h + cfg.hot_coating_height + cfg.gap_height + cfg.cold_coating_height),
        (h + cfg.hot_coating_height + cfg.gap_height + cfg.cold_coating_height + a, cfg.total_height - a),
    ]
    x, y, z = np.zeros(n), np.zeros(n), np.zeros(n)
    for r in range(5):
        lo = bounds[r]
        hi = bounds[r + 1] if r < 4 else n
        x[lo:hi] = radii[r] * root[lo:hi] * cos_t[lo:hi]
        y[lo:hi] = radii[r] * root[lo:hi] * sin_t[lo:hi]
        z[lo:hi] = np.random.uniform(z_ranges[r][0], z_ranges[r][1], hi - lo)
    vx, vy, vz = _maxwell_velocities(cfg)
    return x, y, z, vx, vy, vz


def cube_initial_state(cfg, seed: bool = True):
    """Open_Air_Cube_MC.py:145-172: `remaining_particles` uniform in the cube, then
    `min_num_particles_per_cell` stratified into each of the n_sub^3 cells (x-major, z-minor)."""
    if seed:
        np.random.seed(cfg.seed)
        _pyrandom.seed(cfg.seed)
    ns, m = cfg.num_x_subdivions, cfg.min_num_particles_per_cell
    x0 = cfg.cube_x * np.random.random(cfg.remaining_particles)
    y0 = cfg.cube_y * np.random.random(cfg.remaining_particles)
    z0 = cfg.cube_z * np.random.random(cfg.remaining_particles)
    r = np.random.random(ns * ns * ns * 3 * m).reshape(ns, ns, ns, 3, m)
    ix = np.arange(ns).reshape(ns, 1, 1, 1)
    iy = np.arange(ns).reshape(1, ns, 1, 1)
    iz = np.arange(ns).reshape(1, 1, ns, 1)
    xv = cfg.dx * (r[:, :, :, 0, :] + ix)
    yv = cfg.dy * (r[:, :, :, 1, :] + iy)
    zv = cfg.dz * (r[:, :, :, 2, :] + iz)
    x = np.concatenate((x0, xv.ravel()))
    y = np.concatenate((y0, yv.ravel()))
    z = np.concatenate((z0, zv.ravel()))
    vx, vy, vz = _maxwell_velocities(cfg)
    return x, y, z, vx, vy, vz


# ----------------------------------------------------------------------------- synthetic (configs 4-5)
def synthetic_pore_state(cfg, seed: int = 17, n: int | None = None):
    """Maxwellian argon in the (possibly scaled) pore geometry: region populations as in
    Pore:79-83, positions uniform inside each region shrunk by one argon radius at the walls it
    touches, velocity components N(0, a_shape^2).  NumPy Generator(PCG64) -- not the reference
    streams; use for throughput runs only."""
    rng = np.random.default_rng(seed)
    n = cfg.num_molecules if n is None else int(n)
    a = cfg.argon_radius
    vols = np.array([cfg.open_air_volume, cfg.hot_volume, cfg.gap_volume, cfg.cold_volume, cfg.open_air_volume])
    counts = np.floor(n * vols / vols.sum()).astype(np.int64)
    counts[4] += n - counts.sum()
    h = cfg.open_air_height
    radii = [cfg.open_air_radius - a, cfg.pore_coated_radius - a, cfg.gap_radius - a,
             cfg.pore_coated_radius - a, cfg.open_air_radius - a]
    z_ranges = [(a, h - a), (h, cfg.gap_bottom_height), (cfg.gap_bottom_height + a, cfg.gap_top_height - a),
                (cfg.gap_top_height, cfg.total_height - h), (cfg.total_height - h + a, cfg.total_height - a)]
    x, y, z = np.empty(n), np.empty(n), np.empty(n)
    o = 0
    for r in range(5):
        m = int(counts[r])
        th = rng.uniform(0, 2 * np.pi, m)
        rr = radii[r] * np.sqrt(rng.uniform(0, 1, m))
        x[o:o + m], y[o:o + m] = rr * np.cos(th), rr * np.sin(th)
        z[o:o + m] = rng.uniform(z_ranges[r][0], z_ranges[r][1], m)
        o += m
    v = rng.normal(0.0, cfg.a_shape, (3, n))
    return x, y, z, v[0].copy(), v[1].copy(), v[2].copy()


def synthetic_cube_state(cfg, n: int, seed: int = 127):
    """Config 4: n particles uniform in the cube, Gaussian velocity components."""
    rng = np.random.default_rng(seed)
    x = rng.uniform(0, cfg.cube_x, n)
    y = rng.uniform(0, cfg.cube_y, n)
    z = rng.uniform(0, cfg.cube_z, n)
    v = rng.normal(0.0, cfg.a_shape, (3, n))
    return x, y, z, v[0].copy(), v[1].copy(), v[2].copy()


def synthetic_pore_chunk(cfg, seed: int, chunk_index: int, chunk_size: int):
    """Chunk `chunk_index` of a synthetic Maxwellian pore state that is defined chunk by chunk (each
    chunk has its own Generator seeded by (seed, chunk_index) and draws every particle's region at
    random with the region volumes as weights), so any process can generate any part of a very large
    state without holding the rest.  Returns (first global id, x, y, z, vx, vy, vz)."""
    n_total = cfg.num_molecules
    start = chunk_index * chunk_size
    m = max(0, min(chunk_size, n_total - start))
    rng = np.random.default_rng([seed, chunk_index])
    a = cfg.argon_radius
    vols = np.array([cfg.open_air_volume, cfg.hot_volume, cfg.gap_volume, cfg.cold_volume, cfg.open_air_volume])
    region = rng.choice(5, size=m, p=vols / vols.sum())
    h = cfg.open_air_height
    radii = np.array([cfg.open_air_radius - a, cfg.pore_coated_radius - a, cfg.gap_radius - a,
                      cfg.pore_coated_radius - a, cfg.open_air_radius - a])
    zlo = np.array([a, h, cfg.gap_bottom_height + a, cfg.gap_top_height, cfg.total_height - h + a])
    zhi = np.array([h - a, cfg.gap_bottom_height, cfg.gap_top_height - a, cfg.total_height - h, cfg.total_height - a])
    th = rng.uniform(0, 2 * np.pi, m)
    rr = radii[region] * np.sqrt(rng.uniform(0, 1, m))
    x, y = rr * np.cos(th), rr * np.sin(th)
    z = zlo[region] + (zhi[region] - zlo[region]) * rng.uniform(0, 1, m)
    v = rng.normal(0.0, cfg.a_shape, (3, m))
    return start, x, y, z, v[0].copy(), v[1].copy(), v[2].copy()


def synthetic_pore_chunked(cfg, seed: int = 17, chunk_size: int = 1 << 22, keep=None):
    """Concatenate all chunks (optionally only the particles for which keep(z) is True).
    Returns (global ids, x, y, z, vx, vy, vz)."""
    nchunks = (cfg.num_molecules + chunk_size - 1) // chunk_size
    parts = []
    for c in range(nchunks):
        start, *arrs = synthetic_pore_chunk(cfg, seed, c, chunk_size)
        ids = start + np.arange(len(arrs[0]), dtype=np.int64)
        if keep is not None:
            m = keep(arrs[2])
            ids, arrs = ids[m], [a[m] for a in arrs]
        parts.append((ids, *arrs))
    return tuple(np.concatenate([p[i] for p in parts]) for i in range(7))


# ---------------------------------------------------------------- device-side generation (amc_init_synthetic)
def _spec(n_total, seed, shape, weights, radius, bx, by, z_lo, z_hi, sigma, keep_z=None):
    from .amc import AmcInitSpec, INIT_MAX_REGIONS
    sp = AmcInitSpec()
    k = len(weights)
    if not 1 <= k <= INIT_MAX_REGIONS:
        raise ValueError("1..%d regions" % INIT_MAX_REGIONS)
    cum = np.cumsum(np.asarray(weights, dtype=np.float64) / np.sum(weights))
    cum[-1] = 1.0
    sp.n_total, sp.seed, sp.n_regions, sp.shape, sp.sigma = int(n_total), int(seed), k, int(shape), float(sigma)
    for r in range(k):
        sp.cum_weight[r] = cum[r]
        sp.radius[r], sp.bx[r], sp.by[r] = float(radius[r]), float(bx[r]), float(by[r])
        sp.z_lo[r], sp.z_hi[r] = float(z_lo[r]), float(z_hi[r])
    sp.keep_z_lo, sp.keep_z_hi = (-np.inf, np.inf) if keep_z is None else (float(keep_z[0]), float(keep_z[1]))
    return sp


def pore_spec(cfg, seed: int = 17, keep_z=None):
    """The synthetic Maxwellian pore of synthetic_pore_chunk (same regions, weights and margins) as an
    AmcInitSpec for Simulation.init_synthetic; keep_z = (z_lo, z_hi): only that slab."""
    a, h = cfg.argon_radius, cfg.open_air_height
    vols = [cfg.open_air_volume, cfg.hot_volume, cfg.gap_volume, cfg.cold_volume, cfg.open_air_volume]
    radii = [cfg.open_air_radius - a, cfg.pore_coated_radius - a, cfg.gap_radius - a, cfg.pore_coated_radius - a,
             cfg.open_air_radius - a]
    zlo = [a, h, cfg.gap_bottom_height + a, cfg.gap_top_height, cfg.total_height - h + a]
    zhi = [h - a, cfg.gap_bottom_height, cfg.gap_top_height - a, cfg.total_height - h, cfg.total_height - a]
    return _spec(cfg.num_molecules, seed, 0, vols, radii, [0] * 5, [0] * 5, zlo, zhi, cfg.a_shape, keep_z)


def cube_spec(cfg, n: int, seed: int = 127, keep_z=None):
    """n particles uniform in the cube, Gaussian velocity components (config 4)."""
    return _spec(n, seed, 1, [1.0], [0.0], [cfg.cube_x], [cfg.cube_y], [0.0], [cfg.cube_z], cfg.a_shape, keep_z)
