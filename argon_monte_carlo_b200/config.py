"""Host-side constants ("L0" of the reference) for the three simulation stages.

The reference has no config system: every script starts with a block of module-level
assignments (Open_Air_Cube_MC.py:26-78, Open_Air_Pore_MC.py:25-94,
Temperature_Pore_MC.py:30-113).  The values below are evaluated with the *same Python
expressions* so that every threshold that decides a wall-hit flag or a cell membership is the
same IEEE double the reference compares against (e.g. ``open_air_height + hot_coating_height``
is 1.3000000000000003e-07, not 1.3e-07).  Nothing here touches the GPU.

``scale`` multiplies every length of the pore geometry (synthetic configs 4/5 of
BASELINE.json); ``scale=1`` gives the reference values bit for bit (multiplying by 1.0 is exact).
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from types import SimpleNamespace

import numpy as np

NUM_BINS = 200                      # Pore:93
HIST_RANGE = (0, 10 ** -6)          # Pore:575


def cylinder_volume(radius, height):
    """utils.cylinder_volume (utils.py:3-4)."""
    return np.pi * radius ** 2 * height


@dataclass
class Grid:
    """Collision-cell grid. ``edge[a][k] = (c0+k)*d`` and ``lo[a][k] = edge[a][k] - band``
    are the two sides of the reference's strict layer masks (Pore:527-529, Cube:233-237)."""
    nc: tuple
    c0: tuple
    d: tuple
    band: tuple
    edge: list = field(default_factory=list)
    lo: list = field(default_factory=list)

    def __post_init__(self):
        self.edge, self.lo = [], []
        for a in range(3):
            ks = [self.c0[a] + k for k in range(self.nc[a] + 1)]
            e = np.array([k * self.d[a] for k in ks], dtype=np.float64)
            lo = np.array([k * self.d[a] - self.band[a] for k in ks[:-1]], dtype=np.float64)
            self.edge.append(e)
            self.lo.append(lo)


def _physics(boltzman, temp_ambient):
    p = SimpleNamespace()
    p.argon_mass = 6.63 * 10**-26
    p.ar_molar_mass = 0.039948
    p.molecules_per_mole = 6.02214179 * 10**23
    p.ideal_gas_const = 8.3145
    p.boltzman = boltzman
    p.temp_ambient = temp_ambient
    p.sigma = 3.6 * 10**(-19)
    p.argon_radius = np.sqrt(p.sigma / (4 * np.pi))
    p.collision_radius = p.argon_radius * 1
    p.collision_range = p.collision_radius * 2
    p.pressure = 101325
    p.lambda_mfp = p.boltzman * p.temp_ambient / (np.sqrt(2) * p.sigma * p.pressure)
    p.v_mean = np.sqrt(3 * p.ideal_gas_const * p.temp_ambient / p.ar_molar_mass)
    p.a_shape = np.sqrt(p.boltzman * p.temp_ambient / p.argon_mass)
    p.tau = p.lambda_mfp / p.v_mean
    return p


def cube_config(scale: float = 1.0, n_sub: int | None = None):
    """Open_Air_Cube_MC.py:26-78.  ``scale`` stretches the cube edge (config 4)."""
    c = _physics(1.38 * 10**(-23), 298)
    c.kind = "cube"
    c.cube_x = 100 * 10 ** -9 * scale
    c.cube_y = 100 * 10 ** -9 * scale
    c.cube_z = 100 * 10 ** -9 * scale
    c.cube_volume = c.cube_x * c.cube_y * c.cube_z
    n_sub = 15 if n_sub is None else n_sub
    c.num_x_subdivions = c.num_y_subdivions = c.num_z_subdivions = n_sub
    c.dx = c.cube_x / c.num_x_subdivions
    c.dy = c.cube_y / c.num_y_subdivions
    c.dz = c.cube_z / c.num_z_subdivions
    c.collision_x_overlap = c.dx / 10
    c.collision_y_overlap = c.dy / 10
    c.collision_z_overlap = c.dz / 10
    c.num_moles = c.cube_volume * c.pressure / (c.ideal_gas_const * c.temp_ambient)
    c.num_molecules = int(np.round(c.num_moles * c.molecules_per_mole).astype(int))
    c.Nmft = 20
    c.num_timesteps = c.Nmft * 25
    c.dt = c.Nmft * c.tau / c.num_timesteps
    c.min_num_particles_per_cell = int(np.floor(c.num_molecules / (n_sub * n_sub * n_sub)).astype(int))
    c.N = c.min_num_particles_per_cell * n_sub * n_sub * n_sub
    c.remaining_particles = c.num_molecules - c.N
    c.seed = 127
    c.grid = Grid(nc=(n_sub,) * 3, c0=(0, 0, 0), d=(c.dx, c.dy, c.dz),
                  band=(c.collision_x_overlap, c.collision_y_overlap, c.collision_z_overlap))
    return c


def _debye_integrand(x):
    from mpmath import exp
    return (x**3) / (exp(x) - 1)


def pore_config(temperature: bool = False, scale: float = 1.0, nmft_slice: int = 1000):
    """Open_Air_Pore_MC.py:25-94 (``temperature=False``) or Temperature_Pore_MC.py:30-113.

    Returns a namespace with the reference's names plus ``grid`` and ``geom`` (the thresholds the
    wall masks compare against, evaluated with the reference's expressions)."""
    c = _physics(1.38064852 * 10**(-23) if temperature else 1.38 * 10**(-23),
                 298.0 if temperature else 298)
    c.kind = "temp" if temperature else "pore"
    c.pore_coated_radius = 30 * 10 ** -9 * scale
    c.gap_radius = c.pore_coated_radius + 4 * 10 ** -9 * scale
    c.pore_height = 3000 * 10 ** -9 * scale
    c.hot_coating_height = 30 * 10 ** -9 * scale
    c.gap_height = c.hot_coating_height
    c.cold_coating_height = c.pore_height - c.hot_coating_height - c.gap_height
    c.hot_volume = cylinder_volume(c.pore_coated_radius, c.hot_coating_height)
    c.gap_volume = cylinder_volume(c.gap_radius, c.gap_height)
    c.cold_volume = cylinder_volume(c.pore_coated_radius, c.cold_coating_height)
    c.open_air_radius = 5 * c.pore_coated_radius
    c.open_air_height = 100 * 10 ** -9 * scale
    c.open_air_volume = cylinder_volume(c.open_air_radius, c.open_air_height)
    c.total_volume = c.hot_volume + c.gap_volume + c.cold_volume + c.open_air_volume * 2
    c.total_height = c.pore_height + c.open_air_height * 2
    c.gap_bottom_height = c.open_air_height + c.hot_coating_height
    c.gap_top_height = c.open_air_height + c.hot_coating_height + c.gap_height
    # cell grid: 7 x 7 x 148 per half-axis at scale 1 (Pore:41-46); scaled runs keep the cell edge
    # near 21.4 nm and an even number of cells per axis (the 8 colour groups need both parities)
    c.num_x_subdivions = max(1, int(round(7 * scale)))
    c.num_y_subdivions = max(1, int(round(7 * scale)))
    c.num_z_subdivions = max(2, 2 * int(round(74 * scale)))
    c.dx = c.open_air_radius / c.num_x_subdivions
    c.dy = c.open_air_radius / c.num_y_subdivions
    c.dz = c.total_height / c.num_z_subdivions
    c.num_moles = c.total_volume * c.pressure / (c.ideal_gas_const * c.temp_ambient)
    c.num_molecules = int(np.round(c.num_moles * c.molecules_per_mole).astype(int))
    c.open_air_collision_radius = c.open_air_radius - c.argon_radius
    c.gap_collision_radius = c.gap_radius - c.argon_radius
    c.pore_collision_radius = c.pore_coated_radius - c.argon_radius
    c.Nmft = 20
    c.NMFT_slice = nmft_slice
    c.num_timesteps = c.Nmft * c.NMFT_slice
    c.dt = c.Nmft * c.tau / c.num_timesteps
    c.open_air_particles = int(np.floor(c.num_molecules * (c.open_air_volume / c.total_volume)).astype(int))
    c.cold_pore_particles = int(np.floor(c.num_molecules * (c.cold_volume / c.total_volume)).astype(int))
    c.hot_pore_particles = int(np.floor(c.num_molecules * (c.hot_volume / c.total_volume)).astype(int))
    c.gap_particles = int(np.floor(c.num_molecules * (c.gap_volume / c.total_volume)).astype(int))
    c.remaining_particles = (c.num_molecules - c.gap_particles - c.hot_pore_particles
                             - c.cold_pore_particles - c.open_air_particles * 2)
    c.seed = 17
    cr = c.collision_range
    c.grid = Grid(nc=(2 * c.num_x_subdivions, 2 * c.num_y_subdivions, c.num_z_subdivions),
                  c0=(-c.num_x_subdivions, -c.num_y_subdivions, 0),
                  d=(c.dx, c.dy, c.dz), band=(cr, cr, cr))

    g = SimpleNamespace()
    a = c.argon_radius
    g.argon_mass, g.argon_radius, g.collision_range = c.argon_mass, a, cr
    g.R_oa, g.R_oa_c = c.open_air_radius, c.open_air_collision_radius
    g.R_p, g.R_p_c = c.pore_coated_radius, c.pore_collision_radius
    g.R_g, g.R_g_c = c.gap_radius, c.gap_collision_radius
    g.H, g.oah = c.total_height, c.open_air_height
    g.z_cold = c.total_height - c.open_air_height
    g.z_gb = c.open_air_height + c.hot_coating_height
    g.z_gt_pore = c.total_height - c.open_air_height - c.cold_coating_height
    g.z_gt = c.open_air_height + c.hot_coating_height + c.gap_height
    g.ten_a = 10 * a
    g.R_oa_sq, g.R_g_sq, g.R_p_sq = c.open_air_radius**2, c.gap_radius**2, c.pore_coated_radius**2
    g.zc3 = c.total_height - c.open_air_height + a
    g.zh3 = c.open_air_height - a
    g.zgt_m = c.gap_top_height - a
    g.zgb_p = c.gap_bottom_height + a
    g.R_g_c_sq, g.R_p_c_sq = c.gap_collision_radius**2, c.pore_collision_radius**2
    g.recap_lo = 50 * 10 ** -9 * scale
    g.recap_hi = c.total_height - (50 * 10 ** -9 * scale)
    g.E_cold = g.E_hot = 0.0
    g.alpha_c = g.alpha_g = 0.0
    g.cos85 = math.cos(85 * math.pi / 180)
    if temperature:
        from mpmath import quad
        c.t_cold, c.t_hot = 293.0, 353.0
        c.t_debye_graphene, c.t_debye_alumina = 1813.0, 980.0
        c.coated_accomodation_coeff, c.gap_accomodation_coeff = 0.95, 0.8
        c.num_atoms_unitcell_graphene, c.num_atoms_unitcell_alumina = 2, 10
        c.debye_quadrature_cold = quad(_debye_integrand, [0, c.t_debye_graphene / c.t_cold])
        c.debye_quadrature_hot = quad(_debye_integrand, [0, c.t_debye_graphene / c.t_hot])
        c.surface_energy_cold = (9 * c.t_cold * c.num_atoms_unitcell_graphene * c.boltzman
                                 * (c.t_cold / c.t_debye_graphene)**3 * c.debye_quadrature_cold)
        c.surface_energy_hot = (9 * c.t_hot * c.num_atoms_unitcell_graphene * c.boltzman
                                * (c.t_hot / c.t_debye_graphene)**3 * c.debye_quadrature_hot)
        g.E_cold, g.E_hot = float(c.surface_energy_cold), float(c.surface_energy_hot)
        g.alpha_c, g.alpha_g = c.coated_accomodation_coeff, c.gap_accomodation_coeff
    c.geom = g
    return c


def surface_energy_gap(c, z_value):
    """Temp:143-152: alumina gap wall, temperature linear in z, Debye integral by mpmath.quad.
    Returns an mpf like the reference."""
    from mpmath import quad
    m = (c.t_cold - c.t_hot) / (c.gap_height)
    t_gap = m * (z_value - c.gap_bottom_height) + c.t_hot
    q = quad(_debye_integrand, [0, c.t_debye_alumina / t_gap])
    return 9 * t_gap * c.num_atoms_unitcell_alumina * c.boltzman * (t_gap / c.t_debye_alumina)**3 * q


def gap_energy_chebyshev(c, ncoef: int = 24):
    """Chebyshev fit of ``surface_energy_gap`` over the gap's z range, for the device-RNG mode
    (the reference integrates with mpmath per hit, Temp:519).  Returns (coef, zmid, inv_half);
    the interval is widened by one collision range so contact points that round just outside
    the gap stay inside the fit."""
    zlo = c.gap_bottom_height - c.collision_range
    zhi = c.gap_top_height + c.collision_range
    zmid, half = 0.5 * (zlo + zhi), 0.5 * (zhi - zlo)
    k = np.arange(ncoef)
    nodes = np.cos(np.pi * (k + 0.5) / ncoef)
    vals = np.array([float(surface_energy_gap(c, zmid + half * u)) for u in nodes])
    coef = np.array([2.0 / ncoef * np.sum(vals * np.cos(np.pi * j * (k + 0.5) / ncoef)) for j in range(ncoef)])
    coef[0] *= 0.5
    return coef, float(zmid), float(1.0 / half)
