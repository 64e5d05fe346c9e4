"""Host Mersenne-Twister draws for the energized walls ("parity mode").

The reference scatters a particle off a thermally accommodating wall into a random direction
within 85 degrees of the inward normal (Temperature_Pore_MC.py:132-141), drawing from BOTH global
generators per attempt: ``np.random.uniform`` (cos theta), ``random.uniform`` (phi) and
``np.random.choice`` (sign of y) -- Temperature_Pore_MC.py:119-126.  momentum_energy.csv is only
reproducible if those draws happen in the reference's order (case by case, ascending particle
index, one rejection loop per hit), so in parity mode the device reports the pending hits of a
case, this module draws their directions on the host, and the device applies them.
"""
from __future__ import annotations

import math
import random as _pyrandom

import numpy as np

COS85 = math.cos(85 * math.pi / 180)


def _unit_components():
    costheta = np.random.uniform(low=-1.0, high=1.0)
    phi = _pyrandom.uniform(0, math.pi)
    theta = math.acos(costheta)
    sin_t = math.sin(theta)
    sign = np.random.choice([-1, 1])
    return 1 * math.cos(phi) * sin_t, 1 * math.sin(phi) * sin_t * sign, 1 * math.cos(theta)


def inbound_direction(norm: np.ndarray) -> np.ndarray:
    """One accepted direction about ``norm`` (rejection loop of Temp:132-141)."""
    while True:
        d = np.array(_unit_components())
        dn = np.dot(d, norm)
        if abs(dn) < COS85:
            continue
        return -d if dn < COS85 else d


def directions_for_hits(normals: np.ndarray) -> np.ndarray:
    """Directions for the pending hits of one wall case, in ascending-index order.  Rows whose
    normal is NaN are hits that raised a floating-point error in the reference (no draw)."""
    out = np.zeros_like(normals)
    for k in range(len(normals)):
        if normals[k, 0] == normals[k, 0]:
            out[k] = inbound_direction(normals[k])
    return out
