"""Output writers of the three drivers: the eight histogram text files and momentum_energy.csv,
byte-compatible with what the reference scripts write.

  hist_x_axis_*_data.txt = str(bins[0:200])    (Open_Air_Pore_MC.py:607-630, same in Cube / Temp)
  hist_y_axis_*_data.txt = str(n) with n = np.histogram(..., density=True)   (Pore:575-596)
  momentum_energy.csv    = pandas.DataFrame.from_dict({...}).to_csv          (Temperature_Pore_MC.py:929-933)

The device keeps the histogram *counts* (integers, exact); the density normalisation is the one
np.histogram applies: counts / diff(edges) / counts.sum().
"""
from __future__ import annotations

import os
import sys

import numpy as np

from .config import HIST_RANGE, NUM_BINS

SUFFIXES = ("total", "x", "y", "z")


def bin_edges():
    return np.linspace(HIST_RANGE[0], HIST_RANGE[1], NUM_BINS + 1)


def density(counts):
    """np.histogram(..., density=True) from integer counts (numpy/lib/_histograms_impl.py:873-875)."""
    counts = np.asarray(counts)
    db = np.array(np.diff(bin_edges()), float)
    with np.errstate(divide="ignore", invalid="ignore"):
        return counts / db / counts.sum()


def write_histograms(counts4, directory="."):
    """counts4: [4][200] integer counts in the order total, x, y, z."""
    edges = bin_edges()
    old = np.get_printoptions()
    np.set_printoptions(threshold=sys.maxsize)      # the reference sets this at import (Pore:12)
    try:
        for j, suffix in enumerate(SUFFIXES):
            n = density(counts4[j])
            with open(os.path.join(directory, "hist_x_axis_%s_data.txt" % suffix), "w") as f:
                f.write(str(edges[0:len(n)]))
            with open(os.path.join(directory, "hist_y_axis_%s_data.txt" % suffix), "w") as f:
                f.write(str(n))
    finally:
        np.set_printoptions(**old)


def write_momentum_energy_csv(momentum, energy_cold, energy_hot, path="momentum_energy.csv"):
    """Per-step series as the reference prints them: mpf values through pandas, i.e. 15 significant
    digits (a step without an energized hit stays the Python int 0, as in the reference)."""
    import mpmath
    import pandas as pd

    def conv(v):
        if isinstance(v, (int,)) and v == 0:
            return 0
        return v if isinstance(v, mpmath.mpf) else mpmath.mpf(float(v))
    data = {"Momentum": [conv(v) for v in momentum],
            "EnergyCold": [conv(v) for v in energy_cold],
            "EnergyHot": [conv(v) for v in energy_hot]}
    pd.DataFrame.from_dict(data).to_csv(path)
