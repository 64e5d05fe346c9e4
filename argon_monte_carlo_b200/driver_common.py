"""Shared pieces of the three drop-in driver scripts (drivers/*.py): the optional command line the
reference does not have, the end-of-run report and the output files."""
from __future__ import annotations

import argparse
import os
from time import time

import numpy as np

from . import outputs


def parse_args(default_steps, description, temp=False):
    ap = argparse.ArgumentParser(description=description)
    ap.add_argument("--steps", type=int, default=int(os.environ.get("AMC_STEPS", default_steps)),
                    help="number of timesteps (default: the script's own num_timesteps = %d)" % default_steps)
    ap.add_argument("--outdir", default=os.environ.get("AMC_OUTDIR", "."), help="where the result files go (default: CWD)")
    ap.add_argument("--device", type=int, default=int(os.environ.get("AMC_DEVICE", 0)))
    ap.add_argument("--chunk", type=int, default=100, help="timesteps per device call (device-RNG / specular runs)")
    if temp:
        ap.add_argument("--rng", choices=["host", "device"], default=os.environ.get("AMC_RNG", "host"),
                        help="host: the reference's Mersenne-Twister draws in its order (reproduces "
                             "momentum_energy.csv); device: Philox on the GPU, no host round trips")
    ap.add_argument("--show", action="store_true", help="plt.show() at the end if matplotlib is installed")
    return ap.parse_args()


def final_report(sim, total_errs, total_cols, start, outdir, show=False, labels=("Simulation mean free path: ",
                 "Simulation mean x free path: ", "Simulation mean y free path: ", "Simulation mean z free path: ")):
    counts, n_paths, sums = sim.histograms()
    print(' ', total_errs, ' errors/warnings - potential lost particles')
    print(' ', total_cols, ' collisions')
    print(' ', n_paths, ' completed paths')
    with np.errstate(invalid="ignore", divide="ignore"):
        means = sums / n_paths if n_paths else np.full(4, np.nan)
    for lab, m in zip(labels, means):
        print(lab + str(m))
    print('Num of measured full paths total: ' + str(n_paths))
    print('Runtime: ' + str((time() - start) / 60.0) + ' minutes (pre histogram data rewrite)')
    outputs.write_histograms(counts, outdir)
    return counts, n_paths, means


def maybe_show(show):
    if not show:
        return
    try:
        import matplotlib.pyplot as plt
        plt.show()
    except Exception:
        pass
