"""Helper kept for drop-in compatibility with the reference layout (utils.py next to the scripts)."""
import numpy as np


def cylinder_volume(radius, height):
    return np.pi * radius ** 2 * height
