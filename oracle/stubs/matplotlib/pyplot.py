"""pyplot stub: figure()/add_subplot()/hist()/show() -- hist wraps np.histogram exactly the way
matplotlib.axes.Axes.hist does for a single dataset (density passed straight through)."""
import numpy as np


class _Axes:
    def hist(self, x, bins=None, range=None, density=False, **kw):
        n, edges = np.histogram(np.asarray(x, dtype=float), bins=bins, range=range, density=density)
        return n.astype(float) if not density else n, edges, []

    def __getattr__(self, name):
        return lambda *a, **k: None


class _Figure:
    def add_subplot(self, *a, **k):
        return _Axes()

    def __getattr__(self, name):
        return lambda *a, **k: None


def figure(*a, **k):
    return _Figure()


def show(*a, **k):
    return None
