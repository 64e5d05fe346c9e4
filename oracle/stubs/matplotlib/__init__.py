"""Minimal matplotlib stand-in so the upstream reference scripts import in this image
(matplotlib is not installed; every script does `import matplotlib.pyplot as plt` at line 1
and calls plt.figure at import time).  Test infrastructure only."""
