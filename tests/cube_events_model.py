"""Python model of the event-driven cube sweep (k_sweep_detect + k_sweep_events, amc_kernels.cuh): the serial
lexicographic cell sweep of Open_Air_Cube_MC.py:232-336 visits 3,375 cells per timestep, but a visit changes
nothing unless two members overlap.  The model finds the cells that hold an overlapping pair in one parallel-shaped
pass over the positions as they are before the sweep, then walks only the flagged cells in sweep order; a collision
flags the later cells that hold (or held) one of the two moved particles.  The reference's stale layer masks
(x mask taken once per x layer, y mask once per column, z mask per cell -- Cube:233,235,237) are reproduced through
per-particle snapshots that are refreshed when the walk enters a new layer / column.

Test infrastructure: tests/test_host_round2.py checks this model against the oracle's serial sweep on the CPU, so
the algorithm the CUDA kernels implement is pinned without a GPU."""
from __future__ import annotations

import math

import numpy as np


def axis_cells(grid, a, v):
    """cells k of axis a with lo[k] < v < edge[k+1] (at most two: the owner and the band of the next one)"""
    edge, lo, nc = grid.edge[a], grid.lo[a], grid.nc[a]
    k = int(np.searchsorted(edge, v, side="right")) - 1
    out = []
    for kk in (k - 1, k, k + 1):
        if 0 <= kk < nc and lo[kk] < v < edge[kk + 1]:
            out.append(kk)
    return out


def resolve_cell(st, members, cr, mass, paths, pairs, cell):
    """pairwise_particles_in_cell (Cube:253-324) in plain arithmetic on the members (ascending index)."""
    x, y, z, vx, vy, vz = st["x"], st["y"], st["z"], st["vx"], st["vy"], st["vz"]
    d, dx, dy, dz, fl = st["dist"], st["dist_x"], st["dist_y"], st["dist_z"], st["flag"]
    ncol = 0
    moved = {}
    m = sorted(members)
    for ii in range(len(m)):
        i = m[ii]
        for jj in range(ii):
            j = m[jj]
            x1, x2, y1, y2, z1, z2 = x[j], x[i], y[j], y[i], z[j], z[i]
            ddx, ddy, ddz = x2 - x1, y2 - y1, z2 - z1
            sep = math.sqrt((ddx * ddx + ddy * ddy) + ddz * ddz)
            if not sep < cr:
                continue
            vx1, vx2, vy1, vy2, vz1, vz2 = vx[j], vx[i], vy[j], vy[i], vz[j], vz[i]
            rx, ry, rz = -vx2 + vx1, -vy2 + vy1, -vz2 + vz1
            a = (rx * rx + ry * ry) + rz * rz
            b = 2 * ((ddx * rx + ddy * ry) + ddz * rz)
            c = ((ddx * ddx + ddy * ddy) + ddz * ddz) - cr * cr
            disc = b * b - (4 * a) * c
            if not disc >= 0.0 or a == 0.0:
                continue
            root = math.sqrt(disc)
            t1, t2 = (-b + root) / (2 * a), (-b - root) / (2 * a)
            t = t1 if t1 > t2 else t2
            for (q, wx, wy, wz) in ((j, vx1, vy1, vz1), (i, vx2, vy2, vz2)):
                if fl[q]:
                    sp = math.sqrt((wx * wx + wy * wy) + wz * wz)
                    paths.append((abs(d[q] - abs(sp * t)), abs(dx[q] - abs(wx * t)), abs(dy[q] - abs(wy * t)), abs(dz[q] - abs(wz * t))))
                else:
                    fl[q] = 1
            for q in (i, j):
                moved.setdefault(q, (x[q], y[q], z[q]))
            nx1, ny1, nz1 = x1 - vx1 * t, y1 - vy1 * t, z1 - vz1 * t
            nx2, ny2, nz2 = x2 - vx2 * t, y2 - vy2 * t, z2 - vz2 * t
            n0, n1, n2 = (nx2 - nx1) / cr, (ny2 - ny1) / cr, (nz2 - nz1) / cr
            pp = (((vx1 * n0 + vy1 * n1) + vz1 * n2) - ((vx2 * n0 + vy2 * n1) + vz2 * n2)) / mass
            pm = pp * mass
            wx1, wy1, wz1 = vx1 - pm * n0, vy1 - pm * n1, vz1 - pm * n2
            wx2, wy2, wz2 = vx2 + pm * n0, vy2 + pm * n1, vz2 + pm * n2
            x[j], y[j], z[j] = nx1 + wx1 * t, ny1 + wy1 * t, nz1 + wz1 * t
            x[i], y[i], z[i] = nx2 + wx2 * t, ny2 + wy2 * t, nz2 + wz2 * t
            vx[j], vy[j], vz[j] = wx1, wy1, wz1
            vx[i], vy[i], vz[i] = wx2, wy2, wz2
            d[i] = abs(math.sqrt((wx2 * wx2 + wy2 * wy2) + wz2 * wz2) * t)
            d[j] = abs(math.sqrt((wx1 * wx1 + wy1 * wy1) + wz1 * wz1) * t)
            dx[i], dy[i], dz[i] = abs(wx2 * t), abs(wy2 * t), abs(wz2 * t)
            dx[j], dy[j], dz[j] = abs(wx1 * t), abs(wy1 * t), abs(wz1 * t)
            pairs.append((i, j, cell))
            ncol += 1
    return ncol, moved


def sweep_events(st, grid, cr, mass):
    """One particle-particle pass.  st: dict of NumPy arrays (modified in place).  Returns (collisions,
    reference-equivalent pair tests, completed paths, pairs, cells visited)."""
    nx, ny, nz = grid.nc
    x, y, z = st["x"], st["y"], st["z"]
    n = len(x)
    # ---- pass 1 (k_sweep_detect): per column the members by the x and y masks, per cell the member count and
    # whether two members overlap -- all on the positions before the sweep
    cx = [np.nonzero((grid.lo[0][k] < x) & (x < grid.edge[0][k + 1]))[0] for k in range(nx)]
    col_list = {}
    n0 = np.zeros((nx, ny, nz), dtype=np.int64)
    flagged = np.zeros(nx * ny * nz, dtype=bool)
    cr2 = cr * cr * (1 + 1e-9)
    for xl in range(nx):
        ix = cx[xl]
        for yl in range(ny):
            iy = ix[(grid.lo[1][yl] < y[ix]) & (y[ix] < grid.edge[1][yl + 1])]
            col_list[(xl, yl)] = iy
            for zl in range(nz):
                m = iy[(grid.lo[2][zl] < z[iy]) & (z[iy] < grid.edge[2][zl + 1])]
                n0[xl, yl, zl] = len(m)
                if len(m) >= 2:
                    px, py, pz = x[m], y[m], z[m]
                    d2 = (px[:, None] - px[None, :]) ** 2 + (py[:, None] - py[None, :]) ** 2 + (pz[:, None] - pz[None, :]) ** 2
                    np.fill_diagonal(d2, np.inf)
                    if (d2 < cr2).any():   # conservative: the visit decides exactly
                        flagged[(xl * ny + yl) * nz + zl] = True
    checks = int((n0 * (n0 - 1) // 2).sum())
    # ---- pass 2 (k_sweep_events): the flagged cells in sweep order
    tagged = np.zeros(n, dtype=bool)
    xs, ys = np.zeros(n), np.zeros(n)   # what the current layer's x mask / the current column's y mask saw (tagged particles only)
    moved_list = []
    cur_xl, cur_col = -1, (-1, -1)
    paths, pairs = [], []
    ncol = visits = 0
    c = 0
    ncell = nx * ny * nz
    while True:
        nxt = np.nonzero(flagged[c:])[0]
        if len(nxt) == 0:
            break
        c += int(nxt[0])
        flagged[c] = False
        xl, yl, zl = c // (ny * nz), (c // nz) % ny, c % nz
        if xl != cur_xl:
            for i in moved_list:
                xs[i] = x[i]
            cur_xl = xl
        if (xl, yl) != cur_col:
            for i in moved_list:
                ys[i] = y[i]
            cur_col = (xl, yl)
        members = []
        cand = [i for i in col_list[(xl, yl)] if not tagged[i]] + moved_list
        for i in cand:
            X = xs[i] if tagged[i] else x[i]
            Y = ys[i] if tagged[i] else y[i]
            if (grid.lo[0][xl] < X < grid.edge[0][xl + 1] and grid.lo[1][yl] < Y < grid.edge[1][yl + 1]
                    and grid.lo[2][zl] < z[i] < grid.edge[2][zl + 1]):
                members.append(int(i))
        visits += 1
        nm = len(members)
        checks += nm * (nm - 1) // 2 - int(n0[xl, yl, zl]) * (int(n0[xl, yl, zl]) - 1) // 2
        k, moved = resolve_cell(st, members, cr, mass, paths, pairs, c)
        ncol += k
        for i, old in moved.items():
            if not tagged[i]:
                tagged[i] = True
                xs[i], ys[i] = old[0], old[1]
                moved_list.append(i)
            for (tx, ty, tz) in (old, (x[i], y[i], z[i])):
                czs, cys, cxs = axis_cells(grid, 2, tz), axis_cells(grid, 1, ty), axis_cells(grid, 0, tx)
                for zz in czs:                                   # (i) same column, later cell
                    if zz > zl:
                        flagged[(xl * ny + yl) * nz + zz] = True
                for yy in cys:                                   # (ii) same layer, later column
                    if yy > yl:
                        for zz in czs:
                            flagged[(xl * ny + yy) * nz + zz] = True
                for xx in cxs:                                   # (iii) later layer
                    if xx > xl:
                        for yy in cys:
                            for zz in czs:
                                flagged[(xx * ny + yy) * nz + zz] = True
        c += 1
        if c >= ncell:
            break
    return ncol, checks, paths, pairs, visits
