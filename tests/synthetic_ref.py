"""NumPy restatement of the device-side synthetic generator (k_init_synthetic): Philox4x32-10 keyed by the seed,
counter (i, 0x1417, c), two 53-bit uniforms per call; region by cumulative weight, uniform position in a cylinder /
box, Box-Muller velocities.  Test infrastructure: the transcendental functions differ from CUDA's in the last ulp,
so comparisons use a 1e-12 relative tolerance."""
import numpy as np

M0, M1, W0, W1 = 0xD2511F53, 0xCD9E8D57, 0x9E3779B9, 0xBB67AE85


def philox4x32_10(c, k0, k1):
    c = [np.asarray(v, dtype=np.uint64) for v in c]
    k0, k1 = np.uint64(k0), np.uint64(k1)
    mask = np.uint64(0xFFFFFFFF)
    for _ in range(10):
        p0, p1 = np.uint64(M0) * c[0], np.uint64(M1) * c[2]
        hi0, lo0, hi1, lo1 = p0 >> np.uint64(32), p0 & mask, p1 >> np.uint64(32), p1 & mask
        c = [hi1 ^ c[1] ^ k0, lo1, hi0 ^ c[3] ^ k1, lo0]
        k0, k1 = (k0 + np.uint64(W0)) & mask, (k1 + np.uint64(W1)) & mask
    return c


def u53(hi, lo):
    return ((hi >> np.uint64(5)).astype(np.float64) * 67108864.0 + (lo >> np.uint64(6)).astype(np.float64)) * (1.0 / 9007199254740992.0)


def generate(spec, ids):
    ids = np.asarray(ids, dtype=np.uint64)
    k0, k1 = spec.seed & 0xFFFFFFFF, spec.seed >> 32
    u = []
    for c in range(4):
        r = philox4x32_10([ids & np.uint64(0xFFFFFFFF), ids >> np.uint64(32), np.full(len(ids), 0x1417, np.uint64),
                           np.full(len(ids), c, np.uint64)], k0, k1)
        u += [u53(r[0], r[1]), u53(r[2], r[3])]
    cum = np.array(spec.cum_weight[:spec.n_regions])
    reg = np.minimum(np.searchsorted(cum, u[0], side="right"), spec.n_regions - 1)
    radius, bx, by = np.array(spec.radius[:]), np.array(spec.bx[:]), np.array(spec.by[:])
    zlo, zhi = np.array(spec.z_lo[:]), np.array(spec.z_hi[:])
    if spec.shape == 0:
        rr, th = radius[reg] * np.sqrt(u[1]), 6.283185307179586 * u[2]
        x, y = rr * np.cos(th), rr * np.sin(th)
    else:
        x, y = bx[reg] * u[1], by[reg] * u[2]
    z = zlo[reg] + (zhi[reg] - zlo[reg]) * u[3]
    r1 = spec.sigma * np.sqrt(-2.0 * np.log(1.0 - u[4]))
    r2 = spec.sigma * np.sqrt(-2.0 * np.log(1.0 - u[6]))
    t1, t2 = 6.283185307179586 * u[5], 6.283185307179586 * u[7]
    return x, y, z, r1 * np.cos(t1), r1 * np.sin(t1), r2 * np.cos(t2), reg
