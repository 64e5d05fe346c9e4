"""Host-side logic of the multi-GPU path without GPUs: cut selection, partitioning, and the
torch.distributed transport on 2 gloo ranks with CPU tensors."""
import os
import socket

import numpy as np
import pytest


def test_balanced_cuts_and_owner_layers(pore_cfg, pore_init):
    from argon_monte_carlo_b200 import slab
    z, edges = pore_init[2], pore_cfg.grid.edge[2]
    layer = slab.owner_layer(z, edges)
    assert layer.min() >= 0 and layer.max() <= 147
    k = np.arange(len(z))[::997]
    assert all(edges[layer[i]] <= z[i] < edges[layer[i] + 1] for i in k)
    for nranks in (1, 2, 4, 8):
        cuts = slab.balanced_cuts(z, edges, nranks)
        assert cuts[0] == 0 and cuts[-1] == 148 and len(cuts) == nranks + 1 and (np.diff(cuts) >= 2).all()
        per = np.array([np.count_nonzero((layer >= cuts[r]) & (layer < cuts[r + 1])) for r in range(nranks)])
        assert per.sum() == len(z)
        # 62 % of the particles sit in 6 % of the length: equal-length slabs would be hopeless, these are not
        assert per.max() <= 1.35 * len(z) / nranks + 40000
    with pytest.raises(ValueError):
        slab.balanced_cuts(z, edges, 100)


def test_local_grid_is_a_window_of_the_global_tables(pore_cfg):
    from argon_monte_carlo_b200 import slab
    g = pore_cfg.grid
    lg = slab.local_grid(g, 5, 9)
    assert lg.nc == (14, 14, 4) and lg.c0 == (-7, -7, 5)
    assert np.array_equal(lg.edge[2], g.edge[2][5:10]) and np.array_equal(lg.lo[2], g.lo[2][5:9])
    assert lg.edge[0] is g.edge[0]


def test_chunked_synthetic_state_is_chunk_deterministic():
    from argon_monte_carlo_b200 import config, init_state
    cfg = config.pore_config(True, scale=0.5)
    a = init_state.synthetic_pore_chunked(cfg, seed=5, chunk_size=20000)
    b = init_state.synthetic_pore_chunked(cfg, seed=5, chunk_size=20000, keep=lambda z: z > 1e-6)
    assert len(a[0]) == cfg.num_molecules and np.array_equal(a[0], np.arange(cfg.num_molecules))
    m = a[3] > 1e-6
    assert np.array_equal(b[0], a[0][m]) and np.array_equal(b[4], a[4][m])
    r = np.hypot(a[1], a[2])
    assert r.max() <= cfg.open_air_radius and a[3].min() > 0 and a[3].max() < cfg.total_height


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, out):
    import torch
    import torch.distributed as dist
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from types import SimpleNamespace
    from argon_monte_carlo_b200 import slab
    cap, bcap = 5, 3
    caps = [cap if abs(d - rank) == 1 else 2 for d in range(world)]          # big blocks for neighbours only
    splits = [(c + 1) * slab.REC for c in caps]
    offsets = np.concatenate([[0], np.cumsum(splits)])
    r = SimpleNamespace(rank=rank, xfer_splits=splits, xfer_offsets=offsets,
                        xfer_block=lambda buf, peer: buf[int(offsets[peer]):int(offsets[peer + 1])],
                        xfer_send=torch.zeros(int(offsets[-1]), dtype=torch.float64),
                        xfer_recv=torch.zeros(int(offsets[-1]), dtype=torch.float64),
                        bnd_send_all=None, bnd_recv_all=None,
                        bnd_send_up=torch.full((bcap + 1, slab.REC), 100.0 + rank, dtype=torch.float64),
                        bnd_send_down=torch.full((bcap + 1, slab.REC), 200.0 + rank, dtype=torch.float64),
                        bnd_recv_up=torch.zeros(bcap + 1, slab.REC, dtype=torch.float64),
                        bnd_recv_down=torch.zeros(bcap + 1, slab.REC, dtype=torch.float64))
    for dst in range(world):
        r.xfer_block(r.xfer_send, dst)[:] = 10 * rank + dst          # block (src -> dst) tagged 10*src + dst
    T = slab.DistTransport()
    T.alltoall([r])
    ok = all(bool((r.xfer_block(r.xfer_recv, src) == 10 * src + rank).all()) for src in range(world))
    T.neighbors([r])
    if rank + 1 < world:
        ok &= bool((r.bnd_recv_up == 200.0 + rank + 1).all())     # what the rank above sent down
    else:
        ok &= bool((r.bnd_recv_up == 0).all())
    if rank > 0:
        ok &= bool((r.bnd_recv_down == 100.0 + rank - 1).all())   # what the rank below sent up
    else:
        ok &= bool((r.bnd_recv_down == 0).all())
    tot = T.allreduce_sum(np.array([1.0 + rank, 2.0]))
    ok &= bool(np.allclose(tot, [sum(1.0 + k for k in range(world)), 2.0 * world]))
    # checksum parts add up mod 2^64 (values close to 2^64 wrap), descriptors come back ordered by rank
    big = (1 << 64) - 5
    dg = T.allreduce_u64((big, 7 + rank, 1000 * (rank + 1)))
    ok &= dg == ((big * world) % (1 << 64), sum(7 + k for k in range(world)), sum(1000 * (k + 1) for k in range(world)))
    blobs = T.allgather_bytes(bytes([rank] * (3 + rank)))
    ok &= blobs == [bytes([k] * (3 + k)) for k in range(world)]
    out[rank] = ok
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_dist_transport_gloo(world):
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    out = ctx.Manager().dict()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(k, world, port, out)) for k in range(world)]
    [p.start() for p in procs]
    [p.join(120) for p in procs]
    assert all(p.exitcode == 0 for p in procs)
    assert all(out.get(k) for k in range(world)), dict(out)


def test_balanced_cuts_minimise_the_fullest_slab():
    """Against brute force on small random layer histograms: no choice of cuts has a lighter fullest slab; odd cuts
    are used when they cost less than the tolerance."""
    import itertools
    from argon_monte_carlo_b200 import slab
    rng = np.random.default_rng(3)
    for trial in range(30):
        ncz = int(rng.integers(8, 15))
        nranks = int(rng.integers(2, 4))
        hist = rng.integers(1, 50, ncz) * np.where(rng.random(ncz) < 0.2, 20, 1)     # a few dense layers, as in the end caps
        edges = np.arange(ncz + 1, dtype=np.float64)
        z = np.repeat(np.arange(ncz) + 0.5, hist)
        cum = np.concatenate([[0], np.cumsum(hist)])
        load = lambda c: max(cum[b] - cum[a] for a, b in zip(c, c[1:]))
        best = min(load((0,) + mid + (ncz,)) for mid in itertools.combinations(range(1, ncz), nranks - 1)
                   if all(b - a >= 2 for a, b in zip((0,) + mid, mid + (ncz,))))
        free = slab.balanced_cuts(z, edges, nranks, prefer_odd=False)
        assert load(tuple(free)) == best, (hist, free)
        odd = slab.balanced_cuts(z, edges, nranks, prefer_odd=True)
        assert (np.diff(odd) >= 2).all() and load(tuple(odd)) <= best * 1.03 + 1e-9
        if any(int(c) % 2 == 0 for c in odd[1:-1]):      # an even cut only where every all-odd choice is too heavy
            odd_only = [load((0,) + mid + (ncz,)) for mid in itertools.combinations(range(1, ncz, 2), nranks - 1)
                        if all(b - a >= 2 for a, b in zip((0,) + mid, mid + (ncz,)))]
            assert not odd_only or min(odd_only) > best * 1.03


def test_device_initialiser_restatement_and_spec():
    """tests/synthetic_ref.py (the NumPy restatement the GPU test compares k_init_synthetic with) uses the same
    Philox4x32-10 as the oracle's C implementation, and init_state.pore_spec describes the regions of Pore:79-83."""
    import ctypes as C
    from oracle import oracle as O
    from argon_monte_carlo_b200 import config, init_state
    from synthetic_ref import philox4x32_10, generate
    L = O.lib()
    rng = np.random.default_rng(1)
    ctr = rng.integers(0, 2 ** 32, (50, 4), dtype=np.uint64)
    key = rng.integers(0, 2 ** 32, (50, 2), dtype=np.uint64)
    for c, k in zip(ctr, key):
        out = (C.c_uint32 * 4)()
        L.orc_philox4x32_10((C.c_uint32 * 4)(*[int(v) for v in c]), (C.c_uint32 * 2)(*[int(v) for v in k]), out)
        got = philox4x32_10([np.array([v]) for v in c], int(k[0]), int(k[1]))
        assert [int(g[0]) for g in got] == [int(v) for v in out]
    cfg = config.pore_config(True)
    spec = init_state.pore_spec(cfg, seed=9, keep_z=(1e-7, 2e-7))
    cum = np.array(spec.cum_weight[:5])
    assert spec.n_regions == 5 and spec.n_total == cfg.num_molecules and cum[-1] == 1.0 and (np.diff(cum) > 0).all()
    assert (spec.keep_z_lo, spec.keep_z_hi) == (1e-7, 2e-7)
    x, y, z, vx, vy, vz, reg = generate(spec, np.arange(200000))
    frac = np.bincount(reg, minlength=5) / 200000.0
    vols = np.array([cfg.open_air_volume, cfg.hot_volume, cfg.gap_volume, cfg.cold_volume, cfg.open_air_volume])
    assert np.max(np.abs(frac - vols / vols.sum())) < 5e-3
    assert z.min() >= cfg.argon_radius and z.max() <= cfg.total_height - cfg.argon_radius
    assert abs(np.std(np.concatenate([vx, vy, vz])) / cfg.a_shape - 1) < 5e-3
