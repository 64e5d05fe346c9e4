"""Where does a fused colour-group launch of the multi-GPU step spend its time?  Debug build (-DAMC_SLAB_PROBE), run
under torchrun on N GPUs:  python -m torch.distributed.run --nproc-per-node N ... tools/slab_probe.py [particles per GPU]
Prints, per rank and colour group, ns after the first CTA of the launch started: hand-over from above / below applied,
last cut-adjacent visit done, last visit done, records sent; plus the records applied and the cut-adjacent visits."""
import ctypes as C, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ["AMC_LIBRARY"] = os.path.join(ROOT, "argon_monte_carlo_b200", "libamc_probe.so")
import numpy as np, torch, torch.distributed as dist
import bench
from argon_monte_carlo_b200 import init_state, slab
world, rank, local = int(os.environ["WORLD_SIZE"]), int(os.environ["RANK"]), int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
per = int(sys.argv[1]) if len(sys.argv) > 1 else 12_500_000
cfg, _ = bench.scaled_temp_config(per * world)
zs = init_state.synthetic_pore_chunk(cfg, 17, 0, 1 << 22)[3]
cuts = slab.balanced_cuts(zs, cfg.grid.edge[2], world)
sim = slab.SlabSimulation(cfg, world, zs, transport=slab.DistTransport(), local_ranks=[rank], devices=[local], cuts=cuts,
                          n_total=cfg.num_molecules, seed=17, p2p=True)
sim.init_synthetic(lambda kz: init_state.pore_spec(cfg, 17, keep_z=kz))
sim.step_fused(5, reduce=False)
lib = sim.ranks[0].sim.lib
out = (C.c_ulonglong * 64)()
acc = np.zeros((8, 8))
K = 10
for it in range(K):
    lib.amc_debug_probe(out, 1)
    dist.barrier()
    sim.step_fused(1, reduce=False)
    lib.amc_debug_probe(out, 0)
    a = np.array(out[:], dtype=np.float64).reshape(8, 8)
    for g in range(8):
        t0 = a[g, 0]
        acc[g, 1:6] += np.where(a[g, 1:6] > 0, a[g, 1:6] - t0, 0.0) / 1000.0
        acc[g, 6:] += a[g, 6:]
acc /= K
for r in range(world):
    dist.barrier()
    if r == rank:
        print("rank %d (cuts %s): per group [applied_above applied_below last_cut_visit last_visit sent] us, [records cut_visits]" % (rank, list(cuts)))
        for g in range(8):
            print("   g%d  %s   %s" % (g, " ".join("%6.1f" % v for v in acc[g, 1:6]), " ".join("%5.1f" % v for v in acc[g, 6:])))
        sys.stdout.flush()
sim.close()
dist.destroy_process_group()
