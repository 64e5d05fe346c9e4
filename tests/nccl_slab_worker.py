"""Worker of tests/test_gpu_nccl_slab.py (also usable directly):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        tests/nccl_slab_worker.py OUT.json [--particles M] [--steps K] [--kind temp|pore]

Every rank drives one z slab of ONE synthetic pore on its own GPU -- --mode p2p: amc_slab_step, every transfer a
kernel writing into the peer's buffer over NVLink; --mode nccl: the step-wise entry points with NCCL all-to-all +
neighbour send/recv (slab.DistTransport) -- and rank 0 then repeats the same job as a single domain on its GPU.  Compared: the per-step
counters (summed over ranks) and the order-independent checksum of the id-ordered state (amc_state_digest).
Exit code 0 only if everything is identical."""
import argparse
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

COUNTERS = ("wall_collisions", "pp_collisions", "pair_checks_ref", "oob_after_walls", "oob_after_pp", "errors",
            "completed_paths")


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("out")
    ap.add_argument("--particles", type=int, default=2_000_000)
    ap.add_argument("--steps", type=int, default=6)
    ap.add_argument("--kind", default="temp", choices=["temp", "pore"])
    ap.add_argument("--dense", action="store_true",
                    help="stress case: the small dense pore (overlapping start, thousands of collisions per step) with every cut "
                         "inside an end cap, one-layer slabs included: many hand-over records per colour group, even cuts (pre-round)")
    ap.add_argument("--mode", default="p2p", choices=["p2p", "nccl"],
                    help="p2p: amc_slab_step (records written peer to peer by kernels); nccl: the step-wise entry points over torch.distributed")
    args = ap.parse_args()
    import torch
    import torch.distributed as dist
    from argon_monte_carlo_b200 import amc, config, init_state, slab
    world, rank, local = int(os.environ["WORLD_SIZE"]), int(os.environ["RANK"]), int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    state = None
    if args.dense:
        cfg = config.pore_config(args.kind == "temp", scale=0.5)
        state = init_state.synthetic_pore_state(cfg, seed=11)
        nz = cfg.grid.nc[2]
        cuts = {2: [0, 2, nz], 3: [0, 1, nz - 2, nz], 4: [0, 1, 2, nz - 1, nz]}.get(world)
        if cuts is None:
            cuts = [0, 1, 2, 3] + [nz - (world - 3) + k for k in range(world - 3)] if world > 4 else None
        cuts = np.array(cuts, dtype=np.int32)
        zs = state[2]
        sim = slab.SlabSimulation(cfg, world, zs, transport=slab.DistTransport(), local_ranks=[rank], devices=[local], cuts=cuts,
                                  seed=17, p2p=args.mode == "p2p")
        layer = slab.owner_layer(zs, cfg.grid.edge[2])
        m = (layer >= cuts[rank]) & (layer < cuts[rank + 1])
        sim.set_local_state(np.nonzero(m)[0], *[a[m] for a in state], n_global=len(zs))
    else:
        scale = (args.particles / 557649) ** (1.0 / 3.0)
        cfg = config.pore_config(args.kind == "temp", scale=scale)
        edges = cfg.grid.edge[2]
        zs = init_state.synthetic_pore_chunk(cfg, 17, 0, 1 << 20)[3]
        cuts = slab.balanced_cuts(zs, edges, world)
        sim = slab.SlabSimulation(cfg, world, zs, transport=slab.DistTransport(), local_ranks=[rank], devices=[local], cuts=cuts,
                                  n_total=cfg.num_molecules, seed=17, p2p=args.mode == "p2p")
        sim.init_synthetic(lambda kz: init_state.pore_spec(cfg, 17, keep_z=kz))
    dist.barrier()
    stats = sim.step_fused(args.steps, reduce=True) if args.mode == "p2p" else sim.step(args.steps, reduce=True)
    digest = sim.state_digest()
    per_rank = sim.particles_per_rank()[0]
    sim.close()
    result = {"world": world, "particles": int(cfg.num_molecules if state is None else len(state[0])), "steps": args.steps, "kind": args.kind,
              "mode": args.mode, "dense": bool(args.dense),
              "cuts": [int(c) for c in cuts], "ok": True, "mismatch": []}
    if rank == 0:
        one = amc.Simulation(cfg, seed=17, device=local, max_particles=cfg.num_molecules if state is None else len(state[0]))
        if state is None:
            one.init_synthetic(init_state.pore_spec(cfg, 17))
        else:
            one.set_state(*state)
        ref = one.step(args.steps)
        ref_digest = one.state_digest()
        one.close()
        for k, (a, b) in enumerate(zip(ref, stats)):
            for key in COUNTERS:
                if int(a[key]) != int(b[key]):
                    result["mismatch"].append("step %d %s: single %s, %d ranks %s" % (k, key, a[key], world, b[key]))
            if not np.array_equal(a["wall_hits"], b["wall_hits"]):
                result["mismatch"].append("step %d wall_hits" % k)
            for key in ("dpz", "e_cold", "e_hot"):
                if abs(a[key] - b[key]) > 1e-13 * abs(a[key]):
                    result["mismatch"].append("step %d %s" % (k, key))
        if tuple(ref_digest) != tuple(digest):
            result["mismatch"].append("state digest: single %s, %d ranks %s" % (ref_digest, world, digest))
        result.update(ok=not result["mismatch"], digest=["%016x" % d for d in digest[:2]], digest_count=digest[2],
                      digest_single=["%016x" % d for d in ref_digest[:2]],
                      collisions_per_step=[int(s["collisions"]) for s in stats], resident_rank0=per_rank)
        with open(args.out, "w") as f:
            json.dump(result, f)
        print(json.dumps(result))
    dist.barrier()
    dist.destroy_process_group()
    sys.exit(0 if result["ok"] else 1)


if __name__ == "__main__":
    main()
