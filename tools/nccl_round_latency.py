"""Latency of one boundary hand-over round (torchrun, N ranks): batch P2P vs all_to_all_single."""
import os, time, sys
import torch, torch.distributed as dist
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
dist.init_process_group("nccl", device_id=torch.device("cuda", local)); torch.cuda.set_device(local)
cap = 2049 * 12
su, sd, ru, rd = [torch.zeros(cap, dtype=torch.float64, device="cuda") for _ in range(4)]
ops = []
if rank + 1 < world: ops += [dist.P2POp(dist.isend, su, rank + 1), dist.P2POp(dist.irecv, ru, rank + 1)]
if rank > 0: ops += [dist.P2POp(dist.isend, sd, rank - 1), dist.P2POp(dist.irecv, rd, rank - 1)]
a2a_s, a2a_r = torch.zeros(world, cap, dtype=torch.float64, device="cuda"), torch.zeros(world, cap, dtype=torch.float64, device="cuda")
def run(fn, n=200):
    for _ in range(20): fn()
    torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter(); e0.record()
    for _ in range(n): fn()
    e1.record(); host = time.perf_counter() - t0; torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3, host / n * 1e6
def p2p():
    for r in dist.batch_isend_irecv(ops): r.wait()
def a2a():
    dist.all_to_all_single(a2a_r, a2a_s)
small = torch.zeros(8, device="cuda")
def tiny_kernel():
    small.add_(1)
for name, fn in (("batch_p2p", p2p), ("all_to_all_single", a2a), ("tiny_kernel", tiny_kernel)):
    gpu_us, host_us = run(fn)
    if rank == 0: print(f"{name:20s} gpu {gpu_us:8.1f} us/round   host {host_us:8.1f} us/round", file=sys.stderr)
dist.destroy_process_group()
