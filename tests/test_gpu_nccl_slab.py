"""The real multi-process path on real GPUs: N processes (one per GPU, torch.distributed.run, NCCL) run one
slab-decomposed pore -- through amc_slab_step (peer-to-peer transfers by kernels) and through the step-wise NCCL
path -- and must reproduce the single-GPU run of the same job: per-step counters and the
order-independent checksum of the id-ordered state (tests/nccl_slab_worker.py).  Needs >= 2 GPUs; on a
one-GPU box it is skipped (the device side of the protocol is then covered by tests/test_gpu_slab.py, the
transport by tests/test_slab_transport.py).  profiles/ holds the result files of the 2/4/8-GPU runs."""
import json
import os
import socket
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


@pytest.mark.parametrize("kind,mode", [("temp", "p2p"), ("pore", "p2p"), ("temp", "nccl")])
def test_nccl_ranks_match_single_gpu(tmp_path, kind, mode):
    import torch
    ngpu = torch.cuda.device_count()
    if ngpu < 2:
        pytest.skip("needs >= 2 GPUs (found %d)" % ngpu)
    world = min(ngpu, 8)
    out = tmp_path / "nccl.json"
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(world), "--master-addr",
           "127.0.0.1", "--master-port", str(_free_port()), os.path.join(ROOT, "tests", "nccl_slab_worker.py"), str(out),
           "--kind", kind, "--mode", mode, "--particles", "2000000", "--steps", "6"]
    r = subprocess.run(cmd, cwd=ROOT, capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    res = json.loads(out.read_text())
    assert res["ok"] and res["digest"] == res["digest_single"] and res["digest_count"] == res["particles"]


def test_nccl_dense_cuts_many_handovers(tmp_path):
    """The stress case of tests/test_gpu_slab.py on real GPUs through amc_slab_step: small dense pore, every cut inside an
    end cap (even cuts: the round before the first group is exercised), dozens of hand-over records per colour group."""
    import torch
    ngpu = torch.cuda.device_count()
    if ngpu < 2:
        pytest.skip("needs >= 2 GPUs (found %d)" % ngpu)
    world = min(ngpu, 4)
    out = tmp_path / "nccl_dense.json"
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(world), "--master-addr",
           "127.0.0.1", "--master-port", str(_free_port()), os.path.join(ROOT, "tests", "nccl_slab_worker.py"), str(out),
           "--kind", "pore", "--mode", "p2p", "--dense", "--steps", "30"]
    r = subprocess.run(cmd, cwd=ROOT, capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    res = json.loads(out.read_text())
    assert res["ok"] and res["digest"] == res["digest_single"] and res["digest_count"] == res["particles"]
