"""Two-GPU (NCCL) check of the coupled CG: 2 ranks x 1 angle with a ScalarComm == 1 rank, ptheta = 2.

Skipped unless two GPUs are visible (run with `gpurun --gpus 2 -- python -m pytest tests/test_gpu_dist.py -m gpu`).
"""
import os
import sys

import numpy as np
import pytest
import torch

import workloads
from oracle import numpy_ptycho as O
from util import rel_l2

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _problem():
    w = workloads.synth_angles(2, 200, 220, 64, 64, 5, 1, seed0=11)
    data = np.abs(O.fwd(w["psi"], w["scan"], np.ascontiguousarray(w["probe"][:, 0]), 64)) ** 2
    probe = w["probe"] * (0.9 + 0.1j)
    probe[1] *= 1.3  # make the angles differ so that the global scalars matter
    return data.astype(np.float32), np.ones_like(w["psi"]), w["scan"], probe.astype(np.complex64)


def _worker(rank, world, port, q):
    import torch.distributed as dist
    for p in (ROOT, os.path.join(ROOT, "libtike-cufft_b200")):
        if p not in sys.path:
            sys.path.insert(0, p)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        import libtike.cufft as pt
        from libtike.cufft.dist import ScalarComm
        data, psi, scan, probe = _problem()
        sl = slice(rank, rank + 1)
        with pt.CGPtychoSolver(25, 64, 64, 1, 200, 220) as slv:
            slv.comm = ScalarComm()
            res = slv.run_batch(data[sl], psi[sl], scan[sl], probe[sl], piter=4, model="gaussian",
                                recover_prb=True)
            q.put((rank, res["psi"], res["probe"], slv.history, slv.comm.calls))
        dist.barrier()
    finally:
        dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs")
def test_coupled_cg_two_ranks_equals_ptheta2():
    import torch.multiprocessing as mp
    import libtike.cufft as pt
    data, psi, scan, probe = _problem()
    with pt.CGPtychoSolver(25, 64, 64, 2, 200, 220) as slv:  # one run over both angles
        want = slv.run_batch(data, psi, scan, probe, piter=4, model="gaussian", recover_prb=True)
        hist = slv.history
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, 29750 + os.getpid() % 200, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = sorted([q.get(timeout=300) for _ in range(2)], key=lambda x: x[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for r in range(2):
        assert got[r][3] == hist                       # identical step decisions on every rank
        assert rel_l2(got[r][1], want["psi"][r:r + 1]) < 1e-5
        assert rel_l2(got[r][2], want["probe"][r:r + 1]) < 1e-5
        assert got[r][4] > 0                           # collectives actually ran
