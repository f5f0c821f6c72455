"""Host-side logic of the angle sharding (libtike.cufft.dist) on CPU: world_size 2, gloo backend.

The product solver needs a B200, so the sharded driver is exercised here with a stand-in solver
backed by the NumPy oracle (test infrastructure): what is under test is the partitioning, the
ragged tail, the gather and the ScalarComm reductions, not the kernels.
"""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _load_dist_module():
    """libtike.cufft.dist without importing libtike.cufft (whose import needs the CUDA library to
    be built; it is on the CPU box, but keep this test independent of it)."""
    import importlib.util
    path = os.path.join(ROOT, "libtike-cufft_b200", "libtike", "cufft", "dist.py")
    spec = importlib.util.spec_from_file_location("ptx_dist", path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def test_shard_angles_partition():
    d = _load_dist_module()
    for ntheta, world in [(168, 8), (5, 2), (3, 4), (1, 2), (21, 1)]:
        blocks = [d.shard_angles(ntheta, world, r) for r in range(world)]
        covered = [i for b in blocks for i in range(b.start, b.stop)]
        assert covered == list(range(ntheta))
        assert max(b.stop - b.start for b in blocks) == -(-ntheta // world)
    assert [b.stop - b.start for b in [d.shard_angles(168, 8, r) for r in range(8)]] == [21] * 8


def test_weighted_shards():
    """Blocks sized in proportion to per-rank weights (link rates): still a partition of the angle
    axis, largest-remainder rounding, zero weight = empty block."""
    d = _load_dist_module()
    rates = [23.3, 23.3, 23.3, 23.2, 35.5, 35.5, 35.5, 35.4]  # profiles/r02z_pcie8.txt
    counts = d.weighted_counts(168, rates)
    assert counts == [17, 17, 17, 17, 25, 25, 25, 25]
    blocks = [d.shard_angles(168, 8, r, rates) for r in range(8)]
    assert [i for b in blocks for i in range(b.start, b.stop)] == list(range(168))
    assert [b.stop - b.start for b in blocks] == counts
    assert d.weighted_counts(5, [1, 0, 1]) == [3, 0, 2]
    assert d.weighted_counts(0, [1, 2]) == [0, 0]
    assert d.weighted_counts(7, [1, 1]) == [4, 3]
    for bad in ([], [0, 0], [1, -1], [float("nan"), 1]):
        with pytest.raises(ValueError):
            d.weighted_counts(4, bad)
    with pytest.raises(ValueError):
        d.shard_angles(4, 2, 0, [1, 1, 1])


class _OracleSolver(object):
    """Stand-in with the solver's run_batch contract (ptycho.py:135-162), NumPy oracle inside."""

    def __init__(self, nangles):
        self.nangles = nangles

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        pass

    def run_batch(self, data, psi, scan, probe, piter, model="gaussian", recover_prb=False):
        sys.path.insert(0, ROOT)
        from oracle import numpy_ptycho as O
        out_psi, out_prb = psi.copy(), probe.copy()
        for t in range(scan.shape[0]):  # ptheta = 1: every angle is its own problem
            r = O.cg_run(data[t:t + 1], psi[t:t + 1], scan[t:t + 1], probe[t:t + 1].copy(), piter,
                         model, recover_prb)
            out_psi[t], out_prb[t] = r["psi"][0], r["probe"][0]
        return {"psi": out_psi, "probe": out_prb}


def _problem(ntheta):
    sys.path.insert(0, ROOT)
    import workloads
    from oracle import numpy_ptycho as O
    w = workloads.synth_angles(ntheta, 100, 110, 64, 64, 2, 1, seed0=3)
    data = np.abs(O.fwd(w["psi"], w["scan"], np.ascontiguousarray(w["probe"][:, 0]), 64)) ** 2
    return data.astype(np.float32), np.ones_like(w["psi"]), w["scan"], w["probe"] * (0.9 + 0.1j)


def _worker(rank, world, port, ntheta, q, weights=None):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        d = _load_dist_module()
        data, psi, scan, probe = _problem(ntheta)
        res = d.run_batch_sharded(_OracleSolver, data, psi, scan, probe.astype(np.complex64),
                                  weights=weights, piter=2, model="gaussian", recover_prb=True)
        comm = d.ScalarComm()
        s = comm.sum_(torch.tensor([1.0 + rank, 10.0], dtype=torch.float64))
        m = comm.max_(torch.tensor([float(rank)], dtype=torch.float32))
        g = torch.full((2, 2), 1.0 + 1j * rank, dtype=torch.complex64)
        d.ScalarComm(shared_probe=True).probe_grad_(g)
        if rank == 0:
            q.put((res["psi"], res["probe"], s.tolist(), m.tolist(), g[0, 0].item(), comm.calls))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("ntheta,weights", [(3, None), (1, None), (3, [1.0, 2.6])])
def test_run_batch_sharded_world2_gloo(ntheta, weights):
    """2 ranks, ragged split (2 + 1 angles; 1 + 0 angles; weighted 1 + 2) == the unsharded run."""
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29600 + os.getpid() % 300 + ntheta + (7 if weights else 0)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, ntheta, q, weights)) for r in range(2)]
    for p in procs:
        p.start()
    got = q.get(timeout=240)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    data, psi, scan, probe = _problem(ntheta)
    want = _OracleSolver(ntheta).run_batch(data, psi, scan, probe.astype(np.complex64), piter=2,
                                           model="gaussian", recover_prb=True)
    assert np.array_equal(got[0], want["psi"]) and np.array_equal(got[1], want["probe"])
    assert got[2] == [3.0, 20.0] and got[3] == [1.0]
    assert got[4] == (2 + 1j) and got[5] == 2
