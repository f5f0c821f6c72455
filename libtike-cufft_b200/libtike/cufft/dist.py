"""Angle sharding over the GPUs of one box: one process per GPU, `torch.distributed` for plumbing.

The reference has no multi-GPU code at all: `run_batch` walks the angle axis sequentially in chunks
of `ptheta` (src/libtike/cufft/ptycho.py:143-158) and its users start one process per GPU by hand
(tests/catalyst/test_rec_script.py:177).  Angles are independent problems -- object, probe, scan
and data of different angles never meet -- so the B200 layout is:

  * `shard_angles` / `run_batch_sharded`: rank r owns a contiguous block of ceil(ntheta / world)
    angles, keeps their data resident on its GPU for the whole solve, and runs the ordinary
    `run_batch` on it.  No data-path collective; results are gathered once at the end.
  * `ScalarComm`: only when ONE `run` is to span several GPUs (the reference's `ptheta` > angles per
    GPU, where CG step sizes are global over the chunk, ptycho.py:342-343, 370-371) the solver
    all-reduces its packed CG scalars (sum / max, a few doubles per phase) and, in the optional
    shared-probe mode of BASELINE.json's north_star, the probe gradient.  Over NVLink 5 / NVSwitch
    these are latency-bound, so every phase packs its scalars into one buffer = one NCCL call.

Backends: "nccl" on GPUs; the host-side logic (sharding, gather, ragged tails) is exercised on CPU
with "gloo" in tests/test_dist_cpu.py.
"""
import numpy as np
import torch
import torch.distributed as dist

__all__ = ["shard_angles", "run_batch_sharded", "ScalarComm", "bind_to_gpu"]


def bind_to_gpu(local_rank):
    """Best effort: pin the calling process to the CPU cores (hence, by first touch, the host memory)
    next to GPU `local_rank` -- with one process per GPU the pinned staging buffers and the copies they
    feed otherwise all hang off one NUMA node.  Returns the affinity mask applied, or None (no NVML,
    cores outside the container's cpuset, ...); never raises."""
    try:
        import os
        import pynvml
        pynvml.nvmlInit()
        vis = os.environ.get("CUDA_VISIBLE_DEVICES")
        index = int(vis.split(",")[local_rank]) if vis else local_rank
        handle = pynvml.nvmlDeviceGetHandleByIndex(index)
        want = set()
        words = pynvml.nvmlDeviceGetCpuAffinity(handle, (os.cpu_count() + 63) // 64 + 1)
        for w, bits in enumerate(words):
            for b in range(64):
                if bits >> b & 1:
                    want.add(64 * w + b)
        allowed = os.sched_getaffinity(0)
        mask = want & allowed
        if mask and mask != allowed:
            os.sched_setaffinity(0, mask)
            return sorted(mask)
    except Exception:
        pass
    return None


def shard_angles(ntheta, world, rank):
    """Contiguous block of angles owned by `rank`: ceil(ntheta / world) each, the tail may be short
    or empty (168 angles on 8 GPUs -> 21 each)."""
    per = -(-ntheta // world)
    lo = min(rank * per, ntheta)
    return slice(lo, min(lo + per, ntheta))


def _gather_object(obj, group):
    world = dist.get_world_size(group)
    out = [None] * world
    dist.all_gather_object(out, obj, group=group)
    return out


def run_batch_sharded(make_solver, data, psi, scan, probe, group=None, **kwargs):
    """`run_batch` with the angle axis sharded over the ranks of `group`.

    make_solver(nangles) -> a context-manager solver exposing run_batch(data, psi, scan, probe,
    **kwargs) (normally `lambda n: CGPtychoSolver(nscan, nprb, ndet, 1, nz, n)`).  Every rank passes
    the FULL host arrays (or at least valid views of its own block) and receives the FULL result:
    {'psi': [ntheta, nz, n], 'probe': [ntheta, M, P, P]}.  Ranks with an empty block just take part
    in the gather.
    """
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    ntheta = scan.shape[0]
    mine = shard_angles(ntheta, world, rank)
    n_mine = mine.stop - mine.start
    if n_mine > 0:
        with make_solver(n_mine) as slv:
            res = slv.run_batch(data[mine], psi[mine], scan[mine], probe[mine], **kwargs)
        part = (mine.start, np.asarray(res["psi"]), np.asarray(res["probe"]))
    else:
        part = (mine.start, None, None)
    if world == 1:
        return {"psi": part[1], "probe": part[2]}
    parts = _gather_object(part, group)
    out_psi = np.array(psi, copy=True)
    out_prb = np.array(probe, copy=True)
    for lo, p_psi, p_prb in parts:
        if p_psi is not None:
            out_psi[lo:lo + p_psi.shape[0]] = p_psi
            out_prb[lo:lo + p_prb.shape[0]] = p_prb
    return {"psi": out_psi, "probe": out_prb}


class ScalarComm(object):
    """All-reduce hooks the CG solver calls on its packed device scalars.

    comm = ScalarComm(group)            # group=None -> default process group
    slv = CGPtychoSolver(...); slv.comm = comm
    With a comm attached, a set of ranks that each hold some angles behaves like ONE reference `run`
    over all of them: sums (a, b, Dai-Yuan, line-search costs) and maxima (|probe|, |psi|) are global.
    """

    def __init__(self, group=None, shared_probe=False):
        self.group = group
        self.shared_probe = shared_probe
        self.calls = 0

    @property
    def world(self):
        return dist.get_world_size(self.group) if dist.is_initialized() else 1

    @property
    def rank(self):
        return dist.get_rank(self.group) if dist.is_initialized() else 0

    def sum_(self, t):
        if self.world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.SUM, group=self.group)
            self.calls += 1
        return t

    def max_(self, t):
        if self.world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX, group=self.group)
            self.calls += 1
        return t

    def min_(self, t):
        if self.world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MIN, group=self.group)
            self.calls += 1
        return t

    def probe_grad_(self, g):
        """Shared-probe mode: one probe for every angle -> its gradient is summed over ranks."""
        if self.shared_probe and self.world > 1:
            buf = torch.view_as_real(g)
            dist.all_reduce(buf, op=dist.ReduceOp.SUM, group=self.group)
            self.calls += 1
        return g
