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

__all__ = ["shard_angles", "weighted_counts", "link_rates", "run_batch_sharded", "ScalarComm", "bind_to_gpu"]


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


def weighted_counts(ntheta, weights):
    """Split `ntheta` angles over len(weights) ranks in proportion to `weights` (largest-remainder
    rounding, ties to the lower rank): [count of rank 0, count of rank 1, ...], summing to ntheta."""
    w = np.asarray(weights, dtype=np.float64)
    if w.ndim != 1 or w.size == 0 or not np.all(np.isfinite(w)) or np.any(w < 0) or w.sum() <= 0:
        raise ValueError("weights must be non-negative, finite and not all zero")
    quota = ntheta * w / w.sum()
    counts = np.floor(quota).astype(np.int64)
    order = np.argsort(-(quota - counts), kind="stable")
    counts[order[:ntheta - int(counts.sum())]] += 1
    return [int(c) for c in counts]


def shard_angles(ntheta, world, rank, weights=None):
    """Contiguous block of angles owned by `rank`: ceil(ntheta / world) each, the tail may be short
    or empty (168 angles on 8 GPUs -> 21 each).  With `weights` (one per rank, e.g. `link_rates()`)
    the blocks are sized in proportion to them instead: when the host-array entry points are bound
    by the host-to-device links and those are not alike, equal shards finish with the slowest link."""
    if weights is not None:
        if len(weights) != world:
            raise ValueError("one weight per rank")
        counts = weighted_counts(ntheta, weights)
        lo = sum(counts[:rank])
        return slice(lo, lo + counts[rank])
    per = -(-ntheta // world)
    lo = min(rank * per, ntheta)
    return slice(lo, min(lo + per, ntheta))


def link_rates(group=None, nbytes=64 << 20, reps=8):
    """Host-to-device copy rate (GB/s) of every rank's GPU while ALL ranks copy at once, from pinned
    memory: [rate of rank 0, rate of rank 1, ...], identical on every rank.  On a box whose GPUs share
    PCIe switches or root ports unevenly these differ (measured on an 8 x B200 VM: 23 GB/s on four
    GPUs, 35 GB/s on the other four, 55 GB/s for any one of them alone; profiles/r02z_pcie8.txt) --
    the weights `shard_angles` / `run_batch_sharded` take."""
    src = torch.empty(nbytes, dtype=torch.uint8).pin_memory()
    dst = torch.empty(nbytes, dtype=torch.uint8, device="cuda")
    dst.copy_(src, non_blocking=True)
    torch.cuda.synchronize()
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    if world > 1:
        dist.barrier(group)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        dst.copy_(src, non_blocking=True)
    e1.record()
    e1.synchronize()
    rate = reps * nbytes / (e0.elapsed_time(e1) * 1e-3) / 1e9
    if world == 1:
        return [rate]
    return [float(r) for r in _gather_object(rate, group)]


def _gather_object(obj, group):
    world = dist.get_world_size(group)
    out = [None] * world
    dist.all_gather_object(out, obj, group=group)
    return out


def run_batch_sharded(make_solver, data, psi, scan, probe, group=None, weights=None, **kwargs):
    """`run_batch` with the angle axis sharded over the ranks of `group`.

    make_solver(nangles) -> a context-manager solver exposing run_batch(data, psi, scan, probe,
    **kwargs) (normally `lambda n: CGPtychoSolver(nscan, nprb, ndet, 1, nz, n)`).  Every rank passes
    the FULL host arrays (or at least valid views of its own block) and receives the FULL result:
    {'psi': [ntheta, nz, n], 'probe': [ntheta, M, P, P]}.  Ranks with an empty block just take part
    in the gather.  `weights`: one per rank (the same list on every rank), see `shard_angles`.
    """
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    ntheta = scan.shape[0]
    mine = shard_angles(ntheta, world, rank, weights)
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
