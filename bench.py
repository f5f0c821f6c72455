#!/usr/bin/env python
"""Headline benchmark: diffraction patterns/s through the fused fwd -> residual -> adj pass.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference] [--workload c4|c2]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
        --master-port P bench.py --gpus N --steps K --warmup W

A "step" is one pass of the hot path (a2 fwd + the Gaussian residual of
ptycho.py:351-356 + a3 adj, i.e. the CG object gradient) over the WHOLE angle batch:

  workload c4 (default; BASELINE.json configs[3], the one north_star's targets are quoted on):
      168 angles x 1024 scan positions, 1024x1024 object slices, 256x256 detector, 1 probe mode,
      Gaussian model = 172 032 patterns and 45 GB of measured data per step.  The batch is FIXED:
      N GPUs hold 168 / N angles each ("scaling": "strong"; N = 1 holds all 168 in one B200's HBM).
  workload c2 (configs[1]): 8 angles x 1024 positions, 512x512 object, 128x128 detector per GPU.
      Always measured as a sub-record (`workloads.c2`) of the default line.

Angles are independent problems (ptheta = 1 semantics of every reference test), so the shards need
no data-path collective.  The only collective the path has -- the all-reduce of the CG scalars when
ONE run spans GPUs -- is timed separately (`cg.coupled`) and checked against a single-GPU
ptheta = 2 run (`cg.scalarcomm_check`) whenever N >= 2.

  value     patterns/s with every input resident in HBM (CUDA events, max over ranks)
  e2e       the same pass through the host-array API (`CGPtychoSolver.grad_ptycho_batch`): host
            buffers in, host gradient out, H2D / D2H inside the timed region; `value` with pinned
            host arrays, `pageable` with ordinary NumPy arrays (what a drop-in caller holds)
  e2e_cg    `run_batch(piter=8, recover_prb=True)` -- the reference's own host entry point
            (ptycho.py:135-162) -- in angle-iterations/s through host arrays
  roofline  the fused kernel k_grad<gaussian, object>: SURVEY.md section 8d says FP32-bound, so
            achieved = algorithmic FFT flops per launch / CUDA-event launch time against the nominal
            FP32 peak; the HBM view (algorithmic bytes against MEASURED_PEAKS.json) and the
            L1/shared data-pipe view are sub-keys
  cpu_baseline  the NumPy/pocketfft oracle on the box's host cores, bounded sample (rank 0, N = 1)
  cg        CG iterations/s of `CGPtychoSolver.run` (object + probe recovery), per workload
  workloads c2, the 5-mode C3 shape and the C5 Poisson detector sweep, same pass

--impl reference runs the REFERENCE's own CUDA/cuFFT operators (oracle/_ref, compiled unmodified
from /root/reference/src/cuda) for the same pass and config: fwd -> torch elementwise (the CuPy
statements of ptycho.py:351-356) -> adj, batched at ptheta = 21 angles per call.  The reference has
no CPU implementation of this path; if oracle/_ref did not travel, the NumPy port is timed instead
and the line says so.  Rank 0 alone runs it; each step is a bounded 21-angle sample of the batch.
"""
import argparse
import contextlib
import io
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
PKG = os.path.join(ROOT, "libtike-cufft_b200")
for p in (ROOT, PKG):
    if p not in sys.path:
        sys.path.insert(0, p)

import numpy as np  # noqa: E402
import torch  # noqa: E402

import workloads  # noqa: E402

FP32_NOMINAL_TFLOPS = 148 * 128 * 2 * 1.965e9 / 1e12  # 74.4, SURVEY.md section 8d
C4_ANGLES = 168
REF_ANGLES = 21   # angles per step (and per call) of the reference arm
E2E_ANGLES = 21   # host-array legs: at most this many angles per rank (5.6 GB of pinned data at c4)
CG_ANGLES = 8     # e2e_cg / sharded CG: angles per rank
CG_ITERS = 8


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        return float(json.load(open(path))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler(threading.Thread):
    """SM clock / throttle reasons of one GPU while the timed region runs (NVML, 5 ms period)."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.samples, self.reasons, self.max_mhz = index, [], set(), None
        self.mem, self.power = [], []
        self._stop_evt = threading.Event()
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def run(self):
        if self.nv is None:
            return
        nv = self.nv
        names = {nv.nvmlClocksEventReasonHwSlowdown: "hw_slowdown",
                 nv.nvmlClocksEventReasonHwThermalSlowdown: "hw_thermal_slowdown",
                 nv.nvmlClocksEventReasonSwThermalSlowdown: "sw_thermal_slowdown",
                 nv.nvmlClocksEventReasonSwPowerCap: "sw_power_cap",
                 nv.nvmlClocksEventReasonHwPowerBrakeSlowdown: "hw_power_brake",
                 nv.nvmlClocksEventReasonApplicationsClocksSetting: "applications_clocks_setting"}
        while not self._stop_evt.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                if len(self.samples) % 16 == 1:  # memory clock and power: what differs between two boxes
                    try:                         # whose SM clocks read alike
                        self.mem.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_MEM))
                        self.power.append(nv.nvmlDeviceGetPowerUsage(self.h) / 1000.0)
                    except Exception:
                        pass
                r = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                for bit, name in names.items():
                    if r & bit:
                        self.reasons.add(name)
            except Exception:
                break
            time.sleep(0.005)

    def stop(self):
        self._stop_evt.set()
        self.join(timeout=2)
        s = self.samples
        return {"sm_mhz": float(np.median(s)) if s else None, "sm_max_mhz": self.max_mhz,
                "reasons": sorted(self.reasons), "samples": len(s),
                "mem_mhz": float(np.median(self.mem)) if self.mem else None,
                "power_w": float(np.median(self.power)) if self.power else None}


def physical_gpu_index(local):
    vis = os.environ.get("CUDA_VISIBLE_DEVICES")
    if vis:
        try:
            return int(vis.split(",")[local])
        except Exception:
            return local
    return local


# ---------------------------------------------------------------------------------------------
# workloads and the config record (ONE function for both arms: the dicts must be identical)
# ---------------------------------------------------------------------------------------------
def total_angles(name, world):
    return C4_ANGLES if name == "c4" else 8 * world


def make_config(name, world):
    if name == "c4":
        return {"workload": "c4: 168 angles x 1024 positions, 1024x1024 object slices, 256x256 detector, "
                            "1 mode, gaussian (BASELINE.json configs[3])",
                "patterns_per_step": C4_ANGLES * 1024,
                "angles_per_gpu": C4_ANGLES // world,
                "l2": "45 GB of measured data per step (>> 126 MB L2): no flush needed",
                "parallelism": "angle shards (168 / N per GPU), no data-path collective"}
    return {"workload": "c2: 8 angles/GPU x 1024 positions, 512x512 object, 128x128 detector, 1 mode, "
                        "gaussian (BASELINE.json configs[1])",
            "patterns_per_step": 8 * world * 1024,
            "angles_per_gpu": 8,
            "l2": "537 MB of measured data per GPU and step (> 126 MB L2): no flush needed",
            "parallelism": "angle shards (8 per GPU), no data-path collective"}


def shard(name, world, rank):
    """(first global angle, number of angles) of this rank."""
    if name == "c4":
        per = C4_ANGLES // world
        return rank * per, per
    return rank * 8, 8


def make_angles(name, first, count):
    """Seeded inputs of `count` angles starting at global angle `first` (seed = angle index)."""
    if name == "c4":
        return workloads.synth_angles(count, 1024, 1024, 256, 256, 32, 1, seed0=first)
    return workloads.synth_angles(count, 512, 512, 128, 128, 32, 1, seed0=first)


def algorithmic_bytes(w, T):
    """measured data 4 N^2 + scan 8 per pattern; object read + gradient write
    16 nz n and probe read 8 M P^2 per angle (the fused pass reads the probe once, writes no probe)."""
    N, S, nz, n, P, M = w["ndet"], w["nscan"], w["nz"], w["n"], w["nprb"], w["nmodes"]
    return T * (S * (4 * N * N + 8) + 16 * nz * n + 8 * M * P * P)


def algorithmic_flops(w, T):
    """20 N^2 log2 N per pattern and mode (forward + inverse 2-D FFT), SURVEY.md section 8d."""
    N = w["ndet"]
    return T * w["nscan"] * w["nmodes"] * 20.0 * N * N * np.log2(N)


def dist_setup(ngpus):
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        import torch.distributed as dist
        from libtike.cufft.dist import bind_to_gpu
        bind_to_gpu(local)  # host staging next to the GPU it feeds (NUMA), best effort
        torch.cuda.set_device(local)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    else:
        torch.cuda.set_device(0)
    if world != ngpus and rank == 0 and world == 1 and ngpus > 1:
        raise SystemExit("--gpus %d needs torchrun with %d ranks (see the module docstring)" % (ngpus, ngpus))
    return world, rank, local


def barrier(world):
    if world > 1:
        import torch.distributed as dist
        dist.barrier()
    torch.cuda.synchronize()


def max_over_ranks(x, world):
    if world == 1:
        return x
    import torch.distributed as dist
    t = torch.tensor([x], dtype=torch.float64, device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def timed_steps(step, steps, warmup, world):
    """W warm-ups, then exactly K steps between barrier+synchronize, timed with CUDA events on the
    launching stream; returns seconds (max over ranks)."""
    for _ in range(warmup):
        step()
    barrier(world)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        step()
    e1.record()
    barrier(world)
    return max_over_ranks(e0.elapsed_time(e1), world) * 1e-3


def wall_steps(step, steps, warmup, world):
    """Host-clock variant for the legs whose work includes host-side copies."""
    for _ in range(warmup):
        step()
    barrier(world)
    t0 = time.perf_counter()
    for _ in range(steps):
        step()
    barrier(world)
    return max_over_ranks(time.perf_counter() - t0, world)


def pinned(a):
    return torch.from_numpy(np.ascontiguousarray(a)).pin_memory().numpy()


def cpu_port_rate(w, seconds_target=12.0):
    """NumPy/pocketfft oracle (fwd -> residual -> adj) on the host cores, bounded sample."""
    from oracle import numpy_ptycho as O
    ns = 64 if w["ndet"] <= 128 else 24
    psi, scan, prb = w["psi"][:1], w["scan"][:1, :ns], np.ascontiguousarray(w["probe"][:1, 0])
    N = w["ndet"]
    data = np.abs(O.fwd(psi, scan, prb, N)) ** 2
    psi1 = np.ones_like(psi)
    t0 = time.time()
    reps = 0
    while True:
        f = O.fwd(psi1, scan, prb, N)
        r = f - np.sqrt(data) * f / (np.sqrt(np.abs(f) ** 2) + np.float32(1e-32))
        O.adj(r.astype(np.complex64), scan, prb, w["nz"], w["n"])
        reps += 1
        if time.time() - t0 > seconds_target:
            break
    dt = time.time() - t0
    return reps * ns / dt, "%d x %d patterns of one angle (%.1f s)" % (reps, ns, dt)


def quiet():
    return contextlib.redirect_stdout(io.StringIO())


# ---------------------------------------------------------------------------------------------
# this repository's arm
# ---------------------------------------------------------------------------------------------
def synth_data(slv, psi, scan, probe, chunk=8):
    """|fwd|^2 summed over modes, angle chunks of `chunk` (bounds the temporary far field)."""
    import libtike.cufft as pt
    T, S, N = psi.shape[0], scan.shape[1], slv.ndet
    data = torch.empty((T, S, N, N), dtype=torch.float32, device=psi.device)
    for t0 in range(0, T, chunk):
        t1 = min(T, t0 + chunk)
        with pt.PtychoCuFFT(S, slv.nprb, N, t1 - t0, slv.nz, slv.n) as op:
            acc = None
            for k in range(probe.shape[1]):
                f = op.fwd(psi[t0:t1].contiguous(), scan[t0:t1].contiguous(),
                           probe[t0:t1, k].contiguous()).abs().square_()
                acc = f if acc is None else acc.add_(f)
            data[t0:t1] = acc
    return data


def grad_rate(pt, w, T, steps, warmup, world=1, model=0, counts=0.0):
    """Device-resident rate of the fused gradient pass on workload dict `w` (T angles): returns
    (patterns/s on this rank, mean launch seconds of the dominant kernel, launches per step)."""
    from libtike.cufft.ptychofft import launch_count
    dev = torch.device("cuda", torch.cuda.current_device())
    S, N, nz, n, M = w["nscan"], w["ndet"], w["nz"], w["n"], w["nmodes"]
    psi_true, scan, probe = (torch.from_numpy(w[k]).to(dev) for k in ("psi", "scan", "probe"))
    with pt.CGPtychoSolver(S, w["nprb"], N, T, nz, n) as slv:
        data = synth_data(slv, psi_true, scan, probe)
        if counts:
            data = torch.poisson(data * (counts / data.mean())).contiguous()
        psi = torch.ones_like(psi_true)
        grad = torch.zeros_like(psi)
        inten = torch.empty_like(data) if M > 1 else None
        sc = torch.ones(3, dtype=torch.float32, device=dev)

        def step():
            grad.zero_()
            if M > 1:
                slv._intensity(psi, scan, probe, data, inten, model)
            for k in range(M):
                slv._grad(0, psi, scan, probe, k, data, inten, 1.0, 1.0, 1.0, model, grad, sc=sc)

        for _ in range(warmup):
            step()
        torch.cuda.synchronize()
        l0 = launch_count()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            step()
        e1.record()
        e1.synchronize()
        secs = e0.elapsed_time(e1) * 1e-3
        launches = (launch_count() - l0) // steps
    return T * S * steps / secs, secs / steps, launches


def cg_rates(pt, w, data1, psi1, scan1, probe1, iters=16):
    """CG iterations/s of one device-resident angle, position correction on (the reference's
    behaviour) and off; best of two runs each."""
    S, N, nz, n = w["nscan"], w["ndet"], w["nz"], w["n"]
    out = {"iters": iters, "recover_prb": True}
    with pt.CGPtychoSolver(S, w["nprb"], N, 1, nz, n) as s1:
        for key, on in (("iters_per_s", True), ("iters_per_s_no_position_correction", False)):
            s1.position_correction = on
            with quiet():
                s1.run(data1, psi1, scan1.clone(), probe1.clone(), piter=2, recover_prb=True)
                dt = None
                for _ in range(2):
                    sc1, pr1 = scan1.clone(), probe1.clone()
                    torch.cuda.synchronize()
                    t0 = time.perf_counter()
                    s1.run(data1, psi1, sc1, pr1, piter=iters, recover_prb=True)
                    torch.cuda.synchronize()
                    t1 = time.perf_counter() - t0
                    dt = t1 if dt is None else min(dt, t1)
            out[key] = iters / dt
            if on:
                out["line_search_passes_per_iter"] = len(s1.ls_log) / float(iters)
    return out


def scalarcomm_check(pt, world, rank):
    """2 ranks x 1 angle with a ScalarComm == 1 rank with ptheta = 2 (the NCCL all-reduce of the CG
    scalars), on ranks 0 and 1; returns the record on rank 0."""
    import torch.distributed as dist
    from libtike.cufft.dist import ScalarComm
    w = workloads.synth_angles(2, 200, 220, 64, 64, 5, 1, seed0=11)
    dev = torch.device("cuda", torch.cuda.current_device())
    psi_t, scan, probe = (torch.from_numpy(w[k]).to(dev) for k in ("psi", "scan", "probe"))
    with pt.PtychoCuFFT(25, 64, 64, 2, 200, 220) as op:
        data = op.fwd(psi_t, scan, probe[:, 0].contiguous()).abs().square_().contiguous()
    probe = probe * (0.9 + 0.1j)
    probe[1] *= 1.3
    psi0 = torch.ones_like(psi_t)
    group = dist.new_group([0, 1])
    rec = None
    if rank < 2:
        sl = slice(rank, rank + 1)
        with pt.CGPtychoSolver(25, 64, 64, 1, 200, 220) as slv, quiet():
            slv.comm = ScalarComm(group)
            res = slv.run(data[sl].contiguous(), psi0[sl].contiguous(), scan[sl].clone(),
                          probe[sl].clone(), piter=4, recover_prb=True)
            calls, hist = slv.comm.calls, list(slv.history)
        mine = torch.view_as_real(res["psi"].contiguous())
        parts = [torch.zeros_like(mine) for _ in range(2)]
        dist.all_gather(parts, mine, group=group)
        both = [torch.view_as_complex(x) for x in parts]
        if rank == 0:
            with pt.CGPtychoSolver(25, 64, 64, 2, 200, 220) as slv, quiet():
                want = slv.run(data, psi0, scan.clone(), probe.clone(), piter=4, recover_prb=True)
                same_steps = list(slv.history) == hist
            err = float(torch.linalg.norm(torch.cat(both) - want["psi"]) / torch.linalg.norm(want["psi"]))
            rec = {"rel_l2_psi_vs_ptheta2": err, "same_step_decisions": bool(same_steps),
                   "allreduces": int(calls), "ok": bool(err < 1e-5 and same_steps), "backend": "nccl"}
    dist.barrier()
    return rec


def coupled_cg(pt, w, data, psi, scan, probe, world, iters=CG_ITERS):
    """ONE run spanning all ranks (one angle each, CG scalars all-reduced over NCCL): iterations/s."""
    from libtike.cufft.dist import ScalarComm
    S, N, nz, n = w["nscan"], w["ndet"], w["nz"], w["n"]
    with pt.CGPtychoSolver(S, w["nprb"], N, 1, nz, n) as slv, quiet():
        slv.comm = ScalarComm()
        slv.run(data[:1], psi[:1], scan[:1].clone(), probe[:1].clone(), piter=2, recover_prb=True)
        c0 = slv.comm.calls
        barrier(world)
        t0 = time.perf_counter()
        slv.run(data[:1], psi[:1], scan[:1].clone(), probe[:1].clone(), piter=iters, recover_prb=True)
        barrier(world)
        dt = max_over_ranks(time.perf_counter() - t0, world)
        calls = slv.comm.calls - c0
    return {"iters_per_s": iters / dt, "angles": world, "allreduces_per_iter": calls / float(iters),
            "note": "one CG run over %d angles, one per GPU; sums / maxima of every phase all-reduced" % world}


def run_b200(args, world, rank, local):
    import libtike.cufft as pt
    from libtike.cufft.ptychofft import launch_count
    first, T = shard(args.workload, world, rank)
    w = make_angles(args.workload, first, T)
    S, N, nz, n = w["nscan"], w["ndet"], w["nz"], w["n"]
    npat = T * S
    dev = torch.device("cuda", torch.cuda.current_device())
    psi_true, scan, probe = (torch.from_numpy(w[k]).to(dev) for k in ("psi", "scan", "probe"))
    hbm, peak_src = peaks()
    extra = {}
    with pt.CGPtychoSolver(S, w["nprb"], N, T, nz, n) as slv:
        data = synth_data(slv, psi_true, scan, probe)   # synthetic measurement
        psi = torch.ones_like(psi_true)                 # the solver's starting point (tests/test.py:55)
        grad = torch.zeros_like(psi)
        sc = torch.ones(3, dtype=torch.float32, device=dev)

        def step():
            grad.zero_()
            slv._grad(0, psi, scan, probe, 0, data, None, 1.0, 1.0, 1.0, 0, grad, sc=sc)

        sampler = ClockSampler(physical_gpu_index(local))
        for _ in range(args.warmup):
            step()
        barrier(world)
        sampler.start()
        l1 = launch_count()
        secs = timed_steps(step, args.steps, 0, world)
        launches = launch_count() - l1
        clocks = sampler.stop()
        value = total_angles(args.workload, world) * S * args.steps / secs

        # dominant kernel alone: events straight around each launch (same stream)
        kt = []
        for _ in range(min(args.steps, 10)):
            grad.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            slv._grad(0, psi, scan, probe, 0, data, None, 1.0, 1.0, 1.0, 0, grad, sc=sc)
            e1.record()
            e1.synchronize()
            kt.append(e0.elapsed_time(e1) * 1e-3)
        kavg = float(np.mean(kt))
    abytes, aflops = algorithmic_bytes(w, T), algorithmic_flops(w, T)

    # ---- end to end through the host-array API.  The sample is E angles per rank (world * E in all).
    # This leg is bound by the host-to-device links, and on a multi-GPU box those are not alike when all
    # GPUs copy at once (profiles/r02z_pcie8.txt: 23 GB/s on four GPUs, 35 GB/s on the other four), so for
    # N > 1 the sample is split in proportion to each rank's measured link rate (libtike.cufft.dist).
    E = min(T, E2E_ANGLES)
    counts, rates = [E] * world, None
    if world > 1:
        from libtike.cufft.dist import link_rates, weighted_counts
        rates = link_rates()
        counts = weighted_counts(world * E, rates)
        if min(counts) < 1:  # a link that measured (almost) nothing: fall back to equal shards
            counts, rates = [E] * world, None
    if counts[rank] == E and counts == [E] * world:
        host = {"data": data[:E].cpu().numpy(), "psi": psi[:E].cpu().numpy(),
                "scan": w["scan"][:E], "probe": w["probe"][:E]}
    else:
        # global angles of the sample, in rank order; this rank takes its weighted block of them
        sample = [shard(args.workload, world, r)[0] + k for r in range(world) for k in range(E)]
        lo = sum(counts[:rank])
        mine = sample[lo:lo + counts[rank]]
        assert mine, "a rank with a measured link rate always gets at least one angle"
        parts = {"data": [], "psi": [], "scan": [], "probe": []}
        i = 0
        while i < len(mine):  # contiguous runs of global angles
            j = i
            while j + 1 < len(mine) and mine[j + 1] == mine[j] + 1 and j + 1 - i < 8:
                j += 1
            wr = make_angles(args.workload, mine[i], j + 1 - i)
            tr = [torch.from_numpy(wr[k]).to(dev) for k in ("psi", "scan", "probe")]
            with pt.CGPtychoSolver(S, wr["nprb"], N, j + 1 - i, nz, n) as sr:
                parts["data"].append(synth_data(sr, *tr).cpu().numpy())
            parts["psi"].append(np.ones_like(wr["psi"]))
            parts["scan"].append(wr["scan"])
            parts["probe"].append(wr["probe"])
            del tr
            i = j + 1
        host = {k: np.concatenate(v) for k, v in parts.items()}
        del parts
    e2e = {}
    esteps = max(2, args.steps // 4)
    with pt.CGPtychoSolver(S, w["nprb"], N, 1, nz, n) as s1:
        for kind in ("pinned", "pageable"):
            h = {k: (pinned(v) if kind == "pinned" else np.ascontiguousarray(v)) for k, v in host.items()}

            def e2e_step():
                if h["scan"].shape[0]:
                    s1.grad_ptycho_batch(h["data"], h["psi"], h["scan"], h["probe"], model="gaussian")
            dt = wall_steps(e2e_step, esteps, 1, world)
            e2e[kind] = world * E * S * esteps / dt
            del h
    # bytes per step and rank, averaged over the ranks (every angle is the same size)
    h2d = sum(v.nbytes for v in host.values()) // counts[rank] * E
    d2h = host["psi"].nbytes // counts[rank] * E

    # ---- the reference's host entry point: run_batch (ptycho.py:135-162), angle-iterations/s
    C = min(T, CG_ANGLES, min(counts))
    with pt.CGPtychoSolver(S, w["nprb"], N, 1, nz, n) as s1, quiet():
        hc = {k: v[:C] for k, v in host.items()}
        s1.run_batch(hc["data"][:1], hc["psi"][:1], hc["scan"][:1], hc["probe"][:1], piter=2,
                     recover_prb=True)
        dt = None
        for _ in range(2):  # best of two: a fresh box still pages code in from the image (100 ms stalls)
            barrier(world)
            t0 = time.perf_counter()
            s1.run_batch(hc["data"], hc["psi"], hc["scan"], hc["probe"], piter=CG_ITERS, recover_prb=True)
            barrier(world)
            t1 = max_over_ranks(time.perf_counter() - t0, world)
            dt = t1 if dt is None else min(dt, t1)
    e2e_cg = {"value": world * C * CG_ITERS / dt, "unit": "angle-iterations/s",
              "api": "CGPtychoSolver.run_batch(piter=%d, recover_prb=True), pageable host arrays" % CG_ITERS,
              "sample": "%d angles per GPU, best of two calls" % C}
    del host

    # ---- CG iterations/s
    cg = None
    if rank == 0 and not args.no_cg:
        cg = cg_rates(pt, w, data[:1].contiguous(), psi[:1].contiguous(), scan[:1].contiguous(),
                      probe[:1].contiguous())
        cg["config"] = "one %s angle, device resident" % args.workload
    if world > 1:
        barrier(world)
        coupled = coupled_cg(pt, w, data, psi, scan, probe, world)
        check = scalarcomm_check(pt, world, rank)
        if rank == 0:
            cg = cg or {}
            cg["coupled"] = coupled
            cg["scalarcomm_check"] = check
    del data, grad, psi
    torch.cuda.empty_cache()

    # ---- other configurations of BASELINE.json through the same pass (rank 0)
    if rank == 0 and not args.no_extras:
        def rec(wd, T_, model=0, counts=0.0, steps=5):
            r, ksec, nl = grad_rate(pt, wd, T_, steps, 3, model=model, counts=counts)
            return {"value": r, "unit": "patterns/s", "fp32_frac": algorithmic_flops(wd, T_) / ksec / 1e12 / FP32_NOMINAL_TFLOPS,
                    "hbm_gbs": algorithmic_bytes(wd, T_) / ksec / 1e9, "launches_per_step": nl}
        extras = {}
        if args.workload != "c2":
            w2 = make_angles("c2", 0, 8)
            extras["c2"] = rec(w2, 8, steps=10)
            extras["c2"]["config"] = make_config("c2", 1)["workload"]
            d2 = None
            with pt.CGPtychoSolver(1024, 128, 128, 1, 512, 512) as s2:
                p2, sc2, pr2 = (torch.from_numpy(w2[k][:1]).to(dev) for k in ("psi", "scan", "probe"))
                d2 = synth_data(s2, p2, sc2, pr2)
            extras["c2"]["cg"] = cg_rates(pt, w2, d2, torch.ones_like(p2), sc2, pr2)
            del d2
        c3 = workloads.c3_modes(5, 1100)
        c3["probe"] = c3["probe_init"]
        extras["c3_5modes"] = rec(c3, 1, steps=10)
        extras["c3_5modes"]["config"] = "C3: 1 angle x 1100 positions, 276x600 object, 128x128 detector, 5 modes"
        sweep = {}
        for nd in (64, 128, 256, 512):
            sweep[str(nd)] = rec(workloads.c5_sweep(nd, ntheta=2), 2, model=1, counts=100.0)
        extras["c5_poisson_sweep"] = {"config": "C5: 2 angles x 1024 positions, object (4 N)^2, N^2 detector, "
                                                "Poisson counts (mean 100), sub-pixel positions", "by_detector": sweep}
        extra["workloads"] = extras
    if world > 1:
        barrier(world)
    if rank != 0:
        return None
    pipe_bytes = npat * N * N * 8 * (8 + 4 + 2 + 3.1)   # bytes through the L1/shared pipe
    pipe_peak = 148 * 128 * 1.965e9
    traffic, traffic_src = None, None
    tpath = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tpath):
        try:
            tj = json.load(open(tpath))
            per_pattern = tj.get(args.workload + "_bytes_per_pattern")
            if per_pattern:
                traffic, traffic_src = per_pattern * npat, tj.get("source")
        except Exception:
            pass
    out = {
        "metric": "diffraction patterns/s (fwd+adj)", "value": value, "unit": "patterns/s",
        "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": secs / args.steps * 1e3, "higher_is_better": True,
        "scaling": "strong" if args.workload == "c4" else "weak",
        "vs_baseline": None, "dtype": "f32 (complex64)", "data": "synthetic (seeded, workloads.py)",
        "config": make_config(args.workload, world),
        "clocks": clocks,
        "e2e": {"value": e2e["pinned"], "unit": "patterns/s", "h2d_bytes_per_step": h2d,
                "d2h_bytes_per_step": d2h, "pageable": e2e["pageable"],
                "api": "CGPtychoSolver.grad_ptycho_batch (host arrays in, host gradient out, ptheta=1 chunks)",
                "sample": "%d angles per GPU per step on average (value: pinned host arrays; pageable: plain NumPy)" % E,
                "angles_per_rank": counts,
                "link_gbs": ([round(r, 1) for r in rates] if rates else None),
                "sharding": ("equal" if rates is None else
                             "in proportion to each rank's host-to-device rate with all GPUs copying at once")},
        "e2e_cg": e2e_cg,
        "gpu_launches": int(launches),
        "roofline": {"bound": "fp32", "achieved": aflops / kavg / 1e12, "peak": FP32_NOMINAL_TFLOPS,
                     "unit": "TFLOP/s", "frac": aflops / kavg / 1e12 / FP32_NOMINAL_TFLOPS,
                     "traffic": traffic, "traffic_source": traffic_src,
                     "peak_source": "nominal 148 SM x 128 lanes x 2 x 1.965 GHz (SURVEY.md 8d; "
                                    "MEASURED_PEAKS.json carries no FP32 figure)",
                     "kernel": "k_grad<Plan<%d>, gaussian, object>" % int(np.log2(N)),
                     "launch_ms": kavg * 1e3, "algorithmic_flops_per_launch": aflops,
                     "algorithmic_bytes_per_launch": abytes,
                     "hbm": {"achieved_gbs": abytes / kavg / 1e9, "peak_gbs": hbm,
                             "frac": abytes / kavg / 1e9 / hbm, "peak_source": peak_src},
                     "sm_data_pipe": {"achieved_tbs": pipe_bytes / kavg / 1e12, "peak_tbs": pipe_peak / 1e12,
                                      "frac": pipe_bytes / kavg / pipe_peak,
                                      "note": "L1/shared data pipe, 128 B/clk/SM x 148 SMs x 1.965 GHz"}},
    }
    out.update(extra)
    if world == 1:
        v, sample = cpu_port_rate(w)
        out["cpu_baseline"] = {"value": v, "unit": "patterns/s", "cores": os.cpu_count(),
                               "kind": "port", "sample": sample}
    if cg:
        out["cg"] = cg
    return out


# ---------------------------------------------------------------------------------------------
# the reference arm
# ---------------------------------------------------------------------------------------------
def run_reference(args, world, rank, local):
    """The reference's compiled cuFFT path (oracle/_ref) for the same pass; rank 0 only."""
    if rank != 0:
        return None
    from oracle import ref_gpu
    T = REF_ANGLES if args.workload == "c4" else 8
    w = make_angles(args.workload, 0, T)
    S, N, nz, n = w["nscan"], w["ndet"], w["nz"], w["n"]
    base = {"impl": "reference", "metric": "diffraction patterns/s (fwd+adj)", "unit": "patterns/s",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "higher_is_better": True,
            "scaling": "strong" if args.workload == "c4" else "weak", "vs_baseline": None,
            "dtype": "f32 (complex64)", "data": "synthetic (seeded, workloads.py)",
            "config": make_config(args.workload, world)}
    if not (ref_gpu.available() and torch.cuda.is_available()):
        v, sample = cpu_port_rate(w, 20.0)
        base.update({"value": v, "ms_per_step": None,
                     "notes": "oracle/_ref absent: NumPy port on host cores",
                     "cpu_baseline": {"value": v, "unit": "patterns/s", "cores": os.cpu_count(),
                                      "kind": "port", "sample": sample},
                     "e2e": {"value": v, "unit": "patterns/s", "h2d_bytes_per_step": 0,
                             "d2h_bytes_per_step": 0}})
        return base
    torch.cuda.set_device(0)
    prb_h = np.ascontiguousarray(w["probe"][:, 0])
    with ref_gpu.RefPtychoFFT(S, w["nprb"], N, T, nz, n) as ref:   # one call = all T angles
        psi_t, scan_d, prb_d = (torch.from_numpy(x).cuda() for x in (w["psi"], w["scan"], prb_h))
        data_d = (torch.abs(ref.fwd(psi_t, scan_d, prb_d)) ** 2).contiguous()
        psi_d = torch.ones_like(psi_t)

        def grad_dev(psi, scan, prb, data):
            f = ref.fwd(psi, scan, prb)                                   # ptycho.py:351 (b/a = 1)
            r = f - torch.sqrt(data) * f / (torch.sqrt(torch.abs(f) ** 2) + 1e-32)   # ptycho.py:353-354
            return ref.adj(r, scan, prb)                                  # ptycho.py:352-356

        secs = timed_steps(lambda: grad_dev(psi_d, scan_d, prb_d, data_d), args.steps, args.warmup, 1)
        value = T * S * args.steps / secs
        hdata = data_d.cpu().numpy()
        del data_d
    psi_h = np.ones_like(w["psi"])
    e2e = {}
    nrep = max(2, args.steps // 4)
    with ref_gpu.RefPtychoFFT(S, w["nprb"], N, 1, nz, n) as ref1:
        # the reference's host helpers process one angle per call (ptycho.py:70-78, 143-158):
        # H2D (cp.array), operators, blocking D2H (.get()) per angle
        def grad1(psi, scan, prb, data):
            f = ref1.fwd(psi, scan, prb)
            r = f - torch.sqrt(data) * f / (torch.sqrt(torch.abs(f) ** 2) + 1e-32)
            return ref1.adj(r, scan, prb)
        for kind in ("pinned", "pageable"):
            src = [(pinned(x) if kind == "pinned" else x) for x in (psi_h, w["scan"], prb_h, hdata)]

            def e2e_step():
                out = np.empty_like(psi_h)
                for t in range(T):
                    a = [torch.from_numpy(np.ascontiguousarray(x[t:t + 1])).cuda() for x in src]
                    out[t] = grad1(*a).cpu().numpy()[0]
                return out
            e2e_step()
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            for _ in range(nrep):
                e2e_step()
            torch.cuda.synchronize()
            e2e[kind] = T * S * nrep / (time.perf_counter() - t0)
            del src
    C = min(T, CG_ANGLES)
    with ref_gpu.RefCGPtychoSolver(S, w["nprb"], N, 1, nz, n) as r1, quiet():
        r1.position_correction = True   # unconditional in the reference (ptycho.py:398-403)
        r1.run_batch(hdata[:1], psi_h[:1], w["scan"][:1], w["probe"][:1], piter=2, recover_prb=True,
                     verbose=False)
        dt = None
        for _ in range(2):  # best of two, like the other arm
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            r1.run_batch(hdata[:C], psi_h[:C], w["scan"][:C], w["probe"][:C], piter=CG_ITERS,
                         recover_prb=True, verbose=False)
            torch.cuda.synchronize()
            t1 = time.perf_counter() - t0
            dt = t1 if dt is None else min(dt, t1)
        e2e_cg = C * CG_ITERS / dt
    cg = None
    if not args.no_cg:  # the reference's solver (cp -> torch restatement over its own operators)
        with ref_gpu.RefCGPtychoSolver(S, w["nprb"], N, 1, nz, n) as r1, quiet():
            a = [torch.from_numpy(x[:1]).cuda() for x in (hdata, psi_h, w["scan"], w["probe"])]
            cg = {"iters": 6, "recover_prb": True,
                  "config": "one %s angle, device resident" % args.workload}
            for key, on in (("iters_per_s", True), ("iters_per_s_no_position_correction", False)):
                r1.position_correction = on
                r1.run(a[0], a[1], a[2].clone(), a[3].clone(), piter=2, recover_prb=True, verbose=False)
                torch.cuda.synchronize()
                t0 = time.perf_counter()
                r1.run(a[0], a[1], a[2].clone(), a[3].clone(), piter=6, recover_prb=True, verbose=False)
                torch.cuda.synchronize()
                cg[key] = 6 / (time.perf_counter() - t0)
    if cg:
        base["cg"] = cg
    sample = ("the reference has no CPU path: its CUDA/cuFFT operators compiled unmodified (oracle/_ref) + "
              "torch elementwise standing in for CuPy, on ONE B200; each step = %d angles (%d patterns) "
              "of the batch in one ptheta = %d call" % (T, T * S, T))
    base.update({"value": value, "ms_per_step": secs / args.steps * 1e3, "notes": sample,
                 "cpu_baseline": {"value": value, "unit": "patterns/s", "cores": 1, "kind": "reference",
                                  "sample": sample},
                 "e2e": {"value": e2e["pinned"], "unit": "patterns/s", "pageable": e2e["pageable"],
                         "h2d_bytes_per_step": int(hdata.nbytes + psi_h.nbytes + w["scan"].nbytes + prb_h.nbytes),
                         "d2h_bytes_per_step": int(psi_h.nbytes),
                         "api": "per angle: H2D of the inputs, fwd -> residual -> adj, blocking D2H "
                                "(the reference's _batch helper, ptycho.py:70-78)"},
                 "e2e_cg": {"value": e2e_cg, "unit": "angle-iterations/s",
                            "api": "CGPtychoSolver.run_batch(piter=%d, recover_prb=True) restated over the "
                                   "reference's operators, pageable host arrays" % CG_ITERS,
                            "sample": "%d angles, best of two calls" % C}})
    return base


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="c4", choices=["c2", "c4"])
    ap.add_argument("--no-cg", action="store_true")
    ap.add_argument("--no-extras", action="store_true", help="skip the c2 / C3 / C5 sub-records")
    args = ap.parse_args()
    # stdout carries exactly ONE JSON line: anything libraries print on fd 1 meanwhile (NCCL's version
    # banner, the solver's CSV header) goes to stderr instead
    sys.stdout.flush()
    json_fd = os.dup(1)
    os.dup2(2, 1)
    args.warmup = max(args.warmup, 3)
    if args.impl == "reference":
        rank = int(os.environ.get("RANK", "0"))
        world = int(os.environ.get("WORLD_SIZE", "1"))
        out = run_reference(args, world, rank, 0)
    else:
        if not torch.cuda.is_available():
            raise SystemExit("bench.py needs a B200: there is no CPU fallback for the product path")
        world, rank, local = dist_setup(args.gpus)
        if args.workload == "c4" and C4_ANGLES % world:
            raise SystemExit("c4 shards 168 angles evenly: use 1, 2, 4 or 8 GPUs")
        out = run_b200(args, world, rank, local)
        if world > 1:
            import torch.distributed as dist
            dist.destroy_process_group()
    sys.stdout.flush()
    if out is not None:
        os.write(json_fd, (json.dumps(out) + "\n").encode())
    os.close(json_fd)


if __name__ == "__main__":
    main()
