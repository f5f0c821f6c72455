#!/usr/bin/env python
"""Headline benchmark: diffraction patterns/s through the fused fwd -> residual -> adj pass.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference] [--workload c2|c4]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
        --master-port P bench.py --gpus N --steps K --warmup W

A "step" is one pass of the hot path (a2 fwd + the Gaussian residual of
ptycho.py:351-356 + a3 adj, i.e. the CG object gradient) over one batch of angles:

  workload c2 (default; BASELINE.json configs[1]): per GPU 8 independent angles of a 512x512 object,
      1024 scan positions each, 128x128 detector, 1 probe mode, Gaussian model -> 8192 patterns/step,
      512 MiB of measured data per step (larger than the 126 MB L2, so no flush is needed);
  workload c4 (configs[3]): 1024x1024 object slices, 256x256 detector, 1024 positions per angle.

Angles are independent problems (ptheta = 1 semantics of every reference test), so N GPUs run N
shards with no data-path collective: "scaling": "weak".

  value    -- patterns/s with every input already resident in HBM (CUDA events, max over ranks)
  e2e      -- the same pass through the host-array API (`CGPtychoSolver.grad_ptycho_batch`): pinned
              host buffers in, host gradient out, H2D/D2H inside the timed region
  roofline -- the fused kernel k_grad<.,gaussian,object>: algorithmic HBM bytes per launch / its
              CUDA-event duration against MEASURED_PEAKS.json, plus the FP32-pipe view of the same
              launches (this kernel is FP32-bound)
  cpu_baseline -- the NumPy/pocketfft oracle on the box's host cores, bounded sample (rank 0, N = 1)
  cg       -- CG iterations/s of `CGPtychoSolver.run` on one c2 angle (object + probe recovery)

--impl reference runs the REFERENCE's own CUDA/cuFFT operators (oracle/_ref, compiled unmodified
from /root/reference/src/cuda) for the same pass and config: fwd -> torch elementwise (the CuPy
statements of ptycho.py:351-356) -> adj.  The reference has no CPU implementation of this path;
if oracle/_ref did not travel, the NumPy port is timed instead and the line says so.
"""
import argparse
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
                "reasons": sorted(self.reasons), "samples": len(s)}


def physical_gpu_index(local):
    vis = os.environ.get("CUDA_VISIBLE_DEVICES")
    if vis:
        try:
            return int(vis.split(",")[local])
        except Exception:
            return local
    return local


def make_workload(name, angles):
    if name == "c2":
        w = workloads.c2_single_angle(ntheta=angles)
        label = "c2: %d angles/GPU x 1024 positions, 512x512 object, 128x128 detector, 1 mode, gaussian" % angles
    elif name == "c4":
        w = workloads.c4_catalyst(angles)
        label = "c4: %d angles/GPU x 1024 positions, 1024x1024 object, 256x256 detector, 1 mode, gaussian" % angles
    else:
        raise SystemExit("unknown workload " + name)
    return w, label


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


def timed_steps(step, steps, warmup, world, per_launch_events=False):
    """W warm-ups, then exactly K steps between barrier+synchronize, timed with CUDA events on the
    launching stream; returns (seconds max over ranks, list of per-step ms on this rank)."""
    for _ in range(warmup):
        step()
    barrier(world)
    evs = [torch.cuda.Event(enable_timing=True) for _ in range(steps + 1)]
    evs[0].record()
    for i in range(steps):
        step()
        evs[i + 1].record()
    barrier(world)
    total_ms = evs[0].elapsed_time(evs[-1])
    per = [evs[i].elapsed_time(evs[i + 1]) for i in range(steps)]
    return max_over_ranks(total_ms, world) * 1e-3, per


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


def run_b200(args, world, rank, local):
    import libtike.cufft as pt
    from libtike.cufft.ptychofft import launch_count
    T = args.angles
    w, label = make_workload(args.workload, T)
    S, N, nz, n = w["nscan"], w["ndet"], w["nz"], w["n"]
    npat = T * S
    dev = torch.device("cuda", torch.cuda.current_device())
    psi_true, scan, probe = (torch.from_numpy(w[k]).to(dev) for k in ("psi", "scan", "probe"))
    hbm, peak_src = peaks()
    with pt.CGPtychoSolver(S, w["nprb"], N, T, nz, n) as slv:
        data = slv.fwd(psi_true, scan, probe[:, 0]).abs().square_().contiguous()  # synthetic measurement
        psi = torch.ones_like(psi_true)  # the solver's starting point (tests/test.py:55)
        grad = torch.zeros_like(psi)

        def step():
            grad.zero_()
            slv._grad(0, psi, scan, probe, 0, data, None, 1.0, 1.0, 1.0, 0, grad)

        sampler = ClockSampler(physical_gpu_index(local))
        l0 = launch_count()
        for _ in range(args.warmup):
            step()
        barrier(world)
        sampler.start()
        l1 = launch_count()
        secs, _ = timed_steps(step, args.steps, 0, world)
        launches = launch_count() - l1
        clocks = sampler.stop()
        value = world * npat * args.steps / secs

        # dominant kernel alone: events straight around each launch (same stream)
        kt = []
        for _ in range(args.steps):
            grad.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            slv._grad(0, psi, scan, probe, 0, data, None, 1.0, 1.0, 1.0, 0, grad)
            e1.record()
            e1.synchronize()
            kt.append(e0.elapsed_time(e1) * 1e-3)
        kavg = float(np.mean(kt))
        abytes, aflops = algorithmic_bytes(w, T), algorithmic_flops(w, T)
        # the roof that binds: bytes this algorithm moves through the SM's
        # 128 B/clk L1 / shared-memory pipe -- 2 exchanges x 2 transforms x (store + load) of the
        # 8 N^2-byte tile, 4 bilinear taps + 2 probe reads + scatter staging and reductions
        pipe_bytes = npat * N * N * 8 * (8 + 4 + 2 + 3.1)
        pipe_peak = 148 * 128 * 1.965e9
        pipe = {"achieved_tbs": pipe_bytes / kavg / 1e12, "peak_tbs": pipe_peak / 1e12,
                "frac": pipe_bytes / kavg / pipe_peak,
                "note": "L1/shared data pipe, 128 B/clk/SM x 148 SMs x 1.965 GHz; ncu "
                        "l1tex__data_pipe_lsu_wavefronts reads 58 % busy (profiles/)"}

        # end to end through the host-array API (pinned host memory in, host gradient out)
        h = {k: torch.from_numpy(np.ascontiguousarray(v)).pin_memory().numpy()
             for k, v in (("data", data.cpu().numpy()), ("psi", psi.cpu().numpy()),
                          ("scan", w["scan"]), ("probe", w["probe"]))}
        with pt.CGPtychoSolver(S, w["nprb"], N, 1, nz, n) as s1:
            def e2e_step():
                s1.grad_ptycho_batch(h["data"], h["psi"], h["scan"], h["probe"], model="gaussian")
            for _ in range(max(1, min(args.warmup, 2))):
                e2e_step()
            barrier(world)
            t0 = time.perf_counter()
            for _ in range(args.steps):
                e2e_step()
            barrier(world)
            e2e_secs = max_over_ranks(time.perf_counter() - t0, world)
        h2d = sum(h[k].nbytes for k in h)
        d2h = h["psi"].nbytes

        cg = None
        if rank == 0 and not args.no_cg:
            with pt.CGPtychoSolver(S, w["nprb"], N, 1, nz, n) as s1:
                import contextlib
                import io
                d1 = data[:1].contiguous()
                cg = {"iters": 16, "recover_prb": True,
                      "config": "one %s angle, device resident" % args.workload}
                # position correction (ptycho.py:398-403) is unconditional in the reference: ON is
                # the drop-in configuration; OFF is reported next to it
                for key, on in (("iters_per_s", True), ("iters_per_s_no_position_correction", False)):
                    s1.position_correction = on
                    with contextlib.redirect_stdout(io.StringIO()):
                        s1.run(d1, psi[:1], scan[:1].clone(), probe[:1].clone(), piter=2, recover_prb=True)
                        dt = None
                        for _ in range(2):  # best of two: a stray allocation stall once halved a run
                            sc1, pr1 = scan[:1].clone(), probe[:1].clone()
                            torch.cuda.synchronize()
                            t0 = time.perf_counter()
                            s1.run(d1, psi[:1], sc1, pr1, piter=16, recover_prb=True)
                            torch.cuda.synchronize()
                            t1 = time.perf_counter() - t0
                            dt = t1 if dt is None else min(dt, t1)
                    cg[key] = 16 / dt
                    if on:  # it varies with the data: fused 4-candidate passes
                        cg["line_search_passes_per_iter"] = len(s1.ls_log) / 16.0
    if world > 1:
        barrier(world)
    if rank != 0:
        return None
    out = {
        "metric": "diffraction patterns/s (fwd+adj)", "value": value, "unit": "patterns/s",
        "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": secs / args.steps * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32 (complex64)", "data": "synthetic (seeded, workloads.py)",
        "config": {"workload": label, "patterns_per_step_per_gpu": npat,
                   "l2": "inputs (%.0f MB measured data per step) exceed the 126 MB L2; no flush" %
                         (data.numel() * 4 / 1e6),
                   "parallelism": "angle shards, no collective" if world > 1 else "single GPU"},
        "clocks": clocks,
        "e2e": {"value": world * npat * args.steps / e2e_secs, "unit": "patterns/s",
                "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                "api": "CGPtychoSolver.grad_ptycho_batch (pinned host arrays, ptheta=1 chunks)"},
        "gpu_launches": int(launches),
        "roofline": {"bound": "hbm", "achieved": abytes / kavg / 1e9, "peak": hbm, "unit": "GB/s",
                     "frac": abytes / kavg / 1e9 / hbm, "traffic": None, "peak_source": peak_src,
                     "kernel": "k_grad<gaussian, object>", "launch_ms": kavg * 1e3,
                     "algorithmic_bytes_per_launch": abytes,
                     "fp32": {"achieved_tflops": aflops / kavg / 1e12,
                              "peak_tflops": FP32_NOMINAL_TFLOPS,
                              "frac": aflops / kavg / 1e12 / FP32_NOMINAL_TFLOPS,
                              "note": "nominal peak 148 SM x 128 lanes x 2 x 1.965 GHz, FFT flops "
                                      "20 N^2 log2 N (SURVEY.md 8d)"},
                     "sm_data_pipe": pipe},
    }
    traffic = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(traffic):
        try:
            out["roofline"]["traffic"] = json.load(open(traffic)).get(args.workload)
        except Exception:
            pass
    if world == 1:
        v, sample = cpu_port_rate(w)
        out["cpu_baseline"] = {"value": v, "unit": "patterns/s", "cores": os.cpu_count(),
                               "kind": "port", "sample": sample}
    if cg:
        out["cg"] = cg
    return out


def run_reference(args, world, rank, local):
    """The reference's compiled cuFFT path (oracle/_ref) for the same pass; rank 0 only."""
    if rank != 0:
        return None
    from oracle import ref_gpu
    T = args.angles
    w, label = make_workload(args.workload, T)
    S, N, nz, n = w["nscan"], w["ndet"], w["nz"], w["n"]
    base = {"impl": "reference", "metric": "diffraction patterns/s (fwd+adj)", "unit": "patterns/s",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32 (complex64)",
            "data": "synthetic (seeded, workloads.py)"}
    if not (ref_gpu.available() and torch.cuda.is_available()):
        v, sample = cpu_port_rate(w, 20.0)
        base.update({"value": v, "ms_per_step": None,
                     "config": {"workload": label, "note": "oracle/_ref absent: NumPy port on host cores"},
                     "cpu_baseline": {"value": v, "unit": "patterns/s", "cores": os.cpu_count(),
                                      "kind": "port", "sample": sample},
                     "e2e": {"value": v, "unit": "patterns/s", "h2d_bytes_per_step": 0,
                             "d2h_bytes_per_step": 0}})
        return base
    torch.cuda.set_device(0)
    # the reference processes one angle per call through its host helpers (ptycho.py:70-78, 143-158)
    hdata = None
    with ref_gpu.RefPtychoFFT(S, w["nprb"], N, 1, nz, n) as ref:
        prb_h = np.ascontiguousarray(w["probe"][:, 0])
        hdata = np.stack([np.abs(ref.fwd_ptycho_batch(w["psi"][t:t + 1], w["scan"][t:t + 1],
                                                      prb_h[t:t + 1])[0]) ** 2 for t in range(T)])
        psi_h = np.ones_like(w["psi"])
        dev = [(torch.from_numpy(psi_h[t:t + 1]).cuda(), torch.from_numpy(w["scan"][t:t + 1]).cuda(),
                torch.from_numpy(prb_h[t:t + 1]).cuda(), torch.from_numpy(hdata[t:t + 1]).cuda())
               for t in range(T)]

        def grad_dev(psi, scan, prb, data):
            f = ref.fwd(psi, scan, prb)                                   # ptycho.py:351 (b/a = 1)
            r = f - torch.sqrt(data) * f / (torch.sqrt(torch.abs(f) ** 2) + 1e-32)   # ptycho.py:353-354
            return ref.adj(r, scan, prb)                                  # ptycho.py:352-356

        def step():
            for a in dev:
                grad_dev(*a)

        secs, _ = timed_steps(step, args.steps, args.warmup, 1)
        value = T * S * args.steps / secs

        def e2e_step():  # H2D from pageable numpy (cp.array) and a blocking D2H (.get()) per angle
            out = np.empty_like(psi_h)
            for t in range(T):
                a = [torch.from_numpy(np.ascontiguousarray(x[t:t + 1])).cuda()
                     for x in (psi_h, w["scan"], prb_h, hdata)]
                out[t] = grad_dev(*a).cpu().numpy()[0]
            return out
        e2e_step()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(args.steps):
            e2e_step()
        torch.cuda.synchronize()
        e2e_secs = time.perf_counter() - t0
    e2e = T * S * args.steps / e2e_secs
    cg = None
    if not args.no_cg:  # the reference's solver (cp -> torch restatement over its own operators)
        import contextlib
        import io
        with ref_gpu.RefCGPtychoSolver(S, w["nprb"], N, 1, nz, n) as r1:
            a = dev[0]
            cg = {"iters": 6, "recover_prb": True,
                  "config": "one %s angle, device resident" % args.workload}
            for key, on in (("iters_per_s", True), ("iters_per_s_no_position_correction", False)):
                r1.position_correction = on
                with contextlib.redirect_stdout(io.StringIO()):
                    r1.run(a[3], a[0], a[1].clone(), a[2][:, None].clone(), piter=2, recover_prb=True)
                    torch.cuda.synchronize()
                    t0 = time.perf_counter()
                    r1.run(a[3], a[0], a[1].clone(), a[2][:, None].clone(), piter=6, recover_prb=True)
                    torch.cuda.synchronize()
                    cg[key] = 6 / (time.perf_counter() - t0)
    if cg:
        base["cg"] = cg
    base.update({"value": value, "ms_per_step": secs / args.steps * 1e3,
                 "config": {"workload": label, "patterns_per_step_per_gpu": T * S,
                            "note": "reference CUDA/cuFFT operators compiled unmodified (oracle/_ref) + "
                                    "torch elementwise standing in for CuPy; GPU 0 only"},
                 "cpu_baseline": {"value": e2e, "unit": "patterns/s", "cores": 1, "kind": "reference",
                                  "sample": "the reference has no CPU path: its cuFFT path on one B200, "
                                            "%d steps of %d patterns, host buffers" % (args.steps, T * S)},
                 "e2e": {"value": e2e, "unit": "patterns/s",
                         "h2d_bytes_per_step": int(hdata.nbytes + psi_h.nbytes + w["scan"].nbytes + prb_h.nbytes),
                         "d2h_bytes_per_step": int(psi_h.nbytes)}})
    return base


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="c2", choices=["c2", "c4"])
    ap.add_argument("--angles", type=int, default=0,
                    help="angles per GPU per step (default: 8 for c2, 2 for c4 = 512 MiB of data)")
    ap.add_argument("--no-cg", action="store_true")
    args = ap.parse_args()
    # stdout carries exactly ONE JSON line: anything libraries print on fd 1 meanwhile (NCCL's version
    # banner, the solver's CSV header) goes to stderr instead
    sys.stdout.flush()
    json_fd = os.dup(1)
    os.dup2(2, 1)
    args.warmup = max(args.warmup, 3)
    if args.angles <= 0:
        args.angles = 8 if args.workload == "c2" else 2
    if args.impl == "reference":
        rank = int(os.environ.get("RANK", "0"))
        world = int(os.environ.get("WORLD_SIZE", "1"))
        out = run_reference(args, world, rank, 0)
    else:
        if not torch.cuda.is_available():
            raise SystemExit("bench.py needs a B200: there is no CPU fallback for the product path")
        world, rank, local = dist_setup(args.gpus)
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
