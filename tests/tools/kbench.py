"""Kernel-level timing of every pass (CUDA events, device-resident inputs) next to the reference's
cuFFT path on the same GPU (oracle/_ref -- which is why this lives under tests/: only tests/, smoke()
and bench.py may execute anything from oracle/).  Development tool; the judged numbers come from
bench.py.   usage: python tests/tools/kbench.py <ndet> <angles> [<modes> [poisson]]"""
import ctypes
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "libtike-cufft_b200"))
import workloads  # noqa: E402
import libtike.cufft as pt  # noqa: E402
from libtike.cufft.ptychofft import lib, check, current_stream  # noqa: E402
from oracle import ref_gpu  # noqa: E402


def timeit(fn, reps=10, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ts = []
    for _ in range(reps):
        e0.record()
        fn()
        e1.record()
        e1.synchronize()
        ts.append(e0.elapsed_time(e1))
    return min(ts), float(np.median(ts))


def main(ndet=128, T=4, nside=32, nmodes=1, model=0):
    w = workloads.synth_angles(T, 4 * ndet, 4 * ndet, ndet, ndet, nside, nmodes)
    S = nside * nside
    psi, scan, probe = (torch.from_numpy(w[k]).cuda() for k in ("psi", "scan", "probe"))
    npat = T * S
    print(f"ndet={ndet} T={T} S={S} M={nmodes} patterns={npat} model={'poisson' if model else 'gaussian'}")
    with pt.CGPtychoSolver(S, ndet, ndet, T, 4 * ndet, 4 * ndet) as slv:
        prb0 = probe[:, 0].contiguous()
        g = slv.fwd(psi, scan, prb0)
        data = (g.abs() ** 2).contiguous()
        for k in range(1, nmodes):
            data += slv.fwd(psi, scan, probe[:, k].contiguous()).abs() ** 2
        if model:  # C5: Poisson counts, mean ~100 per pixel
            data = torch.poisson(data * (100.0 / data.mean())).contiguous()
        psi1 = torch.ones_like(psi)
        dpsi = torch.randn_like(psi) * 0.01
        gradpsi = torch.zeros_like(psi)
        gradprb = torch.zeros_like(probe)
        inten = torch.empty_like(data)

        def rate(ms):
            return npat / ms * 1e-3

        rows = []
        rows.append(("fwd (API, far field to HBM)", timeit(lambda: slv.fwd(psi, scan, prb0))))
        rows.append(("adj object (API)", timeit(lambda: slv.adj(g, scan, prb0))))
        rows.append(("adj probe (API)", timeit(lambda: slv.adj_probe(g, scan, psi))))
        rows.append(("cg_intensity", timeit(lambda: slv._intensity(psi1, scan, probe, data, None, model))))
        rows.append(("cg_grad object (fused fwd+adj)", timeit(
            lambda: slv._grad(0, psi1, scan, probe, 0, data, None, 1.0, 1.0, 1.0, model, gradpsi))))
        rows.append(("cg_grad probe (fused fwd+adj_probe)", timeit(
            lambda: slv._grad(1, psi1, scan, probe, 0, data, None, 1.0, 1.0, 1.0, model, gradprb, probe.shape[1] * ndet * ndet))))
        cost = torch.zeros(16, dtype=torch.float64, device="cuda")

        def ls():
            check(lib.ptx_cg_linesearch(slv._h, ctypes.c_void_p(psi1.data_ptr()), ctypes.c_void_p(probe.data_ptr()),
                                        nmodes, 0, ctypes.c_void_p(dpsi.data_ptr()), ctypes.c_void_p(probe.data_ptr()),
                                        nmodes, 0, nmodes, ctypes.c_void_p(scan.data_ptr()),
                                        ctypes.c_void_p(data.data_ptr()), None, None, model, 0, 4, 0, None,
                                        ctypes.c_void_p(cost.data_ptr()), current_stream()))
        rows.append(("cg_linesearch (2 FFT/mode, 4 candidates)", timeit(ls)))
        far = torch.empty((nmodes,) + tuple(data.shape), dtype=torch.complex64, device="cuda")
        for k in range(nmodes):
            slv._grad(0, psi1, scan, probe, k, data, None, 1.0, 1.0, 1.0, model, gradpsi, far_out=far[k])

        def ls_cached():
            check(lib.ptx_cg_linesearch(slv._h, ctypes.c_void_p(psi1.data_ptr()), ctypes.c_void_p(probe.data_ptr()),
                                        nmodes, 0, ctypes.c_void_p(dpsi.data_ptr()), ctypes.c_void_p(probe.data_ptr()),
                                        nmodes, 0, nmodes, ctypes.c_void_p(scan.data_ptr()),
                                        ctypes.c_void_p(data.data_ptr()), None, ctypes.c_void_p(far.data_ptr()),
                                        model, 0, 4, 0, None, ctypes.c_void_p(cost.data_ptr()), current_stream()))
        rows.append(("cg_linesearch, first far field cached", timeit(ls_cached)))
        rows.append(("cg_grad object + far-field cache write", timeit(
            lambda: slv._grad(0, psi1, scan, probe, 0, data, None, 1.0, 1.0, 1.0, model, gradpsi, far_out=far[0]))))
        for name, (best, med) in rows:
            print(f"  {name:44s} best {best:8.3f} ms  median {med:8.3f} ms  {rate(best):8.2f} M patterns/s")

        # CG iteration rate (recover_prb) on angle 0 only
        with pt.CGPtychoSolver(S, ndet, ndet, 1, 4 * ndet, 4 * ndet) as s1:
            d1, sc1, p1 = data[:1].contiguous(), scan[:1].contiguous(), probe[:1].contiguous().clone()
            mname = "poisson" if model else "gaussian"
            s1.run(d1, psi1[:1], sc1.clone(), p1.clone(), piter=2, recover_prb=True, model=mname)
            torch.cuda.synchronize()
            t0 = time.time()
            s1.run(d1, psi1[:1], sc1.clone(), p1.clone(), piter=16, recover_prb=True, model=mname)
            torch.cuda.synchronize()
            dt = time.time() - t0
            print(f"  CG (mine) 16 iterations, 1 angle: {dt*1e3:.1f} ms -> {16/dt:.1f} it/s")

    if ref_gpu.available():
        with ref_gpu.RefCGPtychoSolver(S, ndet, ndet, T, 4 * ndet, 4 * ndet) as ref:
            prb0 = probe[:, 0].contiguous()
            rows = []
            rows.append(("REF fwd (memset+mulop+cuFFT)", timeit(lambda: ref.fwd(psi, scan, prb0))))
            rows.append(("REF adj object (cuFFT+mulop atomics)", timeit(lambda: ref.adj(g, scan, prb0))))
            rows.append(("REF adj probe", timeit(lambda: ref.adj_probe(g, scan, psi))))

            def ref_grad():
                f = ref.fwd(psi1, scan, prb0)
                if model:
                    r = f - data * f / (f.abs() ** 2 + 1e-32)
                else:
                    r = f - torch.sqrt(data) * f / (torch.sqrt(f.abs() ** 2) + 1e-32)
                return ref.adj(r, scan, prb0)
            rows.append(("REF fwd+residual+adj (cuFFT path + torch elementwise)", timeit(ref_grad)))
            for name, (best, med) in rows:
                print(f"  {name:44s} best {best:8.3f} ms  median {med:8.3f} ms  {npat/best*1e-3:8.2f} M patterns/s")
        with ref_gpu.RefCGPtychoSolver(S, ndet, ndet, 1, 4 * ndet, 4 * ndet) as r1:
            d1, sc1, p1 = data[:1].contiguous(), scan[:1].contiguous(), probe[:1].contiguous().clone()
            r1.position_correction = True  # what the reference really executes; mine defaults to it too
            r1.run(d1, psi1[:1], sc1.clone(), p1.clone(), piter=2, recover_prb=True, verbose=False, model=mname)
            torch.cuda.synchronize()
            t0 = time.time()
            r1.run(d1, psi1[:1], sc1.clone(), p1.clone(), piter=16, recover_prb=True, verbose=False, model=mname)
            torch.cuda.synchronize()
            dt = time.time() - t0
            print(f"  CG (REF restatement) 16 iterations, 1 angle: {dt*1e3:.1f} ms -> {16/dt:.1f} it/s; trials {[len(t[1]) for t in r1.last_trials[:16]]}")


if __name__ == "__main__":
    nd = int(sys.argv[1]) if len(sys.argv) > 1 else 128
    T = int(sys.argv[2]) if len(sys.argv) > 2 else 4
    M = int(sys.argv[3]) if len(sys.argv) > 3 else 1
    model = 1 if (len(sys.argv) > 4 and sys.argv[4] == "poisson") else 0
    main(nd, T, 32, M, model)
