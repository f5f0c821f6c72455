"""Timing of the fused gradient / adjoint / registration kernels of the multi-tile plans with and
without the persisting-L2 window on the staging frames (PTX_L2_PERSIST=0 switches it off), plus the
device's persisting-L2 attributes.   usage: python tools/l2_probe.py [ndet=256] [angles=2]"""
import ctypes
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "libtike-cufft_b200")]
import workloads  # noqa: E402
import libtike.cufft as pt  # noqa: E402


def attrs():
    try:
        rt = ctypes.CDLL("libcudart.so.12")
    except OSError:
        import glob
        rt = ctypes.CDLL(glob.glob(os.path.join(os.path.dirname(torch.__file__), "..", "nvidia", "cuda_runtime",
                                                "lib", "libcudart.so*"))[0])
    out = {}
    for name, idx in (("l2_bytes", 38), ("max_persisting_l2", 108), ("max_access_policy_window", 109)):
        v = ctypes.c_int(0)
        rt.cudaDeviceGetAttribute(ctypes.byref(v), idx, 0)
        out[name] = v.value
    return out


def timeit(fn, reps=8, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        e1.synchronize()
        ts.append(e0.elapsed_time(e1))
    return min(ts), float(np.median(ts))


def main(ndet=256, T=2):
    torch.cuda.init()
    print("PTX_L2_PERSIST =", os.environ.get("PTX_L2_PERSIST", "(default on)"), attrs())
    w = workloads.synth_angles(T, 4 * ndet, 4 * ndet, ndet, ndet, 32, 1)
    S = 1024
    psi, scan, probe = (torch.from_numpy(w[k]).cuda() for k in ("psi", "scan", "probe"))
    with pt.CGPtychoSolver(S, ndet, ndet, T, 4 * ndet, 4 * ndet) as slv:
        prb0 = probe[:, 0].contiguous()
        g = slv.fwd(psi, scan, prb0)
        data = (g.abs() ** 2).contiguous()
        psi1 = torch.ones_like(psi)
        grad = torch.zeros_like(psi)
        gprb = torch.zeros_like(probe)
        npat = T * S
        for name, fn in (
                ("fwd", lambda: slv.fwd(psi, scan, prb0)),
                ("adj object", lambda: slv.adj(g, scan, prb0)),
                ("adj probe", lambda: slv.adj_probe(g, scan, psi)),
                ("cg_intensity", lambda: slv._intensity(psi1, scan, probe, data, None, 0)),
                ("cg_grad object", lambda: slv._grad(0, psi1, scan, probe, 0, data, None, 1.0, 1.0, 1.0, 0, grad)),
                ("cg_grad probe", lambda: slv._grad(1, psi1, scan, probe, 0, data, None, 1.0, 1.0, 1.0, 0, gprb,
                                                    ndet * ndet))):
            best, med = timeit(fn)
            print("  %-18s best %8.3f ms  median %8.3f ms  %7.3f M patterns/s" % (name, best, med, npat / best * 1e-3))


if __name__ == "__main__":
    main(*(int(x) for x in sys.argv[1:]))
