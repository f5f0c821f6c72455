"""Time the fused position-correction kernel (ptx_cg_position_shifts) on one angle of c2 / c4."""
import ctypes
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "libtike-cufft_b200")]
import libtike.cufft as pt  # noqa: E402
import workloads  # noqa: E402

for ndet in [int(x) for x in (sys.argv[1:] or ["128", "256"])]:
    S = 1024
    nz = n = 4 * ndet
    w = workloads.synth_angles(1, nz, n, ndet, ndet, 32, 1, seed0=0)
    psi = torch.from_numpy(w["psi"]).cuda()
    psi_b = (psi * (1.0 + 0.01j) + 0.01 * torch.roll(psi, (1, 2), (1, 2))).contiguous()
    scan = torch.from_numpy(w["scan"]).cuda()
    out = torch.empty((S, 2), dtype=torch.float64, device="cuda")
    vp = ctypes.c_void_p
    with pt.CGPtychoSolver(S, ndet, ndet, 1, nz, n) as slv:
        def run():
            pt.ptycho.check(pt.ptycho.lib.ptx_cg_position_shifts(
                slv._h, vp(psi.data_ptr()), vp(psi_b.data_ptr()), vp(scan.data_ptr()), 100,
                vp(out.data_ptr()), pt.ptycho.current_stream()))
        for _ in range(2):
            run()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5):
            run()
        e1.record()
        e1.synchronize()
        ms = e0.elapsed_time(e1) / 5
    U = 150
    macs = S * (160 * ndet * ndet + 160 * 160 * ndet)
    print("ndet %d: %.3f ms per %d patterns; %.2f TFLOP/s fp64-equivalent of the direct products (8 flop per complex MAC, U=160); "
          "largest |shift| %.2f" % (ndet, ms, S, macs * 8 / ms / 1e9, float(out.abs().max())))
