"""Timing of the fused position-correction kernel (ptx_cg_position_shifts) alone.
usage: python tools/reg_probe.py [ndet=128]   (PTX_REG_TMA=1 / PTX_TMA_GATHER=off select the gather)"""
import ctypes
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "libtike-cufft_b200")]
import workloads  # noqa: E402
import libtike.cufft as pt  # noqa: E402
from libtike.cufft.ptychofft import lib, check, current_stream  # noqa: E402

nd = int(sys.argv[1]) if len(sys.argv) > 1 else 128
w = workloads.synth_angles(1, 4 * nd, 4 * nd, nd, nd, 32, 1)
psi, scan = (torch.from_numpy(w[k]).cuda() for k in ("psi", "scan"))
psi_b = psi + 0.01 * torch.roll(psi, 3, dims=2)
shifts = torch.empty((1024, 2), dtype=torch.float64, device="cuda")
with pt.CGPtychoSolver(1024, nd, nd, 1, 4 * nd, 4 * nd) as slv:
    def fn():
        check(lib.ptx_cg_position_shifts(slv._h, ctypes.c_void_p(psi.data_ptr()), ctypes.c_void_p(psi_b.data_ptr()),
                                         ctypes.c_void_p(scan.data_ptr()), 100, ctypes.c_void_p(shifts.data_ptr()),
                                         current_stream()))
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(8):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); e1.synchronize()
        ts.append(e0.elapsed_time(e1))
    print("ndet %d PTX_REG_TMA=%s PTX_TMA_GATHER=%s: %.3f ms per 1024 patterns, checksum %.6f" % (
        nd, os.environ.get("PTX_REG_TMA", "-"), os.environ.get("PTX_TMA_GATHER", "-"), min(ts), float(shifts.sum())))
