"""Timing of the 128^2 object-gradient kernel only (pipelined vs single-role, ablations via
PTX_PIPE / PTX_PIPE_DEBUG).   usage: python tools/pipe_probe.py [angles=4]"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "libtike-cufft_b200")]
import workloads  # noqa: E402
import libtike.cufft as pt  # noqa: E402


def main(T=4):
    w = workloads.synth_angles(T, 512, 512, 128, 128, 32, 1)
    S = 1024
    psi, scan, probe = (torch.from_numpy(w[k]).cuda() for k in ("psi", "scan", "probe"))
    with pt.CGPtychoSolver(S, 128, 128, T, 512, 512) as slv:
        data = (slv.fwd(psi, scan, probe[:, 0].contiguous()).abs() ** 2).contiguous()
        psi1 = torch.ones_like(psi)
        grad = torch.zeros_like(psi)
        sc = torch.ones(3, dtype=torch.float32, device="cuda")
        fn = lambda: slv._grad(0, psi1, scan, probe, 0, data, None, 1.0, 1.0, 1.0, 0, grad, sc=sc)  # noqa: E731
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        ts = []
        for _ in range(8):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            fn()
            e1.record()
            e1.synchronize()
            ts.append(e0.elapsed_time(e1))
        best = min(ts)
        print("PTX_PIPE=%s PTX_PIPE_DEBUG=%s  best %.3f ms  %.2f M patterns/s  %.0f clk/pattern/SM" % (
            os.environ.get("PTX_PIPE", "1"), os.environ.get("PTX_PIPE_DEBUG", "0"), best, T * S / best * 1e-3,
            best * 1e-3 * 1.965e9 / (T * S / 148.0)))


if __name__ == "__main__":
    main(*(int(x) for x in sys.argv[1:]))
