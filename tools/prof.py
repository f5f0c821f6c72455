"""Launch each fused pass a few times on a C2/C4-shaped workload -- the target of the ncu captures.

    ncu --set full --clock-control none --import-source on -k regex:k_grad -s 2 -c 1 \
        -o gpurun_out/grad128 python tools/prof.py 128 4
"""
import ctypes
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "libtike-cufft_b200"))
import workloads  # noqa: E402
import libtike.cufft as pt  # noqa: E402


def main(ndet=128, T=4, reps=3, nmodes=1):
    nside = 32
    w = workloads.synth_angles(T, 4 * ndet, 4 * ndet, ndet, ndet, nside, nmodes)
    S = nside * nside
    psi, scan, probe = (torch.from_numpy(w[k]).cuda() for k in ("psi", "scan", "probe"))
    with pt.CGPtychoSolver(S, ndet, ndet, T, 4 * ndet, 4 * ndet) as slv:
        prb0 = probe[:, 0].contiguous()
        g = slv.fwd(psi, scan, prb0)
        data = (g.abs() ** 2).contiguous()
        psi1 = torch.ones_like(psi)
        dpsi = torch.randn_like(psi) * 0.01
        gradpsi = torch.zeros_like(psi)
        gradprb = torch.zeros_like(probe)
        for _ in range(reps):
            slv.fwd(psi, scan, prb0)
            slv.adj(g, scan, prb0)
            slv.adj_probe(g, scan, psi)
            slv._intensity(psi1, scan, probe, data, None, 0)
            slv._grad(0, psi1, scan, probe, 0, data, None, 1.0, 1.0, 1.0, 0, gradpsi)
            slv._grad(1, psi1, scan, probe, 0, data, None, 1.0, 1.0, 1.0, 0, gradprb,
                      probe.shape[1] * ndet * ndet)
            slv.ls_log = []
            slv._line_search(psi1, probe, nmodes, 0, dpsi, probe, nmodes, 0, nmodes, scan, data,
                             None, 0)
            # cached / a-b variants of the CG loop and the position-correction kernel
            far = torch.empty((nmodes,) + tuple(data.shape), dtype=torch.complex64, device="cuda")
            slv._grad(0, psi1, scan, probe, 0, data, None, 1.0, 1.0, 1.0, 0, gradpsi, far_out=far[0])
            slv._line_search(psi1, probe, nmodes, 0, dpsi, probe, nmodes, 0, nmodes, scan, data,
                             None, 0, far_a=far, want_ab=True)
            shifts = torch.empty((S, 2), dtype=torch.float64, device="cuda")
            psi2 = (psi1 + 0.5 * dpsi).contiguous()
            pt.ptycho.check(pt.ptycho.lib.ptx_cg_position_shifts(
                slv._h, ctypes.c_void_p(psi1.data_ptr()), ctypes.c_void_p(psi2.data_ptr()),
                ctypes.c_void_p(scan.data_ptr()), 100, ctypes.c_void_p(shifts.data_ptr()),
                pt.ptycho.current_stream()))
        torch.cuda.synchronize()
    print("prof done")


if __name__ == "__main__":
    nd = int(sys.argv[1]) if len(sys.argv) > 1 else 128
    T = int(sys.argv[2]) if len(sys.argv) > 2 else 4
    main(nd, T)
