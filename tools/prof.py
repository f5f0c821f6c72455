"""Launch each fused pass a few times on a C2/C4-shaped workload -- the target of the ncu captures.

    ncu --set full --clock-control none --import-source on -k regex:k_grad -s 2 -c 1 \
        -o gpurun_out/grad128 python tools/prof.py 128 4
"""
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
        torch.cuda.synchronize()
    print("prof done")


if __name__ == "__main__":
    nd = int(sys.argv[1]) if len(sys.argv) > 1 else 128
    T = int(sys.argv[2]) if len(sys.argv) > 2 else 4
    main(nd, T)
