"""Distances of the fused solver and of the reference's cuFFT path to the float64 trajectory on the
C5 Poisson problem at 128^2 (1024 positions, 3 iterations, position correction on), for several noise
realisations -- run once per kernel variant (PTX_PIPE=0 / 1).  Development tool (executes the oracle)."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path[:0] = [ROOT, os.path.join(ROOT, "libtike-cufft_b200"), os.path.join(ROOT, "tests")]
import workloads  # noqa: E402
from oracle import ref_gpu  # noqa: E402
from util import rel_l2, ReplaySolver  # noqa: E402


def main(ndet=128, seeds=(5, 6, 7, 8)):
    print("PTX_PIPE =", os.environ.get("PTX_PIPE", "1"))
    w = workloads.c5_sweep(ndet, nside=32)
    psi, scan, probe = w["psi"], w["scan"], w["probe"]
    S, (nz, n) = scan.shape[1], psi.shape[1:]
    cu = lambda x: torch.from_numpy(np.ascontiguousarray(x)).cuda()  # noqa: E731
    with ref_gpu.RefPtychoFFT(S, ndet, ndet, 1, nz, n) as ref:
        clean = torch.abs(ref.fwd(cu(psi), cu(scan), cu(probe[:, 0]))) ** 2
    psi0 = np.ones_like(psi)
    prb0 = (probe * (0.9 + 0.1j)).astype(np.complex64)
    for seed in seeds:
        d = torch.poisson(clean * (100.0 / clean.mean()), generator=torch.Generator("cuda").manual_seed(seed))
        data = d.cpu().numpy()
        with ref_gpu.RefCGPtychoSolver(S, ndet, ndet, 1, nz, n) as r:
            r.position_correction = True
            want = r.run_batch(data, psi0, scan, prb0, piter=3, model="poisson", recover_prb=True, verbose=False)
            steps = [t[2] for t in r.last_trials]
        with ref_gpu.F64CGPtychoSolver(S, ndet, ndet, 1, nz, n) as ex:
            ex.position_correction = True
            ex.forced_steps = list(steps)
            exact = ex.run_batch(data, psi0, scan, prb0, piter=3, model="poisson", recover_prb=True, verbose=False)
        with ReplaySolver(S, ndet, ndet, 1, nz, n) as slv:
            slv.forced_steps = list(steps)
            got = slv.run_batch(data, psi0, scan, prb0, piter=3, model="poisson", recover_prb=True)
        print("seed %d steps %s" % (seed, steps))
        for k in ("psi", "probe"):
            print("   %-5s fused-ref %.2e | ref-f64 %.2e | fused-f64 %.2e" % (
                k, rel_l2(got[k], want[k]), rel_l2(want[k], exact[k]), rel_l2(got[k], exact[k])))


if __name__ == "__main__":
    main()
