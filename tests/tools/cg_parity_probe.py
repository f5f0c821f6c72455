"""CG-parity probe (development tool; lives under tests/ because it executes the oracle).

For the cases where two fp32 implementations drift apart by more than 1e-4 it prints, per case:
  fused vs reference, reference vs a SECOND RUN OF THE REFERENCE (atomic order is non-deterministic),
  and the distances of fused / reference to the float64 restatement (same replayed decisions).
Run once per library build to A/B the per-pixel math:
    python tests/tools/cg_parity_probe.py                     # SFU rsqrt/rcp (shipped)
    PTYCHOFFT_B200_LIB=.../libptychofft_b200_ieee.so python tests/tools/cg_parity_probe.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path[:0] = [ROOT, os.path.join(ROOT, "libtike-cufft_b200"), os.path.join(ROOT, "tests")]
import test_gpu_cg as T  # noqa: E402
from oracle import numpy_ptycho as O  # noqa: E402
from oracle import ref_gpu  # noqa: E402
from util import rel_l2, ReplaySolver  # noqa: E402

CASES = [(1, 36, "gaussian", 24, 128, False), (1, 36, "gaussian", 32, 128, False),
         (1, 49, "poisson", 6, 128, False), (1, 49, "poisson", 6, 128, True)]


def main():
    print("library:", os.environ.get("PTYCHOFFT_B200_LIB", "default"))
    for nmodes, nscan, model, piter, ndet, noisy in CASES:
        data, psi0, scan, prb0 = T._problem(nmodes, nscan, model, ndet, noisy=noisy)
        nz, n = psi0.shape[1:]
        runs = []
        with ref_gpu.RefCGPtychoSolver(nscan, ndet, ndet, 1, nz, n) as ref:
            for rep in range(2):
                want = ref.run_batch(data, psi0, scan, prb0, piter=piter, model=model, recover_prb=True,
                                     verbose=False)
                runs.append((want, [t[2] for t in ref.last_trials]))
        want, steps = runs[0]
        with O.float64_arithmetic():
            exact = O.cg_run(data, psi0, scan, prb0.copy(), piter, model, True, forced_steps=list(steps))
        with ReplaySolver(nscan, ndet, ndet, 1, nz, n) as slv:
            slv.position_correction = False
            slv.forced_steps = list(steps)
            got = slv.run_batch(data, psi0, scan, prb0, piter=piter, model=model, recover_prb=True)
        print("case", (nmodes, nscan, model, piter, ndet, "noisy" if noisy else "clean"),
              "same decisions in both reference runs:", runs[0][1] == runs[1][1])
        for k in ("psi", "probe"):
            print("   %-5s fused-ref %.2e | ref-ref2 %.2e | fused-f64 %.2e | ref-f64 %.2e | ref2-f64 %.2e"
                  % (k, rel_l2(got[k], want[k]), rel_l2(runs[1][0][k], want[k]), rel_l2(got[k], exact[k]),
                     rel_l2(want[k], exact[k]), rel_l2(runs[1][0][k], exact[k])))


if __name__ == "__main__":
    main()
