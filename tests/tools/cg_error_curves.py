"""Error-versus-iteration curves of the long Gaussian CG case (1 mode, 36 positions, 128^2): distance of
the fused solver and of the reference's cuFFT path to the float64 GPU referee after 1..32 iterations
(each point is a fresh run; decisions of that reference run replayed in the other two).
Development tool (executes the oracle); output committed under profiles/."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path[:0] = [ROOT, os.path.join(ROOT, "libtike-cufft_b200"), os.path.join(ROOT, "tests")]
import test_gpu_cg as T  # noqa: E402
from oracle import ref_gpu  # noqa: E402
from util import rel_l2, ReplaySolver  # noqa: E402


def main():
    data, psi0, scan, prb0 = T._problem(1, 36, "gaussian", 128)
    nz, n = psi0.shape[1:]
    print("# piter | psi: fused-ref  ref-f64  fused-f64 | probe: fused-ref  ref-f64  fused-f64")
    for piter in (1, 2, 4, 8, 12, 16, 20, 24, 28, 32):
        with ref_gpu.RefCGPtychoSolver(36, 128, 128, 1, nz, n) as ref:
            want = ref.run_batch(data, psi0, scan, prb0, piter=piter, model="gaussian", recover_prb=True,
                                 verbose=False)
            steps = [t[2] for t in ref.last_trials]
        with ref_gpu.F64CGPtychoSolver(36, 128, 128, 1, nz, n) as ex:
            ex.forced_steps = list(steps)
            exact = ex.run_batch(data, psi0, scan, prb0, piter=piter, model="gaussian", recover_prb=True,
                                 verbose=False)
        with ReplaySolver(36, 128, 128, 1, nz, n) as slv:
            slv.position_correction = False
            slv.forced_steps = list(steps)
            import contextlib
            import io
            with contextlib.redirect_stdout(io.StringIO()):
                got = slv.run_batch(data, psi0, scan, prb0, piter=piter, model="gaussian", recover_prb=True)
        row = []
        for k in ("psi", "probe"):
            row += [rel_l2(got[k], want[k]), rel_l2(want[k], exact[k]), rel_l2(got[k], exact[k])]
        print("%5d   | %.2e  %.2e  %.2e |  %.2e  %.2e  %.2e" % ((piter,) + tuple(row)))


if __name__ == "__main__":
    main()
