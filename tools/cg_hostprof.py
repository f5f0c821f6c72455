"""Host-side profile of CGPtychoSolver.run: where the Python time of an iteration goes (cProfile) and
how far ahead of the GPU the host runs.  Usage: python tools/cg_hostprof.py NDET [ITERS]"""
import cProfile
import io
import os
import pstats
import sys
import time

import numpy as np
import torch

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
for d in (ROOT, os.path.join(ROOT, "tests"), os.path.join(ROOT, "libtike-cufft_b200")):
    sys.path.insert(0, d)
import workloads  # noqa: E402
import libtike.cufft as pt  # noqa: E402

ndet = int(sys.argv[1]) if len(sys.argv) > 1 else 64
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 32
nz = n = 4 * ndet
w = workloads.synth_angles(1, nz, n, ndet, ndet, 32, 1, seed0=0)
psi, scan, probe = w["psi"], w["scan"], w["probe"]
from oracle import numpy_ptycho as O  # noqa: E402  (test-side data synthesis only)
with pt.CGPtychoSolver(scan.shape[1], ndet, ndet, 1, nz, n) as slv:
    data = torch.abs(slv.fwd(torch.as_tensor(psi).cuda(), torch.as_tensor(scan).cuda(),
                             torch.as_tensor(probe[:, 0]).cuda())) ** 2
    args = lambda: (data, torch.ones_like(torch.as_tensor(psi)).cuda(), torch.as_tensor(scan.copy()).cuda(),
                    (torch.as_tensor(probe) * (0.9 + 0.1j)).cuda())
    import contextlib
    with contextlib.redirect_stdout(io.StringIO()):
        slv.run(*args(), piter=4, recover_prb=True)
        waited = [0.0]
        ev_sync, st_sync = torch.cuda.Event.synchronize, torch.cuda.Stream.synchronize

        def timed(f):
            def g(self):
                t = time.perf_counter()
                f(self)
                waited[0] += time.perf_counter() - t
            return g
        torch.cuda.Event.synchronize = timed(ev_sync)
        torch.cuda.Stream.synchronize = timed(st_sync)
        for dls in (True, False):
            slv.device_line_search = dls
            a = args()
            torch.cuda.synchronize()
            waited[0] = 0.0
            t0 = time.perf_counter()
            slv.run(*a, piter=iters, recover_prb=True)
            t1 = time.perf_counter()
            torch.cuda.synchronize()
            t2 = time.perf_counter()
            sys.stderr.write("device_line_search=%s: host returned after %.2f ms, GPU done after %.2f ms "
                             "(%d iterations, %.0f it/s), refits %d; the host spent %.2f ms waiting for the GPU\n"
                             % (dls, 1e3 * (t1 - t0), 1e3 * (t2 - t0), iters, iters / (t2 - t0), slv.ls_refits,
                                1e3 * waited[0]))
        torch.cuda.Event.synchronize, torch.cuda.Stream.synchronize = ev_sync, st_sync
        slv.device_line_search = True
        a = args()
        pr = cProfile.Profile()
        pr.enable()
        slv.run(*a, piter=iters, recover_prb=True)
        torch.cuda.synchronize()
        pr.disable()
    st = io.StringIO()
    pstats.Stats(pr, stream=st).sort_stats("tottime").print_stats(22)
    sys.stderr.write(st.getvalue())
