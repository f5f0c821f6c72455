import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "libtike-cufft_b200")); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
import libtike.cufft as pt
from test_gpu_cg import _problem
nd = int(sys.argv[1]) if len(sys.argv) > 1 else 128
data, psi0, scan, prb0 = _problem(1, 16, "poisson", nd)
ns = scan.shape[1]
nz, n = psi0.shape[1:]
with pt.CGPtychoSolver(ns, nd, nd, 1, nz, n) as slv:
    got = slv.run_batch(data, psi0, scan, prb0, piter=2, model="poisson", recover_prb=True)
    print(slv.history, np.abs(got["psi"]).max())
