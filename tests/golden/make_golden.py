"""Generate golden vectors with the REFERENCE's own compiled CUDA/cuFFT code (run on the GPU box).

    gpurun -- python tests/golden/make_golden.py          # writes gpurun_out/golden/*.npz
    cp gpurun_out/golden/*.npz tests/golden/              # here, then commit

Operators come straight from oracle/_ref/libptychofft_ref.so (= /root/reference/src/cuda/*.cu
compiled unmodified by oracle/Makefile); CG runs go through oracle/ref_gpu.py's statement-by-
statement torch restatement of src/libtike/cufft/ptycho.py:283-488 driving those operators.
Inputs are small seeded cases built from the reference's fixtures (workloads.py).
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import workloads  # noqa: E402
from oracle import ref_gpu  # noqa: E402

OUT = os.path.join(ROOT, "gpurun_out", "golden")


def small_case(ndet, nprb, nz, n, nscan, nmodes, seed):
    rng = np.random.default_rng(seed)
    m = workloads.model()
    obj = (m["initpsiamp"] * np.exp(1j * m["initpsiang"])).astype(np.complex64)
    psi = np.ascontiguousarray(obj[20:20 + nz, 200:200 + n])[None]
    probe = workloads.synth_probe(nprb, nmodes)
    co = m["coords"]
    idx = rng.choice(co.shape[1], nscan, replace=False)
    scan = np.zeros((1, nscan, 2), dtype=np.float32)
    scan[0, :, 0] = np.mod(co[1, idx], nz - nprb - 2)
    scan[0, :, 1] = np.mod(co[0, idx], n - nprb - 2)
    return psi, scan, probe


def ops(name, ndet, nprb, nz, n, nscan, seed, skip_one=True):
    psi, scan, probe = small_case(ndet, nprb, nz, n, nscan, 1, seed)
    if skip_one:
        scan[0, 1] = -1  # the reference's "undefined position" sentinel (tests/test_fsc.py:16-17)
    prb = probe[:, 0]
    with ref_gpu.RefPtychoFFT(nscan, nprb, ndet, 1, nz, n) as ref:
        g = ref.fwd_ptycho_batch(psi, scan, prb)
        f = ref.adj_ptycho_batch(g, scan, prb)
        q = ref.adj_ptycho_batch_prb(g, scan, psi)
    np.savez(os.path.join(OUT, name), psi=psi, scan=scan, probe=prb, ndet=ndet, fwd=g, adj=f,
             adj_probe=q)
    print(name, "fwd", g.shape, "adj", f.shape, "adj_probe", q.shape)


def cg(name, ndet, nz, n, nscan, nmodes, piter, model, seed, noisy=False):
    psi_true, scan, probe = small_case(ndet, ndet, nz, n, nscan, nmodes, seed)
    with ref_gpu.RefCGPtychoSolver(nscan, ndet, ndet, 1, nz, n) as ref:
        data = np.zeros((1, nscan, ndet, ndet), dtype=np.float32)
        for k in range(nmodes):
            data += np.abs(ref.fwd_ptycho_batch(psi_true, scan, probe[:, k])) ** 2
        if noisy:
            rng = np.random.default_rng(seed + 1)
            data = rng.poisson(data * (50.0 / data.mean())).astype(np.float32)
        psi0 = np.ones_like(psi_true)
        probe0 = probe.copy()
        if nmodes == 1:
            probe0 = np.ascontiguousarray(probe0.swapaxes(2, 3))  # tests/test.py:58
        else:
            for k in range(nmodes):
                probe0[:, k] /= np.max(np.abs(probe0[:, k]))      # tests/test_modes.py:41-42
        hist = []
        res = ref.run_batch(data, psi0, scan, probe0, piter=piter, model=model, recover_prb=True,
                            history=hist, verbose=False)
    steps = np.array([t[2] for t in ref.last_trials], dtype=np.float64)  # raw line-search results
    np.savez(os.path.join(OUT, name), data=data, psi0=psi0, probe0=probe0, scan=scan, piter=piter,
             model=model, psi=res["psi"], probe=res["probe"], history=np.array(hist), steps=steps)
    print(name, "history", hist)


if __name__ == "__main__":
    assert torch.cuda.is_available() and ref_gpu.available()
    os.makedirs(OUT, exist_ok=True)
    ops("ref_ops_c1.npz", 128, 128, 160, 260, 6, 0)
    ops("ref_ops_pad.npz", 64, 48, 100, 120, 5, 1)
    cg("ref_cg_gauss.npz", 64, 96, 112, 16, 1, 4, "gaussian", 2)
    cg("ref_cg_modes.npz", 64, 96, 112, 16, 2, 3, "gaussian", 3)
    # Poisson likelihood on noise-free intensities (well conditioned) and on Poisson-noised counts
    # (the reference's d*F/(|F|^2+1e-32) amplifies fp32 FFT rounding there: see test_oracle.py)
    cg("ref_cg_poisson.npz", 64, 96, 112, 16, 1, 3, "poisson", 4)
    cg("ref_cg_poisson_noisy.npz", 64, 96, 112, 16, 1, 3, "poisson", 4, noisy=True)
