"""The float64 GPU referee (oracle/ref_gpu.py: F64PtychoOps / F64CGPtychoSolver).

1. Pinning: the torch float64 restatement equals the NumPy float64 restatement
   (oracle/numpy_ptycho.py under float64_arithmetic(), itself pinned by tests/test_oracle.py) to
   1e-12 on operators and on a short CG run -- so it can stand in for it at sizes NumPy cannot reach.
2. Accuracy of the two fp32 implementations against that exact answer, operator by operator: the
   sm_100a kernels must be at least as accurate as the reference's cuFFT path (within 1.5x), at
   every detector size.  This is what bounds how far two correct fp32 CG trajectories can drift
   apart (tests/test_gpu_cg.py, tests/test_gpu_fullsize.py).
"""
import numpy as np
import pytest
import torch

import workloads
from oracle import numpy_ptycho as O
from oracle import ref_gpu
from util import rel_l2

pytestmark = pytest.mark.gpu


def _cuda(x):
    return torch.from_numpy(np.ascontiguousarray(x)).cuda()


def _case(ndet, nprb, T, S, seed):
    nz, n = nprb + 70, nprb + 90
    w = workloads.synth_angles(T, nz, n, ndet, nprb, 1, 1, seed0=seed)
    rng = np.random.default_rng(seed)
    scan = np.stack([rng.uniform(0, nz - nprb - 1.001, (T, S)),
                     rng.uniform(0, n - nprb - 1.001, (T, S))], axis=-1).astype(np.float32)
    scan[0, 1] = -1.0
    return w["psi"], scan, np.ascontiguousarray(w["probe"][:, 0]), nz, n


@pytest.mark.parametrize("ndet,nprb", [(64, 64), (128, 96)])
def test_f64_gpu_operators_equal_numpy_f64(ndet, nprb):
    psi, scan, prb, nz, n = _case(ndet, nprb, 2, 7, 3)
    with O.float64_arithmetic():
        g0 = O.fwd(psi, scan, prb, ndet)
        f0 = O.adj(g0, scan, prb, nz, n)
        q0 = O.adj_probe(g0, scan, psi, nprb)
    ops = ref_gpu.F64PtychoOps(7, nprb, ndet, 2, nz, n)
    g = ops.fwd(_cuda(psi), _cuda(scan), _cuda(prb))
    assert g.dtype == torch.complex128
    assert rel_l2(g.cpu().numpy(), g0) < 1e-12
    assert rel_l2(ops.adj(_cuda(g0), _cuda(scan), _cuda(prb)).cpu().numpy(), f0) < 1e-12
    assert rel_l2(ops.adj_probe(_cuda(g0), _cuda(scan), _cuda(psi)).cpu().numpy(), q0) < 1e-12


def test_f64_gpu_solver_equals_numpy_f64():
    """Three CG iterations (position correction on, probe recovery) in float64: torch == NumPy."""
    c = workloads.c1_adjoint(nscan=12)
    psi, scan, probe = c["psi"], c["scan"], c["probe"]
    with O.float64_arithmetic():
        data = (np.abs(O.fwd(psi, scan, np.ascontiguousarray(probe[:, 0]), 128)) ** 2)
    data = data.astype(np.float32)
    init = np.ascontiguousarray(probe.swapaxes(2, 3))
    psi0 = np.ones_like(psi)
    with O.float64_arithmetic():
        want = O.cg_run(data, psi0, scan.copy(), init.copy(), 3, "gaussian", True, position_correction=True)
    with ref_gpu.F64CGPtychoSolver(12, 128, 128, 1, 276, 600) as slv:
        slv.position_correction = True
        got = slv.run_batch(data, psi0, scan, init, piter=3, model="gaussian", recover_prb=True, verbose=False)
    assert rel_l2(got["psi"], want["psi"]) < 1e-9
    assert rel_l2(got["probe"], want["probe"]) < 1e-9


@pytest.mark.skipif(not ref_gpu.available(), reason="oracle/_ref not built")
@pytest.mark.parametrize("ndet,S", [(64, 1024), (128, 600), (256, 300), (512, 80)])
def test_operator_accuracy_vs_exact(ndet, S):
    """|| fp32 result - exact || / || exact || of both fp32 implementations, per operator."""
    import libtike.cufft as pt
    psi, scan, prb, nz, n = _case(ndet, ndet, 1, S, ndet)
    psi_d, scan_d, prb_d = _cuda(psi), _cuda(scan), _cuda(prb)
    ex = ref_gpu.F64PtychoOps(S, ndet, ndet, 1, nz, n)
    g_ex = ex.fwd(psi_d, scan_d, prb_d)
    # a generic (not band-limited) far field for the adjoints, exactly representable in fp32
    g_in = (g_ex * torch.exp(1j * 0.37 * torch.arange(ndet, device="cuda"))[None, None, None, :]).to(torch.complex64)
    f_ex = ex.adj(g_in, scan_d, prb_d)
    q_ex = ex.adj_probe(g_in, scan_d, psi_d)

    def err(a, b):
        return float(torch.linalg.norm(a.to(torch.complex128) - b) / torch.linalg.norm(b))
    out = {}
    with pt.PtychoCuFFT(S, ndet, ndet, 1, nz, n) as slv, ref_gpu.RefPtychoFFT(S, ndet, ndet, 1, nz, n) as ref:
        for name, o in (("fused", slv), ("reference", ref)):
            out[name] = (err(o.fwd(psi_d, scan_d, prb_d), g_ex), err(o.adj(g_in, scan_d, prb_d), f_ex),
                         err(o.adj_probe(g_in, scan_d, psi_d), q_ex))
    print("accuracy vs float64, %d^2 x %d: fused fwd %.2e adj %.2e adj_probe %.2e | reference fwd %.2e adj %.2e "
          "adj_probe %.2e" % ((ndet, S) + out["fused"] + out["reference"]))
    for a, b in zip(out["fused"], out["reference"]):
        assert a < 1e-5
        assert a < 1.5 * b + 2e-8
