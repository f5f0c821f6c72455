"""GPU parity of the fused CG solver against the reference's operators + solver restatement.

Bar (north_star): relative L2 <= 1e-4 on psi and probe after a fixed number of CG iterations.
"""
import os

import numpy as np
import pytest
import torch

import workloads
from oracle import numpy_ptycho as O
from oracle import ref_gpu
from util import rel_l2

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
TOL = 1e-4


def _pt():
    import libtike.cufft as pt
    return pt


def _problem(nmodes, nscan, model, ndet=128, seed=0):
    """tests/test.py / tests/test_modes.py style problem on the reference fixtures, scaled down."""
    if ndet == 128:
        c = workloads.c3_modes(nmodes, nscan) if nmodes > 1 else workloads.c1_adjoint(nscan)
        psi, scan, probe = c["psi"], c["scan"], c["probe"]
        init = c["probe_init"] if nmodes > 1 else np.ascontiguousarray(probe.swapaxes(2, 3))
    else:
        w = workloads.synth_angles(1, 200, 220, ndet, ndet, int(np.sqrt(nscan)), nmodes, seed0=seed)
        psi, scan, probe = w["psi"], w["scan"], w["probe"]
        init = probe * (0.9 + 0.1j)
        nscan = scan.shape[1]
    data = np.zeros((1, scan.shape[1], ndet, ndet), dtype=np.float32)
    for k in range(nmodes):
        data += np.abs(O.fwd(psi, scan, np.ascontiguousarray(probe[:, k]), ndet)) ** 2
    if model == "poisson":
        rng = np.random.default_rng(seed + 1)
        data = rng.poisson(data * (50.0 / data.mean())).astype(np.float32)
    return data, np.ones_like(psi), scan, init.astype(np.complex64)


@pytest.mark.skipif(not ref_gpu.available(), reason="oracle/_ref not built")
@pytest.mark.parametrize("nmodes,nscan,model,piter,ndet", [
    (1, 100, "gaussian", 8, 128),
    (1, 100, "gaussian", 32, 128),
    (3, 60, "gaussian", 8, 128),
    (1, 100, "poisson", 8, 128),
    (2, 49, "poisson", 6, 64),
    (1, 64, "gaussian", 8, 64),
])
def test_cg_vs_reference_gpu(nmodes, nscan, model, piter, ndet):
    pt = _pt()
    data, psi0, scan, prb0 = _problem(nmodes, nscan, model, ndet)
    nscan = scan.shape[1]
    nz, n = psi0.shape[1:]
    with ref_gpu.RefCGPtychoSolver(nscan, ndet, ndet, 1, nz, n) as ref:
        hist = []
        want = ref.run_batch(data, psi0, scan, prb0, piter=piter, model=model, recover_prb=True,
                             history=hist, verbose=False)
    with pt.CGPtychoSolver(nscan, ndet, ndet, 1, nz, n) as slv:
        got = slv.run_batch(data, psi0, scan, prb0, piter=piter, model=model, recover_prb=True)
    e_psi, e_prb = rel_l2(got["psi"], want["psi"]), rel_l2(got["probe"], want["probe"])
    print("cg parity", nmodes, nscan, model, piter, ndet, e_psi, e_prb, hist[-1])
    assert e_psi < TOL and e_prb < TOL


def test_cg_vs_numpy_oracle_small():
    """Independent check against the CPU restatement (different FFT, fp64-accumulated adjoint)."""
    pt = _pt()
    data, psi0, scan, prb0 = _problem(1, 12, "gaussian")
    want = O.cg_run(data, psi0, scan, prb0.copy(), 3, "gaussian", True)
    with pt.CGPtychoSolver(12, 128, 128, 1, 276, 600) as slv:
        got = slv.run_batch(data, psi0, scan, prb0, piter=3, model="gaussian", recover_prb=True)
    assert rel_l2(got["psi"], want["psi"]) < TOL
    assert rel_l2(got["probe"], want["probe"]) < TOL


def test_cg_fixed_probe():
    """recover_prb=False branch (tests/test.py:60)."""
    pt = _pt()
    data, psi0, scan, _ = _problem(1, 40, "gaussian")
    prb = workloads.fixture_probe(1)
    want = O.cg_run(data, psi0, scan, prb.copy(), 4, "gaussian", False)
    with pt.CGPtychoSolver(40, 128, 128, 1, 276, 600) as slv:
        got = slv.run_batch(data, psi0, scan, prb, piter=4, model="gaussian", recover_prb=False)
    assert rel_l2(got["psi"], want["psi"]) < TOL
    assert rel_l2(got["probe"], want["probe"]) < TOL


@pytest.mark.parametrize("name", ["ref_cg_gauss.npz", "ref_cg_modes.npz", "ref_cg_poisson.npz"])
def test_cg_vs_golden(name):
    path = os.path.join(GOLD, name)
    if not os.path.exists(path):
        pytest.skip("golden vectors not generated yet")
    pt = _pt()
    z = np.load(path)
    data, scan = z["data"], z["scan"]
    ndet, nscan = data.shape[-1], scan.shape[1]
    nz, n = z["psi0"].shape[1:]
    with pt.CGPtychoSolver(nscan, ndet, ndet, 1, nz, n) as slv:
        got = slv.run_batch(data, z["psi0"], scan, z["probe0"], piter=int(z["piter"]),
                            model=str(z["model"]), recover_prb=True)
    assert rel_l2(got["psi"], z["psi"]) < TOL
    assert rel_l2(got["probe"], z["probe"]) < TOL


def test_cg_cost_decreases_c2():
    """C2 at full size: the Gaussian cost printed by the solver must go down monotonically enough."""
    pt = _pt()
    w = workloads.c2_single_angle()
    psi, scan, probe = w["psi"], w["scan"], w["probe"]
    with pt.CGPtychoSolver(1024, 128, 128, 1, 512, 512) as slv:
        data = np.abs(slv.fwd_ptycho_batch(psi, scan, probe[:, 0])) ** 2
        d, s, p = (torch.from_numpy(x).cuda() for x in (data, scan, probe))
        psi0 = torch.ones((1, 512, 512), dtype=torch.complex64, device="cuda")
        c0 = float(slv._intensity(psi0, s, p, d, None, 0)[2])
        res = slv.run(d, psi0, s, p.clone(), piter=16, model="gaussian", recover_prb=False)
        c1 = float(slv._intensity(res["psi"], s, res["probe"], d, None, 0)[2])
    assert c1 < 0.2 * c0
