"""GPU parity of the operators (through the C ABI) against the oracle and the compiled reference.

Bars (BASELINE.json north_star): bit-exact integer work (patch origin, window offset, skip rule);
relative L2 <= 1e-5 on fwd / adj / adj_probe outputs.
"""
import ctypes
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
TOL = 1e-5  # relative L2, fp32 complex operator outputs (north_star)


def _pt():
    import libtike.cufft as pt
    return pt


def _cuda(x):
    return torch.from_numpy(np.ascontiguousarray(x)).cuda()


def test_library_is_loaded_and_device_ok():
    from libtike.cufft.ptychofft import lib
    assert lib.ptx_device_ok() == 1


@pytest.mark.parametrize("nscan", [1, 7, 100])
def test_c1_operators_vs_oracle(nscan):
    """C1 (tests/test_adjoint.py sizes) against the NumPy/pocketfft restatement."""
    pt = _pt()
    c = workloads.c1_adjoint(nscan=nscan)
    psi, scan, prb = c["psi"], c["scan"], c["probe"][:, 0]
    with pt.PtychoCuFFT(nscan, 128, 128, 1, 276, 600) as slv:
        g = slv.fwd_ptycho_batch(psi, scan, prb)
        g0 = O.fwd(psi, scan, prb, 128)
        assert rel_l2(g, g0) < TOL
        assert rel_l2(slv.adj_ptycho_batch(g0, scan, prb), O.adj(g0, scan, prb, 276, 600)) < TOL
        assert rel_l2(slv.adj_ptycho_batch_prb(g0, scan, psi), O.adj_probe(g0, scan, psi, 128)) < TOL


def test_c1_adjoint_identity():
    """The reference's only quantitative test (tests/test_adjoint.py:42-59), same inputs."""
    pt = _pt()
    c = workloads.c1_adjoint(nscan=100)
    psi0, scan, prb0 = c["psi"], c["scan"], c["probe"]
    with pt.PtychoCuFFT(100, 128, 128, 1, 276, 600) as slv:
        t1 = slv.fwd_ptycho_batch(psi0, scan, prb0[:, 0])
        t2 = slv.adj_ptycho_batch(t1, scan, prb0[:, 0])
        t3 = slv.adj_ptycho_batch_prb(t1, scan, psi0)
    a = np.sum(psi0 * np.conj(t2))
    b = np.sum(t1 * np.conj(t1))
    c_ = np.sum(prb0[:, 0] * np.conj(t3))
    assert abs(a - b) / abs(a) < 1e-5 and abs(a - c_) / abs(a) < 1e-5
    assert abs(a.real - 60304.69) < 0.5


@pytest.mark.skipif(not ref_gpu.available(), reason="oracle/_ref not built")
@pytest.mark.parametrize("cfg", [(128, 128, 276, 600, 64), (128, 96, 276, 600, 16),
                                 (64, 64, 200, 220, 40), (64, 40, 200, 220, 9),
                                 (256, 256, 300, 420, 12), (256, 200, 300, 420, 5),
                                 (512, 512, 700, 640, 4), (512, 384, 700, 640, 3)])
def test_operators_vs_compiled_reference(cfg):
    """Same inputs through the reference's own CUDA/cuFFT code on the same GPU."""
    pt = _pt()
    ndet, nprb, nz, n, nscan = cfg
    w = workloads.synth_angles(2, nz, n, ndet, nprb, 1, 1, seed0=5)
    rng = np.random.default_rng(7)
    scan = np.stack([rng.uniform(0, nz - nprb - 1.5, (2, nscan)),
                     rng.uniform(0, n - nprb - 1.5, (2, nscan))], axis=-1).astype(np.float32)
    scan[1, 0] = -1.0
    psi = _cuda(w["psi"])
    prb = _cuda(w["probe"][:, 0])
    scan = _cuda(scan)
    with pt.PtychoCuFFT(nscan, nprb, ndet, 2, nz, n) as slv, \
            ref_gpu.RefPtychoFFT(nscan, nprb, ndet, 2, nz, n) as ref:
        g_ref = ref.fwd(psi, scan, prb)
        g = slv.fwd(psi, scan, prb)
        assert rel_l2(g.cpu().numpy(), g_ref.cpu().numpy()) < TOL
        f_ref = ref.adj(g_ref, scan, prb).cpu().numpy()
        f = slv.adj(g_ref, scan, prb).cpu().numpy()
        assert rel_l2(f, f_ref) < TOL
        q_ref = ref.adj_probe(g_ref, scan, psi).cpu().numpy()
        q = slv.adj_probe(g_ref, scan, psi).cpu().numpy()
        assert rel_l2(q, q_ref) < TOL


@pytest.mark.parametrize("name", ["ref_ops_c1.npz", "ref_ops_pad.npz"])
def test_operators_vs_golden(name):
    path = os.path.join(GOLD, name)
    if not os.path.exists(path):
        pytest.skip("golden vectors not generated yet")
    pt = _pt()
    z = np.load(path)
    psi, scan, prb, ndet = z["psi"], z["scan"], z["probe"], int(z["ndet"])
    nz, n = psi.shape[1:]
    with pt.PtychoCuFFT(scan.shape[1], prb.shape[-1], ndet, 1, nz, n) as slv:
        assert rel_l2(slv.fwd_ptycho_batch(psi, scan, prb), z["fwd"]) < TOL
        assert rel_l2(slv.adj_ptycho_batch(z["fwd"], scan, prb), z["adj"]) < TOL
        assert rel_l2(slv.adj_ptycho_batch_prb(z["fwd"], scan, psi), z["adj_probe"]) < TOL


@pytest.mark.parametrize("ndet,nprb", [(128, 128), (128, 100), (64, 64), (64, 30), (256, 256),
                                       (256, 130), (512, 512), (512, 78)])
def test_integer_work_bit_exact(ndet, nprb):
    """Patch origin, window offset (ndet-nprb)/2 and the skip rule, bit for bit.

    With integer scan positions and a probe of ones every near-plane value is an object value
    times the power of two 1/ndet, so the comparison with the oracle is exact.
    """
    from libtike.cufft.ptychofft import lib, check, current_stream
    pt = _pt()
    nz, n, nscan = max(150, nprb + 22), max(170, nprb + 42), 12
    rng = np.random.default_rng(3)
    psi = (rng.integers(-512, 512, (1, nz, n)) + 1j * rng.integers(-512, 512, (1, nz, n))).astype(np.complex64)
    prb = np.ones((1, nprb, nprb), dtype=np.complex64)
    scan = np.stack([rng.integers(0, nz - nprb - 1, (1, nscan)),
                     rng.integers(0, n - nprb - 1, (1, nscan))], axis=-1).astype(np.float32)
    scan[0, 3] = (-1.0, 5.0)
    scan[0, 4] = (2.0, -1.0)
    scan[0, 5] = (nz - nprb - 1, n - nprb - 1)   # largest valid origin
    patches, keep = O.gather_patches(psi[0], scan[0], nprb)
    o = (ndet - nprb) // 2
    want = np.zeros((1, nscan, ndet, ndet), dtype=np.complex64)
    want[0, :, o:o + nprb, o:o + nprb] = patches * np.float32(1.0 / ndet)
    want[0, ~keep] = 0
    with pt.PtychoCuFFT(nscan, nprb, ndet, 1, nz, n) as slv:
        near = torch.full((1, nscan, ndet, ndet), 7.0, dtype=torch.complex64, device="cuda")
        a, b, c_ = _cuda(psi), _cuda(scan), _cuda(prb)  # keep the device buffers alive
        check(lib.ptx_debug_nearplane(slv._h, ctypes.c_void_p(near.data_ptr()),
                                      ctypes.c_void_p(a.data_ptr()), ctypes.c_void_p(b.data_ptr()),
                                      ctypes.c_void_p(c_.data_ptr()), 0, current_stream()))
        torch.cuda.synchronize()
        got = near.cpu().numpy()
    assert np.array_equal(got.view(np.uint32), want.view(np.uint32))


def test_subpixel_nearplane_close():
    """Fractional positions: same near plane as the oracle to fp32 rounding (weights are fp32 products)."""
    from libtike.cufft.ptychofft import lib, check, current_stream
    pt = _pt()
    c = workloads.c1_adjoint(nscan=20)
    psi, scan, prb = c["psi"], c["scan"], c["probe"][:, 0]
    patches, _ = O.gather_patches(psi[0], scan[0], 128)
    want = (patches * prb[0][None]) * np.float32(1 / 128)
    with pt.PtychoCuFFT(20, 128, 128, 1, 276, 600) as slv:
        near = torch.zeros((1, 20, 128, 128), dtype=torch.complex64, device="cuda")
        a, b, c_ = _cuda(psi), _cuda(scan), _cuda(prb)
        check(lib.ptx_debug_nearplane(slv._h, ctypes.c_void_p(near.data_ptr()),
                                      ctypes.c_void_p(a.data_ptr()), ctypes.c_void_p(b.data_ptr()),
                                      ctypes.c_void_p(c_.data_ptr()), 0, current_stream()))
        got = near.cpu().numpy()[0]
    assert rel_l2(got, want) < 5e-7


def test_edge_positions_zero_extension():
    """Origins whose (P+1)-window crosses the object edge: the object is zero-extended (Q11)."""
    pt = _pt()
    nz, n, P = 140, 150, 128
    rng = np.random.default_rng(11)
    psi = (rng.random((1, nz, n)) + 1j * rng.random((1, nz, n))).astype(np.complex64)
    prb = (rng.random((1, P, P)) + 1j * rng.random((1, P, P))).astype(np.complex64)
    scan = np.array([[[nz - P, n - P], [nz - P - 0.5, n - P - 0.25], [0.0, 0.0]]], dtype=np.float32)
    big = np.zeros((1, nz + 2, n + 2), dtype=np.complex64)
    big[:, :nz, :n] = psi
    with pt.PtychoCuFFT(3, P, P, 1, nz, n) as slv:
        g = slv.fwd_ptycho_batch(psi, scan, prb)
        g0 = O.fwd(big, scan, prb, P)
        assert rel_l2(g, g0) < TOL
        f = slv.adj_ptycho_batch(g0, scan, prb)
        f0 = O.adj(g0, scan, prb, nz + 2, n + 2)[:, :nz, :n]
        assert rel_l2(f, f0) < TOL


@pytest.mark.parametrize("ndet", [256, 512])
def test_large_detector_operators_vs_oracle(ndet):
    """256^2 / 512^2 detectors (C4, C5): frame split into 4 / 16 tiles by the cross stage."""
    pt = _pt()
    nscan = 6 if ndet == 256 else 3
    w = workloads.synth_angles(1, ndet + 60, ndet + 90, ndet, ndet, 2, 1, seed0=4)
    psi, prb = w["psi"], w["probe"][:, 0]
    rng = np.random.default_rng(ndet)
    scan = np.stack([rng.uniform(0, 58.5, (1, nscan)), rng.uniform(0, 88.5, (1, nscan))],
                    axis=-1).astype(np.float32)
    scan[0, 1] = (58.25, 88.75)  # largest valid origin, fractional
    with pt.PtychoCuFFT(nscan, ndet, ndet, 1, ndet + 60, ndet + 90) as slv:
        g0 = O.fwd(psi, scan, prb, ndet)
        assert rel_l2(slv.fwd_ptycho_batch(psi, scan, prb), g0) < TOL
        assert rel_l2(slv.adj_ptycho_batch(g0, scan, prb), O.adj(g0, scan, prb, ndet + 60, ndet + 90)) < TOL
        assert rel_l2(slv.adj_ptycho_batch_prb(g0, scan, psi), O.adj_probe(g0, scan, psi, ndet)) < TOL


def test_raw_pointer_class_surface():
    """`ptychofft` keeps the reference's native surface: ctor kwargs, read-only attrs, free()."""
    from libtike.cufft.ptychofft import ptychofft
    o = ptychofft(ptheta=1, nz=276, n=600, nscan=10, detector_shape=128, probe_shape=128)
    assert (o.ptheta, o.nz, o.n, o.nscan, o.ndet, o.nprb) == (1, 276, 600, 10, 128, 128)
    with pytest.raises(AttributeError):
        o.ndet = 3
    c = workloads.c1_adjoint(nscan=10)
    psi, scan, prb = _cuda(c["psi"]), _cuda(c["scan"]), _cuda(c["probe"][:, 0])
    g = torch.zeros((1, 10, 128, 128), dtype=torch.complex64, device="cuda")
    o.fwd(g.data_ptr(), psi.data_ptr(), scan.data_ptr(), prb.data_ptr())
    f = torch.zeros_like(psi)
    o.adj(f.data_ptr(), g.data_ptr(), scan.data_ptr(), prb.data_ptr(), 0)
    assert rel_l2(g.cpu().numpy(), O.fwd(c["psi"], c["scan"], c["probe"][:, 0], 128)) < TOL
    assert float(f.abs().sum()) > 0
    o.free()
    o.free()  # idempotent (ptychofft.cu:49-57)
    from libtike.cufft.ptychofft import PtxError
    with pytest.raises(PtxError):
        o.fwd(g.data_ptr(), psi.data_ptr(), scan.data_ptr(), prb.data_ptr())


def test_unsupported_sizes_fail_loudly():
    pt = _pt()
    from libtike.cufft.ptychofft import PtxError
    with pytest.raises(PtxError):
        pt.PtychoCuFFT(10, 112, 112, 1, 276, 600)   # not a built size: no fallback
    with pytest.raises(PtxError):
        pt.PtychoCuFFT(10, 200, 128, 1, 276, 600)   # probe larger than detector


def test_multi_angle_and_strided_probe_view():
    """ptheta = 2 with probe[:, k] views of a [T, M, P, P] array (Q10)."""
    pt = _pt()
    w = workloads.synth_angles(2, 300, 310, 128, 128, 3, nmodes=3, seed0=2)
    psi, scan, probe = w["psi"], w["scan"], w["probe"]
    probe[1] *= 0.5 + 0.25j
    with pt.PtychoCuFFT(9, 128, 128, 2, 300, 310) as slv:
        pg = _cuda(probe)
        for k in range(3):
            g = slv.fwd(_cuda(psi), _cuda(scan), pg[:, k]).cpu().numpy()
            g0 = O.fwd(psi, scan, np.ascontiguousarray(probe[:, k]), 128)
            assert rel_l2(g, g0) < TOL


def test_c2_full_size_properties():
    """C2 (512^2 object, 1024 positions): linearity and the adjoint identity at full size."""
    pt = _pt()
    w = workloads.c2_single_angle()
    psi, scan, prb = _cuda(w["psi"]), _cuda(w["scan"]), _cuda(w["probe"][:, 0])
    with pt.PtychoCuFFT(1024, 128, 128, 1, 512, 512) as slv:
        g1 = slv.fwd(psi, scan, prb)
        psi2 = torch.roll(psi, 7, dims=2) * (0.3 - 0.8j)
        g2 = slv.fwd(psi2, scan, prb)
        g12 = slv.fwd(psi + psi2, scan, prb)
        assert float(torch.linalg.norm(g12 - g1 - g2) / torch.linalg.norm(g12)) < 1e-6
        f = slv.adj(g1, scan, prb)
        q = slv.adj_probe(g1, scan, psi)
        a = torch.sum(psi * torch.conj(f))
        b = torch.sum(g1 * torch.conj(g1))
        c = torch.sum(prb * torch.conj(q))
        assert abs(a - b) / abs(a) < 1e-5 and abs(a - c) / abs(a) < 1e-5


@pytest.mark.parametrize("ndet,nmodes,model", [(128, 1, "gaussian"), (64, 3, "poisson"), (256, 1, "gaussian")])
def test_grad_ptycho_batch_vs_oracle(ndet, nmodes, model):
    """The host-array fused gradient (what bench.py's e2e leg times): for every angle
    sum_k Q_k* F* [F Q_k psi (1 - sqrt(d)/sqrt(I))]  (gaussian; 1 - d/I for poisson), against the
    NumPy restatement of fwd / adj.  3 angles through a ptheta = 1 plan = 3 pipelined chunks."""
    pt = _pt()
    T, side = 3, 3
    nz, n = ndet + 40, ndet + 52
    w = workloads.synth_angles(T, nz, n, ndet, ndet, side, nmodes, seed0=9)
    psi, scan, probe = w["psi"], w["scan"], w["probe"]
    rng = np.random.default_rng(5)
    data = np.zeros((T, side * side, ndet, ndet), dtype=np.float32)
    for k in range(nmodes):
        data += np.abs(O.fwd(psi, scan, np.ascontiguousarray(probe[:, k]), ndet)) ** 2
    data *= rng.uniform(0.7, 1.3, data.shape).astype(np.float32)  # make the residual non-trivial
    psi1 = (psi * (0.8 + 0.3j)).astype(np.complex64)
    want = np.zeros_like(psi1)
    inten = np.zeros_like(data)
    far = [O.fwd(psi1, scan, np.ascontiguousarray(probe[:, k]), ndet) for k in range(nmodes)]
    for f in far:
        inten += np.abs(f) ** 2
    for k, f in enumerate(far):
        if model == "gaussian":
            r = f - np.sqrt(data) * f / (np.sqrt(inten) + np.float32(1e-32))
        else:
            r = f - data * f / (inten + np.float32(1e-32))
        want += O.adj(r.astype(np.complex64), scan, np.ascontiguousarray(probe[:, k]), nz, n)
    with pt.CGPtychoSolver(side * side, ndet, ndet, 1, nz, n) as slv:
        got = slv.grad_ptycho_batch(data, psi1, scan, probe, model=model)
        got2 = slv.grad_ptycho_batch(data, psi1, scan, probe, model=model)  # buffers are reused
    assert rel_l2(got, want) < 2e-5
    assert rel_l2(got2, want) < 2e-5
