"""CPU tests of the oracle itself (no GPU): the reference's own adjoint check and golden vectors."""
import os

import numpy as np
import pytest

import workloads
from oracle import numpy_ptycho as O
from util import rel_l2

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def test_adjoint_identity_reference_fixture():
    """tests/test_adjoint.py:42-59 of the reference, on its own fixtures (tolerance 1e-3 there)."""
    c = workloads.c1_adjoint(nscan=100)
    psi, scan, prb = c["psi"], c["scan"], c["probe"]
    t1 = O.fwd(psi, scan, prb[:, 0], 128)
    t2 = O.adj(t1, scan, prb[:, 0], 276, 600)
    t3 = O.adj_probe(t1, scan, psi, 128)
    a = np.sum(psi * np.conj(t2))
    b = np.sum(t1 * np.conj(t1))
    c_ = np.sum(prb[:, 0] * np.conj(t3))
    assert abs(a - b) / abs(a) < 1e-5
    assert abs(a - c_) / abs(a) < 1e-5
    # survey-time sanity value of the three inner products
    assert abs(a.real - 60304.69) < 0.5


def test_patch_origin_integer_work():
    scan = np.array([[[0.0, 0.0], [3.75, 7.25], [-1.0, 4.0], [5.0, -1.0], [-0.5, 2.5],
                      [147.999, 471.5]]], dtype=np.float32)
    R, C, keep = O.patch_origin(scan)
    assert R.tolist() == [[0, 3, -1, 5, 0, 147]]
    assert C.tolist() == [[0, 7, 4, -1, 2, 471]]
    # the sentinel -1 skips; (-1, 0) does not (modff yields -0.0, kernels.cu:39)
    assert keep.tolist() == [[True, True, False, False, True, True]]


def test_skipped_positions_give_zero():
    c = workloads.c1_adjoint(nscan=4)
    scan = c["scan"].copy()
    scan[0, 1] = -1
    g = O.fwd(c["psi"], scan, c["probe"][:, 0], 128)
    assert np.all(g[0, 1] == 0) and np.any(g[0, 0] != 0)


def test_zero_padding_is_centred():
    """ndet > nprb: probe window at offset (ndet-nprb)/2 in both axes (kernels.cu:48-57, Q16)."""
    rng = np.random.default_rng(0)
    psi = (rng.random((1, 80, 90)) + 1j * rng.random((1, 80, 90))).astype(np.complex64)
    prb = (rng.random((1, 24, 24)) + 1j * rng.random((1, 24, 24))).astype(np.complex64)
    scan = np.array([[[10.0, 20.0]]], dtype=np.float32)
    g = O.fwd(psi, scan, prb, 32)
    near = np.fft.ifft2(g[0, 0])
    want = np.zeros((32, 32), dtype=np.complex64)
    want[4:28, 4:28] = prb[0] * psi[0, 10:34, 20:44] / 32
    assert rel_l2(near, want) < 1e-6


@pytest.mark.parametrize("name", ["ref_ops_c1.npz", "ref_ops_pad.npz"])
def test_oracle_matches_reference_golden_ops(name):
    """Golden vectors produced by the reference's compiled CUDA/cuFFT code on a B200
    (tests/golden/make_golden.py) pin the NumPy restatement."""
    path = os.path.join(GOLD, name)
    if not os.path.exists(path):
        pytest.skip("golden vectors not generated yet (parity unpinned)")
    z = np.load(path)
    psi, scan, prb = z["psi"], z["scan"], z["probe"]
    ndet = int(z["ndet"])
    nz, n = psi.shape[1:]
    g = O.fwd(psi, scan, prb, ndet)
    assert rel_l2(g, z["fwd"]) < 1e-5
    assert rel_l2(O.adj(z["fwd"], scan, prb, nz, n), z["adj"]) < 1e-5
    assert rel_l2(O.adj_probe(z["fwd"], scan, psi, prb.shape[-1]), z["adj_probe"]) < 1e-5


def _replay(z, **kw):
    # the reference's own line-search decisions are replayed: with fp32 cost sums (Poisson costs are
    # ~1e7 with steps of 1) a near-tie may be decided either way, which would fork the trajectory
    forced = z["steps"].tolist() if "steps" in z.files else None
    return O.cg_run(z["data"], z["psi0"], z["scan"], z["probe0"].copy(), int(z["piter"]),
                    str(z["model"]), True, forced_steps=forced, **kw)


@pytest.mark.parametrize("name", ["ref_cg_gauss.npz", "ref_cg_modes.npz"])
def test_oracle_cg_matches_reference_golden(name):
    path = os.path.join(GOLD, name)
    if not os.path.exists(path):
        pytest.skip("golden vectors not generated yet (parity unpinned)")
    z = np.load(path)
    hist = []
    res = _replay(z, history=hist)
    assert rel_l2(res["psi"], z["psi"]) < 1e-4
    assert rel_l2(res["probe"], z["probe"]) < 1e-4
    # the printed cost (function min) follows the reference too
    assert np.allclose([h[3] for h in hist], z["history"][:, 3], rtol=2e-5)
    # free-running decisions agree as well on these cases
    free = O.cg_run(z["data"], z["psi0"], z["scan"], z["probe0"].copy(), int(z["piter"]),
                    str(z["model"]), True)
    assert rel_l2(free["psi"], z["psi"]) < 1e-4


@pytest.mark.parametrize("name", ["ref_cg_poisson.npz", "ref_cg_poisson_noisy.npz"])
def test_oracle_poisson_is_rounding_limited(name):
    """Poisson likelihood: the reference's gradient d*F/(|F|^2 + 1e-32) (ptycho.py:360, 438)
    divides by the far field, so wherever the model intensity is far below the data (weak pixels
    that recorded a photon; the flat starting object of tests/test.py) the fp32 rounding error of
    the FFT is amplified and two correct fp32 implementations (cuFFT, pocketfft) differ by 1e-4 to
    1e-2 after 3 iterations.  The bar there is anchored on the float64 restatement: the fp32
    restatement must be as close to the exact trajectory as the reference's own run is."""
    path = os.path.join(GOLD, name)
    if not os.path.exists(path):
        pytest.skip("golden vectors not generated yet (parity unpinned)")
    z = np.load(path)
    hist = []
    res32 = _replay(z, history=hist)
    assert np.allclose([h[3] for h in hist], z["history"][:, 3], rtol=2e-4)  # cost follows the reference
    with O.float64_arithmetic():
        res64 = _replay(z)
    for key in ("psi", "probe"):
        e_ref = rel_l2(z[key], res64[key])
        e_32 = rel_l2(res32[key], res64[key])
        print(name, key, "reference vs f64 %.2e   oracle(f32) vs f64 %.2e" % (e_ref, e_32))
        assert e_32 < max(3 * e_ref, 1e-4)
        assert e_ref < 5e-2  # and the reference itself stays in the neighbourhood of the exact run
