"""GPU parity of the position-correction kernels (ptx_register_translation, ptx_cg_position_shifts)
against the restatements of src/libtike/cufft/ptycho.py:163-248 and :398-403.

Shifts are multiples of 1/upsample_factor picked by an argmax, so parity is EXACT equality wherever
the peak is not a near-tie; the tests use images with a clear peak and allow no mismatch there.
"""
import numpy as np
import pytest
import torch

import workloads
from oracle import numpy_ptycho as O
from oracle import ref_gpu
from test_register_oracle import smooth_images, fourier_shift
from util import rel_l2, ReplaySolver

pytestmark = pytest.mark.gpu


def _pt():
    import libtike.cufft as pt
    return pt


def _shifts(S, seed):
    rng = np.random.default_rng(seed)
    sh = rng.uniform(-6, 6, size=(S, 2))
    sh[0] = 0
    sh[1] = (0.37, -1.62)
    return sh


@pytest.mark.parametrize("N", [64, 128, 256, 512])
@pytest.mark.parametrize("space", ["fourier", "real"])
def test_register_translation_vs_oracle(N, space):
    pt = _pt()
    S = 9 if N <= 128 else 4
    img = smooth_images(S, N, seed=N)
    F = np.fft.fft2(img).astype(np.complex64)
    true = _shifts(S, N)
    G = fourier_shift(F, true)
    if space == "real":
        a, b = img, np.fft.ifft2(G).astype(np.complex64)
    else:
        a, b = F, G
    want = O.register_translation_batch(a, b, 100, space)
    got = pt.register_translation_batch(torch.from_numpy(a).cuda(), torch.from_numpy(b).cuda(), 100,
                                        space).cpu().numpy()
    assert got.dtype == np.float64 and got.shape == (S, 2)
    assert np.abs(want + true).max() < 0.0051
    assert np.array_equal(got, want), (got - want)
    whole = pt.register_translation_batch(torch.from_numpy(a).cuda(), torch.from_numpy(b).cuda(), 1,
                                          space).cpu().numpy()
    assert np.array_equal(whole, O.register_translation_batch(a, b, 1, space))
    coarse = pt.register_translation_batch(torch.from_numpy(a).cuda(), torch.from_numpy(b).cuda(), 10,
                                           space).cpu().numpy()
    assert np.array_equal(coarse, O.register_translation_batch(a, b, 10, space))


def test_register_vs_reference_restatement_on_gpu():
    """cp -> torch restatement (cuFFT + complex128 einsum) on the same device."""
    pt = _pt()
    S, N = 64, 128
    img = smooth_images(S, N, seed=11)
    F = np.fft.fft2(img).astype(np.complex64)
    G = fourier_shift(F, _shifts(S, 5))
    Fd, Gd = torch.from_numpy(F).cuda(), torch.from_numpy(G).cuda()
    want = ref_gpu.register_translation_batch(Fd, Gd, 100, "fourier").cpu().numpy()
    got = pt.register_translation_batch(Fd, Gd, 100, "fourier").cpu().numpy()
    # same inputs, but torch's complex64 product and cuFFT's ifft2 round differently from the fused
    # kernel: a true shift that sits half way between two 0.01 px grid points is a coin toss
    bad = np.abs(got - want).max(axis=1) > 0
    print("positions that differ:", int(bad.sum()), "of", S, "worst", np.abs(got - want).max())
    assert np.abs(got - want).max() <= 0.0100001 and bad.sum() <= 3


def test_register_batch_of_one_and_errors():
    pt = _pt()
    img = smooth_images(1, 64)
    F = np.fft.fft2(img).astype(np.complex64)
    G = fourier_shift(F, np.array([[2.25, -3.5]]))
    got = pt.register_translation_batch(torch.from_numpy(F).cuda(), torch.from_numpy(G).cuda(), 100,
                                        "fourier")
    assert np.array_equal(got.cpu().numpy(), np.zeros((1, 2)))
    with pytest.raises(pt.ptycho.PtxError):
        pt.register_translation_batch(torch.from_numpy(F).cuda(), torch.from_numpy(G).cuda(), 1000,
                                      "fourier")
    with pytest.raises(ValueError):
        pt.register_translation_batch(torch.from_numpy(F).cuda(), torch.from_numpy(G[:, :32]).cuda())


@pytest.mark.parametrize("ndet,nprb", [(64, 64), (128, 128), (128, 96), (256, 256)])
def test_cg_position_shifts_vs_oracle(ndet, nprb):
    """The fused step of ptycho.py:398-403: fwd(psi, ones), fwd(psi', ones), register -- psi' is psi
    moved by a known sub-pixel amount, plus a skipped position (scan < 0)."""
    pt = _pt()
    import ctypes
    side = 3
    nz, n = ndet + 40, ndet + 52
    w = workloads.synth_angles(1, nz, n, nprb, ndet, side, 1, seed0=4)
    psi, scan = w["psi"], w["scan"].copy()
    scan[0, 4] = -1.0
    mv = np.array([[0.43, -0.27]])
    k0, k1 = np.fft.fftfreq(nz), np.fft.fftfreq(n)
    ph = np.exp(-2j * np.pi * (mv[0, 0] * k0[:, None] + mv[0, 1] * k1[None, :]))
    psi_b = np.fft.ifft2(np.fft.fft2(psi[0]) * ph).astype(np.complex64)[None]
    ones = np.ones((1, nprb, nprb), dtype=np.complex64)
    t1 = O.fwd(psi, scan, ones, ndet)[0]
    t2 = O.fwd(psi_b, scan, ones, ndet)[0]
    want = O.register_translation_batch(t1, t2, 100, "fourier")
    S = side * side
    with pt.CGPtychoSolver(S, nprb, ndet, 1, nz, n) as slv:
        out = torch.empty((S, 2), dtype=torch.float64, device="cuda")
        d = [torch.from_numpy(x).cuda() for x in (psi, psi_b, scan)]
        pt.ptycho.check(pt.ptycho.lib.ptx_cg_position_shifts(
            slv._h, ctypes.c_void_p(d[0].data_ptr()), ctypes.c_void_p(d[1].data_ptr()),
            ctypes.c_void_p(d[2].data_ptr()), 100, ctypes.c_void_p(out.data_ptr()),
            pt.ptycho.current_stream()))
        got = out.cpu().numpy()
    print("shifts", want[:3], got[:3])
    assert np.array_equal(got[4], [-0.75, -0.75]) and np.array_equal(want[4], [-0.75, -0.75])
    bad = np.abs(got - want).max(axis=1) > 0
    # float32 far fields from two different FFTs: allow a near-tie to fall one 0.01 step apart
    assert np.abs(got - want).max() <= 0.0100001 and bad.sum() <= 1, (got - want)


@pytest.mark.parametrize("model", ["gaussian", "poisson"])
def test_cg_with_position_correction_vs_oracle(model):
    """`run` exactly as the reference executes it (position correction ON, Q5): psi, probe and the
    corrected scan positions against the NumPy restatement, 4 iterations."""
    pt = _pt()
    ndet, side = 64, 4
    w = workloads.synth_angles(1, 200, 220, ndet, ndet, side, 1, seed0=6)
    psi, scan, probe = w["psi"], w["scan"], w["probe"]
    data = (np.abs(O.fwd(psi, scan, np.ascontiguousarray(probe[:, 0]), ndet)) ** 2).astype(np.float32)
    init = np.ones_like(psi)
    prb0 = (probe * (0.9 + 0.1j)).astype(np.complex64)
    scan_o = scan.copy()
    log = []
    want = O.cg_run(data, init, scan_o, prb0.copy(), 4, model, True, position_correction=True,
                    shift_log=log)
    with pt.CGPtychoSolver(side * side, ndet, ndet, 1, 200, 220) as slv:
        assert slv.position_correction  # the reference's behaviour is the default
        slv.log_shifts = True
        d_scan = torch.from_numpy(scan.copy()).cuda()
        got = slv.run(torch.from_numpy(data).cuda(), torch.from_numpy(init).cuda(), d_scan,
                      torch.from_numpy(prb0.copy()).cuda(), 4, model=model, recover_prb=True)
        glog = [s.cpu().numpy() for s in slv.shift_log]
    assert len(glog) == len(log) == 3
    for a, b in zip(glog, log):
        print("shift step: max |ours - oracle| %.3f, max |oracle| %.3f" % (np.abs(a - b).max(), np.abs(b).max()))
    assert np.abs(d_scan.cpu().numpy() - scan_o).max() <= 0.0100001  # caller's scan mutated in place
    assert rel_l2(got["psi"].cpu().numpy(), want["psi"]) < 1e-4
    assert rel_l2(got["probe"].cpu().numpy(), want["probe"]) < 1e-4


def test_register_algorithms_agree():
    """The three forms of the upsampled matrix DFT -- low-rank Jacobi-Anger factorisation (default),
    the reference's two direct matrix products on the FP64 tensor cores, and the same on the scalar
    FP64 pipe -- must pick identical shifts (PTX_REG_ALGO is read once per process)."""
    import json
    import os
    import subprocess
    import sys
    here = os.path.dirname(os.path.abspath(__file__))
    code = (
        "import sys, json, numpy as np, torch\n"
        "sys.path[:0] = [%r, %r, %r]\n"
        "import libtike.cufft as pt\n"
        "from test_register_oracle import smooth_images, fourier_shift\n"
        "out = {}\n"
        "for N in (64, 128, 256):\n"
        "    img = smooth_images(24, N, seed=N + 1)\n"
        "    F = np.fft.fft2(img).astype(np.complex64)\n"
        "    sh = np.random.default_rng(N).uniform(-5, 5, size=(24, 2))\n"
        "    G = fourier_shift(F, sh)\n"
        "    out[N] = pt.register_translation_batch(torch.from_numpy(F).cuda(), torch.from_numpy(G).cuda(),\n"
        "                                           100, 'fourier').cpu().numpy().tolist()\n"
        "print('RESULT' + json.dumps(out))\n" % (here, os.path.dirname(here),
                                                   os.path.join(os.path.dirname(here), "libtike-cufft_b200")))
    res = {}
    for algo in ("lowrank", "dmma", "dfma"):
        env = dict(os.environ, PTX_REG_ALGO=algo)
        p = subprocess.run([sys.executable, "-c", code], env=env, capture_output=True, text=True, timeout=600)
        assert p.returncode == 0, p.stderr[-2000:]
        line = [ln for ln in p.stdout.splitlines() if ln.startswith("RESULT")][0]
        res[algo] = json.loads(line[6:])
    assert res["lowrank"] == res["dmma"] == res["dfma"]


@pytest.mark.skipif(not ref_gpu.available(), reason="oracle/_ref not built")
def test_c2_full_size_position_correction_vs_reference():
    """BASELINE.json configs[1] at full size (1024 positions, 128^2 detector, 512^2 object): three CG
    iterations exactly as the reference executes them (position correction on) against the
    reference's cuFFT operators + restated solver on the same GPU."""
    pt = _pt()
    w = workloads.c2_single_angle() if hasattr(workloads, "c2_single_angle") else None
    if w is None:
        w = workloads.synth_angles(1, 512, 512, 128, 128, 32, 1, seed0=0)
    psi, scan, probe = w["psi"], w["scan"], w["probe"]
    S = scan.shape[1]
    data = (np.abs(O.fwd(psi, scan, np.ascontiguousarray(probe[:, 0]), 128)) ** 2).astype(np.float32)
    init = np.ones_like(psi)
    prb0 = (probe * (0.9 + 0.1j)).astype(np.complex64)
    with ref_gpu.RefCGPtychoSolver(S, 128, 128, 1, 512, 512) as ref:
        ref.position_correction = True
        ref.shift_log = []
        want = ref.run_batch(data, init, scan, prb0, piter=3, model="gaussian", recover_prb=True,
                             verbose=False)
        steps = [t[2] for t in ref.last_trials]
        rlog = ref.shift_log
    with ReplaySolver(S, 128, 128, 1, 512, 512) as slv:
        slv.forced_steps = list(steps)
        got = slv.run_batch(data, init, scan, prb0, piter=3, model="gaussian", recover_prb=True)
        glog = [x.cpu().numpy() for x in slv.shift_log]
    nbad = sum(int((np.abs(a - b).max(axis=1) > 0).sum()) for a, b in zip(glog, rlog))
    worst = max(float(np.abs(a - b).max()) for a, b in zip(glog, rlog))
    print("c2 full size: %d of %d shifts differ (worst %.3f px, largest shift %.2f px); psi %.2e probe %.2e"
          % (nbad, S * len(rlog), worst, max(float(np.abs(b).max()) for b in rlog),
             rel_l2(got["psi"], want["psi"]), rel_l2(got["probe"], want["probe"])))
    assert worst <= 0.0100001 and nbad <= S * len(rlog) // 50
    assert rel_l2(got["psi"], want["psi"]) < 1e-4 and rel_l2(got["probe"], want["probe"]) < 1e-4
