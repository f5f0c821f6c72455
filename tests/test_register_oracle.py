"""CPU checks of the position-correction oracle (reference: src/libtike/cufft/ptycho.py:163-248).

The reference has no test for register_translation_batch; its restatements are pinned by known
answers (a Fourier-shifted image must come back with exactly that shift, to 1/upsample_factor) and
against each other (NumPy restatement vs the cp -> torch restatement used on the GPU box).
"""
import numpy as np
import torch
from scipy import ndimage

from oracle import numpy_ptycho as O
from oracle import ref_gpu


def smooth_images(S, N, seed=0, sigma=2.0):
    rng = np.random.default_rng(seed)
    img = rng.normal(size=(S, N, N)) + 1j * rng.normal(size=(S, N, N))
    img = np.stack([ndimage.gaussian_filter(i.real, sigma, mode="wrap")
                    + 1j * ndimage.gaussian_filter(i.imag, sigma, mode="wrap") for i in img])
    return img.astype(np.complex64)


def fourier_shift(F, shifts):
    """F: spectra [S,N,N]; returns the spectra of the images moved by `shifts` (row, col)."""
    N = F.shape[-1]
    k = np.fft.fftfreq(N)
    ph = np.exp(-2j * np.pi * (shifts[:, 0, None, None] * k[None, :, None]
                               + shifts[:, 1, None, None] * k[None, None, :]))
    return (F * ph).astype(np.complex64)


TRUE = np.array([[0.37, -1.62], [3.0, 2.5], [-0.01, 0.0], [10.33, -7.77], [0.0, 0.0], [-20.5, 30.99]])


def test_known_shifts_fourier_and_real():
    img = smooth_images(len(TRUE), 64)
    F = np.fft.fft2(img).astype(np.complex64)
    G = fourier_shift(F, TRUE)
    got = O.register_translation_batch(F, G, upsample_factor=100, space="fourier")
    # registering target onto source returns minus the applied shift; 0.01 px grid
    assert np.abs(got + TRUE).max() < 0.0051
    got_r = O.register_translation_batch(img, np.fft.ifft2(G).astype(np.complex64), 100, "real")
    assert np.abs(got_r - got).max() < 1e-12
    whole = O.register_translation_batch(F, G, upsample_factor=1, space="fourier")
    assert np.array_equal(whole, np.round(whole)) and np.abs(whole + TRUE).max() <= 0.5


def test_batch_of_one_is_zeroed():
    """ptycho.py:243-245 indexes the batch axis: one image -> zero shifts (reproduced)."""
    img = smooth_images(1, 64)
    F = np.fft.fft2(img).astype(np.complex64)
    G = fourier_shift(F, np.array([[2.25, -3.5]]))
    assert np.array_equal(O.register_translation_batch(F, G, 100, "fourier"), np.zeros((1, 2)))


def test_numpy_and_torch_restatements_agree():
    img = smooth_images(4, 64, seed=3)
    F = np.fft.fft2(img).astype(np.complex64)
    G = fourier_shift(F, np.array([[0.37, -1.62], [5.5, 0.25], [-0.8, 0.8], [0.0, 0.0]]))
    a = O.register_translation_batch(F, G, 100, "fourier")
    b = ref_gpu.register_translation_batch(torch.from_numpy(F), torch.from_numpy(G), 100, "fourier")
    assert np.abs(a - b.numpy()).max() < 1e-12


def test_cg_position_correction_moves_scan_only_for_angle0():
    import workloads
    w = workloads.synth_angles(2, 100, 104, 64, 64, 3, 1, seed0=2)
    psi, scan, probe = w["psi"], w["scan"], w["probe"]
    data = np.abs(O.fwd(psi, scan, np.ascontiguousarray(probe[:, 0]), 64)) ** 2
    s0 = scan.copy()
    log = []
    O.cg_run(data, np.ones_like(psi), scan, probe, 3, "gaussian", True, position_correction=True,
             shift_log=log)
    assert len(log) == 2 and log[0].shape == (9, 2)
    assert np.array_equal(scan[1], s0[1])
    assert np.allclose(scan[0], (s0[0].astype(np.float64) + log[0] + log[1]), atol=1e-5)
