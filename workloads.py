"""Seeded synthetic inputs for the configurations of BASELINE.json.

Shared by tests/, bench.py and tests/golden/make_golden.py.  Pure NumPy/SciPy; no
GPU, no oracle, no reference access (the reference's input fixtures were
re-packed into tests/golden/model.npz by tests/golden/make_inputs.py).

Array conventions (reference tests/test_adjoint.py:24-39):
  psi   [ntheta, nz, n]            complex64
  scan  [ntheta, nscan, 2]         float32, (row, col) pairs
  probe [ntheta, nmodes, nprb, nprb] complex64
"""
import os

import numpy as np
from scipy import ndimage

_HERE = os.path.dirname(os.path.abspath(__file__))
_MODEL = os.path.join(_HERE, "tests", "golden", "model.npz")


def model():
    """The reference's input fixtures (tests/model/*.tiff, coords.npy), re-packed."""
    return np.load(_MODEL)


def fixture_probe(nmodes=1, multi=False):
    m = model()
    if multi:
        p = m["probes_amp"][:nmodes] * np.exp(1j * m["probes_ang"][:nmodes])
    else:
        assert nmodes == 1
        p = (m["prbamp"] * np.exp(1j * m["prbang"]))[None]
    return p.astype(np.complex64)[None]  # [1, M, 128, 128]


def fixture_object():
    m = model()
    return (m["initpsiamp"] * np.exp(1j * m["initpsiang"])).astype(np.complex64)[None]


def fixture_scan(nscan, stride=1):
    """scan[0,:,0] = coords[1] (vertical), scan[0,:,1] = coords[0] (tests/test_adjoint.py:28-33)."""
    c = np.moveaxis(model()["coords"], 0, 1)[: nscan * stride: stride]
    scan = np.ones((1, nscan, 2), dtype=np.float32)
    scan[0, :, 0] = c[:, 1]
    scan[0, :, 1] = c[:, 0]
    return scan


def c1_adjoint(nscan=100):
    """C1: tests/test_adjoint.py:15-39 -- n=600, nz=276, nscan=100, nprb=ndet=128, 1 mode."""
    return dict(psi=fixture_object(), scan=fixture_scan(nscan), probe=fixture_probe(1),
                ndet=128, nprb=128, nz=276, n=600, nscan=nscan, nmodes=1)


def c3_modes(nmodes=5, nscan=1100):
    """C3: tests/test_modes.py:18-52 with all five stored modes; every 5th coordinate."""
    probe = fixture_probe(nmodes, multi=True)
    init = probe.copy()
    for k in range(nmodes):
        init[:, k] /= np.max(np.abs(init[:, k]))
    return dict(psi=fixture_object(), scan=fixture_scan(nscan, 5), probe=probe,
                probe_init=init, ndet=128, nprb=128, nz=276, n=600, nscan=nscan, nmodes=nmodes)


def synth_object(nz, n, seed):
    """amp in [0.8,1], phase in [-0.5,0.5], both low-pass filtered (sigma 4 px)."""
    rng = np.random.default_rng(seed)
    amp = ndimage.gaussian_filter(rng.random((nz, n)), 4.0, mode="wrap")
    ph = ndimage.gaussian_filter(rng.random((nz, n)), 4.0, mode="wrap")

    def unit(a):
        return (a - a.min()) / (a.max() - a.min())

    amp = 0.8 + 0.2 * unit(amp)
    ph = unit(ph) - 0.5
    return (amp * np.exp(1j * ph)).astype(np.complex64)


def synth_probe(nprb, nmodes=1):
    """Fixture probe(s) zoomed (order-1 spline on re/im) to nprb x nprb."""
    p = fixture_probe(nmodes, multi=(nmodes > 1))[0]
    if nprb == p.shape[-1]:
        return p[None].copy()
    z = nprb / p.shape[-1]
    out = np.stack([ndimage.zoom(q.real, z, order=1) + 1j * ndimage.zoom(q.imag, z, order=1)
                    for q in p])
    return out.astype(np.complex64)[None]


def raster_scan(nz, n, nprb, nside, seed, jitter=3.0):
    """nside x nside raster + U(-jitter, jitter), clipped to [0, dim-nprb-1]; always fractional."""
    rng = np.random.default_rng(seed)
    step_r = (nz - nprb - 2 - 2 * jitter) / max(nside - 1, 1)
    step_c = (n - nprb - 2 - 2 * jitter) / max(nside - 1, 1)
    rr, cc = np.meshgrid(np.arange(nside) * step_r + jitter, np.arange(nside) * step_c + jitter,
                         indexing="ij")
    r = rr.ravel() + rng.uniform(-jitter, jitter, nside * nside)
    c = cc.ravel() + rng.uniform(-jitter, jitter, nside * nside)
    r = np.clip(r, 0.0, nz - nprb - 1.001) + 1e-3 * rng.random(nside * nside)
    c = np.clip(c, 0.0, n - nprb - 1.001) + 1e-3 * rng.random(nside * nside)
    return np.stack([r, c], axis=-1).astype(np.float32)


def synth_angles(ntheta, nz, n, ndet, nprb, nside, nmodes=1, seed0=0):
    """ntheta independent angles of a (nz x n) object, nside^2 scan positions each (C2, C4, C5)."""
    psi = np.stack([synth_object(nz, n, seed0 + t) for t in range(ntheta)])
    scan = np.stack([raster_scan(nz, n, nprb, nside, 1000 + seed0 + t) for t in range(ntheta)])
    probe = np.repeat(synth_probe(nprb, nmodes), ntheta, axis=0)
    return dict(psi=psi, scan=scan, probe=probe, ndet=ndet, nprb=nprb, nz=nz, n=n,
                nscan=nside * nside, nmodes=nmodes)


def c2_single_angle(ntheta=1, nside=32):
    """C2: 512x512 object, 128x128 detector, 1 mode, 1024 scan positions per angle."""
    return synth_angles(ntheta, 512, 512, 128, 128, nside, 1)


def c4_catalyst(ntheta, nside=32, nmodes=1):
    """C4 shard: 1024x1024 object slices, 256x256 detector, 1024 positions per angle."""
    return synth_angles(ntheta, 1024, 1024, 256, 256, nside, nmodes)


def c5_sweep(ndet, ntheta=1, nside=32):
    """C5: detector-size sweep, object (4 ndet)^2, subpixel scan."""
    return synth_angles(ntheta, 4 * ndet, 4 * ndet, ndet, ndet, nside, 1)
