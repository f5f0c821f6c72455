"""CPU oracle for the ptychography hot path -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

A NumPy / scipy.fft (pocketfft) restatement of the reference's algorithm for the
path named by BASELINE.json:north_star.  Only tests/, __graft_entry__.smoke()
and bench.py's cpu_baseline leg may import this module; the product
(libtike.cufft) never does and fails loudly without its CUDA library.

Parity pinning: the reference stores NO golden vectors;
its only quantitative assertion is the adjoint identity of
tests/test_adjoint.py:47-59.  This oracle is pinned by (1) that identity on the
reference's own fixtures (tests/test_oracle.py) and (2) golden vectors produced
by the reference's compiled CUDA/cuFFT code (oracle/_ref, built from
/root/reference/src/cuda/*.cu where they lie) on a B200, committed under
tests/golden/ref_*.npz together with the script that made them
(tests/golden/make_golden.py).

What follows what (paths relative to /root/reference):
  gather_patch / weights .......... src/cuda/kernels.cu:19-46, 97-104
  fwd ............................. src/cuda/kernels.cu:95-107 + src/cuda/ptychofft.cu:60-73
  adj (object, flg 0) ............. src/cuda/ptychofft.cu:76-88 + src/cuda/kernels.cu:69-81
  adj_probe (flg 1) ............... src/cuda/ptychofft.cu:76-88 + src/cuda/kernels.cu:82-94
  register_translation_batch ...... src/libtike/cufft/ptycho.py:163-248
  line_search_sqr ................. src/libtike/cufft/ptycho.py:253-281
  cg_run .......................... src/libtike/cufft/ptycho.py:283-488
cuFFT (closed source, CUDA toolkit 12.9 -> cuFFT 11.4.1; call sites
src/cuda/ptychofft.cu:14,72,85) is restated by its published definition: the
unnormalised DFT with exp(-2 pi i ...) forward and exp(+2 pi i ...) inverse.
"""
import warnings

import numpy as np
import scipy.fft as sfft

F32 = np.float32
C64 = np.complex64


class float64_arithmetic(object):
    """Context manager: run the restatement in float64 / complex128 instead of the reference's
    float32 / complex64.  Used by the parity tests as the exact answer against which the rounding
    error of BOTH the reference's cuFFT path and the sm_100a kernels is measured where the
    reference's formula is ill-conditioned (Poisson gradient on noisy data: d * F / (|F|^2 + 1e-32)
    amplifies the fp32 FFT rounding error at weak-signal pixels that recorded a photon)."""

    def __enter__(self):
        global F32, C64
        self._saved = (F32, C64)
        F32, C64 = np.float64, np.complex128
        return self

    def __exit__(self, *exc):
        global F32, C64
        F32, C64 = self._saved


def _split_scan(scan_t):
    """modff split of one angle's scan positions (kernels.cu:27-28, 39).

    scan[..., 0] is the row (vertical) coordinate, scan[..., 1] the column; the
    kernel reads the pair as float2 with .x=row -> sy and .y=col -> sx.
    Returns integer origins (row R, col C), fractions (rho, gam) and the keep
    mask (positions whose integer part is negative are skipped).
    """
    r = scan_t[:, 0].astype(F32)
    c = scan_t[:, 1].astype(F32)
    R = np.trunc(r)
    C = np.trunc(c)
    rho = (r - R).astype(F32)
    gam = (c - C).astype(F32)
    keep = ~((R < 0) | (C < 0))
    return R.astype(np.int64), C.astype(np.int64), rho, gam, keep


def patch_origin(scan):
    """Integer work of the path: (R, C, keep) for scan [T,S,2] -- bit-exact parity target."""
    T = scan.shape[0]
    out = [_split_scan(scan[t]) for t in range(T)]
    R = np.stack([o[0] for o in out])
    C = np.stack([o[1] for o in out])
    keep = np.stack([o[4] for o in out])
    return R, C, keep


def _weights(rho, gam):
    """Bilinear weights in the reference's order (kernels.cu:97-100).

    w00 -> f[idx], w01 -> f[idx+1] (next column), w10 -> f[idx+N] (next row),
    w11 -> f[idx+1+N];  sxf = gam (column fraction), syf = rho (row fraction).
    """
    one = F32(1)
    w00 = ((one - gam) * (one - rho)).astype(F32)
    w01 = (gam * (one - rho)).astype(F32)
    w10 = ((one - gam) * rho).astype(F32)
    w11 = (gam * rho).astype(F32)
    return w00, w01, w10, w11


def gather_patches(psi_t, scan_t, nprb):
    """Bilinear patches [S,P,P] of one angle's object (kernels.cu:42-46, 97-104)."""
    R, C, rho, gam, keep = _split_scan(scan_t)
    S = scan_t.shape[0]
    out = np.zeros((S, nprb, nprb), dtype=C64)
    w00, w01, w10, w11 = _weights(rho, gam)
    for s in range(S):
        if not keep[s]:
            continue
        r0, c0 = R[s], C[s]
        a = psi_t[r0:r0 + nprb + 1, c0:c0 + nprb + 1]
        out[s] = (a[:-1, :-1] * w00[s] + a[:-1, 1:] * w01[s]
                  + a[1:, :-1] * w10[s] + a[1:, 1:] * w11[s])
    return out, keep


def fwd(psi, scan, probe, ndet, workers=-1):
    """g = FFT2(pad(c * probe * patch)), c = 1/ndet (ptychofft.cu:60-73, kernels.cu:48-66, 95-107).

    psi [T,nz,n] c64, scan [T,S,2] f32, probe [T,P,P] c64 -> [T,S,ndet,ndet] c64.
    """
    T, S = scan.shape[:2]
    P = probe.shape[-1]
    o = (ndet - P) // 2
    c = F32(1.0) / F32(ndet)
    g = np.zeros((T, S, ndet, ndet), dtype=C64)
    for t in range(T):
        patches, keep = gather_patches(psi[t], scan[t], P)
        near = np.zeros((S, ndet, ndet), dtype=C64)
        near[:, o:o + P, o:o + P] = (patches * probe[t][None]) * c
        near[~keep] = 0
        g[t] = sfft.fft2(near, axes=(-2, -1), workers=workers)
    return g


def _near_from_far(g_t, P, workers=-1):
    ndet = g_t.shape[-1]
    o = (ndet - P) // 2
    # cuFFT inverse is unnormalised: scipy's ifft2 with norm="forward" applies no 1/N^2
    near = sfft.ifft2(g_t, axes=(-2, -1), norm="forward", workers=workers)
    return near[:, o:o + P, o:o + P].astype(C64)


def adj(g, scan, probe, nz, n, workers=-1):
    """Object adjoint (flg 0): scatter-add of c*conj(probe)*IFFT2(g) (kernels.cu:69-81)."""
    T, S = scan.shape[:2]
    P = probe.shape[-1]
    ndet = g.shape[-1]
    c = F32(1.0) / F32(ndet)
    out = np.zeros((T, nz, n), dtype=C64)
    for t in range(T):
        near = _near_from_far(g[t], P, workers)
        tmp = (np.conj(probe[t])[None] * near * c).astype(C64)
        R, C, rho, gam, keep = _split_scan(scan[t])
        w00, w01, w10, w11 = _weights(rho, gam)
        acc = np.zeros((nz, n), dtype=np.complex128)  # order-free reference sum
        for s in range(S):
            if not keep[s]:
                continue
            r0, c0 = R[s], C[s]
            acc[r0:r0 + P, c0:c0 + P] += tmp[s] * w00[s]
            acc[r0:r0 + P, c0 + 1:c0 + P + 1] += tmp[s] * w01[s]
            acc[r0 + 1:r0 + P + 1, c0:c0 + P] += tmp[s] * w10[s]
            acc[r0 + 1:r0 + P + 1, c0 + 1:c0 + P + 1] += tmp[s] * w11[s]
        out[t] = acc.astype(C64)
    return out


def adj_probe(g, scan, psi, nprb, workers=-1):
    """Probe adjoint (flg 1): sum_s c * IFFT2(g) * conj(patch) (kernels.cu:82-94)."""
    T, S = scan.shape[:2]
    ndet = g.shape[-1]
    c = F32(1.0) / F32(ndet)
    out = np.zeros((T, nprb, nprb), dtype=C64)
    for t in range(T):
        near = _near_from_far(g[t], nprb, workers)
        patches, keep = gather_patches(psi[t], scan[t], nprb)
        prod = near * np.conj(patches) * c
        prod[~keep] = 0
        out[t] = prod.sum(axis=0, dtype=np.complex128).astype(C64)
    return out


# ----------------------------------------------------------------------------
# Position correction (ptycho.py:163-248), "next" row f1 of SURVEY.md section 8.
# ----------------------------------------------------------------------------

def _upsampled_dft_batch(data, ups, upsample_factor=1, axis_offsets=None):
    """ptycho.py:163-190 -- matrix-multiply DFT of an `ups` x `ups` window of the upsampled inverse
    transform, complex128 (float64 kernels times complex64 data).

    data [S,N,N]; axis_offsets [S,2] (row, col).  rec[i, jr, jc] =
      sum_{r,c} exp(-2 pi i (jr - off_r) f[r]) exp(-2 pi i (jc - off_c) f[c]) data[i, r, c],
    f = fftfreq(N, upsample_factor).  Both kernels use data.shape[2] (the reference's square-frame
    assumption, ptycho.py:182, 185).
    """
    im2pi = 1j * 2 * np.pi
    ups = int(ups)
    S = data.shape[0]
    freq = np.fft.fftfreq(data.shape[2], upsample_factor)
    kernel = (np.tile(np.arange(ups), (S, 1)) - axis_offsets[:, 1:2])[:, :, None] * freq
    kernel = np.exp(-im2pi * kernel)
    tdata = np.einsum('ijk,ipk->ijp', kernel, data)
    kernel = (np.tile(np.arange(ups), (S, 1)) - axis_offsets[:, 0:1])[:, :, None] * freq
    kernel = np.exp(-im2pi * kernel)
    return np.einsum('ijk,ipk->ijp', kernel, tdata)


def register_translation_batch(src_image, target_image, upsample_factor=1, space="real"):
    """ptycho.py:192-248 -- batched phase correlation with an upsampled matrix DFT around the
    whole-pixel peak.  Returns float64 shifts [S,2] (row, col).  Statement by statement, including
    the trailing `shape[dim] == 1` loop, which for a batch of ONE image zeroes that image's shifts
    (it indexes the batch axis; ptycho.py:243-245)."""
    if space.lower() == 'fourier':
        src_freq = src_image
        target_freq = target_image
    elif space.lower() == 'real':
        src_freq = sfft.fft2(src_image, axes=(-2, -1))
        target_freq = sfft.fft2(target_image, axes=(-2, -1))
    shape = src_freq.shape
    image_product = src_freq * target_freq.conj()
    cross_correlation = sfft.ifft2(image_product, axes=(-2, -1))
    A = np.abs(cross_correlation)
    maxima = A.reshape(A.shape[0], -1).argmax(1)
    maxima = np.column_stack(np.unravel_index(maxima, A[0, :, :].shape))
    midpoints = np.array([np.fix(axis_size / 2) for axis_size in shape[1:]])
    shifts = np.array(maxima, dtype=np.float64)
    ids = np.where(shifts[:, 0] > midpoints[0])
    shifts[ids[0], 0] -= shape[1]
    ids = np.where(shifts[:, 1] > midpoints[1])
    shifts[ids[0], 1] -= shape[2]
    if upsample_factor > 1:
        shifts = np.round(shifts * upsample_factor) / upsample_factor
        upsampled_region_size = np.ceil(upsample_factor * 1.5)
        dftshift = np.fix(upsampled_region_size / 2.0)
        normalization = (src_freq[0].size * upsample_factor ** 2)
        sample_region_offset = dftshift - shifts * upsample_factor
        cross_correlation = _upsampled_dft_batch(image_product.conj(), upsampled_region_size,
                                                 upsample_factor, sample_region_offset).conj()
        cross_correlation /= normalization
        A = np.abs(cross_correlation)
        maxima = A.reshape(A.shape[0], -1).argmax(1)
        maxima = np.column_stack(np.unravel_index(maxima, A[0, :, :].shape))
        maxima = np.array(maxima, dtype=np.float64) - dftshift
        shifts = shifts + maxima / upsample_factor
    for dim in range(src_freq.ndim):
        if shape[dim] == 1:
            shifts[dim] = 0
    return shifts


# ----------------------------------------------------------------------------
# CG solver (ptycho.py:250-488), float32 arithmetic like CuPy's.
# ----------------------------------------------------------------------------

def line_search_sqr(f, p1, p2, p3, step_length=1, step_shrink=0.5, forced=None):
    """ptycho.py:253-281 -- backtracking on f(p1 + g^2 p2 + g p3), m = 0.

    `forced` (diagnostics): a list of reference decisions consumed in call order; when given the
    search result is overridden so that near-tie decisions (fp32 summation noise) cannot fork the
    trajectory of a long parity run.
    """
    assert 0 < step_shrink < 1
    if forced:
        return forced.pop(0)
    m = 0
    fp1 = f(p1)
    while f(p1 + F32(step_length ** 2) * p2 + F32(step_length) * p3) > fp1 + step_shrink * m:
        if step_length < 1e-32:
            warnings.warn("Line search failed for conjugate gradient.")
            return 0
        step_length *= step_shrink
    return step_length


def cg_run(data, psi, scan, probe, piter, model="gaussian", recover_prb=False,
           ndet=None, verbose=False, history=None, forced_steps=None, position_correction=False,
           shift_log=None):
    """Statement-by-statement restatement of CGPtychoSolver.run (ptycho.py:283-488).

    Deviations, all deliberate and documented in DESIGN.md:
      * Q1: the Poisson object branch reads `fpsi` before assignment
        (ptycho.py:357-363); the evident missing line `fpsi = fwd(...)` is added.
      * Q5: the position-correction block (ptycho.py:398-403), unconditional in the reference,
        is behind `position_correction` (off = the primary parity configuration).  It mutates
        the caller's `scan` (angle 0 only), like the reference.
      * Q7: the dead `sfpsi` recompute (ptycho.py:476-480) is skipped.
    data [T,S,N,N] f32, psi [T,nz,n] c64, scan [T,S,2] f32, probe [T,M,P,P] c64.
    """
    assert probe.ndim == 4
    T, S = scan.shape[:2]
    nz, n = psi.shape[1:]
    M, P = probe.shape[1], probe.shape[-1]
    ndet = data.shape[-1] if ndet is None else ndet
    data = data.astype(F32)
    psi = psi.astype(C64).copy()
    probe = probe.astype(C64).copy()

    def _fwd(x, p):
        return fwd(x, scan, np.ascontiguousarray(p), ndet)

    def minf(fpsi):
        if model == "gaussian":
            return np.linalg.norm(np.sqrt(np.abs(fpsi)) - np.sqrt(data)) ** 2
        return np.sum(np.abs(fpsi) - data * np.log(np.abs(fpsi) + F32(1e-32)))

    dprb = dpsi = gradprb0 = gradpsi0 = 0
    gammaprb = 0
    for i in range(piter):
        absfpsi = data * 0
        for k in range(M):
            absfpsi += np.abs(_fwd(psi, probe[:, k])) ** 2
        a = np.sum(np.sqrt(absfpsi * data))
        b = np.sum(absfpsi)
        probe *= (a / b)
        absfpsi *= (a / b) ** 2
        gradpsi = np.zeros((T, nz, n), dtype=C64)
        for k in range(M):
            if model == "gaussian":
                fpsi = _fwd(psi, probe[:, k]) * (b / a)
                res = fpsi - np.sqrt(data) * fpsi / (np.sqrt(absfpsi) + F32(1e-32))
            else:
                fpsi = _fwd(psi, probe[:, k])  # Q1: line missing in the reference
                res = fpsi - data * fpsi / (absfpsi + F32(1e-32))
            gradpsi += adj(res.astype(C64), scan, np.ascontiguousarray(probe[:, k]), nz, n) \
                / (np.max(np.abs(probe[:, k])) ** 2)
        if i == 0:
            dpsi = -gradpsi
        else:
            dpsi = -gradpsi + (np.linalg.norm(gradpsi) ** 2
                               / (np.sum(np.conj(dpsi) * (gradpsi - gradpsi0))) * dpsi)
        dpsi = dpsi.astype(C64)
        gradpsi0 = gradpsi
        p1 = data * 0
        p2 = data * 0
        p3 = data * 0
        for k in range(M):
            tmp1 = _fwd(psi, probe[:, k])
            tmp2 = _fwd(dpsi, probe[:, k])
            p1 += np.abs(tmp1) ** 2
            p2 += np.abs(tmp2) ** 2
            p3 += 2 * (tmp1.real * tmp2.real + tmp1.imag * tmp2.imag)
        gammapsi = 0.5 * line_search_sqr(minf, p1, p2, p3, forced=forced_steps)
        if position_correction and i > 0:  # ptycho.py:398-403
            ones = probe[:, 0] * 0 + 1
            tmp1 = _fwd(psi, ones)[0]
            tmp2 = _fwd((psi + F32(gammapsi) * dpsi).astype(C64), ones)[0]
            shifts = register_translation_batch(tmp1, tmp2, upsample_factor=100, space='fourier')
            if shift_log is not None:
                shift_log.append(shifts.copy())
            scan[0, :] += shifts
        psi = (psi + F32(gammapsi) * dpsi).astype(C64)

        if recover_prb:
            if i == 0:
                gradprb = probe * 0
                gradprb0 = probe * 0
                dprb = probe * 0
            for m in range(M):
                fprb = _fwd(psi, probe[:, m])
                absfprb = data * 0
                for k in range(M):
                    absfprb += np.abs(_fwd(psi, probe[:, k])) ** 2
                if model == "gaussian":
                    res = fprb - np.sqrt(data) * fprb / (np.sqrt(absfprb) + F32(1e-32))
                    gradprb[:, m] = adj_probe(res.astype(C64), scan, psi, P) \
                        / np.max(np.abs(psi)) ** 2 / S * M
                else:
                    res = fprb - data * fprb / (absfprb + F32(1e-32))
                    gradprb[:, m] = adj_probe(res.astype(C64), scan, psi, P) \
                        / np.max(np.abs(psi)) ** 2 / S
                if i == 0:
                    dprb[:, m] = -gradprb[:, m]
                else:
                    dprb[:, m] = -gradprb[:, m] + (
                        np.linalg.norm(gradprb[:, m]) ** 2
                        / (np.sum(np.conj(dprb[:, m]) * (gradprb[:, m] - gradprb0[:, m])))
                        * dprb[:, m])
                gradprb0[:, m] = gradprb[:, m]
                p1 = data * 0
                for k in range(M):
                    p1 += np.abs(_fwd(psi, probe[:, k])) ** 2
                tmp1 = _fwd(psi, probe[:, m])
                tmp2 = _fwd(psi, dprb[:, m])
                p2 = np.abs(tmp2) ** 2
                p3 = 2 * (tmp1.real * tmp2.real + tmp1.imag * tmp2.imag)
                gammaprb = 0.5 * line_search_sqr(minf, p1, p2, p3, step_length=1, forced=forced_steps)
                probe[:, m] = probe[:, m] + F32(gammaprb) * dprb[:, m]
        if history is not None:
            history.append((i, float(gammapsi), float(gammaprb), float(minf(absfpsi))))
        if verbose and i % 32 == 0:
            print("%4d, %.3e, %.3e, %.7e" % (i, gammapsi, gammaprb, minf(absfpsi)))
    return {"psi": psi, "probe": probe}
