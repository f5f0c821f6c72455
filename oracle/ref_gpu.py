"""GPU oracle: the reference's own compiled CUDA/cuFFT code + a torch restatement of its Python layer.

TEST INFRASTRUCTURE, NOT PRODUCT CODE.  Only tests/, tests/golden/make_golden.py,
__graft_entry__.smoke() and bench.py's reference arm may import this module.

* `RefPtychoFFT` drives oracle/_ref/libptychofft_ref.so, i.e. the UNMODIFIED
  /root/reference/src/cuda/ptychofft.cu + kernels.cu compiled where they lie by
  oracle/Makefile (cuFFT 11.4 from CUDA 12.9), through raw device pointers, as
  src/libtike/cufft/ptycho.py:80-123 does with CuPy's `.data.ptr`.
* `RefCGPtychoSolver.run` restates src/libtike/cufft/ptycho.py:283-488 statement by
  statement with `cp.` -> `torch.` (CuPy is not installed in this image, so the
  reference's Python layer cannot be imported -- SURVEY.md section 8c).  Deviations:
    Q1  the missing `fpsi = self.fwd(...)` line of the Poisson object branch
        (ptycho.py:357-363) is added (the reference raises UnboundLocalError);
    Q5  the position-correction block (ptycho.py:398-403), unconditional in the reference, is
        behind `position_correction` (False = the primary parity configuration);
        `register_translation_batch` / `_upsampled_dft_batch` (ptycho.py:163-248) are restated
        with cp -> torch (cuFFT + complex128 einsum, like CuPy);
    Q7  the dead `sfpsi` recompute (ptycho.py:476-480) is skipped;
    Q10 `probe[:, k]` views are made contiguous before their pointer is taken (the
        reference passes the view's base pointer, which is only right for ptheta = 1).
"""
import ctypes
import os
import warnings

import numpy as np
import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = os.path.join(_HERE, "_ref", "libptychofft_ref.so")


def available():
    return os.path.exists(_LIB)


_lib = None


def _load():
    global _lib
    if _lib is None:
        lib = ctypes.CDLL(_LIB)
        sz, vp = ctypes.c_size_t, ctypes.c_void_p
        lib.ref_create.restype = vp
        lib.ref_create.argtypes = [sz] * 6
        lib.ref_fwd.argtypes = [vp, sz, sz, sz, sz]
        lib.ref_adj.argtypes = [vp, sz, sz, sz, sz, ctypes.c_int]
        lib.ref_free.argtypes = [vp]
        lib.ref_destroy.argtypes = [vp]
        lib.ref_last_cuda_error.restype = ctypes.c_int
        lib.ref_device_sync.restype = ctypes.c_int
        _lib = lib
    return _lib


class RefPtychoFFT(object):
    """PtychoCuFFT of the reference (ptycho.py:34-129) on torch tensors, legacy default stream."""

    def __init__(self, nscan, probe_shape, detector_shape, ntheta, nz, n):
        self._lib = _load()
        self.ptheta, self.nz, self.n = ntheta, nz, n
        self.nscan, self.ndet, self.nprb = nscan, detector_shape, probe_shape
        self._h = ctypes.c_void_p(self._lib.ref_create(ntheta, nz, n, nscan, detector_shape,
                                                       probe_shape))

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.free()

    def free(self):
        if self._h:
            self._lib.ref_free(self._h)

    def __del__(self):
        try:
            if self._h:
                self._lib.ref_destroy(self._h)
                self._h = None
        except Exception:
            pass

    def _check(self):
        err = self._lib.ref_last_cuda_error()
        if err:
            raise RuntimeError("reference CUDA error %d" % err)

    def fwd(self, psi, scan, probe):
        assert psi.dtype == torch.complex64 and scan.dtype == torch.float32
        assert probe.dtype == torch.complex64
        psi, scan, probe = psi.contiguous(), scan.contiguous(), probe.contiguous()
        farplane = torch.zeros([self.ptheta, self.nscan, self.ndet, self.ndet],
                               dtype=torch.complex64, device="cuda")
        self._lib.ref_fwd(self._h, farplane.data_ptr(), psi.data_ptr(), scan.data_ptr(),
                          probe.data_ptr())
        self._check()
        return farplane

    def adj(self, farplane, scan, probe):
        farplane, scan, probe = farplane.contiguous(), scan.contiguous(), probe.contiguous()
        psi = torch.zeros([self.ptheta, self.nz, self.n], dtype=torch.complex64, device="cuda")
        self._lib.ref_adj(self._h, psi.data_ptr(), farplane.data_ptr(), scan.data_ptr(),
                          probe.data_ptr(), 0)
        self._check()
        return psi

    def adj_probe(self, farplane, scan, psi):
        farplane, scan, psi = farplane.contiguous(), scan.contiguous(), psi.contiguous()
        probe = torch.zeros([self.ptheta, self.nprb, self.nprb], dtype=torch.complex64,
                            device="cuda")
        self._lib.ref_adj(self._h, psi.data_ptr(), farplane.data_ptr(), scan.data_ptr(),
                          probe.data_ptr(), 1)
        self._check()
        return probe

    # host-array helpers exactly as the reference runs them: one angle per call,
    # pageable H2D (cp.array) and a blocking D2H (.get()) per angle (ptycho.py:70-78)
    def _batch(self, function, output, *inputs):
        for ids in range(0, inputs[0].shape[0]):
            inputs_gpu = [torch.from_numpy(np.ascontiguousarray(x[ids:ids + 1])).cuda()
                          for x in inputs]
            output[ids] = function(*inputs_gpu).cpu().numpy()
        return output

    def fwd_ptycho_batch(self, psi, scan, probe):
        data = np.zeros([scan.shape[0], self.nscan, self.ndet, self.ndet], dtype="complex64")
        return self._batch(self.fwd, data, psi, scan, probe)

    def adj_ptycho_batch(self, farplane, scan, probe):
        psi = np.zeros([scan.shape[0], self.nz, self.n], dtype="complex64")
        return self._batch(self.adj, psi, farplane, scan, probe)

    def adj_ptycho_batch_prb(self, farplane, scan, psi):
        probe = np.zeros([scan.shape[0], self.nprb, self.nprb], dtype="complex64")
        return self._batch(self.adj_probe, probe, farplane, scan, psi)


def _upsampled_dft_batch(data, ups, upsample_factor=1, axis_offsets=None):
    """ptycho.py:163-190 with cp -> torch; the kernels are float64 -> complex128 and CuPy's einsum
    promotes the complex64 data to complex128, which torch needs spelled out."""
    ups = int(ups)
    S = data.shape[0]
    dev = data.device
    freq = torch.fft.fftfreq(data.shape[2], upsample_factor, dtype=torch.float64, device=dev)
    tdata = data.to(torch.complex128)
    ar = torch.arange(ups, dtype=torch.float64, device=dev).repeat(S, 1)
    kernel = (ar - axis_offsets[:, 1:2])[:, :, None] * freq
    kernel = torch.exp(-2j * np.pi * kernel)
    tdata = torch.einsum('ijk,ipk->ijp', kernel, tdata)
    kernel = (ar - axis_offsets[:, 0:1])[:, :, None] * freq
    kernel = torch.exp(-2j * np.pi * kernel)
    return torch.einsum('ijk,ipk->ijp', kernel, tdata)


def register_translation_batch(src_image, target_image, upsample_factor=1, space="real"):
    """ptycho.py:192-248 with cp -> torch.  Returns float64 shifts [S,2] on the device."""
    if space.lower() == 'fourier':
        src_freq = src_image
        target_freq = target_image
    elif space.lower() == 'real':
        src_freq = torch.fft.fft2(src_image)
        target_freq = torch.fft.fft2(target_image)
    shape = src_freq.shape
    image_product = src_freq * target_freq.conj()
    cross_correlation = torch.fft.ifft2(image_product)
    A = torch.abs(cross_correlation)
    maxima = A.reshape(A.shape[0], -1).argmax(1)
    maxima = torch.stack((maxima // shape[2], maxima % shape[2]), dim=1)
    midpoints = [np.fix(axis_size / 2) for axis_size in shape[1:]]
    shifts = maxima.to(torch.float64)
    shifts[:, 0] = torch.where(shifts[:, 0] > midpoints[0], shifts[:, 0] - shape[1], shifts[:, 0])
    shifts[:, 1] = torch.where(shifts[:, 1] > midpoints[1], shifts[:, 1] - shape[2], shifts[:, 1])
    if upsample_factor > 1:
        shifts = torch.round(shifts * upsample_factor) / upsample_factor
        upsampled_region_size = np.ceil(upsample_factor * 1.5)
        dftshift = np.fix(upsampled_region_size / 2.0)
        normalization = (src_freq[0].numel() * upsample_factor ** 2)
        sample_region_offset = dftshift - shifts * upsample_factor
        cross_correlation = _upsampled_dft_batch(image_product.conj(), upsampled_region_size,
                                                 upsample_factor, sample_region_offset).conj()
        cross_correlation = cross_correlation / normalization
        A = torch.abs(cross_correlation)
        maxima = A.reshape(A.shape[0], -1).argmax(1)
        ups = A.shape[2]
        maxima = torch.stack((maxima // ups, maxima % ups), dim=1)
        maxima = maxima.to(torch.float64) - dftshift
        shifts = shifts + maxima / upsample_factor
    for dim in range(src_freq.ndim):
        if shape[dim] == 1:
            shifts[dim] = 0
    return shifts


class RefCGPtychoSolver(RefPtychoFFT):
    """torch restatement of CGPtychoSolver (ptycho.py:250-488)."""

    position_correction = False
    cdtype = torch.complex64
    rdtype = torch.float32
    #: optional list of raw line-search results, consumed in call order, that REPLACE the decisions
    #: (used to run a second arithmetic -- e.g. the float64 referee -- along the same trajectory)
    forced_steps = None

    @staticmethod
    def line_search_sqr(f, p1, p2, p3, step_length=1, step_shrink=0.5, trials=None, forced=None):
        """ptycho.py:253-281; `trials` (diagnostics) collects (f(p1), [(step, f(step)), ...], result)."""
        assert step_shrink > 0 and step_shrink < 1
        if forced is not None:
            if trials is not None:
                trials.append((float("nan"), [], forced))
            return forced
        m = 0
        fp1 = f(p1)
        seen = []
        while True:
            fs = f(p1 + step_length ** 2 * p2 + step_length * p3)
            seen.append((step_length, float(fs)))
            if not (fs > fp1 + step_shrink * m):
                break
            if step_length < 1e-32:
                warnings.warn("Line search failed for conjugate gradient.")
                step_length = 0
                break
            step_length *= step_shrink
        if trials is not None:
            trials.append((float(fp1), seen, step_length))
        return step_length

    def run(self, data, psi, scan, probe, piter, model="gaussian", recover_prb=False,
            ortho_prb=False, history=None, verbose=True):
        assert probe.ndim == 4, "probe needs 4 dimensions, not %d" % probe.ndim

        def minf(fpsi):
            if model == "gaussian":
                f = torch.linalg.norm(torch.sqrt(torch.abs(fpsi)) - torch.sqrt(data)) ** 2
            elif model == "poisson":
                f = torch.sum(torch.abs(fpsi) - data * torch.log(torch.abs(fpsi) + 1e-32))
            return f

        dprb = 0
        dpsi = 0
        gradprb0 = 0
        gradpsi0 = 0
        if verbose:
            print("# congujate gradient parameters\n"
                  "iteration, step size object, step size probe, function min")
        gammaprb = 0
        trials = []
        forced = list(self.forced_steps) if self.forced_steps else None

        def next_forced():
            return forced.pop(0) if forced else None
        for i in range(piter):
            absfpsi = data * 0
            for k in range(probe.shape[1]):
                tmp = self.fwd(psi, scan, probe[:, k])
                absfpsi += torch.abs(tmp) ** 2
            a = torch.sum(torch.sqrt(absfpsi * data))
            b = torch.sum(absfpsi)
            probe *= (a / b)
            absfpsi *= (a / b) ** 2
            gradpsi = torch.zeros([self.ptheta, self.nz, self.n], dtype=self.cdtype,
                                  device="cuda")
            if model == "gaussian":
                for k in range(probe.shape[1]):
                    fpsi = self.fwd(psi, scan, probe[:, k]) * (b / a)
                    gradpsi += self.adj(
                        fpsi - torch.sqrt(data) * fpsi / (torch.sqrt(absfpsi) + 1e-32),
                        scan, probe[:, k]) / (torch.max(torch.abs(probe[:, k])) ** 2)
            elif model == "poisson":
                for k in range(probe.shape[1]):
                    fpsi = self.fwd(psi, scan, probe[:, k])  # Q1
                    gradpsi += self.adj(
                        fpsi - data * fpsi / (absfpsi + 1e-32),
                        scan, probe[:, k]) / (torch.max(torch.abs(probe[:, k])) ** 2)
            if i == 0:
                dpsi = -gradpsi
            else:
                dpsi = -gradpsi + (
                    torch.linalg.norm(gradpsi) ** 2 /
                    (torch.sum(torch.conj(dpsi) * (gradpsi - gradpsi0))) * dpsi)
            gradpsi0 = gradpsi
            p1 = data * 0
            p2 = data * 0
            p3 = data * 0
            for k in range(probe.shape[1]):
                tmp1 = self.fwd(psi, scan, probe[:, k])
                tmp2 = self.fwd(dpsi, scan, probe[:, k])
                p1 += torch.abs(tmp1) ** 2
                p2 += torch.abs(tmp2) ** 2
                p3 += 2 * (tmp1.real * tmp2.real + tmp1.imag * tmp2.imag)
            gammapsi = 0.5 * self.line_search_sqr(minf, p1, p2, p3, trials=trials, forced=next_forced())
            if self.position_correction and i > 0:  # ptycho.py:398-403
                tmp1 = self.fwd(psi, scan, probe[:, 0] * 0 + 1)[0]
                tmp2 = self.fwd(psi + gammapsi * dpsi, scan, probe[:, 0] * 0 + 1)[0]
                shifts = register_translation_batch(tmp1, tmp2, upsample_factor=100, space='fourier')
                if getattr(self, "shift_log", None) is not None:
                    self.shift_log.append(shifts.cpu().numpy())
                scan[0, :] += shifts
            psi = psi + gammapsi * dpsi

            if recover_prb:
                if i == 0:
                    gradprb = probe * 0
                    gradprb0 = probe * 0
                    dprb = probe * 0
                for m in range(0, probe.shape[1]):
                    fprb = self.fwd(psi, scan, probe[:, m])
                    absfprb = data * 0
                    for k in range(probe.shape[1]):
                        tmp = self.fwd(psi, scan, probe[:, k])
                        absfprb += torch.abs(tmp) ** 2
                    if model == "gaussian":
                        gradprb[:, m] = self.adj_probe(
                            fprb - torch.sqrt(data) * fprb / (torch.sqrt(absfprb) + 1e-32),
                            scan, psi) / torch.max(torch.abs(psi)) ** 2 / self.nscan * probe.shape[1]
                    elif model == "poisson":
                        gradprb[:, m] = self.adj_probe(
                            fprb - data * fprb / (absfprb + 1e-32),
                            scan, psi) / torch.max(torch.abs(psi)) ** 2 / self.nscan
                    if i == 0:
                        dprb[:, m] = -gradprb[:, m]
                    else:
                        dprb[:, m] = -gradprb[:, m] + (
                            torch.linalg.norm(gradprb[:, m]) ** 2 /
                            (torch.sum(torch.conj(dprb[:, m]) * (gradprb[:, m] - gradprb0[:, m])))
                            * dprb[:, m])
                    gradprb0[:, m] = gradprb[:, m]
                    p1 = data * 0
                    p2 = data * 0
                    p3 = data * 0
                    for k in range(probe.shape[1]):
                        tmp1 = self.fwd(psi, scan, probe[:, k])
                        p1 += torch.abs(tmp1) ** 2
                    tmp1 = self.fwd(psi, scan, probe[:, m])
                    tmp2 = self.fwd(psi, scan, dprb[:, m])
                    p2 = torch.abs(tmp2) ** 2
                    p3 = 2 * (tmp1.real * tmp2.real + tmp1.imag * tmp2.imag)
                    gammaprb = 0.5 * self.line_search_sqr(minf, p1, p2, p3, step_length=1,
                                                          trials=trials, forced=next_forced())
                    probe[:, m] = probe[:, m] + gammaprb * dprb[:, m]
            if history is not None:
                history.append((i, float(gammapsi), float(gammaprb), float(minf(absfpsi))))
            if np.mod(i, 32) == 0 and verbose:
                print("%4d, %.3e, %.3e, %.7e" % (i, gammapsi, gammaprb, minf(absfpsi)))
        self.last_trials = trials
        return {"psi": psi, "probe": probe}

    def run_batch(self, data, psi, scan, probe, **kwargs):
        """ptycho.py:135-162"""
        assert probe.ndim == 4
        psi = psi.copy()
        probe = probe.copy()
        for k in range(0, scan.shape[0] // self.ptheta):
            ids = np.arange(k * self.ptheta, (k + 1) * self.ptheta)
            psi_gpu = torch.from_numpy(psi[ids]).cuda().to(self.cdtype)
            scan_gpu = torch.from_numpy(scan[ids]).cuda()  # positions stay float32 data in every arithmetic
            prb_gpu = torch.from_numpy(probe[ids]).cuda().to(self.cdtype)
            data_gpu = torch.from_numpy(data[ids]).cuda().to(self.rdtype)
            result = self.run(data_gpu, psi_gpu, scan_gpu, prb_gpu, **kwargs)
            psi[ids], probe[ids] = result["psi"].cpu().numpy(), result["probe"].cpu().numpy()
        return {"psi": psi, "probe": probe}


class F64PtychoOps(object):
    """The operators of kernels.cu:19-107 + ptychofft.cu:60-88 restated in torch with float64 /
    complex128 arithmetic on the GPU (cuFFT in double precision): the EXACT answer, to ~1e-15, that
    both fp32 implementations -- the reference's cuFFT path and the sm_100a kernels -- are measured
    against where the reference's formulas amplify rounding (tests/, tests/tools/).  Same statements
    as oracle/numpy_ptycho.py under `float64_arithmetic()`, which pins it (tests/test_gpu_f64_oracle.py);
    positions must lie in the reference's valid domain (no zero extension here)."""

    def __init__(self, nscan, probe_shape, detector_shape, ntheta, nz, n):
        self.ptheta, self.nz, self.n = ntheta, nz, n
        self.nscan, self.ndet, self.nprb = nscan, detector_shape, probe_shape

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        pass

    def free(self):
        pass

    @staticmethod
    def _split(scan_t):
        r, c = scan_t[:, 0].double(), scan_t[:, 1].double()
        R, C = torch.trunc(r), torch.trunc(c)
        keep = ~((R < 0) | (C < 0))
        return (torch.where(keep, R, 0 * R).long(), torch.where(keep, C, 0 * C).long(), r - R, c - C, keep)

    def _patches(self, psi_t, scan_t):
        P = self.nprb
        R, C, rho, gam, keep = self._split(scan_t)
        ar = torch.arange(P + 1, device=psi_t.device)
        a = psi_t[(R[:, None] + ar)[:, :, None], (C[:, None] + ar)[:, None, :]]
        w = lambda x: x[:, None, None]  # noqa: E731
        out = (a[:, :-1, :-1] * w((1 - gam) * (1 - rho)) + a[:, :-1, 1:] * w(gam * (1 - rho))
               + a[:, 1:, :-1] * w((1 - gam) * rho) + a[:, 1:, 1:] * w(gam * rho))
        out[~keep] = 0
        return out, keep

    def fwd(self, psi, scan, probe):
        psi, probe = psi.to(torch.complex128), probe.to(torch.complex128)
        N, P = self.ndet, self.nprb
        o = (N - P) // 2
        g = torch.zeros((self.ptheta, self.nscan, N, N), dtype=torch.complex128, device=psi.device)
        for t in range(self.ptheta):
            patches, _ = self._patches(psi[t], scan[t])
            near = torch.zeros((self.nscan, N, N), dtype=torch.complex128, device=psi.device)
            near[:, o:o + P, o:o + P] = patches * probe[t][None] / N
            g[t] = torch.fft.fft2(near)
        return g

    def _near(self, g_t):
        N, P = self.ndet, self.nprb
        o = (N - P) // 2
        return torch.fft.ifft2(g_t.to(torch.complex128), norm="forward")[:, o:o + P, o:o + P]

    def adj(self, farplane, scan, probe):
        probe = probe.to(torch.complex128)
        P = self.nprb
        out = torch.zeros((self.ptheta, self.nz, self.n), dtype=torch.complex128, device=farplane.device)
        ar = torch.arange(P, device=farplane.device)
        for t in range(self.ptheta):
            R, C, rho, gam, keep = self._split(scan[t])
            tmp = self._near(farplane[t]) * torch.conj(probe[t])[None] / self.ndet
            tmp[~keep] = 0
            rows, cols = (R[:, None] + ar)[:, :, None], (C[:, None] + ar)[:, None, :]
            re, im = torch.zeros_like(out[t].real), torch.zeros_like(out[t].real)
            w = lambda x: x[:, None, None]  # noqa: E731
            for dy, dx, wt in ((0, 0, (1 - gam) * (1 - rho)), (0, 1, gam * (1 - rho)),
                               (1, 0, (1 - gam) * rho), (1, 1, gam * rho)):
                v = tmp * w(wt)
                idx = ((rows + dy).expand(-1, P, P), (cols + dx).expand(-1, P, P))
                re.index_put_(idx, v.real, accumulate=True)
                im.index_put_(idx, v.imag, accumulate=True)
            out[t] = torch.complex(re, im)
        return out

    def adj_probe(self, farplane, scan, psi):
        psi = psi.to(torch.complex128)
        out = torch.zeros((self.ptheta, self.nprb, self.nprb), dtype=torch.complex128, device=psi.device)
        for t in range(self.ptheta):
            patches, keep = self._patches(psi[t], scan[t])
            near = self._near(farplane[t])
            near[~keep] = 0
            out[t] = (near * torch.conj(patches)).sum(0) / self.ndet
        return out


class F64CGPtychoSolver(F64PtychoOps, RefCGPtychoSolver):
    """The restated solver (ptycho.py:250-488) over the float64 operators: the referee trajectory."""
    cdtype = torch.complex128
    rdtype = torch.float64

    def __init__(self, *args):
        F64PtychoOps.__init__(self, *args)

    def run_batch(self, data, psi, scan, probe, **kwargs):
        return RefCGPtychoSolver.run_batch(self, data.astype(np.float64), psi.astype(np.complex128),
                                           scan, probe.astype(np.complex128), **kwargs)
