"""Ptychography operators and the conjugate-gradient solver, B200-native.

Same public surface as the reference module src/libtike/cufft/ptycho.py
(`PtychoCuFFT`, `CGPtychoSolver`, context managers, `fwd` / `adj` /
`adj_probe` on device arrays, the `*_batch` host-array helpers, `run_batch`
and `run`), re-backed by the fused sm_100a kernels of libptychofft_b200.so:

  * device arrays are torch CUDA tensors (complex64 / float32) instead of CuPy
    arrays (CuPy is not part of this stack); anything exposing
    `__cuda_array_interface__` is accepted too;
  * the ~60 CuPy elementwise / reduction kernels per CG iteration of
    ptycho.py:308-482 are gone: every pass over the [ptheta, nscan, ndet, ndet]
    far-field volume is one fused kernel that keeps the far field on chip
    (ptx_cg_intensity / ptx_cg_grad / ptx_cg_linesearch), and the object- and
    probe-sized vector updates are the ptx_vec_* kernels.

There is no CPU path: without the CUDA library or a compute-capability-10.x GPU
the operators raise.

Documented deviations from the reference (see DESIGN.md "Quirks"):
  Q1  model='poisson' object gradient: the reference raises UnboundLocalError
      (ptycho.py:357-363 reads `fpsi` before assignment); the evident missing
      line `fpsi = self.fwd(psi, scan, probe[:, k])` is supplied.
  Q5  the position-correction block (ptycho.py:398-403), unconditional in the
      reference, is one fused kernel (ptx_cg_position_shifts) and can be switched
      off with the solver attribute `position_correction` (default True = the
      reference's behaviour; False = the primary CG parity configuration).
  Q9  `_batch` chunks by ptheta (identical at ptheta = 1, the only value the
      reference's helper is valid for).
"""
import concurrent.futures
import ctypes
import warnings
import weakref

import numpy as np
import torch

from libtike.cufft.ptychofft import ptychofft, lib, check, current_stream, PtxError  # noqa: F401

__all__ = ["PtychoCuFFT", "CGPtychoSolver", "register_translation_batch", "line_search_gammas",
           "release_registration_plans"]

MODELS = {"gaussian": 0, "poisson": 1}


def _dev_tensor(x, dtype=None):
    """torch CUDA view of a device array (torch tensor or __cuda_array_interface__ object)."""
    if not isinstance(x, torch.Tensor):
        x = torch.as_tensor(x, device="cuda")
    if dtype is not None and x.dtype != dtype:
        raise AssertionError(f"{x.dtype}")
    return x


def _ptr(t):
    return ctypes.c_void_p(t.data_ptr())


_HOST_REGISTERED = {}  # base address -> (nbytes, weakref.finalize)
#: page-lock large caller-owned NumPy arrays in place the first time a host-array entry point sees
#: them (cudaHostRegister, cached per array, released when the array is garbage collected): their
#: chunks are then DMA'd straight from the caller's memory instead of going through a staging copy
HOST_REGISTER_MIN_BYTES = 8 << 20
#: cap of torch's intra-op CPU thread pool while `run_batch` is in flight (restored afterwards)
HOST_POOL_THREADS = 1


def _host_tensor(a):
    """CPU tensor view of (a chunk of) a caller's NumPy array, page-locked in place when possible."""
    a = np.ascontiguousarray(a)
    t = torch.from_numpy(a)
    if t.is_pinned() or HOST_REGISTER_MIN_BYTES is None:
        return t
    base = a
    while isinstance(base.base, np.ndarray):
        base = base.base
    if base.nbytes < HOST_REGISTER_MIN_BYTES or not base.flags.c_contiguous or not base.flags.writeable:
        return t
    ptr = base.ctypes.data
    if ptr not in _HOST_REGISTERED or _HOST_REGISTERED[ptr][0] < base.nbytes:
        rt = torch.cuda.cudart()
        if ptr in _HOST_REGISTERED:
            _HOST_REGISTERED.pop(ptr)[1]()
        if int(rt.cudaHostRegister(ptr, base.nbytes, 0)) != 0:
            return t  # e.g. memory that cannot be locked: stays pageable, the caller stages it

        def release(p=ptr):
            if _HOST_REGISTERED.pop(p, None) is not None:
                torch.cuda.cudart().cudaHostUnregister(p)
        _HOST_REGISTERED[ptr] = (base.nbytes, weakref.finalize(base, release))
    return t


class _TorchArrayModule(object):
    """Minimal stand-in for the reference's `array_module = cp` class attribute (ptycho.py:55)."""

    complex64 = torch.complex64
    float32 = torch.float32

    @staticmethod
    def array(x):
        return torch.as_tensor(np.ascontiguousarray(x)).cuda()

    @staticmethod
    def zeros(shape, dtype="complex64"):
        dt = getattr(torch, dtype) if isinstance(dtype, str) else dtype
        return torch.zeros(*shape, dtype=dt, device="cuda")

    @staticmethod
    def asnumpy(x):
        return x.detach().cpu().numpy()


class PtychoCuFFT(ptychofft):
    """Base class for ptychography solvers (reference: ptycho.py:34-162).

    Attributes
    ----------
    nscan : int   number of scan positions at each angular view
    nprb : int    pixel width and height of the probe illumination
    ndet : int    pixel width and height of the detector
    ptheta : int  number of angles processed per call (the ctor argument `ntheta`, Q15)
    n, nz : int   pixel width and height of the reconstructed grid
    """

    array_module = _TorchArrayModule
    asnumpy = staticmethod(_TorchArrayModule.asnumpy)

    def __init__(self, nscan, probe_shape, detector_shape, ntheta, nz, n):
        """Argument order of the reference (ptycho.py:58-60)."""
        super().__init__(ntheta, nz, n, nscan, detector_shape, probe_shape)

    def __enter__(self):
        return self

    def __exit__(self, type, value, traceback):
        self.free()

    # ------------------------------------------------------------------ device operators
    def _probe_arg(self, probe):
        """(tensor, angle stride) for a [ptheta, P, P] probe, possibly a probe[:, k] view (Q10)."""
        P = self.nprb
        if probe.dim() == 3 and probe.stride(2) == 1 and probe.stride(1) == P and \
                (probe.shape[0] == 1 or probe.stride(0) >= P * P):
            return probe, (probe.stride(0) if probe.shape[0] > 1 else P * P)
        probe = probe.contiguous()
        return probe, P * P

    def fwd(self, psi, scan, probe):
        """Ptychography transform (FQ), ptycho.py:80-89."""
        psi = _dev_tensor(psi)
        scan = _dev_tensor(scan)
        probe = _dev_tensor(probe)
        assert psi.dtype == torch.complex64, f"{psi.dtype}"
        assert scan.dtype == torch.float32, f"{scan.dtype}"
        assert probe.dtype == torch.complex64, f"{probe.dtype}"
        psi = psi.contiguous()
        scan = scan.contiguous()
        probe, ts = self._probe_arg(probe)
        # fwd overwrites every element of g, so the reference's zero fill (ptycho.py:85) is dropped
        farplane = torch.empty((self.ptheta, self.nscan, self.ndet, self.ndet),
                               dtype=torch.complex64, device=psi.device)
        ptychofft.fwd(self, farplane.data_ptr(), psi.data_ptr(), scan.data_ptr(),
                      probe.data_ptr(), ts)
        return farplane

    def adj(self, farplane, scan, probe):
        """Adjoint ptychography transform (Q*F*), ptycho.py:97-106."""
        farplane = _dev_tensor(farplane)
        scan = _dev_tensor(scan)
        probe = _dev_tensor(probe)
        assert farplane.dtype == torch.complex64, f"{farplane.dtype}"
        assert scan.dtype == torch.float32, f"{scan.dtype}"
        assert probe.dtype == torch.complex64, f"{probe.dtype}"
        farplane = farplane.contiguous()
        scan = scan.contiguous()
        probe, ts = self._probe_arg(probe)
        psi = torch.zeros((self.ptheta, self.nz, self.n), dtype=torch.complex64,
                          device=farplane.device)
        ptychofft.adj(self, psi.data_ptr(), farplane.data_ptr(), scan.data_ptr(),
                      probe.data_ptr(), 0, ts)
        return psi

    def adj_probe(self, farplane, scan, psi):
        """Adjoint ptychography probe transform (O*F*), object is fixed; ptycho.py:113-123."""
        farplane = _dev_tensor(farplane)
        scan = _dev_tensor(scan)
        psi = _dev_tensor(psi)
        assert farplane.dtype == torch.complex64, f"{farplane.dtype}"
        assert scan.dtype == torch.float32, f"{scan.dtype}"
        assert psi.dtype == torch.complex64, f"{psi.dtype}"
        farplane = farplane.contiguous()
        scan = scan.contiguous()
        psi = psi.contiguous()
        probe = torch.zeros((self.ptheta, self.nprb, self.nprb), dtype=torch.complex64,
                            device=farplane.device)
        ptychofft.adj(self, psi.data_ptr(), farplane.data_ptr(), scan.data_ptr(),
                      probe.data_ptr(), 1, 0)
        return probe

    # ------------------------------------------------------------------ host-array helpers
    def _batch(self, function, output, *inputs):
        """Host <-> device shuffle, ptycho.py:70-78, chunked by ptheta (Q9).

        Inputs are copied (pageable, like the reference's `cp.array`) on torch's current stream;
        the result lands in `output` (a host array) chunk by chunk.
        """
        T = self.ptheta
        ntheta = inputs[0].shape[0]
        if ntheta % T:
            raise ValueError(f"leading dimension {ntheta} is not a multiple of ptheta={T}")
        out_t = torch.from_numpy(output)
        for ids in range(0, ntheta, T):
            inputs_gpu = [torch.from_numpy(np.ascontiguousarray(x[ids:ids + T])).cuda(non_blocking=True)
                          for x in inputs]
            res = function(*inputs_gpu)
            out_t[ids:ids + T].copy_(res)
        return output

    def fwd_ptycho_batch(self, psi, scan, probe):
        """Batch of ptychography transforms (FQ) on host arrays, ptycho.py:91-95."""
        data = np.zeros([scan.shape[0], self.nscan, self.ndet, self.ndet], dtype="complex64")
        return self._batch(self.fwd, data, psi, scan, probe)

    def adj_ptycho_batch(self, farplane, scan, probe):
        """Batch of adjoint transforms (Q*F*) on host arrays, ptycho.py:108-111."""
        psi = np.zeros([scan.shape[0], self.nz, self.n], dtype="complex64")
        return self._batch(self.adj, psi, farplane, scan, probe)

    def adj_ptycho_batch_prb(self, farplane, scan, psi):
        """Batch of probe adjoints (O*F*) on host arrays, ptycho.py:125-129."""
        probe = np.zeros([scan.shape[0], self.nprb, self.nprb], dtype="complex64")
        return self._batch(self.adj_probe, probe, farplane, scan, psi)

    def run(self, data, psi, scan, probe, **kwargs):
        """Placeholder for a child's solving function (ptycho.py:131-133)."""
        raise NotImplementedError("Cannot run a base class.")

    def run_batch(self, data, psi, scan, probe, **kwargs):
        """Run by dividing the work into batches of ptheta angles, ptycho.py:135-162.

        Same contract as the reference (host arrays in, copies of psi / probe out, trailing
        `ntheta % ptheta` angles left untouched, Q9).  The host <-> device shuffle is pipelined: the
        inputs of chunk k+1 are staged through pinned memory and copied on a side stream while
        chunk k is being reconstructed, and results are read back asynchronously.
        """
        assert probe.ndim == 4, "probe needs 4 dimensions, not %d" % probe.ndim
        psi_in, probe_in = psi, probe  # staged from the caller's arrays (page-locked in place once)
        T = self.ptheta
        nchunk = scan.shape[0] // T
        if nchunk == 0:
            return {"psi": psi.copy(), "probe": probe.copy()}
        copy_stream = torch.cuda.Stream()
        main = torch.cuda.current_stream()
        device = torch.cuda.current_device()
        # Host-side copies run on ONE worker thread, in order: first the copies of psi / probe the call
        # returns (1.4 GB at 168 angles of 1024^2), then each chunk's results as they land -- the thread
        # that queues kernels never stops to move host memory while the GPU has nothing queued.
        out = {}

        def copy_inputs():
            torch.cuda.set_device(device)
            out["psi"], out["probe"] = psi_in.copy(), probe_in.copy()

        def drain(ids, h_psi, h_prb, landed, keep):
            landed.synchronize()
            out["psi"][ids], out["probe"][ids] = h_psi.numpy(), h_prb.numpy()

        def stage(k):
            ids = slice(k * T, (k + 1) * T)
            with torch.cuda.stream(copy_stream):
                dev = []
                for x in (data, psi_in, scan, probe_in):
                    # large arrays are DMA'd straight from the caller's (page-locked) memory; small ones
                    # (positions, probes) go as they are -- a pinned staging buffer would cost more
                    dev.append(_host_tensor(x[ids]).cuda(non_blocking=True))
                ev = torch.cuda.Event()
                ev.record(copy_stream)
            for t in dev:  # allocated on copy_stream, consumed by kernels on `main`: keep the blocks
                t.record_stream(main)  # out of copy_stream's pool until main is done with them
            return ids, dev, ev

        R = 3  # pinned result buffers in rotation (kept on the solver: pinning costs milliseconds)
        shp_psi, shp_prb = (T,) + tuple(psi.shape[1:]), (T,) + tuple(probe.shape[1:])
        cache = getattr(self, "_h_out", None)
        if cache is None or cache[0] != (shp_psi, shp_prb):
            cache = ((shp_psi, shp_prb),
                     [(torch.empty(shp_psi, dtype=torch.complex64).pin_memory(),
                       torch.empty(shp_prb, dtype=torch.complex64).pin_memory()) for _ in range(R)])
            self._h_out = cache
        h_out = cache[1]
        # keep torch's intra-op CPU pool out of the way of the two threads that matter here (the one that
        # queues kernels and the copy worker) while the batch is in flight
        pool_threads = torch.get_num_threads()
        if pool_threads > HOST_POOL_THREADS:
            torch.set_num_threads(HOST_POOL_THREADS)
        try:
            return self._run_batch_chunks(nchunk, stage, copy_inputs, drain, out, h_out, R, main, kwargs)
        finally:
            if pool_threads > HOST_POOL_THREADS:
                torch.set_num_threads(pool_threads)

    def _run_batch_chunks(self, nchunk, stage, copy_inputs, drain, out, h_out, R, main, kwargs):
        with concurrent.futures.ThreadPoolExecutor(max_workers=1) as worker:
            jobs = [worker.submit(copy_inputs)]
            drained = [None] * R
            nxt = stage(0)
            for k in range(nchunk):
                ids, (data_gpu, psi_gpu, scan_gpu, prb_gpu), ev = nxt
                main.wait_event(ev)
                if k + 1 < nchunk:
                    nxt = stage(k + 1)
                result = self.run(data_gpu, psi_gpu, scan_gpu, prb_gpu, **kwargs)
                if drained[k % R] is not None:
                    drained[k % R].result()  # chunk k - R has left this buffer
                h_psi, h_prb = h_out[k % R]
                h_psi.copy_(result["psi"], non_blocking=True)
                h_prb.copy_(result["probe"], non_blocking=True)
                landed = torch.cuda.Event()
                landed.record(main)
                drained[k % R] = worker.submit(drain, ids, h_psi, h_prb, landed, result)
                jobs.append(drained[k % R])
            for j in jobs:
                j.result()  # (re-raises what a job raised)
        return {"psi": out["psi"], "probe": out["probe"]}


_REG_PLANS = {}


def release_registration_plans():
    """Free the per-(detector size, device) plans `register_translation_batch` keeps between calls."""
    for plan in _REG_PLANS.values():
        plan.free()
    _REG_PLANS.clear()


def register_translation_batch(src_image, target_image, upsample_factor=1, space="real"):
    """Batched sub-pixel image registration by phase correlation (reference: ptycho.py:192-248).

    src_image, target_image : [S, N, N] complex64 device arrays (N in {64, 128, 256, 512});
    space='fourier' means they already are Fourier transforms.  Returns the float64 shifts [S, 2]
    (row, col) as a device tensor, refined to 1/upsample_factor of a pixel by the reference's
    upsampled matrix DFT (ptycho.py:163-190) -- one fused kernel instead of two cupy.fft calls, two
    argmax passes and two complex128 einsums.  Like the reference, a batch of ONE image comes back
    as zeros (its trailing `shape[dim] == 1` loop indexes the batch axis, ptycho.py:243-245).
    """
    src = _dev_tensor(src_image)
    tgt = _dev_tensor(target_image)
    if space.lower() not in ("fourier", "real"):
        raise ValueError("space must be 'real' or 'fourier'")
    if src.dtype != torch.complex64 or tgt.dtype != torch.complex64:
        raise TypeError("register_translation_batch needs complex64 images")
    if src.ndim != 3 or src.shape != tgt.shape or src.shape[1] != src.shape[2]:
        raise ValueError("register_translation_batch needs two [S, N, N] stacks of equal shape")
    if int(upsample_factor) != upsample_factor:
        raise ValueError("upsample_factor must be an integer")
    src, tgt = src.contiguous(), tgt.contiguous()
    S, N = src.shape[0], src.shape[2]
    key = (N, src.device.index)
    plan = _REG_PLANS.get(key)
    if plan is None:
        # a plan carries twiddles and the per-CTA scratch its kernels touch (allocated on first use:
        # a registration-only plan never pays for the solver's accumulators); object size is irrelevant
        plan = _REG_PLANS[key] = ptychofft(1, N + 1, N + 1, 4096, N, N)
    shifts = torch.empty((S, 2), dtype=torch.float64, device=src.device)
    check(lib.ptx_register_translation(plan._h, _ptr(src), _ptr(tgt), S,
                                       1 if space.lower() == "fourier" else 0,
                                       int(upsample_factor), _ptr(shifts), current_stream()))
    if S == 1:
        shifts[0] = 0
    return shifts


def line_search_gammas(c0, ncand):
    return [2.0 ** -(c0 + c) for c in range(ncand)]


class CGPtychoSolver(PtychoCuFFT):
    """Solve the ptychography problem using conjugate gradient (reference: ptycho.py:250-488)."""

    #: Q5 -- execute the reference's position-correction block (ptycho.py:398-403); the reference
    #: has no switch (always on).  False = the primary CG parity configuration.
    position_correction = True
    #: upsampling factor of the position correction (ptycho.py:401)
    position_upsample = 100
    #: keep a copy of every position-correction step's shifts in `shift_log` (diagnostics; one device
    #: copy per iteration, off in production)
    log_shifts = False
    #: step candidates evaluated per fused line-search pass
    ls_candidates = 4
    #: decide the first pass of every line search on the device and let the host read the outcome one
    #: gradient pass later (see `run`); False = the host reads the costs of a pass before it queues
    #: anything else (what an instrumented subclass that overrides `_ls_decide` needs)
    device_line_search = True
    #: queue work ahead of a device-decided search: "adaptive" = not for a search kind (object, probe
    #: mode m) whose previous first pass accepted nothing; True = always; False = never
    ls_run_ahead = "adaptive"
    #: keep F(psi, probe_k) of each gradient pass in HBM (8 N^2 B per pattern and mode) so that the
    #: line search that follows does not gather and transform it again
    cache_far_field = True
    #: single-mode runs with probe recovery: take the a, b sums (and cost) that open the next
    #: iteration (ptycho.py:330-343) from the probe line search, which has just evaluated exactly
    #: that intensity, instead of running the intensity pass again
    reuse_line_search_sums = True
    #: several modes with probe recovery: after the probe line search of mode m, form the summed
    #: intensity that mode m + 1 starts from as I + g^2 p2 + g p3 (one pass over the intensity map)
    #: instead of M forward operators (ptycho.py:424-428)
    incremental_intensity = True
    #: optional libtike.cufft.dist.ScalarComm: when set, the CG scalars (and, in shared-probe mode,
    #: the probe gradient) are all-reduced over its process group, so that ranks holding different
    #: angles behave like ONE reference run over all of them
    comm = None

    def _sum(self, t):
        return self.comm.sum_(t) if self.comm is not None else t

    def _max(self, t):
        return self.comm.max_(t) if self.comm is not None else t

    @staticmethod
    def line_search_sqr(f, p1, p2, p3, step_length=1, step_shrink=0.5):
        """Backtracking line search on f(p1 + g^2 p2 + g p3), verbatim semantics of ptycho.py:253-281."""
        assert step_shrink > 0 and step_shrink < 1
        m = 0
        fp1 = f(p1)
        while f(p1 + step_length ** 2 * p2 + step_length * p3) > fp1 + step_shrink * m:
            if step_length < 1e-32:
                warnings.warn("Line search failed for conjugate gradient.")
                return 0
            step_length *= step_shrink
        return step_length

    # ------------------------------------------------------------------ fused passes
    def _scalars(self, *vals):
        return torch.tensor(vals, dtype=torch.float32, device="cuda")

    def _intensity(self, psi, scan, probe, data, inten, model, iscale=None, red=None):
        """`red`: 3 zeroed device doubles to accumulate into (allocated when not given)."""
        if red is None:
            red = torch.zeros(3, dtype=torch.float64, device=psi.device)
        sc = self._scalars(iscale) if iscale is not None else None
        check(lib.ptx_cg_intensity(self._h, _ptr(psi), _ptr(scan), _ptr(probe), probe.shape[1],
                                   _ptr(data), _ptr(inten) if inten is not None else None,
                                   _ptr(sc) if sc is not None else None, model, _ptr(red),
                                   current_stream()))
        return self._sum(red)

    def _grad(self, what, psi, scan, probe, mode, data, inten, fscale, iscale, gscale, model, out,
              out_stride=0, sc=None, far_out=None):
        """`sc`: device tensor {fscale, iscale, gscale} (then the three host values are ignored).
        `far_out`: [T,S,N,N] complex64 that receives F(psi, probe[:, mode]) for the next line search."""
        if sc is None:
            sc = self._scalars(fscale, iscale, gscale)
        check(lib.ptx_cg_grad(self._h, what, _ptr(psi), _ptr(scan), _ptr(probe), probe.shape[1],
                              mode, _ptr(data), _ptr(inten) if inten is not None else None,
                              _ptr(sc), model, _ptr(out), out_stride,
                              _ptr(far_out) if far_out is not None else None, current_stream()))

    def _line_search(self, obj_a, prb_a, nm_a, m_a, obj_b, prb_b, nm_b, m_b, npairs, scan, data,
                     p1, model, far_a=None, want_ab=False, p23=None, slots=None, gam=None, carry=None):
        """Fused line_search_sqr: evaluates `ls_candidates` halvings per pass (ptycho.py:272-281).
        `far_a`: [npairs, T,S,N,N] cached first far fields (see `_grad(far_out=...)`).
        `want_ab`: also reduce a = sum sqrt(I data), b = sum I for every candidate intensity; after
        the call `self._ls_ab` holds (a, b, cost) of the intensity at HALF the returned step -- the
        update the solver applies (ptycho.py:393, 461).  `self._ls_ab_dev` then is (device cost buffer,
        index of a, of b, of the cost) for a device-side hand-over (ptx_cg_pick3).  So that the half
        step is always among the candidates of the deciding pass, a pass with `want_ab` decides on its
        first K - 1 candidates only (the order in which steps are tried is unchanged).
        `slots`: [npass, 16] zeroed device doubles, one row per fused pass (allocated beyond that).
        `gam`: one device float.  When given, the first pass is decided ON THE DEVICE
        (ptx_cg_ls_decide: half the accepted step -> gam, {a, b, cost} -> `carry`) and the call returns
        at once with a function `finish() -> (step, refit)` instead of the step: the caller queues the
        work that follows against `gam` and calls finish() when it needs the host-side value.
        `refit` is True when no candidate of the first pass was accepted; finish() has then run the
        remaining passes the host-driven way and `gam` / `carry` still hold the no-op values."""
        K = int(self.ls_candidates)
        kdec = K - 1 if want_ab else K
        assert kdec >= 1
        self._ls_begin()
        self._ls_ab = None
        self._ls_ab_dev = None
        npass = [0]
        if getattr(self, "_h_cost", None) is None:
            self._h_cost = torch.empty(16, dtype=torch.float64).pin_memory()
            self._h_rows = torch.empty((4, 16), dtype=torch.float64).pin_memory()
            self._h_row_ev = [torch.cuda.Event() for _ in range(4)]
            self._h_row_i = 0

        def one_pass(c0):
            # 5 costs (+ 5 a + 5 b) of this pass
            cost = (slots[npass[0]] if slots is not None and npass[0] < slots.shape[0]
                    else torch.zeros(16, dtype=torch.float64, device=obj_a.device))
            npass[0] += 1
            check(lib.ptx_cg_linesearch(self._h, _ptr(obj_a), _ptr(prb_a), nm_a, m_a, _ptr(obj_b),
                                        _ptr(prb_b), nm_b, m_b, npairs, _ptr(scan), _ptr(data),
                                        _ptr(p1) if p1 is not None else None,
                                        _ptr(far_a) if far_a is not None else None, model, c0, K,
                                        1 if want_ab else 0, _ptr(p23) if p23 is not None else None,
                                        _ptr(cost), current_stream()))
            return self._sum(cost)

        def done(step, c, cost, c0):
            self.ls_steps.append(step)
            if want_ab:
                j = 0 if step == 0 else int(round(-np.log2(step))) - c0 + 2  # slot of step / 2
                if 0 <= j <= K:
                    self._ls_ab = (c[5 + j], c[10 + j], c[j])
                    self._ls_ab_dev = (cost, 5 + j, 10 + j, j)
            return step

        def host_loop(c0):
            while True:
                cost = one_pass(c0)
                # the one host read of a pass: pinned buffer, no pageable staging, no allocation
                self._h_cost.copy_(cost, non_blocking=True)
                torch.cuda.current_stream().synchronize()
                c = self._h_cost.numpy().copy()
                self.ls_log.append((c0, c[:1 + K].copy()))
                step = self._ls_decide(c0, c, kdec)
                if step is not None:
                    return done(step, c, cost, c0)
                c0 += kdec

        if gam is None:
            return host_loop(0)
        cost = one_pass(0)
        check(lib.ptx_cg_ls_decide(_ptr(cost), 0, kdec, _ptr(gam), _ptr(carry) if carry is not None else None,
                                   current_stream()))
        # (at most one search is in flight at a time: four rotating rows are ample)
        row = self._h_rows[self._h_row_i % 4]
        landed = self._h_row_ev[self._h_row_i % 4]
        self._h_row_i += 1
        row.copy_(cost, non_blocking=True)
        landed.record()

        def finish():
            landed.synchronize()
            c = row.numpy().copy()
            self.ls_log.append((0, c[:1 + K].copy()))
            jj = int(c[15])
            if jj >= 0:
                return done(2.0 ** -jj, c, cost, 0), False
            return host_loop(kdec), True

        return finish

    def _ls_begin(self):
        """Called once per line search, before its first pass (a seam for instrumented subclasses)."""

    def _ls_decide(self, c0, c, K):
        """The decision line_search_sqr (ptycho.py:272-281) takes on the K candidate costs c[1:1+K] of
        steps 2^-c0 .. 2^-(c0+K-1) against c[0] = f(p1): the accepted step, 0 when the search failed,
        or None to keep halving."""
        for j in range(K):
            step = 2.0 ** -(c0 + j)
            if not (c[1 + j] > c[0]):
                return step
            if step < 1e-32:
                warnings.warn("Line search failed for conjugate gradient.")
                return 0
        return None

    def _dai_yuan(self, grad, grad0, d, first, red=None, zero_grad=False):
        """`red`: 3 zeroed device doubles; `zero_grad`: leave `grad` zeroed for the next pass to
        accumulate into (saves the fill launch)."""
        if red is None:
            red = torch.zeros(3, dtype=torch.float64, device=grad.device)
        n = grad.numel()
        if not first:
            check(lib.ptx_vec_dai_yuan_reduce(_ptr(grad), _ptr(grad0), _ptr(d), n, _ptr(red),
                                              current_stream()))
            self._sum(red)
        check(lib.ptx_vec_dai_yuan_update(_ptr(grad), _ptr(grad0), _ptr(d), n, _ptr(red),
                                          (1 if first else 0) | (2 if zero_grad else 0), current_stream()))

    def _axpy(self, y, x, alpha):
        check(lib.ptx_vec_axpy_s(_ptr(y), _ptr(x), y.numel(), float(alpha), current_stream()))

    def _absmax(self, x, out=None):
        """`out`: one zeroed device float (allocated when not given)."""
        if out is None:
            out = torch.zeros(1, dtype=torch.float32, device=x.device)
        if x.is_contiguous():
            check(lib.ptx_vec_absmax(_ptr(x), x.numel(), _ptr(out), current_stream()))
        else:  # probe[:, k] with ptheta > 1: one launch per angle, same accumulator
            for t in range(x.shape[0]):
                check(lib.ptx_vec_absmax(_ptr(x[t]), x[t].numel(), _ptr(out), current_stream()))
        return self._max(out)

    # ------------------------------------------------------------------ fused gradient, host arrays
    def grad_ptycho_batch(self, data, psi, scan, probe, model="gaussian"):
        """Object gradient of the data-fit cost for a batch of angles given as HOST arrays.

        Per angle chunk of `ptheta`: sum_k Q_k* F* [ F Q_k psi * (1 - sqrt(d)/(sqrt(I)+1e-32)) ]
        (gaussian; `d/(I+1e-32)` for poisson), I = sum_k |F Q_k psi|^2 -- the forward and adjoint
        chain of ptycho.py:347-363 without the CG normalisations -- computed by ONE fused kernel
        per mode with the far field kept on chip.  Inputs are staged through pinned buffers on two
        streams so that the H2D copy of chunk i+1 and the D2H copy of chunk i-1 overlap the
        kernels of chunk i.  Returns a host array [ntheta, nz, n] complex64.
        """
        assert probe.ndim == 4, "probe needs 4 dimensions, not %d" % probe.ndim
        mdl = MODELS[model]
        T, M = self.ptheta, probe.shape[1]
        ntheta = scan.shape[0]
        if ntheta % T:
            raise ValueError(f"leading dimension {ntheta} is not a multiple of ptheta={T}")
        out = np.empty((ntheta, self.nz, self.n), dtype=np.complex64)
        dev = torch.device("cuda", torch.cuda.current_device())
        st = getattr(self, "_pipe", None)
        if st is None or st["M"] != M:
            def pin(shape, dtype):
                return torch.empty(shape, dtype=dtype).pin_memory()
            # {fscale, iscale, gscale} = 1: made ONCE -- torch.tensor(..., device=) is a pageable, hence
            # stream-synchronous, copy that would stall the host behind the chunk's 67 MB upload
            st = {"M": M, "streams": [torch.cuda.Stream(), torch.cuda.Stream()], "slots": [],
                  "sc": torch.ones(3, dtype=torch.float32, device=dev)}
            for _ in range(2):
                st["slots"].append({
                    "h": (pin((T, self.nscan, self.ndet, self.ndet), torch.float32),
                          pin((T, self.nz, self.n), torch.complex64),
                          pin((T, self.nscan, 2), torch.float32),
                          pin((T, M, self.nprb, self.nprb), torch.complex64)),
                    "d": (torch.empty((T, self.nscan, self.ndet, self.ndet), dtype=torch.float32, device=dev),
                          torch.empty((T, self.nz, self.n), dtype=torch.complex64, device=dev),
                          torch.empty((T, self.nscan, 2), dtype=torch.float32, device=dev),
                          torch.empty((T, M, self.nprb, self.nprb), dtype=torch.complex64, device=dev)),
                    "g": torch.empty((T, self.nz, self.n), dtype=torch.complex64, device=dev),
                    "gh": pin((T, self.nz, self.n), torch.complex64),
                    "inten": (torch.empty((T, self.nscan, self.ndet, self.ndet), dtype=torch.float32,
                                          device=dev) if M > 1 else None),
                    "done": None, "ids": None})
            self._pipe = st
        nchunk = ntheta // T

        def drain(slot):
            if slot["done"] is not None:
                slot["done"].synchronize()
                out[slot["ids"]] = slot["gh"].numpy()
                slot["done"] = None

        computed = None  # event after the kernels of the previous chunk
        for c in range(nchunk):
            slot, stream = st["slots"][c % 2], st["streams"][c % 2]
            drain(slot)  # the pinned buffers of this slot are free again
            ids = slice(c * T, (c + 1) * T)
            srcs = []
            for hbuf, src in zip(slot["h"], (data, psi, scan, probe)):
                t = _host_tensor(src[ids])
                if not t.is_pinned():  # pageable caller memory is staged; pinned memory is DMA'd as is
                    hbuf.copy_(t)
                    t = hbuf
                srcs.append(t)
            with torch.cuda.stream(stream):
                for t, dbuf in zip(srcs, slot["d"]):
                    dbuf.copy_(t, non_blocking=True)
                d_data, d_psi, d_scan, d_prb = slot["d"]
                slot["g"].zero_()
                # the plan's per-CTA scratch (staging frames, accumulators) is indexed by blockIdx
                # only: kernels of two chunks must never overlap, only copies do
                if computed is not None:
                    stream.wait_event(computed)
                if M > 1:
                    self._intensity(d_psi, d_scan, d_prb, d_data, slot["inten"], mdl)
                for k in range(M):
                    self._grad(0, d_psi, d_scan, d_prb, k, d_data, slot["inten"], 1.0, 1.0, 1.0, mdl,
                               slot["g"], sc=st["sc"])
                computed = torch.cuda.Event()
                computed.record(stream)
                slot["gh"].copy_(slot["g"], non_blocking=True)
                slot["done"] = torch.cuda.Event()
                slot["done"].record(stream)
                slot["ids"] = ids
        for slot in st["slots"]:
            drain(slot)
        return out

    # ------------------------------------------------------------------ the solver
    def run(self, data, psi, scan, probe, piter, model="gaussian", recover_prb=False,
            ortho_prb=False):
        """Conjugate gradients for ptychography (reference: ptycho.py:283-488).

        Parameters
        ----------
        model : str, 'gaussian' or 'poisson' -- the noise model used for the gradient.
        piter : int -- number of gradient steps.
        recover_prb : bool -- recover the probe or keep the given one.
        ortho_prb : accepted and ignored, like the reference (Q8).
        """
        assert probe.ndim == 4, "probe needs 4 dimensions, not %d" % probe.ndim
        if model not in MODELS:
            raise ValueError("model must be 'gaussian' or 'poisson'")
        mdl = MODELS[model]
        data = _dev_tensor(data).contiguous()
        scan_in = _dev_tensor(scan)  # position correction updates the caller's array (ptycho.py:403)
        scan = scan_in if scan_in.is_contiguous() else scan_in.contiguous()
        psi = _dev_tensor(psi).contiguous().clone()  # the reference rebinds psi (ptycho.py:405)
        probe_in = _dev_tensor(probe)
        probe = probe_in if probe_in.is_contiguous() else probe_in.contiguous()  # mutated in place (Q6)
        assert data.dtype == torch.float32 and psi.dtype == torch.complex64
        assert probe.dtype == torch.complex64 and scan.dtype == torch.float32
        T, S, M, P = self.ptheta, self.nscan, probe.shape[1], self.nprb
        dev = psi.device
        multi = M > 1
        inten = torch.empty_like(data) if multi else None
        # F(psi, probe_k) of the gradient passes, re-read by the line searches that follow them
        # (both shortcuts are dropped, not shrunk, when their arrays would not fit comfortably)
        free_bytes = torch.cuda.mem_get_info(dev)[0]
        fits = [bool(self.cache_far_field and 8 * M * data.numel() < 0.4 * free_bytes),
                bool(multi and recover_prb and self.incremental_intensity
                     and 8 * (M + 1) * data.numel() < 0.6 * free_bytes)]
        if self.comm is not None:
            # every rank must take the same shortcuts: they decide how many collectives an iteration
            # issues (a rank that skips an intensity pass skips its all-reduce)
            f = torch.tensor([float(x) for x in fits], dtype=torch.float64, device=dev)
            fits = [bool(x > 0.5) for x in self.comm.min_(f).cpu().tolist()]
        far = (torch.empty((M,) + tuple(data.shape), dtype=torch.complex64, device=dev)
               if fits[0] else None)
        p23 = (torch.empty(tuple(data.shape) + (2,), dtype=torch.float32, device=dev)
               if fits[1] else None)
        sum_data = float(self._sum(data.sum(dtype=torch.float64).reshape(1))) if mdl == 0 else 0.0

        gradpsi = torch.zeros_like(psi)
        gradpsi0 = torch.zeros_like(psi)
        dpsi = torch.zeros_like(psi)
        # probe-side CG state is kept mode-major [M, T, P, P] so that x[m] is contiguous
        gradprb = torch.zeros((M, T, P, P), dtype=torch.complex64, device=dev)
        gradprb0 = torch.zeros_like(gradprb)
        dprb = torch.zeros_like(gradprb)

        s_dev = torch.zeros(1, dtype=torch.float32, device=dev)
        sc_obj = torch.ones(3, dtype=torch.float32, device=dev)   # {fscale, iscale, gscale}, object pass
        sc_prb = torch.ones(3, dtype=torch.float32, device=dev)   # {1, 1, gscale}, probe pass

        # Per-iteration device scalars live in ONE preallocated buffer that a single memset clears at
        # the top of an iteration; every reduction of the iteration accumulates into its own slot, so
        # the loop allocates nothing and launches no torch fill / clone / pageable copy.
        LSP = 6                                     # fused line-search passes with a preallocated row
        nd = 3 + 3 + 16 * LSP + M * (3 + 3 + 16 * LSP)
        buf = torch.zeros(8 * nd + 4 * (2 * M), dtype=torch.uint8, device=dev)
        ws = buf[:8 * nd].view(torch.float64)
        wf = buf[8 * nd:].view(torch.float32)        # absmax slots: M for |probe_k|, M for |psi|
        red_int, dy_obj = ws[0:3], ws[3:6]
        ls_obj = ws[6:6 + 16 * LSP].view(LSP, 16)
        o = 6 + 16 * LSP
        per = 3 + 3 + 16 * LSP
        red_prb = [ws[o + m * per:o + m * per + 3] for m in range(M)]
        dy_prb = [ws[o + m * per + 3:o + m * per + 6] for m in range(M)]
        ls_prb = [ws[o + m * per + 6:o + (m + 1) * per].view(LSP, 16) for m in range(M)]
        red_carried = torch.zeros(3, dtype=torch.float64, device=dev)  # survives the memset
        psi_new = torch.empty_like(psi) if self.position_correction else None
        shifts = torch.empty((S, 2), dtype=torch.float64, device=dev) if self.position_correction else None
        nbuf = buf.numel()

        print("# congujate gradient parameters\n"
              "iteration, step size object, step size probe, function min")  # csv column headers
        reuse = bool(self.reuse_line_search_sums) and recover_prb and (M == 1 or p23 is not None)
        self.history = []  # (iteration, step size object, step size probe) -- diagnostics only
        self.ls_log = []   # (first candidate exponent, [f(0), f(2^-c0), ...]) per fused pass
        self.ls_steps = []  # raw result of every line search, in call order (replayable by an instrumented subclass)
        self.shift_log = []  # device [S,2] float64 shifts of every position-correction step (log_shifts)
        self.ls_refits = 0  # line searches whose first fused pass accepted nothing (device_line_search)

        # Line searches are decided on the device (ptx_cg_ls_decide) and the updates behind them read the
        # step from device memory, so the host never waits for a cost before it queues the next kernels.
        # It looks at the outcome one gradient pass LATER (`settle`): by then the answer has long landed
        # in pinned memory and the GPU has work queued.  What is queued ahead of that look is chosen so
        # that it is harmless when the first pass accepted nothing (step 0: the updates are no-ops, the
        # gradient computed from the unchanged state is thrown away and queued again after the remaining
        # passes have run the host-driven way).  Where no such work exists the host settles at once.
        dls = bool(self.device_line_search)
        gam = torch.zeros(1 + M, dtype=torch.float32, device=dev)  # half steps: object, probe modes
        gam_obj, gam_prb = gam[0:1], [gam[1 + m:2 + m] for m in range(M)]
        probe_k = [probe[:, k] for k in range(M)]           # views made once, not once per launch
        wf_prb, wf_psi = [wf[k:k + 1] for k in range(M)], [wf[M + m:M + m + 1] for m in range(M)]
        gprb_m, gprb0_m, dprb_m = list(gradprb), list(gradprb0), list(dprb)
        far_m = list(far) if far is not None else [None] * M
        prb_tm = [[probe[t, m] for t in range(T)] for m in range(M)]
        dprb_mt = [[dprb[m, t] for t in range(T)] for m in range(M)]
        state = {"carried": False, "pending": None}
        # a search kind (object, probe mode m) whose last first pass accepted nothing is not run ahead of:
        # on problems whose steps are habitually below 2^-3 the discarded gradient would be pure loss
        trust = [self.ls_run_ahead is not False] * (1 + M)

        def learn(kind, refit):
            if self.ls_run_ahead == "adaptive":
                trust[kind] = not refit
        recs = []  # per-iteration records still waiting for a step size

        def emit():
            while recs and recs[0]["gpsi"] is not None and recs[0]["gprb"] is not None:
                r = recs.pop(0)
                self.history.append((r["i"], r["gpsi"], r["gprb"]))
                # check convergence (ptycho.py:474-482)
                if r["i"] % 32 == 0:
                    print("%4d, %.3e, %.3e, %.7e" % (r["i"], r["gpsi"], r["gprb"], r["fmin"]))

        def settle(redo=None):
            """Host-side end of the line search still in flight (if any): log it, and when its first
            pass accepted nothing, apply the update with the step the remaining passes found (`fix`)
            and queue the work that was issued ahead of this call again (`redo`)."""
            if state["pending"] is None:
                return
            finish, fix = state["pending"]
            state["pending"] = None
            step, refit = finish()
            self.ls_refits += int(refit)
            fix(step, refit)
            if refit and redo is not None:
                redo()
            emit()

        def submit(finish, fix, speculate):
            state["pending"] = (finish, fix)
            if not speculate:
                settle()

        def open_object_pass():
            """Top of an iteration up to the object gradient (ptycho.py:327-363)."""
            check(lib.ptx_vec_zero(_ptr(buf), nbuf, current_stream()))
            # 1) object retrieval subproblem with fixed probes (ptycho.py:327-345).  a, b, the probe
            # rescaling and the gradient scalars stay on the device: no host round trip here
            if state["carried"]:  # a, b, cost of this very intensity were left on the device by the last line search
                red = red_carried
            else:
                red = self._intensity(psi, scan, probe, data, inten, mdl, red=red_int)
            check(lib.ptx_cg_prep_scale(_ptr(red), mdl, _ptr(s_dev), _ptr(sc_obj), current_stream()))
            check(lib.ptx_vec_scale(_ptr(probe), probe.numel(), _ptr(s_dev), current_stream()))
            # gradient (ptycho.py:346-363); gradpsi is zero here (initially, then left so by _dai_yuan)
            for k in range(M):
                check(lib.ptx_cg_prep_gscale(_ptr(self._absmax(probe_k[k], out=wf_prb[k])), 1.0, _ptr(sc_obj),
                                             current_stream()))
                self._grad(0, psi, scan, probe, k, data, inten, 0, 0, 0, mdl, gradpsi, sc=sc_obj,
                           far_out=far_m[k])
            return red

        def reopen_object_pass():
            check(lib.ptx_vec_zero(_ptr(gradpsi), gradpsi.numel() * 8, current_stream()))
            return open_object_pass()

        def open_probe_pass(m):
            """Probe subproblem of mode m up to its gradient (ptycho.py:420-441)."""
            if multi and not (m > 0 and p23 is not None):
                self._intensity(psi, scan, probe, data, inten, mdl, red=red_prb[m])
            kg = (float(M) if mdl == 0 else 1.0) / S      # Q13: * nmodes only for gaussian
            check(lib.ptx_cg_prep_gscale(_ptr(self._absmax(psi, out=wf_psi[m])), kg, _ptr(sc_prb),
                                         current_stream()))
            # gradprb[m] is zero here (initially, then left so by _dai_yuan)
            self._grad(1, psi, scan, probe, m, data, inten, 0, 0, 0, mdl, gprb_m[m], P * P,
                       sc=sc_prb, far_out=far_m[m])
            if self.comm is not None:
                self.comm.probe_grad_(gprb_m[m])

        def reopen_probe_pass(m):
            check(lib.ptx_vec_zero(_ptr(wf_psi[m]), 4, current_stream()))
            check(lib.ptx_vec_zero(_ptr(gprb_m[m]), gprb_m[m].numel() * 8, current_stream()))
            open_probe_pass(m)

        def update_probe(m, g):
            """probe[:, m] += g dprb[m] (ptycho.py:463); g a host value or one device float"""
            for t in range(T):
                if torch.is_tensor(g):
                    check(lib.ptx_vec_axpy(_ptr(prb_tm[m][t]), _ptr(dprb_mt[m][t]), P * P, _ptr(g), current_stream()))
                else:
                    self._axpy(prb_tm[m][t], dprb_mt[m][t], g)

        for i in range(piter):
            rec = {"i": i, "gpsi": None, "gprb": None if recover_prb else 0, "fmin": None}
            recs.append(rec)
            red = open_object_pass()
            if state["pending"] is not None:
                # the probe line search of the previous iteration: `red` was queued ahead of its outcome
                box = []
                settle(lambda: box.append(reopen_object_pass()))
                if box:
                    red = box[0]
            state["carried"] = False
            if i % 32 == 0:  # cost of this iteration's absfpsi, printed by emit() (ptycho.py:481-482)
                if mdl == 0:
                    r = red.cpu().numpy()
                    s = float(np.float32(np.float32(r[0]) / np.float32(r[1])))
                    rec["fmin"] = s ** 2 * r[1] - 2.0 * s * r[0] + sum_data
                else:
                    rec["fmin"] = float(self._intensity(psi, scan, probe, data, None, mdl)[2])
            # Dai-Yuan direction (ptycho.py:364-372)
            self._dai_yuan(gradpsi, gradpsi0, dpsi, i == 0, red=dy_obj, zero_grad=True)
            # line search (ptycho.py:374-393)
            ls = self._line_search(psi, probe, M, 0, dpsi, probe, M, 0, M, scan, data, None, mdl,
                                   far_a=far, slots=ls_obj, gam=gam_obj if dls else None)
            correct = self.position_correction and i > 0
            rank0 = self.comm is None or self.comm.rank == 0  # angle 0 of the run lives on rank 0

            def propose(g, psi=psi, psi_new=psi_new):
                # psi + gamma dpsi and its registration against psi (ptycho.py:398-402): nothing is
                # overwritten, so this can be queued before the step is known to be final
                if torch.is_tensor(g):
                    check(lib.ptx_vec_axpy_out_dev(_ptr(psi_new), _ptr(psi), _ptr(dpsi), psi.numel(), _ptr(g),
                                                   current_stream()))
                else:
                    check(lib.ptx_vec_axpy_out(_ptr(psi_new), _ptr(psi), _ptr(dpsi), psi.numel(), float(g),
                                               current_stream()))
                if rank0 and S > 1:
                    check(lib.ptx_cg_position_shifts(self._h, _ptr(psi), _ptr(psi_new), _ptr(scan),
                                                     int(self.position_upsample), _ptr(shifts),
                                                     current_stream()))

            if dls:
                if correct:
                    def fix(step, refit, rec=rec):
                        rec["gpsi"] = 0.5 * step
                        learn(0, refit)
                    if trust[0]:
                        propose(gam_obj)
                        submit(ls, fix, True)
                        settle(lambda rec=rec: propose(rec["gpsi"]))
                    else:
                        submit(ls, fix, False)
                        propose(rec["gpsi"])
                else:
                    # update psi (ptycho.py:405)
                    check(lib.ptx_vec_axpy(_ptr(psi), _ptr(dpsi), psi.numel(), _ptr(gam_obj), current_stream()))

                    def fix(step, refit, rec=rec, psi=psi):
                        rec["gpsi"] = 0.5 * step
                        learn(0, refit)
                        if refit:
                            self._axpy(psi, dpsi, 0.5 * step)
                    # the probe gradient that follows only reads psi; a new iteration would rescale the probe
                    submit(ls, fix, recover_prb and trust[0])
            else:
                rec["gpsi"] = 0.5 * ls
                if correct:
                    propose(rec["gpsi"])
                else:
                    self._axpy(psi, dpsi, rec["gpsi"])
            if correct:
                # position correction (ptycho.py:398-403): the scan positions of angle 0 move by the shifts
                # between psi and psi + gamma dpsi -- the caller's scan array is updated in place, like
                # the reference's
                if rank0:
                    if S > 1:
                        check(lib.ptx_cg_apply_shifts(_ptr(scan), _ptr(shifts), S, current_stream()))
                    if self.log_shifts:  # (a batch of ONE position gets zero shifts: ptycho.py:243-245)
                        self.shift_log.append(shifts.clone() if S > 1 else torch.zeros_like(shifts))
                psi, psi_new = psi_new, psi

            if recover_prb:
                for m in range(M):
                    # 2) probe retrieval subproblem with fixed object (ptycho.py:420-441)
                    open_probe_pass(m)
                    settle(lambda m=m: reopen_probe_pass(m))
                    # Dai-Yuan direction (ptycho.py:442-450)
                    self._dai_yuan(gprb_m[m], gprb0_m[m], dprb_m[m], i == 0, red=dy_prb[m], zero_grad=True)
                    # line search (ptycho.py:451-461)
                    last = m == M - 1
                    want = reuse and last
                    ls = self._line_search(psi, probe, M, m, psi, dprb_m[m], 1, 0, 1, scan, data, inten, mdl,
                                           far_a=far_m[m],
                                           want_ab=want, p23=p23, slots=ls_prb[m],
                                           gam=gam_prb[m] if dls else None,
                                           carry=red_carried if (dls and want) else None)

                    def after_search(step, refit, m=m, last=last, want=want, rec=rec):
                        """What the host does once a probe line search has a step it did not get on the device."""
                        g = 0.5 * step
                        if last:
                            rec["gprb"] = g
                        learn(1 + m, refit)
                        if not refit:
                            return
                        # a, b, cost of the intensity the NEXT iteration opens with: only the last mode's
                        state["carried"] = bool(want and self._ls_ab_dev is not None)
                        if state["carried"]:
                            cbuf, ia, ib, ic = self._ls_ab_dev
                            check(lib.ptx_cg_pick3(_ptr(red_carried), _ptr(cbuf), ia, ib, ic, current_stream()))
                        if p23 is not None and (m + 1 < M or state["carried"]):
                            # the intensity mode m + 1 (or the next iteration) will start from
                            check(lib.ptx_cg_intensity_step(_ptr(inten), _ptr(p23), inten.numel(), float(g),
                                                            current_stream()))
                        update_probe(m, g)

                    if dls:
                        g_dev = gam_prb[m]
                        state["carried"] = want  # the device left a, b, cost in red_carried ({1, 1, 0} if it accepted nothing)
                        if p23 is not None and (m + 1 < M or want):
                            check(lib.ptx_cg_intensity_step_dev(_ptr(inten), _ptr(p23), inten.numel(), _ptr(g_dev),
                                                                current_stream()))
                        update_probe(m, g_dev)
                        # ahead of the outcome: the next mode's gradient only reads the probe; a new iteration
                        # rescales it, which is exact only with the device's {1, 1, 0} hand-over
                        submit(ls, after_search, ((not last) or want) and trust[1 + m])
                    else:
                        after_search(ls, True)
            emit()
        settle()
        emit()

        if probe is not probe_in:
            probe_in.copy_(probe)
        if scan is not scan_in:
            scan_in.copy_(scan)
        return {"psi": psi, "probe": probe_in}
