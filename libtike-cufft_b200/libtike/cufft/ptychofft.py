"""Native operator class `ptychofft`, bound to the B200 C ABI through ctypes.

This module takes the place of the compiled SWIG / pybind11 extension module
`libtike.cufft.ptychofft` of the reference (src/cuda/swig/ptychofft.i:1-26,
src/cuda/pybind11/ptychofft.cxx:1-27): same class name, same constructor
(`ptychofft(ptheta, nz, n, nscan, detector_shape, probe_shape)`), the same six
read-only attributes, and `fwd(g_, f_, scan_, prb_)`, `adj(f_, g_, scan_, prb_,
flg)`, `free()` taking raw device addresses as Python ints
(src/include/ptychofft.cuh:34-43).

The shared library is libptychofft_b200.so (include/ptychofft_b200.h), built
in-tree by `__graft_entry__.build()`.  There is NO fallback of any kind: if the
library is missing, or the device is not compute capability 10.x, importing or
constructing raises.
"""
import ctypes
import os

__all__ = ["ptychofft", "PtxError", "lib", "current_stream", "launch_count"]

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_NAME = "libptychofft_b200.so"

# every symbol include/ptychofft_b200.h declares: (name, restype, argtypes)
_vp, _sz, _i, _fp, _dp = (ctypes.c_void_p, ctypes.c_size_t, ctypes.c_int,
                          ctypes.c_void_p, ctypes.c_void_p)
SYMBOLS = [
    ("ptx_last_error", ctypes.c_char_p, []),
    ("ptx_device_ok", _i, []),
    ("ptx_launch_count", ctypes.c_ulonglong, []),
    ("ptx_create", _i, [ctypes.POINTER(_vp), _sz, _sz, _sz, _sz, _sz, _sz]),
    ("ptx_free", _i, [_vp]),
    ("ptx_destroy", _i, [_vp]),
    ("ptx_dim", _sz, [_vp, _i]),
    ("ptx_fwd", _i, [_vp, _vp, _vp, _vp, _vp, _sz, _vp]),
    ("ptx_debug_nearplane", _i, [_vp, _vp, _vp, _vp, _vp, _sz, _vp]),
    ("ptx_adj", _i, [_vp, _vp, _vp, _vp, _vp, _sz, _i, _vp]),
    ("ptx_cg_intensity", _i, [_vp, _vp, _vp, _vp, _i, _fp, _fp, _fp, _i, _dp, _vp]),
    ("ptx_cg_grad", _i, [_vp, _i, _vp, _vp, _vp, _i, _i, _fp, _fp, _fp, _i, _vp, _sz, _vp, _vp]),
    ("ptx_cg_linesearch", _i, [_vp, _vp, _vp, _i, _i, _vp, _vp, _i, _i, _i, _vp, _fp, _fp, _vp, _i, _i,
                               _i, _i, _vp, _dp, _vp]),
    ("ptx_cg_intensity_step", _i, [_fp, _vp, _sz, ctypes.c_float, _vp]),
    ("ptx_register_translation", _i, [_vp, _vp, _vp, _sz, _i, _i, _dp, _vp]),
    ("ptx_cg_position_shifts", _i, [_vp, _vp, _vp, _vp, _i, _dp, _vp]),
    ("ptx_prepare_data", _i, [_fp, _vp, _sz, _sz, ctypes.c_float, _i, _fp, _vp]),
    ("ptx_vec_dai_yuan_reduce", _i, [_vp, _vp, _vp, _sz, _dp, _vp]),
    ("ptx_vec_dai_yuan_update", _i, [_vp, _vp, _vp, _sz, _dp, _i, _vp]),
    ("ptx_vec_axpy", _i, [_vp, _vp, _sz, _fp, _vp]),
    ("ptx_vec_axpy_s", _i, [_vp, _vp, _sz, ctypes.c_float, _vp]),
    ("ptx_vec_axpy_out", _i, [_vp, _vp, _vp, _sz, ctypes.c_float, _vp]),
    ("ptx_vec_zero", _i, [_vp, _sz, _vp]),
    ("ptx_cg_apply_shifts", _i, [_fp, _dp, _sz, _vp]),
    ("ptx_cg_pick3", _i, [_dp, _dp, _i, _i, _i, _vp]),
    ("ptx_cg_ls_decide", _i, [_dp, _i, _i, _fp, _dp, _vp]),
    ("ptx_vec_axpy_out_dev", _i, [_vp, _vp, _vp, _sz, _fp, _vp]),
    ("ptx_cg_intensity_step_dev", _i, [_fp, _vp, _sz, _fp, _vp]),
    ("ptx_cg_prep_scale", _i, [_dp, _i, _fp, _fp, _vp]),
    ("ptx_cg_prep_gscale", _i, [_fp, ctypes.c_double, _fp, _vp]),
    ("ptx_vec_scale", _i, [_vp, _sz, _fp, _vp]),
    ("ptx_vec_absmax", _i, [_vp, _sz, _fp, _vp]),
]


class PtxError(RuntimeError):
    """A C-ABI call returned a non-zero status (the reference checks none)."""


def _load():
    path = os.environ.get("PTYCHOFFT_B200_LIB", os.path.join(_HERE, _LIB_NAME))
    if not os.path.exists(path):
        raise ImportError(
            f"{path} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(there is no CPU or cuFFT fallback for libtike.cufft)")
    handle = ctypes.CDLL(path)
    for name, restype, argtypes in SYMBOLS:
        fn = getattr(handle, name)  # AttributeError if the library is stale
        fn.restype = restype
        fn.argtypes = argtypes
    return handle


lib = _load()


def check(rc):
    if rc != 0:
        raise PtxError("libptychofft_b200: error %d: %s" % (rc, lib.ptx_last_error().decode()))


def current_stream():
    """cudaStream_t of torch's current stream (0 = legacy default, as the reference uses)."""
    import torch
    try:  # the raw handle without building a torch.cuda.Stream object: this sits in front of every launch
        return ctypes.c_void_p(torch._C._cuda_getCurrentRawStream(torch._C._cuda_getDevice()))
    except AttributeError:
        return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def launch_count():
    return int(lib.ptx_launch_count())


class ptychofft(object):
    """Forward / adjoint ptychography operators on raw device pointers.

    Mirrors class ptychofft of src/include/ptychofft.cuh:6-44.  Unlike the reference
    (single legacy stream, no status checks) every call runs on torch's current
    stream and raises PtxError on failure.
    """

    def __init__(self, ptheta, nz, n, nscan, detector_shape, probe_shape):
        self._h = _vp()
        self._closed = True
        h = _vp()
        check(lib.ptx_create(ctypes.byref(h), int(ptheta), int(nz), int(n), int(nscan),
                             int(detector_shape), int(probe_shape)))
        self._h = h
        self._closed = False

    # read-only attributes (pybind11/ptychofft.cxx:17-22, swig/ptychofft.i:11-17)
    ptheta = property(lambda self: int(lib.ptx_dim(self._h, 0)))
    nz = property(lambda self: int(lib.ptx_dim(self._h, 1)))
    n = property(lambda self: int(lib.ptx_dim(self._h, 2)))
    nscan = property(lambda self: int(lib.ptx_dim(self._h, 3)))
    ndet = property(lambda self: int(lib.ptx_dim(self._h, 4)))
    nprb = property(lambda self: int(lib.ptx_dim(self._h, 5)))

    def fwd(self, g_, f_, scan_, prb_, prb_angle_stride=0):
        """g = FQ f (ptychofft.cu:60-73); arguments are device addresses (ints)."""
        check(lib.ptx_fwd(self._h, _vp(g_), _vp(f_), _vp(scan_), _vp(prb_),
                          int(prb_angle_stride), current_stream()))

    def adj(self, f_, g_, scan_, prb_, flg, prb_angle_stride=0):
        """flg 0: f += Q*F*g ; flg 1: prb += O*F*g (ptychofft.cu:76-88)."""
        check(lib.ptx_adj(self._h, _vp(f_), _vp(g_), _vp(scan_), _vp(prb_),
                          int(prb_angle_stride), int(flg), current_stream()))

    def free(self):
        """Release device memory; idempotent like the reference (ptychofft.cu:49-57)."""
        if self._h:
            check(lib.ptx_free(self._h))

    def __del__(self):
        try:
            if not self._closed and self._h:
                lib.ptx_destroy(self._h)
                self._closed = True
        except Exception:
            pass
