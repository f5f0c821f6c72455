"""CPU-side tests: C-ABI surface, host logic of the Python mirror, and the CTA-FFT emulation."""
import ctypes
import os
import re
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "ptychofft_b200.h")
LIBPATH = os.path.join(ROOT, "libtike-cufft_b200", "libtike", "cufft", "libptychofft_b200.so")


def _declared_symbols():
    text = open(HEADER).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(ptx_[a-z_0-9]+)\s*\(", text)))


@pytest.fixture(scope="module")
def built():
    if not os.path.exists(LIBPATH):
        sys.path.insert(0, ROOT)
        import __graft_entry__
        __graft_entry__.build()
    return LIBPATH


def test_library_exports_every_declared_symbol(built):
    lib = ctypes.CDLL(built)
    names = _declared_symbols()
    assert len(names) >= 17
    for name in names:
        assert hasattr(lib, name), name


def test_python_binding_table_matches_header(built):
    from libtike.cufft import ptychofft as m
    bound = sorted(n for n, _, _ in m.SYMBOLS)
    assert bound == _declared_symbols()


def test_no_gpu_fails_loudly(built):
    """Without a compute-capability-10 device the constructor raises; nothing falls back to the CPU."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    import libtike.cufft as pt
    from libtike.cufft.ptychofft import PtxError, lib
    assert lib.ptx_device_ok() == 0
    with pytest.raises(PtxError):
        pt.CGPtychoSolver(100, 128, 128, 1, 276, 600)


def test_reference_import_surface(built):
    import libtike.cufft as pt
    from libtike.cufft.ptycho import PtychoCuFFT, CGPtychoSolver
    from libtike.cufft.ptychofft import ptychofft
    assert issubclass(PtychoCuFFT, ptychofft) and issubclass(CGPtychoSolver, PtychoCuFFT)
    for name in ("fwd", "adj", "adj_probe", "fwd_ptycho_batch", "adj_ptycho_batch",
                 "adj_ptycho_batch_prb", "run", "run_batch", "__enter__", "__exit__", "free"):
        assert hasattr(CGPtychoSolver, name), name
    assert hasattr(pt, "__version__")
    # ctor argument order of the reference wrapper (ptycho.py:58) and the native class (cxx:9-16)
    import inspect
    assert list(inspect.signature(PtychoCuFFT.__init__).parameters)[1:] == \
        ["nscan", "probe_shape", "detector_shape", "ntheta", "nz", "n"]
    assert list(inspect.signature(ptychofft.__init__).parameters)[1:] == \
        ["ptheta", "nz", "n", "nscan", "detector_shape", "probe_shape"]
    assert list(inspect.signature(CGPtychoSolver.run).parameters)[1:] == \
        ["data", "psi", "scan", "probe", "piter", "model", "recover_prb", "ortho_prb"]
    # module-level registration function of the reference (ptycho.py:192-193) and its defaults
    assert list(inspect.signature(pt.register_translation_batch).parameters) == \
        ["src_image", "target_image", "upsample_factor", "space"]
    sig = inspect.signature(pt.register_translation_batch)
    assert sig.parameters["upsample_factor"].default == 1 and sig.parameters["space"].default == "real"
    # the reference runs its position-correction block unconditionally (ptycho.py:398-403)
    assert CGPtychoSolver.position_correction is True and CGPtychoSolver.position_upsample == 100


def test_line_search_sqr_matches_reference_semantics(built):
    """ptycho.py:253-281: halve from 1 while the cost increases; give up below 1e-32 with 0."""
    from libtike.cufft.ptycho import CGPtychoSolver
    from oracle.numpy_ptycho import line_search_sqr
    f = lambda x: float(np.sum((x - 3.0) ** 2))
    p1, p2, p3 = np.array([1.0]), np.array([0.0]), np.array([16.0])
    for fn in (CGPtychoSolver.line_search_sqr, line_search_sqr):
        assert fn(f, p1, p2, p3) == 0.25
        assert fn(f, p1, p2, p3 / 4) == 1  # a non-increasing full step is accepted as is
    with pytest.warns(UserWarning):
        assert CGPtychoSolver.line_search_sqr(lambda x: float(x[0]), p1 * 0, p2 * 0 + 1.0, p3 * 0 + 1.0) == 0


def test_cta_fft_emulation():
    """fft_tile.cuh on the CPU: every thread of a CTA run in turn, forward + inverse, 64^2 and 128^2."""
    exe = os.path.join("/tmp", "emu_fft_%d" % os.getpid())
    subprocess.check_call(["g++", "-std=c++17", "-O2", "-I",
                           os.path.join(ROOT, "libtike-cufft_b200", "csrc"),
                           os.path.join(ROOT, "tests", "emu_fft.cpp"), "-o", exe])
    try:
        out = subprocess.run([exe], capture_output=True, text=True, timeout=300)
    finally:
        os.remove(exe)
    assert out.returncode == 0, out.stdout + out.stderr
    assert "EMU OK" in out.stdout
    assert "worst_bank_conflict=1" in out.stdout


def test_workloads_are_seeded():
    import workloads
    a, b = workloads.c2_single_angle(), workloads.c2_single_angle()
    assert np.array_equal(a["scan"], b["scan"]) and np.array_equal(a["psi"], b["psi"])
    s = a["scan"]
    assert s.shape == (1, 1024, 2) and s.min() >= 0 and s.max() <= 512 - 128 - 1
    assert np.all(s != np.trunc(s))  # always sub-pixel
