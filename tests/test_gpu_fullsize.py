"""Full-size GPU parity: every detector-size plan with MORE PATTERNS THAN THE PERSISTENT GRID.

The kernels run one CTA per SM (four for the 64^2 plan) in a persistent loop over patterns.  With
fewer patterns than CTAs each CTA sees at most one pattern, and everything that only happens
between two patterns of the same CTA -- the TMA prefetch of the next measured-data tile and object
patch, mbarrier phase flips, reuse of the staging frame and of the thread-private accumulators, the
flush of the probe accumulators at an angle boundary, and at 64^2 the scatter loop that carries the
NEXT pattern's gather (scatter_gather_impl) -- never executes.  These tests run BASELINE.json's
configurations C2-C5 at (or near) their stated sizes against the reference's own CUDA/cuFFT code
compiled where it lies (oracle/_ref) on the same GPU:

  operators   rel. L2 <= 1e-5 on fwd / adj / adj_probe        (kernels.cu:19-107, ptychofft.cu:60-88)
  gradient    the host-array fused gradient against fwd -> residual -> adj of the reference
  solver      3 CG iterations exactly as the reference runs them (position correction ON, probe
              recovery, its line-search decisions replayed) <= 1e-4 on psi and probe
              (tests/test_modes.py:18-60, tests/test_fsc.py:74-120, ptycho.py:283-488)
"""
import numpy as np
import pytest
import torch

import workloads
from oracle import numpy_ptycho as O
from oracle import ref_gpu
from util import rel_l2, ReplaySolver

pytestmark = [pytest.mark.gpu,
              pytest.mark.skipif(not ref_gpu.available(), reason="oracle/_ref not built")]
TOL_OP = 1e-5
TOL_CG = 1e-4


def _pt():
    import libtike.cufft as pt
    return pt


def _cuda(x):
    return torch.from_numpy(np.ascontiguousarray(x)).cuda()


def _positions(rng, T, S, nz, n, nprb, skip_every=0):
    """Uniformly random sub-pixel positions over the whole valid domain of the reference
    (R + P <= nz - 1, C + P <= n - 1), the largest valid origin included; every `skip_every`-th
    position is flagged -1 (skipped, tests/test_fsc.py:16-17)."""
    scan = np.stack([rng.uniform(0, nz - nprb - 1.001, (T, S)),
                     rng.uniform(0, n - nprb - 1.001, (T, S))], axis=-1).astype(np.float32)
    scan[:, 1] = (nz - nprb - 1, n - nprb - 1)        # largest valid origin, integer
    scan[:, 2] = (nz - nprb - 1.25, n - nprb - 1.5)   # ... and fractional
    scan[:, 3] = (0.0, 0.0)
    if skip_every:
        scan[:, 5::skip_every] = -1.0
        scan[0, 0] = -1.0           # first pattern of the batch
        scan[T - 1, S - 1] = (-1.0, 7.5)  # last pattern of the batch, only the row negative
    return scan


# (ndet, nprb, T, S, nz, n, skip_every): every case has T * S > grid (148 CTAs; 592 at 64^2)
OPS = [
    (64, 64, 1, 2048, 256, 256, 0),      # Plan<6>, interior full windows only: scatter_gather_impl on every pattern
    (64, 64, 2, 1024, 200, 220, 37),     # Plan<6>, skips between interior patterns, two angles
    (64, 40, 2, 1024, 200, 220, 41),     # Plan<6>, probe window smaller than the detector
    (128, 128, 2, 1024, 512, 512, 53),   # Plan<7>, C2 shape
    (128, 96, 1, 600, 276, 600, 29),     # Plan<7>, window, C1/C3 object shape
    (256, 256, 1, 1024, 1024, 1024, 0),  # Plan<8>, C4 shape
    (256, 256, 2, 512, 1024, 1024, 61),  # Plan<8>, two angles, skips
    (256, 200, 1, 400, 600, 640, 31),    # Plan<8>, window
    (512, 512, 1, 320, 1200, 1100, 43),  # Plan<9>
]


@pytest.mark.parametrize("cfg", OPS, ids=lambda c: "N%d_P%d_T%d_S%d" % c[:4])
def test_operators_full_size_vs_compiled_reference(cfg):
    pt = _pt()
    ndet, nprb, T, S, nz, n, skip_every = cfg
    w = workloads.synth_angles(T, nz, n, ndet, nprb, 1, 1, seed0=40 + ndet)
    rng = np.random.default_rng(ndet * 7 + S)
    scan = _cuda(_positions(rng, T, S, nz, n, nprb, skip_every))
    psi = _cuda(w["psi"])
    prb = w["probe"][:, 0].copy()
    prb[T - 1] *= 0.75 - 0.5j  # angles must not share a probe
    prb = _cuda(prb)
    with pt.PtychoCuFFT(S, nprb, ndet, T, nz, n) as slv, \
            ref_gpu.RefPtychoFFT(S, nprb, ndet, T, nz, n) as ref:
        g_ref = ref.fwd(psi, scan, prb)
        g = slv.fwd(psi, scan, prb)
        nrm = torch.linalg.norm(g_ref)
        e_fwd = float(torch.linalg.norm(g - g_ref) / nrm)
        del g
        # adjoints of a far field that is NOT the forward image of psi (generic input)
        g_in = g_ref * torch.exp(1j * 0.3 * torch.arange(ndet, device="cuda", dtype=torch.float32))[None, None, None, :]
        g_in = g_in.contiguous()
        f_ref = ref.adj(g_in, scan, prb)
        f = slv.adj(g_in, scan, prb)
        e_adj = float(torch.linalg.norm(f - f_ref) / torch.linalg.norm(f_ref))
        q_ref = ref.adj_probe(g_in, scan, psi)
        q = slv.adj_probe(g_in, scan, psi)
        e_prb = float(torch.linalg.norm(q - q_ref) / torch.linalg.norm(q_ref))
    print("full-size operators %s: fwd %.2e adj %.2e adj_probe %.2e" % (cfg, e_fwd, e_adj, e_prb))
    assert e_fwd < TOL_OP and e_adj < TOL_OP and e_prb < TOL_OP


def test_plan6_zero_extension_between_interior_patterns():
    """64^2, more patterns than CTAs, with windows that CROSS the object edge (outside the
    reference's valid domain, where it reads out of bounds: SURVEY Q11 -> zero extension) mixed
    between interior ones, so that consecutive patterns of a CTA alternate between the fused
    scatter+gather loop and the predicated path.  Checker: the NumPy oracle on a zero-padded object."""
    pt = _pt()
    nz, n, P, S = 150, 170, 64, 1400
    rng = np.random.default_rng(64)
    psi = (rng.random((1, nz, n)) + 1j * rng.random((1, nz, n))).astype(np.complex64)
    prb = (rng.random((1, P, P)) + 1j * rng.random((1, P, P))).astype(np.complex64)
    scan = _positions(rng, 1, S, nz, n, P, 0)
    scan[0, 7::3, 0] = rng.uniform(nz - P, nz - P + 0.9, scan[0, 7::3, 0].shape)   # last row of taps outside
    scan[0, 8::5, 1] = rng.uniform(n - P, n - P + 0.9, scan[0, 8::5, 1].shape)     # last column outside
    scan[0, 11::97] = -1.0
    big = np.zeros((1, nz + 2, n + 2), dtype=np.complex64)
    big[:, :nz, :n] = psi
    g0 = O.fwd(big, scan, prb, P)
    with pt.PtychoCuFFT(S, P, P, 1, nz, n) as slv:
        assert rel_l2(slv.fwd_ptycho_batch(psi, scan, prb), g0) < TOL_OP
        f0 = O.adj(g0, scan, prb, nz + 2, n + 2)[:, :nz, :n]
        assert rel_l2(slv.adj_ptycho_batch(g0, scan, prb), f0) < TOL_OP
    # the same mix through the fused gradient kernel (where scatter_gather_impl lives)
    data = (np.abs(g0) ** 2 * rng.uniform(0.6, 1.4, g0.shape)).astype(np.float32)
    psi1 = (psi * (0.7 + 0.2j)).astype(np.complex64)
    big1 = np.zeros_like(big)
    big1[:, :nz, :n] = psi1
    f = O.fwd(big1, scan, prb, P)
    r = f - np.sqrt(data) * f / (np.abs(f) + np.float32(1e-32))
    want = O.adj(r.astype(np.complex64), scan, prb, nz + 2, n + 2)[:, :nz, :n]
    with pt.CGPtychoSolver(S, P, P, 1, nz, n) as slv:
        got = slv.grad_ptycho_batch(data, psi1, scan, prb[:, None], model="gaussian")
    assert rel_l2(got, want) < 2e-5


def _ref_gradient(ref, psi, scan, probe, data, model):
    """sum_k Q_k* F* [F Q_k psi (1 - sqrt(d)/sqrt(I))] (gaussian; 1 - d/I poisson) with the
    reference's operators and the CuPy statements of ptycho.py:351-363 in torch."""
    M = probe.shape[1]
    far = [ref.fwd(psi, scan, probe[:, k].contiguous()) for k in range(M)]
    inten = sum(torch.abs(f) ** 2 for f in far)
    out = None
    for k, f in enumerate(far):
        if model == "gaussian":
            r = f - torch.sqrt(data) * f / (torch.sqrt(inten) + 1e-32)
        else:
            r = f - data * f / (inten + 1e-32)
        g = ref.adj(r, scan, probe[:, k].contiguous())
        out = g if out is None else out + g
    return out


@pytest.mark.parametrize("ndet,nmodes,model,S,nz,n", [
    (256, 2, "gaussian", 1024, 1024, 1024),   # chunks that fill the grid, M > 1 (accumulators, frames)
    (128, 1, "poisson", 1024, 512, 512),
    (64, 3, "gaussian", 2048, 256, 256),
])
def test_grad_ptycho_batch_full_size(ndet, nmodes, model, S, nz, n):
    """The pipelined host-array gradient with chunks large enough that the kernels of consecutive
    chunks WOULD overlap if their streams were not ordered (they share the plan's per-CTA scratch):
    three angles through a ptheta = 1 plan, against the reference's operators on the same GPU."""
    pt = _pt()
    T = 3
    w = workloads.synth_angles(T, nz, n, ndet, ndet, 1, nmodes, seed0=77)
    rng = np.random.default_rng(S + ndet)
    scan = _positions(rng, T, S, nz, n, ndet, 67)
    psi, probe = w["psi"], w["probe"].copy()
    for k in range(nmodes):
        probe[:, k] *= (1.0 - 0.3 * k)
    psi1 = (psi * (0.8 + 0.3j)).astype(np.complex64)
    data = np.empty((T, S, ndet, ndet), dtype=np.float32)
    want = np.empty_like(psi1)
    with ref_gpu.RefPtychoFFT(S, ndet, ndet, 1, nz, n) as ref:
        for t in range(T):
            p_t, s_t = _cuda(probe[t:t + 1]), _cuda(scan[t:t + 1])
            far = [ref.fwd(_cuda(psi[t:t + 1]), s_t, p_t[:, k].contiguous()) for k in range(nmodes)]
            d = sum(torch.abs(f) ** 2 for f in far)
            d = d * torch.from_numpy(rng.uniform(0.7, 1.3, (1, S, 1, 1)).astype(np.float32)).cuda()
            data[t] = d.cpu().numpy()[0]
            want[t] = _ref_gradient(ref, _cuda(psi1[t:t + 1]), s_t, p_t, d, model).cpu().numpy()[0]
    with pt.CGPtychoSolver(S, ndet, ndet, 1, nz, n) as slv:
        got = slv.grad_ptycho_batch(data, psi1, scan, probe, model=model)
        got2 = slv.grad_ptycho_batch(data, psi1, scan, probe, model=model)
    e = [rel_l2(got[t], want[t]) for t in range(T)] + [rel_l2(got2[t], want[t]) for t in range(T)]
    print("full-size grad_ptycho_batch %s: %s" % ((ndet, nmodes, model), ["%.1e" % x for x in e]))
    assert max(e) < 2e-5


def _cg_case(name):
    """(data, psi0, scan, probe0, ndet, model) of one BASELINE.json configuration at stated size;
    the measured data come from the reference's own forward operator."""
    if name == "c3":      # tests/test_modes.py:18-60 with all five modes, 1100 positions
        c = workloads.c3_modes(5, 1100)
        psi, scan, probe, init, model, counts = c["psi"], c["scan"], c["probe"], c["probe_init"], "gaussian", 0
    elif name == "c4":    # 1024^2 slice, 256^2 detector, 1024 positions
        w = workloads.c4_catalyst(1)
        psi, scan, probe, model, counts = w["psi"], w["scan"], w["probe"], "gaussian", 0
        init = probe * (0.9 + 0.1j)
    elif name == "c4m2":  # ... two probe modes
        w = workloads.c4_catalyst(1, nmodes=2)
        psi, scan, probe, model, counts = w["psi"], w["scan"], w["probe"].copy(), "gaussian", 0
        probe[:, 1] *= 0.5
        init = probe * (0.9 + 0.1j)
    else:                 # C5: detector sweep, Poisson counts (mean ~100), sub-pixel positions
        ndet = int(name[2:])
        side = 32 if ndet < 512 else 18
        w = workloads.c5_sweep(ndet, nside=side) if ndet < 512 else \
            workloads.synth_angles(1, 1200, 1100, 512, 512, side, 1)
        psi, scan, probe, model, counts = w["psi"], w["scan"], w["probe"], "poisson", 100.0
        init = probe * (0.9 + 0.1j)
    ndet = probe.shape[-1]
    S = scan.shape[1]
    nz, n = psi.shape[1:]
    with ref_gpu.RefPtychoFFT(S, ndet, ndet, 1, nz, n) as ref:
        d = sum(torch.abs(ref.fwd(_cuda(psi), _cuda(scan), _cuda(probe[:, k]))) ** 2
                for k in range(probe.shape[1]))
        if counts:
            d = torch.poisson(d * (counts / d.mean()), generator=torch.Generator("cuda").manual_seed(5))
        data = d.cpu().numpy()
    return data, np.ones_like(psi), scan, np.ascontiguousarray(init).astype(np.complex64), ndet, model


@pytest.mark.parametrize("name", ["c3", "c4", "c4m2", "c564", "c5128", "c5256", "c5512"])
def test_cg_full_size_vs_reference(name):
    """Three CG iterations of `run` as the reference executes it (position correction on, probe
    recovery), line-search decisions replayed, against the reference's cuFFT operators + restated
    solver on the same GPU."""
    data, psi0, scan, prb0, ndet, model = _cg_case(name)
    S = scan.shape[1]
    nz, n = psi0.shape[1:]
    piter = 3
    with ref_gpu.RefCGPtychoSolver(S, ndet, ndet, 1, nz, n) as ref:
        ref.position_correction = True
        ref.shift_log = []
        want = ref.run_batch(data, psi0, scan, prb0, piter=piter, model=model, recover_prb=True,
                             verbose=False)
        steps = [t[2] for t in ref.last_trials]
        rlog = ref.shift_log
    torch.cuda.empty_cache()
    with ReplaySolver(S, ndet, ndet, 1, nz, n) as slv:
        assert slv.position_correction is True  # the solver's default = the reference's behaviour
        slv.forced_steps = list(steps)
        got = slv.run_batch(data, psi0, scan, prb0, piter=piter, model=model, recover_prb=True)
        glog = [x.cpu().numpy() for x in slv.shift_log]
    nbad = sum(int((np.abs(a - b).max(axis=1) > 0).sum()) for a, b in zip(glog, rlog))
    worst = max(float(np.abs(a - b).max()) for a, b in zip(glog, rlog))
    e_psi, e_prb = rel_l2(got["psi"], want["psi"]), rel_l2(got["probe"], want["probe"])
    print("full-size CG %s (%d modes, %d positions, %d^2, %s): psi %.2e probe %.2e; %d of %d shifts "
          "differ, worst %.3f px" % (name, prb0.shape[1], S, ndet, model, e_psi, e_prb, nbad,
                                     S * len(rlog), worst))
    assert len(glog) == len(rlog) == piter - 1
    # shifts are argmax picks on a 0.01 px grid of a correlation peak that is nearly flat at that
    # scale (the corrections themselves are <= 0.02 px here): a rounding-level difference moves a pick
    # by one grid step on a few per cent of the positions, never by more
    assert worst <= 0.0100001 and nbad <= max(1, S * len(rlog) // 25)
    if max(e_psi, e_prb) >= TOL_CG:
        # Poisson counts on a large detector: d F / (|F|^2 + 1e-32) at the many weak-signal pixels
        # that recorded a photon amplifies the fp32 rounding of ANY FFT (SURVEY.md Q1 note).  The float64
        # GPU restatement of the same statements, same replayed decisions, is the referee: the fused
        # solver must be as close to the exact trajectory as the reference's own cuFFT run is.
        with ref_gpu.F64CGPtychoSolver(S, ndet, ndet, 1, nz, n) as ex:
            ex.position_correction = True
            ex.forced_steps = list(steps)
            exact = ex.run_batch(data, psi0, scan, prb0, piter=piter, model=model, recover_prb=True,
                                 verbose=False)
        # ... and the reference does not reproduce itself here (unordered atomics): a second run of its
        # own cuFFT path, same inputs and decisions, tells how far two "correct" fp32 runs lie apart
        with ref_gpu.RefCGPtychoSolver(S, ndet, ndet, 1, nz, n) as ref2:
            ref2.position_correction = True
            ref2.forced_steps = list(steps)
            want2 = ref2.run_batch(data, psi0, scan, prb0, piter=piter, model=model, recover_prb=True,
                                   verbose=False)
        for k in ("psi", "probe"):
            e_ref = max(rel_l2(want[k], exact[k]), rel_l2(want2[k], exact[k]), rel_l2(want2[k], want[k]))
            e_got = rel_l2(got[k], exact[k])
            print("   %s: reference vs f64 %.2e / %.2e, reference vs its re-run %.2e; fused vs f64 %.2e" % (
                k, rel_l2(want[k], exact[k]), rel_l2(want2[k], exact[k]), rel_l2(want2[k], want[k]), e_got))
            # factor 3 on a single draw: see tests/test_gpu_cg.py::_assert_parity for the measurements
            assert e_got < max(3 * e_ref, TOL_CG), (k, e_got, e_ref)

def test_pipelined_kernel_parity_subprocess():
    """The opt-in warp-specialised 128^2 object-gradient kernel (PTX_PIPE=1, csrc/ptycho_pipe.cuh): the
    same full-size gradient check as above (interior, edge and skipped positions, two probe modes so
    that the multi-mode variant runs too), in a fresh interpreter because the switch is read once."""
    import os
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    code = (
        "import sys; sys.path[:0] = [%r, %r, %r]\n"
        "import test_gpu_fullsize as T\n"
        "T.test_grad_ptycho_batch_full_size(128, 1, 'poisson', 1024, 512, 512)\n"
        "T.test_grad_ptycho_batch_full_size(128, 2, 'gaussian', 700, 300, 320)\n"
        "from libtike.cufft.ptychofft import lib\n"
        "print('PIPE OK')\n" % (root, os.path.join(root, "libtike-cufft_b200"), os.path.join(root, "tests")))
    p = subprocess.run([sys.executable, "-c", code], env=dict(os.environ, PTX_PIPE="1"), capture_output=True,
                       text=True, timeout=900)
    assert p.returncode == 0 and "PIPE OK" in p.stdout, p.stdout[-2000:] + p.stderr[-3000:]
