"""run_batch(piter=8) on 8 angles of the c4 shape, repeated: wall time of every call and of every run() inside it,
to see where the slow calls lose their time.   usage: python tools/runbatch_trace.py [reps=8]"""
import contextlib, gc, io, os, sys, time
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "libtike-cufft_b200")]
import workloads, libtike.cufft as pt
reps = int(sys.argv[1]) if len(sys.argv) > 1 else 8
nd, A = 256, 8
w = workloads.synth_angles(A, 4 * nd, 4 * nd, nd, nd, 32, 1)
psi_t, scan, probe = (torch.from_numpy(w[k]).cuda() for k in ("psi", "scan", "probe"))
with pt.CGPtychoSolver(1024, nd, nd, 1, 4 * nd, 4 * nd) as s1, contextlib.redirect_stdout(io.StringIO()):
    data = torch.cat([(s1.fwd(psi_t[t:t + 1], scan[t:t + 1], probe[t:t + 1, 0].contiguous()).abs() ** 2) for t in range(A)])
    h = {"data": data.cpu().numpy(), "psi": np.ones_like(w["psi"]), "scan": w["scan"], "probe": w["probe"]}
    del data
    inner = s1.run
    marks = []

    def timed_run(*a, **k):
        t0 = time.perf_counter()
        r = inner(*a, **k)
        marks.append((t0, time.perf_counter()))
        return r
    s1.run = timed_run
    s1.run_batch(h["data"][:1], h["psi"][:1], h["scan"][:1], h["probe"][:1], piter=2, recover_prb=True)
    gc_ms, gc_t0 = [0.0, 0], [0.0]

    def gc_cb(phase, info):
        if phase == "start":
            gc_t0[0] = time.perf_counter()
        else:
            gc_ms[0] += 1e3 * (time.perf_counter() - gc_t0[0])
            gc_ms[1] += 1 if info["generation"] == 2 else 0
    gc.callbacks.append(gc_cb)
    for rep in range(2 * reps):
        if rep == reps:
            gc.collect()
            gc.disable()
            sys.stderr.write("-- cyclic GC disabled from here on\n")
        gc_ms[0], gc_ms[1] = 0.0, 0
        marks.clear()
        torch.cuda.synchronize(); t0 = time.perf_counter()
        s1.run_batch(h["data"], h["psi"], h["scan"], h["probe"], piter=8, recover_prb=True)
        torch.cuda.synchronize(); t1 = time.perf_counter()
        runs = ["%.1f" % (1e3 * (b - a)) for a, b in marks]
        gaps = ["%.1f" % (1e3 * (marks[i + 1][0] - marks[i][1])) for i in range(len(marks) - 1)]
        sys.stderr.write("call %d: %.1f ms = %.1f angle-it/s; GC %.1f ms (%d full); %.2f GB allocated; first run starts at %.1f ms; run() ms %s; gaps ms %s\n" % (
            rep, 1e3 * (t1 - t0), 64 / (t1 - t0), gc_ms[0], gc_ms[1], torch.cuda.memory_allocated() / 1e9, 1e3 * (marks[0][0] - t0), runs, gaps))
