import numpy as np


def rel_l2(a, b):
    a = np.asarray(a)
    b = np.asarray(b)
    return float(np.linalg.norm((a - b).ravel()) / np.linalg.norm(b.ravel()))


def ReplaySolver(*args, **kwargs):
    """CGPtychoSolver whose line-search DECISIONS can be replayed from a list (`forced_steps`, raw
    line_search_sqr results consumed in call order).  Test instrumentation only: the costs are still
    evaluated and logged by the product code, only the accept/halve decision is overridden, so that
    two fp32 implementations can be compared along the same trajectory (near-tie decisions are
    rounding noise).  With `forced_steps = None` it is the product solver."""
    import libtike.cufft as pt

    class _Replay(pt.CGPtychoSolver):
        forced_steps = None
        log_shifts = True  # the parity tests compare the position-correction shifts step by step

        @property
        def device_line_search(self):
            # replayed decisions are taken by `_ls_decide` on the host; free-running it is the product path
            return self.forced_steps is None

        def _ls_begin(self):
            self._forced = self.forced_steps.pop(0) if self.forced_steps else None

        def _ls_decide(self, c0, c, K):
            f = self._forced
            if f is None:
                return super()._ls_decide(c0, c, K)
            if f == 0 or f >= 2.0 ** -(c0 + K - 1):
                return f
            return None

    return _Replay(*args, **kwargs)
