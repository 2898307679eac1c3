"""The parity criterion of the GPU tests (BASELINE.json north_star): same convergence flag, iteration count within
+-2, ||x - x_ref||inf <= tol (1 + ||x_ref||inf) -- plus a *same-trajectory* criterion for the one situation in which
the exit iteration is not a property of the algorithm.

With RunTests.jl:50-53 settings (eps 1e-7, adaptive rho) several generator classes end on the convAdmm test
`|dx|inf <= 1e-9 and |dz|inf <= 1e-9` (SolveQuadraticProgram.jl:34,105) after the iterates have stalled: what is left
of dx is then the rounding / inner-solve noise of the x~ step, and the check at which it dips under 1e-9 changes with
the summation order of a dot product.  Measured on equalityConstrainedQp, n = 10, seed 1235: oracle mode D exits at
iteration 75, the Python oracle (mode J) at 375, the C oracle at 100 / 175 / 225 depending on how its dot products
were blocked, the GPU at 75 or 225 depending on the order in which the grid reduction adds its partials -- all with
the same x to 1e-9.  When (and only when) a test passes `resolve_gpu` / `resolve_ref` and the exit iterations differ,
the criterion becomes: both runs converged (flag 2 or 3), the final solutions agree to `tol`, and at the EARLIER of the
two exit iterations the two iterates agree to `tol` with the same number of rho updates (the run that went on is
re-solved with the iteration cap set to that iteration).  Every case where the exits agree -- all but a handful -- is
held to the strict criterion, and a warning names each case that took the trajectory route.
"""
import warnings

import numpy as np


def _close(a, b, tol):
    err = float(np.max(np.abs(a - b))) if len(a) else 0.0
    return err, err <= tol * (1.0 + (float(np.max(np.abs(b))) if len(b) else 0.0))


def assert_parity(x, flag, info, x_ref, flag_ref, info_ref, tol=1e-6, resolve_gpu=None, resolve_ref=None,
                  rho_updates=False, what=""):
    """info_ref: the oracle's info dict, or just its iteration count.  resolve_*(k) -> (x_k, info_k) re-solve the
    problem with numIterations = k (same other settings).  Returns "strict" or "trajectory"."""
    it = int(info["iterations"])
    it_ref = int(info_ref["iterations"]) if isinstance(info_ref, dict) else int(info_ref)
    if abs(it - it_ref) <= 2 or resolve_gpu is None or resolve_ref is None:
        assert int(flag) == int(flag_ref), f"flag {int(flag)} vs oracle {int(flag_ref)}"
        assert abs(it - it_ref) <= 2, f"iterations {it} vs oracle {it_ref}"
        err, ok = _close(x, x_ref, tol)
        assert ok, f"|x - x_ref|inf = {err:.3e}"
        if rho_updates:
            assert int(info["rho_updates"]) == int(info_ref["rho_updates"]), "rho updates differ"
        return "strict"
    # the exit checks differ: legitimate only if both runs converged to the same point along the same trajectory
    assert int(flag) in (2, 3) and int(flag_ref) in (2, 3), \
        f"exit at {it} (flag {int(flag)}) vs oracle {it_ref} (flag {int(flag_ref)}): one of the runs did not converge"
    err, ok = _close(x, x_ref, tol)
    assert ok, f"exit at {it} vs oracle {it_ref} and the solutions differ: |x - x_ref|inf = {err:.3e}"
    k = min(it, it_ref)
    xg, ig = (x, info) if it == k else resolve_gpu(k)
    xr, ir = (x_ref, info_ref) if it_ref == k else resolve_ref(k)
    err_k, ok = _close(xg, xr, tol)
    assert ok, f"exit at {it} vs oracle {it_ref}; at iteration {k} the iterates differ: {err_k:.3e}"
    if isinstance(ir, dict) and "rho_updates" in ir:
        assert int(ig["rho_updates"]) == int(ir["rho_updates"]), f"rho updates differ at iteration {k}"
    warnings.warn(f"{what} exit checks differ (GPU {it}, oracle {it_ref}; stop test fired on rounding noise): same trajectory at "
                  f"iteration {k} (|dx| {err_k:.1e}), same solution (|dx| {err:.1e})")
    return "trajectory"
