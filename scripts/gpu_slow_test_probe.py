"""Where do the 186 s of tests/test_gpu_direct.py::test_exact_solve_agrees_with_tight_pcg_and_reuses_the_factor go?
Times the legs of that test separately (exact-solve handle: create, first and second solve; tight-PCG solve; oracle mode D)."""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import qp_oracle  # noqa: E402
from quadraticprogramsolver_b200 import solver as S  # noqa: E402
from workloads.problems import config_sparse  # noqa: E402


def lap(what, t0, extra=""):
    print(f"{what}: {time.time() - t0:.2f} s {extra}", flush=True)
    return time.time()


t = time.time()
P, q, A, l, u = config_sparse(1200, 1800, 5e-3, seed=5)
t = lap("generate", t)
kw = dict(numIterations=1000, rho=0.1, adptRho=True)
with S.QPB200Solver(P, q, A, l, u, linSolver="cholesky", **kw) as s:
    t = lap("create (cholesky)", t)
    x1 = np.zeros(P.shape[0])
    s.solve(x1)
    i = dict(s.info)
    t = lap("solve 1", t, f"iterations {i['iterations']} rho_updates {i['rho_updates']} solve_ms {i['solve_ms']:.1f} launches {i['kernel_launches']}")
    x2 = np.zeros(P.shape[0])
    s.solve(x2)
    i = dict(s.info)
    t = lap("solve 2", t, f"solve_ms {i['solve_ms']:.1f} launches {i['kernel_launches']}")
for eps in (1e-10, 1e-12):
    xp, fp, ip = S.SolveQuadraticProgram(P, q, A, l, u, epsPcg=eps, **kw)
    t = lap(f"tight PCG epsPcg={eps:g}", t, f"iterations {ip['iterations']} pcg_iters {ip['pcg_iters_total']} solve_ms {ip['solve_ms']:.1f}")
xr, fr, ir = qp_oracle.solve(P, q, A, l, u, mode="D", **kw)
t = lap("oracle mode D", t)
