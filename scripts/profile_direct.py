"""Target program for ncu: one exact-solve (linSolver = "cholesky") of the RunTests lasso problem at n = 100 (10 200 variables):
the dense factorisation kernels (build_k / gj_pivot / gj_panel / gj_update) and a short admm_kernel<DIRECT> launch."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from quadraticprogramsolver_b200 import solver as S  # noqa: E402
from workloads.problems import GenerateRandomQP, ProblemClass  # noqa: E402

P, q, A, l, u = GenerateRandomQP(ProblemClass.lassoOptimization, 100, seed=1234)
x, flag, info = S.SolveQuadraticProgram(P, q, A, l, u, linSolver="cholesky", numIterations=50, rho=0.1)
print("direct", int(flag), info["iterations"], info["solve_ms"], info["kernel_launches"])
