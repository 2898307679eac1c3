"""Small end-to-end run of every kernel family for compute-sanitizer (memcheck / racecheck)."""
import os, sys
import numpy as np
import scipy.sparse as sp
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from quadraticprogramsolver_b200 import solver as S
from workloads.problems import config_cfg1, config_cfg3_batch, config_sparse

P, q, A, l, u = config_cfg1()
for loader in ("tma", "ldg", "tma_pipe"):
    x, flag, info = S.SolveQuadraticProgram(P, q, A, l, u, spmvLoader=loader, numIterations=50, adptRho=True, rho=0.1)
    print(loader, int(flag), info["iterations"], info["pcg_iters_total"])
P, q, A, l, u = config_sparse(3000, 6000, 2e-3, seed=5)
with S.QPB200Solver(P, q, A, l, u, numIterations=25) as s:
    x = np.zeros(3000); s.solve(x)
    y = s.apply(3, np.ones(3000))
    print("sparse", s.info["iterations"], float(np.abs(y).max()))
rng = np.random.default_rng(0)
n = 1500
M = rng.standard_normal((n, n)) / np.sqrt(n)
Pd = sp.csc_matrix(M.T @ M + 0.01 * np.eye(n)); Ad = sp.csc_matrix(rng.standard_normal((40, n)))
with S.QPB200Solver(Pd, rng.standard_normal(n), Ad, -np.ones(40), np.ones(40), numIterations=3) as s:   # rows longer than a tile
    x = np.zeros(n); s.solve(x); print("longrows", s.info["pcg_iters_total"])
Pb, qb, Ab, lb, ub = config_cfg3_batch(12, 64, 96, seed=1)
X, flags, iters, info = S.SolveQuadraticProgramBatch(Pb, qb, Ab, lb, ub, numIterations=100, adptRho=True, rho=0.1)
print("dense", flags.tolist(), iters.tolist())
Pb, qb, Ab, lb, ub = config_cfg3_batch(5, 30, 45, seed=2)
X, flags, iters, info = S.SolveQuadraticProgramBatch(Pb, qb, Ab, lb, ub, numIterations=50, unblockedCholesky=True)
print("dense padded", flags.tolist(), iters.tolist())
