"""Diagnostic: dense variants vs the C oracle on padded shapes with the RunTests settings."""
import os, sys, json
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import c_oracle
from quadraticprogramsolver_b200 import solver as S
from workloads.problems import config_cfg3_batch
kw = dict(rho=0.1, adptRho=True, epsAbs=1e-7, epsRel=1e-7, numIterations=50000)
for n, m in [(40, 96), (64, 96), (17, 94), (64, 93)]:
    P, q, A, l, u = config_cfg3_batch(160, n, m, seed=77)
    X0 = np.random.default_rng(3).standard_normal((160, n))
    Xr, fr, ir, _, rc = c_oracle.solve_dense_batch(P, q, A, l, u, x0=X0, **kw)
    out = {"n": n, "m": m}
    res = {}
    for v in ("smem", "regs"):
        X, f, it, _ = S.SolveQuadraticProgramBatch(P, q, A, l, u, X0=X0, denseVariant=v, **kw)
        res[v] = X
        err = np.max(np.abs(X - Xr), axis=1) / (1 + np.max(np.abs(Xr), axis=1))
        b = int(np.argmax(err))
        out[v] = dict(max_rel_err=float(err.max()), worst=b, flags_equal=bool(np.array_equal(f, fr)), it_diff=int(np.max(np.abs(it - ir))),
                      worst_iters=int(it[b]), worst_flag=int(f[b]), n_over_1e6=int((err > 1e-6).sum()))
    out["regs_vs_smem"] = float(np.max(np.abs(res["regs"] - res["smem"])))
    print(json.dumps(out), flush=True)
