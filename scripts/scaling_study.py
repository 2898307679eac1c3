"""Effect of Ruiz equilibration (numItrScaling) on time-to-convergence, GPU path only.
usage: python scripts/scaling_study.py [out.json]"""
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from quadraticprogramsolver_b200 import solver as S
from workloads import problems   # noqa: E402

cases = {
    "cfg1 (n=100, m=50)": problems.config_cfg1(seed=1234),
    "cfg1 badly scaled (variables 1e-2..1e2, rows 1e-3..1e3)": problems.badly_scaled(problems.config_cfg1(seed=1234), seed=0),
    "cfg4 constrained LS, scale 0.2 (n=10000)": problems.config_cfg4(seed=1234, scale=0.2),
    "cfg4 scale 0.2, badly scaled (variables 1e-1..1e1, rows 1e-2..1e2)":
        problems.badly_scaled(problems.config_cfg4(seed=1234, scale=0.2), seed=2, var_decades=1.0, con_decades=2.0),
    "cfg2 (n=1e4, m=2e4)": problems.config_cfg2(seed=1234),
}
rows = []
for name, (P, q, A, l, u) in cases.items():
    for k in (0, 10):
        kw = dict(rho=0.1, adptRho=True, numIterations=20000, numItrScaling=k)
        x = np.zeros(P.shape[0])
        with S.QPB200Solver(P, q, A, l, u, **kw) as s:
            s.solve(x)                       # warm-up
            x[:] = 0.0
            flag = s.solve(x, want_zy=True)
            i = s.info
            Ax = A @ x
            viol = float(np.max(np.maximum(np.maximum(l - Ax, Ax - u), 0.0))) if A.shape[0] else 0.0
            stat = float(np.max(np.abs(P @ x + q + A.T @ i["y"])))
            rows.append({"case": name, "numItrScaling": k, "flag": int(flag), "iterations": int(i["iterations"]),
                         "cg_iterations": int(i["pcg_iters_total"]), "rho_updates": int(i["rho_updates"]),
                         "solve_ms": round(float(i["solve_ms"]), 2), "setup_ms": round(float(i["setup_ms"]), 2),
                         "primal_violation": viol, "stationarity": stat})
            print(json.dumps(rows[-1]), flush=True)
if len(sys.argv) > 1:
    json.dump(rows, open(sys.argv[1], "w"), indent=1)
