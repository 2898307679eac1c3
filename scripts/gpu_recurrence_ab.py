"""A/B of the two (P)CG recurrences of admm_kernel on small and mid-size problems (latency-bound regime).
usage: gpu_recurrence_ab.py tag"""
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from quadraticprogramsolver_b200 import solver as S  # noqa: E402
from workloads import problems  # noqa: E402

tag = sys.argv[1]
cases = {"cfg1": lambda: problems.config_cfg1(1234), "cfg2": lambda: problems.config_cfg2(1234),
         "cfg4": lambda: problems.config_cfg4(1234), "cfg5_s0.1": lambda: problems.config_cfg5(1234, scale=0.1)}
for name, gen in cases.items():
    P, q, A, l, u = gen()
    for recur in ("one_reduction", "standard"):
        with S.QPB200Solver(P, q, A, l, u, numIterations=200, cgRecurrence=recur) as s:
            best = None
            for _ in range(3):
                x = np.zeros(P.shape[0]); s.solve(x)
                if best is None or s.info["solve_ms"] < best["solve_ms"]:
                    best = dict(s.info)
            print(json.dumps({"tag": tag, "case": name, "n": P.shape[0], "m": A.shape[0], "nnz": int(P.nnz + 2 * A.nnz), "recur": recur,
                              "solve_ms": round(best["solve_ms"], 3), "iters": best["iterations"], "pcg": best["pcg_iters_total"],
                              "us_per_cg": round(1e3 * best["solve_ms"] / max(1, best["pcg_iters_total"]), 2)}), flush=True)
