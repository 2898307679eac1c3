"""Latency-bound regime (cfg2, cfg4, cfg1): solve time vs persistent grid size (QPB200_GRID) and tile loader.
One process; the grid override is read by qpb200_create at every call.  Output: one JSON line per combination."""
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from quadraticprogramsolver_b200 import solver as S                      # noqa: E402
from workloads.problems import config_cfg1, config_cfg2, config_cfg4   # noqa: E402

probs = {"cfg2": (config_cfg2(), dict(numIterations=500)),
         "cfg4": (config_cfg4(), dict(numIterations=200)),
         "cfg1": (config_cfg1(), dict(numIterations=2000, epsAbs=1e-12, epsRel=1e-12))}
for name, (prob, kw) in probs.items():
    P, q, A, l, u = prob
    base = None
    for loader in ("tma", "ldg"):
        for g in ("auto", "74", "148", "296", "444"):
            if g == "auto":
                os.environ.pop("QPB200_GRID", None)
            else:
                os.environ["QPB200_GRID"] = g
            with S.QPB200Solver(P, q, A, l, u, spmvLoader=loader, **kw) as s:
                best = None
                for _ in range(3):
                    x = np.zeros(P.shape[0])
                    s.solve(x)
                    best = s.info["solve_ms"] if best is None else min(best, s.info["solve_ms"])
                info = s.info
            if base is None:
                base = x.copy()
            print(json.dumps({"problem": name, "loader": loader, "grid": g, "ms": round(best, 3), "it": info["iterations"],
                              "pcg": info["pcg_iters_total"], "us_per_pcg": round(1e3 * best / max(1, info["pcg_iters_total"]), 3),
                              "max_dx_vs_first": float(np.max(np.abs(x - base)))}), flush=True)
os.environ.pop("QPB200_GRID", None)
