"""Timing of the exact-solve path (linSolver = "cholesky" on a sparse handle) next to the PCG path: device ms of a
whole solve (factorisation included), iterations, refactorisations.  Writes gpurun_out/r2_direct_probe.jsonl."""
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from quadraticprogramsolver_b200 import solver as S  # noqa: E402
from workloads.problems import GenerateRandomQP, ProblemClass, config_cfg1, config_cfg2  # noqa: E402

RUNTESTS_KW = dict(numIterations=50000, epsAbs=1e-7, epsRel=1e-7, rho=0.1, adptRho=True)
cases = [("cfg1 n=100 defaults", config_cfg1(seed=1234), {}),
         ("cfg1 n=100 RunTests settings", config_cfg1(seed=1234), RUNTESTS_KW),
         ("inequality n=100 m=1000", GenerateRandomQP(ProblemClass.inequalityConstrainedQp, 100, seed=1234), RUNTESTS_KW),
         ("lasso n=100 (10200 vars)", GenerateRandomQP(ProblemClass.lassoOptimization, 100, seed=1234), RUNTESTS_KW),
         ("svm n=100 (10100 vars)", GenerateRandomQP(ProblemClass.supportVectorMachine, 100, seed=1234), RUNTESTS_KW),
         ("cfg2 n=10000 m=20000", config_cfg2(seed=1234), dict(numIterations=2000))]
out = open(os.path.join("gpurun_out", "r2_direct_probe.jsonl"), "w")
for name, (P, q, A, l, u), kw in cases:
    rec = {"case": name, "n": P.shape[0], "m": A.shape[0]}
    for mode, extra in (("cholesky", {}), ("pcg", dict(epsPcg=1e-10))):
        with S.QPB200Solver(P, q, A, l, u, linSolver=mode, **kw, **extra) as s:
            best = None
            for rep in range(3):
                x = np.zeros(P.shape[0])
                flag = s.solve(x)
                if mode == "cholesky" and rep == 0:
                    first = s.info["solve_ms"]           # includes the initial factorisation
                best = s.info["solve_ms"] if best is None else min(best, s.info["solve_ms"])
            rec[mode] = {"flag": int(flag), "iterations": int(s.info["iterations"]), "rho_updates": int(s.info["rho_updates"]),
                         "solve_ms": best, "launches": int(s.info["kernel_launches"]), "pcg_iters": int(s.info["pcg_iters_total"])}
            if mode == "cholesky":
                rec[mode]["first_solve_ms_with_initial_factorisation"] = first
    print(json.dumps(rec), flush=True)
    out.write(json.dumps(rec) + "\n")
