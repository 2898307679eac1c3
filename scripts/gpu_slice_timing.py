"""Time the rank-local operators of the row-partitioned path on ONE GPU (rank 0's slice for R = 1, 2, 4, 8)."""
import json, os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from quadraticprogramsolver_b200 import solver as S, partition
from workloads.problems import config_cfg5
P, q, A, l, u = config_cfg5(seed=1234)
for R in (1, 2, 4, 8):
    P_r, A_r, l_r, u_r, rows, cols = partition.slice_problem(P, A, l, u, 0, R)
    with S.QPB200Solver(P_r, q, A_r, l_r, u_r, numIterations=20) as s:
        out = {"R": R, "nnzP_r": int(P_r.nnz), "nnzA_r": int(A_r.nnz)}
        for which, name in ((1, "A_r"), (4, "H_r")):
            out[name + "_ms"] = round(min(s.time_apply(which, reps=20, flush_l2=True) for _ in range(2)), 4)
            out[name + "_ms_warm"] = round(min(s.time_apply(which, reps=20, flush_l2=False) for _ in range(2)), 4)
        x = np.zeros(P.shape[0]); s.solve(x)
        out["solve20_ms"] = round(s.info["solve_ms"], 1); out["pcg"] = s.info["pcg_iters_total"]
        out["us_per_pcg"] = round(1e3 * s.info["solve_ms"] / s.info["pcg_iters_total"], 1)
    print(json.dumps(out), flush=True)
