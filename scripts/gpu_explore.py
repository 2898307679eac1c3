"""Exploration on a B200: SpMV loader variants at cfg5 scale, solves of cfg2 / cfg5.  Not a bench."""
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from quadraticprogramsolver_b200 import solver as S  # noqa: E402
from workloads.problems import config_cfg5, config_sparse  # noqa: E402

scale = float(sys.argv[1]) if len(sys.argv) > 1 else 1.0
out = {}
t0 = time.time()
P, q, A, l, u = config_cfg5(seed=1234, scale=scale)
out["gen_s"] = time.time() - t0
out["n"], out["m"], out["nnzP"], out["nnzA"] = P.shape[0], A.shape[0], int(P.nnz), int(A.nnz)
print(out, flush=True)
for loader in ("ldg", "tma", "tma_pipe"):
    t0 = time.time()
    with S.QPB200Solver(P, q, A, l, u, spmvLoader=loader, numIterations=100) as s:
        out[f"{loader}_create_s"] = time.time() - t0
        for which, name in ((1, "A"), (4, "H"), (5, "Hsplit"), (3, "K")):
            ms = s.time_apply(which, reps=20, flush_l2=True)
            gb = s.apply_bytes(which) / 1e9
            out[f"{loader}_{name}_ms"] = ms
            out[f"{loader}_{name}_GBs"] = gb / (ms * 1e-3)
        x = np.zeros(P.shape[0])
        flag = s.solve(x)
        info = dict(s.info)
        info["solve_GBs"] = s.apply_bytes(100) / 1e9 / (info["solve_ms"] * 1e-3)
        out[f"{loader}_solve100"] = info
    print(json.dumps(out, default=float), flush=True)
P, q, A, l, u = config_sparse(10000, 20000, 1e-3, seed=1234)
for loader in ("ldg", "tma"):
    with S.QPB200Solver(P, q, A, l, u, spmvLoader=loader) as s:
        x = np.zeros(P.shape[0])
        s.solve(x)
        x = np.zeros(P.shape[0])
        flag = s.solve(x)
        out[f"cfg2_{loader}"] = dict(s.info, flag=int(flag))
print(json.dumps(out, default=float))
json.dump(out, open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "gpurun_out", "explore.json"), "w"), default=float, indent=1)
