"""A/B of stand-alone SpMV + 100-iteration solve for one libqpb200 build (select with QPB200_LIB).
usage: QPB200_LIB=... [QPB200_CTAS_PER_SM=3] gpu_spmv_ab.py tag [scale]"""
import json, os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from quadraticprogramsolver_b200 import solver as S
from workloads.problems import config_cfg5
tag = sys.argv[1]
scale = float(sys.argv[2]) if len(sys.argv) > 2 else 1.0
P, q, A, l, u = config_cfg5(seed=1234, scale=scale)
out = {"tag": tag}
for loader in ("tma",):
    with S.QPB200Solver(P, q, A, l, u, spmvLoader=loader, numIterations=100) as s:
        for which, name in ((1, "A"), (4, "H")):
            ms = min(s.time_apply(which, reps=20, flush_l2=True) for _ in range(2))
            out[f"{loader}_{name}_ms"] = round(ms, 4)
            out[f"{loader}_{name}_GBs"] = round(s.apply_bytes(which) / 1e6 / ms, 1)
        x = np.zeros(P.shape[0]); s.solve(x)
        x = np.zeros(P.shape[0]); s.solve(x)
        out[f"{loader}_solve_ms"] = round(s.info["solve_ms"], 1)
        out[f"{loader}_pcg"] = s.info["pcg_iters_total"]
print(json.dumps(out), flush=True)
