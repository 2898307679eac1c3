"""Stand-alone SpMV roofline of the tile engine on matrices WITH and WITHOUT column locality (one libqpb200 build,
select with QPB200_LIB): cfg5 (uniformly random columns), cfg4 (constrained least squares, P = A'A), banded (cfg5's sizes
and non-zeros per row, columns within +-256 of the diagonal).  Matrices are generated once and cached as .npz under /tmp.
usage: gpu_spmv_friendly.py tag [workloads, comma separated] [loaders, comma separated]"""
import json
import os
import sys
import time

import numpy as np
import scipy.sparse as sp

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from quadraticprogramsolver_b200 import solver as S  # noqa: E402
from workloads import problems  # noqa: E402

tag = sys.argv[1]
workloads = (sys.argv[2] if len(sys.argv) > 2 else "cfg5,cfg4,banded").split(",")
loaders = (sys.argv[3] if len(sys.argv) > 3 else "tma").split(",")
peak = 6551.7
try:
    peak = float(json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")))["hbm_gbs"])
except Exception:
    pass


def get(name):
    path = f"/tmp/qpb_{name}.npz"
    if os.path.exists(path):
        d = np.load(path)
        P = sp.csc_matrix((d["Pv"], d["Pi"], d["Pp"]), shape=(int(d["n"]), int(d["n"])))
        A = sp.csc_matrix((d["Av"], d["Ai"], d["Ap"]), shape=(int(d["m"]), int(d["n"])))
        return P, d["q"], A, d["l"], d["u"]
    t0 = time.time()
    if name == "cfg5":
        prob = problems.config_cfg5(seed=1234)
    elif name == "cfg4":
        prob = problems.config_cfg4(seed=1234)
    elif name == "cfg4x8":
        prob = problems.config_cfg4(seed=1234, scale=8.0)
    elif name == "banded":
        prob = problems.config_banded()
    else:
        raise SystemExit(name)
    P, q, A, l, u = prob
    P = sp.csc_matrix(P); A = sp.csc_matrix(A)
    np.savez(path, Pv=P.data, Pi=P.indices, Pp=P.indptr, Av=A.data, Ai=A.indices, Ap=A.indptr, q=q, l=l, u=u,
             n=P.shape[0], m=A.shape[0])
    print(f"# generated {name} in {time.time() - t0:.1f} s", file=sys.stderr)
    return P, q, A, l, u


for name in workloads:
    P, q, A, l, u = get(name)
    for loader in loaders:
        out = {"tag": tag, "workload": name, "loader": loader, "n": P.shape[0], "m": A.shape[0], "nnzP": int(P.nnz),
               "nnzA": int(A.nnz)}
        out["recur"] = os.environ.get("QPB_RECUR", "auto")
        with S.QPB200Solver(P, q, A, l, u, spmvLoader=loader, numIterations=25, cgRecurrence=out["recur"]) as s:
            for which, nm in ((1, "A"), (4, "H")):
                ms = min(s.time_apply(which, reps=20, flush_l2=True) for _ in range(2))
                gb = s.apply_bytes(which) / 1e9
                out[f"{nm}_ms"] = round(ms, 4)
                out[f"{nm}_GBs"] = round(gb / (ms * 1e-3), 1)
                out[f"{nm}_frac"] = round(gb / (ms * 1e-3) / peak, 3)
            x = np.zeros(P.shape[0]); s.solve(x)
            x = np.zeros(P.shape[0]); s.solve(x)
            out["solve25_ms"] = round(s.info["solve_ms"], 2)
            out["pcg"] = s.info["pcg_iters_total"]
            out["kernel_frac"] = round(s.apply_bytes(100) / 1e9 / (s.info["solve_ms"] * 1e-3) / peak, 3)
        print(json.dumps(out), flush=True)
