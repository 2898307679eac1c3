"""Target program for ncu: stand-alone SpMV launches (A, H) on a cached workload (cfg5 / banded / cfg4).
usage: profile_spmv.py workload [loader]"""
import os
import sys

import numpy as np
import scipy.sparse as sp

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from quadraticprogramsolver_b200 import solver as S  # noqa: E402
from workloads import problems  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else "banded"
loader = sys.argv[2] if len(sys.argv) > 2 else "tma"
path = f"/tmp/qpb_{name}.npz"
if os.path.exists(path):
    d = np.load(path)
    P = sp.csc_matrix((d["Pv"], d["Pi"], d["Pp"]), shape=(int(d["n"]), int(d["n"])))
    A = sp.csc_matrix((d["Av"], d["Ai"], d["Ap"]), shape=(int(d["m"]), int(d["n"])))
    q, l, u = d["q"], d["l"], d["u"]
else:
    P, q, A, l, u = {"cfg5": problems.config_cfg5, "banded": problems.config_banded, "cfg4": problems.config_cfg4}[name]()
    P = sp.csc_matrix(P); A = sp.csc_matrix(A)
    np.savez(path, Pv=P.data, Pi=P.indices, Pp=P.indptr, Av=A.data, Ai=A.indices, Ap=A.indptr, q=q, l=l, u=u,
             n=P.shape[0], m=A.shape[0])
with S.QPB200Solver(P, q, A, l, u, spmvLoader=loader, numIterations=1) as s:
    for which in (1, 4):
        ms = s.time_apply(which, reps=1, flush_l2=True)     # 2 warm-up + 1 timed launch each
        print("apply", name, which, ms, "ms", s.apply_bytes(which) / 1e6 / ms, "GB/s")
