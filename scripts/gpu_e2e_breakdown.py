"""Where does the end-to-end time of one SolveQuadraticProgram call on cfg5 go?  (QPB200_TIMING=1 for the C side.)"""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from quadraticprogramsolver_b200 import solver as S
from workloads.problems import config_cfg5
P, q, A, l, u = config_cfg5(seed=1234)
n, m = P.shape[0], A.shape[0]
Parr, Aarr = S.csc_arrays_int64(P), S.csc_arrays_int64(A)
for rep in range(3):
    t0 = time.perf_counter()
    x = np.zeros(n)
    flag, info = S.solve_csc_arrays(n, m, Parr, q, Aarr, l, u, x, want_zy=True, numIterations=100)
    t1 = time.perf_counter()
    print(f"rep {rep}: total {1e3 * (t1 - t0):.1f} ms (create {info['setup_ms']:.1f}, solve device {info['solve_ms']:.1f})", flush=True)
