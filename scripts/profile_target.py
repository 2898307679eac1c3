"""Target program for ncu: cfg5-scale problem, 3 launches of each stand-alone SpMV kernel (A, H, H split),
then one short solve (the persistent ADMM kernel).  usage: profile_target.py [scale] [admm_iters] [loader]"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from quadraticprogramsolver_b200 import solver as S  # noqa: E402
from workloads.problems import config_cfg5  # noqa: E402

scale = float(sys.argv[1]) if len(sys.argv) > 1 else 1.0
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 2
loader = sys.argv[3] if len(sys.argv) > 3 else "tma"
P, q, A, l, u = config_cfg5(seed=1234, scale=scale)
with S.QPB200Solver(P, q, A, l, u, spmvLoader=loader, numIterations=iters) as s:
    for which in (1, 4, 5):
        ms = s.time_apply(which, reps=1, flush_l2=True)     # 2 warm-up + 1 timed launch each
        print("apply", which, ms, "ms", s.apply_bytes(which) / 1e6 / ms, "GB/s")
    x = np.zeros(P.shape[0])
    flag = s.solve(x)
    print("solve", int(flag), s.info)
