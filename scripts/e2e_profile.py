"""Where does the end-to-end time of one create -> solve -> destroy cycle go?  (host wall clock per C call)
usage: python scripts/e2e_profile.py [scale] [iters] [reps]"""
import ctypes as C
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from quadraticprogramsolver_b200 import _lib, solver as S
from workloads import problems   # noqa: E402

scale = float(sys.argv[1]) if len(sys.argv) > 1 else 1.0
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 100
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 4

P, q, A, l, u = problems.config_cfg5(seed=1234, scale=scale)
n, m = P.shape[0], A.shape[0]
Pa, Aa = S._csc_arrays(P), S._csc_arrays(A)
lib = _lib.load()
rows = []
for rep in range(reps):
    t = [time.perf_counter()]
    settings = S.make_settings(numIterations=iters)
    h = C.c_void_p()
    x = np.zeros(n)
    z = np.empty(m)
    y = np.empty(m)
    t.append(time.perf_counter())
    _lib.check(lib.qpb200_create(C.byref(h), n, m, S._p64(Pa[0]), S._p64(Pa[1]), S._pd(Pa[2]), S._p64(Aa[0]), S._p64(Aa[1]),
                                 S._pd(Aa[2]), S._pd(q), S._pd(l), S._pd(u), C.byref(settings), 0))
    t.append(time.perf_counter())
    info = S.Info()
    _lib.check(lib.qpb200_solve(h, S._pd(x), S._pd(z), S._pd(y), C.byref(info)))
    t.append(time.perf_counter())
    lib.qpb200_destroy(h)
    t.append(time.perf_counter())
    d = info.as_dict()
    rows.append({"rep": rep, "alloc_host_ms": 1e3 * (t[1] - t[0]), "create_ms": 1e3 * (t[2] - t[1]), "solve_call_ms": 1e3 * (t[3] - t[2]),
                 "solve_device_ms": d["solve_ms"], "destroy_ms": 1e3 * (t[4] - t[3]), "iterations": d["iterations"]})
    print(json.dumps(rows[-1]), flush=True)
