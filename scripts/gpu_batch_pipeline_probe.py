"""Probe for the next round: can the end-to-end time of the dense batch (configs[2]) hide the solve behind the
upload?  Uses only the existing C entry points: the batch is cut into chunks, every chunk gets its own handle; one
Python thread creates (= uploads) chunk i+1 while another solves chunk i (ctypes releases the GIL; distinct handles
may be driven from distinct host threads, include/qpb200.h).  Prints the serial end-to-end time, the pipelined one and
whether the results agree bit for bit.
usage: python scripts/gpu_batch_pipeline_probe.py [batch] [chunks]"""
import json
import os
import queue
import sys
import threading
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from quadraticprogramsolver_b200 import solver as S                       # noqa: E402
from workloads.problems import config_cfg3_batch        # noqa: E402

batch = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
chunks = int(sys.argv[2]) if len(sys.argv) > 2 else 8
P, q, A, l, u = config_cfg3_batch(batch, 64, 96, seed=1234)


def serial():
    t0 = time.perf_counter()
    X, flags, iters, info = S.SolveQuadraticProgramBatch(P, q, A, l, u)
    return time.perf_counter() - t0, X, iters


def pipelined():
    bounds = [batch * c // chunks for c in range(chunks + 1)]
    X = np.zeros((batch, 64))
    iters = np.zeros(batch, dtype=np.int64)
    ready = queue.Queue(maxsize=2)

    def uploader():
        for c in range(chunks):
            lo, hi = bounds[c], bounds[c + 1]
            ready.put((lo, hi, S.QPB200Batch(P[lo:hi], q[lo:hi], A[lo:hi], l[lo:hi], u[lo:hi])))
        ready.put(None)

    t0 = time.perf_counter()
    th = threading.Thread(target=uploader)
    th.start()
    while True:
        item = ready.get()
        if item is None:
            break
        lo, hi, b = item
        Xc, _, ic = b.solve()
        X[lo:hi], iters[lo:hi] = Xc, ic
        b.close()
    th.join()
    return time.perf_counter() - t0, X, iters


serial()                                   # warm-up (context, module load, page-locked ring, block caches)
ts, Xs, its = serial()
pipelined()
tp, Xp, itp = pipelined()
print(json.dumps({"batch": batch, "chunks": chunks, "serial_e2e_s": round(ts, 4), "pipelined_e2e_s": round(tp, 4),
                  "serial_solves_per_s": round(batch / ts, 1), "pipelined_solves_per_s": round(batch / tp, 1),
                  "bitwise_equal": bool(np.array_equal(Xs, Xp) and np.array_equal(its, itp))}))
