"""Exploration on a B200: the batched dense path at configs[2] scale.  usage: gpu_explore_batch.py [batch]"""
import json, os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from quadraticprogramsolver_b200 import solver as S
from workloads.problems import config_cfg3_batch
from oracle import c_oracle

batch = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
t0 = time.time()
P, q, A, l, u = config_cfg3_batch(batch, 64, 96, seed=1234)
out = {"batch": batch, "gen_s": time.time() - t0}
for unblocked in (False, True):
    t0 = time.time()
    with S.QPB200Batch(P, q, A, l, u, unblockedCholesky=unblocked) as b:
        create_s = time.time() - t0
        for rep in range(2):
            t0 = time.time()
            X, flags, iters = b.solve()
            wall = time.time() - t0
        info = b.info
    out["unblocked" if unblocked else "blocked"] = dict(create_s=create_s, solve_wall_s=wall, solve_ms=info["solve_ms"],
        iters_total=int(iters.sum()), iters_mean=float(iters.mean()), iters_max=int(iters.max()),
        flags={int(k): int(v) for k, v in zip(*np.unique(flags, return_counts=True))},
        solves_per_s=batch / (info["solve_ms"] * 1e-3), admm_iters_per_s=float(iters.sum()) / (info["solve_ms"] * 1e-3))
    print(json.dumps(out), flush=True)
ns = 1024
Xr, fr, ir, sec, rc = c_oracle.solve_dense_batch(P[:ns], q[:ns], A[:ns], l[:ns], u[:ns])
out["cpu"] = dict(sample=ns, seconds=sec, solves_per_s=ns / sec, threads=c_oracle.num_threads(),
                  parity_flags=bool(np.array_equal(fr, flags[:ns])), parity_iters=int(np.max(np.abs(ir - iters[:ns]))),
                  parity_x=float(np.max(np.abs(Xr - X[:ns]) / (1 + np.max(np.abs(Xr), axis=1, keepdims=True)))))
print(json.dumps(out))
json.dump(out, open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "gpurun_out", "explore_batch.json"), "w"), indent=1)
