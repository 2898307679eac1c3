"""Dense batch kernel A/B (configs[2] shape n = 64, m = 96): shared-memory products vs A (and K^-1) held in registers.
usage: python scripts/gpu_dense_variant_ab.py [batch]"""
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from quadraticprogramsolver_b200 import solver as S                       # noqa: E402
from workloads.problems import config_cfg3_batch        # noqa: E402

batch = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
P, q, A, l, u = config_cfg3_batch(batch, 64, 96, seed=1234)
ref = None
for variant in ("smem", "regs", "regs_ak", "smem", "regs", "regs_ak"):
    with S.QPB200Batch(P, q, A, l, u, denseVariant=variant) as b:
        X, flags, iters = b.solve()
        info = b.info
    if ref is None:
        ref = (X.copy(), flags.copy(), iters.copy())
    print(json.dumps({"variant": variant, "batch": batch, "solve_ms": round(info["solve_ms"], 3),
                      "admm_iters": int(iters.sum()), "solves_per_s": round(batch / (info["solve_ms"] * 1e-3), 1),
                      "ns_per_iter_per_qp_slot": round(1e6 * info["solve_ms"] / max(1, int(iters.sum())), 3),
                      "flags_equal": bool(np.array_equal(flags, ref[1])), "iters_equal": bool(np.array_equal(iters, ref[2])),
                      "max_dx": float(np.max(np.abs(X - ref[0])))}), flush=True)
