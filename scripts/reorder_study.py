"""Would a create-time row/column reordering of H = [P A'] pay on cfg5?    CPU-only study of the
quantity that bounds the SpMV on this matrix: 32-byte sectors of the gathered vector touched per 1024-non-zero tile
(every sector is one L2->L1 request), natural order vs reverse Cuthill-McKee on the symmetric
pattern of the KKT matrix [P A'; A 0].  usage: reorder_study.py [scale]"""
import json
import sys
import time

import numpy as np
import scipy.sparse as sp
from scipy.sparse.csgraph import reverse_cuthill_mckee

sys.path.insert(0, ".")
from workloads.problems import config_cfg5, config_banded  # noqa: E402


def sectors_per_nnz(H, tile=1024):
    """distinct 32-byte sectors (4 doubles) of the gathered vector per tile of consecutive non-zeros, per non-zero"""
    H = sp.csr_matrix(H)
    sec = H.indices // 4
    nt = (len(sec) + tile - 1) // tile
    tot = 0
    for t in range(nt):
        tot += len(np.unique(sec[t * tile:(t + 1) * tile]))
    return tot / len(sec)


scale = float(sys.argv[1]) if len(sys.argv) > 1 else 0.1
out = {}
for name, prob in (("cfg5", config_cfg5(seed=1234, scale=scale)), ("banded", config_banded(n=int(1e6 * scale), m=int(2e6 * scale)))):
    P, q, A, l, u = prob
    n, m = P.shape[0], A.shape[0]
    H = sp.hstack([sp.csr_matrix(P), sp.csr_matrix(A.T)], format="csr")
    K = sp.bmat([[sp.csr_matrix(P), sp.csr_matrix(A.T)], [sp.csr_matrix(A), None]], format="csr")
    t0 = time.time()
    perm = reverse_cuthill_mckee(K, symmetric_mode=True)
    t_rcm = time.time() - t0
    Kp = K[perm][:, perm]
    # rows of the permuted KKT matrix that came from the x-block, gathered from the permuted [x; y] pair
    rec = {"n": n, "m": m, "nnz_H": int(H.nnz), "sectors_per_nnz_H_natural": sectors_per_nnz(H),
           "sectors_per_nnz_kkt_natural": sectors_per_nnz(K), "sectors_per_nnz_kkt_rcm": sectors_per_nnz(Kp),
           "rcm_seconds": round(t_rcm, 1),
           "bandwidth_natural": int(np.max(np.abs(K.tocoo().row - K.tocoo().col))),
           "bandwidth_rcm": int(np.max(np.abs(Kp.tocoo().row - Kp.tocoo().col)))}
    out[name] = rec
    print(name, json.dumps(rec), flush=True)
json.dump(out, open("profiles/r2_reorder_study.json", "w"), indent=1)
