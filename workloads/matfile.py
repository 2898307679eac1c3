"""``QpModel.mat``: the problem file the reference's scripts exchange between MATLAB and Julia.

``SolveQuadraticProgramUnitTest.m:84`` writes it (``save('QpModel', 'mP', 'vQ', 'mA', 'vL', 'vU')``) and
``SolveQuadraticProgramUnitTest.jl:47-55`` / ``SolveQuadraticProgramBenchMark.jl:42-50`` read it back with
``matread`` (``dataSource = dataSourceLoaded``), dropping the singleton dimension MATLAB gives the vectors.
Caller-side input format, like the generators in this package: not part of the product library.
"""
from __future__ import annotations

import numpy as np
import scipy.io
import scipy.sparse as sp

FIELDS = ("mP", "vQ", "mA", "vL", "vU")


def load_qp_model(path):
    """``(mP, vQ, mA, vL, vU)`` from a MATLAB v5/v7 ``.mat`` file: matrices as ``csc_matrix`` float64 (sparse or full in
    the file), vectors as 1-D float64 (``dropdims(...; dims = 2)`` of the Julia readers); +-Inf bounds pass through."""
    d = scipy.io.loadmat(path, spmatrix=True)       # explicit: scipy is moving the default to sparse arrays
    missing = [k for k in FIELDS if k not in d]
    if missing:
        raise KeyError(f"{path}: missing variable(s) {missing}; expected {FIELDS}")
    P, A = sp.csc_matrix(d["mP"], dtype=np.float64), sp.csc_matrix(d["mA"], dtype=np.float64)
    q, l, u = (np.asarray(d[k].todense() if sp.issparse(d[k]) else d[k], dtype=np.float64).reshape(-1) for k in ("vQ", "vL", "vU"))
    n = P.shape[0]
    if A.shape == (0, 0) or A.shape[0] == 0:            # MATLAB writes an empty constraint matrix as 0 x 0
        A = sp.csc_matrix((0, n), dtype=np.float64)
    if P.shape != (n, n) or q.shape != (n,) or A.shape[1] != n or l.shape != (A.shape[0],) or u.shape != (A.shape[0],):
        raise ValueError(f"{path}: inconsistent sizes P {P.shape}, q {q.shape}, A {A.shape}, l {l.shape}, u {u.shape}")
    P.sort_indices()
    A.sort_indices()
    return P, q, A, l, u


def save_qp_model(path, P, q, A, l, u):
    """Write the five variables the way MATLAB's ``save`` does: sparse double matrices, column vectors."""
    col = lambda v: np.asarray(v, dtype=np.float64).reshape(-1, 1)
    scipy.io.savemat(path, {"mP": sp.csc_matrix(P, dtype=np.float64), "vQ": col(q), "mA": sp.csc_matrix(A, dtype=np.float64),
                            "vL": col(l), "vU": col(u)}, do_compression=True)
