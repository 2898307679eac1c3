"""Host-side partitioning of ONE large sparse QP over R ranks.

Rank r owns a contiguous, nnz-balanced block of rows I_r of A (and of l, u, z, y) and a contiguous,
nnz-balanced block of columns J_r of P.  With H_r = [P[:, J_r]  A_r'] the reduced KKT operator is
    K u = sum_r ( P[:, J_r] u[J_r] + rho A_r' (A_r u) ) + sigma u ,
one all-reduce(sum) of an n-vector per application; x, q and the CG vectors are replicated.
"""
from __future__ import annotations

import numpy as np
import scipy.sparse as sp


def balanced_blocks(weights: np.ndarray, parts: int) -> np.ndarray:
    """Boundaries b[0..parts] of contiguous blocks with (nearly) equal total weight (+1 per item, so
    that empty rows / columns are spread too)."""
    w = np.asarray(weights, dtype=np.float64) + 1.0
    c = np.concatenate([[0.0], np.cumsum(w)])
    targets = c[-1] * np.arange(1, parts) / parts
    inner = np.searchsorted(c, targets, side="left")
    b = np.concatenate([[0], inner, [len(w)]]).astype(np.int64)
    return np.maximum.accumulate(b)


def plan(P, A, nranks: int):
    """Row blocks of A and column blocks of P for every rank: (row_bounds[R+1], col_bounds[R+1])."""
    A = sp.csr_matrix(A)
    P = sp.csc_matrix(P)
    return balanced_blocks(np.diff(A.indptr), nranks), balanced_blocks(np.diff(P.indptr), nranks)


def slice_problem(P, A, l, u, rank: int, nranks: int, bounds=None):
    """The slice rank ``rank`` passes to ``qpb200_dist_create``: ``(P_r, A_r, l_r, u_r, (i0, i1), (j0, j1))``
    where ``P_r`` is n x n holding only the columns [j0, j1) and ``A_r`` is (i1 - i0) x n."""
    Pc = sp.csc_matrix(P)
    Ar = sp.csr_matrix(A)
    rb, cb = bounds if bounds is not None else plan(Pc, Ar, nranks)
    i0, i1 = int(rb[rank]), int(rb[rank + 1])
    j0, j1 = int(cb[rank]), int(cb[rank + 1])
    n = Pc.shape[0]
    indptr = np.zeros(n + 1, dtype=Pc.indptr.dtype)
    lo, hi = Pc.indptr[j0], Pc.indptr[j1]
    indptr[j0:j1 + 1] = Pc.indptr[j0:j1 + 1] - lo
    indptr[j1 + 1:] = hi - lo
    P_r = sp.csc_matrix((Pc.data[lo:hi], Pc.indices[lo:hi], indptr), shape=(n, n))
    A_r = sp.csc_matrix(Ar[i0:i1, :])
    A_r.sort_indices()
    return P_r, A_r, np.ascontiguousarray(l[i0:i1]), np.ascontiguousarray(u[i0:i1]), (i0, i1), (j0, j1)
