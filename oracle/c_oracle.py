"""ORACLE -- TEST INFRASTRUCTURE ONLY.  ctypes binding of oracle/libqp_oracle.so (qp_oracle.c).

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline legs may import this.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np
import scipy.sparse as sp

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


class OracleSettings(C.Structure):
    _fields_ = [("max_iter", C.c_int64), ("eps_abs", C.c_double), ("eps_rel", C.c_double),
                ("rho", C.c_double), ("sigma", C.c_double), ("alpha", C.c_double),
                ("adaptive_rho", C.c_int32), ("rho_factor", C.c_double), ("check_every", C.c_int64),
                ("pcg_eps", C.c_double), ("pcg_max_iter", C.c_int64), ("precond", C.c_int32),
                ("time_limit_s", C.c_double)]


class OracleInfo(C.Structure):
    _fields_ = [("conv_flag", C.c_int32), ("iterations", C.c_int64), ("rho_final", C.c_double),
                ("res_prim", C.c_double), ("res_dual", C.c_double), ("rho_updates", C.c_int64),
                ("cg_iters_total", C.c_int64), ("solve_seconds", C.c_double)]


def build(force: bool = False) -> str:
    so = os.path.join(_HERE, "libqp_oracle.so")
    src = os.path.join(_HERE, "qp_oracle.c")
    if force or not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        subprocess.check_call(["make", "-s", "-C", _HERE, "libqp_oracle.so"])
    return so


def lib():
    global _LIB
    if _LIB is None:
        _LIB = C.CDLL(build())
    return _LIB


def settings(numIterations=5000, epsAbs=1e-6, epsRel=1e-6, rho=1.0, sigma=1e-6, alpha=1.6, adptRho=False,
             fctrRho=5.0, numItrConv=25, epsPcg=1e-6, numItrPcg=1000, precond=0, time_limit_s=0.0):
    return OracleSettings(numIterations, epsAbs, epsRel, rho, sigma, alpha, int(adptRho), fctrRho, numItrConv,
                          epsPcg, numItrPcg, int(precond), time_limit_s)


def _p(a, t):
    return a.ctypes.data_as(C.POINTER(t))


def solve_sparse(mP, vQ, mA, vL, vU, x0=None, **kw):
    """Modes M (precond=0) / J (precond=1).  Returns (x, flag, info dict)."""
    mP = sp.csc_matrix(mP); mA = sp.csc_matrix(mA)
    n = mP.shape[0]; m = mA.shape[0]
    Pp = mP.indptr.astype(np.int64); Pi = mP.indices.astype(np.int64); Pv = np.ascontiguousarray(mP.data, np.float64)
    Ap = mA.indptr.astype(np.int64); Ai = mA.indices.astype(np.int64); Av = np.ascontiguousarray(mA.data, np.float64)
    x = np.zeros(n) if x0 is None else np.array(x0, dtype=np.float64)
    z = np.zeros(m); y = np.zeros(m)
    st = settings(**kw)
    info = OracleInfo()
    q = np.ascontiguousarray(vQ, np.float64); l = np.ascontiguousarray(vL, np.float64); u = np.ascontiguousarray(vU, np.float64)
    rc = lib().oracle_solve_sparse(C.c_int64(n), C.c_int64(m), _p(Pp, C.c_int64), _p(Pi, C.c_int64), _p(Pv, C.c_double),
                                   _p(Ap, C.c_int64), _p(Ai, C.c_int64), _p(Av, C.c_double),
                                   _p(q, C.c_double), _p(l, C.c_double), _p(u, C.c_double), C.c_int64(0),
                                   C.byref(st), _p(x, C.c_double), _p(z, C.c_double), _p(y, C.c_double), C.byref(info))
    assert rc == 0
    d = {k: getattr(info, k) for k, _ in OracleInfo._fields_}
    d.update(z=z, y=y)
    return x, info.conv_flag, d


def solve_dense_batch(P, q, A_cm, l, u, x0=None, **kw):
    """Batched dense mode D.  P[b,n,n] (col-major blocks), A_cm[b,n,m] (= col-major m x n blocks)."""
    batch, n, _ = P.shape
    m = A_cm.shape[2]
    X = np.zeros((batch, n)) if x0 is None else np.array(x0, dtype=np.float64)
    flags = np.zeros(batch, np.int32); iters = np.zeros(batch, np.int64)
    st = settings(**kw)
    sec = C.c_double(0.0)
    P = np.ascontiguousarray(P); A_cm = np.ascontiguousarray(A_cm)
    q = np.ascontiguousarray(q); l = np.ascontiguousarray(l); u = np.ascontiguousarray(u)
    rc = lib().oracle_solve_dense_batch(C.c_int64(batch), C.c_int64(n), C.c_int64(m), _p(P, C.c_double), _p(A_cm, C.c_double),
                                        _p(q, C.c_double), _p(l, C.c_double), _p(u, C.c_double), C.byref(st),
                                        _p(X, C.c_double), _p(flags, C.c_int32), _p(iters, C.c_int64), C.byref(sec))
    return X, flags, iters, sec.value, rc


def num_threads() -> int:
    return int(lib().oracle_num_threads())


def use_all_cores() -> int:
    """All cores this process may run on, regardless of OMP_NUM_THREADS (torchrun sets it to 1)."""
    try:
        nt = len(os.sched_getaffinity(0))
    except AttributeError:
        nt = os.cpu_count() or 1
    return int(lib().oracle_set_num_threads(C.c_int(nt)))
