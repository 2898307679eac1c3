"""Host-side mirror of the reference's second solver interface (``/root/reference/ProxQP.jl``) over the C ABI.

    min 0.5 x'Px + q'x   s.t.   A x = b,   C x <= d

The reference holds the problem and the iterates in ``struct ProxQP`` (ProxQP.jl:8-65: ``vX, mP, vQ, mA, vB, mC, vD,
vY, vZ, vS`` plus work buffers and the Cholesky factor) and solves it with
``SolveQuadraticProgram!(sQpProb::ProxQP; numIterations = 2000, ϵAbs = 1e-7, ϵRel = 1e-6, numItrConv = 50, ρ = 1e2,
σ = 1e-2, adptΡ = true, τ = 10)`` (:118-173), returning a report dictionary.  Here :class:`ProxQP` keeps the same
field names (the factor and the buffers live on the GPU behind a ``qpb200_handle``), and
:func:`SolveQuadraticProgram_` takes the same keywords and returns the same report keys.  The whole iteration runs in
``proxqp_kernel`` (csrc/proxqp_kernels.cuh); there is no CPU path.
"""
from __future__ import annotations

import ctypes as C

import numpy as np
import scipy.sparse as sp

from . import _lib
from .solver import _ALIASES, _csc_arrays, _p64, _pd

_PROX_DEFAULTS = dict(numIterations=2000, epsAbs=1e-7, epsRel=1e-6, numItrConv=50, rho=1e2, sigma=1e-2, adptRho=True, tau=10.0)


class ProxQP:
    """``ProxQP(mP, vQ, mA, vB, mC, vD[, vX, vY, vZ, vS])`` (ProxQP.jl:36, :91).  The start point defaults to zeros
    with ``vS = max(vD - mC vX, 0)`` (:109); pass ``vX, vY`` to reproduce the equality-constrained warm start of the
    dense convenience constructor (:91-113)."""

    def __init__(self, mP, vQ, mA, vB, mC, vD, vX=None, vY=None, vZ=None, vS=None, device=-1):
        lib = _lib.load()
        self.mP = sp.csc_matrix(mP, dtype=np.float64)
        self.mA = sp.csc_matrix(mA, dtype=np.float64)
        self.mC = sp.csc_matrix(mC, dtype=np.float64)
        self.vQ = np.ascontiguousarray(vQ, dtype=np.float64)
        self.vB = np.ascontiguousarray(vB, dtype=np.float64)
        self.vD = np.ascontiguousarray(vD, dtype=np.float64)
        self.dataDim, self.numEq, self.numInEq = self.mP.shape[0], self.mA.shape[0], self.mC.shape[0]
        if self.mA.shape[1] != self.dataDim or self.mC.shape[1] != self.dataDim or self.vQ.shape != (self.dataDim,) \
                or self.vB.shape != (self.numEq,) or self.vD.shape != (self.numInEq,):
            raise ValueError("ProxQP: dimension mismatch")
        self.vX = np.zeros(self.dataDim) if vX is None else np.array(vX, dtype=np.float64)
        self.vY = np.zeros(self.numEq) if vY is None else np.array(vY, dtype=np.float64)
        self.vZ = np.zeros(self.numInEq) if vZ is None else np.array(vZ, dtype=np.float64)
        self.vS = None if vS is None else np.array(vS, dtype=np.float64)
        # the handle holds the stacked constraint matrix [A; C]; u = [b; d] (l is not used by this solver)
        stacked = sp.vstack([self.mA, self.mC], format="csc")
        Pp, Pi, Pv = _csc_arrays(self.mP)
        Ap, Ai, Av = _csc_arrays(stacked)
        bd = np.concatenate([self.vB, self.vD])
        lo = np.concatenate([self.vB, np.full(self.numInEq, -np.inf)])
        s = _lib.Settings()
        lib.qpb200_proxqp_default_settings(C.byref(s))
        s.device = int(device)
        self._h = C.c_void_p()
        _lib.check(lib.qpb200_create(C.byref(self._h), self.dataDim, self.numEq + self.numInEq, _p64(Pp), _p64(Pi), _pd(Pv),
                                     _p64(Ap), _p64(Ai), _pd(Av), _pd(self.vQ), _pd(lo), _pd(bd), C.byref(s), 0))

    def close(self):
        if getattr(self, "_h", None) is not None and self._h.value:
            _lib.load().qpb200_destroy(self._h)
            self._h = C.c_void_p()

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def SolveQuadraticProgram_(sQpProb: ProxQP, **kw):
    """``SolveQuadraticProgram!(sQpProb::ProxQP; ...)`` (ProxQP.jl:118): iterates ``sQpProb.vX, vY, vZ, vS`` in place and
    returns ``dReport`` with the reference's keys (``Converged, Iterations, ρ, σ, PrimalResidual, DualResidual``)."""
    opts = dict(_PROX_DEFAULTS)
    for k, v in kw.items():
        k = {"τ": "tau"}.get(k, _ALIASES.get(k, k))
        if k not in opts:
            raise TypeError(f"SolveQuadraticProgram!(::ProxQP): unknown keyword argument {k!r}")
        opts[k] = v
    lib = _lib.load()
    s = _lib.Settings()
    lib.qpb200_proxqp_default_settings(C.byref(s))
    s.max_iter = int(opts["numIterations"]); s.eps_abs = float(opts["epsAbs"]); s.eps_rel = float(opts["epsRel"])
    s.check_every = int(opts["numItrConv"]); s.rho = float(opts["rho"]); s.sigma = float(opts["sigma"])
    s.adaptive_rho = int(bool(opts["adptRho"])); s.rho_factor = float(opts["tau"])
    rep = _lib.ProxReport()
    p = sQpProb
    init_slack = p.vS is None                        # ProxQP.jl:109: vS = max(vD - mC vX, 0), computed on the device
    if init_slack:
        p.vS = np.zeros(p.numInEq)
    _lib.check(lib.qpb200_proxqp_solve(p._h, p.numEq, C.byref(s), _pd(p.vX), _pd(p.vY), _pd(p.vZ), _pd(p.vS), int(init_slack),
                                       C.byref(rep)))
    return {"Converged": bool(rep.converged), "Iterations": int(rep.iterations), "ρ": float(rep.rho), "σ": float(rep.sigma),
            "PrimalResidual": float(rep.res_prim), "DualResidual": float(rep.res_dual),
            "rho_updates": int(rep.rho_updates), "solve_ms": float(rep.solve_ms), "kernel_launches": int(rep.kernel_launches)}
