"""CPU restatement (TEST INFRASTRUCTURE -- never imported by the product path) of the reference's second solver,
``/root/reference/ProxQP.jl``:

    min 0.5 x'Px + q'x   s.t.  A x = b,  C x <= d

``SolveQuadraticProgram!(sQpProb::ProxQP; ...)`` (ProxQP.jl:118-173) with its helpers ``UpdateDecomposition!``
(:191-205), ``CalculateRhs!`` (:207-218), ``UpdateX!/S!/Y!/Z!`` (:220-247) and ``CheckConvergence!`` (:250-296),
statement for statement (including the two-step dual updates and the absence of a ``break`` at :153).

PARITY UNPINNED: the reference ships no vectors for this solver and Julia cannot run here.
The start point is explicit, as in the inner constructor ``ProxQP(mP, vQ, mA, vB, mC, vD, vX, vY, vZ, vS)`` (:36).
"""
from __future__ import annotations

import numpy as np
import scipy.linalg as sla
import scipy.sparse as sp


def _inf(v):
    return float(np.max(np.abs(v))) if v.size else 0.0


def proxqp_check_convergence(P, q, A, b, C, d, x, y, z, s, epsAbs, epsRel, rho, adptRho, tau):
    """``CheckConvergence!`` (ProxQP.jl:250-296).  Returns (convFlag, rp, rd, rho, scaleRatio, updatedRho)."""
    MIN_VAL_RHO, MAX_VAL_RHO = 1e-5, 1e5                          # :253-254
    vX1 = P @ x                                                   # :259
    vX2 = A.T @ y
    vX3 = C.T @ z
    vBb = A @ x
    vDb = C @ x
    rp = max(_inf(vBb - b), _inf(vDb - d + s))                    # :264
    rd = _inf(vX1 + vX2 + vX3 + q)                                # :265
    maxP = max(_inf(vBb), _inf(b), _inf(vDb), _inf(d), _inf(s))   # :267
    maxD = max(_inf(vX1), _inf(vX2), _inf(vX3), _inf(q))          # :268
    updated, scale = False, 1.0
    if adptRho:                                                   # :274-283
        with np.errstate(divide="ignore", invalid="ignore"):
            ratio = np.float64(rp * maxD) / np.float64(rd * maxP)
            if (ratio > tau) or (1.0 / ratio > tau):
                updated = True
                rr = float(np.sqrt(np.sqrt(ratio)))
                rhorho = rho * rr
                rhorho = MAX_VAL_RHO if rhorho > MAX_VAL_RHO else (MIN_VAL_RHO if rhorho < MIN_VAL_RHO else rhorho)
                scale = rho / rhorho
                rho = rhorho
    conv = (rp < epsAbs + epsRel * maxP) and (rd < epsAbs + epsRel * maxD)   # :286-291
    return bool(conv), rp, rd, float(rho), scale, updated


def proxqp_solve(P, q, A, b, C, d, x0=None, y0=None, z0=None, s0=None, *, numIterations=2000, epsAbs=1e-7, epsRel=1e-6,
                 numItrConv=50, rho=1e2, sigma=1e-2, adptRho=True, tau=10.0):
    """``SolveQuadraticProgram!(sQpProb::ProxQP; ...)`` (ProxQP.jl:118-173).  Returns ``(x, y, z, s, report)`` with the
    reference's report keys plus ``rho_updates``."""
    P = sp.csr_matrix(P, dtype=np.float64); A = sp.csr_matrix(A, dtype=np.float64); C = sp.csr_matrix(C, dtype=np.float64)
    q = np.asarray(q, dtype=np.float64); b = np.asarray(b, dtype=np.float64); d = np.asarray(d, dtype=np.float64)
    n = P.shape[0]
    x = np.zeros(n) if x0 is None else np.array(x0, dtype=np.float64)
    y = np.zeros(A.shape[0]) if y0 is None else np.array(y0, dtype=np.float64)
    z = np.zeros(C.shape[0]) if z0 is None else np.array(z0, dtype=np.float64)
    s = np.maximum(d - C @ x, 0.0) if s0 is None else np.array(s0, dtype=np.float64)   # :109
    mK = (A.T @ A + C.T @ C).toarray()                            # :40-45
    Pd = P.toarray()

    def factor(rho):                                              # UpdateM! / UpdateDecomposition! (:175-205)
        M = Pd + rho * mK
        M[np.diag_indices(n)] += sigma
        return sla.cho_factor(M, lower=True)

    rho = float(rho)
    rho1 = 1.0 / rho                                              # :129
    fac = factor(rho)                                             # :131
    report = {"Converged": False, "Iterations": int(numIterations), "ρ": rho, "σ": float(sigma),
              "PrimalResidual": float("inf"), "DualResidual": float("inf"), "rho_updates": 0}
    conv = False
    for ii in range(1, int(numIterations) + 1):                   # :135
        r = -q + sigma * x                                        # CalculateRhs! :210-217
        r = r + A.T @ (rho * b - y)
        r = r + C.T @ (rho * (d - s) - z)
        x = sla.cho_solve(fac, r)                                 # UpdateX! :223
        s = np.maximum((d - rho1 * z) - C @ x, 0.0)               # UpdateS! :229-231
        y = (y - rho * b) + rho * (A @ x)                         # UpdateY! :237-238
        z = np.maximum((z + rho * (s - d)) + rho * (C @ x), 0.0)  # UpdateZ! :244-246
        if ii % numItrConv == 0:                                  # :151
            conv, rp, rd, rho_new, _, updated = proxqp_check_convergence(P, q, A, b, C, d, x, y, z, s, epsAbs, epsRel, rho,
                                                                         adptRho, tau)
            report["PrimalResidual"], report["DualResidual"] = rp, rd
            if conv:
                report["Iterations"] = ii                         # :156 (no break: :157 is commented out)
            if updated:                                           # :159-165
                rho = rho_new
                rho1 = 1.0 / rho
                fac = factor(rho)
                report["ρ"] = rho
                report["rho_updates"] += 1
    report["Converged"] = bool(conv)                              # :169
    return x, y, z, s, report


def random_proxqp(n, m_eq, m_in, seed=0, density=1.0):
    """A feasible, strongly convex test problem in ProxQP's form (dense by default, like the reference's unit test)."""
    rng = np.random.default_rng(seed)
    M = rng.standard_normal((n, n)) * (rng.random((n, n)) < density)
    P = M.T @ M + 1e-2 * np.eye(n)
    P = 0.5 * (P + P.T)
    q = rng.standard_normal(n)
    A = rng.standard_normal((m_eq, n)) * (rng.random((m_eq, n)) < density)
    C = rng.standard_normal((m_in, n)) * (rng.random((m_in, n)) < density)
    xs = rng.standard_normal(n)
    b = A @ xs
    d = C @ xs + rng.random(m_in)
    return sp.csc_matrix(P), q, sp.csc_matrix(A), b, sp.csc_matrix(C), d
