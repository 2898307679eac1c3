"""ORACLE -- TEST INFRASTRUCTURE ONLY.  Not part of the product path.

CPU restatement (numpy / scipy) of the reference's OSQP-style ADMM hot path:

* ``SolveQuadraticProgram!`` and ``CheckConvergence``   -- SolveQuadraticProgram.jl:14-112
* the KKT-solve plugin pairs ``(Init, Sol!)``          -- LinearSystemSolvers.jl:16-229
* ``IterativeSolvers.cg!`` (third-party, NOT in /root/reference, version un-pinned: there is no
  Manifest.toml and Project.toml does not list it) -- restated from the published algorithm of
  IterativeSolvers.jl v0.9.x (``cg_iterator!`` / ``CGIterable`` / ``PCGIterable``):
  ``tol = max(reltol*||b - A x0||, abstol)``, ``reltol = sqrt(eps)``, stop when
  ``||r||_2 <= tol`` or ``iter >= maxiter``.  Anchored on the reference call sites
  LinearSystemSolvers.jl:137,181,224 (``abstol = 1e-6, maxiter = 1000``, warm start).

PARITY UNPINNED: the reference ships no golden vectors, no known-answer tests and no pinned
dependency versions, and neither Julia nor MATLAB/Octave exists in this environment, so this
restatement cannot be checked against outputs of the reference itself.  It is validated instead by
(1) an independent KKT optimality certificate of every solution, (2) agreement between the four
linear-solver modes to the reference's own threshold 1e-5 (RunTests.jl:58), and (3) agreement with
the independently written C restatement ``qp_oracle.c``.  See DESIGN.md.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline legs may import this.

Modes (plugin pairs):
  D  ``la_ldl_init / la_ldl``          exact KKT solve (LaLdl/QDLdl/FacLdl, LinearSystemSolvers.jl:16-107)
  C  ``itr_sol_cg_init / itr_sol_cg``  CG on the explicit mL = P + sigma I + rho A'A (:110-142)
  M  ``lin_op_cg_init / lin_op_cg``    CG on the matrix-free operator (:145-186, same as :188-229)
  J  ``jacobi_pcg_init / jacobi_pcg``  M with ``Pl = Diagonal(diag(K))`` -- the Jacobi-PCG that
     BASELINE.json's north_star adds (not in the reference); IterativeSolvers' PCGIterable semantics.
"""
from __future__ import annotations

import enum
import math

import numpy as np
import scipy.sparse as sp
import scipy.sparse.linalg as spla


class ConvergenceFlag(enum.IntEnum):
    """``@enum ConvergenceFlag convNumItr = 1 convAdmm convPrimDual`` (SolveQuadraticProgram.jl:12)."""

    convNumItr = 1
    convAdmm = 2
    convPrimDual = 3


def _clamp(x, lo, hi):
    """Julia ``clamp(x, lo, hi) = ifelse(x > hi, hi, ifelse(x < lo, lo, x))`` (NaN passes through)."""
    if isinstance(x, np.ndarray):
        return np.where(x > hi, hi, np.where(x < lo, lo, x))
    if x > hi:
        return hi
    if x < lo:
        return lo
    return x


def _norm_inf(v) -> float:
    return float(np.max(np.abs(v))) if v.size else 0.0


# --------------------------------------------------------------------------------------------
# IterativeSolvers.cg! restatement
# --------------------------------------------------------------------------------------------

def cg(x, apply_A, b, *, abstol=0.0, reltol=math.sqrt(np.finfo(np.float64).eps), maxiter=None,
       diag_precond=None):
    """In-place ``cg!(x, A, b; abstol, reltol, maxiter, Pl)``.  Returns (#iterations, #mv-products).

    Un-preconditioned (``CGIterable``): beta = res^2/prev_res^2; u = r + beta u; c = A u;
    alpha = res^2 / (u.c); x += alpha u; r -= alpha c; res = ||r||.
    Preconditioned (``PCGIterable``, Pl = Diagonal(d)): c = r ./ d; rho = c.r; beta = rho/rho_prev;
    u = c + beta u; c = A u; alpha = rho/(u.c); x += alpha u; r -= alpha c; res = ||r||.
    """
    n = b.shape[0]
    if maxiter is None:
        maxiter = n
    u = np.zeros(n)
    r = b.copy()
    c = apply_A(x)
    r -= c
    mv = 1
    residual = float(np.linalg.norm(r))
    tol = max(reltol * residual, abstol)
    it = 0
    if diag_precond is None:
        prev_residual = 1.0
        while it < maxiter and not (residual <= tol):
            beta = residual ** 2 / prev_residual ** 2
            u = r + beta * u
            c = apply_A(u)
            mv += 1
            alpha = residual ** 2 / float(np.dot(u, c))
            x += alpha * u
            r -= alpha * c
            prev_residual = residual
            residual = float(np.linalg.norm(r))
            it += 1
    else:
        rho = 1.0
        while it < maxiter and not (residual <= tol):
            c = r / diag_precond
            rho_prev = rho
            rho = float(np.dot(c, r))
            beta = rho / rho_prev
            u = c + beta * u
            c = apply_A(u)
            mv += 1
            alpha = rho / float(np.dot(u, c))
            x += alpha * u
            r -= alpha * c
            residual = float(np.linalg.norm(r))
            it += 1
    return it, mv


# --------------------------------------------------------------------------------------------
# Plugin pairs (LinearSystemSolvers.jl).  Init(vX,mP,vQ,mA,rho,rho1,sigma,n,m) -> (vXX,vZZ,state)
# Sol(state,vXX,vZZ,vX,mP,vQ,mA,vZ,vY,rho,rho1,sigma,n,m,changedRho) mutates vXX, vZZ.
# --------------------------------------------------------------------------------------------

def _kkt(mP, mA, rho1, sigma, n, m):
    """rho1: the scalar 1/rho of the reference, or the vector 1 ./ (rho * rhoScale)."""
    lower_right = -sp.diags(np.broadcast_to(np.asarray(rho1, dtype=np.float64), (m,)), format="csc")
    return sp.bmat([[mP + sigma * sp.identity(n, format="csc"), mA.T],
                    [mA, lower_right]], format="csc")


def _rho_vec(st, rho, rho1):
    """(rho_i, 1/rho_i): the reference's scalars unless the driver stored a per-constraint scale
    (``rhoScale``, not in the reference) in the plugin state."""
    rs = st.get("rho_scale")
    if rs is None:
        return rho, rho1
    rv = rho * rs
    return rv, 1.0 / rv


def la_ldl_init(vX, mP, vQ, mA, rho, rho1, sigma, n, m):
    """LaLdlInit / QDLdlInit / FacLdlInit (LinearSystemSolvers.jl:16-26,47-57,78-89): factor the
    quasi-definite KKT matrix once; vXX, vZZ are views of one (n+m) buffer."""
    hDL = spla.splu(_kkt(mP, mA, rho1, sigma, n, m))
    vV = np.zeros(n + m)
    return vV[:n], vV[n:], {"hDL": hDL, "vV": vV, "n_factor": 1, "cg_iters": 0}


def la_ldl(st, vXX, vZZ, vX, mP, vQ, mA, vZ, vY, rho, rho1, sigma, n, m, changedRho):
    """LaLdl! / QDLdl! / FacLdl! (LinearSystemSolvers.jl:28-44,59-75,91-107)."""
    rho, rho1 = _rho_vec(st, rho, rho1)
    if changedRho or st.pop("refactor", False):
        st["hDL"] = spla.splu(_kkt(mP, mA, rho1, sigma, n, m))     # :30-32 full refactor
        st["n_factor"] += 1
    vV = st["vV"]
    vXX[:] = sigma * vX - vQ                                        # :37
    vZZ[:] = vZ - rho1 * vY                                         # :38
    vV[:] = st["hDL"].solve(vV)                                     # :39
    vZZ[:] = vZ + rho1 * (vZZ - vY)                                 # :40


def itr_sol_cg_init(vX, mP, vQ, mA, rho, rho1, sigma, n, m):
    """ItrSolCgInit (LinearSystemSolvers.jl:110-123): explicit mL."""
    mAA = (mA.T @ mA).tocsr()
    mPI = (mP + sigma * sp.identity(n, format="csc")).tocsr()
    mL = (mPI + rho * mAA).tocsr()
    return np.zeros(n), np.zeros(m), {"mL": mL, "mPI": mPI, "mAA": mAA, "vT": np.zeros(n), "cg_iters": 0,
                                      "eps_pcg": 1e-6, "num_itr_pcg": 1000}


def itr_sol_cg(st, vXX, vZZ, vX, mP, vQ, mA, vZ, vY, rho, rho1, sigma, n, m, changedRho):
    """ItrSolCg! (LinearSystemSolvers.jl:125-142)."""
    if changedRho:
        st["mL"] = (st["mPI"] + rho * st["mAA"]).tocsr()            # :127-129
    mL = st["mL"]
    vT = st["vT"]
    vZZ[:] = rho * vZ - vY                                          # :134
    vT[:] = mA.T @ vZZ                                              # :135
    vT[:] = sigma * vX - vQ + vT                                    # :136
    it, _ = cg(vXX, lambda w: mL @ w, vT, abstol=st["eps_pcg"], maxiter=st["num_itr_pcg"])   # :137
    st["cg_iters"] += it
    vZZ[:] = mA @ vXX                                               # :139


def _matfree_state(mP, mA, n, m):
    return {"mPr": sp.csr_matrix(mP), "mAr": sp.csr_matrix(mA), "mAt": sp.csr_matrix(mA.T),
            "vT": np.zeros(n), "cg_iters": 0, "eps_pcg": 1e-6, "num_itr_pcg": 1000}


def lin_op_cg_init(vX, mP, vQ, mA, rho, rho1, sigma, n, m):
    """LinOpCgInit / LinMapsCgInit (LinearSystemSolvers.jl:145-162,188-205)."""
    return np.zeros(n), np.zeros(m), _matfree_state(mP, mA, n, m)


def _apply_K(st, vZZ, w, rho, sigma):
    """The closure at LinearSystemSolvers.jl:152-157: vZZ = A w; u = A' vZZ; u = P w + rho u;
    u += sigma w.  (It uses the plugin's vZZ as its scratch m-vector.)"""
    vZZ[:] = st["mAr"] @ w
    if st.get("rho_scale") is None:
        u = st["mAt"] @ vZZ
        u = st["mPr"] @ w + rho * u
    else:                                          # K = P + sigma I + A' diag(rho_i) A
        u = st["mAt"] @ ((rho * st["rho_scale"]) * vZZ)
        u = st["mPr"] @ w + u
    u = u + sigma * w
    return u


def lin_op_cg(st, vXX, vZZ, vX, mP, vQ, mA, vZ, vY, rho, rho1, sigma, n, m, changedRho):
    """LinOpCg! / LinMapsCg! (LinearSystemSolvers.jl:164-186,207-229)."""
    vT = st["vT"]
    vZZ[:] = _rho_vec(st, rho, rho1)[0] * vZ - vY                   # :178
    vT[:] = st["mAt"] @ vZZ                                         # :179
    vT[:] = sigma * vX - vQ + vT                                    # :180
    it, _ = cg(vXX, lambda w: _apply_K(st, vZZ, w, rho, sigma), vT,
               abstol=st["eps_pcg"], maxiter=st["num_itr_pcg"])     # :181
    st["cg_iters"] += it
    vZZ[:] = st["mAr"] @ vXX                                        # :183


def jacobi_pcg_init(vX, mP, vQ, mA, rho, rho1, sigma, n, m):
    """Mode J (north_star's Jacobi-PCG; not in the reference): diag(K) = diag(P) + sigma +
    rho * colsumsq(A)."""
    st = _matfree_state(mP, mA, n, m)
    st["dP"] = np.asarray(sp.csc_matrix(mP).diagonal(), dtype=np.float64)
    mAc = sp.csc_matrix(mA)
    st["dAA"] = np.asarray(mAc.multiply(mAc).sum(axis=0)).ravel()
    return np.zeros(n), np.zeros(m), st


def jacobi_pcg(st, vXX, vZZ, vX, mP, vQ, mA, vZ, vY, rho, rho1, sigma, n, m, changedRho):
    vT = st["vT"]
    vZZ[:] = _rho_vec(st, rho, rho1)[0] * vZ - vY
    vT[:] = st["mAt"] @ vZZ
    vT[:] = sigma * vX - vQ + vT
    if st.get("rho_scale") is not None and "dAA_scaled" not in st:
        mAc = sp.csc_matrix(mA)
        st["dAA_scaled"] = np.asarray(sp.diags(st["rho_scale"]).dot(mAc.multiply(mAc)).sum(axis=0)).ravel()
    d = st["dP"] + sigma + rho * (st["dAA"] if st.get("rho_scale") is None else st["dAA_scaled"])
    it, _ = cg(vXX, lambda w: _apply_K(st, vZZ, w, rho, sigma), vT,
               abstol=st["eps_pcg"], maxiter=st["num_itr_pcg"], diag_precond=d)
    st["cg_iters"] += it
    vZZ[:] = st["mAr"] @ vXX


PLUGINS = {
    "D": (la_ldl_init, la_ldl),
    "C": (itr_sol_cg_init, itr_sol_cg),
    "M": (lin_op_cg_init, lin_op_cg),
    "J": (jacobi_pcg_init, jacobi_pcg),
}


# --------------------------------------------------------------------------------------------
# ADMM driver
# --------------------------------------------------------------------------------------------

def check_convergence(vX, mP, vQ, mA, vZ, vY, vXP, vZP, rho, rhorho, adptRho, epsAbs, epsRel, epsAdmm, convFlag,
                      scaling=None):
    """``CheckConvergence`` (SolveQuadraticProgram.jl:79-112).  Returns (rhorho, convFlag, norms).

    ``scaling = (D, Dinvc, Einv, normQ)`` (not in the reference): the iterates belong to the equilibrated
    problem and every norm is taken on the residual of the UNSCALED problem -- ``|.| * Einv`` on
    constraint-space vectors, ``|.| * Dinvc`` (= 1/(c D)) on dual-residual vectors, ``|dx| * D``."""
    MIN_VAL_RHO = 1e-3                                              # :81
    MAX_VAL_RHO = 1e6                                               # :82
    vAx = mA @ vX
    vPx = mP @ vX
    vAty = mA.T @ vY
    if scaling is None:
        wX = wD = wE = 1.0
        normQ = _norm_inf(vQ)
    else:
        wX, wD, wE, normQ = scaling
    normResPrim = _norm_inf(np.abs(vAx - vZ) * wE)                  # :85
    normResDual = _norm_inf(np.abs(vPx + vQ + vAty) * wD)           # :86
    maxNormPrim = max(_norm_inf(np.abs(vAx) * wE), _norm_inf(np.abs(vZ) * wE))            # :88
    maxNormDual = max(_norm_inf(np.abs(vPx) * wD), _norm_inf(np.abs(vAty) * wD), normQ)   # :89
    if adptRho:                                                     # :92-96
        numeratorVal = normResPrim * maxNormDual
        denominatorVal = normResDual * maxNormPrim
        with np.errstate(divide="ignore", invalid="ignore"):
            ratio = np.float64(numeratorVal) / np.float64(denominatorVal)
            rhorho = float(_clamp(rho * float(np.sqrt(ratio)), MIN_VAL_RHO, MAX_VAL_RHO))
    epsPrim = epsAbs + epsRel * maxNormPrim                         # :99
    epsDual = epsAbs + epsRel * maxNormDual                         # :100
    if (normResPrim < epsPrim) and (normResDual < epsDual):         # :102
        convFlag = ConvergenceFlag.convPrimDual
    if (_norm_inf(np.abs(vX - vXP) * wX) <= epsAdmm) and (_norm_inf(np.abs(vZ - vZP) * wE) <= epsAdmm):   # :105
        convFlag = ConvergenceFlag.convAdmm
    return rhorho, convFlag, (normResPrim, normResDual)


def solve_quadratic_program(vX, mP, vQ, mA, vL, vU, LinSysSolInit, LinSysSol, *,
                            numIterations=5000, epsAbs=1e-6, epsRel=1e-6, rho=1.0, sigma=1e-6, alpha=1.6,
                            delta=1e-6, adptRho=False, fctrRho=5.0, numItrConv=25, numItrPolish=10,
                            epsMinres=1e-6, numItrMinres=500, epsPcg=None, numItrPcg=None, trace=None, scaling=None,
                            rhoScale=None, polish=False):
    """``SolveQuadraticProgram!`` (SolveQuadraticProgram.jl:14-76).  Mutates ``vX``.

    Returns ``(convFlag, info)``; the reference returns only the flag -- ``info`` carries the
    iteration count, final rho, residuals and CG iteration total for the parity tests.
    ``epsPcg`` / ``numItrPcg`` override the plugin kwargs the reference driver never forwards
    (LinearSystemSolvers.jl:125 vs SolveQuadraticProgram.jl:54); None keeps 1e-6 / 1000.
    ``delta, numItrPolish, epsMinres, numItrMinres`` are accepted and unused, as in the reference, unless
    ``polish=True`` (not a reference keyword): then ``polish_solution`` below runs after the loop.
    ``rhoScale`` (m positive factors, not in the reference: OSQP's rho vector, README.md:71-72 TODO): constraint i
    uses ``rho * rhoScale[i]`` wherever :56-61 and the plugin use the scalar; ``CheckConvergence`` and the
    adaptive update keep the scalar.  Supported by modes D, M and J.
    """
    mP = sp.csc_matrix(mP)
    mA = sp.csc_matrix(mA)
    n = vX.shape[0]
    m = mA.shape[0]
    rho = float(rho)
    rho1 = 1.0 / rho                                                # :30
    alpha1 = 1.0 - alpha                                            # :31
    convFlag = ConvergenceFlag.convNumItr                           # :33
    epsAdmm = min(epsAbs, epsRel) * 1e-2                            # :34

    vXX, vZZ, st = LinSysSolInit(vX, mP, vQ, mA, rho, rho1, sigma, n, m)   # :36
    if epsPcg is not None and "eps_pcg" in st:
        st["eps_pcg"] = epsPcg
    if numItrPcg is not None and "num_itr_pcg" in st:
        st["num_itr_pcg"] = numItrPcg
    if rhoScale is not None:
        if "mL" in st:
            raise ValueError("rhoScale is not implemented for the explicit-matrix CG plugin (mode C)")
        st["rho_scale"] = np.asarray(rhoScale, dtype=np.float64)
        st["refactor"] = True                      # direct plugin: the Init factor used the scalar rho

    vXP = np.zeros(n)                                               # :38
    vZ = np.zeros(m)
    vY = np.zeros(m)
    vZP = np.zeros(m)                                               # :41
    mAr = sp.csr_matrix(mA)
    mPr = sp.csr_matrix(mP)

    rhorho = rho                                                    # :43
    norms = (float("nan"), float("nan"))
    n_rho_updates = 0
    ii = 0
    for ii in range(1, numIterations + 1):                          # :45
        changedRho = False
        if adptRho and ((rhorho * fctrRho < rho) or (rhorho > fctrRho * rho)):   # :47
            rho = rhorho
            rho1 = 1.0 / rho
            changedRho = True
            n_rho_updates += 1

        LinSysSol(st, vXX, vZZ, vX, mP, vQ, mA, vZ, vY, rho, rho1, sigma, n, m, changedRho)   # :54

        vXP[:] = vX                                                 # :56
        vX[:] = alpha * vXX + alpha1 * vX                           # :57
        vZP[:] = vZ                                                 # :59
        rho_v, rho1_v = _rho_vec(st, rho, rho1)
        vZ[:] = _clamp(alpha * vZZ + alpha1 * vZ + rho1_v * vY, vL, vU)   # :60
        vY[:] = vY + rho_v * (alpha * vZZ + alpha1 * vZP - vZ)      # :61

        if trace is not None:
            trace(ii, vX, vZ, vY, rho)

        if ii % numItrConv == 0:                                    # :63
            rhorho, convFlag, norms = check_convergence(vX, mPr, vQ, mAr, vZ, vY, vXP, vZP, rho, rhorho,
                                                        adptRho, epsAbs, epsRel, epsAdmm, convFlag, scaling)
            if convFlag != ConvergenceFlag.convNumItr:              # :66
                break

    info = {"iterations": ii, "rho": rho, "res_prim": norms[0], "res_dual": norms[1],
            "rho_updates": n_rho_updates, "cg_iters": st.get("cg_iters", 0), "z": vZ, "y": vY}
    if polish:
        info["polish"] = polish_solution(vX, mPr, vQ, mAr, vL, vU, vZ, vY, delta, numItrPolish, epsMinres, numItrMinres)
    return convFlag, info


# --------------------------------------------------------------------------------------------
# Solution polish.  The Julia function reserves the keyword arguments and never uses them
# (SolveQuadraticProgram.jl:16-17); the algorithm is the MATLAB twin's, SolveQuadraticProgram.m:289-325:
#   active sets from the multipliers, g = [-q; l_L; u_U], K = [P A_L' A_U'; A_L 0 0; A_U 0 0],
#   KK = K + blkdiag(delta I, -delta I); numPolishItr rounds of iterative refinement
#       tt = minres(KK, g - K t, tol, maxit, x0 = tt);  stop at the first failure;  t += tt
#   and x = t[1:n] only if the last minres call converged.
# One deliberate change: the MATLAB code picks the active rows by the SIGN of y, which puts rows whose
# multiplier is rounding noise around 0 into an active set with whichever bound the noise points at; here a row is
# active only if its multiplier exceeds its distance to the bound (OSQP's rule, Stellato et al. 2020, section 5.4):
#   lower active  <=>  z_i - l_i < -y_i        upper active  <=>  u_i - z_i < y_i .
# The reduced system is kept at full size: inactive rows carry the equation -delta nu_i = 0, so their entries stay
# exactly zero and the operator is two masked passes over A and H = [P A'].
# --------------------------------------------------------------------------------------------

def minres(op, b, tol, maxit, x0):
    """MINRES (Paige & Saunders 1975) without preconditioner, stopping rule of MATLAB's ``minres``: estimated
    ``||b - A x|| <= tol ||b||``.  Returns ``(x, flag, iterations)`` with flag 0 = converged, 1 = maxit reached."""
    x = np.array(x0, dtype=np.float64)
    bnorm = float(np.sqrt(np.dot(b, b)))
    Y = [None, None, None]
    Y[0] = b - op(x)
    beta1 = float(np.sqrt(np.dot(Y[0], Y[0])))
    tolb = tol * bnorm
    if beta1 <= tolb:
        return x, 0, 0
    W = [np.zeros_like(x), np.zeros_like(x), np.zeros_like(x)]
    oldb, beta, dbar, epsln, phibar, cs, sn = 0.0, beta1, 0.0, 0.0, beta1, -1.0, 0.0
    v = Y[0] / beta
    for k in range(1, int(maxit) + 1):
        y = op(v)
        if k >= 2:
            y = y - (beta / oldb) * Y[(k - 2) % 3]
        alfa = float(np.dot(v, y))
        y = y - (alfa / beta) * Y[(k - 1) % 3]
        Y[k % 3] = y
        oldb = beta
        beta = float(np.sqrt(np.dot(y, y)))
        oldeps = epsln
        delta = cs * dbar + sn * alfa
        gbar = sn * dbar - cs * alfa
        epsln = sn * beta
        dbar = -cs * beta
        gamma = max(float(np.sqrt(gbar * gbar + beta * beta)), np.finfo(np.float64).eps)
        cs = gbar / gamma
        sn = beta / gamma
        phi = cs * phibar
        phibar = sn * phibar
        W[k % 3] = (v - oldeps * W[(k - 2) % 3] - delta * W[(k - 1) % 3]) * (1.0 / gamma)
        x = x + phi * W[k % 3]
        if phibar <= tolb or beta == 0.0:
            return x, 0, k
        v = y / beta
    return x, 1, int(maxit)


def polish_active_sets(vL, vU, vZ, vY):
    """0 = inactive, 1 = lower bound active, 2 = upper bound active."""
    lower = (vZ - vL) < -vY
    upper = ~lower & ((vU - vZ) < vY)
    return np.where(lower, 1, np.where(upper, 2, 0))


def polish_solution(vX, mP, vQ, mA, vL, vU, vZ, vY, delta, numItrPolish, epsMinres, numItrMinres):
    """Polish ``vX`` in place; returns ``{"applied", "minres_iters", "n_active"}``."""
    n, m = mP.shape[0], mA.shape[0]
    act = polish_active_sets(vL, vU, vZ, vY)
    mask = (act != 0).astype(np.float64)
    g = np.concatenate([-np.asarray(vQ, dtype=np.float64),
                        np.where(act == 1, vL, np.where(act == 2, vU, 0.0))])
    mAt = sp.csr_matrix(mA.T)

    def K(t, reg):
        top = mP @ t[:n] + mAt @ t[n:] + reg * t[:n]
        bot = mask * (mA @ t[:n]) - reg * t[n:]
        return np.concatenate([top, bot])

    t = np.zeros(n + m)
    tt = np.zeros(n + m)
    flag, total = -1, 0
    for _ in range(int(numItrPolish)):
        tt, flag, its = minres(lambda v: K(v, delta), g - K(t, 0.0), epsMinres, numItrMinres, tt)
        total += its
        if flag:
            break
        t = t + tt
    applied = flag == 0
    if applied:
        vX[:] = t[:n]
    return {"applied": bool(applied), "minres_iters": total, "n_active": int(np.count_nonzero(act)), "nu": t[n:]}


def _limit_scaling(v):
    return np.where(v < 1e-4, 1.0, np.where(v > 1e4, 1e4, v))


def _row_max_abs(M):
    """inf-norm of every row of a CSR matrix (0 for empty rows)."""
    out = np.zeros(M.shape[0])
    nz = np.diff(M.indptr) > 0
    if M.nnz:
        out[nz] = np.maximum.reduceat(np.abs(M.data), M.indptr[:-1][nz])
    return out


def ruiz_equilibrate(mP, vQ, mA, vL, vU, numItr):
    """Modified Ruiz equilibration of ``[P A'; A 0]`` with cost scaling -- Stellato et al., "OSQP: an operator
    splitting solver for quadratic programs" (2020), Algorithm 2.  NOT in the reference (its README.md:71-72
    lists scaling as a TODO): SURVEY.md 8(f) row 1.  Returns ``(P_s, q_s, A_s, l_s, u_s, D, E, c)`` with
    ``P_s = c D P D``, ``A_s = E A D``, ``q_s = c D q``, ``l_s = E l``, ``u_s = E u``.

    Per iteration: delta = 1/sqrt(limit(column inf-norms of the KKT matrix)), scale, then
    gamma = 1/max(limit(mean column norm of P_s), limit(|q_s|inf)); limit(v) = 1 if v < 1e-4, min(v, 1e4).
    Entries are scaled as ``a_ij * (d_j * e_i)`` and the mean is a left-to-right sum, like the C++ host code."""
    P = sp.csr_matrix(mP, dtype=np.float64, copy=True)
    A = sp.csr_matrix(mA, dtype=np.float64, copy=True)
    P.sort_indices(); A.sort_indices()
    n, m = P.shape[0], A.shape[0]
    q = np.array(vQ, dtype=np.float64)
    D, E, c = np.ones(n), np.ones(m), 1.0
    prow = np.repeat(np.arange(n), np.diff(P.indptr))
    arow = np.repeat(np.arange(m), np.diff(A.indptr))
    for _ in range(int(numItr)):
        At = sp.csr_matrix(A.T)
        dn = np.maximum(_row_max_abs(P), _row_max_abs(At) if m else 0.0)
        en = _row_max_abs(A)
        dn = 1.0 / np.sqrt(_limit_scaling(dn))
        en = 1.0 / np.sqrt(_limit_scaling(en))
        P.data *= dn[prow] * dn[P.indices]
        A.data *= dn[A.indices] * en[arow]
        q *= dn
        D *= dn
        E *= en
        pn = _row_max_abs(P)
        mean = float(np.cumsum(pn)[-1]) / n                      # serial sum
        mean = float(_limit_scaling(np.float64(mean)))
        qn = float(_limit_scaling(np.float64(_norm_inf(q))))
        gamma = 1.0 / max(mean, qn)
        P.data *= gamma
        q *= gamma
        c *= gamma
    return sp.csc_matrix(P), q, sp.csc_matrix(A), E * np.asarray(vL, dtype=np.float64), E * np.asarray(vU, dtype=np.float64), D, E, c


def solve(mP, vQ, mA, vL, vU, mode="D", x0=None, numItrScaling=0, **kw):
    """Convenience wrapper: ``(x, flag, info)`` for a plugin mode letter.  ``numItrScaling > 0`` iterates the
    Ruiz-equilibrated problem and tests convergence on the unscaled residuals (see ``ruiz_equilibrate``)."""
    init, sol = PLUGINS[mode]
    vX = np.zeros(mP.shape[0]) if x0 is None else np.array(x0, dtype=np.float64)
    if not numItrScaling:
        flag, info = solve_quadratic_program(vX, mP, vQ, mA, vL, vU, init, sol, **kw)
        return vX, flag, info
    Ps, qs, As, ls, us, D, E, c = ruiz_equilibrate(mP, vQ, mA, vL, vU, numItrScaling)
    vXs = vX / D
    scaling = (D, 1.0 / (c * D), 1.0 / E, _norm_inf(np.asarray(vQ, dtype=np.float64)))
    flag, info = solve_quadratic_program(vXs, Ps, qs, As, ls, us, init, sol, scaling=scaling, **kw)
    info["z"] = info["z"] / E
    info["y"] = E * info["y"] / c
    info["scaling"] = {"D": D, "E": E, "c": c}
    return vXs * D, flag, info


# --------------------------------------------------------------------------------------------
# Independent optimality certificate (replaces the absent OSQP / Gurobi cross-check of RunTests.jl)
# --------------------------------------------------------------------------------------------

def kkt_certificate(mP, vQ, mA, vL, vU, vX, vY):
    """Residuals of the QP's KKT conditions at (x, y): stationarity ||Px+q+A'y||inf, primal
    infeasibility max(l-Ax, Ax-u, 0), and the sign/complementarity violation of y."""
    vAx = mA @ vX
    stat = _norm_inf(mP @ vX + vQ + mA.T @ vY)
    pinf = float(np.max(np.maximum(np.maximum(vL - vAx, vAx - vU), 0.0))) if vAx.size else 0.0
    # y_i > 0 requires Ax_i at the upper bound; y_i < 0 at the lower bound
    gap_u = np.where(np.isfinite(vU), np.abs(vU - vAx), np.inf)
    gap_l = np.where(np.isfinite(vL), np.abs(vAx - vL), np.inf)
    yp = np.maximum(vY, 0.0)
    ym = np.maximum(-vY, 0.0)
    comp = np.maximum(np.minimum(yp, gap_u), np.minimum(ym, gap_l))   # min(|y|, distance to its bound)
    comp = float(np.max(comp)) if comp.size else 0.0
    return {"stationarity": stat, "primal_infeasibility": pinf, "complementarity": comp}
