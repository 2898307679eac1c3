"""Synthetic QP generators: the input side of the hot path.

Restates ``GenerateRandomQP`` (reference ``GenerateQuadraticProgram.jl:8-115``) with numpy
RNG streams (Julia's RNG stream cannot be reproduced outside Julia, so problems are
identified by ``(class, n, m, density, seed)`` here) and adds the density-parameterised
configurations named in ``BASELINE.json`` (cfg2..cfg5).

All matrices are returned as ``scipy.sparse.csc_matrix`` with float64 values -- the layout
Julia's ``SparseMatrixCSC{Float64,Int64}`` has -- and dense float64 vectors.

    min 0.5 x'Px + q'x   s.t.  l <= Ax <= u
"""
from __future__ import annotations

import enum

import numpy as np
import scipy.sparse as sp


class ProblemClass(enum.IntEnum):
    """Mirror of ``@enum ProblemClass`` (GenerateQuadraticProgram.jl:6), 1-based like Julia."""

    randomQp = 1
    inequalityConstrainedQp = 2
    equalityConstrainedQp = 3
    optimalControl = 4
    portfolioOptimization = 5
    lassoOptimization = 6
    huberFitting = 7
    supportVectorMachine = 8
    isotonicRegression = 9


def sprandn(rng: np.random.Generator, m: int, n: int, density: float) -> sp.csc_matrix:
    """``sprandn(m, n, density)``: every entry is non-zero with probability ``density``
    (independently), non-zeros ~ N(0,1).  Sampled as Binomial(m*n, density) positions drawn
    uniformly (duplicates dropped), which scales to m*n ~ 1e12 (cfg5)."""
    total = int(m) * int(n)
    if total == 0 or density <= 0.0:
        return sp.csc_matrix((m, n), dtype=np.float64)
    if density >= 1.0:
        return sp.csc_matrix(rng.standard_normal((m, n)))
    if total <= 4_000_000:
        mask = rng.random((m, n)) < density
        rows, cols = np.nonzero(mask)
    else:
        k = int(rng.binomial(total, density)) if total < 2**62 else int(round(total * density))
        rows = rng.integers(0, m, size=k, dtype=np.int64)
        cols = rng.integers(0, n, size=k, dtype=np.int64)
        key = np.unique(cols * np.int64(m) + rows)          # sorted column-major, deduped
        cols, rows = np.divmod(key, np.int64(m))
    vals = rng.standard_normal(rows.shape[0])
    mat = sp.csc_matrix((vals, (rows, cols)), shape=(m, n), dtype=np.float64)
    mat.sum_duplicates()
    mat.sort_indices()
    return mat


def _speye(n: int, scale: float = 1.0) -> sp.csc_matrix:
    return sp.identity(n, dtype=np.float64, format="csc") * scale


def _finish(mP, vQ, mA, vL, vU):
    mP = sp.csc_matrix(mP, dtype=np.float64)
    mA = sp.csc_matrix(mA, dtype=np.float64)
    mP.sum_duplicates(); mP.sort_indices()
    mA.sum_duplicates(); mA.sort_indices()
    return (mP, np.ascontiguousarray(vQ, dtype=np.float64), mA,
            np.ascontiguousarray(vL, dtype=np.float64), np.ascontiguousarray(vU, dtype=np.float64))


def GenerateRandomQP(problemClass: ProblemClass, numElements: int = 1000, *, numConstraints: int = 0,
                     seed: int = 1234, densityFctr: float | None = None):
    """Restatement of ``GenerateRandomQP`` (GenerateQuadraticProgram.jl:8-115).

    ``densityFctr`` overrides the class's hard-coded density (0.15 / 0.5 / 0.25) so the same
    recipe yields the sparse configurations of BASELINE.json; ``None`` keeps the reference value.
    Returns ``(mP, vQ, mA, vL, vU)``.
    """
    rng = np.random.default_rng(seed)
    pc = ProblemClass(problemClass)
    n = int(numElements)
    m = int(numConstraints)

    if pc in (ProblemClass.randomQp, ProblemClass.inequalityConstrainedQp,
              ProblemClass.equalityConstrainedQp, ProblemClass.optimalControl):
        d = 0.15 if densityFctr is None else densityFctr            # :11
        a = 1e-2                                                    # :12
        mM = sprandn(rng, n, n, d)                                  # :14
        mP = (mM.T @ mM) + _speye(n, a)                             # :15
        vQ = rng.standard_normal(n)                                 # :16
        if pc == ProblemClass.inequalityConstrainedQp:
            m = m or 10 * n                                         # :18
            mA = sprandn(rng, m, n, d)
            vL = -rng.random(m)
            vU = rng.random(m)
        elif pc == ProblemClass.equalityConstrainedQp:
            m = m or n // 2                                         # :23
            mA = sprandn(rng, m, n, d)
            vL = rng.standard_normal(m)
            vU = vL.copy()
        else:
            m = m or n // 2                                         # :28
            mA = sprandn(rng, m, n, d)
            vL = -rng.random(m)
            vU = rng.random(m)
            vI = rng.random(m) <= 0.15
            vL[vI] = vU[vI]                                         # :33  equality rows
            vI = rng.random(m) <= 0.15
            vU[vI] = 1.0                                            # :35  `vU[vI] .= vI[vI]` == true == 1.0
        return _finish(mP, vQ, mA, vL, vU)

    if pc == ProblemClass.portfolioOptimization:                    # :37-47
        d = 0.5 if densityFctr is None else densityFctr
        m = m or max(5, n // 100)
        mD = sp.diags(rng.random(n) * np.sqrt(m), format="csc")
        mP = sp.block_diag([mD, _speye(m)], format="csc")
        vQ = np.concatenate([rng.standard_normal(n), np.zeros(m)])
        mF = sprandn(rng, n, m, d)
        mA = sp.bmat([[mF.T, -_speye(m)],
                      [sp.csc_matrix(np.ones((1, n))), sp.csc_matrix((1, m))],
                      [_speye(n), sp.csc_matrix((n, m))]], format="csc")
        vL = np.concatenate([np.zeros(m), [1.0], np.zeros(n)])
        vU = np.concatenate([np.zeros(m), [1.0], np.ones(n)])
        return _finish(mP, vQ, mA, vL, vU)

    if pc == ProblemClass.lassoOptimization:                        # :48-61
        d = 0.15 if densityFctr is None else densityFctr
        m = m or n * 100
        mAd = sprandn(rng, m, n, d)
        vXX = (rng.standard_normal(n) / np.sqrt(n)) * (rng.random(n) > 0.5)
        vB = mAd @ vXX + rng.standard_normal(m)
        lam = np.linalg.norm(mAd.T @ vB, np.inf) / 5.0
        mP = sp.block_diag([sp.csc_matrix((n, n)), _speye(m, 2.0), sp.csc_matrix((n, n))], format="csc")
        vQ = np.concatenate([np.zeros(n + m), lam * np.ones(n)])
        mA = sp.bmat([[mAd, -_speye(m), sp.csc_matrix((m, n))],
                      [_speye(n), sp.csc_matrix((n, m)), -_speye(n)],
                      [_speye(n), sp.csc_matrix((n, m)), _speye(n)]], format="csc")
        vL = np.concatenate([vB, -np.inf * np.ones(n), np.zeros(n)])
        vU = np.concatenate([vB, np.zeros(n), np.inf * np.ones(n)])
        return _finish(mP, vQ, mA, vL, vU)

    if pc == ProblemClass.huberFitting:                             # :62-76
        d = 0.15 if densityFctr is None else densityFctr
        m = m or n * 100
        mAd = sprandn(rng, m, n, d)
        vXX = rng.standard_normal(n) / np.sqrt(n)
        vI = rng.random(m) < 0.95
        vB = (mAd @ vXX) + (0.5 * vI * rng.standard_normal(m)) + (10.0 * (~vI) * rng.random(m))
        mP = sp.block_diag([sp.csc_matrix((n, n)), _speye(m, 2.0), sp.csc_matrix((2 * m, 2 * m))], format="csc")
        vQ = np.concatenate([np.zeros(n + m), 2.0 * np.ones(2 * m)])
        mIm = _speye(m)
        mA = sp.vstack([sp.hstack([mAd, -mIm, -mIm, mIm]),
                        sp.hstack([sp.csc_matrix((m, n + m)), mIm, sp.csc_matrix((m, m))]),
                        sp.hstack([sp.csc_matrix((m, n + 2 * m)), mIm])], format="csc")
        vL = np.concatenate([vB, np.zeros(2 * m)])
        vU = np.concatenate([vB, np.inf * np.ones(2 * m)])
        return _finish(mP, vQ, mA, vL, vU)

    if pc == ProblemClass.supportVectorMachine:                     # :77-92
        d = 0.15 if densityFctr is None else densityFctr
        m = m or n * 100
        numClassA = m // 2
        m = 2 * numClassA  # the reference assumes an even count (vB has 2*numClassA rows)
        lam = 1.0
        vB = np.concatenate([np.ones(numClassA), -np.ones(numClassA)])
        mAu = sprandn(rng, numClassA, n, d)
        mAl = sprandn(rng, numClassA, n, d)
        mAuNz = mAu.copy(); mAuNz.data[:] = 1.0
        mAlNz = mAl.copy(); mAlNz.data[:] = 1.0
        mAd = sp.vstack([(mAu / np.sqrt(m)) + (mAuNz / m), (mAl / np.sqrt(m)) - (mAlNz / m)], format="csc")
        mP = sp.block_diag([_speye(n, 2.0), sp.csc_matrix((m, m))], format="csc")
        vQ = lam * np.concatenate([np.zeros(n), np.ones(m)])
        mA = sp.bmat([[sp.diags(vB) @ mAd, -_speye(m)],
                      [sp.csc_matrix((m, n)), _speye(m)]], format="csc")
        vL = np.concatenate([-np.inf * np.ones(m), np.zeros(m)])
        vU = np.concatenate([-np.ones(m), np.inf * np.ones(m)])
        return _finish(mP, vQ, mA, vL, vU)

    if pc == ProblemClass.isotonicRegression:                       # :93-110
        d = 0.25 if densityFctr is None else densityFctr
        a = 1e-2
        mM = sprandn(rng, n, n, d)
        mP = (mM.T @ mM) + _speye(n, a)
        vQ = rng.standard_normal(n)
        if rng.random() >= 0.5:
            mA = sp.diags([np.ones(n - 1), -np.ones(n - 1)], [0, 1], shape=(n - 1, n), format="csc")
        else:
            mA = sp.diags([-np.ones(n - 1), np.ones(n - 1)], [0, 1], shape=(n - 1, n), format="csc")
        vL = np.zeros(n - 1)
        vU = 10.0 * np.ones(n - 1)
        return _finish(mP, vQ, mA, vL, vU)

    raise ValueError(f"unknown problem class {problemClass!r}")


# ---------------------------------------------------------------------------------------
# BASELINE.json configurations
# ---------------------------------------------------------------------------------------

def config_cfg1(seed: int = 1234):
    """configs[0]: GenerateRandomQP(randomQp, 100): n=100, m=50, d=0.15."""
    return GenerateRandomQP(ProblemClass.randomQp, 100, seed=seed)


def config_sparse(n: int, m: int, density: float, seed: int = 1234, feasible: bool = True):
    """Same recipe as randomQp with the density parameterised (cfg2: n=1e4, m=2e4, d=1e-3;
    cfg5: n=1e6, m=2e6, d=5e-6).

    With m = 2n two-sided rows around 0 plus 15 % equality rows -- and, at ~5 non-zeros per row, empty
    rows of A whose equality value is not 0 -- the literal recipe is primal infeasible (ADMM then runs
    to the iteration cap with ||Ax - z||inf stuck near 1).  ``feasible=True`` (default) keeps the recipe's
    structure but centres the bounds on A x* for a random x*: l = A x* - rand, u = A x* + rand,
    equality rows l = u = A x* (15 %), and the reference's ``u = 1`` quirk becomes u = A x* + 1 (15 %),
    so x* is feasible by construction (as BASELINE.md does for cfg4)."""
    mP, vQ, mA, vL, vU = GenerateRandomQP(ProblemClass.randomQp, n, numConstraints=m, seed=seed, densityFctr=density)
    if feasible:
        rng = np.random.default_rng(seed + 7919)
        centre = mA @ rng.standard_normal(n)
        mm = mA.shape[0]
        vL = centre - rng.random(mm)
        vU = centre + rng.random(mm)
        vI = rng.random(mm) <= 0.15
        vL[vI] = centre[vI]
        vU[vI] = centre[vI]
        vI = rng.random(mm) <= 0.15
        vU[vI] = centre[vI] + 1.0
    return mP, vQ, mA, vL, vU


def config_cfg2(seed: int = 1234, feasible: bool = True):
    return config_sparse(10_000, 20_000, 1e-3, seed, feasible)


def config_cfg5(seed: int = 1234, scale: float = 1.0, feasible: bool = True):
    """configs[4]: n=1M, m=2M, d=5e-6.  ``scale`` < 1 shrinks n, m and raises the density so that
    the non-zeros per row stay the same (5 per row of M and of A) -- used by tests."""
    n = int(round(1_000_000 * scale))
    m = 2 * n
    return config_sparse(n, m, 5.0 / n, seed, feasible)


def config_cfg4(seed: int = 1234, scale: float = 1.0):
    """configs[3]: constrained least squares of the README form (README.md:22-28)

        min 0.5 ||A x - b||^2   s.t.  B x <= c,  D x = e

    with A 200000x50000, B 25000x50000, D 5000x50000, all at ~5 non-zeros per row, reformulated as
    P = A'A, q = -A'b, constraint matrix [B; D], l = [-Inf; e], u = [c; e].  Feasible by
    construction (c = B x* + rand, e = D x*)."""
    rng = np.random.default_rng(seed)
    n = int(round(50_000 * scale))
    ra, rb, rd = 4 * n, n // 2, n // 10
    d = 5.0 / n
    mAls = sprandn(rng, ra, n, d)
    mB = sprandn(rng, rb, n, d)
    mD = sprandn(rng, rd, n, d)
    xs = rng.standard_normal(n)
    vB = mAls @ xs + 0.1 * rng.standard_normal(ra)
    vC = mB @ xs + rng.random(rb)
    vE = mD @ xs
    mP = mAls.T @ mAls
    vQ = -(mAls.T @ vB)
    mA = sp.vstack([mB, mD], format="csc")
    vL = np.concatenate([-np.inf * np.ones(rb), vE])
    vU = np.concatenate([vC, vE])
    return _finish(mP, vQ, mA, vL, vU)


def config_cfg3_batch(batch: int, n: int = 64, m: int = 96, seed: int = 1234):
    """configs[2]: ``batch`` independent dense QPs (recipe with d=1), MPC-style.

    Returns dense arrays in the layout the batched C-ABI takes: ``P[batch, n, n]`` and
    ``A[batch, n, m]`` such that each problem's block is *column-major* n x n / m x n (i.e.
    ``A[b, j, i] = A_b[i, j]``), plus ``q[batch, n]``, ``l, u[batch, m]``.
    One vectorised stream seeded by ``seed`` (not one seed per problem, for generation speed)."""
    rng = np.random.default_rng(seed)
    mM = rng.standard_normal((batch, n, n))
    P = np.matmul(mM.transpose(0, 2, 1), mM)
    P += 1e-2 * np.eye(n)[None]
    P = 0.5 * (P + P.transpose(0, 2, 1))              # exactly symmetric, so col-major == row-major
    A_rm = rng.standard_normal((batch, m, n))         # row-major m x n
    q = rng.standard_normal((batch, n))
    l = -rng.random((batch, m))
    u = rng.random((batch, m))
    vI = rng.random((batch, m)) <= 0.15
    l[vI] = u[vI]
    vI = rng.random((batch, m)) <= 0.15
    u[vI] = 1.0
    A_cm = np.ascontiguousarray(A_rm.transpose(0, 2, 1))   # [b, j, i] -> column-major m x n block
    return np.ascontiguousarray(P), q, A_cm, l, u


def config_cfg3_shared(batch: int, n: int = 64, m: int = 96, seed: int = 1234):
    """configs[2] as an MPC controller sees it: ONE plant model -- one P[n, n], one A (column-major
    m x n as ``A_cm[n, m]``) -- and ``batch`` different right-hand sides q[batch, n], l, u[batch, m] (the recipe of
    ``config_cfg3_batch`` for the vectors).  ``np.broadcast_to(P, (batch, n, n))`` gives the per-problem layout."""
    rng = np.random.default_rng(seed)
    mM = rng.standard_normal((n, n))
    P = mM.T @ mM + 1e-2 * np.eye(n)
    P = 0.5 * (P + P.T)
    A_rm = rng.standard_normal((m, n))
    q = rng.standard_normal((batch, n))
    l = -rng.random((batch, m))
    u = rng.random((batch, m))
    vI = rng.random((batch, m)) <= 0.15
    l[vI] = u[vI]
    vI = rng.random((batch, m)) <= 0.15
    u[vI] = 1.0
    return np.ascontiguousarray(P), q, np.ascontiguousarray(A_rm.T), l, u


def config_banded(n: int = 1_000_000, m: int = 2_000_000, nnz_row_p: int = 26, nnz_row_a: int = 5, band: int = 256,
                  seed: int = 1234):
    """A QP with the SAME sizes and non-zeros per row as cfg5 but with every row's columns inside a band of
    +-``band`` around the (scaled) diagonal -- i.e. the gathers of x have locality.  Used only to separate the
    SpMV engine's streaming ability from cfg5's uniformly random gather pattern; P is made
    symmetric and diagonally dominant so the QP is well posed."""
    rng = np.random.default_rng(seed)

    def band_matrix(rows, cols, k):
        r = np.repeat(np.arange(rows, dtype=np.int64), k)
        centre = (r * cols) // max(rows, 1)
        c = np.clip(centre + rng.integers(-band, band + 1, size=r.shape[0]), 0, cols - 1)
        v = rng.standard_normal(r.shape[0])
        mat = sp.csr_matrix((v, (r, c)), shape=(rows, cols))
        mat.sum_duplicates()
        return mat

    T = band_matrix(n, n, max(1, nnz_row_p // 2))
    P = (T + T.T) * 0.5
    P = P + sp.diags(np.asarray(abs(P).sum(axis=1)).ravel() + 1e-2)
    A = band_matrix(m, n, nnz_row_a)
    q = rng.standard_normal(n)
    centre = A @ rng.standard_normal(n)
    l = centre - rng.random(m)
    u = centre + rng.random(m)
    return _finish(P, q, A, l, u)


def badly_scaled(problem, seed: int = 0, var_decades: float = 2.0, con_decades: float = 3.0):
    """The same QP in badly chosen units: variables x_j = d_j x'_j with d_j = 10^U(-var, var) and constraint rows
    multiplied by e_i = 10^U(-con, con):  P' = D P D, q' = D q, A' = E A D, l' = E l, u' = E u.
    ADMM with a scalar rho stalls on such a problem; Ruiz equilibration (numItrScaling) recovers it."""
    mP, vQ, mA, vL, vU = problem
    rng = np.random.default_rng(seed)
    n, m = mP.shape[0], mA.shape[0]
    d = 10.0 ** rng.uniform(-var_decades, var_decades, n)
    e = 10.0 ** rng.uniform(-con_decades, con_decades, m)
    D, E = sp.diags(d), sp.diags(e)
    return sp.csc_matrix(D @ mP @ D), d * vQ, sp.csc_matrix(E @ mA @ D), e * vL, e * vU

