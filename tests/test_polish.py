"""Solution polish (MATLAB twin SolveQuadraticProgram.m:289-325): CPU tests of the oracle's
restatement, GPU parity of polish_kernels.cuh against it."""
import numpy as np
import pytest
import scipy.sparse.linalg as spla

from oracle import qp_oracle
from workloads.problems import GenerateRandomQP, ProblemClass, config_cfg1, config_sparse

LOOSE = dict(epsAbs=1e-4, epsRel=1e-4)
CASES = {"cfg1": lambda: config_cfg1(seed=1234),
         "svm_inf_bounds": lambda: GenerateRandomQP(ProblemClass.supportVectorMachine, 10, seed=3),
         "lasso_inf_bounds": lambda: GenerateRandomQP(ProblemClass.lassoOptimization, 10, seed=3),
         "equality": lambda: GenerateRandomQP(ProblemClass.equalityConstrainedQp, 100, numConstraints=50, seed=5)}


def test_minres_matches_scipy_on_an_indefinite_system():
    rng = np.random.default_rng(0)
    n = 60
    M = rng.standard_normal((n, n))
    S = M + M.T + np.diag(np.r_[np.full(n // 2, 8.0), np.full(n - n // 2, -8.0)])     # symmetric indefinite
    b = rng.standard_normal(n)
    x, flag, its = qp_oracle.minres(lambda v: S @ v, b, 1e-12, 500, np.zeros(n))
    assert flag == 0 and its <= 500
    assert np.linalg.norm(S @ x - b) <= 1e-9 * np.linalg.norm(b)
    xs, info = spla.minres(S, b, rtol=1e-12, maxiter=500)
    assert np.max(np.abs(x - xs)) <= 1e-7
    x1, flag1, its1 = qp_oracle.minres(lambda v: S @ v, b, 1e-12, 3, np.zeros(n))
    assert flag1 == 1 and its1 == 3                                                     # iteration cap -> flag 1
    x2, flag2, its2 = qp_oracle.minres(lambda v: S @ v, b, 1e-6, 500, x)                # converged start point
    assert flag2 == 0 and its2 == 0 and np.array_equal(x2, x)


@pytest.mark.parametrize("case", sorted(CASES))
def test_oracle_polish_recovers_the_tight_solution(case):
    """A loose ADMM solve (1e-4) + polish lands on the solution a 1e-10 solve finds."""
    P, q, A, l, u = CASES[case]()
    x0, f0, i0 = qp_oracle.solve(P, q, A, l, u, mode="D", **LOOSE)
    x1, f1, i1 = qp_oracle.solve(P, q, A, l, u, mode="D", polish=True, **LOOSE)
    xs, fs, _ = qp_oracle.solve(P, q, A, l, u, mode="D", epsAbs=1e-10, epsRel=1e-10, numIterations=200000)
    assert i1["polish"]["applied"]
    e0, e1 = np.max(np.abs(x0 - xs)), np.max(np.abs(x1 - xs))
    assert e1 <= 1e-7 * (1 + np.max(np.abs(xs))) and e1 < 1e-2 * e0
    assert int(f0) == int(f1) and i0["iterations"] == i1["iterations"]                  # the ADMM loop is untouched


def test_oracle_polish_active_set_rule_and_failure_leaves_x_untouched():
    l = np.array([0.0, 0.0, 0.0, -np.inf, 1.0])
    u = np.array([1.0, 1.0, 1.0, 2.0, 1.0])
    z = np.array([0.0, 1.0, 0.5, 2.0, 1.0])
    y = np.array([-0.3, 0.2, -1e-17, 1e-17, 0.0])       # noise-sized multipliers do not activate a row
    assert qp_oracle.polish_active_sets(l, u, z, y).tolist() == [1, 2, 0, 2, 0]
    P, q, A, l, u = config_cfg1(seed=1234)
    x0, _, _ = qp_oracle.solve(P, q, A, l, u, mode="D", **LOOSE)
    x1, _, i1 = qp_oracle.solve(P, q, A, l, u, mode="D", polish=True, numItrMinres=3, **LOOSE)
    assert not i1["polish"]["applied"] and np.array_equal(x0, x1)


# ---------------------------------------------------------------------------------------------------
# GPU parity (through the C ABI): same active set, same polished x
# ---------------------------------------------------------------------------------------------------
@pytest.mark.gpu
@pytest.mark.parametrize("lin", ["cholesky", "pcg"])
@pytest.mark.parametrize("case", sorted(CASES))
def test_gpu_polish_matches_oracle(lib, case, lin):
    from quadraticprogramsolver_b200 import solver as S
    P, q, A, l, u = CASES[case]()
    kw = dict(LOOSE, polish=True, epsMinres=1e-9, numItrMinres=2000)
    extra = dict(linSolver="cholesky") if lin == "cholesky" else dict(epsPcg=1e-11)
    x, flag, info = S.SolveQuadraticProgram(P, q, A, l, u, **kw, **extra)
    xr, fr, ir = qp_oracle.solve(P, q, A, l, u, mode="D" if lin == "cholesky" else "J", **kw,
                                 **({} if lin == "cholesky" else dict(epsPcg=1e-11)))
    assert int(flag) == int(fr) and info["iterations"] == ir["iterations"]
    assert info["polish_status"] == (1 if ir["polish"]["applied"] else 2)
    assert info["polish_active"] == ir["polish"]["n_active"]
    assert np.max(np.abs(x - xr)) <= 1e-6 * (1 + np.max(np.abs(xr)))
    xs, _, _ = qp_oracle.solve(P, q, A, l, u, mode="D", epsAbs=1e-10, epsRel=1e-10, numIterations=200000)
    assert np.max(np.abs(x - xs)) <= 1e-6 * (1 + np.max(np.abs(xs)))                    # and it is the tight solution


@pytest.mark.gpu
def test_gpu_polish_off_by_default_failure_and_sparse_problem(lib):
    from quadraticprogramsolver_b200 import solver as S
    P, q, A, l, u = config_cfg1(seed=1234)
    x0, f0, i0 = S.SolveQuadraticProgram(P, q, A, l, u, linSolver="cholesky", **LOOSE)
    assert i0["polish_status"] == 0                                                     # the reference's behaviour
    x1, f1, i1 = S.SolveQuadraticProgram(P, q, A, l, u, linSolver="cholesky", polish=True, numItrMinres=3, **LOOSE)
    assert i1["polish_status"] == 2 and np.array_equal(x0, x1)                          # MINRES failed: x untouched
    # a sparse problem with 7 000 rows, equilibrated: polish runs on the scaled problem, x comes back unscaled
    P, q, A, l, u = config_sparse(3000, 4000, 2e-3, seed=9)
    kw = dict(epsAbs=1e-4, epsRel=1e-4, numItrScaling=10, polish=True, epsMinres=1e-9, numItrMinres=5000, epsPcg=1e-11)
    x, flag, info = S.SolveQuadraticProgram(P, q, A, l, u, **kw)
    xr, fr, ir = qp_oracle.solve(P, q, A, l, u, mode="J", **kw)
    assert info["polish_status"] == (1 if ir["polish"]["applied"] else 2)
    assert info["polish_active"] == ir["polish"]["n_active"]
    assert np.max(np.abs(x - xr)) <= 1e-6 * (1 + np.max(np.abs(xr)))
