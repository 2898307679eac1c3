"""GPU parity of the EXACT x~ step (linSolver = "cholesky" on a sparse handle: dense K inverted on the device,
csrc/direct_kernels.cuh) against oracle mode D -- the reference's direct plugins (LaLdl!/QDLdl!/FacLdl!,
LinearSystemSolvers.jl:16-107), i.e. the configuration the reference's own tests run (RunTests.jl:55-56, settings
RunTests.jl:50-53).  Criterion: same flag, iteration count within 2 (equal in practice), the same number of
refactorisations, ||x - x_ref||inf <= 1e-6 (1 + ||x_ref||inf)."""
import numpy as np
import pytest
import scipy.sparse as sp

from oracle import qp_oracle
from parity_util import assert_parity
from workloads.problems import GenerateRandomQP, ProblemClass, config_cfg1, config_sparse

pytestmark = pytest.mark.gpu

RUNTESTS_KW = dict(numIterations=50000, epsAbs=1e-7, epsRel=1e-7, rho=0.1, adptRho=True)   # RunTests.jl:50-53


def _solver():
    from quadraticprogramsolver_b200 import solver
    return solver


def _check(x, flag, info, xr, fr, ir, tol=1e-6):
    assert_parity(x, flag, info, xr, fr, ir, tol=tol)                   # strict criterion (tests/parity_util.py)


def _check_sweep(S, prob, x, flag, info, xr, fr, ir, what):
    """Strict parity incl. the refactorisation count; where the exit checks differ because the stop test fired on
    rounding noise, the same-trajectory criterion of tests/parity_util.py."""
    def gpu_at(k):
        xk, _, ik = S.SolveQuadraticProgram(*prob, linSolver="cholesky", **dict(RUNTESTS_KW, numIterations=k))
        return xk, ik

    def ref_at(k):
        xk, _, ik = qp_oracle.solve(*prob, mode="D", **dict(RUNTESTS_KW, numIterations=k))
        return xk, ik

    assert_parity(x, flag, info, xr, fr, ir, resolve_gpu=gpu_at, resolve_ref=ref_at, rho_updates=True, what=what)


def _problem(pc, n, seed):
    m = (5 if n == 10 else 50) if pc == ProblemClass.equalityConstrainedQp else 0
    return GenerateRandomQP(pc, n, numConstraints=m, seed=seed)


@pytest.mark.parametrize("seed", [1234, 1235, 1236])
@pytest.mark.parametrize("pc", list(ProblemClass))
def test_runtests_sweep_n10_exact_solve(lib, pc, seed):
    S = _solver()
    P, q, A, l, u = _problem(pc, 10, seed)
    x, flag, info = S.SolveQuadraticProgram(P, q, A, l, u, linSolver="cholesky", **RUNTESTS_KW)
    xr, fr, ir = qp_oracle.solve(P, q, A, l, u, mode="D", **RUNTESTS_KW)
    _check_sweep(S, (P, q, A, l, u), x, flag, info, xr, fr, ir, f"{pc.name} n=10 seed={seed} (exact solve):")
    assert info["pcg_iters_total"] == 0


@pytest.mark.parametrize("seed", [1234, 1235])
@pytest.mark.parametrize("pc", [pc for pc in ProblemClass if pc != ProblemClass.huberFitting])
def test_runtests_sweep_n100_exact_solve(lib, pc, seed):
    """n = 100 of RunTests.jl:62-99: up to 10 200 variables (lasso) -> an 833 MB dense inverse on the device."""
    if seed != 1234 and pc in (ProblemClass.lassoOptimization, ProblemClass.supportVectorMachine):
        pytest.skip("one seed for the 10k-variable classes (the oracle's Python loop dominates the run time)")
    S = _solver()
    P, q, A, l, u = _problem(pc, 100, seed)
    x, flag, info = S.SolveQuadraticProgram(P, q, A, l, u, linSolver="cholesky", **RUNTESTS_KW)
    xr, fr, ir = qp_oracle.solve(P, q, A, l, u, mode="D", **RUNTESTS_KW)
    _check_sweep(S, (P, q, A, l, u), x, flag, info, xr, fr, ir, f"{pc.name} n=100 seed={seed} (exact solve):")


@pytest.mark.parametrize("seed", [1234, 1235, 1236])
def test_cfg1_default_settings_exact_solve(lib, seed):
    """configs[0] with the reference's default keyword arguments (SolveQuadraticProgram.jl:15-17)."""
    S = _solver()
    P, q, A, l, u = config_cfg1(seed=seed)
    x, flag, info = S.SolveQuadraticProgram(P, q, A, l, u, linSolver="cholesky")
    xr, fr, ir = qp_oracle.solve(P, q, A, l, u, mode="D")
    _check(x, flag, info, xr, fr, ir)
    z, y = info["z"], info["y"]
    assert np.max(np.abs(z - ir["z"])) <= 1e-6 * (1 + np.max(np.abs(ir["z"])))
    assert np.max(np.abs(y - ir["y"])) <= 1e-5 * (1 + np.max(np.abs(ir["y"])))


def test_exact_solve_agrees_with_tight_pcg_and_reuses_the_factor(lib):
    S = _solver()
    P, q, A, l, u = config_sparse(1200, 1800, 5e-3, seed=5)
    kw = dict(numIterations=1000, rho=0.1, adptRho=True)
    with S.QPB200Solver(P, q, A, l, u, linSolver="cholesky", **kw) as s:
        x1 = np.zeros(P.shape[0])
        f1 = s.solve(x1)
        i1 = dict(s.info)
        x2 = np.zeros(P.shape[0])
        f2 = s.solve(x2)
        i2 = dict(s.info)
    assert np.array_equal(x1, x2) and int(f1) == int(f2) and i1["iterations"] == i2["iterations"]   # bit-reproducible
    xp, fp, ip = S.SolveQuadraticProgram(P, q, A, l, u, epsPcg=1e-12, **kw)
    assert int(f1) == int(fp) and abs(i1["iterations"] - ip["iterations"]) <= 25
    assert np.max(np.abs(x1 - xp)) <= 1e-5 * (1 + np.max(np.abs(xp)))
    xr, fr, ir = qp_oracle.solve(P, q, A, l, u, mode="D", **kw)
    _check(x1, f1, i1, xr, fr, ir)
    assert i1["rho_updates"] == ir["rho_updates"]


def test_exact_solve_with_equilibration_and_rho_vector(lib):
    S = _solver()
    P, q, A, l, u = GenerateRandomQP(ProblemClass.equalityConstrainedQp, 100, numConstraints=50, seed=7)
    kw = dict(numIterations=4000, epsAbs=1e-7, epsRel=1e-7)
    rs = S.equality_rho_scale(l, u, 1e3)
    x, flag, info = S.SolveQuadraticProgram(P, q, A, l, u, linSolver="cholesky", rhoScale=rs, numItrScaling=10, **kw)
    xr, fr, ir = qp_oracle.solve(P, q, A, l, u, mode="D", rhoScale=rs, numItrScaling=10, **kw)
    _check(x, flag, info, xr, fr, ir)


def test_exact_solve_argument_and_breakdown_errors(lib):
    S = _solver()
    n = 40000                                         # beyond the dense-inverse limit
    P = sp.identity(n, format="csc")
    A = sp.identity(n, format="csc")
    with pytest.raises(S.QPB200Error):
        S.QPB200Solver(P, np.zeros(n), A, -np.ones(n), np.ones(n), linSolver="cholesky")
    n = 50                                            # K indefinite: the sweep meets a non-positive pivot
    P = sp.identity(n, format="csc") * -5.0
    A = sp.identity(n, format="csc")
    with S.QPB200Solver(P, np.ones(n), A, -np.ones(n), np.ones(n), linSolver="cholesky") as s:
        with pytest.raises(S.QPB200Error):
            s.solve(np.zeros(n))


def test_huber_n100_exact_solve_30k_variables(lib):
    """The largest RunTests problem: huber fitting at n = 100 has 30 100 variables -> a 7.3 GB dense inverse,
    refactorised at every rho change."""
    S = _solver()
    P, q, A, l, u = _problem(ProblemClass.huberFitting, 100, 1234)
    x, flag, info = S.SolveQuadraticProgram(P, q, A, l, u, linSolver="cholesky", **RUNTESTS_KW)
    xr, fr, ir = qp_oracle.solve(P, q, A, l, u, mode="D", **RUNTESTS_KW)
    _check(x, flag, info, xr, fr, ir)
    assert info["rho_updates"] == ir["rho_updates"]
    print(f"huber n=100: {info['iterations']} iterations, {info['rho_updates']} refactorisations, device {info['solve_ms']:.0f} ms")
