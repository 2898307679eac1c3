"""CPU tests of the oracle itself (the reference's RunTests.jl strategy with the absent OSQP/Gurobi
cross-check replaced by an independent KKT certificate) and of the golden fixtures."""
import glob
import os

import numpy as np
import pytest
import scipy.sparse as sp

from oracle import c_oracle, qp_oracle
from workloads.problems import (GenerateRandomQP, ProblemClass, config_cfg1, config_cfg3_batch,
                                                  config_cfg4, config_cfg5)

# RunTests.jl:50-56
RUNTESTS_KW = dict(numIterations=50000, epsAbs=1e-7, epsRel=1e-7, rho=0.1, adptRho=True)
ABS_DEV_THR = 1e-5   # RunTests.jl:58

GOLDEN = sorted(glob.glob(os.path.join(os.path.dirname(__file__), "golden", "*.npz")))


def _dims(pc, n):
    # RunTests.jl:29-47: equality class uses m = 5 / 50, the rest the OSQP-paper defaults
    return {ProblemClass.equalityConstrainedQp: {10: 5, 100: 50}[n]}.get(pc, 0)


@pytest.mark.parametrize("pc", list(ProblemClass))
@pytest.mark.parametrize("seed", [1234, 1235, 1236])
def test_direct_mode_kkt_certificate_n10(pc, seed):
    P, q, A, l, u = GenerateRandomQP(pc, 10, numConstraints=_dims(pc, 10), seed=seed)
    empty = np.diff(sp.csr_matrix(A).indptr) == 0
    if np.any(empty & ((l > 0) | (u < 0))):
        pytest.skip("structurally infeasible instance (empty row of A with 0 outside [l, u])")
    x, flag, info = qp_oracle.solve(P, q, A, l, u, mode="D", **RUNTESTS_KW)
    assert flag != qp_oracle.ConvergenceFlag.convNumItr
    cert = qp_oracle.kkt_certificate(P, q, A, l, u, x, info["y"])
    scale = 1.0 + max(np.max(np.abs(q)), np.max(np.abs(x)))
    assert cert["stationarity"] <= 1e-5 * scale
    assert cert["primal_infeasibility"] <= 1e-5 * scale
    assert cert["complementarity"] <= 1e-4 * scale


@pytest.mark.parametrize("pc", [ProblemClass.randomQp, ProblemClass.inequalityConstrainedQp,
                                ProblemClass.equalityConstrainedQp, ProblemClass.portfolioOptimization,
                                ProblemClass.isotonicRegression])
def test_modes_agree_n100(pc):
    """Modes D (what the reference's tests run), C, M (its CG plugins) and J agree to RunTests' 1e-5."""
    P, q, A, l, u = GenerateRandomQP(pc, 100, numConstraints=_dims(pc, 100), seed=1234)
    xs = {mode: qp_oracle.solve(P, q, A, l, u, mode=mode, **RUNTESTS_KW) for mode in "DCMJ"}
    for mode in "CMJ":
        assert np.max(np.abs(xs[mode][0] - xs["D"][0])) <= ABS_DEV_THR, mode
        assert xs[mode][1] != qp_oracle.ConvergenceFlag.convNumItr


@pytest.mark.parametrize("pc", [ProblemClass.lassoOptimization, ProblemClass.huberFitting,
                                ProblemClass.supportVectorMachine])
def test_modes_agree_inf_bounds(pc):
    """The classes with +-Inf bounds (GenerateQuadraticProgram.jl:60-61,76,91-92)."""
    P, q, A, l, u = GenerateRandomQP(pc, 10, seed=1234)
    assert np.isinf(l).any() or np.isinf(u).any()
    xd = qp_oracle.solve(P, q, A, l, u, mode="D", **RUNTESTS_KW)[0]
    xm = qp_oracle.solve(P, q, A, l, u, mode="M", **RUNTESTS_KW)[0]
    assert np.max(np.abs(xm - xd)) <= ABS_DEV_THR


@pytest.mark.parametrize("precond", [0, 1])
def test_c_restatement_matches_python(precond):
    """qp_oracle.c vs qp_oracle.py (independently written; tight inner solve so the trajectory is
    well defined)."""
    for seed in (1234, 7):
        P, q, A, l, u = config_cfg1(seed)
        kw = dict(numIterations=5000, epsAbs=1e-6, epsRel=1e-6, rho=1.0, epsPcg=1e-10)
        x, f, i = qp_oracle.solve(P, q, A, l, u, mode="J" if precond else "M", **kw)
        xc, fc, ic = c_oracle.solve_sparse(P, q, A, l, u, precond=precond, **kw)
        assert int(f) == fc
        assert i["iterations"] == ic["iterations"]
        assert np.max(np.abs(x - xc)) <= 1e-9 * (1 + np.max(np.abs(x)))


def test_c_dense_batch_matches_python_direct():
    P, q, A, l, u = config_cfg3_batch(6, n=16, m=24, seed=3)
    X, flags, iters, _, rc = c_oracle.solve_dense_batch(P, q, A, l, u)
    assert rc == 0
    for b in range(6):
        x, f, i = qp_oracle.solve(sp.csc_matrix(P[b]), q[b], sp.csc_matrix(A[b].T), l[b], u[b], mode="D")
        assert int(f) == flags[b] and i["iterations"] == iters[b]
        assert np.max(np.abs(x - X[b])) <= 1e-9 * (1 + np.max(np.abs(x)))


def test_adaptive_rho_nan_is_ignored():
    """0/0 in the rho update gives NaN; Julia's clamp passes it through and the trigger's comparisons
    are false (SolveQuadraticProgram.jl:47,95)."""
    assert np.isnan(qp_oracle._clamp(float("nan"), 1e-3, 1e6))
    n = 4
    P = sp.identity(n, format="csc")
    A = sp.identity(n, format="csc")
    x, flag, info = qp_oracle.solve(P, np.zeros(n), A, -np.ones(n), np.ones(n), mode="D", adptRho=True)
    assert np.all(x == 0.0) and info["rho"] == 1.0 and flag != qp_oracle.ConvergenceFlag.convNumItr


def test_config_shapes():
    P, q, A, l, u = config_cfg5(scale=0.002)
    assert P.shape == (2000, 2000) and A.shape == (4000, 2000)
    assert abs(A.nnz / 4000 - 5) < 1.0
    P, q, A, l, u = config_cfg4(scale=0.01)
    assert P.shape == (500, 500) and A.shape == (300, 500)
    assert np.isinf(l[:250]).all() and (l[250:] == u[250:]).all()


@pytest.mark.skipif(not GOLDEN, reason="no golden fixtures")
@pytest.mark.parametrize("path", GOLDEN, ids=[os.path.basename(p) for p in GOLDEN])
def test_golden_fixtures_reproduce(path):
    """The committed fixtures (tests/golden/make_golden.py) pin the oracle against regressions."""
    g = np.load(path)
    P = sp.csc_matrix((g["P_data"], g["P_indices"], g["P_indptr"]), shape=tuple(g["P_shape"]))
    A = sp.csc_matrix((g["A_data"], g["A_indices"], g["A_indptr"]), shape=tuple(g["A_shape"]))
    kw = {k[3:]: g[k].item() for k in g.files if k.startswith("kw_")}
    if "rhoEqScale" in kw:                       # the mirrors' shorthand for a rhoScale vector
        kw["rhoScale"] = np.where(g["l"] == g["u"], float(kw.pop("rhoEqScale")), 1.0)
    mode = str(g["mode"])
    x, flag, info = qp_oracle.solve(P, g["q"], A, g["l"], g["u"], mode=mode, **kw)
    assert int(flag) == int(g["flag"])
    assert info["iterations"] == int(g["iterations"])
    assert np.max(np.abs(x - g["x"])) <= 1e-9 * (1 + np.max(np.abs(g["x"])))


# ---------------------------------------------------------------------------------------------------
# Ruiz equilibration (not in the reference; the oracle defines it, see ruiz_equilibrate)
# ---------------------------------------------------------------------------------------------------
def test_ruiz_equilibrate_balances_the_kkt_matrix():
    from workloads.problems import badly_scaled
    P, q, A, l, u = badly_scaled(config_cfg1(seed=1234), seed=3)
    Ps, qs, As, ls, us, D, E, c = qp_oracle.ruiz_equilibrate(P, q, A, l, u, 15)
    assert np.all(D > 0) and np.all(E > 0) and c > 0
    # the scaled data are exactly what the definition says (up to rounding of the repeated products)
    ref_P = (c * sp.diags(D) @ P @ sp.diags(D)).toarray()
    ref_A = (sp.diags(E) @ A @ sp.diags(D)).toarray()
    np.testing.assert_allclose(Ps.toarray(), ref_P, rtol=1e-12, atol=0)
    np.testing.assert_allclose(As.toarray(), ref_A, rtol=1e-12, atol=0)
    np.testing.assert_allclose(qs, c * D * q, rtol=1e-12)
    np.testing.assert_allclose(ls, E * l, rtol=1e-15)
    # column inf-norms of [P A'; A 0]: spread over 10 decades before (cost scaling aside), within 2x after
    K0 = sp.bmat([[P, A.T], [A, None]]).tocsc()
    K1 = sp.bmat([[Ps / c, As.T], [As, None]]).tocsc()
    n0 = np.array([np.abs(K0[:, j].data).max() for j in range(K0.shape[1])])
    n1 = np.array([np.abs(K1[:, j].data).max() for j in range(K1.shape[1])])
    assert n0.max() / n0.min() > 1e6
    assert n1.max() / n1.min() < 2.0


@pytest.mark.parametrize("mode", ["D", "J"])
def test_scaling_leaves_the_solution_unchanged(mode):
    P, q, A, l, u = config_cfg1(seed=1235)
    kw = dict(rho=0.1, adptRho=True, epsAbs=1e-7, epsRel=1e-7, numIterations=50000, epsPcg=1e-11)
    x0, f0, i0 = qp_oracle.solve(P, q, A, l, u, mode=mode, **kw)
    x1, f1, i1 = qp_oracle.solve(P, q, A, l, u, mode=mode, numItrScaling=10, **kw)
    assert int(f0) == int(f1) == 3
    assert np.max(np.abs(x0 - x1)) <= 1e-5 * (1 + np.max(np.abs(x0)))          # RunTests.jl:58 threshold
    cert = qp_oracle.kkt_certificate(P, q, A, l, u, x1, i1["y"])               # unscaled problem, unscaled y
    assert max(cert.values()) < 1e-5
    np.testing.assert_allclose(A @ x1, i1["z"], atol=1e-5)


def test_scaling_rescues_a_badly_scaled_problem():
    from workloads.problems import badly_scaled
    P, q, A, l, u = badly_scaled(config_cfg1(seed=1234), seed=0)
    kw = dict(rho=0.1, adptRho=True, numIterations=4000)
    x0, f0, i0 = qp_oracle.solve(P, q, A, l, u, mode="D", **kw)
    x1, f1, i1 = qp_oracle.solve(P, q, A, l, u, mode="D", numItrScaling=10, **kw)
    assert int(f0) == 1 and i0["iterations"] == 4000                           # stalls without equilibration
    assert int(f1) == 3 and i1["iterations"] <= 500
    # termination is tested on the UNSCALED residuals: the reference's criterion holds for the original data
    rp = np.max(np.abs(A @ x1 - i1["z"]))
    rd = np.max(np.abs(P @ x1 + q + A.T @ i1["y"]))
    eps_p = 1e-6 + 1e-6 * max(np.max(np.abs(A @ x1)), np.max(np.abs(i1["z"])))
    eps_d = 1e-6 + 1e-6 * max(np.max(np.abs(P @ x1)), np.max(np.abs(A.T @ i1["y"])), np.max(np.abs(q)))
    assert rp < 1.01 * eps_p and rd < 1.01 * eps_d


# ---------------------------------------------------------------------------------------------------
# per-constraint rho (rhoScale; OSQP's rho vector -- not in the reference)
# ---------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("pc,n", [(ProblemClass.randomQp, 100), (ProblemClass.equalityConstrainedQp, 100),
                                  (ProblemClass.supportVectorMachine, 10)])
def test_rho_scale_modes_agree_and_solution_is_unchanged(pc, n):
    """rho_i = rho * s_i changes the path of the iteration, not its fixed point: modes D and J agree with each
    other, the KKT certificate holds, and the solution equals the scalar-rho one."""
    P, q, A, l, u = GenerateRandomQP(pc, n, numConstraints=_dims(pc, n), seed=3)
    rs = np.where(l == u, 1e3, 1.0)
    rs[::3] *= 2.0                                     # an arbitrary positive vector, not only the equality rule
    kw = dict(rho=0.1, adptRho=True, numIterations=20000, epsAbs=1e-8, epsRel=1e-8)
    xd, fd, infod = qp_oracle.solve(P, q, A, l, u, mode="D", rhoScale=rs, **kw)
    xj, fj, infoj = qp_oracle.solve(P, q, A, l, u, mode="J", rhoScale=rs, epsPcg=1e-10, **kw)
    x0, f0, _ = qp_oracle.solve(P, q, A, l, u, mode="D", **kw)
    assert int(fd) != 1 and int(fd) == int(fj) and infod["iterations"] == infoj["iterations"]
    sc = 1.0 + np.max(np.abs(x0))
    assert np.max(np.abs(xd - xj)) <= 1e-7 * sc
    assert np.max(np.abs(xd - x0)) <= 1e-5 * sc
    assert max(qp_oracle.kkt_certificate(P, q, A, l, u, xd, infod["y"]).values()) < 1e-5


def test_rho_scale_of_ones_is_the_scalar_iteration():
    P, q, A, l, u = config_cfg1(seed=1234)
    kw = dict(rho=0.1, adptRho=True, numIterations=3000)
    for mode in ("D", "J"):
        x0, f0, i0 = qp_oracle.solve(P, q, A, l, u, mode=mode, **kw)
        x1, f1, i1 = qp_oracle.solve(P, q, A, l, u, mode=mode, rhoScale=np.ones(A.shape[0]), **kw)
        assert int(f0) == int(f1) and i0["iterations"] == i1["iterations"]
        assert np.max(np.abs(x0 - x1)) <= 1e-8 * (1 + np.max(np.abs(x0)))   # rounding order differs, CG amplifies it


def test_equality_rho_scale_cuts_iterations_on_an_equality_constrained_qp():
    from quadraticprogramsolver_b200.solver import equality_rho_scale
    P, q, A, l, u = GenerateRandomQP(ProblemClass.equalityConstrainedQp, 100, seed=1)
    kw = dict(rho=0.1, numIterations=20000)
    _, f0, i0 = qp_oracle.solve(P, q, A, l, u, mode="D", **kw)
    _, f1, i1 = qp_oracle.solve(P, q, A, l, u, mode="D", rhoScale=equality_rho_scale(l, u), **kw)
    assert int(f1) != 1 and i1["iterations"] < i0["iterations"]


def test_noise_triggered_exit_and_the_same_trajectory_criterion():
    """equalityConstrainedQp, n = 10, seed 1235 with RunTests.jl:50-53 settings ends on convAdmm (|dx|, |dz| <= 1e-9)
    after the iterates have stalled at rho = 1e6: three CPU restatements of the same iteration exit at three different
    checks with the same x.  tests/parity_util.py's same-trajectory criterion (used by the GPU sweeps for exactly this
    case) accepts such a pair and still rejects a wrong iterate."""
    import warnings

    from oracle import c_oracle
    from parity_util import assert_parity

    P, q, A, l, u = GenerateRandomQP(ProblemClass.equalityConstrainedQp, 10, numConstraints=5, seed=1235)
    kw = dict(numIterations=50000, epsAbs=1e-7, epsRel=1e-7, rho=0.1, adptRho=True, epsPcg=1e-11)
    xc, fc, ic = c_oracle.solve_sparse(P, q, A, l, u, precond=1, **kw)
    xj, fj, ij = qp_oracle.solve(P, q, A, l, u, mode="J", **kw)
    xd, fd, idd = qp_oracle.solve(P, q, A, l, u, mode="D", **{k: v for k, v in kw.items() if k != "epsPcg"})
    its = {int(ic["iterations"]), int(ij["iterations"]), int(idd["iterations"])}
    assert int(fc) == int(fj) == int(fd) == 2
    assert max(np.max(np.abs(xc - xj)), np.max(np.abs(xc - xd))) <= 1e-8
    if len(its) == 1:
        pytest.skip("the restatements happen to agree on this host")

    def c_at(k):
        xk, _, ik = c_oracle.solve_sparse(P, q, A, l, u, precond=1, **dict(kw, numIterations=k))
        return xk, ik

    def j_at(k):
        xk, _, ik = qp_oracle.solve(P, q, A, l, u, mode="J", **dict(kw, numIterations=k))
        return xk, ik

    with warnings.catch_warnings(record=True) as w:
        warnings.simplefilter("always")
        assert assert_parity(xc, fc, ic, xj, fj, ij, resolve_gpu=c_at, resolve_ref=j_at, rho_updates=True) == "trajectory"
    assert w and "same trajectory" in str(w[0].message)
    with pytest.raises(AssertionError):                       # without the re-solve hooks the criterion is the strict one
        assert_parity(xc, fc, ic, xj, fj, ij)
    with pytest.raises(AssertionError):                       # a wrong iterate at the common iteration is still rejected
        assert_parity(xc, fc, ic, xj, fj, ij, resolve_gpu=lambda k: (c_at(k)[0] + 1e-3, c_at(k)[1]),
                      resolve_ref=lambda k: (j_at(k)[0] - 1e-3, j_at(k)[1]))


def test_runtests_sweep_python_restatement_vs_c_restatement(capsys):
    """The RunTests-shaped sweep (RunTests.jl:62-99 settings, 9 classes x 3 seeds at n = 10, four classes at n = 100)
    between the two CPU restatements, tight inner solve: the same criterion the GPU sweep is held to -- strict parity,
    or, where the stop test fired on rounding noise, the same-trajectory criterion -- so the GPU-vs-oracle statistics of
    tests/test_gpu_parity.py::test_runtests_sweep can be read next to CPU-vs-CPU ones."""
    import warnings

    from parity_util import assert_parity

    kw = dict(numIterations=50000, epsAbs=1e-7, epsRel=1e-7, rho=0.1, adptRho=True, epsPcg=1e-11)
    cases = [(pc, 10, s) for pc in ProblemClass for s in (1234, 1235, 1236)]
    cases += [(pc, 100, 1234) for pc in (ProblemClass.randomQp, ProblemClass.equalityConstrainedQp,
                                         ProblemClass.portfolioOptimization, ProblemClass.isotonicRegression)]
    routes = {"strict": 0, "trajectory": 0}
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        for pc, n, seed in cases:
            m = (5 if n == 10 else 50) if pc == ProblemClass.equalityConstrainedQp else 0
            P, q, A, l, u = GenerateRandomQP(pc, n, numConstraints=m, seed=seed)
            xc, fc, ic = c_oracle.solve_sparse(P, q, A, l, u, precond=1, **kw)
            xj, fj, ij = qp_oracle.solve(P, q, A, l, u, mode="J", **kw)

            def c_at(k):
                xk, _, ik = c_oracle.solve_sparse(P, q, A, l, u, precond=1, **dict(kw, numIterations=k))
                return xk, ik

            def j_at(k):
                xk, _, ik = qp_oracle.solve(P, q, A, l, u, mode="J", **dict(kw, numIterations=k))
                return xk, ik

            routes[assert_parity(xc, fc, ic, xj, fj, ij, resolve_gpu=c_at, resolve_ref=j_at, rho_updates=True,
                                 what=f"{pc.name} n={n} seed={seed}")] += 1
    with capsys.disabled():
        print(f"\n[C port vs Python oracle, RunTests sweep] strict parity: {routes['strict']}, same trajectory (noise-triggered exit): "
              f"{routes['trajectory']} of {len(cases)}")
    assert routes["trajectory"] <= 3


def test_qp_model_mat_file_round_trip(tmp_path):
    """QpModel.mat, the file the reference's MATLAB and Julia scripts exchange problems through
    (SolveQuadraticProgramUnitTest.m:84, SolveQuadraticProgramUnitTest.jl:47-55): write, read back bit for bit
    (incl. +-Inf bounds and a problem without constraints), and solve what was read."""
    import scipy.io

    from workloads.matfile import load_qp_model, save_qp_model

    for k, prob in enumerate((GenerateRandomQP(ProblemClass.supportVectorMachine, 10, seed=3), config_cfg1(1234),
                              (sp.identity(4, format="csc") * 2.0, np.ones(4), sp.csc_matrix((0, 4)), np.zeros(0), np.zeros(0)))):
        path = str(tmp_path / f"QpModel{k}.mat")
        save_qp_model(path, *prob)
        P, q, A, l, u = load_qp_model(path)
        assert (P != sp.csc_matrix(prob[0])).nnz == 0 and (A != sp.csc_matrix(prob[2])).nnz == 0
        assert np.array_equal(q, prob[1]) and np.array_equal(l, prob[3]) and np.array_equal(u, prob[4])
        assert q.ndim == 1 and l.ndim == 1
    P, q, A, l, u = load_qp_model(str(tmp_path / "QpModel1.mat"))
    x, flag, info = qp_oracle.solve(P, q, A, l, u, mode="D")
    x0, flag0, _ = qp_oracle.solve(*config_cfg1(1234), mode="D")
    assert int(flag) == int(flag0) and np.array_equal(x, x0)
    scipy.io.savemat(str(tmp_path / "bad.mat"), {"mP": np.eye(2)})
    with pytest.raises(KeyError):
        load_qp_model(str(tmp_path / "bad.mat"))
    # full (dense) matrices in the file, as MATLAB writes them when the caller never made them sparse
    scipy.io.savemat(str(tmp_path / "full.mat"), {"mP": np.eye(3), "vQ": np.ones((3, 1)), "mA": np.ones((2, 3)),
                                                  "vL": -np.ones((2, 1)), "vU": np.array([[1.0], [np.inf]])})
    P, q, A, l, u = load_qp_model(str(tmp_path / "full.mat"))
    assert sp.issparse(P) and A.shape == (2, 3) and np.isinf(u[1]) and q.shape == (3,)
