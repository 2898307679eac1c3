"""GPU parity tests (run with `-m gpu` on a B200): every call goes through the C ABI of libqpb200.so.

Criterion (BASELINE.json north_star): ||x - x_ref||inf <= 1e-6 (1 + ||x_ref||inf) in Float64, same
convergence flag, iteration count within +-2 (checks happen every 25 iterations, so in practice equal).
x_ref comes from the oracle in the matching linear-solver mode: M (the reference's matrix-free CG)
for precond="none", J for precond="jacobi".  Unless a test says otherwise the inner solve is tight
(epsPcg 1e-10): with the reference's loose default (1e-6) two *CPU* implementations of the same
algorithm already disagree on the exit check ("iteration-count parity").
"""
import glob
import os

import numpy as np
import pytest
import scipy.sparse as sp

from oracle import c_oracle, qp_oracle
from parity_util import assert_parity
from workloads.problems import (GenerateRandomQP, ProblemClass, config_cfg1, config_cfg4,
                                                  config_cfg5, config_sparse, sprandn)

pytestmark = pytest.mark.gpu

GOLDEN = sorted(glob.glob(os.path.join(os.path.dirname(__file__), "golden", "*.npz")))
LOADERS = ["ldg", "tma"]


def _solver():
    from quadraticprogramsolver_b200 import solver
    return solver


def _assert_parity(x, flag, info, x_ref, flag_ref, it_ref, tol=1e-6):
    assert_parity(x, flag, info, x_ref, flag_ref, it_ref, tol=tol)      # strict criterion (tests/parity_util.py)


# ---------------------------------------------------------------------------------------------------
# operators: P x, A x, A' x, K x  (bit-for-bit is not defined for FP64 sums of different order: 1e-12)
# ---------------------------------------------------------------------------------------------------
def _op_case(name):
    rng = np.random.default_rng(11)
    if name == "cfg1":
        return config_cfg1()
    if name == "sparse_20k":
        return config_sparse(20000, 40000, 2.5e-4, seed=3)
    if name == "long_rows":      # P rows longer than a tile (dense 5000 x 5000 rows > 2048 nnz)
        n, m = 2500, 300
        M = rng.standard_normal((n, n)) / np.sqrt(n)
        P = sp.csc_matrix(M.T @ M + 0.01 * np.eye(n))
        A = sp.csc_matrix(rng.standard_normal((m, n)))
        return P, rng.standard_normal(n), A, -np.ones(m), np.ones(m)
    if name == "empty_rows_cols":
        n, m = 3000, 5000
        P = sp.identity(n, format="csc") * 2.0
        A = sprandn(rng, m, n, 2e-4)     # most rows and many columns of A are empty
        return P, rng.standard_normal(n), A, -np.ones(m), np.ones(m)
    if name == "no_constraints_m0":
        n = 500
        M = sprandn(rng, n, n, 0.02)
        return (M.T @ M + sp.identity(n)).tocsc(), rng.standard_normal(n), sp.csc_matrix((0, n)), np.zeros(0), np.zeros(0)
    raise KeyError(name)


@pytest.mark.parametrize("loader", LOADERS)
@pytest.mark.parametrize("case", ["cfg1", "sparse_20k", "long_rows", "empty_rows_cols", "no_constraints_m0"])
def test_operators_match_scipy(lib, case, loader):
    S = _solver()
    P, q, A, l, u = _op_case(case)
    n, m = P.shape[0], A.shape[0]
    rng = np.random.default_rng(1)
    x = rng.standard_normal(n)
    w = rng.standard_normal(m)
    rho, sigma = 0.7, 1e-3
    with S.QPB200Solver(P, q, A, l, u, spmvLoader=loader, rho=rho, sigma=sigma) as s:
        def close(a, b):
            scale = np.max(np.abs(b)) if b.size else 1.0
            assert a.shape == b.shape
            assert np.max(np.abs(a - b), initial=0.0) <= 1e-12 * (1.0 + scale)
        close(s.apply(0, x), P @ x)
        close(s.apply(1, x), A @ x)
        close(s.apply(2, w), A.T @ w)
        close(s.apply(3, x), P @ x + rho * (A.T @ (A @ x)) + sigma * x)


# ---------------------------------------------------------------------------------------------------
# whole solves vs the oracle
# ---------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("loader", LOADERS)
@pytest.mark.parametrize("precond,mode", [("none", "M"), ("jacobi", "J")])
@pytest.mark.parametrize("seed", [1234, 1235, 1236])
def test_cfg1_default_settings(lib, seed, precond, mode, loader):
    """configs[0]: GenerateRandomQP(randomQp, 100), reference defaults (rho = 1, adaptive off)."""
    S = _solver()
    P, q, A, l, u = config_cfg1(seed)
    kw = dict(epsPcg=1e-10)
    x, flag, info = S.SolveQuadraticProgram(P, q, A, l, u, precond=precond, spmvLoader=loader, **kw)
    x_ref, flag_ref, info_ref = qp_oracle.solve(P, q, A, l, u, mode=mode, **kw)
    _assert_parity(x, flag, info, x_ref, flag_ref, info_ref["iterations"])
    assert np.max(np.abs(info["z"] - info_ref["z"])) <= 1e-6 * (1 + np.max(np.abs(info_ref["z"])))
    assert np.max(np.abs(info["y"] - info_ref["y"])) <= 1e-5 * (1 + np.max(np.abs(info_ref["y"])))


@pytest.mark.parametrize("seed", [1234, 1235])
def test_cfg1_runtests_settings_adaptive_rho(lib, seed):
    """RunTests.jl:50-56: rho = 0.1, adaptive rho on, eps 1e-7, 50000 iterations."""
    S = _solver()
    P, q, A, l, u = config_cfg1(seed)
    kw = dict(numIterations=50000, ϵAbs=1e-7, ϵRel=1e-7, ρ=0.1, adptΡ=True, epsPcg=1e-11)
    x = np.zeros(P.shape[0])
    flag = S.SolveQuadraticProgram_(x, P, q, A, l, u, S.B200Init, S.B200Sol, **kw)
    okw = dict(numIterations=50000, epsAbs=1e-7, epsRel=1e-7, rho=0.1, adptRho=True, epsPcg=1e-11)
    x_ref, flag_ref, info_ref = qp_oracle.solve(P, q, A, l, u, mode="J", **okw)
    assert int(flag) == int(flag_ref)
    assert np.max(np.abs(x - x_ref)) <= 1e-6 * (1 + np.max(np.abs(x_ref)))
    # and against the exact-solve plugin the reference's own test runs, at its own threshold 1e-5
    x_d, _, _ = qp_oracle.solve(P, q, A, l, u, mode="D", **okw)
    assert np.max(np.abs(x - x_d)) <= 1e-5


@pytest.mark.parametrize("pc", list(ProblemClass))
def test_all_problem_classes_n10(lib, pc):
    """The nine generator classes at RunTests' small size (incl. +-Inf bounds, equality rows)."""
    S = _solver()
    m = 5 if pc == ProblemClass.equalityConstrainedQp else 0
    P, q, A, l, u = GenerateRandomQP(pc, 10, numConstraints=m, seed=1234)
    kw = dict(numIterations=50000, epsAbs=1e-7, epsRel=1e-7, rho=0.1, adptRho=True, epsPcg=1e-11)
    x, flag, info = S.SolveQuadraticProgram(P, q, A, l, u, **kw)
    x_ref, flag_ref, info_ref = qp_oracle.solve(P, q, A, l, u, mode="J", **kw)

    def gpu_at(k):
        xk, _, ik = S.SolveQuadraticProgram(P, q, A, l, u, **dict(kw, numIterations=k))
        return xk, ik

    def ref_at(k):
        xk, _, ik = qp_oracle.solve(P, q, A, l, u, mode="J", **dict(kw, numIterations=k))
        return xk, ik

    assert_parity(x, flag, info, x_ref, flag_ref, info_ref, resolve_gpu=gpu_at, resolve_ref=ref_at, rho_updates=True,
                  what=f"{pc.name} n=10:")


@pytest.mark.skipif(not GOLDEN, reason="no golden fixtures")
@pytest.mark.parametrize("path", GOLDEN, ids=[os.path.basename(p) for p in GOLDEN])
def test_golden_fixtures(lib, path):
    S = _solver()
    g = np.load(path)
    P = sp.csc_matrix((g["P_data"], g["P_indices"], g["P_indptr"]), shape=tuple(g["P_shape"]))
    A = sp.csc_matrix((g["A_data"], g["A_indices"], g["A_indptr"]), shape=tuple(g["A_shape"]))
    kw = {k[3:]: g[k].item() for k in g.files if k.startswith("kw_")}
    precond = {"M": "none", "J": "jacobi"}[str(g["mode"])]
    x, flag, info = S.SolveQuadraticProgram(P, g["q"], A, g["l"], g["u"], precond=precond, **kw)
    if int(g["flag"]) == 1:
        # did not converge within numIterations: the iterate is mid-trajectory, compare loosely
        assert int(flag) == 1 and info["iterations"] == int(g["iterations"])
        assert np.max(np.abs(x - g["x"])) <= 1e-4 * (1 + np.max(np.abs(g["x"])))
    else:
        _assert_parity(x, flag, info, g["x"], int(g["flag"]), int(g["iterations"]))


def test_reference_default_inner_tolerance(lib):
    """The reference's own inner tolerance (cg! abstol = 1e-6, LinearSystemSolvers.jl:125) with the
    un-preconditioned CG: solutions agree to the reference test threshold; exit checks may differ by one
    check interval (the two CPU oracles already do -- DESIGN.md)."""
    S = _solver()
    P, q, A, l, u = config_cfg1(1234)
    x, flag, info = S.SolveQuadraticProgram(P, q, A, l, u, precond="none")
    x_ref, flag_ref, info_ref = qp_oracle.solve(P, q, A, l, u, mode="M")
    assert int(flag) != 1 and int(flag_ref) != 1
    assert abs(info["iterations"] - info_ref["iterations"]) <= 25
    assert np.max(np.abs(x - x_ref)) <= 1e-5


def test_start_point_is_used_and_x_is_mutated_in_place(lib):
    S = _solver()
    P, q, A, l, u = config_cfg1(1234)
    kw = dict(epsPcg=1e-10)
    x0 = np.random.default_rng(0).standard_normal(P.shape[0])
    x = x0.copy()
    flag = S.SolveQuadraticProgram_(x, P, q, A, l, u, **kw)
    x_ref, flag_ref, info_ref = qp_oracle.solve(P, q, A, l, u, mode="J", x0=x0, **kw)
    assert int(flag) == int(flag_ref)
    assert np.max(np.abs(x - x_ref)) <= 1e-6 * (1 + np.max(np.abs(x_ref)))
    assert not np.array_equal(x, x0)


def test_iteration_cap_returns_convNumItr(lib):
    S = _solver()
    P, q, A, l, u = config_cfg1(1234)
    x, flag, info = S.SolveQuadraticProgram(P, q, A, l, u, numIterations=30, epsPcg=1e-10)
    x_ref, flag_ref, info_ref = qp_oracle.solve(P, q, A, l, u, mode="J", numIterations=30, epsPcg=1e-10)
    assert int(flag) == 1 == int(flag_ref) and info["iterations"] == 30
    assert np.max(np.abs(x - x_ref)) <= 1e-6 * (1 + np.max(np.abs(x_ref)))


def test_bitwise_reproducible_and_handle_reuse(lib):
    """Static tile ownership + fixed-order reductions: two solves give identical bits; a handle can be
    re-solved and its vectors updated (qpb200_update_vectors)."""
    S = _solver()
    P, q, A, l, u = config_sparse(3000, 6000, 2e-3, seed=5)
    with S.QPB200Solver(P, q, A, l, u, numIterations=200) as s:
        x1 = np.zeros(3000); s.solve(x1); i1 = dict(s.info)
        x2 = np.zeros(3000); s.solve(x2); i2 = dict(s.info)
        assert np.array_equal(x1, x2) and i1["iterations"] == i2["iterations"]
        assert i1["pcg_iters_total"] == i2["pcg_iters_total"]
        q2 = q * 0.5
        s.update_vectors(vQ=q2)
        x3 = np.zeros(3000); s.solve(x3)
    with S.QPB200Solver(P, q2, A, l, u, numIterations=200) as s:
        x4 = np.zeros(3000); s.solve(x4)
    assert np.array_equal(x3, x4)


@pytest.mark.parametrize("cfg", ["cfg2", "cfg4_small", "cfg5_small"])
def test_sparse_configs_against_c_oracle_and_kkt(lib, cfg):
    """Mid-size sparse configurations: parity with the compiled oracle (same mode J, tight inner solve)
    and the size-independent KKT optimality certificate."""
    S = _solver()
    if cfg == "cfg2":
        P, q, A, l, u = config_sparse(10000, 20000, 1e-3, seed=1234)
    elif cfg == "cfg4_small":
        P, q, A, l, u = config_cfg4(seed=1234, scale=0.04)
    else:
        P, q, A, l, u = config_cfg5(seed=1234, scale=0.02)
    kw = dict(numIterations=2000, epsPcg=1e-10)
    x, flag, info = S.SolveQuadraticProgram(P, q, A, l, u, **kw)
    xc, fc, ic = c_oracle.solve_sparse(P, q, A, l, u, precond=1, **kw)
    _assert_parity(x, flag, info, xc, fc, ic["iterations"])
    if int(flag) != 1:
        cert = qp_oracle.kkt_certificate(P, q, A, l, u, x, info["y"])
        scale = 1.0 + max(np.max(np.abs(q)), np.max(np.abs(x)))
        assert cert["stationarity"] <= 1e-4 * scale
        assert cert["primal_infeasibility"] <= 1e-4 * scale


def test_pcg_reduces_residual_of_kkt_system(lib):
    """K x~ = rhs solved by the device PCG: one ADMM iteration from x = 0, z = y = 0 gives
    x = alpha * K^{-1}(-q) (SolveQuadraticProgram.jl:57, LinearSystemSolvers.jl:178-181)."""
    S = _solver()
    P, q, A, l, u = config_sparse(5000, 10000, 1e-3, seed=9)
    rho, sigma, alpha = 1.0, 1e-6, 1.6
    x, flag, info = S.SolveQuadraticProgram(P, q, A, l, u, numIterations=1, epsPcg=1e-11, numItrPcg=5000)
    K = (P + sigma * sp.identity(5000) + rho * (A.T @ A)).tocsc()
    import scipy.sparse.linalg as spla
    xt = spla.spsolve(K, -q)
    assert np.max(np.abs(x - alpha * xt)) <= 1e-8 * (1 + np.max(np.abs(xt)))


def test_one_based_indices_like_julia(lib):
    """The Julia shim passes SparseMatrixCSC fields as they are (1-based Int64, index_base = 1): the raw C-ABI
    call with shifted arrays must give bit-identical results to the 0-based call."""
    import ctypes as C
    from quadraticprogramsolver_b200 import _lib
    S = _solver()
    P, q, A, l, u = config_cfg1(1236)
    n, m = P.shape[0], A.shape[0]
    Pp, Pi, Pv = S.csc_arrays_int64(P)
    Ap, Ai, Av = S.csc_arrays_int64(A)
    xs = []
    for base in (0, 1):
        h = C.c_void_p()
        s = S.make_settings(epsPcg=1e-10)
        _lib.check(lib.qpb200_create(C.byref(h), n, m, S._p64(Pp + base), S._p64(Pi + base), S._pd(Pv), S._p64(Ap + base),
                                     S._p64(Ai + base), S._pd(Av), S._pd(q), S._pd(l), S._pd(u), C.byref(s), base))
        x = np.zeros(n)
        info = _lib.Info()
        _lib.check(lib.qpb200_solve(h, S._pd(x), None, None, C.byref(info)))
        lib.qpb200_destroy(h)
        xs.append((x, info.conv_flag, info.iterations))
    assert np.array_equal(xs[0][0], xs[1][0]) and xs[0][1:] == xs[1][1:]


def test_update_settings_and_raw_array_call(lib):
    """qpb200_update_settings on a live handle (rho / adaptive rho change) equals a fresh handle; the raw CSC
    array call sequence of julia/QPB200.jl (solve_csc_arrays) equals the object wrapper."""
    S = _solver()
    P, q, A, l, u = config_cfg1(1234)
    n, m = P.shape[0], A.shape[0]
    kw2 = dict(rho=0.1, adptRho=True, epsAbs=1e-7, epsRel=1e-7, numIterations=50000, epsPcg=1e-11)
    with S.QPB200Solver(P, q, A, l, u, epsPcg=1e-10) as s:
        x1 = np.zeros(n); s.solve(x1)
        s.update_settings(**kw2)
        x2 = np.zeros(n); f2 = s.solve(x2); i2 = dict(s.info)
    x3 = np.zeros(n)
    f3, i3 = S.solve_csc_arrays(n, m, S.csc_arrays_int64(P), q, S.csc_arrays_int64(A), l, u, x3, **kw2)
    assert int(f2) == int(f3) and i2["iterations"] == i3["iterations"] and i2["rho_updates"] == i3["rho_updates"]
    assert np.array_equal(x2, x3) and not np.array_equal(x1, x2)


def test_dense_batch_is_bitwise_reproducible(lib):
    """Dynamic work queue, but every QP is solved by one CTA with fixed-order sums: results do not depend on
    which CTA picks which problem."""
    S = _solver()
    from workloads.problems import config_cfg3_batch
    P, q, A, l, u = config_cfg3_batch(700, 64, 96, seed=4)
    X1, f1, i1, _ = S.SolveQuadraticProgramBatch(P, q, A, l, u)
    X2, f2, i2, _ = S.SolveQuadraticProgramBatch(P, q, A, l, u)
    assert np.array_equal(X1, X2) and np.array_equal(f1, f2) and np.array_equal(i1, i2)


# ---------------------------------------------------------------------------------------------------
# Ruiz equilibration (numItrScaling).  Oracle = qp_oracle.solve(..., numItrScaling=k),
# which scales with the same formulas and tests convergence on the unscaled residuals.
# ---------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("case", ["cfg1", "cfg1_badly_scaled", "svm_inf_bounds", "sparse_5k"])
def test_equilibrated_solve_matches_oracle(lib, case):
    from workloads.problems import badly_scaled
    S = _solver()
    if case == "cfg1":
        prob = config_cfg1(seed=1236)
    elif case == "cfg1_badly_scaled":
        prob = badly_scaled(config_cfg1(seed=1234), seed=0)
    elif case == "svm_inf_bounds":
        prob = GenerateRandomQP(ProblemClass.supportVectorMachine, 10, seed=5)
    else:
        prob = badly_scaled(config_sparse(5000, 10000, 1e-3, seed=9), seed=1, var_decades=1.0, con_decades=2.0)
    P, q, A, l, u = prob
    # sparse_5k does not converge quickly: compare the iterates at a 200-iteration cap instead
    kw = dict(rho=0.1, adptRho=True, numIterations=200 if case == "sparse_5k" else 3000, epsPcg=1e-11, numItrScaling=10)
    x_ref, flag_ref, info_ref = qp_oracle.solve(P, q, A, l, u, mode="J", **kw)
    x = np.zeros(P.shape[0])
    with S.QPB200Solver(P, q, A, l, u, **kw) as s:
        flag = s.solve(x, want_zy=True)
        info = dict(s.info)
        with pytest.raises(S.QPB200Error):          # the handle holds the scaled operators
            s.apply(0, np.ones(P.shape[0]))
    _assert_parity(x, flag, info, x_ref, flag_ref, info_ref["iterations"])
    assert info["rho_updates"] == info_ref["rho_updates"]
    # z and y come back unscaled too
    sc = 1.0 + np.max(np.abs(info_ref["y"]))
    assert np.max(np.abs(info["y"] - info_ref["y"])) <= 1e-6 * sc
    assert np.max(np.abs(info["z"] - info_ref["z"])) <= 1e-6 * (1.0 + np.max(np.abs(info_ref["z"])))
    if int(flag) == 3:      # converged: the reported residuals are those of the unscaled problem
        assert abs(info["res_prim"] - np.max(np.abs(A @ x - info["z"]))) <= 1e-9 * (1 + np.max(np.abs(info["z"])))


def test_equilibration_rescues_a_badly_scaled_problem_and_update_vectors(lib):
    from workloads.problems import badly_scaled
    S = _solver()
    P, q, A, l, u = badly_scaled(config_cfg1(seed=1234), seed=0)
    kw = dict(rho=0.1, adptRho=True, numIterations=4000)
    x0, x1 = np.zeros(P.shape[0]), np.zeros(P.shape[0])
    with S.QPB200Solver(P, q, A, l, u, **kw) as s:
        f0 = s.solve(x0)
        it0 = s.info["iterations"]
    with S.QPB200Solver(P, q, A, l, u, numItrScaling=10, **kw) as s:
        f1 = s.solve(x1, want_zy=True)
        it1, y1 = s.info["iterations"], s.info["y"].copy()
        # parametric re-solve on the equilibrated handle: new q, l, u are scaled with the stored D, E, c
        q2 = 0.5 * q
        s.update_vectors(q2, l, u)
        x2 = np.zeros(P.shape[0])
        f2 = s.solve(x2, want_zy=True)
        y2 = s.info["y"].copy()
    assert int(f0) == 1 and it0 == 4000
    assert int(f1) == 3 and it1 <= 500
    assert max(qp_oracle.kkt_certificate(P, q, A, l, u, x1, y1).values()) < 1e-4
    assert int(f2) == 3
    assert max(qp_oracle.kkt_certificate(P, q2, A, l, u, x2, y2).values()) < 1e-4


def test_scaling_is_rejected_where_it_is_not_implemented(lib):
    S = _solver()
    rng = np.random.default_rng(0)
    n, m, b = 8, 12, 4
    M = rng.standard_normal((b, n, n))
    P = np.einsum("bij,bkj->bik", M, M) + np.eye(n)
    A_cm = rng.standard_normal((b, n, m))
    with pytest.raises(S.QPB200Error):
        S.QPB200Batch(P, rng.standard_normal((b, n)), A_cm, -np.ones((b, m)), np.ones((b, m)), numItrScaling=5)



# ---------------------------------------------------------------------------------------------------
# Per-constraint rho (second half): qpb200_set_rho_scale against the oracle's rhoScale
# ---------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("case", ["cfg1", "equality_n100", "svm_inf_bounds", "sparse_5k_scaled"])
def test_rho_scale_matches_oracle(lib, case):
    S = _solver()
    kw = dict(rho=0.1, adptRho=True, numIterations=3000, epsPcg=1e-10)
    if case == "cfg1":
        P, q, A, l, u = config_cfg1(seed=1235)
    elif case == "equality_n100":
        P, q, A, l, u = GenerateRandomQP(ProblemClass.equalityConstrainedQp, 100, numConstraints=50, seed=3)
        kw["adptRho"] = False                       # rho_i = 1e3 rho already; a larger rho only hurts the inner CG
    elif case == "svm_inf_bounds":
        P, q, A, l, u = GenerateRandomQP(ProblemClass.supportVectorMachine, 10, seed=5)
    else:                                            # together with the equilibration, iterates compared at a cap
        P, q, A, l, u = config_sparse(5000, 10000, 1e-3, seed=9)
        kw.update(numIterations=50, numItrScaling=10)
    # (the sparse case uses a milder factor: 1e3 costs ~500 CG iterations per ADMM iteration in the CPU oracle)
    rs = S.equality_rho_scale(l, u, 10.0 if case == "sparse_5k_scaled" else 1e3)
    rs[::3] *= 2.0
    x_ref, flag_ref, info_ref = qp_oracle.solve(P, q, A, l, u, mode="J", rhoScale=rs, **kw)
    x = np.zeros(P.shape[0])
    with S.QPB200Solver(P, q, A, l, u, rhoScale=rs, **kw) as s:
        flag = s.solve(x, want_zy=True)
        info = dict(s.info)
        # clearing the vector restores the scalar iteration bit for bit
        s.set_rho_scale(None)
        x_a = np.zeros(P.shape[0])
        flag_a = s.solve(x_a)
        it_a = s.info["iterations"]
    _assert_parity(x, flag, info, x_ref, flag_ref, info_ref["iterations"])
    assert info["rho_updates"] == info_ref["rho_updates"]
    assert np.max(np.abs(info["y"] - info_ref["y"])) <= 1e-6 * (1.0 + np.max(np.abs(info_ref["y"])))
    assert np.max(np.abs(info["z"] - info_ref["z"])) <= 1e-6 * (1.0 + np.max(np.abs(info_ref["z"])))
    x_b = np.zeros(P.shape[0])
    with S.QPB200Solver(P, q, A, l, u, **kw) as s:
        flag_b = s.solve(x_b)
        it_b = s.info["iterations"]
    assert int(flag_a) == int(flag_b) and it_a == it_b and np.array_equal(x_a, x_b)


def test_rho_eq_scale_keyword_and_argument_checks(lib):
    S = _solver()
    P, q, A, l, u = GenerateRandomQP(ProblemClass.equalityConstrainedQp, 100, numConstraints=50, seed=1)
    kw = dict(rho=0.1, numIterations=20000, epsPcg=1e-10)
    x0, x1 = np.zeros(100), np.zeros(100)
    with S.QPB200Solver(P, q, A, l, u, **kw) as s:
        s.solve(x0)
        it0 = s.info["iterations"]
        with pytest.raises(S.QPB200Error):
            s.set_rho_scale(np.zeros(50))            # factors must be positive
        with pytest.raises(S.QPB200Error):
            s.set_rho_scale(np.full(50, np.inf))
    with S.QPB200Solver(P, q, A, l, u, rhoEqScale=1e3, **kw) as s:
        f1 = s.solve(x1, want_zy=True)
        it1, y1 = s.info["iterations"], s.info["y"]
    assert int(f1) != 1 and it1 < it0
    assert max(qp_oracle.kkt_certificate(P, q, A, l, u, x1, y1).values()) < 1e-4


# ---------------------------------------------------------------------------------------------------
# RunTests.jl-shaped sweep (RunTests.jl:62-99): 9 problem classes x n in {10, 100} x 3 seeds with the settings of
# RunTests.jl:50-56 (rho = 0.1, adaptive rho, eps 1e-7, 50 000 iterations).  The reference compares with OSQP at
# 1e-5 (:58,93); here: the compiled oracle in the same linear-solver mode (J, tight inner solve -> same flag,
# iteration count within 2, x to 1e-6) AND the exact-solve oracle (mode D = the FacLdl plugin RunTests runs) at the
# reference's own 1e-5 wherever the (n+m) x (n+m) factorisation is cheap.
# ---------------------------------------------------------------------------------------------------
RUNTESTS_KW = dict(numIterations=50000, epsAbs=1e-7, epsRel=1e-7, rho=0.1, adptRho=True)


def _runtests_problem(pc, n, seed):
    m = (5 if n == 10 else 50) if pc == ProblemClass.equalityConstrainedQp else 0
    return GenerateRandomQP(pc, n, numConstraints=m, seed=seed)


@pytest.mark.parametrize("seed", [1234, 1235, 1236])
@pytest.mark.parametrize("n", [10, 100])
@pytest.mark.parametrize("pc", list(ProblemClass))
def test_runtests_sweep(lib, pc, n, seed):
    S = _solver()
    P, q, A, l, u = _runtests_problem(pc, n, seed)
    kw = dict(RUNTESTS_KW, epsPcg=1e-11)
    x, flag, info = S.SolveQuadraticProgram(P, q, A, l, u, **kw)
    xc, fc, ic = c_oracle.solve_sparse(P, q, A, l, u, precond=1, **kw)

    def gpu_at(k):
        xk, _, ik = S.SolveQuadraticProgram(P, q, A, l, u, **dict(kw, numIterations=k))
        return xk, ik

    def ref_at(k):
        xk, _, ik = c_oracle.solve_sparse(P, q, A, l, u, precond=1, **dict(kw, numIterations=k))
        return xk, ik

    # strict parity; where the exit checks differ (stop test fired on rounding noise) the same-trajectory criterion
    assert_parity(x, flag, info, xc, fc, ic, resolve_gpu=gpu_at, resolve_ref=ref_at, rho_updates=True,
                  what=f"{pc.name} n={n} seed={seed}:")
    if P.shape[0] + A.shape[0] <= 2500:
        xd, fd, idd = qp_oracle.solve(P, q, A, l, u, mode="D", **RUNTESTS_KW)
        assert int(fd) != 1 and int(flag) != 1
        assert np.max(np.abs(x - xd)) <= 1e-5, "RunTests.jl:58 absDevThr against the exact-solve plugin"


@pytest.mark.parametrize("precond,cmode", [("none", 0), ("jacobi", 1)])
def test_runtests_sweep_reference_default_inner_tolerance(lib, precond, cmode):
    """The same sweep at the reference's own inner tolerance (cg! abstol = 1e-6, LinearSystemSolvers.jl:125), where the
    exit check is sensitive to summation order: quantified rather than pinned.  Over 9 classes x n in {10, 100} the
    GPU and the compiled oracle must agree on the flag, differ by at most a few check intervals and agree on x to the
    reference's 1e-5 in nearly all cases; the spread is printed (scripts/gpu_default_eps_sweep.py commits it)."""
    S = _solver()
    diffs, bad_x, bad_flag = [], 0, 0
    for pc in ProblemClass:
        for n in (10, 100):
            P, q, A, l, u = _runtests_problem(pc, n, 1234)
            x, flag, info = S.SolveQuadraticProgram(P, q, A, l, u, precond=precond, **RUNTESTS_KW)
            xc, fc, ic = c_oracle.solve_sparse(P, q, A, l, u, precond=cmode, **RUNTESTS_KW)
            diffs.append(abs(int(info["iterations"]) - int(ic["iterations"])) // 25)
            bad_flag += int(int(flag) != int(fc))
            bad_x += int(np.max(np.abs(x - xc)) > 1e-5 * (1 + np.max(np.abs(xc))))
    print(f"default inner tolerance, precond={precond}: exit-check differences (intervals of 25) {sorted(diffs)}, "
          f"flag mismatches {bad_flag}, x beyond 1e-5: {bad_x} of {len(diffs)}")
    assert bad_flag <= 2 and bad_x <= 2
    assert sorted(diffs)[len(diffs) // 2] <= 1          # median: the same check or the neighbouring one


# ---------------------------------------------------------------------------------------------------
# The benchmarked configurations AT THEIR FULL SIZE against the compiled oracle, iteration count capped so that the
# CPU side finishes in seconds (the trajectories are compared mid-flight: same x after the same K ADMM iterations).
# ---------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("cfg,iters", [("cfg4", 200), ("cfg5", 15)])
def test_full_size_configs_against_c_oracle(lib, cfg, iters):
    S = _solver()
    c_oracle.use_all_cores()
    P, q, A, l, u = config_cfg4(seed=1234) if cfg == "cfg4" else config_cfg5(seed=1234)
    kw = dict(numIterations=iters, epsPcg=1e-10)
    x, flag, info = S.SolveQuadraticProgram(P, q, A, l, u, **kw)
    xc, fc, ic = c_oracle.solve_sparse(P, q, A, l, u, precond=1, **kw)
    _assert_parity(x, flag, info, xc, fc, ic["iterations"])
    assert abs(info["pcg_iters_total"] - ic["cg_iters_total"]) <= max(3, ic["cg_iters_total"] // 200)
    assert np.max(np.abs(info["z"] - ic["z"])) <= 1e-6 * (1 + np.max(np.abs(ic["z"])))
    assert np.max(np.abs(info["y"] - ic["y"])) <= 1e-5 * (1 + np.max(np.abs(ic["y"])))


@pytest.mark.parametrize("recur", ["standard", "one_reduction"])
def test_cg_recurrences_agree_with_oracle(lib, recur):
    """Both arrangements of the (P)CG recurrence (QPB200_RSV_CG_RECURRENCE) against the oracle on a mid-size problem."""
    S = _solver()
    P, q, A, l, u = config_sparse(10000, 20000, 1e-3, seed=1234)
    kw = dict(numIterations=400, epsPcg=1e-10)
    x, flag, info = S.SolveQuadraticProgram(P, q, A, l, u, cgRecurrence=recur, **kw)
    xc, fc, ic = c_oracle.solve_sparse(P, q, A, l, u, precond=1, **kw)
    _assert_parity(x, flag, info, xc, fc, ic["iterations"])
    assert abs(info["pcg_iters_total"] - ic["cg_iters_total"]) <= max(3, ic["cg_iters_total"] // 200)


def test_validation_of_updates_and_start_point(lib):
    """qpb200_update_vectors re-checks l <= u against the stored counterpart, qpb200_solve rejects a non-finite start
    point, qpb200_update_settings keeps the creation keywords that are not named and refuses the immutable ones."""
    S = _solver()
    P, q, A, l, u = config_cfg1(1234)
    with S.QPB200Solver(P, q, A, l, u, numIterations=50, adptRho=True, epsPcg=1e-10) as s:
        with pytest.raises(S.QPB200Error):
            s.update_vectors(vL=u + 1.0)                     # l > stored u
        with pytest.raises(S.QPB200Error):
            s.update_vectors(vU=l - 1.0)                     # u < stored l
        s.update_vectors(vL=l - 1.0, vU=u + 1.0)            # a consistent pair is accepted
        x = np.full(P.shape[0], np.nan)
        with pytest.raises(S.QPB200Error):
            s.solve(x)
        s.update_settings(rho=0.1)
        assert s.settings.max_iter == 50 and s.settings.adaptive_rho == 1 and s.settings.pcg_eps == 1e-10
        assert s.settings.rho == 0.1
        with pytest.raises(ValueError):
            s.update_settings(numItrScaling=5)
