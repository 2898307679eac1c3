"""GPU parity tests of the batched dense path (configs[2]) against the oracle's exact-solve mode D --
the configuration the reference's own tests run (FacLdl, RunTests.jl:55-56)."""
import numpy as np
import pytest
import scipy.sparse as sp

from oracle import c_oracle, qp_oracle
from workloads.problems import config_cfg3_batch

pytestmark = pytest.mark.gpu


def _S():
    from quadraticprogramsolver_b200 import solver
    return solver


def _check(X, flags, iters, Xr, fr, ir, tol=1e-6):
    assert np.array_equal(flags, fr), f"flags differ at {np.nonzero(flags != fr)[0][:8]}"
    assert np.max(np.abs(iters - ir)) <= 2, f"iterations differ: {np.max(np.abs(iters - ir))}"
    err = np.max(np.abs(X - Xr), axis=1)
    ref = 1.0 + np.max(np.abs(Xr), axis=1)
    # a problem that ran into the iteration cap (flag 1) has no limit point to agree on: after tens of thousands of
    # steps of a non-convergent sequence two correct implementations differ by more than the converged tolerance
    tol_b = np.where(fr == 1, max(tol, 1e-4), tol)
    assert np.all(err <= tol_b * ref), f"max |x - x_ref| / (1+|x_ref|) = {np.max(err / ref):.3e}"


@pytest.mark.parametrize("unblocked", [False, True])
@pytest.mark.parametrize("kw", [dict(), dict(rho=0.1, adptRho=True, epsAbs=1e-7, epsRel=1e-7, numIterations=50000)],
                         ids=["defaults", "runtests_adaptive_rho"])
def test_cfg3_shape_matches_oracle(lib, kw, unblocked):
    """n = 64, m = 96: the configs[2] shape; both Cholesky variants (DMMA-blocked / unblocked)."""
    P, q, A, l, u = config_cfg3_batch(96, 64, 96, seed=1234)
    X, flags, iters, info = _S().SolveQuadraticProgramBatch(P, q, A, l, u, unblockedCholesky=unblocked, **kw)
    Xr, fr, ir, _, rc = c_oracle.solve_dense_batch(P, q, A, l, u, **kw)
    assert rc == 0
    _check(X, flags, iters, Xr, fr, ir)
    assert info["iterations"] == int(iters.sum())


@pytest.mark.parametrize("variant", ["regs", "regs_ak"])
@pytest.mark.parametrize("n,m", [(64, 96), (64, 93), (40, 96), (17, 94)])
@pytest.mark.parametrize("kw", [dict(), dict(rho=0.1, adptRho=True, epsAbs=1e-7, epsRel=1e-7, numIterations=50000)],
                         ids=["defaults", "runtests_adaptive_rho"])
def test_register_resident_variant_matches_oracle_and_smem_variant(lib, n, m, kw, variant):
    """The kernel variants that keep A (and K^-1) in registers during the iterations (shapes padded to 64 x 96): same
    flags and iteration counts as the oracle and as the shared-memory variant, x to 1e-6 / 1e-8."""
    P, q, A, l, u = config_cfg3_batch(160, n, m, seed=77)
    X0 = np.random.default_rng(3).standard_normal((160, n))
    Xa, fa, ia, _ = _S().SolveQuadraticProgramBatch(P, q, A, l, u, X0=X0, denseVariant=variant, **kw)
    Xb, fb, ib, _ = _S().SolveQuadraticProgramBatch(P, q, A, l, u, X0=X0, denseVariant="smem", **kw)
    Xr, fr, ir, _, rc = c_oracle.solve_dense_batch(P, q, A, l, u, x0=X0, **kw)
    assert rc == 0
    _check(Xa, fa, ia, Xr, fr, ir)
    _check(Xa, fa, ia, Xb, fb, ib, tol=1e-8)
    Xc, fc, ic, _ = _S().SolveQuadraticProgramBatch(P, q, A, l, u, X0=X0, denseVariant=variant, **kw)
    assert np.array_equal(Xa, Xc) and np.array_equal(ia, ic)          # bitwise reproducible


def test_register_resident_variant_needs_the_cfg3_row_padding(lib):
    from quadraticprogramsolver_b200 import _lib
    P, q, A, l, u = config_cfg3_batch(3, 8, 4, seed=1)
    with pytest.raises(_lib.QPB200Error) as e:
        _S().SolveQuadraticProgramBatch(P, q, A, l, u, denseVariant="regs")
    assert e.value.code == _lib.ERR_ARG


@pytest.mark.parametrize("n,m", [(8, 4), (30, 45), (64, 128), (63, 97), (1, 1)])
def test_padded_shapes(lib, n, m):
    """n < 64 and m not a multiple of 4 are zero padded inside the kernel."""
    P, q, A, l, u = config_cfg3_batch(20, n, m, seed=5)
    X, flags, iters, info = _S().SolveQuadraticProgramBatch(P, q, A, l, u)
    Xr, fr, ir, _, rc = c_oracle.solve_dense_batch(P, q, A, l, u)
    _check(X, flags, iters, Xr, fr, ir)


def test_against_python_direct_plugin(lib):
    """Against qp_oracle.py mode D = the KKT-matrix LDL' solve of LinearSystemSolvers.jl:16-107."""
    P, q, A, l, u = config_cfg3_batch(4, 64, 96, seed=9)
    X, flags, iters, _ = _S().SolveQuadraticProgramBatch(P, q, A, l, u)
    for b in range(4):
        x, f, i = qp_oracle.solve(sp.csc_matrix(P[b]), q[b], sp.csc_matrix(A[b].T), l[b], u[b], mode="D")
        assert int(f) == flags[b] and abs(i["iterations"] - iters[b]) <= 2
        assert np.max(np.abs(x - X[b])) <= 1e-6 * (1 + np.max(np.abs(x)))


def test_start_points_and_more_problems_than_ctas(lib):
    """More QPs than resident CTAs (persistent loop over the batch) and non-zero start points."""
    P, q, A, l, u = config_cfg3_batch(1500, 16, 24, seed=2)
    X0 = np.random.default_rng(0).standard_normal((1500, 16))
    X, flags, iters, _ = _S().SolveQuadraticProgramBatch(P, q, A, l, u, X0=X0)
    Xr, fr, ir, _, rc = c_oracle.solve_dense_batch(P, q, A, l, u, x0=X0)
    _check(X, flags, iters, Xr, fr, ir)


def test_breakdown_is_reported(lib):
    """An indefinite P makes a pivot non-positive: QPB200_ERR_FACTOR, not a silent wrong answer."""
    from quadraticprogramsolver_b200 import _lib
    P, q, A, l, u = config_cfg3_batch(3, 8, 4, seed=1)
    P[1] = -10.0 * np.eye(8)
    with pytest.raises(_lib.QPB200Error) as e:
        _S().SolveQuadraticProgramBatch(P, q, A, l, u, sigma=1e-6)
    assert e.value.code == _lib.ERR_FACTOR


def test_batch_argument_validation(lib):
    from quadraticprogramsolver_b200 import _lib
    P, q, A, l, u = config_cfg3_batch(3, 8, 4, seed=1)
    q2 = q.copy(); q2[2, 1] = np.inf
    with pytest.raises(_lib.QPB200Error) as e:
        _S().SolveQuadraticProgramBatch(P, q2, A, l, u)
    assert e.value.code == _lib.ERR_NONFINITE
    with pytest.raises(ValueError):
        _S().QPB200Batch(P, q[:, :-1], A, l, u)


def test_batch_update_vectors_resolves_without_reupload(lib):
    """MPC-style: same P, A on the device, new q / l / u per solve -- equals a fresh handle bit for bit."""
    from quadraticprogramsolver_b200 import _lib
    S = _S()
    P, q, A, l, u = config_cfg3_batch(64, 64, 96, seed=4)
    rng = np.random.default_rng(1)
    q2 = q + 0.1 * rng.standard_normal(q.shape)
    l2, u2 = l - 0.05, u + 0.05
    with S.QPB200Batch(P, q, A, l, u) as b:
        b.solve()
        b.update_vectors(q2, l2, u2)
        X1, f1, i1 = b.solve()
        with pytest.raises(_lib.QPB200Error):
            b.update_vectors(l=u2 + 1.0, u=u2)           # l > u
        b.update_vectors(q=q)                             # only q back: bounds stay
        X3, f3, i3 = b.solve()
    X2, f2, i2, _ = S.SolveQuadraticProgramBatch(P, q2, A, l2, u2)
    assert np.array_equal(X1, X2) and np.array_equal(f1, f2) and np.array_equal(i1, i2)
    X4, f4, i4, _ = S.SolveQuadraticProgramBatch(P, q, A, l2, u2)
    assert np.array_equal(X3, X4) and np.array_equal(i3, i4)


def test_pipelined_one_shot_solve_equals_resident_solve_bit_for_bit(lib):
    """qpb200_batch_solve_once (chunked upload overlapped with the solve; 5 chunks, the last one ragged) against
    create + solve with everything resident."""
    P, q, A, l, u = config_cfg3_batch(1100, 64, 96, seed=77)
    rng = np.random.default_rng(5)
    X0 = rng.standard_normal((1100, 64))
    Xa, fa, ia, infa = _S().SolveQuadraticProgramBatch(P, q, A, l, u, X0=X0, batchChunk=256)
    Xb, fb, ib, infb = _S().SolveQuadraticProgramBatch(P, q, A, l, u, X0=X0, pipelined=False)
    assert np.array_equal(Xa, Xb) and np.array_equal(fa, fb) and np.array_equal(ia, ib)
    assert infa["kernel_launches"] == 5 and infa["iterations"] == infb["iterations"] == int(ia.sum())
    Xc, fc, ic, _ = _S().SolveQuadraticProgramBatch(P[:100], q[:100], A[:100], l[:100], u[:100], X0=X0[:100])   # one chunk
    assert np.array_equal(Xc, Xb[:100]) and np.array_equal(ic, ib[:100])


# ---------------------------------------------------------------------------------------------------
# Shared-(P, A) batch (MPC-style): one factor, 16 problems per CTA as GEMM columns
# ---------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("shape", [(64, 96), (64, 50), (40, 96), (17, 23)], ids=lambda s: f"n{s[0]}_m{s[1]}")
def test_shared_matrix_batch_matches_oracle_and_per_problem_kernel(lib, shape):
    from workloads.problems import config_cfg3_shared
    n, m = shape
    batch = 150                                   # 10 tiles of 16 columns with a ragged tail, slots refilled as columns finish
    P, q, A, l, u = config_cfg3_shared(batch, n, m, seed=11)
    rng = np.random.default_rng(3)
    X0 = rng.standard_normal((batch, n)) * 0.1
    S = _S()
    with S.QPB200Batch(P, q, A, l, u) as b:
        assert b.shared
        X, flags, iters = b.solve(X0.copy())
        info = dict(b.info)
        X2, f2, i2 = b.solve(X0.copy())
        assert np.array_equal(X, X2) and np.array_equal(iters, i2)            # bit-reproducible
    Pb = np.ascontiguousarray(np.broadcast_to(P, (batch, n, n)))
    Ab = np.ascontiguousarray(np.broadcast_to(A, (batch, n, m)))
    Xr, fr, ir, _, rc = c_oracle.solve_dense_batch(Pb, q, Ab, l, u, x0=X0)     # oracle mode D, problem by problem
    assert rc == 0
    _check(X, flags, iters, Xr, fr, ir)
    assert info["iterations"] == int(iters.sum())
    Xp, fp, ip, _ = S.SolveQuadraticProgramBatch(Pb, q, Ab, l, u, X0=X0)       # the per-problem-factor kernel
    assert np.array_equal(flags, fp) and np.max(np.abs(iters - ip)) <= 2
    assert np.max(np.abs(X - Xp)) <= 1e-8 * (1 + np.max(np.abs(Xp)))


def test_shared_matrix_batch_iteration_cap_resolve_and_errors(lib):
    from workloads.problems import config_cfg3_shared
    S = _S()
    P, q, A, l, u = config_cfg3_shared(40, 64, 96, seed=5)
    Pb = np.ascontiguousarray(np.broadcast_to(P, (40, 64, 64)))
    Ab = np.ascontiguousarray(np.broadcast_to(A, (40, 64, 96)))
    for cap in (60, 75):                          # a cap between two check points / on a check point
        X, flags, iters, _ = S.SolveQuadraticProgramBatch(P, q, A, l, u, numIterations=cap)
        Xr, fr, ir, _, _ = c_oracle.solve_dense_batch(Pb, q, Ab, l, u, numIterations=cap)
        assert np.array_equal(flags, fr) and np.array_equal(iters, ir) and np.all(iters <= cap)
        assert np.max(np.abs(X - Xr)) <= 1e-9 * (1 + np.max(np.abs(Xr)))
    with S.QPB200Batch(P, q, A, l, u) as b:       # MPC-style re-solve: new vectors, same factor
        b.solve()
        q2 = q[::-1].copy()
        b.update_vectors(q=q2)
        X, flags, iters = b.solve()
    Xr, fr, ir, _, _ = c_oracle.solve_dense_batch(Pb, q2, Ab, l, u)
    _check(X, flags, iters, Xr, fr, ir)
    with pytest.raises(S.QPB200Error):
        S.QPB200Batch(P, q, A, l, u, adptRho=True)                             # one factor needs one rho
    with pytest.raises(S.QPB200Error):
        S.QPB200Batch(-P, q, A * 0.0, l, u)                                    # K not positive definite
