"""CPU tests of the host side: the C-ABI library loads and exports every symbol of include/qpb200.h,
argument validation / error codes, the settings mirror, and the SpMV tile plan (emulated in numpy)."""
import ctypes as C
import os
import re

import numpy as np
import pytest
import scipy.sparse as sp

from quadraticprogramsolver_b200 import _lib
from workloads.problems import config_cfg1, sprandn

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_header_symbols_exported(lib):
    hdr = open(os.path.join(ROOT, "include", "qpb200.h")).read()
    declared = set(re.findall(r"\b(qpb200_[a-z0-9_]+)\s*\(", hdr))
    assert declared, "no declarations parsed"
    assert declared == set(_lib.EXPORTS)
    for sym in declared:
        assert hasattr(lib, sym), sym
    assert lib.qpb200_version() == 100


def test_struct_layout_matches_header(lib, tmp_path):
    src = tmp_path / "sz.c"
    src.write_text('#include "qpb200.h"\n#include <stdio.h>\n#include <stddef.h>\nint main(){printf("%zu %zu %zu %zu %zu %zu\\n", '
                   'sizeof(qpb200_settings), sizeof(qpb200_info), offsetof(qpb200_settings, pcg_eps), '
                   'offsetof(qpb200_info, solve_ms), sizeof(qpb200_proxqp_report), offsetof(qpb200_info, polish_active));return 0;}')
    exe = tmp_path / "sz"
    import subprocess
    subprocess.check_call(["/usr/bin/gcc", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe)])
    a, b, c, d, e, f = map(int, subprocess.check_output([str(exe)]).split())
    assert a == C.sizeof(_lib.Settings) and b == C.sizeof(_lib.Info)
    assert c == _lib.Settings.pcg_eps.offset and d == _lib.Info.solve_ms.offset
    assert e == C.sizeof(_lib.ProxReport) and f == _lib.Info.polish_active.offset


def test_default_settings_match_reference_kwargs(lib):
    s = _lib.default_settings()
    # SolveQuadraticProgram.jl:15-17 and LinearSystemSolvers.jl:125
    assert (s.max_iter, s.eps_abs, s.eps_rel, s.rho, s.sigma, s.alpha) == (5000, 1e-6, 1e-6, 1.0, 1e-6, 1.6)
    assert (s.adaptive_rho, s.rho_factor, s.check_every) == (0, 5.0, 25)
    assert (s.pcg_eps, s.pcg_max_iter) == (1e-6, 1000)
    assert (s.delta, s.polish_iter, s.minres_eps, s.minres_iter) == (1e-6, 10, 1e-6, 500)


def test_settings_mirror_accepts_reference_keyword_names(lib):
    from quadraticprogramsolver_b200.solver import make_settings
    s = make_settings(numIterations=50000, ϵAbs=1e-7, ϵRel=1e-7, ρ=0.1, adptΡ=True, σ=1e-5, α=1.5, fctrΡ=4, numItrConv=10)
    assert (s.max_iter, s.eps_abs, s.eps_rel, s.rho, s.adaptive_rho) == (50000, 1e-7, 1e-7, 0.1, 1)
    assert (s.sigma, s.alpha, s.rho_factor, s.check_every) == (1e-5, 1.5, 4.0, 10)
    with pytest.raises(TypeError):
        make_settings(notAKeyword=1)
    # new keyword: off by default = the reference's behaviour; header slot QPB200_RSV_SCALING_ITERS
    assert list(make_settings().reserved_i) == [0] * 7
    assert make_settings(numItrScaling=10).reserved_i[2] == 10
    hdr = open(os.path.join(os.path.dirname(__file__), "..", "include", "qpb200.h")).read()
    assert "#define QPB200_RSV_SCALING_ITERS 2" in hdr and "#define QPB200_RSV_DIST_MODE 1" in hdr


def _create(lib, P, q, A, l, u, base=0, settings=None):
    from quadraticprogramsolver_b200.solver import _csc_arrays, _p64, _pd
    Pp, Pi, Pv = _csc_arrays(P)
    Ap, Ai, Av = _csc_arrays(A)
    h = C.c_void_p()
    s = settings or _lib.default_settings()
    rc = lib.qpb200_create(C.byref(h), P.shape[0], A.shape[0], _p64(Pp + base), _p64(Pi + base), _pd(Pv),
                           _p64(Ap + base), _p64(Ai + base), _pd(Av), _pd(q), _pd(l), _pd(u), C.byref(s), base)
    return rc, h


def test_argument_validation_error_codes(lib):
    """Validation runs before the device is touched, so these codes are observable without a GPU."""
    P, q, A, l, u = config_cfg1()
    qbad = q.copy(); qbad[3] = np.nan
    rc, _ = _create(lib, P, qbad, A, l, u)
    assert rc == _lib.ERR_NONFINITE and b"q[3]" in lib.qpb200_last_error()
    lbad = l.copy(); lbad[0] = u[0] + 1.0
    rc, _ = _create(lib, P, q, A, lbad, u)
    assert rc == _lib.ERR_NONFINITE
    Pbad = P.copy(); Pbad.data[0] = np.inf
    rc, _ = _create(lib, Pbad, q, A, l, u)
    assert rc == _lib.ERR_NONFINITE
    rc, _ = _create(lib, P, q, A, l, u, base=2)
    assert rc == _lib.ERR_ARG
    s = _lib.default_settings(); s.rho = -1.0
    rc, _ = _create(lib, P, q, A, l, u, settings=s)
    assert rc in (_lib.ERR_ARG, _lib.ERR_DEVICE)   # settings are checked right after the device probe
    # good arguments: on a machine without a B200 the library must refuse loudly (no CPU fallback)
    rc, h = _create(lib, P, q, A, l, u, base=1)
    msg = lib.qpb200_last_error()
    if lib.qpb200_device_count() <= 0:
        assert rc == _lib.ERR_DEVICE and b"no CPU fallback" in msg
    else:
        assert rc == 0
        lib.qpb200_destroy(h)


def test_python_mirror_rejects_foreign_plugins_and_bad_dims(lib):
    from quadraticprogramsolver_b200.solver import QPB200Solver, SolveQuadraticProgram_
    P, q, A, l, u = config_cfg1()
    with pytest.raises(TypeError):
        SolveQuadraticProgram_(np.zeros(100), P, q, A, l, u, object(), object())
    with pytest.raises(ValueError):
        QPB200Solver(P, q[:-1], A, l, u)


# ---------------------------------------------------------------------------------------------------
# tile plan: emulate exactly what the kernel does with (tiles, cta_begin) and compare with scipy
# ---------------------------------------------------------------------------------------------------
def _plan(lib, M, grid):
    M = sp.csr_matrix(M)
    ptr = np.ascontiguousarray(M.indptr, dtype=np.int32)
    cap = M.shape[0] + M.nnz // 64 + 16
    tiles = np.zeros((cap, 4), dtype=np.int32)
    cta = np.zeros(grid + 1, dtype=np.int32)
    lpr = C.c_int32(0)
    p32 = C.POINTER(C.c_int32)
    nt = lib.qpb200_debug_tile_plan(M.shape[0], ptr.ctypes.data_as(p32), grid, tiles.ctypes.data_as(p32), cap,
                                    cta.ctypes.data_as(p32), C.byref(lpr))
    assert 0 <= nt <= cap
    return tiles[:nt], cta, lpr.value


def _emulate(M, x, tiles, cta, T):
    M = sp.csr_matrix(M)
    y = np.full(M.shape[0], np.nan)
    written = np.zeros(M.shape[0], dtype=int)
    assert cta[0] == 0 and cta[-1] == len(tiles) and np.all(np.diff(cta) >= 0)
    next_row, next_k = 0, 0
    for b in range(len(cta) - 1):
        carry = None
        for t in range(cta[b], cta[b + 1]):
            row0, nrows, k0, w = (int(v) for v in tiles[t])
            nk = w & ((1 << 24) - 1)
            from_prev, to_next = bool(w & (1 << 30)), bool(w & (1 << 29))
            assert nk <= T and k0 == next_k
            next_k = k0 + nk
            prod = M.data[k0:k0 + nk] * x[M.indices[k0:k0 + nk]]
            if from_prev or to_next:
                assert nrows == 1
                assert from_prev == (carry is not None), "a long row must stay on one CTA"
                s = (carry or 0.0) + prod.sum()
                if to_next:
                    carry = s
                else:
                    y[row0] = s; written[row0] += 1; carry = None
                    assert next_k == M.indptr[row0 + 1]
                    next_row = row0 + 1
            else:
                assert carry is None and row0 == next_row and k0 == M.indptr[row0]
                for r in range(row0, row0 + nrows):
                    a, e = M.indptr[r] - k0, M.indptr[r + 1] - k0
                    assert 0 <= a <= e <= nk
                    y[r] = prod[a:e].sum(); written[r] += 1
                next_row = row0 + nrows
        assert carry is None
    assert next_row == M.shape[0] and next_k == M.nnz
    assert np.all(written == 1)
    return y


@pytest.mark.parametrize("case", ["random", "long_rows", "empty_rows", "dense", "tiny", "all_empty"])
@pytest.mark.parametrize("grid", [1, 7, 296])
def test_tile_plan_covers_matrix_exactly_once(lib, case, grid):
    rng = np.random.default_rng(5)
    T = lib.qpb200_debug_tile_nnz()
    if case == "random":
        M = sprandn(rng, 3000, 2000, 0.01).tocsr()
    elif case == "long_rows":     # rows longer than a tile, incl. exact multiples of the tile size
        M = sprandn(rng, 40, 3 * T + 5, 0.02).tolil()
        M[3, :] = rng.standard_normal(3 * T + 5)
        M[4, :2 * T] = rng.standard_normal(2 * T)
        M[17, :T + 1] = 1.0
        M[39, :] = 2.0
        M = M.tocsr()
    elif case == "empty_rows":
        M = sprandn(rng, 20000, 50, 0.002).tocsr()
    elif case == "dense":
        M = sp.csr_matrix(rng.standard_normal((300, 300)))
    elif case == "tiny":
        M = sp.csr_matrix(np.array([[1.0, 0.0], [0.0, 0.0], [2.0, 3.0]]))
    else:
        M = sp.csr_matrix((5000, 10))
    x = rng.standard_normal(M.shape[1])
    tiles, cta, lpr = _plan(lib, M, grid)
    assert lpr in (1, 2, 4, 8, 16, 32)
    y = _emulate(M, x, tiles, cta, T)
    ref = M @ x
    assert np.allclose(y, ref, rtol=1e-12, atol=1e-12)


def test_tile_plan_is_balanced(lib):
    rng = np.random.default_rng(6)
    M = sprandn(rng, 200000, 100000, 5e-5).tocsr()
    tiles, cta, _ = _plan(lib, M, 296)
    nk = tiles[:, 3] & ((1 << 24) - 1)
    per_cta = np.array([nk[cta[b]:cta[b + 1]].sum() for b in range(296)])
    assert per_cta.sum() == M.nnz
    assert per_cta.max() <= 1.25 * per_cta.mean() + lib.qpb200_debug_tile_nnz()


@pytest.mark.parametrize("case", ["cfg1", "cfg1_badly_scaled", "svm_inf_bounds", "m0"])
def test_host_equilibration_matches_oracle(lib, case):
    """qpb200_create's host-side Ruiz equilibration (no GPU needed) against oracle/qp_oracle.ruiz_equilibrate."""
    import scipy.sparse as sp
    from oracle import qp_oracle
    from workloads.problems import GenerateRandomQP, ProblemClass, badly_scaled, config_cfg1
    from quadraticprogramsolver_b200.solver import _csc_arrays, _p64, _pd
    if case == "cfg1":
        P, q, A, l, u = config_cfg1(seed=1234)
    elif case == "cfg1_badly_scaled":
        P, q, A, l, u = badly_scaled(config_cfg1(seed=1234), seed=0)
    elif case == "svm_inf_bounds":
        P, q, A, l, u = GenerateRandomQP(ProblemClass.supportVectorMachine, 10, seed=5)
    else:
        P, q, A, l, u = config_cfg1(seed=1234)
        A, l, u = sp.csc_matrix((0, P.shape[0])), np.zeros(0), np.zeros(0)
    n, m = P.shape[0], A.shape[0]
    Pa, Aa = _csc_arrays(P), _csc_arrays(A)
    D, E, c, qs = np.zeros(n), np.zeros(max(m, 1)), C.c_double(0.0), np.zeros(n)
    Pv, Av = np.zeros(len(Pa[2])), np.zeros(max(len(Aa[2]), 1))
    q = np.ascontiguousarray(q, dtype=np.float64)
    rc = lib.qpb200_debug_equilibrate(n, m, _p64(Pa[0]), _p64(Pa[1]), _pd(Pa[2]), _p64(Aa[0]), _p64(Aa[1]), _pd(Aa[2]), _pd(q),
                                      10, 0, _pd(D), _pd(E), C.byref(c), _pd(qs), _pd(Pv), _pd(Av))
    assert rc == 0, lib.qpb200_last_error()
    Ps, qs_ref, As, ls, us, D_ref, E_ref, c_ref = qp_oracle.ruiz_equilibrate(P, q, A, l, u, 10)
    # same formulas, same operation order: agreement to the last couple of bits
    np.testing.assert_allclose(D, D_ref, rtol=1e-14)
    np.testing.assert_allclose(E[:m], E_ref, rtol=1e-14)
    assert abs(c.value - c_ref) <= 1e-14 * c_ref
    np.testing.assert_allclose(qs, qs_ref, rtol=1e-13, atol=1e-300)
    Ps, As = sp.csc_matrix(Ps), sp.csc_matrix(As)
    Ps.sort_indices(); As.sort_indices()
    np.testing.assert_allclose(Pv, Ps.data, rtol=1e-13)
    np.testing.assert_allclose(Av[:As.nnz], As.data, rtol=1e-13)


def _assemble_and_compare(lib, P, A, base):
    """qpb200_debug_assemble_h (the host conversion of qpb200_create) against scipy: same row pointers, split points,
    columns in ascending order, bit-identical values; plus diag(P) and the column square sums of A."""
    from quadraticprogramsolver_b200.solver import _p64, _pd
    n, m = P.shape[0], A.shape[0]
    P, A = sp.csc_matrix(P), sp.csc_matrix(A)
    for X in (P, A):
        X.sort_indices()
    arrs = []
    for X in (P, A):
        arrs += [np.ascontiguousarray(X.indptr, dtype=np.int64) + base, np.ascontiguousarray(X.indices, dtype=np.int64) + base,
                 np.ascontiguousarray(X.data, dtype=np.float64)]
    nnz = P.nnz + A.nnz
    ptr, mid = np.zeros(n + 1, np.int32), np.zeros(n, np.int32)
    col, val = np.zeros(max(nnz, 1), np.int32), np.zeros(max(nnz, 1))
    dP, dAA = np.zeros(n), np.zeros(n)
    p32 = lambda a: a.ctypes.data_as(C.POINTER(C.c_int32))
    rc = lib.qpb200_debug_assemble_h(n, m, _p64(arrs[0]), _p64(arrs[1]), _pd(arrs[2]), _p64(arrs[3]), _p64(arrs[4]), _pd(arrs[5]),
                                     base, p32(ptr), p32(mid), p32(col), _pd(val), _pd(dP), _pd(dAA))
    assert rc == 0, lib.qpb200_last_error()
    H = sp.hstack([sp.csr_matrix(P), sp.csr_matrix(A.T)], format="csr")
    H.sort_indices()
    assert np.array_equal(ptr, H.indptr)
    assert np.array_equal(col[:nnz], H.indices)
    assert np.array_equal(val[:nnz], H.data)
    assert np.array_equal(mid, H.indptr[:-1] + np.diff(sp.csr_matrix(P).indptr))
    np.testing.assert_array_equal(dP, P.diagonal())
    np.testing.assert_allclose(dAA, np.asarray(A.multiply(A).sum(axis=0)).ravel(), rtol=1e-14, atol=0)


@pytest.mark.parametrize("n,m,dens,base", [(60, 40, 0.2, 0), (60, 0, 0.2, 1), (3000, 5000, 0.02, 1), (40000, 70000, 3e-4, 0)])
def test_host_operator_assembly_matches_scipy(lib, n, m, dens, base):
    """CSC inputs -> H = [P A'] row-major: bucketed parallel transpose for the large case, serial counting sort for the
    small ones."""
    rng = np.random.default_rng(n + m)
    M = sprandn(rng, n, n, dens)
    P = sp.csc_matrix(M.T @ M + 0.01 * sp.identity(n))          # symmetric
    P = sp.csc_matrix(P + sp.triu(sprandn(rng, n, n, dens / 4), 1))   # ... and a non-symmetric part: rows != columns
    A = sp.csc_matrix(sprandn(rng, m, n, dens)) if m else sp.csc_matrix((0, n))
    _assemble_and_compare(lib, P, A, base)


@pytest.mark.parametrize("case", ["arrowhead", "top_rows_only", "dense_512", "dense_600", "last_column_only", "one_row_of_A"])
def test_host_operator_assembly_skewed_patterns(lib, case):
    """The bucketed transpose (nnz >= 2^18) on patterns that put everything into one row block, one column block or one
    thread's share: a dense row and column, all entries in the first rows, dense matrices at the row-block boundaries
    (512 rows = one row per block, 600 rows = two), a single non-empty column, and a constraint matrix of one dense row."""
    rng = np.random.default_rng(17)
    A = None
    if case == "arrowhead":
        n = 150000
        r = np.arange(1, n)
        P = sp.coo_matrix((rng.standard_normal(3 * n - 2), (np.concatenate([np.arange(n), np.zeros(n - 1, int), r]),
                                                           np.concatenate([np.arange(n), r, np.zeros(n - 1, int)]))), shape=(n, n))
    elif case == "top_rows_only":
        n = 5000
        P = sp.vstack([sp.csr_matrix(rng.standard_normal((100, n))), sp.csr_matrix((n - 100, n))])
    elif case in ("dense_512", "dense_600"):
        n = int(case.split("_")[1])
        P = sp.csr_matrix(rng.standard_normal((n, n)))
    elif case == "last_column_only":
        n = 300000
        P = sp.coo_matrix((rng.standard_normal(n), (np.arange(n), np.full(n, n - 1))), shape=(n, n))
    else:
        n = 300000
        P = sp.identity(n, format="csc") * 2.0
        A = sp.csr_matrix(rng.standard_normal((1, n)))
    if A is None:
        A = sprandn(rng, 7, n, 0.3)
    _assemble_and_compare(lib, P, A, 1 if case in ("arrowhead", "dense_600") else 0)


def test_bench_reference_arm_prints_exactly_one_json_line():
    """The driver's contract: `bench.py --impl reference` (the CPU port of the path on the host cores; Julia is not
    available) prints ONE JSON line on stdout with the agreed keys -- nothing else, whatever native code prints."""
    import json
    import subprocess
    import sys
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--scale", "0.005",
                          "--steps", "1", "--warmup", "0"], capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [ln for ln in out.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1, out.stdout[-2000:]
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "admm_iters_per_s" and d["unit"] == "iter/s"
    assert d["higher_is_better"] is True and d["dtype"] == "f64" and d["n_gpus"] == 1
    assert d["value"] > 0 and d["e2e"]["value"] == d["value"]
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["sample"]
    assert "workload" in d["config"]


# ---------------------------------------------------------------------------------------------------
# The partition behind qpb200_dist_create_full (dist_partition.cpp) against partition.py, without a device
# ---------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("nranks", [1, 2, 3, 8])
@pytest.mark.parametrize("base", [0, 1])
def test_in_library_partition_and_slices_match_partition_py(lib, nranks, base):
    from quadraticprogramsolver_b200 import partition
    from quadraticprogramsolver_b200.solver import _csc_arrays
    from workloads.problems import config_sparse
    P, q, A, l, u = config_sparse(900, 1700, 6e-3, seed=21)
    A = sp.csc_matrix(sp.vstack([A[:400], sp.csr_matrix((37, 900)), A[400:]]))      # a run of empty rows
    n, m = P.shape[0], A.shape[0]
    (Pp, Pi, Pv), (Ap, Ai, Av) = _csc_arrays(P), _csc_arrays(A)
    Pp, Pi, Ap, Ai = Pp + base, Pi + base, Ap + base, Ai + base
    p64 = lambda a: a.ctypes.data_as(C.POINTER(C.c_int64))
    pd = lambda a: a.ctypes.data_as(C.POINTER(C.c_double))
    rb = np.zeros(nranks + 1, dtype=np.int64); cb = np.zeros(nranks + 1, dtype=np.int64)
    assert lib.qpb200_debug_partition(n, m, p64(Pp), p64(Ap), p64(Ai), base, nranks, p64(rb), p64(cb)) == nranks
    rb_py, cb_py = partition.plan(P, A, nranks)
    assert np.array_equal(rb, rb_py) and np.array_equal(cb, cb_py)
    assert rb[0] == 0 and rb[-1] == m and cb[0] == 0 and cb[-1] == n
    for rank in range(nranks):
        b4 = np.zeros(4, dtype=np.int64); poff = C.c_int64()
        Pc = np.zeros(n + 1, dtype=np.int64); Ac = np.zeros(n + 1, dtype=np.int64)
        Ar = np.zeros(A.nnz + 1, dtype=np.int64); Avv = np.zeros(A.nnz + 1)
        assert lib.qpb200_debug_slice(n, m, p64(Pp), p64(Ap), p64(Ai), pd(Av), base, rank, nranks, p64(b4), C.byref(poff),
                                      p64(Pc), p64(Ac), p64(Ar), pd(Avv), A.nnz + 1) == 0
        P_r, A_r, l_r, u_r, (i0, i1), (j0, j1) = partition.slice_problem(P, A, l, u, rank, nranks)
        assert b4.tolist() == [i0, i1, j0, j1]
        nnzp = Pc[-1] - base
        mine = sp.csc_matrix((Pv[poff.value:poff.value + nnzp], Pi[poff.value:poff.value + nnzp] - base, Pc - base), shape=(n, n))
        assert (mine != P_r).nnz == 0
        nnza = Ac[-1] - base
        mine = sp.csc_matrix((Avv[:nnza], Ar[:nnza] - base, Ac - base), shape=(i1 - i0, n))
        assert (mine != A_r).nnz == 0 and mine.nnz == A_r.nnz


def test_julia_shim_binds_exported_symbols_and_mirrors_the_struct_layouts(lib):
    """Julia is absent here, so the shim cannot run; what can be checked statically is: every symbol it ccalls is exported,
    and its three mutable structs list the same fields in the same order as the ctypes mirrors (whose sizes and offsets
    test_struct_layout_matches_header checks against the header)."""
    src = open(os.path.join(ROOT, "julia", "QPB200.jl")).read()
    syms = set(re.findall(r"ccall\(\(:(qpb200_[a-z0-9_]+)", src))
    assert syms and syms <= set(_lib.EXPORTS)
    for fam in ("qpb200_create", "qpb200_batch_solve_once", "qpb200_batch_create_shared", "qpb200_dist_create_full",
                "qpb200_proxqp_solve", "qpb200_set_rho_scale"):
        assert fam in syms, fam

    def fields(name):
        body = re.search(r"mutable struct %s\n(.*?)\n    %s\(\) = new\(\)" % (name, name), src, re.S).group(1)
        return [ln.split("::")[0].strip() for ln in body.splitlines() if "::" in ln]

    assert fields("Settings") == [f for f, _ in _lib.Settings._fields_]
    assert fields("Info") == [f for f, _ in _lib.Info._fields_]
    assert fields("ProxReport") == [f for f, _ in _lib.ProxReport._fields_]


def test_tile_plan_property_random_row_lengths(lib):
    """Property test of the tile plan (the static work distribution every matrix pass of the kernels follows): for
    random row-length profiles -- empty rows, rows around the tile size and its multiples, a few very long rows -- and
    random grid sizes, every non-zero is covered exactly once, in order, rows longer than a tile stay on one CTA, and
    the emulated pass reproduces M x."""
    from hypothesis import given, settings, strategies as st

    T = lib.qpb200_debug_tile_nnz()
    length = st.one_of(st.just(0), st.integers(0, 12), st.integers(T - 3, T + 3), st.sampled_from([2 * T, 3 * T, 2 * T + 1]),
                       st.integers(0, 5 * T))
    profile = st.lists(length, min_size=1, max_size=120)

    @settings(max_examples=120, deadline=None, derandomize=True)
    @given(profile, st.integers(1, 600), st.integers(0, 3))
    def run(lengths, grid, repeat):
        lengths = np.array(lengths * (1 + repeat), dtype=np.int64)
        ptr = np.concatenate([[0], np.cumsum(lengths)])
        nnz = int(ptr[-1])
        ncols = max(1, int(lengths.max()))
        idx = np.concatenate([np.arange(k) for k in lengths]) if nnz else np.zeros(0, dtype=np.int64)
        rng = np.random.default_rng(nnz + grid)
        M = sp.csr_matrix((rng.standard_normal(nnz), idx.astype(np.int32), ptr.astype(np.int32)), shape=(len(lengths), ncols))
        x = rng.standard_normal(ncols)
        tiles, cta, lpr = _plan(lib, M, grid)
        assert lpr in (1, 2, 4, 8, 16, 32)
        y = _emulate(M, x, tiles, cta, T)
        assert np.allclose(y, M @ x, rtol=1e-12, atol=1e-12)

    run()
